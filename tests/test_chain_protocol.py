"""Model check of the launch chain of nlbm_dense_step_n (k_dense_chain, neon_b200/csrc/lbm_step.cuh) — no GPU.

The chain runs iteration t+1 while iteration t drains: a launch becomes eligible once EVERY block of its predecessor has started
(programmatic dependent launch); the blocks of the first `early` planes start when the per-plane counters say that planes z-1, z,
z+1 of the previous iteration are complete, the others when the previous launch is over (griddepcontrol.wait); blocks of planes
0..early publish their completion.  Iteration t reads field t % 2 (planes z-1..z+1 of a tile of plane z) and writes plane z of
the other field.

This test simulates that protocol as a discrete-event system under a RANDOM scheduler — any eligible block may take any free slot
of the chip, in any order, with random durations — and checks, for several shapes, `early` values, slot counts and seeds, that
  * no block reads a plane before all its writers of the previous iteration have finished (RAW),
  * no block overwrites a plane while a reader of the previous iteration is still running or yet to run (WAR),
  * the chain always terminates (no deadlock with blocks that spin while they hold a slot),
and that deliberately broken variants are caught: polling two planes instead of three, late blocks that do not wait, publishing
planes 0..early-1 only, and a dependent launch released before all blocks of its predecessor have started.
"""
import random

import pytest


class Violation(Exception):
    pass


def simulate(nz, tiles, iters, early, slots, seed, poll=(-1, 0, 1), late_waits=True, publish_upto=None, release_when_all_started=True,
             max_events=200000):
    rng = random.Random(seed)
    publish_upto = early if publish_upto is None else publish_upto
    counters = [0] * nz
    started = [[[False] * tiles for _ in range(nz)] for _ in range(iters)]
    finished = [[[False] * tiles for _ in range(nz)] for _ in range(iters)]
    n_started = [0] * iters
    n_finished = [0] * iters
    per_iter = nz * tiles
    waiting = []   # blocks that hold a slot and spin: (t, z, i)
    running = []   # blocks past their wait: (t, z, i)
    next_block = [0] * iters  # blocks of iteration t not yet dispatched are taken in a RANDOM order (worst case for the hardware)
    order = [rng.sample(range(per_iter), per_iter) for _ in range(iters)]

    def plane_done(t, z):
        return all(finished[t][z])

    def may_run(t, z):
        if t == 0:
            return True
        if z < early:
            return all(counters[min(max(z + d, 0), nz - 1)] >= t * tiles for d in poll)
        return (n_finished[t - 1] == per_iter) if late_waits else True

    def check_start(t, z):
        if t == 0:
            return
        for zz in (z - 1, z, z + 1):
            if 0 <= zz < nz:
                if not plane_done(t - 1, zz):  # RAW on the input field, and WAR: those tiles read the plane this one overwrites
                    raise Violation(f"tile ({t},{z}) runs before plane {zz} of iteration {t - 1} is complete")

    events = 0
    while n_finished[iters - 1] < per_iter:
        events += 1
        if events > max_events:
            raise Violation("no progress (event limit)")
        actions = []
        # dispatch a new block into a free slot: iteration t is eligible when its predecessor has released it
        if len(waiting) + len(running) < slots:
            for t in range(iters):
                if next_block[t] < per_iter:
                    released = t == 0 or (n_started[t - 1] == per_iter if release_when_all_started else n_started[t - 1] > 0)
                    if released:  # (with the real rule at most one iteration has blocks left to dispatch)
                        actions.append(("dispatch", t))
        for b in waiting:
            if may_run(b[0], b[1]):
                actions.append(("go", b))
        for b in running:
            actions.append(("finish", b))
        if not actions:
            raise Violation(f"deadlock: {len(waiting)} blocks spin, none can proceed")
        kind, arg = rng.choice(actions)
        if kind == "dispatch":
            t = arg
            k = order[t][next_block[t]]
            next_block[t] += 1
            z, i = divmod(k, tiles)
            started[t][z][i] = True
            n_started[t] += 1
            waiting.append((t, z, i))
        elif kind == "go":
            waiting.remove(arg)
            check_start(arg[0], arg[1])
            running.append(arg)
        else:
            t, z, i = arg
            running.remove(arg)
            finished[t][z][i] = True
            n_finished[t] += 1
            if z <= publish_upto:
                counters[z] += 1
    return events


SHAPES = [  # nz, tiles per plane, iterations, early, slots
    (8, 2, 4, 3, 6), (8, 2, 4, 8, 5), (6, 3, 5, 1, 4), (12, 1, 4, 4, 3), (5, 4, 3, 2, 40), (16, 2, 3, 5, 9), (4, 1, 6, 4, 1), (3, 2, 5, 2, 2),
]


@pytest.mark.parametrize("nz,tiles,iters,early,slots", SHAPES)
def test_chain_protocol_is_safe_and_live(nz, tiles, iters, early, slots):
    for seed in range(60):
        simulate(nz, tiles, iters, min(early, nz), slots, seed)


def _caught(**broken):
    hits = 0
    for nz, tiles, iters, early, slots in SHAPES:
        for seed in range(60):
            try:
                simulate(nz, tiles, iters, min(early, nz), slots, seed, **broken)
            except Violation:
                hits += 1
    return hits


def test_broken_variants_are_caught():
    assert _caught(poll=(-1, 0)) > 0, "polling planes z-1 and z only must lose a RAW/WAR dependency on plane z+1"
    assert _caught(late_waits=False) > 0, "late blocks that do not wait for the previous launch must be caught"
    # planes 0..early-1 publish only: the tile of plane early-1 waits for a counter nobody increments
    hits = 0
    for nz, tiles, iters, early, slots in SHAPES:
        e = min(early, nz)
        if e < nz:
            try:
                simulate(nz, tiles, iters, e, slots, 1, publish_upto=e - 1)
            except Violation as v:
                hits += "deadlock" in str(v) or "no progress" in str(v)
    assert hits > 0
    assert _caught(release_when_all_started=False) > 0, "a dependent launch released too early can fill the chip with spinning blocks"
