"""Model check of the producer/consumer stage ring of the TMA kernel (neon_b200/csrc/lbm_step_tma.cuh), no GPU.

Consumer groups wait only on the stages of their own tiles, so a bare parity wait can pass one phase early when the
group count does not divide the stage count (found on the B200 as a hang with 3 stages x 2 groups).  The kernel
therefore tags every stage with the index of the tile it carries; this randomised interleaving model shows that the
tagged protocol never deadlocks or reads a stale stage, and that the untagged one does.
"""
import random

import pytest


def simulate(stages, groups, tiles, seed, check_id=True, max_steps=100000):
    rnd = random.Random(seed)
    full = [0] * stages            # completed phases of full[s]
    empty = [[0, 0] for _ in range(stages)]  # [completed phases, pending arrivals]
    tile_id = [-1] * stages
    inflight = []                  # TMA loads issued and not yet landed (land in random order)
    passed = lambda phases, parity: (phases & 1) != parity   # mbarrier.try_wait.parity
    prod = 0
    warps = [[g, g] for g in range(groups) for _ in range(4)]  # [group, next tile]
    stale = 0
    for _ in range(max_steps):
        agents = ["producer", "land"] + list(range(len(warps)))
        rnd.shuffle(agents)
        for a in agents:
            if a == "land":
                if inflight and rnd.random() < 0.5:
                    full[inflight.pop(rnd.randrange(len(inflight)))] += 1
            elif a == "producer":
                if prod < tiles:
                    s, ph = prod % stages, (prod // stages) & 1
                    if passed(empty[s][0], ph ^ 1):
                        tile_id[s] = prod
                        inflight.append(s)
                        prod += 1
            else:
                g, i = warps[a]
                if i < tiles:
                    s, par = i % stages, (i // stages) & 1
                    if passed(full[s], par) and (not check_id or tile_id[s] == i):
                        stale += full[s] != i // stages + 1
                        empty[s][1] += 1
                        if empty[s][1] == 4:
                            empty[s] = [empty[s][0] + 1, 0]
                        warps[a][1] += groups
        if prod >= tiles and not inflight and all(w[1] >= tiles for w in warps):
            return "done", stale
    return "stuck", stale


@pytest.mark.parametrize("stages", [2, 3, 4, 5, 8])
@pytest.mark.parametrize("groups", [1, 2, 3])
def test_tagged_ring_is_safe(stages, groups):
    for seed in range(12):
        assert simulate(stages, groups, 70, seed) == ("done", 0)


def test_untagged_ring_aliases():
    outcomes = {simulate(3, 2, 70, seed, check_id=False) for seed in range(12)}
    assert any(o[0] == "stuck" or o[1] > 0 for o in outcomes)
