#!/usr/bin/env python
"""bench.py — LBM MLUPS of the B200-native Neon hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one LBM iteration (fused pull-stream + BGK collide over every cell, plus the halo update when the box is
split over several GPUs).  MLUPS = Nx*Ny*Nz*steps / elapsed_us, all cells counted, walls included
(benchmarks/lbm-lid-driven-cavity-flow/src/Metrics.h:39-42).

Workloads (Re=100, ulb=0.04, synthetic — there is no input data):
  N=1 default : D3Q19 fp32 512^3 dGrid lid-driven cavity        (BASELINE.json configs[1])
  N>1 default : D3Q19 fp32 1024 x 1024 x (128*N), z-slab partitioned, OCC overlap of the halo exchange: the same
                134 M cells per GPU as 512^3 (weak scaling); at N=8 it is the 1024^3 box of configs[2]
  --workload  : cavity512 | slab1024 | cavity1024 (configs[2], strong: 1024^3 over N GPUs) | sphere (configs[3], bGrid
                1024x512x512, strong) | d3q27f64 (configs[4], 768 x 768 x 96 per GPU, weak) | cavity<N> | bcavity<N>

The JSON line (headline workload):
  value          device-resident throughput, CUDA events on the launching stream, max over ranks.
  e2e            the same job through the public host API with HOST buffers (the reference benchmark's flow,
                 RunCavityTwoPop.cu:159-275): pinned host mirror -> updateDeviceData -> wall mask -> K iterations ->
                 updateHostData, every copy inside the timed region, with the seconds of each phase.
  roofline       algorithmic bytes (2*Q*sizeof(T) per cell) / measured kernel time vs MEASURED_PEAKS.json.
  arith_reference  the same workload in REFERENCE arithmetic (the reference's bits; the headline runs FAST arithmetic).
  extra_configs  the other BASELINE.json configs measured in the same job on the same N GPUs: cavity1024 (strong),
                 sphere on bGrid, D3Q27 fp64 — each with its own roofline (and e2e for the bGrid one).
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref/ref_lbm: Neon's own LbmIterationD3Q19 on its
                 CPU backend, built by oracle/Makefile.ref) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="")
    ap.add_argument("--arith", default="fast", choices=["fast", "reference"])
    ap.add_argument("--occ", default="standard", choices=["none", "standard"])
    ap.add_argument("--transport", default="auto", choices=["auto", "packed", "views", "ipc", "fused"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "direct", "tma"])
    ap.add_argument("--vec", type=int, default=0)
    ap.add_argument("--tma-l2promo", type=int, default=0)
    ap.add_argument("--tma-groups", type=int, default=0)
    ap.add_argument("--rows-log2", type=int, default=0)
    ap.add_argument("--flags-summary-first", action="store_true")
    ap.add_argument("--rpw", type=int, default=0, help="rows per warp selector: 0 default, else log2(rows)+1")
    ap.add_argument("--experiment", type=int, default=0, choices=[0, 3],
                    help="MEASUREMENT ONLY (wrong results): 3 = flags honoured, wall fix-ups and kept wall values skipped")
    ap.add_argument("--no-xface-prefetch", action="store_true")
    ap.add_argument("--opts-extra", type=lambda v: int(v, 0), default=0,
                    help="OR-ed into the step options (include/neon_lbm.h: 1<<28 flag words with the populations, 1<<29 no "
                         "speculative x-face fix-up operands, 1<<30 literal transcription in REFERENCE arithmetic); never changes results")
    ap.add_argument("--no-pipeline", action="store_true", help="N>1: halo update in front of the consumer instead of pushed after BOUNDARY")
    ap.add_argument("--graph-iters", type=int, default=-1,
                    help="one device: iterations per host call — a CUDA-graph replay, or one multi-iteration launch with --persistent "
                         "(0 = plain launches; default: 0 for boxes above 2^22 cells, else 10)")
    ap.add_argument("--persistent", action="store_true", help="nlbm_dense_step_n called directly (G iterations per call, no CUDA graph)")
    ap.add_argument("--chain-early", type=int, default=0, help="with --persistent: NLBM_OPT_CHAIN_EARLY (0 default, 1..14: 2^(e-1) planes start on the plane counters, 15: all)")
    ap.add_argument("--chain-graph", action="store_true", help="graph replay of the launch chain whatever the box size")
    ap.add_argument("--no-chain", action="store_true", help="graph replay of single-iteration launches whatever the box size")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline line only: no arith_reference, no extra_configs")
    ap.add_argument("--extra-steps", type=int, default=30, help="timed iterations of each extra config (warm-up 5)")
    ap.add_argument("--cpu-n", type=int, default=64, help="edge of the cpu_baseline box (configs[0]: 64)")
    ap.add_argument("--cpu-iters", type=int, default=100, help="iterations of the cpu_baseline run incl. 10 warm-up (configs[0]: 100)")
    ap.add_argument("--ref-n", type=int, default=128, help="--impl reference: edge of the box one sample step iterates")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workloads
def workload(name: str, n: int):
    name = name or ("cavity512" if n == 1 else "slab1024")
    if name == "slab1024":
        return dict(key=name, name=f"lid-driven cavity D3Q19 fp32 1024x1024x{128 * n} dGrid, z-slab over {n} GPU(s)", q=19, dtype="float32",
                    dim=(1024, 1024, 128 * n), scaling="weak")
    if name == "cavity1024":
        return dict(key=name, name=f"lid-driven cavity D3Q19 fp32 1024^3 dGrid, z-slab over {n} GPU(s)", q=19, dtype="float32",
                    dim=(1024, 1024, 1024), scaling="strong")
    if name == "d3q27f64":
        return dict(key=name, name=f"lid-driven cavity D3Q27 fp64 768x768x{96 * n} dGrid, z-slab over {n} GPU(s)", q=27, dtype="float64",
                    dim=(768, 768, 96 * n), scaling="weak")
    if name == "sphere":
        # BASELINE.json configs[3]: flow over a sphere on bGrid, 1024 x 512 x 512; integers recorded in SURVEY.md §8d
        return dict(key=name, name=f"flow over sphere D3Q19 fp32 1024x512x512 bGrid (8^3 blocks, bounce-back), z block layers over {n} GPU(s)", q=19,
                    dtype="float32", dim=(1024, 512, 512), scaling="strong", grid="bGrid", geom=2, sphere=(392.0, 277.0, 256.0, 60.0),
                    omega=1.0 / (3.0 * 0.04 * 60.0 / 100.0 + 0.5))
    if name.startswith("box"):  # box<NX>x<NY>x<NZ>: lid-driven cavity in an arbitrary box (layout experiments)
        e = [int(v) for v in name[3:].split("x")]
        return dict(key=name, name=f"lid-driven cavity D3Q19 fp32 {e[0]}x{e[1]}x{e[2]} dGrid", q=19, dtype="float32", dim=tuple(e),
                    scaling="strong" if n > 1 else "weak")
    if name.startswith("bcavity"):
        e = int(name[len("bcavity"):])
        return dict(key=name, name=f"lid-driven cavity D3Q19 fp32 {e}^3 bGrid (8^3 blocks)", q=19, dtype="float32", dim=(e, e, e),
                    scaling="strong" if n > 1 else "weak", grid="bGrid")
    if name.startswith("cavity"):
        e = int(name[len("cavity"):])
        return dict(key=name, name=f"lid-driven cavity D3Q19 fp32 {e}^3 dGrid" + (f", z-slab over {n} GPUs" if n > 1 else ""), q=19,
                    dtype="float32", dim=(e, e, e), scaling="strong" if n > 1 else "weak")
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            p = [v.strip() for v in line.split(",")]
            if len(p) < 7:
                continue
            try:
                smax = max(smax, float(p[1]))
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(p[0]))
                    for nme, v in zip(names, p[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference(n: int, iters: int, warm: int, fp: str = "float"):
    """Times the UNMODIFIED reference CPU backend (oracle/_ref/ref_lbm) or, if that binary is absent, the C oracle
    port, on this box's host cores.  The reference's dense OpenMP executor is serial on Linux
    (libNeonSet/include/Neon/set/LambdaExecutor.h:68,84 — pragmas commented out): one core does the work."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_lbm")
    sample = f"lid-driven cavity D3Q19 {'fp32' if fp == 'float' else 'fp64'} {n}^3 dGrid, {warm}+{iters - warm} iterations"
    if os.path.exists(ref):
        try:
            out = subprocess.run([ref, "--n", str(n), "--iters", str(iters), "--bench", str(warm), "--fp", fp], capture_output=True,
                                 text=True, timeout=1200, cwd="/tmp")
            for line in out.stdout.splitlines():
                if line.startswith("{") and "ref_bench" in line:
                    r = json.loads(line)
                    return {"value": r["mlups"], "unit": "MLUPS", "cores": 1, "kind": "reference", "sample": sample,
                            "threads_available": os.cpu_count(), "ms_per_iteration_of_sample": r["elapsed_us"] / max(1, r["timed_iters"]) / 1e3}
        except (OSError, subprocess.SubprocessError, ValueError):
            pass
    import numpy as np
    from oracle import oracle as O
    cls = O.classify(0, n, n, n)
    mask = O.wall_mask(19, cls)
    dt = np.float32 if fp == "float" else np.float64
    a = O.init_pop(19, cls, dt)
    b = a.copy()
    om = O.omega_cavity(n)
    for _ in range(warm):
        O.step(19, a, b, cls, mask, om)
        a, b = b, a
    t0 = time.perf_counter()
    for _ in range(iters - warm):
        O.step(19, a, b, cls, mask, om)
        a, b = b, a
    dt_s = time.perf_counter() - t0
    return {"value": n ** 3 * (iters - warm) / (dt_s * 1e6), "unit": "MLUPS", "cores": 1, "kind": "port", "sample": sample,
            "threads_available": os.cpu_count(), "ms_per_iteration_of_sample": dt_s * 1e3 / max(1, iters - warm)}


def reference_gpu_backend(n: int = 256, iters: int = 60, warm: int = 10):
    """The UNMODIFIED reference on ITS OWN CUDA backend (generic lambda kernel, nvcc -arch=sm_100) on this GPU — the
    reference GPU number BASELINE.md §4.3 asks for.  Reported next to the bench line, never part of any timed region."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_lbm")
    if not os.path.exists(ref):
        return None
    try:
        out = subprocess.run([ref, "--device", "gpu", "--n", str(n), "--iters", str(iters), "--bench", str(warm), "--fp", "float",
                              "--grid", "dGrid"], capture_output=True, text=True, timeout=300, cwd="/tmp")
        for line in out.stdout.splitlines():
            if line.startswith("{") and "ref_bench" in line:
                r = json.loads(line)
                return {"value": r["mlups"], "unit": "MLUPS", "kind": "reference CUDA backend (Neon v0.3.3, unmodified, sm_100)",
                        "sample": f"lid-driven cavity D3Q19 fp32 {n}^3 dGrid, {warm}+{iters - warm} iterations, 1 GPU"}
    except (OSError, subprocess.SubprocessError, ValueError):
        pass
    return None


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (unmodified Neon, CPU backend) on the config of
    the B200 arm.  One step = one LBM iteration of the reference on a BOUNDED SAMPLE of the workload — the same lid-driven
    cavity on a box of --ref-n^3 cells (the reference needs ~0.13 us per cell update on its one effective core, so a
    512^3 iteration would take 18 s) — for exactly W warm-up + K timed steps.  MLUPS (cells of the sample x K / time) is the
    workload's metric; on the CPU it does not depend on the box size beyond cache effects."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload, args.gpus)
    n, k, w = args.ref_n, max(1, args.steps), max(0, args.warmup)
    cb = cpu_reference(n, k + w, w)
    cb["sample"] = (f"each step = one iteration of the unmodified reference on a {n}^3 box of the same lid-driven cavity "
                    f"(D3Q19 fp32, Re=100, ulb=0.04): {w} warm-up + {k} timed steps")
    line = {"impl": "reference", "metric": "LBM MLUPS (D3Q19 fp32)", "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": k, "warmup": w, "ms_per_step": cb.get("ms_per_iteration_of_sample"), "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": cb["sample"],
                       "note": "reference = Autodesk/Neon's own LbmIterationD3Q19 on its CPU/OpenMP backend (serial executor), unmodified"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
class Job:
    """Everything one bench process shares between the workloads it measures."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import neon_b200 as nb
        self.args, self.torch, self.dist, self.nb = args, torch, dist, nb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: neon_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.bk = nb.Backend()
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, measured)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        self.traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                self.traffic = json.load(open(tpath))
            except (OSError, ValueError):
                self.traffic = {}

    def barrier(self):
        self.bk.syncAll()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.bk.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.bk.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def opts(self):
        a, nb = self.args, self.nb
        from neon_b200._capi import opt_tma
        return nb.opt_vec(a.vec) | nb.opt_rows_log2(a.rows_log2) | nb.opt_kernel({"auto": 0, "direct": 1, "tma": 2}[a.kernel]) \
            | opt_tma(a.tma_l2promo, a.tma_groups) | ((1 << 20) if a.flags_summary_first else 0) | ((a.rpw & 7) << 21) \
            | ((a.experiment & 7) << 24) | ((1 << 27) if a.no_xface_prefetch else 0) | a.opts_extra | ((a.chain_early & 15) << 16)

    def release(self, *objs):
        """Frees device memory between workloads: peer mappings of the neighbours' fields first, then the tensors."""
        if self.world > 1:
            from neon_b200 import ipc
            self.barrier()
            ipc.close_all()
            self.barrier()
        for o in objs:
            for attr in ("data", "words", "_wall_cache", "view4", "view3", "cells"):
                if hasattr(o, attr):
                    setattr(o, attr, None)
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()

    def l2_copy(self, nbytes: int, ms_step: float):
        """A -> B, B -> A copies of ``nbytes`` (one population field) replayed from a CUDA graph: the time a bare copy of one
        iteration's bytes takes when everything stays in L2, and the iteration's time relative to it."""
        torch = self.torch
        n = nbytes // 4
        a, b = torch.empty(n, dtype=torch.float32, device=self.bk.device), torch.ones(n, dtype=torch.float32, device=self.bk.device)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                a.copy_(b)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            g.capture_begin()
            for _ in range(25):
                a.copy_(b)
                b.copy_(a)
            g.capture_end()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10):
                g.replay()
            e1.record(s)
            s.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 500
        del g, a, b
        return {"us_per_copy": us, "GB/s": 2 * n * 4 / us / 1e3, "iteration_us": ms_step * 1e3, "iteration_over_copy": ms_step * 1e3 / us,
                "note": "library copy kernel (torch) moving the bytes of one iteration between two L2-resident buffers, 50 per graph replay"}

    # ------------------------------------------------------------------------------------------ one device-resident run
    def measure(self, wl, arith_name: str, steps: int, warmup: int, sample_clocks: bool = False):
        import numpy as np
        from neon_b200 import problems as P
        nb, bk, args, torch = self.nb, self.bk, self.args, self.torch
        q, dtype, dim = wl["q"], np.dtype(wl["dtype"]), wl["dim"]
        cells = dim[0] * dim[1] * dim[2]
        omega = wl.get("omega", nb.omega_from_re(dim[0]))
        arith = nb.ARITH_FAST if arith_name == "fast" else nb.ARITH_REFERENCE
        opts = self.opts()
        occ = nb.Occ.standard if args.occ == "standard" else nb.Occ.none
        is_block = wl.get("grid", "dGrid") == "bGrid"
        grid = nb.bGrid(bk, dim) if is_block else nb.dGrid(bk, dim)
        pop0, pop1, flag = P.setup_device(grid, q, dtype, wl.get("geom", P.CAVITY), wl.get("sphere"))
        # one device, small boxes: the launch of one iteration costs as much as the iteration (64^3: 12.6 us per step of
        # which ~5 us is launch overhead), so G iterations are captured once into a CUDA graph and replayed — as G launches of the
        # step kernel, or (boxes of more than one chip-load of blocks, LbmIteration.chainPays) as the launch chain of
        # nlbm_dense_step_n, whose iterations overlap tail and ramp-up
        graph_iters = args.graph_iters
        if graph_iters < 0:
            graph_iters = 10 if (self.world == 1 and cells <= (1 << 24)) else 0
            # the launch chain starts every replay with an ordinary launch (behind a memset of its counters): longer replays, fewer drains
            if graph_iters and not is_block and 400_000 < cells and not args.no_chain and steps >= 100:
                graph_iters = 50
        if self.world > 1:
            graph_iters = 0
        it = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q,
                             arith=arith, opts=opts, halo_transport=args.transport, pipelined=not args.no_pipeline)
        main_stream = bk.stream(0)
        per_call = 1
        runner = it.run
        small_mode = "plain launches"
        if graph_iters > 1 and not is_block and args.persistent:
            per_call = graph_iters + (graph_iters & 1)  # LbmIteration.runMany: G iterations per library call, not captured
            small_mode = "nlbm_dense_step_n launch chain, direct calls"

            def runner():
                it.runMany(per_call)
        elif graph_iters > 1:
            per_call = graph_iters + (graph_iters & 1)
            many = False if (args.no_chain or is_block) else (True if args.chain_graph else None)
            it.runGraph(per_call, many=many)  # builds the graph (LbmIteration.runGraph: G iterations per host call)
            chain = (many is True) or (many is None and it.chainPays())
            small_mode = "CUDA graph of the nlbm_dense_step_n launch chain" if chain else "CUDA graph of step launches"

            def runner():
                it.runGraph(per_call, many=many)
        steps = max(per_call, steps // per_call * per_call)
        # untimed iterations in front of the timed region: the W the caller asked for, and never fewer than 20 when the box is split
        # (the first iterations of a multi-GPU run also warm the peer mappings and the NVLink paths up: 3.22 against 3.20 ms per
        # iteration at 8 GPUs when only 5 precede the timer)
        untimed = max(warmup, 20) if self.world > 1 else warmup
        for _ in range(max(1, untimed // per_call)):
            runner()
        self.barrier()
        sampler = ClockSampler(self.local) if (self.rank == 0 and sample_clocks) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(main_stream)
        for _ in range(steps // per_call):
            runner()
        e1.record(main_stream)
        self.barrier()
        t1 = time.time()
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop(t0, t1) if sampler else None
        if it.timeouts() != 0:  # a face that never arrived: the numbers would come from stale ghost planes
            raise SystemExit(f"rank {self.rank}: {it.timeouts()} halo wait(s) timed out — no result")
        ms_step = ms_total / steps
        mlups = cells * steps / (ms_total * 1e3)

        # launches of OUR kernels per step on this rank: the step kernel per view (+ halo kernels per neighbour)
        dn, up = grid.neighbours()
        nnb = (dn is not None) + (up is not None)
        pipelined = self.world > 1 and not args.no_pipeline and args.transport in ("auto", "ipc") and occ != nb.Occ.none
        if self.world == 1:
            launches_step = 1
        elif args.transport == "fused":
            launches_step = 1 + nnb  # step+push kernel, one flag wait per neighbour
        elif pipelined:
            launches_step = 2 + 2  # INTERNAL + BOUNDARY, one face-push launch (both faces + signals), one two-flag wait
        else:
            per_nb = {"auto": 3, "ipc": 3, "packed": 2, "views": 0}[args.transport]
            launches_step = (2 if occ != nb.Occ.none else 1) + per_nb * nnb

        # --- roofline of the dominant kernel: algorithmic bytes / measured launch duration -------------------------------
        bytes_cell = 2 * q * dtype.itemsize
        cells_rank = grid.n_blocks * 512 if is_block else dim[0] * dim[1] * grid.nz_local
        if self.world == 1:
            kern_ms = ms_step  # the timed region holds exactly K launches of the step kernel and nothing else
        else:
            c = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop0, pop1, flag, omega, q, None, arith, opts)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            bk.syncAll()
            a0.record(main_stream)
            for _ in range(5):
                c.run(0, nb.DataView.STANDARD)
            a1.record(main_stream)
            bk.syncAll()
            kern_ms = a0.elapsed_time(a1) / 5
        achieved = bytes_cell * cells_rank / (kern_ms * 1e-3) / 1e9
        key = f"d3q{q}_{'f32' if dtype.itemsize == 4 else 'f64'}_{dim[0]}x{dim[1]}x{cells_rank // (dim[0] * dim[1])}" + ("_bgrid" if is_block else "")
        traffic = self.traffic.get(key, {}).get("dram_bytes_per_launch") if arith_name == "fast" else None
        l2_note = "inputs exceed L2 (two population fields of %.2f GB per GPU)" % (q * cells_rank * dtype.itemsize / 1e9)
        roofline = {"bound": "hbm", "kernel": "k_block_step" if is_block else ("k_dense_chain" if "chain" in small_mode else "k_dense_step"), "achieved": achieved, "peak": self.peak,
                    "unit": "GB/s", "frac": achieved / self.peak, "traffic": traffic, "bytes_per_cell": bytes_cell,
                    "cells_per_launch": cells_rank, "kernel_ms": kern_ms, "peak_source": self.peak_src,
                    "frac_of_nominal_8TBps": achieved / 8000.0}
        if 2 * q * cells_rank * dtype.itemsize < 100e6:
            # both fields fit the 126 MB L2: HBM is not the ceiling of such a box, the L2 and the launch rate are
            roofline["note"] = ("two population fields of %.1f MB fit the 126 MB L2: the HBM fraction is reported for continuity, "
                                "it is not a bound here" % (2 * q * cells_rank * dtype.itemsize / 1e6))
            l2_note = "inputs FIT L2 (%.1f MB): not an HBM measurement" % (2 * q * cells_rank * dtype.itemsize / 1e6)
            # the ceiling that applies instead: what a bare copy of the same bytes reaches when both buffers stay in L2 (torch's
            # copy kernel, the one behind MEASURED_PEAKS.json's HBM number; 50 copies per CUDA-graph replay), measured here
            try:
                roofline["l2_resident_copy"] = self.l2_copy(q * cells_rank * dtype.itemsize, ms_step)
            except Exception as ex:  # a ceiling that cannot be measured must not sink the line
                roofline["l2_resident_copy"] = {"error": f"{type(ex).__name__}: {ex}"}
        res = {"workload": wl["name"], "key": wl["key"], "metric": f"LBM MLUPS (D3Q{q} {'fp32' if dtype.itemsize == 4 else 'fp64'})",
               "value": mlups, "unit": "MLUPS", "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "scaling": wl["scaling"],
               "dtype": "f32" if dtype.itemsize == 4 else "f64", "arith": arith_name, "dim": list(dim), "lattice": f"D3Q{q}",
               "grid": "bGrid" if is_block else "dGrid", "roofline": roofline, "gpu_launches": int(round(launches_step * steps)),
               "graph_iters": per_call if graph_iters > 1 else 0,
               "iterations_per_launch": 1,
               "issue": small_mode, "untimed_iterations": max(1, untimed // per_call) * per_call, "l2": l2_note, "clocks": clocks,
               "partition": ((f"{grid.n_blocks} blocks per GPU" if is_block else f"z-slabs of {grid.nz_local} planes")
                             if self.world > 1 else "single partition")}
        del it, runner
        self.release(pop0, pop1, flag)
        return res

    # ------------------------------------------------------------------------------------------ end to end, host buffers
    def e2e_dense(self, wl, arith_name: str, steps: int):
        """Host mirror of every rank's slab (pinned) -> device -> K iterations -> host, every copy inside the timed region.
        The population mirror crosses the bus once (the second field of the two-field scheme is a device copy of the first,
        dField.copyFrom), classes travel as one byte per cell."""
        import numpy as np
        from neon_b200 import problems as P
        nb, bk, args, torch = self.nb, self.bk, self.args, self.torch
        q, dtype, dim = wl["q"], np.dtype(wl["dtype"]), wl["dim"]
        cells = dim[0] * dim[1] * dim[2]
        omega = wl.get("omega", nb.omega_from_re(dim[0]))
        arith = nb.ARITH_FAST if arith_name == "fast" else nb.ARITH_REFERENCE
        occ = nb.Occ.standard if args.occ == "standard" else nb.Occ.none
        grid = nb.dGrid(bk, dim)
        nx, ny, nz = dim
        nzl, z0, zh = grid.nz_local, grid.z_origin, grid.z_halo
        lo, hi = max(0, z0 - zh), min(nz, z0 + nzl + zh)  # global planes this rank uploads
        cls3 = P.host_classes(P.CAVITY, (nx, ny, 3))  # plane 0: z-wall, plane 1: interior
        pop3 = P.host_populations(q, cls3, dtype)
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        pop_h = torch.empty((q, hi - lo, ny, nx), dtype=tdt, pin_memory=True)
        cls_h = torch.empty((hi - lo, ny, nx), dtype=torch.uint8, pin_memory=True)
        pop_np, cls_np = pop_h.numpy(), cls_h.numpy()
        for gz in range(lo, hi):
            w = 0 if gz in (0, nz - 1) else 1
            cls_np[gz - lo] = cls3[w]
            pop_np[:, gz - lo] = pop3[:, w]
        out_h = pop_h[:, z0 - lo:z0 - lo + nzl]  # the result lands in the same pinned buffer (the upload is done by then)
        f0, f1 = grid.newField("pop0", q, dtype), grid.newField("pop1", q, dtype)
        fl = grid.newFlagField("flag", like=f0)
        self.barrier()
        w0 = time.perf_counter()
        fl.setClasses(cls_h, host_z0=lo)
        f0.updateDeviceData(pop_h, host_z0=lo)
        f1.copyFrom(f0)
        fl.computeWallNghMask(q)
        bk.syncAll()
        w1 = time.perf_counter()
        it2 = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, f0, f1, fl, omega, lattice_q=q, arith=arith,
                              opts=self.opts(), halo_transport=args.transport, pipelined=not args.no_pipeline)
        for _ in range(steps):
            it2.run()
        bk.syncAll()
        w2 = time.perf_counter()
        it2.getInput().updateHostDataInto(out_h)
        self.barrier()
        w3 = time.perf_counter()
        if it2.timeouts() != 0:
            raise SystemExit(f"rank {self.rank}: {it2.timeouts()} halo wait(s) timed out in the end-to-end run — no result")
        secs = self.max_over_ranks(w3 - w0)
        h2d = self.sum_over_ranks(float(pop_h.numel() * dtype.itemsize + cls_h.numel()))
        d2h = self.sum_over_ranks(float(out_h.numel() * dtype.itemsize))
        phases = {"h2d_and_setup_s": self.max_over_ranks(w1 - w0), "iterate_s": self.max_over_ranks(w2 - w1),
                  "d2h_s": self.max_over_ranks(w3 - w2)}
        checksum = float(out_h[q // 2, nzl // 2, ny // 2, nx // 2])  # the result is on the host: read one value of it
        del it2
        self.release(f0, f1, fl)
        return {"value": cells * steps / (secs * 1e6), "unit": "MLUPS", "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
                "seconds": secs, "steps": steps, **phases, "host_gbps": {"h2d": h2d / phases["h2d_and_setup_s"] / 1e9, "d2h": d2h / phases["d2h_s"] / 1e9},
                "result_probe": checksum,
                "note": "whole job through the host API, max over ranks: pinned host mirror (populations once + one byte of class per "
                        "cell) of every rank's slab -> updateDeviceData -> device copy into the second field -> wall mask -> K iterations "
                        "(with halo updates) -> updateHostData of the result field; an LBM iteration has no per-step host input, so "
                        "bytes are job totals over all ranks / K"}

    def e2e_block(self, wl, arith_name: str, steps: int):
        """bGrid: the host mirror is kept in the field's own block layout, as Neon's bField keeps it (bField host mirror:
        libNeonDomain/include/Neon/domain/details/bGrid/bField_imp.h).  Classes are made on the device (the sphere is
        analytic); the populations are initialised on the HOST from the classes read back, uploaded, iterated, downloaded."""
        import numpy as np
        from neon_b200 import problems as P
        nb, bk, args, torch = self.nb, self.bk, self.args, self.torch
        q, dtype, dim = wl["q"], np.dtype(wl["dtype"]), wl["dim"]
        cells = dim[0] * dim[1] * dim[2]
        omega = wl.get("omega", nb.omega_from_re(dim[0]))
        arith = nb.ARITH_FAST if arith_name == "fast" else nb.ARITH_REFERENCE
        occ = nb.Occ.standard if args.occ == "standard" else nb.Occ.none
        grid = nb.bGrid(bk, dim)
        f0, f1 = grid.newField("pop0", q, dtype), grid.newField("pop1", q, dtype)
        fl = grid.newFlagField("flag", like=f0)
        fl.classify(wl.get("geom", P.CAVITY), wl.get("sphere"))
        bk.syncAll()
        nba = grid.n_blocks_alloc
        cls_b = ((fl.cells[:nba].cpu().numpy().view(np.uint32) >> 28) & 3).astype(np.uint8)  # [blocks, 512]
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        pop_h = torch.empty((q, nba, 512), dtype=tdt, pin_memory=True)
        pop_np = pop_h.numpy()
        L = __import__("neon_b200.lattice", fromlist=["lattice"]).lattice(q)
        bulk, moving = cls_b == nb.BULK, cls_b == nb.MOVING_WALL
        for k in range(q):
            wall = dtype.type(-6.0 * L.t[k] * 0.04 * (L.c[k, 0] * 1.0 + L.c[k, 1] * 0.0 + L.c[k, 2] * 0.0))
            pop_np[k] = np.where(bulk, dtype.type(L.t[k]), np.where(moving, wall, dtype.type(0)))
        self.barrier()
        w0 = time.perf_counter()
        f0.updateDeviceBlocks(pop_h)
        f1.copyFrom(f0)
        fl.computeWallNghMask(q)
        bk.syncAll()
        w1 = time.perf_counter()
        it2 = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, f0, f1, fl, omega, lattice_q=q, arith=arith,
                              opts=self.opts(), halo_transport=args.transport, pipelined=not args.no_pipeline)
        for _ in range(steps):
            it2.run()
        bk.syncAll()
        w2 = time.perf_counter()
        it2.getInput().updateHostBlocksInto(pop_h)
        self.barrier()
        w3 = time.perf_counter()
        if it2.timeouts() != 0:
            raise SystemExit(f"rank {self.rank}: {it2.timeouts()} halo wait(s) timed out in the end-to-end run — no result")
        secs = self.max_over_ranks(w3 - w0)
        h2d = self.sum_over_ranks(float(pop_h.numel() * dtype.itemsize))
        d2h = self.sum_over_ranks(float(q * grid.n_blocks * 512 * dtype.itemsize))
        phases = {"h2d_and_setup_s": self.max_over_ranks(w1 - w0), "iterate_s": self.max_over_ranks(w2 - w1),
                  "d2h_s": self.max_over_ranks(w3 - w2)}
        checksum = float(pop_h[q // 2, grid.n_blocks // 2, 256])
        del it2
        self.release(f0, f1, fl)
        return {"value": cells * steps / (secs * 1e6), "unit": "MLUPS", "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
                "seconds": secs, "steps": steps, **phases, "result_probe": checksum,
                "note": "host mirror in block layout (pinned) -> device -> device copy into the second field -> wall mask -> K "
                        "iterations -> host; classes are made on the device, populations on the host"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    job = Job(args)
    wl = workload(args.workload, args.gpus)
    head = job.measure(wl, args.arith, args.steps, args.warmup, sample_clocks=True)

    e2e = None
    if not args.no_e2e:
        e2e = job.e2e_block(wl, args.arith, head["steps"]) if head["grid"] == "bGrid" else job.e2e_dense(wl, args.arith, head["steps"])

    arith_ref, extras = None, []
    if not args.no_extras:
        ks, kw = max(2, min(args.extra_steps, args.steps)), min(5, args.warmup)
        if args.arith == "fast":
            r = job.measure(wl, "reference", ks, kw)
            arith_ref = {"value": r["value"], "unit": "MLUPS", "ms_per_step": r["ms_per_step"], "frac": r["roofline"]["frac"],
                         "frac_of_nominal_8TBps": r["roofline"]["frac_of_nominal_8TBps"], "steps": r["steps"],
                         "note": "same workload, NLBM_ARITH_REFERENCE: bit-exact with the reference's CPU build (tests/), evaluated "
                                 "through the conversion-lean re-arrangement of csrc/lbm_collide_exact.cuh"}
        for name in ("cavity1024", "sphere", "d3q27f64", "cavity64", "cavity128"):
            if name == wl["key"]:
                continue
            if name in ("cavity64", "cavity128") and job.world > 1:
                continue
            w2 = workload(name, args.gpus)
            try:
                small = name in ("cavity64", "cavity128")  # 11 / 58 us per iteration: 400 of them, or the timer measures the graph launches
                r = job.measure(w2, args.arith, 400 if small else ks, 40 if small else kw)
                if w2.get("grid") == "bGrid" and not args.no_e2e:
                    r["e2e"] = job.e2e_block(w2, args.arith, r["steps"])
                r.pop("clocks", None)
                extras.append(r)
            except Exception as ex:  # an extra that cannot run (e.g. not enough planes for N partitions) must not sink the headline
                extras.append({"workload": w2["name"], "key": name, "error": f"{type(ex).__name__}: {ex}"})

    cpu = ref_gpu = None
    if job.rank == 0 and not args.no_cpu and job.world == 1:
        # configs[0] exactly: 64^3, 10 warm-up + 90 timed iterations on the reference's CPU backend
        cpu = cpu_reference(args.cpu_n, args.cpu_iters, 10 if args.cpu_iters > 10 else 1)
        ref_gpu = reference_gpu_backend()

    if job.rank == 0:
        line = {"metric": head["metric"], "value": head["value"], "unit": "MLUPS", "n_gpus": job.world, "steps": head["steps"],
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": wl["scaling"],
                "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
                "config": {"workload": wl["name"], "dim": head["dim"], "lattice": head["lattice"], "arith": args.arith, "kernel": args.kernel,
                           **({"EXPERIMENT_wrong_results": args.experiment} if args.experiment else {}),
                           "occ": args.occ if job.world > 1 else "n/a (1 partition)",
                           "halo_transport": args.transport if job.world > 1 else "n/a", "l2": head["l2"], "partition": head["partition"],
                           "graph_iters": head["graph_iters"], "iterations_per_launch": head["iterations_per_launch"], "issue": head["issue"],
                           "untimed_iterations": head["untimed_iterations"]},
                "roofline": head["roofline"], "cpu_baseline": cpu, "reference_gpu": ref_gpu, "e2e": e2e,
                "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "arith_reference": arith_ref, "extra_configs": extras}
        print(json.dumps(line), flush=True)
    if job.world > 1:
        job.dist.destroy_process_group()


if __name__ == "__main__":
    main()
