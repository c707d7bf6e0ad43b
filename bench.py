#!/usr/bin/env python
"""bench.py — LBM MLUPS of the B200-native Neon hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one LBM iteration (fused pull-stream + BGK collide over every cell, plus the halo update when the box is
split over several GPUs).  MLUPS = Nx*Ny*Nz*steps / elapsed_us, all cells counted, walls included
(benchmarks/lbm-lid-driven-cavity-flow/src/Metrics.h:39-42).

Workloads (lid-driven cavity, Re=100, ulb=0.04, synthetic — there is no input data):
  N=1 default : D3Q19 fp32 512^3 dGrid                         (BASELINE.json configs[1])
  N>1 default : D3Q19 fp32 1024 x 1024 x (128*N), z-slab partitioned, OCC overlap of the halo exchange: the same
                134 M cells per GPU as 512^3 (weak scaling); at N=8 it is the 1024^3 box of configs[2]
  --workload  : cavity512 | slab1024 | cavity1024 (strong: 1024^3 over N GPUs) | d3q27f64 (768 x 768 x 96 per GPU)
                | cavity<N> (cube of edge N)

value     : device-resident throughput, CUDA events on the launching stream, max over ranks.
e2e       : the same job through the public host API with HOST buffers (the reference benchmark's flow,
            RunCavityTwoPop.cu:159-275): pinned host arrays -> updateDeviceData -> wall mask -> K iterations ->
            updateHostData, every copy inside the timed region.
roofline  : algorithmic bytes (2*Q*sizeof(T) per cell) / measured kernel time vs MEASURED_PEAKS.json.
cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref/ref_lbm: Neon's own LbmIterationD3Q19 on its
            CPU backend, built by oracle/Makefile.ref) timed on this box's host cores on a bounded sample.
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
                         "speculative x-face fix-up operands); never changes results")
    ap.add_argument("--no-pipeline", action="store_true", help="N>1: halo update in front of the consumer instead of pushed after BOUNDARY")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=128, help="edge of the CPU sample box")
    ap.add_argument("--cpu-iters", type=int, default=20)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workloads
def workload(args):
    n = args.gpus
    name = args.workload or ("cavity512" if n == 1 else "slab1024")
    if name == "slab1024":
        return dict(name=f"lid-driven cavity D3Q19 fp32 1024x1024x{128 * n} dGrid, z-slab over {n} GPU(s)", q=19, dtype="float32",
                    dim=(1024, 1024, 128 * n), scaling="weak")
    if name == "cavity1024":
        return dict(name=f"lid-driven cavity D3Q19 fp32 1024^3 dGrid, z-slab over {n} GPU(s)", q=19, dtype="float32",
                    dim=(1024, 1024, 1024), scaling="strong")
    if name == "d3q27f64":
        return dict(name=f"lid-driven cavity D3Q27 fp64 768x768x{96 * n} dGrid, z-slab over {n} GPU(s)", q=27, dtype="float64",
                    dim=(768, 768, 96 * n), scaling="weak")
    if name == "sphere":
        # BASELINE.json configs[3]: flow over a sphere on bGrid, 1024 x 512 x 512; integers recorded in SURVEY.md §8d
        return dict(name=f"flow over sphere D3Q19 fp32 1024x512x512 bGrid (8^3 blocks, bounce-back), z block layers over {n} GPU(s)", q=19,
                    dtype="float32", dim=(1024, 512, 512), scaling="strong", grid="bGrid", geom=2, sphere=(392.0, 277.0, 256.0, 60.0),
                    omega=1.0 / (3.0 * 0.04 * 60.0 / 100.0 + 0.5))
    if name.startswith("bcavity"):
        e = int(name[len("bcavity"):])
        return dict(name=f"lid-driven cavity D3Q19 fp32 {e}^3 bGrid (8^3 blocks)", q=19, dtype="float32", dim=(e, e, e),
                    scaling="strong" if n > 1 else "weak", grid="bGrid")
    if name.startswith("cavity"):
        e = int(name[len("cavity"):])
        return dict(name=f"lid-driven cavity D3Q19 fp32 {e}^3 dGrid" + (f", z-slab over {n} GPUs" if n > 1 else ""), q=19,
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args)
    # bounded sample of the workload: the reference needs ~0.26 us per cell update on one core
    n = args.cpu_n
    per_step = max(1, args.cpu_iters // 10)
    iters = args.warmup * 0 + min(args.steps, 100) * per_step
    iters = max(2, min(iters, 60))
    warm = 1
    cb = cpu_reference(n, iters + warm, warm)
    line = {"impl": "reference", "metric": "LBM MLUPS (D3Q19 fp32)", "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb.get("ms_per_iteration_of_sample"), "higher_is_better": True,
            "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": cb["sample"],
                       "note": "reference = Autodesk/Neon's own LbmIterationD3Q19 on its CPU/OpenMP backend (serial executor), unmodified"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import neon_b200 as nb
    from neon_b200 import problems as P
    from neon_b200._capi import opt_tma as capi_opt_tma

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: neon_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workload(args)
    q, dtype, dim = wl["q"], np.dtype(wl["dtype"]), wl["dim"]
    cells = dim[0] * dim[1] * dim[2]
    omega = wl.get("omega", nb.omega_from_re(dim[0]))
    arith = nb.ARITH_FAST if args.arith == "fast" else nb.ARITH_REFERENCE
    opts = nb.opt_vec(args.vec) | nb.opt_rows_log2(args.rows_log2) | nb.opt_kernel({"auto": 0, "direct": 1, "tma": 2}[args.kernel]) \
        | capi_opt_tma(args.tma_l2promo, args.tma_groups) | ((1 << 20) if args.flags_summary_first else 0) | ((args.rpw & 7) << 21) | ((args.experiment & 7) << 24) | ((1 << 27) if args.no_xface_prefetch else 0) | args.opts_extra
    occ = nb.Occ.standard if args.occ == "standard" else nb.Occ.none

    bk = nb.Backend()
    is_block = wl.get("grid", "dGrid") == "bGrid"
    grid = nb.bGrid(bk, dim) if is_block else nb.dGrid(bk, dim)
    pop0, pop1, flag = P.setup_device(grid, q, dtype, wl.get("geom", P.CAVITY), wl.get("sphere"))
    it = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q,
                         arith=arith, opts=opts, halo_transport=args.transport, pipelined=not args.no_pipeline)
    main_stream = bk.stream(0)

    def barrier():
        bk.syncAll()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        it.run()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(main_stream)
    for _ in range(args.steps):
        it.run()
    e1.record(main_stream)
    barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=bk.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    if it.timeouts() != 0:  # a face that never arrived: the numbers would come from stale ghost planes
        raise SystemExit(f"rank {rank}: {it.timeouts()} halo wait(s) timed out — no result")
    ms_step = ms_total / args.steps
    mlups = cells * args.steps / (ms_total * 1e3)

    # launches of OUR kernels per step on this rank: the step kernel per view (+ pack/unpack per neighbour)
    dn, up = grid.neighbours()
    nnb = (dn is not None) + (up is not None)
    # halo per neighbour: ipc = push + signal + wait kernels, packed = pack + unpack kernels (+ NCCL's own), views = NCCL only
    per_nb = {"auto": 3, "ipc": 3, "packed": 2, "views": 0, "fused": 1}[args.transport]
    launches_step = 1 if world == 1 else ((1 if args.transport == "fused" else (2 if occ != nb.Occ.none else 1)) + per_nb * nnb)

    # --- roofline of the dominant kernel (k_dense_step): algorithmic bytes / measured launch duration ------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_cell = 2 * q * dtype.itemsize
    cells_rank = grid.n_blocks * 512 if is_block else dim[0] * dim[1] * grid.nz_local
    # per-launch duration of the step kernel, measured live: at N=1 the timed region holds exactly K launches of it
    kern_ms = None
    if world == 1:
        kern_ms = ms_step
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            key = f"d3q{q}_{'f32' if dtype.itemsize == 4 else 'f64'}_{dim[0]}x{dim[1]}x{cells_rank // (dim[0] * dim[1])}" + ("_bgrid" if is_block else "")
            traffic = tj.get(key, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_block_step" if is_block else "k_dense_step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "bytes_per_cell": bytes_cell, "cells_per_launch": cells_rank, "kernel_ms": kern_ms,
                "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0}

    # --- e2e: host buffers -> device -> K iterations -> host ----------------------------------------------------------
    # Every rank holds the host mirror of ITS slab (plus the in-box ghost planes), as one process per GPU implies.
    e2e = None
    if not args.no_e2e and not is_block:
        del it
        if world > 1:  # unmap the neighbours' fields before their owners free them
            from neon_b200 import ipc
            barrier()
            ipc.close_all()
            barrier()
        pop0.data = pop1.data = None
        torch.cuda.empty_cache()
        nx, ny, nz = dim
        nzl, z0, zh = grid.nz_local, grid.z_origin, grid.z_halo
        lo, hi = max(0, z0 - zh), min(nz, z0 + nzl + zh)  # global planes this rank uploads
        cls3 = P.host_classes(P.CAVITY, (nx, ny, 3))  # plane 0: z-wall, plane 1: interior
        pop3 = P.host_populations(q, cls3, dtype)
        tdt = torch.float32 if dtype.itemsize == 4 else torch.float64
        pop_h = torch.empty((q, hi - lo, ny, nx), dtype=tdt, pin_memory=True)
        pop_np = pop_h.numpy()
        cls = np.empty((hi - lo, ny, nx), np.int32)
        for gz in range(lo, hi):
            w = 0 if gz in (0, nz - 1) else 1
            cls[gz - lo] = cls3[w]
            pop_np[:, gz - lo] = pop3[:, w]
        out_h = pop_h[:, z0 - lo:z0 - lo + nzl]  # the result lands in the same pinned buffer (uploads are done by then)
        f0, f1 = grid.newField("pop0", q, dtype), grid.newField("pop1", q, dtype)
        fl = grid.newFlagField("flag", like=f0)
        barrier()
        w0 = time.perf_counter()
        fl.setClasses(cls, host_z0=lo)
        f0.updateDeviceData(pop_h, host_z0=lo)
        f1.updateDeviceData(pop_h, host_z0=lo)
        fl.computeWallNghMask(q)
        it2 = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, f0, f1, fl, omega, lattice_q=q, arith=arith,
                              opts=opts, halo_transport=args.transport)
        for _ in range(args.steps):
            it2.run()
        it2.getInput().updateHostDataInto(out_h)
        barrier()
        w1 = time.perf_counter()
        if it2.timeouts() != 0:
            raise SystemExit(f"rank {rank}: {it2.timeouts()} halo wait(s) timed out in the end-to-end run — no result")
        secs = torch.tensor([w1 - w0], dtype=torch.float64, device=bk.device)
        traffic_hd = torch.tensor([2.0 * pop_h.numel() * dtype.itemsize + cls.size * 4, float(out_h.numel() * dtype.itemsize)],
                                  dtype=torch.float64, device=bk.device)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            dist.all_reduce(traffic_hd, op=dist.ReduceOp.SUM)
        secs = float(secs.item())
        h2d, d2h = float(traffic_hd[0].item()), float(traffic_hd[1].item())
        e2e = {"value": cells * args.steps / (secs * 1e6), "unit": "MLUPS", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "seconds": secs, "steps": args.steps,
               "note": "whole job through the host API, max over ranks: pinned host populations+classes of every rank's slab -> "
                       "updateDeviceData -> wall mask -> K iterations (with halo updates) -> updateHostData of the result field; "
                       "an LBM iteration has no per-step host input, so bytes are job totals over all ranks / K"}

    cpu = ref_gpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        cpu = cpu_reference(args.cpu_n, args.cpu_iters + 1, 1)
        ref_gpu = reference_gpu_backend()

    if rank == 0:
        line = {"metric": f"LBM MLUPS (D3Q{q} {'fp32' if dtype.itemsize == 4 else 'fp64'})", "value": mlups, "unit": "MLUPS",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32" if dtype.itemsize == 4 else "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "dim": list(dim), "lattice": f"D3Q{q}", "arith": args.arith, "kernel": args.kernel, **({"EXPERIMENT_wrong_results": args.experiment} if args.experiment else {}),
                           "occ": args.occ if world > 1 else "n/a (1 partition)", "halo_transport": args.transport if world > 1 else "n/a",
                           "l2": "inputs exceed L2 (two population fields of %.1f GB per GPU)" % (q * cells_rank * dtype.itemsize / 1e9),
                           "partition": (f"{grid.n_blocks} blocks per GPU" if is_block else f"z-slabs of {grid.nz_local} planes") if world > 1 else "single partition"},
                "roofline": roofline, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "e2e": e2e, "gpu_launches": launches_step * args.steps, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
