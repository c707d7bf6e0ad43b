#!/usr/bin/env python
"""Device-side timeline of ONE multi-GPU LBM iteration (no profiler): a CUDA event in front of and behind every node the
Skeleton issues, per rank.  Shows whether the halo traffic and the BOUNDARY kernel overlap the INTERNAL kernel (OCC).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/halo_timeline.py [--occ none] [--no-pipeline]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neon_b200 as nb  # noqa: E402
from neon_b200 import problems as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--occ", default="standard", choices=["none", "standard"])
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--planes", type=int, default=128, help="z planes per GPU of the 1024 x 1024 x (planes*N) box")
    args = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bk = nb.Backend()
    dim = (1024, 1024, args.planes * world)
    grid = nb.dGrid(bk, dim)
    pop0, pop1, flag = P.setup_device(grid, 19, np.float32, P.CAVITY)
    occ = nb.Occ.standard if args.occ == "standard" else nb.Occ.none
    it = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, pop0, pop1, flag, nb.omega_from_re(dim[0]),
                         pipelined=not args.no_pipeline)
    for _ in range(20):
        it.run()
    bk.syncAll()
    dist.barrier()
    rows = []
    for rep in range(4):  # two iterations of each parity, back to back, the way run() alternates them
        sk = it.lbmTwoPop[it.parity]
        rows.append(sk.timeline())
        it.parity ^= 1
    bk.syncAll()
    assert it.timeouts() == 0
    out = [None] * world
    dist.all_gather_object(out, rows)
    if rank == 0:
        print(json.dumps({"world": world, "dim": dim, "occ": args.occ, "pipelined": not args.no_pipeline}))
        for r, per_rank in enumerate(out):
            for rep, tl in enumerate(per_rank):
                t_end = max(e for *_, e in tl)
                print(f"rank {r} iteration {rep}: {t_end:.3f} ms")
                for stream, kind, name, view, t0, t1 in tl:
                    print(f"    stream {stream} {kind:10s} {(view or '-'):9s} {t0:8.3f} -> {t1:8.3f} ms  ({t1 - t0:6.3f})  {name[:60]}")
    dist.destroy_process_group()


main()
