"""torchrun --nproc-per-node N tests/mgpu_check.py — multi-GPU parity + halo transports on a real multi-GPU box.

Not collected by pytest (needs N GPUs): run under `gpurun --gpus N`.  Checks, for each transport and OCC mode, that N
ranks with z-slab partitions reproduce the single-partition CPU oracle bit for bit (REFERENCE arithmetic), then times the
halo update alone and a short OCC / no-OCC loop at benchmark size.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neon_b200 as nb  # noqa: E402
from neon_b200 import problems as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bk = nb.Backend()
    out = {}
    # ---- parity
    from oracle import oracle as O
    dim, iters, omega = (40, 24, 8 * world + 3), 7, 1.3
    for q, dt in ((19, np.float32), (27, np.float64)):
        cls = O.classify(1, *dim)
        mask = O.wall_mask(q, cls)
        ref = O.run(q, O.init_pop(q, cls, dt), cls, mask, omega, iters) if rank == 0 else None
        for transport in ("packed", "views", "ipc", "fused"):
            for occ in (nb.Occ.none, nb.Occ.standard):
                grid = nb.dGrid(bk, dim)
                pop0, pop1, flag = P.setup_device(grid, q, dt, P.CAVITY_SPHERE)
                it = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q,
                                     arith=nb.ARITH_REFERENCE, halo_transport=transport)
                for _ in range(iters):
                    it.run()
                bk.syncAll()
                got = it.getInput().gather()
                if rank == 0:
                    out[f"parity_q{q}_{transport}_{occ.value}"] = bool(np.array_equal(got.view(np.uint8), ref.view(np.uint8)))
                dist.barrier()
    # ---- timing at benchmark size: 1024 x 1024 x 128 per GPU, D3Q19 fp32
    dim = (1024, 1024, 128 * world)
    grid = nb.dGrid(bk, dim)
    pop0, pop1, flag = P.setup_device(grid, 19, np.float32, P.CAVITY)
    om = nb.omega_from_re(dim[0])
    for transport in ("packed", "ipc", "fused"):
        halo = pop0.newHaloUpdate(nb.StencilSemantic.streaming, nb.TransferMode.get, 19, "ipc" if transport == "fused" else transport)
        for _ in range(3):
            halo.run(0)
        bk.syncAll(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(bk.stream(0))
        for _ in range(20):
            halo.run(0)
        e1.record(bk.stream(0))
        bk.syncAll()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=bk.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[f"halo_ms_{transport}"] = float(ms.item())
        out[f"halo_bytes_per_dir"] = halo.bytesPerDirection(+1)
        for occ in (nb.Occ.none, nb.Occ.standard):
            it = nb.LbmIteration(nb.StencilSemantic.streaming, occ, nb.TransferMode.get, pop0, pop1, flag, om, halo_transport=transport)
            for _ in range(5):
                it.run()
            bk.syncAll(); dist.barrier()
            e0.record(bk.stream(0))
            for _ in range(30):
                it.run()
            e1.record(bk.stream(0))
            bk.syncAll()
            ms = torch.tensor([e0.elapsed_time(e1) / 30], device=bk.device)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[f"iter_ms_{transport}_{occ.value}"] = float(ms.item())
            out[f"mlups_{transport}_{occ.value}"] = dim[0] * dim[1] * dim[2] / (float(ms.item()) * 1e3)
            out[f"timeouts_{transport}_{occ.value}"] = it.timeouts()
            dist.barrier()
    if rank == 0:
        print(json.dumps(out, indent=1), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
