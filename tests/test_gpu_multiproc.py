"""Multi-PROCESS halo exchange on the GPU (one process per partition, as bench.py --gpus N runs it), on ONE device.

The reference tests multi-device by oversubscribing one GPU with several partitions
(libNeonDomain/tests/domain-halos/src/runHelper.h:64-67).  The counterpart here: several processes share cuda:0, each
owns one z-slab, the control plane is gloo and the data plane the peer-store transport (CUDA-IPC mapping of the
neighbour's field + device-side flags, neon_b200/ipc.py) — NCCL refuses two ranks on one device.  Result after N
iterations with OCC == the single-partition oracle, bit for bit (parity oracle for n>1).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, dtype_name, dim, iters, occ_name, kind, results, transport="ipc", pipelined=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import neon_b200 as nb
        from neon_b200 import problems as P
        dtype = np.dtype(dtype_name)
        bk = nb.Backend(devices=[0] * world)
        grid = nb.dGrid(bk, dim) if kind == "dGrid" else nb.bGrid(bk, dim)
        pop0, pop1, flag = P.setup_device(grid, q, dtype, P.CAVITY_SPHERE)
        it = nb.LbmIteration(nb.StencilSemantic.streaming, getattr(nb.Occ, occ_name), nb.TransferMode.get, pop0, pop1, flag, 1.25,
                             lattice_q=q, arith=nb.ARITH_REFERENCE, halo_transport=transport, pipelined=pipelined)
        if transport == "ipc":
            kinds = [k for _, k, _, _ in it.lbmTwoPop[0].schedule()]
            assert ("halo_push" in kinds and "halo_wait" in kinds) == pipelined and ("halo" in kinds) == (not pipelined), kinds
        for _ in range(iters):
            it.run()
        bk.syncAll()
        timeouts = it.timeouts()
        out = it.getInput().gather()
        cls, mask = flag.gather()
        if rank == 0:
            results["pop"], results["mask"], results["timeouts"] = out, mask, timeouts
        bk.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("q,dtype,world,occ,kind,pipelined", [
    (19, "float32", 2, "standard", "dGrid", True), (19, "float32", 3, "none", "dGrid", True), (19, "float32", 3, "standard", "dGrid", True),
    (27, "float64", 2, "standard", "dGrid", True), (19, "float32", 2, "standard", "bGrid", True), (27, "float64", 2, "none", "bGrid", True),
    (19, "float32", 2, "standard", "dGrid", False), (19, "float32", 3, "none", "dGrid", False), (19, "float32", 2, "standard", "bGrid", False)])
def test_ipc_halo_across_processes(oracle, q, dtype, world, occ, kind, pipelined):
    """pipelined: faces pushed right after the BOUNDARY kernel that computed them, the consumer only waits (Options.pipelinedHalo);
    otherwise the whole update sits in front of the consumer, as the reference schedules it."""
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    dim, iters = ((36, 20, 23) if kind == "dGrid" else (36, 20, 37)), 8
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, dtype, dim, iters, occ, kind, results, "ipc", pipelined), nprocs=world, join=True)
    nx, ny, nz = dim
    cls = oracle.classify(1, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    ref = oracle.run(q, oracle.init_pop(q, cls, np.dtype(dtype)), cls, mask, 1.25, iters)
    assert results["timeouts"] == 0
    assert np.array_equal(results["mask"], mask)
    assert np.array_equal(results["pop"].view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("q,dtype,world", [(19, "float32", 2), (19, "float32", 3), (27, "float64", 2)])
def test_fused_step_and_halo_kernel_across_processes(oracle, q, dtype, world):
    """nlbm_dense_step_push: ONE kernel per iteration updates the partition and stores the face-crossing populations of its
    boundary planes into the neighbours' ghost planes (peer memory), ordered by device-side counters.  Same bits as the
    single-partition oracle."""
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    dim, iters = (36, 20, 23), 9
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, dtype, dim, iters, "none", "dGrid", results, "fused"), nprocs=world, join=True)
    nx, ny, nz = dim
    cls = oracle.classify(1, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    ref = oracle.run(q, oracle.init_pop(q, cls, np.dtype(dtype)), cls, mask, 1.25, iters)
    assert results["timeouts"] == 0
    assert np.array_equal(results["pop"].view(np.uint8), ref.view(np.uint8))


def _worker_full(rank, world, port, dim, iters, half, r, results, transport):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import neon_b200 as nb
        from neon_b200 import problems as P
        bk = nb.Backend(devices=[0] * world)
        grid = nb.dGrid(bk, dim)
        pop0, pop1, flag = P.setup_device(grid, 19, np.float32, P.CAVITY)
        it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.standard, nb.TransferMode.get, pop0, pop1, flag, 1.3, arith=nb.ARITH_REFERENCE,
                             halo_transport=transport)
        for _ in range(iters):
            it.run()
        bk.syncAll()
        # the planes of this rank inside [mid - half, mid + half), mid = the partition face between rank 0 and rank 1
        mid = grid.z_origin + grid.nz_local if rank == 0 else grid.z_origin
        z0, z1 = max(mid - half, grid.z_origin), min(mid + half, grid.z_origin + grid.nz_local)
        part = None
        if z1 > z0:
            f = it.getInput()
            part = (z0, f.view4[:, z0 - grid.z_origin + grid.z_halo:z1 - grid.z_origin + grid.z_halo, :r, :r].cpu().numpy())
        results[rank] = (part, it.timeouts(), mid)
        bk.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["ipc", "fused"])
def test_partition_face_of_a_full_size_box_matches_the_oracle(oracle, transport):
    """512^3 (configs[1]'s box) as two z-slabs in two processes, OCC standard, pipelined peer-store halo (and the fused step + push
    kernel): the cells around the partition face, next to the x = 0 / y = 0 walls, hold the bits of the CPU oracle.  By the locality
    of the scheme a 48^3 oracle box around that place suffices: after K iterations its artificial z walls have reached K + 1 cells
    inwards (tests/test_gpu_full_size.py does the same for the corners of single-partition boxes)."""
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    S, R, K = 48, 32, 6
    half = S // 2 - (K + 1)  # planes on either side of the face that the artificial walls have not reached
    dim = (512, 512, 512)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_full, args=(2, _free_port(), dim, K, half, R, results, transport), nprocs=2, join=True)
    mid = results[0][2]
    assert mid == 256 and results[0][1] == 0 and results[1][1] == 0
    got = np.concatenate([results[0][0][1], results[1][0][1]], axis=1)  # [19, 2 * half, R, R], planes mid - half .. mid + half
    assert results[0][0][0] == mid - half and results[1][0][0] == mid and got.shape[1] == 2 * half
    cls = oracle.classify(0, S, S, S)
    mask = oracle.wall_mask(19, cls)
    ref = oracle.run(19, oracle.init_pop(19, cls, np.float32), cls, mask, 1.3, K)
    ref = ref[:, S // 2 - half:S // 2 + half, :R, :R]  # small box = [0, S) x [0, S) x [mid - S/2, mid + S/2) of the big one
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))
