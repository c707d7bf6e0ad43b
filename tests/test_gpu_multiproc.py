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
