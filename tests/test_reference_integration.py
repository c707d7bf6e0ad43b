"""INTEGRATION.md §3 executed: the reference's OWN host code — Neon::Backend, dGrid / dField, Loader tokens, Skeleton with its
OCC split and halo update, LbmIterationD3Q19 — driving libneon_lbm.so.

oracle/_ref/ref_lbm_b200 (oracle/Makefile.refb200) is oracle/ref_driver.cu, the driver that produced the golden dumps from
the unmodified reference, compiled against an include overlay in which ONLY the body of LbmContainers::iteration
(benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:283-325) is replaced by the call into the C ABI
(integration/lbm_iteration_b200.inc), and linked with the reference libraries built from /root/reference.  Boxes whose rows are a
multiple of 512 bytes are laid out by Neon's dField exactly as the library wants them, so no
library source changes at all.  The binary is built in the container (where /root/reference exists) and travels to the GPU
box with the repository snapshot; the test skips if it is absent.
"""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_lbm_b200")


def run(tmp, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([BIN, "--device", "gpu", *[str(a) for a in args]], cwd=tmp, capture_output=True, text=True, timeout=600, env=e)


@pytest.fixture(scope="module")
def have_bin():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_lbm_b200 not built (make -f oracle/Makefile.refb200 refb200, needs /root/reference)")
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    return torch.cuda.device_count()


@pytest.mark.parametrize("dim,geom,iters,fp,ndev,occ", [
    ((128, 20, 18), "sphere", 12, "float", 1, "none"),
    ((128, 12, 24), "cavity", 10, "float", 2, "none"),      # two partitions: the reference's halo update feeds the kernel's ghost planes
    ((128, 12, 24), "sphere", 10, "float", 2, "standard"),  # the reference's OCC split: INTERNAL and BOUNDARY views of the kernel
    ((128, 10, 36), "sphere", 8, "float", 3, "standard"),
    ((64, 16, 16), "sphere", 10, "double", 1, "none"),      # fp64: 64 cells are 512 bytes
    ((64, 12, 24), "sphere", 8, "double", 2, "standard"),
])
def test_reference_host_code_over_the_library_reproduces_the_oracle(have_bin, oracle, tmp_path, dim, geom, iters, fp, ndev, occ):
    """REFERENCE arithmetic (the shim's default): the bits of the reference's own CPU run.  Several partitions are placed on
    device 0 when the box has fewer GPUs (an oversubscribed device list, as the reference's own tests use)."""
    nx, ny, nz = dim
    dump = os.path.join(str(tmp_path), "d.bin")
    extra = [] if ndev <= have_bin else ["--same-gpu"]
    r = run(str(tmp_path), "--nx", nx, "--ny", ny, "--nz", nz, "--iters", iters, "--fp", fp, "--geom", geom, "--ndev", ndev, "--occ", occ,
            "--dump", dump, *extra)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = oracle.read_ref_dump(dump)
    dt = np.float32 if fp == "float" else np.float64
    cls = oracle.classify(1 if geom == "sphere" else 0, nx, ny, nz)
    mask = oracle.wall_mask(19, cls)
    assert np.array_equal(d["cls"], cls) and np.array_equal(d["mask"], mask)
    ref = oracle.run(19, oracle.init_pop(19, cls, dt), cls, mask, d["omega"], iters)
    assert np.array_equal(d["pop"].view(np.uint8), ref.view(np.uint8))


def test_reference_host_code_over_the_library_throughput(have_bin, tmp_path):
    """The reference benchmark's metric (Metrics.h:39-42) with its own Skeleton issuing the library's kernel; FAST arithmetic.
    Printed for the record (pytest -s) and kept loose: it must at least beat the reference's own CUDA backend on the same GPU
    (18.3 GLUPS at 256^3, profiles/r01m_reference_gpu_backend.log)."""
    r = run(str(tmp_path), "--n", 256, "--iters", 110, "--bench", 10, "--fp", "float", env={"NLBM_SHIM_ARITH": "fast"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = next(l for l in r.stdout.splitlines() if l.startswith("{") and "ref_bench" in l)
    j = json.loads(line)
    print("reference host code over libneon_lbm.so, 256^3 fp32:", j)
    assert j["mlups"] > 25000
