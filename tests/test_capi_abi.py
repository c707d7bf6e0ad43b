"""The C-ABI library loads on a machine without a GPU and exports every symbol include/*.h declares; host-only entry
points work; compute entry points fail loudly (no CPU fallback).  No GPU."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from neon_b200 import build as B
    B.build()
    from neon_b200 import _capi
    return _capi


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(nlbm_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_every_declared_symbol_is_exported(capi):
    lib = C.CDLL(capi.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 19
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the Python binding knows every one of them
    assert set(decl) == set(capi.exported_symbols())


def test_abi_version_and_layout(capi):
    assert capi.lib().nlbm_abi_version() == 3
    d = capi.DenseDesc()
    d.nx, d.ny, d.nz_local, d.z_halo = 100, 7, 5, 1
    pb, fb = C.c_size_t(), C.c_size_t()
    capi.call("nlbm_dense_layout", C.byref(d), 19, 4, C.byref(pb), C.byref(fb))
    assert d.pitch_y == 128 and d.pitch_z == 128 * 7 and d.pitch_q == 128 * 7 * 7   # 512-byte rows
    assert pb.value == 19 * d.pitch_q * 4
    assert fb.value >= d.pitch_q * 4
    capi.call("nlbm_dense_layout", C.byref(d), 27, 8, C.byref(pb), C.byref(fb))
    assert d.pitch_y == 128 and pb.value == 27 * d.pitch_q * 8
    d.nx = 0
    assert capi.lib().nlbm_dense_layout(C.byref(d), 19, 4, None, None) == capi.ERR_INVALID
    assert "size" in capi.last_error()


def test_invalid_arguments_are_reported_not_thrown(capi):
    lib = capi.lib()
    assert lib.nlbm_d3q19_f32_dense_step(None, 1.0, 0, 0, None) == capi.ERR_INVALID
    with pytest.raises(capi.NeonException):
        capi.call("nlbm_dense_wall_mask", None, 19, None, None)


def test_no_cpu_fallback(capi):
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.lib().nlbm_device_count() <= 0
    import neon_b200 as nb
    with pytest.raises(RuntimeError):
        nb.Backend()  # Runtime.stream needs CUDA
    import numpy as np
    bk = nb.Backend(runtime=nb.Runtime.openmp)
    grid = nb.dGrid(bk, (16, 8, 8))
    a, b = grid.newField("a", 19, np.float32), grid.newField("b", 19, np.float32)
    flag = grid.newFlagField()
    c = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, a, b, flag, 1.0)
    with pytest.raises(nb.NeonException) as e:
        c.run(0, nb.DataView.STANDARD)
    assert e.value.status == capi.ERR_CUDA


def test_product_does_not_import_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "neon_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "oracle" not in src.replace("# oracle", ""), f"{path} refers to the oracle"
