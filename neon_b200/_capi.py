"""ctypes binding of the C ABI in include/neon_lbm.h (libneon_lbm.so, hand-written sm_100a kernels).

This is the ONLY compute path of the package: there is no CPU or PyTorch fallback.  If the library has not been
built (``python -m neon_b200.build``) importing the symbols fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libneon_lbm.so")

OK, ERR_INVALID, ERR_CUDA, ERR_GEOMETRY, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
VIEW_STANDARD, VIEW_INTERNAL, VIEW_BOUNDARY = 0, 1, 2
BOUNCE_BACK, MOVING_WALL, BULK, UNDEFINED = 0, 1, 2, 3
FLAG_MASK_BITS = 0x07FFFFFF
FLAG_CLASS_SHIFT = 28
ARITH_REFERENCE, ARITH_FAST = 0, 1
GEOM_CAVITY, GEOM_CAVITY_SPHERE, GEOM_FLOW_SPHERE = 0, 1, 2


def opt_vec(v: int) -> int:
    return (v & 0xF) << 4


def opt_rows_log2(r: int) -> int:
    return (r & 0xF) << 8


KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA = 0, 1, 2


def opt_kernel(k: int) -> int:
    return (k & 0xF) << 12


def opt_tma(l2promo: int = 0, groups: int = 0) -> int:
    return ((l2promo & 3) << 16) | ((groups & 3) << 18)


# direct-kernel variants (never change results; include/neon_lbm.h bits 20, 27, 28, 29)
OPT_FLAGS_SUMMARY_FIRST = 1 << 20
OPT_NO_XFACE_PREFETCH = 1 << 27
OPT_FLAG_WORDS = 1 << 28
OPT_NO_XFACE_FIXUP_PREFETCH = 1 << 29
OPT_REF_LITERAL = 1 << 30  # REFERENCE arithmetic: operand-for-operand transcription instead of the conversion-lean evaluation


class NeonException(RuntimeError):
    """Counterpart of Neon::NeonException (libNeonCore/include/Neon/core/types/Exceptions.h:19-24): every non-zero
    status of the C layer is converted into this, as the reference does for every failed CUDA call."""

    def __init__(self, where: str, status: int, text: str):
        super().__init__(f"[{where}] status {status}: {text}")
        self.status = status


class DenseDesc(C.Structure):
    """struct nlbm_dense_desc"""
    _fields_ = [
        ("pop_in", C.c_void_p), ("pop_out", C.c_void_p), ("flags", C.c_void_p),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz_local", C.c_int32), ("z_halo", C.c_int32),
        ("pitch_y", C.c_int64), ("pitch_z", C.c_int64), ("pitch_q", C.c_int64),
        ("z_origin", C.c_int32), ("gnx", C.c_int32), ("gny", C.c_int32), ("gnz", C.c_int32),
        ("wall_cache", C.c_void_p),
    ]

    def clone(self) -> "DenseDesc":
        d = DenseDesc()
        C.memmove(C.byref(d), C.byref(self), C.sizeof(DenseDesc))
        return d


class BlockDesc(C.Structure):
    """struct nlbm_block_desc"""
    _fields_ = [
        ("pop_in", C.c_void_p), ("pop_out", C.c_void_p), ("flags", C.c_void_p), ("info", C.c_void_p),
        ("n_blocks", C.c_uint32), ("n_blocks_alloc", C.c_uint32), ("n_down", C.c_uint32), ("n_up", C.c_uint32),
        ("gnx", C.c_int32), ("gny", C.c_int32), ("gnz", C.c_int32),
    ]

    def clone(self) -> "BlockDesc":
        d = BlockDesc()
        C.memmove(C.byref(d), C.byref(self), C.sizeof(BlockDesc))
        return d


class PeerDesc(C.Structure):
    """struct nlbm_peer_desc"""
    _fields_ = [
        ("down_field", C.c_void_p), ("up_field", C.c_void_p), ("down_nz_local", C.c_int32), ("up_nz_local", C.c_int32),
        ("down_flag", C.c_void_p), ("up_flag", C.c_void_p), ("counters", C.c_void_p), ("value", C.c_uint32),
    ]


_P = C.c_void_p
_D = C.POINTER(DenseDesc)
_B = C.POINTER(BlockDesc)
_SIGNATURES = {
    "nlbm_abi_version": (C.c_int, []),
    "nlbm_last_error": (C.c_char_p, []),
    "nlbm_device_count": (C.c_int, []),
    "nlbm_dense_layout": (C.c_int, [_D, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "nlbm_dense_wall_cache_layout": (C.c_int, [_D, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "nlbm_dense_wall_cache_build": (C.c_int, [_D, C.c_int, C.c_int, _P]),
    "nlbm_dense_classify": (C.c_int, [_D, C.c_int, C.POINTER(C.c_double), _P]),
    "nlbm_dense_flags_commit": (C.c_int, [_D, _P]),
    "nlbm_selftest_exact": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "nlbm_dense_flags_from_classes": (C.c_int, [_D, _P, C.c_int, C.c_int, _P]),
    "nlbm_dense_wall_mask": (C.c_int, [_D, C.c_int, _P, _P]),
    "nlbm_dense_init_pop_f32": (C.c_int, [_D, C.c_int, C.c_double, _P]),
    "nlbm_dense_init_pop_f64": (C.c_int, [_D, C.c_int, C.c_double, _P]),
    "nlbm_d3q19_f32_dense_step": (C.c_int, [_D, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q19_f64_dense_step": (C.c_int, [_D, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q19_f32c64_dense_step": (C.c_int, [_D, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q27_f32_dense_step": (C.c_int, [_D, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q27_f64_dense_step": (C.c_int, [_D, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_dense_step_n": (C.c_int, [C.c_int, _D, _P, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_dense_step_push": (C.c_int, [C.c_int, _D, C.POINTER(PeerDesc), C.c_double, C.c_int, _P]),
    "nlbm_d3q19_f32_dense_rho_u": (C.c_int, [_D, _P, _P, _P]),
    "nlbm_d3q19_f64_dense_rho_u": (C.c_int, [_D, _P, _P, _P]),
    "nlbm_dense_halo_push": (C.c_int, [_D, _P, _D, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "nlbm_dense_halo_pack": (C.c_int, [_D, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_size_t), _P]),
    "nlbm_dense_halo_unpack": (C.c_int, [_D, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nlbm_block_classify": (C.c_int, [_B, C.c_int, C.POINTER(C.c_double), _P, _P]),
    "nlbm_block_wall_mask": (C.c_int, [_B, C.c_int, _P, _P]),
    "nlbm_block_init_pop_f32": (C.c_int, [_B, C.c_int, C.c_double, _P]),
    "nlbm_block_init_pop_f64": (C.c_int, [_B, C.c_int, C.c_double, _P]),
    "nlbm_d3q19_f32_block_step": (C.c_int, [_B, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q19_f64_block_step": (C.c_int, [_B, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q27_f32_block_step": (C.c_int, [_B, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_d3q27_f64_block_step": (C.c_int, [_B, C.c_double, C.c_int, C.c_int, _P]),
    "nlbm_block_halo_push": (C.c_int, [_B, _P, _B, _P, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "nlbm_ipc_export": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_uint64)]),
    "nlbm_ipc_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "nlbm_ipc_close": (C.c_int, [_P]),
    "nlbm_enable_peer_access": (C.c_int, [C.c_int]),
    "nlbm_flag_signal": (C.c_int, [_P, C.c_uint32, _P]),
    "nlbm_flag_wait": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    "nlbm_flag_wait2": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "nlbm_dense_halo_push2": (C.c_int, [_D, _P, _P, C.c_int32, _P, _P, C.c_int32, _P, _P, C.c_uint32, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the sm_100a kernels with `python -m neon_b200.build` "
                "(neon_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def last_error() -> str:
    return lib().nlbm_last_error().decode(errors="replace")


def check(status: int, where: str) -> None:
    if status != OK:
        raise NeonException(where, status, last_error())


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)
