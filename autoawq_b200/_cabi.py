"""ctypes binding of libb200awq.so (include/b200awq.h).  No fallback: if the CUDA library is missing
or a call fails, this raises - the product path never routes through the CPU oracle."""
from __future__ import annotations

import ctypes
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200awq.so")

OK = 0
ABI_VERSION = 1

_c_void_p, _c_int, _c_i64, _c_size_t, _c_float = (
    ctypes.c_void_p,
    ctypes.c_int,
    ctypes.c_int64,
    ctypes.c_size_t,
    ctypes.c_float,
)



class Op(ctypes.Structure):
    """b200awq_op_t (include/b200awq.h): one recorded operator call of a decode program."""

    _fields_ = [
        ("kind", ctypes.c_int32), ("M", ctypes.c_int32), ("K", ctypes.c_int32), ("N", ctypes.c_int32),
        ("group_size", ctypes.c_int32), ("eps", ctypes.c_float), ("ldx", ctypes.c_int64),
        ("x", ctypes.c_void_p), ("qweight", ctypes.c_void_p), ("scales", ctypes.c_void_p),
        ("qzeros", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("weight", ctypes.c_void_p), ("y", ctypes.c_void_p),
    ]


OP_RMSNORM, OP_LINEAR_GEMM, OP_SILU_AND_MUL = 1, 2, 3
EUNSUPPORTED = 2

# name -> (restype, argtypes); mirrors include/b200awq.h one to one
SIGNATURES = {
    "b200awq_abi_version": (_c_int, []),
    "b200awq_error_string": (ctypes.c_char_p, [_c_int]),
    "b200awq_last_cuda_error": (ctypes.c_char_p, []),
    "b200awq_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "b200awq_dequantize_gemm": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "b200awq_gemm_forward": (
        _c_int,
        [_c_void_p, _c_i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
         _c_void_p, _c_size_t, _c_void_p],
    ),
    "b200awq_gemv_forward": (
        _c_int,
        [_c_void_p, _c_i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
         _c_void_p, _c_size_t, _c_void_p],
    ),
    "b200awq_fast_forward": (
        _c_int,
        [_c_void_p, _c_i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
         _c_void_p, _c_size_t, _c_void_p],
    ),
    "b200awq_rmsnorm": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_float, _c_void_p]),
    "b200awq_silu_and_mul": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "b200awq_set_knob": (_c_int, [_c_int, _c_int]),
    "b200awq_get_knob": (_c_int, [_c_int]),
    "b200awq_debug_read": (_c_int, [_c_void_p, _c_size_t]),
    "b200awq_tcq_plan": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "b200awq_topk_softmax": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "b200awq_moe_align_block_size": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                              _c_void_p]),
    "b200awq_grouped_gemm_forward": (
        _c_int,
        [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
         _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p],
    ),
    "b200awq_comm_create": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_void_p)]),
    "b200awq_comm_ipc_handle": (_c_int, [_c_void_p, _c_void_p]),
    "b200awq_comm_open": (_c_int, [_c_void_p, _c_void_p]),
    "b200awq_comm_all_reduce": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "b200awq_comm_error": (_c_int, [_c_void_p]),
    "b200awq_comm_destroy": (_c_int, [_c_void_p]),
    "b200awq_program_create": (_c_int, [ctypes.POINTER(Op), _c_int, ctypes.POINTER(_c_void_p)]),
    "b200awq_program_num_ops": (_c_int, [_c_void_p]),
    "b200awq_program_kind": (_c_int, [_c_void_p]),
    "b200awq_stream_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "b200awq_stream_pack": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "b200awq_program_run": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "b200awq_program_destroy": (_c_int, [_c_void_p]),
}


class B200AwqError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (the reference's kernels raise
    RuntimeError through TORCH_CHECK; same class hierarchy here)."""


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    v = lib.b200awq_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"libb200awq ABI {v} != expected {ABI_VERSION}")
    return lib


lib = _load()


def check(code: int, what: str) -> None:
    if code != OK:
        msg = lib.b200awq_error_string(code).decode()
        cu = lib.b200awq_last_cuda_error().decode()
        raise B200AwqError(f"{what}: {msg}" + (f" [{cu}]" if code == 4 and cu else ""))


def lib_path() -> str:
    return _LIB_PATH
