"""ctypes binding of libtiktak_cuda.so -- every symbol include/tiktak_cuda.h declares.

The library is built in-tree (tiktak_b200/csrc/Makefile -> tiktak_b200/libtiktak_cuda.so) by
__graft_entry__.build().  There is no Python/CPU fallback: if the shared object is missing or no
CUDA device is present the calls fail loudly (TikTakError).
"""
import ctypes as C
import os

from ._abi import TiktakConfig, c_double_p, c_int32_p, c_int64_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TIKTAK_CUDA_LIB") or os.path.join(_HERE, "libtiktak_cuda.so")  # override: experiment builds

#: name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "tiktak_cuda_config_defaults": (C.c_int, [C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_double_p, c_int32_p, c_int32_p]),
    "tiktak_cuda_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(TiktakConfig)]),
    "tiktak_cuda_destroy": (None, [C.c_void_p]),
    "tiktak_cuda_sobol_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "tiktak_cuda_objfun": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tiktak_cuda_pretest": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "tiktak_cuda_pretest_result": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "tiktak_cuda_pretest_values_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_int64_p]),
    "tiktak_cuda_pretest_values_shard_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_int64_p]),
    "tiktak_cuda_pretest_values_wait": (C.c_int, [C.c_void_p]),
    "tiktak_cuda_pretest_values": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "tiktak_cuda_pretest_wave": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "tiktak_cuda_pretest_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "tiktak_cuda_pretest_find": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_int64_p]),
    "tiktak_cuda_pretest_set_cut": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "tiktak_cuda_choose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tiktak_cuda_choose_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tiktak_cuda_choose_candidates": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tiktak_cuda_local": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tiktak_cuda_local_enqueue": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tiktak_cuda_local_ingest": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "tiktak_cuda_local_fetch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tiktak_cuda_local_missing": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "tiktak_cuda_local_status": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_nrescue": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_async_supported": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_async_clock": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_async_info": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int32_p, c_int32_p]),
    "tiktak_cuda_local_ingest_rounds": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, c_int32_p]),
    "tiktak_cuda_local_capacity": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p]),
    "tiktak_cuda_local_launch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_int32_p]),
    "tiktak_cuda_comm_unique_id": (C.c_int, [C.c_void_p, C.c_int32]),
    "tiktak_cuda_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "tiktak_cuda_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32]),
    "tiktak_cuda_push_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "tiktak_cuda_best": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int32_p]),
    "tiktak_cuda_last_search": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int32_p]),
    "tiktak_cuda_device_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), c_int64_p]),
    "tiktak_cuda_launch_count": (C.c_int64, [C.c_void_p]),
    "tiktak_cuda_stage_ms": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p]),
    "tiktak_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "tiktak_cuda_strerror": (C.c_char_p, [C.c_int]),
    "tiktak_cuda_last_error": (C.c_char_p, []),
    "tiktak_cuda_version": (C.c_char_p, []),
    "tiktak_cuda_fp64_peak": (C.c_int, [C.c_int, c_double_p, c_double_p]),
}


class TikTakError(RuntimeError):
    """A negative status from libtiktak_cuda (the reference would have called exitState / stop 0)."""

    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: status {code}: {detail}")


_lib = None


def load():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TikTakError(-4, "load", f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, where):
    if code != 0:
        lib = load()
        msg = lib.tiktak_cuda_strerror(code).decode()
        det = lib.tiktak_cuda_last_error().decode()
        raise TikTakError(code, where, f"{msg}; {det}")
