"""ctypes binding of the C ABI declared in include/ddcb200.h (the only path from Python to the kernels)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDCB200_LIB", os.path.join(_HERE, "libddcb200.so"))

OK, EINVAL, ECUDA, ENOMEM, ETOOSHORT = 0, -1, -2, -3, -4

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "ddcb200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int]),
    "ddcb200_destroy": (None, [C.c_void_p]),
    "ddcb200_set_taps": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "ddcb200_set_decimation": (C.c_int, [C.c_void_p, C.c_int]),
    "ddcb200_out_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "ddcb200_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "ddcb200_tensor_engine_geometry": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "ddcb200_run_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "ddcb200_run_packed10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_void_p]),
    "ddcb200_unpack10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ddcb200_pack10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "ddcb200_run_short_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int64, C.c_void_p, C.c_void_p]),
    "ddcb200_mix_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ddcb200_fir_c64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ddcb200_decimate_c64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ddcb200_run_host_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int64,
                                       C.c_void_p, C.c_int64]),
    "ddcb200_run_host_f32_c128": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int64, C.c_void_p]),
    "ddcb200_run_host_packed10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                            C.c_int64, C.c_void_p, C.c_int64]),
    "ddcb200_cwg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double,
                              C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_void_p]),
    "ddcb200_session_open": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_void_p)]),
    "ddcb200_session_open_packed10": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_void_p)]),
    "ddcb200_session_push_packed10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                                C.POINTER(C.c_int64), C.c_void_p]),
    "ddcb200_session_push_host_packed10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                                     C.POINTER(C.c_int64)]),
    "ddcb200_session_close": (None, [C.c_void_p]),
    "ddcb200_session_reset": (C.c_int, [C.c_void_p, C.c_int64]),
    "ddcb200_session_pending": (C.c_int64, [C.c_void_p]),
    "ddcb200_session_position": (C.c_int64, [C.c_void_p]),
    "ddcb200_session_out_len": (C.c_int64, [C.c_void_p, C.c_int64]),
    "ddcb200_session_push_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                           C.POINTER(C.c_int64), C.c_void_p]),
    "ddcb200_session_push_host_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                                C.POINTER(C.c_int64)]),
    "ddcb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "ddcb200_host_free": (None, [C.c_void_p]),
    "ddcb200_sync": (C.c_int, [C.c_void_p]),
    "ddcb200_stream": (C.c_void_p, [C.c_void_p]),
    "ddcb200_last_error": (C.c_char_p, []),
    "ddcb200_version": (C.c_int, []),
    "ddcb200_launch_count": (C.c_int64, [C.c_void_p]),
    "ddcb200_last_variant": (C.c_char_p, [C.c_void_p]),
    "ddcb200_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
}

_lib = None


class DdcLibraryError(RuntimeError):
    """libddcb200.so is missing or failed; there is deliberately no CPU fallback."""


def load():
    """Load libddcb200.so once and attach the signatures. Raises DdcLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DdcLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dc_sand_b200/csrc`). dc_sand_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the machine
        raise DdcLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().ddcb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "ddcb200"):
    """Map a status code to the exception the reference-facing wrapper promises."""
    if rc == OK:
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc in (EINVAL, ETOOSHORT):
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
