"""ctypes binding of include/reve_cuda.h (exactly the symbols the header declares)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REVE_LIB") or os.path.join(_HERE, "libreve_cuda.so")   # REVE_LIB: A/B runs against another build


class reve_profile(C.Structure):
    _fields_ = [("launches_conv0", C.c_uint64), ("launches_body", C.c_uint64),
                ("launches_tail", C.c_uint64), ("ms_conv0", C.c_double), ("ms_body", C.c_double),
                ("ms_tail", C.c_double), ("timed_body", C.c_uint64), ("timed_frames", C.c_uint64),
                ("frames", C.c_uint64), ("body_frames", C.c_uint64), ("launches_yuv", C.c_uint64),
                ("body_layer_frames", C.c_uint64)]


class reve_ctx_options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("layers_per_launch", C.c_int),
                ("max_batch", C.c_int), ("debug_flags", C.c_uint32), ("debug_grid", C.c_int), ("trace", C.c_int),
                ("trace_launch", C.c_int), ("trace_chain", C.c_int)]


# name -> (restype, argtypes); mirrors include/reve_cuda.h one to one
SIGNATURES = {
    "reve_version": (C.c_int, []),
    "reve_strerror": (C.c_char_p, [C.c_int]),
    "reve_last_error": (C.c_char_p, [C.c_void_p]),
    "reve_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "reve_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    "reve_model_load_ncnn": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "reve_model_random": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "reve_model_from_arrays": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_void_p)]),
    "reve_model_save_ncnn": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]),
    "reve_model_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "reve_model_free": (None, [C.c_void_p]),
    "reve_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "reve_ctx_create_ex": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(reve_ctx_options), C.POINTER(C.c_void_p)]),
    "reve_ctx_destroy": (None, [C.c_void_p]),
    "reve_ctx_launch_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "reve_device_recover": (C.c_int, [C.c_int]),
    "reve_ctx_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 5),
    "reve_ctx_set_output_format": (C.c_int, [C.c_void_p, C.c_int]),
    "reve_ctx_output_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "reve_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "reve_host_free": (None, [C.c_void_p]),
    "reve_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint64]),
    "reve_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "reve_sync": (C.c_int, [C.c_void_p]),
    "reve_upscale_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "reve_ctx_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "reve_ctx_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "reve_ctx_get_profile": (C.c_int, [C.c_void_p, C.POINTER(reve_profile), C.c_int]),
    "reve_debug_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "reve_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "reve_debug_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float)]),
    "reve_debug_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float)]),
    "reve_geometry": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int)] * 2 + [C.c_void_p] * 4 + [C.c_size_t]),
    "reve_launch_plan": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int)] * 4),
}

_lib = None


def load() -> C.CDLL:
    """Loads libreve_cuda.so from the package directory.  No fallback: a missing library is an
    error (build it with ``python -c 'import __graft_entry__ as g; g.build()'`` or
    ``make -C reve_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first "
                              "(make -C reve_b200/csrc); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
