"""ctypes binding of libjds.so (include/jds.h) - the only door between the Python
mirror of the reference API and the CUDA implementation.

There is deliberately no fallback: if the library is missing or fails to load the
import raises, and every compute call needs a CUDA device.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JDS_LIB", os.path.join(HERE, "libjds.so"))

JDS_ABI_VERSION = 1
JDS_OK, JDS_ERR_INVALID, JDS_ERR_UNSUPPORTED, JDS_ERR_CUDA, JDS_ERR_NOMEM = 0, -1, -2, -3, -4
JDS_ERR_CAPACITY = -5
JDS_SUB_444, JDS_SUB_422, JDS_SUB_420 = 0, 1, 2
JDS_EXACT, JDS_FAST = 0, 1
JDS_HOST, JDS_DEVICE = 0, 1
JDS_OUT_RECON, JDS_OUT_COEFFS, JDS_OUT_ERR_Y, JDS_OUT_ERR_RGB = 1, 2, 4, 8
JDS_OUT_HIST, JDS_OUT_SSIM, JDS_OUT_PSNR = 16, 32, 64
JDS_VALUE_HIST_BINS = 2048
JDS_RECORD_FIELDS, JDS_SWEEP_RECORDS_MAX = 13, 128

SUBSAMPLING = {"4:4:4": JDS_SUB_444, "4:2:2": JDS_SUB_422, "4:2:0": JDS_SUB_420}
PRECISION = {"exact": JDS_EXACT, "fast": JDS_FAST}


class JdsParams(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("quality", C.c_int32),
                ("subsampling", C.c_int32), ("prefilter", C.c_int32),
                ("precision", C.c_int32), ("outputs", C.c_uint32), ("reserved", C.c_int32)]


class JdsMetrics(C.Structure):
    _fields_ = [("sse_rgb", C.c_uint64), ("sse_y", C.c_double), ("ssim_sum", C.c_double * 4),
                ("ssim_count", C.c_uint64), ("coeff_bits", C.c_uint64), ("nnz", C.c_uint64),
                ("total_coeffs", C.c_uint64), ("luma_blocks", C.c_uint64),
                ("hist50", C.c_int64 * 50), ("gpu_ms", C.c_double), ("reserved", C.c_uint64 * 3)]


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libjds error {code}: {message}")
        self.code = code
        self.message = message


#: name -> (restype, argtypes); exactly the entry points include/jds.h declares
PROTOTYPES = {
    "jds_abi_version": (C.c_int, []),
    "jds_last_error": (C.c_char_p, []),
    "jds_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "jds_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "jds_ctx_destroy": (C.c_int, [C.c_void_p]),
    "jds_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jds_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "jds_ctx_wait_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jds_ctx_record_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jds_ctx_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "jds_ctx_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "jds_ctx_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "jds_quant_table": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "jds_coeff_count": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "jds_plane_dims": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jds_roundtrip": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.POINTER(JdsMetrics)]),
    "jds_roundtrip_batch": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.POINTER(JdsMetrics)]),
    "jds_sweep": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.POINTER(C.c_int32), C.c_int,
                            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(JdsMetrics)]),
    "jds_sweep_records": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.POINTER(C.c_int32), C.c_int,
                                    C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "jds_aliasing_demo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(JdsMetrics), C.POINTER(JdsMetrics)]),
    "jds_aliasing_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(JdsMetrics), C.POINTER(JdsMetrics)]),
    "jds_entropy_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_uint64)]),
    "jds_roundtrip_batch_begin": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_int, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_int, C.POINTER(JdsMetrics)]),
    "jds_ctx_finish": (C.c_int, [C.c_void_p]),
    "jds_roundtrip_band": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(JdsMetrics)]),
    "jds_entropy_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64)]),
    "jds_jfif_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_double), C.c_void_p, C.c_uint64,
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "jds_roundtrip_batch_records": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "jds_plot_payload": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.POINTER(JdsMetrics)]),
    "jds_preview_size": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]),
    "jds_resize_area": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int]),
    "jds_color_convert": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "jds_subsample_plane": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "jds_upsample_plane": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_int]),
    "jds_compare_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(JdsMetrics)]),
    "jds_bitrate_partials": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "jds_block_op": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "jds_selected_block": (C.c_int, [C.c_void_p, C.POINTER(JdsParams), C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
}

_lib = None


def load():
    """Load libjds.so (once).  Raises if it is missing - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python jpeg_dsp_studio_b200/build.py` "
            "(nvcc, sm_100a). jpeg_dsp_studio_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)          # ctypes releases the GIL around every call
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.jds_abi_version() != JDS_ABI_VERSION:
        raise ImportError(f"libjds.so ABI {lib.jds_abi_version()} != binding {JDS_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != JDS_OK:
        msg = load().jds_last_error()
        raise NativeError(rc, msg.decode() if msg else "")
