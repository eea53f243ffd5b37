"""ctypes binding of the C ABI in include/vrt_b200.h (libvrt_b200.so, built by csrc/Makefile).

The library is the product: if it is missing or fails to load this module raises -- there is no
Python or CPU fallback anywhere in the package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRT_B200_LIB") or os.path.join(_HERE, "libvrt_b200.so")   # override only for tuning experiments

VRT_OK, VRT_ERR_INVALID, VRT_ERR_CUDA, VRT_ERR_NOMEM, VRT_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
VRT_F32, VRT_I16, VRT_U32 = 0, 1, 2
VRT_TRACE_LIVE_TRANSLUCENCY, VRT_TRACE_PATHS, VRT_TRACE_ROUND_HOST = 1, 2, 4
VRT_SCENE_BORROW = 1
VRT_SCENE_LAYOUT_BRICK = 2
VRT_SCENE_KEEP_I16 = 4
VRT_SCENE_LAYOUT_TEXTURE = 8
VRT_SCENE_LAYOUT_PAIR = 16
VRT_OPT_KERNEL, VRT_OPT_BLOCK_THREADS, VRT_OPT_REFILL, VRT_OPT_CHUNK_RAYS, VRT_OPT_STEPS_PER_POLL, VRT_OPT_MAX_CTAS_PER_SM = 0, 1, 2, 3, 4, 5
VRT_OPT_REGION_LOG2, VRT_OPT_REGION_ROUNDS = 6, 7
VRT_OPT_WAVE_LOG2, VRT_OPT_WAVE_MARGIN, VRT_OPT_WAVE_CHECK, VRT_OPT_WAVE_TAIL_PERMILLE, VRT_OPT_WAVE_CTAS_PER_SM = 8, 9, 10, 11, 12
VRT_OPT_WAVE_REFILL = 14
VRT_OPT_ALL_CLEAR_KERNEL = 18
VRT_INFO_ALL_CLEAR = 103
VRT_INFO_WAVE_ROUNDS = 102
VRT_INFO_EMPTY_PERMILLE = 100
VRT_INFO_NUM_SMS = 101
VRT_INFO_STAT_BASE = 200

# every symbol include/vrt_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "vrt_last_error", "vrt_version", "vrt_device_count", "vrt_scene_create", "vrt_scene_create_interleaved",
    "vrt_scene_create_device", "vrt_scene_create_from_ior", "vrt_scene_destroy", "vrt_scene_info",
    "vrt_scene_download", "vrt_scene_export_device", "vrt_scene_set_option", "vrt_scene_get_option", "vrt_trace", "vrt_trace_device",
    "vrt_normalise_rays_device", "vrt_measure_gather_bandwidth", "vrt_selftest_division", "vrt_launch_count",
    "vrt_scene_storage_info", "vrt_scene_replicate", "vrt_comm_unique_id", "vrt_comm_create", "vrt_comm_destroy", "vrt_scene_broadcast",
    "vrt_trace_cap_hit",
]


class VrtError(RuntimeError):
    """Mirrors the reference's error convention: failures surface as std::runtime_error (cu:54-63), which
    pybind11 maps to RuntimeError."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `make -C volumeraytracer_b200/csrc` "
                              "(or python -c 'import __graft_entry__ as g; g.build()'); there is no fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.vrt_last_error.restype = C.c_char_p
        L.vrt_version.restype = C.c_char_p
        L.vrt_launch_count.restype = C.c_uint64
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.vrt_device_count.argtypes = [C.POINTER(i32)]
        L.vrt_scene_create.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp, C.c_uint]
        L.vrt_scene_create_interleaved.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp, C.c_uint]
        L.vrt_scene_create_device.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp, C.c_uint]
        L.vrt_scene_create_from_ior.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp, i32, C.c_uint]
        L.vrt_scene_destroy.argtypes = [vp]
        L.vrt_scene_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
        L.vrt_scene_download.argtypes = [vp, vp, vp]
        L.vrt_scene_export_device.argtypes = [vp, vp, vp, vp]
        L.vrt_scene_set_option.argtypes = [vp, i32, C.c_int64]
        L.vrt_scene_get_option.argtypes = [vp, i32, C.POINTER(C.c_int64)]
        L.vrt_trace.argtypes = [vp, u64, vp, vp, i32, vp, u32, u32, C.c_uint, vp, vp, vp, vp, vp]
        L.vrt_trace_device.argtypes = [vp, u64, vp, vp, i32, vp, u32, u32, C.c_uint, vp, vp, vp, vp, vp, vp]
        L.vrt_normalise_rays_device.argtypes = [vp, u64, vp, vp, i32, C.POINTER(C.c_int64), vp]
        L.vrt_measure_gather_bandwidth.argtypes = [i32, u64, i32, i32, C.POINTER(C.c_double)]
        L.vrt_selftest_division.argtypes = [i32, C.POINTER(C.c_uint64)]
        L.vrt_scene_storage_info.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_uint), C.POINTER(u64), C.POINTER(vp)]
        L.vrt_scene_replicate.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(vp), C.POINTER(C.c_double)]
        L.vrt_comm_unique_id.argtypes = [vp]
        L.vrt_comm_create.argtypes = [C.POINTER(vp), i32, i32, i32, vp]
        L.vrt_comm_destroy.argtypes = [vp]
        L.vrt_scene_broadcast.argtypes = [vp, i32, vp, C.POINTER(vp), C.POINTER(C.c_double)]
        L.vrt_trace_cap_hit.restype = i32
        _lib = L
    return _lib


def check(rc):
    if rc != VRT_OK:
        raise VrtError(rc, lib().vrt_last_error().decode(errors="replace"))


def launch_count():
    return int(lib().vrt_launch_count())
