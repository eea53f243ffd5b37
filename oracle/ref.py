"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/libvrt_ref.so -- the UNMODIFIED reference
CPU implementation (see ref_harness.cpp / ref_harness_scene.cpp / Makefile).  Importable only
where that library has been built (here: from /root/reference; on the GPU box: the prebuilt file
that travels with the snapshot).  Never imported by the product package.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvrt_ref.so")
# the same harness compiled by nvcc for sm_100: TraceRaysCu<> then runs the reference's own CUDA kernel
# (trace_rays_gpu, cu:397-414) for more than 0x80 rays -- the GPU-side comparator on the B200 box
CUDA_LIB_PATH = os.path.join(_HERE, "_ref", "libvrt_ref_cuda.so")

# the reference's UNMODIFIED image_util/util/io_util/serialize objects linked against OUR TraceRaysCu<> drop-in
# (volumeraytracer_b200/csrc/trace_rays_cu_dropin.cpp -> libvrt_b200.so): RefScene(..., which="dropin") then runs the
# reference's C++ API with the B200 marcher underneath
DROPIN_LIB_PATH = os.path.join(_HERE, "_ref", "libvrt_dropin.so")

_libs = {}


def _path(which):
    return {False: LIB_PATH, "cpu": LIB_PATH, True: CUDA_LIB_PATH, "cuda": CUDA_LIB_PATH, "dropin": DROPIN_LIB_PATH}[which]


def _host_has_avx512():
    """oracle/_ref is built with -march=x86-64-v4 (the reference needs AVX-512BW/VL, cu:176): do not load it elsewhere."""
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f in flags for f in ("avx512f", "avx512bw", "avx512vl", "avx2", "fma"))
    except OSError:
        return False


def available(cuda=False):
    return os.path.exists(_path(cuda)) and _host_has_avx512()


def lib(cuda=False):
    if cuda not in _libs:
        if not available(cuda):
            raise RuntimeError("reference harness not built: run `make -C oracle ref ref_cuda` where /root/reference exists")
        l = C.CDLL(_path(cuda))
        l.vrtref_last_error.restype = C.c_char_p
        _libs[cuda] = l
    return _libs[cuda]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc, cuda=False):
    if rc != 0:
        raise RuntimeError(lib(cuda).vrtref_last_error().decode())


def omp_max_threads():
    return lib().vrtref_omp_max_threads()


_DIR = {np.dtype(np.float32): "f32", np.dtype(np.int16): "i16"}


class RefScene:
    """Reference RaytraceScene<float,float,float> (ior float32) or <ior_t,iorlog_t,diff_t> (ior uint32).
    image_util.cpp:501-643 (ctor), :645-772 (trace_rays)."""

    def __init__(self, bounds, ior, translucency, loglevel=0, which=False):
        self.which = which
        self.bounds = np.asarray(bounds, dtype=np.uint64)
        self.dim = len(self.bounds)
        ior = np.ascontiguousarray(ior).reshape(-1)
        tr = np.ascontiguousarray(translucency, dtype=np.uint32).reshape(-1)
        self.kind = "f32" if ior.dtype == np.float32 else "u32"
        if self.kind == "u32":
            ior = ior.astype(np.uint32, copy=False)
        self.h = C.c_void_p()
        fn = getattr(lib(self.which), "vrtref_scene_new_" + self.kind)
        _check(fn(C.byref(self.h), _p(self.bounds), self.dim, _p(ior), _p(tr), loglevel), self.which)
        db = np.zeros(self.dim, dtype=np.uint64)
        getattr(lib(self.which), "vrtref_scene_diff_bounds_" + self.kind)(self.h, _p(db))
        self.diff_bounds = db
        self.diff_dtype = np.float32 if self.kind == "f32" else np.int16
        self.dir_dtype = np.float32 if self.kind == "f32" else np.int16

    def close(self):
        if self.h:
            getattr(lib(self.which), "vrtref_scene_delete_" + self.kind)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nvox(self):
        return int(np.prod(self.diff_bounds))

    def interleaved(self):
        out = np.zeros((self.nvox, self.dim + 1), dtype=self.diff_dtype)
        getattr(lib(self.which), "vrtref_scene_interleaved_" + self.kind)(self.h, _p(out))
        return out

    def diff(self, axis):
        out = np.zeros(self.nvox, dtype=self.diff_dtype)
        getattr(lib(self.which), "vrtref_scene_diff_" + self.kind)(self.h, axis, _p(out))
        return out

    def iorlog(self):
        out = np.zeros(int(np.prod(self.bounds)), dtype=np.float32 if self.kind == "f32" else np.int32)
        getattr(lib(self.which), "vrtref_scene_iorlog_" + self.kind)(self.h, _p(out))
        return out

    def translucency_cropped(self):
        out = np.zeros(self.nvox, dtype=np.uint32)
        getattr(lib(self.which), "vrtref_scene_translucency_cropped_" + self.kind)(self.h, _p(out))
        return out

    def trace(self, pos, dir, invscale, min_brightness, iterations, trace_path=False, max_cpu=0):
        pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, self.dim)
        dir = np.ascontiguousarray(dir, dtype=self.dir_dtype).reshape(-1, self.dim)
        n = pos.shape[0]
        isc = np.ascontiguousarray(invscale, dtype=np.float32)
        epos = np.zeros_like(pos); edir = np.zeros_like(dir)
        eit = np.zeros(n, dtype=np.uint32); light = np.zeros(n, dtype=np.uint32)
        path = np.zeros((n, iterations, self.dim), dtype=np.uint32) if trace_path else None
        fn = getattr(lib(self.which), "vrtref_scene_trace_" + self.kind)
        _check(fn(self.h, C.c_size_t(n), _p(pos), _p(dir), _p(isc), C.c_uint32(min_brightness), C.c_uint32(iterations),
                  int(trace_path), int(max_cpu), _p(epos), _p(edir), _p(eit), _p(light), _p(path)), self.which)
        return epos, edir, eit, light, path


class RefTracer:
    """Reference TraceRaysCu<float|diff_t> (cuda_volume_raytracer.h:61-115) on planar gradient arrays."""

    def __init__(self, bounds, diff_planes, translucency_cropped, cuda=False):
        self.cuda = cuda
        self.bounds = np.asarray(bounds, dtype=np.uint64)
        self.dim = len(self.bounds)
        self.planes = [np.ascontiguousarray(d).reshape(-1) for d in diff_planes]
        self.kind = _DIR[self.planes[0].dtype]
        tr = np.ascontiguousarray(translucency_cropped, dtype=np.uint32).reshape(-1)
        ptrs = (C.c_void_p * self.dim)(*[p.ctypes.data for p in self.planes])
        self.h = C.c_void_p()
        _check(getattr(lib(cuda), "vrtref_tracer_new_" + self.kind)(C.byref(self.h), _p(self.bounds), self.dim, ptrs, _p(tr)), cuda)

    def close(self):
        if self.h:
            getattr(lib(self.cuda), "vrtref_tracer_delete_" + self.kind)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def interleaved(self):
        out = np.zeros((int(np.prod(self.bounds)), self.dim + 1), dtype=self.planes[0].dtype)
        getattr(lib(), "vrtref_tracer_interleaved_" + self.kind)(self.h, _p(out))
        return out

    def trace(self, pos, dir, invscale, min_brightness, iterations, trace_path=False, max_cpu=0):
        pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, self.dim)
        dir = np.ascontiguousarray(dir).reshape(-1, self.dim)
        dk = _DIR[dir.dtype]
        n = pos.shape[0]
        isc = np.ascontiguousarray(invscale, dtype=np.float32)
        epos = np.zeros_like(pos); edir = np.zeros_like(dir)
        eit = np.zeros(n, dtype=np.uint32); light = np.zeros(n, dtype=np.uint32)
        path = np.zeros((n, iterations, self.dim), dtype=np.uint32) if trace_path else None
        fn = getattr(lib(self.cuda), "vrtref_tracer_trace_%s_%s" % (self.kind, dk))
        _check(fn(self.h, C.c_size_t(n), _p(pos), _p(dir), _p(isc), C.c_uint32(min_brightness), C.c_uint32(iterations),
                  int(trace_path), int(max_cpu), _p(epos), _p(edir), _p(eit), _p(light), _p(path)), self.cuda)
        return epos, edir, eit, light, path


def trace_live(volume, translucency, bounds, invscale, pos, dir, iterations, min_brightness, trace_path=False, threads=0):
    """Reference trace_rays_cpu (cu:376-394) instantiated with a LIVE translucency plane and brightness_t
    (cu:337-341); `volume` is the interleaved [d..,extra] array of the scene's DiffType."""
    bounds = np.asarray(bounds, dtype=np.uint64)
    dim = len(bounds)
    volume = np.ascontiguousarray(volume)
    vk = _DIR[volume.dtype]
    tr = np.ascontiguousarray(translucency, dtype=np.uint32).reshape(-1) if translucency is not None else None   # None: shipped (compiled-out) variant
    pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, dim)
    dir = np.ascontiguousarray(dir).reshape(-1, dim)
    dk = _DIR[dir.dtype]
    n = pos.shape[0]
    isc = np.ascontiguousarray(invscale, dtype=np.float32)
    epos = np.zeros_like(pos); edir = np.zeros_like(dir)
    eit = np.zeros(n, dtype=np.uint32); light = np.zeros(n, dtype=np.uint32)
    path = np.zeros((n, iterations, dim), dtype=np.uint32) if trace_path else None
    if threads <= 0:
        threads = omp_max_threads()
    fn = getattr(lib(), "vrtref_trace_live_%s_%s" % (vk, dk))
    _check(fn(_p(volume), _p(tr), _p(bounds), dim, _p(isc), C.c_size_t(n), _p(pos), _p(dir), C.c_uint32(iterations),
              C.c_uint32(min_brightness), int(trace_path), int(threads), _p(epos), _p(edir), _p(eit), _p(light), _p(path)))
    return epos, edir, eit, light, path


def interpolate(img, bounds, pos):
    """Reference host interpolator<T> (image_util.h:348-431)."""
    bounds = np.asarray(bounds, dtype=np.uint64)
    dim = len(bounds)
    img = np.ascontiguousarray(img).reshape(-1)
    pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, dim)
    kind = {np.dtype(np.float32): "f32", np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32"}[img.dtype]
    out = np.zeros(pos.shape[0], dtype=img.dtype)
    _check(getattr(lib(), "vrtref_interpolate_" + kind)(_p(img), _p(bounds), dim, _p(pos), C.c_size_t(pos.shape[0]), _p(out)))
    return out
