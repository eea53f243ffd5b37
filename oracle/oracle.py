"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/libvrt_oracle.so (vrt_oracle.c), the plain-C CPU
restatement of the reference hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvrt_oracle.so")

ROUND_DEVICE = 0
ROUND_HOST = 1


class _TraceArgs(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("volume_is_i16", C.c_int), ("dir_is_i16", C.c_int), ("round_mode", C.c_int),
        ("bounds", C.c_uint32 * 3), ("invscale", C.c_float * 3),
        ("iterations", C.c_uint32), ("min_brightness", C.c_uint32),
        ("volume", C.c_void_p), ("translucency", C.c_void_p), ("threads", C.c_int),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.vrt_oracle_interp_f32.restype = C.c_float
        _lib.vrt_oracle_interp_u32.restype = C.c_uint32
        _lib.vrt_oracle_interp_i32.restype = C.c_int32
        _lib.vrt_oracle_normalise_f32.restype = C.c_long
        _lib.vrt_oracle_normalise_u32.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def trace(volume, bounds, pos, dir, invscale, iterations, translucency=None, min_brightness=0,
          trace_path=False, round_mode=ROUND_DEVICE, threads=0):
    """vrt_oracle_trace.  volume: interleaved [nvox, dim+1] float32|int16; pos [n, dim] uint32; dir [n, dim] float32|int16.
    translucency=None reproduces the shipped (compiled-out) behaviour; an array makes the plane live."""
    bounds = [int(b) for b in bounds]
    dim = len(bounds)
    volume = np.ascontiguousarray(volume)
    assert volume.dtype in (np.float32, np.int16)
    assert volume.size == int(np.prod(bounds)) * (dim + 1), "volume size does not match bounds"
    pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, dim)
    dir = np.ascontiguousarray(dir).reshape(-1, dim)
    assert dir.dtype in (np.float32, np.int16)
    n = pos.shape[0]
    a = _TraceArgs()
    a.dim = dim
    a.volume_is_i16 = int(volume.dtype == np.int16)
    a.dir_is_i16 = int(dir.dtype == np.int16)
    a.round_mode = round_mode
    for d in range(dim):
        a.bounds[d] = bounds[d]
        a.invscale[d] = float(invscale[d])
    a.iterations = iterations
    a.min_brightness = min_brightness
    a.volume = volume.ctypes.data
    tr = None
    if translucency is not None:
        tr = np.ascontiguousarray(translucency, dtype=np.uint32).reshape(-1)
        assert tr.size == int(np.prod(bounds))
        a.translucency = tr.ctypes.data
    a.threads = threads
    epos = np.zeros_like(pos); edir = np.zeros_like(dir)
    eit = np.zeros(n, dtype=np.uint32); light = np.zeros(n, dtype=np.uint32)
    path = np.zeros((n, iterations, dim), dtype=np.uint32) if trace_path else None
    rc = lib().vrt_oracle_trace(C.byref(a), C.c_size_t(n), _p(pos), _p(dir), _p(epos), _p(edir), _p(eit), _p(light), _p(path))
    if rc != 0:
        raise ValueError("vrt_oracle_trace: bad arguments")
    return epos, edir, eit, light, path


def fold(diff_planes, translucency_cropped):
    """TraceRaysCu ctor: [d0..d(dim-1), extra] interleave (cu:654-669)."""
    planes = [np.ascontiguousarray(d).reshape(-1) for d in diff_planes]
    dim = len(planes)
    nvox = planes[0].size
    tr = np.ascontiguousarray(translucency_cropped, dtype=np.uint32).reshape(-1)
    out = np.zeros((nvox, dim + 1), dtype=planes[0].dtype)
    ptrs = (C.c_void_p * dim)(*[p.ctypes.data for p in planes])
    fn = lib().vrt_oracle_fold_f32 if planes[0].dtype == np.float32 else lib().vrt_oracle_fold_i16
    fn(dim, C.c_size_t(nvox), ptrs, _p(tr), _p(out))
    return out


def prep(bounds, ior, translucency):
    """Scene prep (f1): returns (diff_bounds, iorlog, [diff planes], translucency_cropped)."""
    bounds_a = np.asarray(bounds, dtype=np.uint64)
    dim = len(bounds_a)
    ior = np.ascontiguousarray(ior).reshape(-1)
    tr = np.ascontiguousarray(translucency, dtype=np.uint32).reshape(-1)
    isf = ior.dtype == np.float32
    if not isf:
        ior = ior.astype(np.uint32, copy=False)
    ob = [int(b) - 2 for b in bounds_a]
    nout = int(np.prod(ob))
    iorlog = np.zeros(ior.size, dtype=np.float32 if isf else np.int32)
    planes = [np.zeros(nout, dtype=np.float32 if isf else np.int16) for _ in range(dim)]
    trc = np.zeros(nout, dtype=np.uint32)
    ptrs = (C.c_void_p * dim)(*[p.ctypes.data for p in planes])
    fn = lib().vrt_oracle_prep_f32 if isf else lib().vrt_oracle_prep_u32
    rc = fn(dim, _p(bounds_a), _p(ior), _p(tr), _p(iorlog), ptrs, _p(trc))
    if rc == -2:
        raise RuntimeError("refraction-index underflow/overflow")
    if rc == -3:
        raise RuntimeError("differention overflow")
    if rc != 0:
        raise ValueError("bad arguments")
    return ob, iorlog, planes, trc


def normalise(bounds, ior, pos, dir):
    """Ray pre-processing (f2), returns new (pos, dir) in cropped coordinates with dir *= n(pos)."""
    bounds_a = np.asarray(bounds, dtype=np.uint64)
    dim = len(bounds_a)
    ior = np.ascontiguousarray(ior).reshape(-1)
    pos = np.array(pos, dtype=np.uint32).reshape(-1, dim).copy()
    n = pos.shape[0]
    if ior.dtype == np.float32:
        dir = np.array(dir, dtype=np.float32).reshape(-1, dim).copy()
        bad = lib().vrt_oracle_normalise_f32(dim, _p(bounds_a), _p(ior), C.c_size_t(n), _p(pos), _p(dir))
    else:
        ior = ior.astype(np.uint32, copy=False)
        dir = np.array(dir, dtype=np.int16).reshape(-1, dim).copy()
        ovf = C.c_long(0)
        bad = lib().vrt_oracle_normalise_u32(dim, _p(bounds_a), _p(ior), C.c_size_t(n), _p(pos), _p(dir), C.byref(ovf))
        if bad == 0 and ovf.value:
            raise RuntimeError("Normalize length failed (ray %d)" % (ovf.value - 1))
    if bad:
        raise RuntimeError("ray %d: is not in 0 to bounds" % (bad - 1))
    return pos, dir


def interp(img, bounds, pos):
    bounds_a = np.asarray(bounds, dtype=np.uint64)
    dim = len(bounds_a)
    img = np.ascontiguousarray(img).reshape(-1)
    pos = np.ascontiguousarray(pos, dtype=np.uint32).reshape(-1, dim)
    fn = {np.dtype(np.float32): lib().vrt_oracle_interp_f32, np.dtype(np.uint32): lib().vrt_oracle_interp_u32,
          np.dtype(np.int32): lib().vrt_oracle_interp_i32}[img.dtype]
    out = np.zeros(pos.shape[0], dtype=img.dtype)
    for i in range(pos.shape[0]):
        out[i] = fn(dim, _p(bounds_a), _p(img), _p(pos[i]))
    return out
