"""Reader/writer for the reference's `debug_*_instance` dumps ("next" row f3), so that they can be replayed as fixtures.

Format (reference src/serialize.h:12-79, image_util.cpp:35-144): little-endian raw values; a vector is a uint64 count
followed by its elements; bool is one byte.
  scene instance   : bound_vec (u64[]), ior (u32[] 16.16 | f32[]), translucency (u32[])
  ray instance     : start_position (u32[]), start_direction (i16[] | f32[]), invscale (f32[]), minimum_brightness u32,
                     iterations u32, trace_path bool, normalize_length bool
  combined instance: the scene fields followed by the ray fields (python_binding.cpp:21-34 writes this one)
"""
import struct

import numpy as np


def _write_vec(f, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
    f.write(struct.pack("<Q", a.size))
    f.write(a.tobytes())


def _read_vec(f, dtype):
    (n,) = struct.unpack("<Q", _exact(f, 8))
    dt = np.dtype(dtype)
    return np.frombuffer(_exact(f, n * dt.itemsize), dtype=dt).copy()


def _exact(f, n):
    b = f.read(n)
    if len(b) != n:
        raise RuntimeError("Bad input Stream")        # serialize.h:16
    return b


def write_scene(f, bound_vec, ior, translucency, float_scene=False):
    _write_vec(f, bound_vec, np.uint64)
    _write_vec(f, ior, np.float32 if float_scene else np.uint32)
    _write_vec(f, translucency, np.uint32)


def read_scene(f, float_scene=False):
    return dict(bound_vec=_read_vec(f, np.uint64), ior=_read_vec(f, np.float32 if float_scene else np.uint32),
                translucency=_read_vec(f, np.uint32))


def write_rays(f, start_position, start_direction, invscale, minimum_brightness, iterations, trace_path, normalize_length=True,
               float_dirs=False):
    _write_vec(f, start_position, np.uint32)
    _write_vec(f, start_direction, np.float32 if float_dirs else np.int16)
    _write_vec(f, invscale, np.float32)
    f.write(struct.pack("<II??", int(minimum_brightness), int(iterations), bool(trace_path), bool(normalize_length)))


def read_rays(f, float_dirs=False):
    d = dict(start_position=_read_vec(f, np.uint32), start_direction=_read_vec(f, np.float32 if float_dirs else np.int16),
             invscale=_read_vec(f, np.float32))
    mb, it, tp, nl = struct.unpack("<II??", _exact(f, 10))
    d.update(minimum_brightness=mb, iterations=it, trace_path=tp, normalize_length=nl)
    return d


def write_instance(f, scene, rays, float_scene=False, float_dirs=False):
    write_scene(f, scene["bound_vec"], scene["ior"], scene["translucency"], float_scene)
    write_rays(f, rays["start_position"], rays["start_direction"], rays["invscale"], rays["minimum_brightness"], rays["iterations"],
               rays["trace_path"], rays.get("normalize_length", True), float_dirs)


def read_instance(f, float_scene=False, float_dirs=False):
    d = read_scene(f, float_scene)
    d.update(read_rays(f, float_dirs))
    return d
