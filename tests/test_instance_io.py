"""f3: the reference's debug-instance format.  Python round trip everywhere; where the reference build exists, its own
writer/reader is the counterpart (what it writes we read; what we write it reads and traces)."""
import ctypes as C
import io as _io

import numpy as np

from tests import scenes as S
from volumeraytracer_b200 import instance_io as iio


def _example():
    shape = (14, 12, 16)
    ior, tr = S.random_scene(shape, seed=3, kind="u32")
    pos, d = S.random_rays(shape, 300, seed=1, dir_kind="i16")
    scene = dict(bound_vec=np.array(shape, np.uint64), ior=ior.reshape(-1), translucency=tr.reshape(-1))
    rays = dict(start_position=pos.reshape(-1), start_direction=d.reshape(-1), invscale=np.array([1, 1, 1], np.float32),
                minimum_brightness=7, iterations=200, trace_path=False, normalize_length=True)
    return shape, scene, rays


def test_python_round_trip():
    shape, scene, rays = _example()
    buf = _io.BytesIO()
    iio.write_instance(buf, scene, rays)
    buf.seek(0)
    back = iio.read_instance(buf)
    for k in ("bound_vec", "ior", "translucency", "start_position", "start_direction", "invscale"):
        assert np.array_equal(back[k], np.asarray({**scene, **rays}[k]).reshape(-1))
    assert (back["minimum_brightness"], back["iterations"], back["trace_path"], back["normalize_length"]) == (7, 200, False, True)
    assert buf.read() == b""


def test_against_the_reference_serializer(refimpl, oracle, tmp_path):
    shape, scene, rays = _example()
    lib = refimpl.lib()
    # (i) the reference writes a scene instance, we read it
    p1 = str(tmp_path / "debug_scene_instance")
    b = np.array(shape, np.uint64)
    assert lib.vrtref_write_scene_instance_u32(p1.encode(), b.ctypes.data_as(C.c_void_p), 3, scene["ior"].ctypes.data_as(C.c_void_p),
                                               scene["translucency"].ctypes.data_as(C.c_void_p)) == 0
    got = iio.read_scene(open(p1, "rb"))
    assert np.array_equal(got["bound_vec"], b) and np.array_equal(got["ior"], scene["ior"]) and np.array_equal(got["translucency"], scene["translucency"])
    # (ii) we write a combined instance, the reference reads and traces it; the oracle agrees with what it computed
    p2 = str(tmp_path / "debug_raytrace_instance")
    iio.write_instance(open(p2, "wb"), scene, rays)
    n = C.c_size_t(0)
    epos = np.zeros(900, np.uint32); edir = np.zeros(900, np.int16); eit = np.zeros(300, np.uint32)
    assert lib.vrtref_replay_instance_u32(p2.encode(), C.byref(n), epos.ctypes.data_as(C.c_void_p), edir.ctypes.data_as(C.c_void_p),
                                          eit.ctypes.data_as(C.c_void_p), C.c_size_t(300)) == 0, lib.vrtref_last_error()
    assert n.value == 300
    ob, _, planes, trc = oracle.prep(shape, scene["ior"], scene["translucency"])
    p, d = oracle.normalise(shape, scene["ior"], rays["start_position"], rays["start_direction"])
    want = oracle.trace(oracle.fold(planes, trc), ob, p, d, [1, 1, 1], 200, round_mode=oracle.ROUND_HOST)
    assert np.array_equal(epos.reshape(-1, 3), want[0] + np.uint32(0x10000)) and np.array_equal(eit, want[2])
