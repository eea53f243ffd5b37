"""Pins the oracle against committed outputs of the unmodified reference CPU build (tests/golden/*.npz, produced by
tests/golden/make_golden.py).  Runs anywhere -- no /root/reference needed."""
import os

import numpy as np
import pytest

from tests import scenes as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EQ = np.array_equal


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


@pytest.mark.parametrize("kind", ["u32", "f32"])
def test_scaling_test_golden(oracle, kind):
    g = load("scaling_test_" + kind)
    inp = S.scaling_test_inputs(kind)
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    assert EQ(planes[0].reshape(ob)[:, 3, 3], g["diff0_x_profile"])
    vol = oracle.fold(planes, trc)
    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    ep, ed, ei, li, pa = oracle.trace(vol, ob, p2, d2, inp["invscale"], inp["iterations"], trace_path=True, round_mode=oracle.ROUND_HOST)
    assert EQ(ep + np.uint32(0x10000), g["end_position"]) and EQ(ed, g["end_direction"])
    assert EQ(ei, g["end_iteration"]) and EQ(li, g["remaining_light"])
    assert EQ((pa + np.uint32(0x10000))[:, ::97, :], g["path_every_97"])
    assert ei.tolist() == S.SCALING_KNOWN[kind]["eit"]


@pytest.mark.parametrize("kind,dirk", [("f32", "f32"), ("u32", "i16")])
@pytest.mark.parametrize("shape", [(24, 20, 28), (33, 17)])
def test_api_level_golden(oracle, kind, dirk, shape):
    g = load("api_%s_%dd" % (kind, len(shape)))
    ior, tr = S.random_scene(shape, seed=11 + len(shape), kind=kind, opaque_fraction=0.01)
    pos, d = S.random_rays(shape, 1500, seed=5, dir_kind=dirk)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    vol = oracle.fold(planes, trc)
    assert EQ(vol, g["volume"])
    isc = [1.0, 0.75, 1.5][:len(shape)]
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    ep, ed, ei, li, _ = oracle.trace(vol, ob, p2, d2, isc, 400, round_mode=oracle.ROUND_HOST)
    assert EQ(ep + np.uint32(0x10000), g["end_position"]) and EQ(ed, g["end_direction"])
    assert EQ(ei, g["end_iteration"]) and EQ(li, g["remaining_light"])


@pytest.mark.parametrize("volk", ["f32", "i16"])
def test_live_translucency_golden(oracle, volk):
    g = load("live_" + volk)
    ob = [18, 24, 20]
    ep, ed, ei, li, _ = oracle.trace(g["volume"], ob, g["start_position"], g["start_direction"], [1, 1, 1], 300,
                                     translucency=g["translucency"], min_brightness=0x40000000, round_mode=oracle.ROUND_HOST)
    assert EQ(ep, g["end_position"]) and EQ(ed, g["end_direction"]) and EQ(ei, g["end_iteration"]) and EQ(li, g["remaining_light"])
    assert np.any(li < 0x40000000)


def test_step_count_semantics(oracle):
    """cu:333-335,350-351,953-956: k+1 when the ray leaves after k body steps, k when stopped in body k, `iterations` at
    the cap, 1 for a ray that starts out of bounds."""
    b = [8, 8, 8]
    vol = np.zeros((512, 4), np.float32); vol[:, 3] = -32768.0
    pos = np.array([[0x10000, 0x30000, 0x30000], [0xFFFF0000, 0x30000, 0x30000], [0x10000, 0x30000, 0x30000]], np.uint32)
    d = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0]], np.float32)
    ep, ed, ei, li, _ = oracle.trace(vol, b, pos, d, [1, 1, 1], 1000)
    step = 16896                                     # rni(0x42000000 / 65536): 0.2578 voxel per step at |T| = 1
    k = -(-(7 * 65536 - 0x10000) // step)            # body steps until pos >> 16 reaches bounds - 1 = 7
    assert ei[0] == k + 1 and ei[1] == 1 and np.all(li == 0xFFFFFFFF)
    ep2, _, ei2, _, _ = oracle.trace(vol, b, pos[:1], d[:1], [1, 1, 1], 5)
    assert ei2[0] == 5 and ep2[0, 0] == 0x10000 + 4 * step      # cap: iterations-1 body steps, reports `iterations`
    vol2 = vol.copy(); vol2[:, 3] = 1.0              # opaque everywhere: stopped in body 1
    _, _, ei3, _, _ = oracle.trace(vol2, b, pos[:1], d[:1], [1, 1, 1], 1000)
    assert ei3[0] == 1
