"""Pins the C restatement (oracle/vrt_oracle.c) against the UNMODIFIED reference CPU build
(oracle/_ref/libvrt_ref.so, compiled from /root/reference/src by oracle/Makefile): bit-exact on the
reference's own known-answer inputs and on seeded random scenes.  Skipped where the reference build is
absent; tests/test_golden.py then covers the same ground from committed fixtures."""
import numpy as np
import pytest

from tests import scenes as S

ALL_EQ = np.array_equal


@pytest.mark.parametrize("kind", ["u32", "f32"])
def test_scaling_test_bit_exact(oracle, refimpl, kind):
    inp = S.scaling_test_inputs(kind)
    sc = refimpl.RefScene(inp["bounds"], inp["ior"], inp["translucency"])
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    assert list(ob) == [int(x) for x in sc.diff_bounds]
    assert ALL_EQ(iorlog, sc.iorlog())
    for a in range(3):
        assert ALL_EQ(planes[a], sc.diff(a))
    assert ALL_EQ(trc, sc.translucency_cropped())
    vol = oracle.fold(planes, trc)
    assert ALL_EQ(vol, sc.interleaved())

    rep, red, rei, rli, rpa = sc.trace(inp["pos"], inp["dir"], inp["invscale"], 0, inp["iterations"], trace_path=True)
    known = S.SCALING_KNOWN[kind]
    assert rep.ravel().tolist() == known["epos"] and rei.tolist() == known["eit"]

    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    ep, ed, ei, li, pa = oracle.trace(vol, ob, p2, d2, inp["invscale"], inp["iterations"], trace_path=True,
                                      round_mode=oracle.ROUND_HOST)
    assert ALL_EQ(ep + np.uint32(0x10000), rep)
    assert ALL_EQ(ed, red)
    assert ALL_EQ(ei, rei)
    assert ALL_EQ(li, rli)
    assert ALL_EQ(pa + np.uint32(0x10000), rpa)
    # the reference's own assertions (cuda_volume_raytracer_test.h:48-52)
    assert abs(int(ei[0]) - 46718) <= 100 and abs(int(ei[1]) - 46718) <= 100


def test_interpolation_test(oracle, refimpl):
    """image_util_test.h:4-35: 5^3 ramp, exact integer equality of the host interpolator."""
    b = [5, 5, 5]
    img = np.arange(125, dtype=np.int32) * 0x10000
    pos = np.array([[0, 0, 0], [0x10000, 0, 0], [0, 0x10000, 0], [0, 0, 0x10000], [0x8000, 0, 0], [0, 0x8000, 0],
                    [0, 0, 0x8000], [0x8000, 0x8000, 0x8000], [0x38000, 0x18000, 0x2C000]], dtype=np.uint32)
    assert ALL_EQ(oracle.interp(img, b, pos), refimpl.interpolate(img, b, pos))
    rng = np.random.default_rng(7)
    pos = rng.integers(0, 4 * 0x10000, size=(500, 3), dtype=np.uint32)
    for dt in (np.int32, np.uint32, np.float32):
        im = rng.integers(0, 1 << 20, size=125).astype(dt)
        assert ALL_EQ(oracle.interp(im, b, pos), refimpl.interpolate(im, b, pos))


@pytest.mark.parametrize("kind,dirk", [("f32", "f32"), ("u32", "i16")])
@pytest.mark.parametrize("shape", [(24, 20, 28), (33, 17)])
def test_api_level_random(oracle, refimpl, kind, dirk, shape):
    """RaytraceScene<> end to end (prep + normalise + march + coordinate shift) on seeded random scenes."""
    ior, tr = S.random_scene(shape, seed=11 + len(shape), kind=kind, opaque_fraction=0.01)
    pos, d = S.random_rays(shape, 3000, seed=5, dir_kind=dirk, scale=1.0)
    sc = refimpl.RefScene(shape, ior, tr)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    for a in range(len(shape)):
        assert ALL_EQ(planes[a], sc.diff(a)), "stencil axis %d" % a
    vol = oracle.fold(planes, trc)
    assert ALL_EQ(vol, sc.interleaved())
    isc = [1.0, 0.75, 1.5][:len(shape)]
    rep, red, rei, rli, rpa = sc.trace(pos, d, isc, 0, 400, trace_path=True)
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    ep, ed, ei, li, pa = oracle.trace(vol, ob, p2, d2, isc, 400, trace_path=True, round_mode=oracle.ROUND_HOST)
    assert ALL_EQ(ei, rei)
    assert ALL_EQ(ep + np.uint32(0x10000), rep)
    assert ALL_EQ(ed, red)
    assert ALL_EQ(li, rli)
    assert ALL_EQ(pa + np.uint32(0x10000), rpa)
    assert len(np.unique(ei)) > 3          # rays really terminate at different steps (exit / opaque / cap)


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
@pytest.mark.parametrize("shape", [(20, 26, 22), (40, 31)])
def test_boundary_level_all_four_combos(oracle, refimpl, volk, dirk, shape):
    """TraceRaysCu<Diff>::trace_rays_cu<Dir> for all four explicit instantiations (cu:992-1049)."""
    dim = len(shape)
    ior, tr = S.random_scene(shape, seed=3, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.01)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    t = refimpl.RefTracer(ob, planes, trc)
    vol = oracle.fold(planes, trc)
    assert ALL_EQ(vol, t.interleaved())
    pos, d = S.random_rays(ob, 2000, seed=9, dir_kind=dirk, scale=1.3)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)      # cropped coordinates, includes cells at index 0
    isc = [1.0, 1.25, 0.8][:dim]
    rep, red, rei, rli, rpa = t.trace(pos, d, isc, 12345, 300, trace_path=True)
    ep, ed, ei, li, pa = oracle.trace(vol, ob, pos, d, isc, 300, trace_path=True, round_mode=oracle.ROUND_HOST)
    assert ALL_EQ(ei, rei) and ALL_EQ(ep, rep) and ALL_EQ(ed, red) and ALL_EQ(li, rli) and ALL_EQ(pa, rpa)
    assert np.all(li == 0xFFFFFFFF)        # shipped behaviour: attenuation compiled out (cu:853...)


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
@pytest.mark.parametrize("shape", [(20, 26, 22), (40, 31)])
def test_live_translucency(oracle, refimpl, volk, dirk, shape):
    """The live-translucency / minimum-brightness instantiation of trace_rays_cpu (cu:337-341)."""
    dim = len(shape)
    ior, tr = S.random_scene(shape, seed=21, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.003)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    trc = (trc.astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    trc[trc != 0] -= np.uint32(1 << 24)                       # strong absorption so min_brightness triggers
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 2000, seed=10, dir_kind=dirk, scale=1.1)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    isc = [1.0] * dim
    minb = 0x40000000
    rep, red, rei, rli, rpa = refimpl.trace_live(vol, trc, ob, isc, pos, d, 300, minb, trace_path=True)
    ep, ed, ei, li, pa = oracle.trace(vol, ob, pos, d, isc, 300, translucency=trc, min_brightness=minb,
                                      trace_path=True, round_mode=oracle.ROUND_HOST)
    assert ALL_EQ(ei, rei) and ALL_EQ(ep, rep) and ALL_EQ(ed, red) and ALL_EQ(li, rli) and ALL_EQ(pa, rpa)
    assert np.any(li < minb) and np.any(li == 0xFFFFFFFF) or np.any(li < 0xFFFFFFFF)


def test_device_vs_host_rounding_within_tolerance(oracle):
    """ROUND_DEVICE (cvt.rni, the reference's CUDA build) and ROUND_HOST (std::round, its CPU build) differ only
    on exact .5 ties: on a smooth field end positions agree to << 1e-3 voxel and step counts are identical."""
    shape = (48, 40, 44)
    from volumeraytracer_b200 import workloads as W
    ior = W.ior_random_smooth(shape, 2, lo=1.0, hi=1.1)       # smooth field (north_star: "on smooth fields")
    tr = np.full(shape, 0xFFFFFFFF, np.uint32)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 4000, seed=4)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    a = oracle.trace(vol, ob, pos, d, [1, 1, 1], 500, round_mode=oracle.ROUND_HOST)
    b = oracle.trace(vol, ob, pos, d, [1, 1, 1], 500, round_mode=oracle.ROUND_DEVICE)
    same = a[2] == b[2]
    assert np.mean(~same) < 0.002          # a tie can move an exit by one step on a handful of rays
    dp = np.abs(a[0].astype(np.int64) - b[0].astype(np.int64))[same].max() / 65536.0
    assert dp <= 1e-3                      # north_star tolerance: 1e-3 voxel
    assert np.abs(a[1] - b[1])[same].max() <= 1e-5   # 1e-5 rad on unit-length directions
