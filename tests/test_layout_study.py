"""Layout-study variants (z-pair layout, point-sampled 3-D texture): they lost the study and are compiled only into
volumeraytracer_b200/libvrt_b200_study.so (`make -C volumeraytracer_b200/csrc study`, -DVRT_STUDY).  The shipped library answers
VRT_ERR_UNSUPPORTED for them.  The parity tests below run against the study build in a child process (the ctypes binding loads one
library per process: VRT_B200_LIB selects it)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu
EQ = np.array_equal
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUDY_LIB = os.path.join(ROOT, "volumeraytracer_b200", "libvrt_b200_study.so")
IN_STUDY = os.environ.get("VRT_B200_LIB", "").endswith("_study.so")
needs_study = pytest.mark.skipif(not IN_STUDY, reason="runs in the child process started by test_study_variants_in_the_study_build")


@pytest.fixture(scope="module")
def vrt():
    import volumeraytracer_b200 as v
    v.lib()
    return v


def _assert_same(got, want, what=""):
    names = ["end_position", "end_direction", "end_iteration", "remaining_light", "path"]
    for g, w, nme in zip(got, want, names):
        if w is None:
            continue
        assert EQ(g, w), "%s %s differs (%d of %d)" % (what, nme, int(np.sum(g != w)), g.size)


@pytest.mark.skipif(IN_STUDY, reason="parent-process test")
def test_shipped_library_rejects_study_layouts(vrt, oracle):
    shape = (12, 12, 12)
    ior, tr = S.random_scene(shape, seed=1, kind="f32")
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    for kw in ({"paired": True}, {"texture": True}):
        with pytest.raises(vrt.VrtError) as e:
            vrt.TraceRaysCu(ob, planes, trc, **kw)
        assert e.value.code == 4 and "VRT_STUDY" in str(e.value)


@pytest.mark.skipif(IN_STUDY, reason="parent-process test")
def test_study_variants_in_the_study_build():
    if not os.path.exists(STUDY_LIB):
        pytest.skip("libvrt_b200_study.so not built (make -C volumeraytracer_b200/csrc study)")
    env = dict(os.environ, VRT_B200_LIB=STUDY_LIB)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


@needs_study
@pytest.mark.parametrize("volk", ["f32", "i16"])
def test_texture_layout_is_bit_identical(vrt, oracle, volk):
    """VRT_SCENE_LAYOUT_TEXTURE: corners point-sampled from a block-linear CUDA 3-D array -- same bits out."""
    shape = (31, 36, 29)
    ior, tr = S.random_scene(shape, seed=19, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.004)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 8000, seed=3, dir_kind="f32", scale=1.2)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.25, 0.8], 300, round_mode=oracle.ROUND_DEVICE)
    t = vrt.TraceRaysCu(ob, planes, trc, texture=True)
    for refill in (0, 1, 32):
        t.set_option(vrt.VRT_OPT_REFILL, refill)
        _assert_same(t.trace_rays_cu(pos, d, [1.0, 1.25, 0.8], 0, 300), want[:4], "texture refill=%d" % refill)


@needs_study
@pytest.mark.parametrize("volk", ["f32", "i16"])
def test_pair_layout_is_bit_identical(vrt, oracle, volk):
    """VRT_SCENE_LAYOUT_PAIR: {voxel, z neighbour} per cell, one 256-bit load per corner row -- same bits out, with and without
    live translucency, with path output, in region mode, through host chunking; the download gives back the reference layout."""
    shape = (31, 36, 29)
    ior, tr = S.random_scene(shape, seed=23, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.004)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    trl = trc.copy(); trl[trl != 0] -= np.uint32(1 << 22)
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 9000, seed=5, dir_kind="f32", scale=1.3)
    pos = pos - np.uint32(0x10000) + np.uint32(0x2345)
    isc = [1.0, 1.25, 0.8]
    t = vrt.TraceRaysCu(ob, planes, trc, paired=True)
    got_vol, got_tr = t.download_volume()
    assert EQ(np.asarray(got_vol).reshape(-1), np.asarray(vol).reshape(-1)) and EQ(got_tr, trc)
    want = oracle.trace(vol, ob, pos, d, isc, 300, round_mode=oracle.ROUND_DEVICE)
    for refill, poll, chunk in ((32, 128, 0), (0, 7, 0), (1, 1, 1000), (8, 32, 0)):
        t.set_option(vrt.VRT_OPT_REFILL, refill); t.set_option(vrt.VRT_OPT_STEPS_PER_POLL, poll); t.set_option(vrt.VRT_OPT_CHUNK_RAYS, chunk)
        _assert_same(t.trace_rays_cu(pos, d, isc, 0, 300), want[:4], "pair refill=%d poll=%d chunk=%d" % (refill, poll, chunk))
    t.set_option(vrt.VRT_OPT_CHUNK_RAYS, 0)
    wantp = oracle.trace(vol, ob, pos[:500], d[:500], isc, 120, trace_path=True, round_mode=oracle.ROUND_DEVICE)
    _assert_same(t.trace_rays_cu(pos[:500], d[:500], isc, 0, 120, trace_paths=True), wantp, "pair paths")
    t.close()
    tl = vrt.TraceRaysCu(ob, planes, trl, paired=True)
    wantl = oracle.trace(oracle.fold(planes, trl), ob, pos, d, isc, 300, translucency=trl, min_brightness=0x40000000, round_mode=oracle.ROUND_DEVICE)
    _assert_same(tl.trace_rays_cu(pos, d, isc, 0x40000000, 300, live_translucency=True), wantl[:4], "pair live")
    wantr = oracle.trace(oracle.fold(planes, trl), ob, pos, d, isc, 300, round_mode=oracle.ROUND_DEVICE)
    tl.set_option(vrt.VRT_OPT_REGION_LOG2, 5)
    _assert_same(tl.trace_rays_cu(pos, d, isc, 0, 300), wantr[:4], "pair region mode")
    tl.close()


