"""Parity tests proper: the CUDA path, called through the C ABI (libvrt_b200.so), against
  (a) the CPU oracle in ROUND_DEVICE mode -- bit-exact (positions, directions, step counts, light, paths),
  (b) the reference's own CUDA kernel compiled for sm_100 (oracle/_ref/libvrt_ref_cuda.so) -- bit-exact,
  (c) the reference's CPU build through the oracle's ROUND_HOST mode -- within the north_star tolerance
      (1e-3 voxel, 1e-5 rad on smooth fields).
Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu
EQ = np.array_equal

KERNELS = [1, 2, 3]
REFILLS = [0, 1, 16, 32]


@pytest.fixture(scope="module")
def vrt():
    import volumeraytracer_b200 as v
    v.lib()
    return v


def _mk(vrt, oracle, shape, seed, volk, opaque=0.01, absorb=False):
    ior, tr = S.random_scene(shape, seed=seed, kind="f32" if volk == "f32" else "u32", opaque_fraction=opaque)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    if absorb:
        trc = trc.copy()
        trc[trc != 0] -= np.uint32(1 << 24)
    vol = oracle.fold(planes, trc)
    tracer = vrt.TraceRaysCu(ob, planes, trc)
    return ob, planes, trc, vol, tracer


def _assert_same(got, want, what=""):
    names = ["end_position", "end_direction", "end_iteration", "remaining_light", "path"]
    for g, w, nme in zip(got, want, names):
        if w is None:
            continue
        assert EQ(g, w), "%s %s differs (%d of %d)" % (what, nme, int(np.sum(g != w)), g.size)


@pytest.mark.parametrize("kind", ["u32", "f32"])
def test_scaling_test_known_answer(vrt, oracle, kind):
    """The reference's scaling_test (cuda_volume_raytracer_test.h:4-75) through the boundary class."""
    inp = S.scaling_test_inputs(kind)
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    vol = oracle.fold(planes, trc)
    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    want = oracle.trace(vol, ob, p2, d2, inp["invscale"], inp["iterations"], trace_path=True, round_mode=oracle.ROUND_DEVICE)
    t = vrt.TraceRaysCu(ob, planes, trc)
    got = t.trace_rays_cu(p2, d2, inp["invscale"], 0, inp["iterations"], trace_paths=True)
    _assert_same(got, want, "scaling_test")
    # the reference's own assertions: step count 46718 +- 100 and |T| = n at the exit (cuda_volume_raytracer_test.h:48-52)
    assert abs(int(got[2][0]) - 46718) <= 100 and abs(int(got[2][1]) - 46718) <= 100
    unit = 1.0 if kind == "f32" else 256.0
    ratio = got[1][:, 0].astype(np.float64) / inp["dir"][:, 0].astype(np.float64)
    n_end = oracle.interp(inp["ior"], inp["bounds"], got[0] + np.uint32(0x10000)).astype(np.float64) / (1.0 if kind == "f32" else 65536.0)
    assert np.all(np.abs(ratio - n_end) <= 1e-5 + (0 if kind == "f32" else 1.0 / 256) + 2e-3), (ratio, n_end, unit)
    # and against the reference CPU's known answers: same step counts, positions within 1e-3 voxel
    known = S.SCALING_KNOWN[kind]
    assert got[2].tolist() == known["eit"]
    diff = ((got[0].ravel() + np.uint32(0x10000)) - np.array(known["epos"], dtype=np.uint32)).astype(np.int32)   # modular: ray 1 exits below 0
    assert np.abs(diff).max() <= 66


@pytest.mark.parametrize("kver", KERNELS)
@pytest.mark.parametrize("refill", REFILLS)
def test_kernel_variants_bit_exact(vrt, oracle, kver, refill):
    ob, planes, trc, vol, t = _mk(vrt, oracle, (40, 36, 44), 5, "f32")
    pos, d = S.random_rays(ob, 20000, seed=77)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1.0, 0.9, 1.1], 600, round_mode=oracle.ROUND_DEVICE)
    t.set_option(vrt.VRT_OPT_KERNEL, kver)
    t.set_option(vrt.VRT_OPT_REFILL, refill)
    for block in (64, 128, 256):
        t.set_option(vrt.VRT_OPT_BLOCK_THREADS, block)
        got = t.trace_rays_cu(pos, d, [1.0, 0.9, 1.1], 0, 600)
        _assert_same(got, want, "kver=%d refill=%d block=%d" % (kver, refill, block))
    assert len(np.unique(want[2])) > 20


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
@pytest.mark.parametrize("live", [False, True])
def test_all_type_combinations_and_paths(vrt, oracle, volk, dirk, live):
    ob, planes, trc, vol, t = _mk(vrt, oracle, (30, 34, 28), 9, volk, opaque=0.004, absorb=live)
    pos, d = S.random_rays(ob, 5000, seed=3, dir_kind=dirk, scale=1.2)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    isc = [1.0, 1.25, 0.8]
    minb = 0x40000000
    want = oracle.trace(vol, ob, pos, d, isc, 300, translucency=trc if live else None, min_brightness=minb, trace_path=True,
                        round_mode=oracle.ROUND_DEVICE)
    got = t.trace_rays_cu(pos, d, isc, minb, 300, trace_paths=True, live_translucency=live)
    _assert_same(got, want, "paths")
    for kver in KERNELS:
        t.set_option(vrt.VRT_OPT_KERNEL, kver)
        got = t.trace_rays_cu(pos, d, isc, minb, 300, live_translucency=live)
        _assert_same(got, want[:4], "kver %d" % kver)
    if live:
        assert np.any(got[3] < minb) and np.any(got[3] >= minb)      # both exit classes are exercised
    else:
        assert np.all(got[3] == 0xFFFFFFFF)


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
@pytest.mark.parametrize("live", [False, True])
def test_two_dimensional(vrt, oracle, volk, dirk, live):
    """dim == 2 (f4), including the reference's second-x-lerp quirk (cu:207-208)."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (60, 47), 13, volk, opaque=0.004, absorb=live)
    pos, d = S.random_rays(ob, 4000, seed=8, dir_kind=dirk, scale=1.1)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.1], 250, translucency=trc if live else None, min_brightness=0x40000000,
                        trace_path=True, round_mode=oracle.ROUND_DEVICE)
    got = t.trace_rays_cu(pos, d, [1.0, 1.1], 0x40000000, 250, trace_paths=True, live_translucency=live)
    _assert_same(got, want, "2-D")


def test_edge_cases(vrt, oracle):
    ob, planes, trc, vol, t = _mk(vrt, oracle, (20, 22, 24), 1, "f32")
    # empty batch
    out = t.trace_rays_cu(np.zeros((0, 3), np.uint32), np.zeros((0, 3), np.float32), [1, 1, 1], 0, 100)
    assert out[0].shape == (0, 3) and out[2].shape == (0,)
    # ragged sizes: 1, 31, 33, 129 rays; rays starting out of bounds (report 1 step), zero direction (NaN step),
    # iterations == 1 (cap immediately), negative-going rays that wrap below zero
    for n in (1, 31, 33, 129):
        pos, d = S.random_rays(ob, n, seed=n)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
        pos[0] = [0xFFFF0000, 0x20000, 0x20000]
        if n > 2:
            d[1] = 0.0
            d[2] = [-1.0, 0.0, 0.0]
            pos[2] = [0x1000, 0x28000, 0x28000]
        for its in (1, 2, 50):
            want = oracle.trace(vol, ob, pos, d, [1, 1, 1], its, round_mode=oracle.ROUND_DEVICE)
            got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, its)
            _assert_same(got, want, "n=%d its=%d" % (n, its))
            assert got[2][0] == 1
    # argument errors are reported, not crashed on
    with pytest.raises(vrt.VrtError):
        t.trace_rays_cu(np.zeros(5, np.uint32), np.zeros(5, np.float32), [1, 1, 1], 0, 10)
    with pytest.raises(vrt.VrtError):
        vrt.TraceRaysCu([4, 4, 4, 4], [np.zeros(256, np.float32)] * 4, np.zeros(256, np.uint32))


def test_host_call_chunking_and_in_place(vrt, oracle):
    """vrt_trace pipelines chunks over two streams and writes results back in place on the device."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (36, 30, 33), 4, "f32")
    pos, d = S.random_rays(ob, 30011, seed=12)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1, 1, 1], 400, round_mode=oracle.ROUND_DEVICE)
    for chunk in (0, 1000, 4096, 30011):
        t.set_option(vrt.VRT_OPT_CHUNK_RAYS, chunk)
        got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, 400)
        _assert_same(got, want, "chunk %d" % chunk)


def test_gpu_scene_prep_and_api_level(vrt, oracle):
    """f1 + f2 on the GPU: RaytraceScene (prep, normalise, march, coordinate shifts) vs the oracle's restatement."""
    import torch
    for kind, dirk, shape in (("f32", "f32", (26, 30, 22)), ("u32", "i16", (26, 30, 22)), ("f32", "f32", (40, 33)), ("u32", "i16", (40, 33))):
        ior, tr = S.random_scene(shape, seed=31, kind=kind, opaque_fraction=0.01)
        ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
        vol = oracle.fold(planes, trc)
        sc = vrt.RaytraceScene(shape, ior, tr)
        co = sc._calculation_object
        assert co._output_sizes == list(ob)
        host, host_tr = co.download_volume()
        assert EQ(host_tr, trc)
        mism = np.sum(host.reshape(vol.shape) != vol)
        assert mism <= max(1, vol.size // 1000000), "GPU scene prep differs from the oracle in %d of %d values" % (mism, vol.size)
        pos, d = S.random_rays(shape, 3000, seed=6, dir_kind=dirk)
        isc = [1.0, 0.75, 1.5][:len(shape)]
        p2, d2 = oracle.normalise(shape, ior, pos, d)
        want = oracle.trace(host.reshape(vol.shape), ob, p2, d2, isc, 300, trace_path=True, round_mode=oracle.ROUND_DEVICE)
        got = sc.trace_rays(pos, d, isc, 0, 300, trace_path=True)
        assert EQ(got[0], want[0] + np.uint32(0x10000))
        assert EQ(got[1], want[1]) and EQ(got[2], want[2]) and EQ(got[3], want[3])
        assert EQ(got[4], want[4] + np.uint32(0x10000))
        # out-of-range start is an error, as in the reference (image_util.cpp:686-691)
        bad = pos.copy(); bad[5, 0] = 0x8000
        with pytest.raises(vrt.VrtError):
            sc.trace_rays(bad, d, isc, 0, 10)
        sc.close()


def test_against_reference_cuda_kernel(vrt, oracle):
    """Bit-exact against the reference's OWN CUDA kernel (trace_rays_gpu, cu:397-414) compiled for sm_100."""
    from oracle import ref
    if not ref.available(cuda=True):
        pytest.skip("oracle/_ref/libvrt_ref_cuda.so not built")
    for volk, dirk in (("f32", "f32"), ("i16", "i16"), ("f32", "i16"), ("i16", "f32")):
        ob, planes, trc, vol, t = _mk(vrt, oracle, (34, 30, 38), 17, volk)
        pos, d = S.random_rays(ob, 9000, seed=23, dir_kind=dirk)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
        rt = ref.RefTracer(ob, planes, trc, cuda=True)
        want = rt.trace(pos, d, [1, 1, 1], 0, 500)
        got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, 500)
        _assert_same(got, want[:4], "vs reference CUDA %s/%s" % (volk, dirk))
        rt.close()


def test_reference_boundary_class_on_the_dropin(vrt, oracle):
    """The reference's own class TraceRaysCu<float|diff_t> (cuda_volume_raytracer.h:61-115), constructed and called exactly as
    image_util.cpp does, with its member functions DEFINED by the drop-in shim: all four volume x direction type combinations,
    3-D and 2-D, with and without path output, equal the reference's CUDA build and the oracle bit for bit."""
    from oracle import ref
    if not ref.available("dropin"):
        pytest.skip("oracle/_ref/libvrt_dropin.so not built")
    for volk, dirk in (("f32", "f32"), ("i16", "i16"), ("f32", "i16"), ("i16", "f32")):
        for shape in ((34, 30, 38), (60, 45)):
            ob, planes, trc, vol, t = _mk(vrt, oracle, shape, 17, volk)
            t.close()
            pos, d = S.random_rays(ob, 40000, seed=23, dir_kind=dirk)           # > 32 768: more than one reference chunk
            pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
            isc = [1.0] * len(shape)
            rt = ref.RefTracer(ob, planes, trc, cuda="dropin")
            got = rt.trace(pos, d, isc, 0, 300)
            want = oracle.trace(vol, ob, pos, d, isc, 300, round_mode=oracle.ROUND_DEVICE)
            _assert_same(got, want[:4], "TraceRaysCu<> on the drop-in %s/%s %s" % (volk, dirk, shape))
            gotp = rt.trace(pos[:300], d[:300], isc, 0, 50, trace_path=True)
            wantp = oracle.trace(vol, ob, pos[:300], d[:300], isc, 50, trace_path=True, round_mode=oracle.ROUND_DEVICE)
            _assert_same(gotp, wantp, "TraceRaysCu<> on the drop-in, paths %s/%s %s" % (volk, dirk, shape))
            rt.close()


def test_within_tolerance_of_reference_cpu_on_smooth_field(vrt, oracle):
    """north_star tolerance vs the reference's CPU trace (ROUND_HOST = bit-exact restatement of it): end positions
    within 1e-3 voxel, directions within 1e-5 rad, identical termination step counts, on a smooth analytic field."""
    from volumeraytracer_b200 import workloads as W
    size = 64
    ior = W.ior_sines(size, base=1.3, amp=0.05, period=48.0)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    ob, iorlog, planes, trc = oracle.prep(ior.shape, ior, tr)
    vol = oracle.fold(planes, trc)
    t = vrt.TraceRaysCu(ob, planes, trc)
    pos, d = W.rays_parallel_x(96, 96, 3.0, size - 4.0, x0=2.0)
    pos, d = oracle.normalise(ior.shape, ior, pos, d)
    cpu = oracle.trace(vol, ob, pos, d, [1, 1, 1], 1024, round_mode=oracle.ROUND_HOST)
    got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, 1024)
    assert EQ(got[2], cpu[2]), "termination step counts differ on %d rays" % int(np.sum(got[2] != cpu[2]))
    dp = np.abs(got[0].astype(np.int64) - cpu[0].astype(np.int64)).max() / 65536.0
    assert dp <= 1e-3, dp
    a, b = got[1].astype(np.float64), cpu[1].astype(np.float64)
    cosang = np.sum(a * b, axis=1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    ang = np.arccos(np.clip(cosang, -1, 1))
    assert ang.max() <= 1e-5, ang.max()


def test_config1_constant_index(vrt, oracle):
    """BASELINE config 1: 64^3 constant-index volume, 4096 straight rays -- directions unchanged bit-exactly, y/z
    unchanged, all rays the same step count, remaining_light 0xFFFFFFFF; float and int16 variants."""
    from volumeraytracer_b200 import workloads as W
    size = 64
    for kind in ("f32", "u32"):
        ior = W.ior_constant(size, 1.0)
        if kind == "u32":
            ior = W.ior_to_u32(ior)
        tr = np.full((size,) * 3, 0xFFFFFFFF, np.uint32)
        pos, d = W.rays_parallel_x(64, 64, 1.5, 61.5, x0=1.5)   # < bound-2: the last cell row is outside the marcher's test (cu:335)
        if kind == "u32":
            d = W.dirs_to_i16(d)
        sc = vrt.RaytraceScene((size,) * 3, ior, tr)
        epos, edir, eit, light, _ = sc.trace_rays(pos, d, [1, 1, 1], 0, 1024)
        assert EQ(edir, d), "directions must be unchanged in a constant-index volume"
        assert EQ(epos[:, 1:], pos[:, 1:])
        assert len(np.unique(eit)) == 1 and 200 < int(eit[0]) < 280
        assert np.all(light == 0xFFFFFFFF)
        ob, iorlog, planes, trc = oracle.prep((size,) * 3, ior, tr)
        p2, d2 = oracle.normalise((size,) * 3, ior, pos, d)
        want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, [1, 1, 1], 1024, round_mode=oracle.ROUND_DEVICE)
        assert EQ(epos, want[0] + np.uint32(0x10000)) and EQ(eit, want[2])
        sc.close()


def test_size_independent_properties_large(vrt, oracle):
    """At a size the CPU oracle cannot cover in seconds: (i) all kernel variants and refill policies agree bit for bit
    (2M rays, 256^3 Luneburg lens = BASELINE config 2 geometry), (ii) a 1/64 subsample matches the oracle exactly,
    (iii) tracing in two halves equals tracing at once (ray independence), (iv) permuting the rays permutes the output."""
    from volumeraytracer_b200 import workloads as W
    size = 256
    ior = W.ior_luneburg(size, 100.0)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    sc = vrt.RaytraceScene(ior.shape, ior, tr)
    co = sc._calculation_object
    pos, d = W.rays_parallel_x(1024, 1024, 30.0, 225.0, x0=2.0)
    # normalise on the host oracle for the subsample, on the GPU for everything
    import torch
    dev0 = torch.device("cuda", co.device)     # not .cuda(): the reference's CUDA build (an earlier test) leaves another device current on multi-GPU boxes
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev0); tdir = torch.from_numpy(d.reshape(-1)).to(dev0)
    co.normalise_rays_device(tpos, tdir)
    results = []
    for kver, refill in ((1, 0), (2, 16), (3, 1), (3, 32)):
        co.set_option(vrt.VRT_OPT_KERNEL, kver); co.set_option(vrt.VRT_OPT_REFILL, refill)
        out = co.trace_device(tpos, tdir, [1, 1, 1], 0, 4096)
        torch.cuda.synchronize()
        results.append([o.cpu().numpy() for o in out])
    for r in results[1:]:
        for a, b in zip(results[0], r):
            assert EQ(a, b)
    base = results[0]
    n = pos.shape[0]
    # (ii) subsample vs oracle (volume read back from the device so both sides consume identical bits)
    sel = np.arange(0, n, 64)
    host_vol, _ = co.download_volume()
    p_h = tpos.cpu().numpy().view(np.uint32).reshape(-1, 3)[sel]; d_h = tdir.cpu().numpy().reshape(-1, 3)[sel]
    want = oracle.trace(host_vol, co._output_sizes, p_h, d_h, [1, 1, 1], 4096, round_mode=oracle.ROUND_DEVICE)
    assert EQ(base[0].view(np.uint32).reshape(-1, 3)[sel], want[0])
    assert EQ(base[1].reshape(-1, 3)[sel], want[1])
    assert EQ(base[2].view(np.uint32)[sel], want[2])
    # physics sanity of config 2: rays through the lens converge near the far focus (x ~ 227.5, y = z = 127.5)
    # (iii) halves
    h = n // 2
    o1 = co.trace_device(tpos[:h * 3].clone(), tdir[:h * 3].clone(), [1, 1, 1], 0, 4096)
    o2 = co.trace_device(tpos[h * 3:].clone(), tdir[h * 3:].clone(), [1, 1, 1], 0, 4096)
    torch.cuda.synchronize()
    assert EQ(np.concatenate([o1[0].cpu().numpy(), o2[0].cpu().numpy()]), base[0])
    assert EQ(np.concatenate([o1[2].cpu().numpy(), o2[2].cpu().numpy()]), base[2])
    # (iv) permutation
    perm = torch.randperm(n, device=dev0, generator=torch.Generator(device=dev0).manual_seed(3))
    pp = tpos.view(-1, 3)[perm].contiguous().view(-1); dd = tdir.view(-1, 3)[perm].contiguous().view(-1)
    op = co.trace_device(pp, dd, [1, 1, 1], 0, 4096)
    torch.cuda.synchronize()
    assert EQ(op[0].view(-1, 3).cpu().numpy(), base[0].reshape(-1, 3)[perm.cpu().numpy()])
    assert EQ(op[2].cpu().numpy(), base[2][perm.cpu().numpy()])
    sc.close()


@pytest.mark.parametrize("kind", ["u32", "f32"])
def test_dropin_behind_the_reference_cpp_api(vrt, oracle, kind):
    """The reference's UNMODIFIED C++ API (RaytraceScene<> from image_util.o: its own scene prep, normalisation and
    coordinate shifts) linked against our TraceRaysCu<> drop-in: scaling_test runs on the B200 marcher and gives what the
    oracle predicts for device rounding, and satisfies the reference's own assertions."""
    from oracle import ref
    if not ref.available("dropin"):
        pytest.skip("oracle/_ref/libvrt_dropin.so not built")
    inp = S.scaling_test_inputs(kind)
    sc = ref.RefScene(inp["bounds"], inp["ior"], inp["translucency"], which="dropin")
    got = sc.trace(inp["pos"], inp["dir"], inp["invscale"], 0, inp["iterations"], trace_path=True)
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    vol = oracle.fold(planes, trc)
    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    want = oracle.trace(vol, ob, p2, d2, inp["invscale"], inp["iterations"], trace_path=True, round_mode=oracle.ROUND_DEVICE)
    assert EQ(got[0], want[0] + np.uint32(0x10000)) and EQ(got[1], want[1]) and EQ(got[2], want[2]) and EQ(got[3], want[3])
    assert EQ(got[4], want[4] + np.uint32(0x10000))
    assert abs(int(got[2][0]) - 46718) <= 100 and abs(int(got[2][1]) - 46718) <= 100       # cuda_volume_raytracer_test.h:51-52
    # a batch large enough for the reference to have used its GPU path (> Options::_minimum_gpu = 0x80 rays)
    shape = (30, 28, 26)
    ior, tr = S.random_scene(shape, seed=41, kind=kind, opaque_fraction=0.01)
    pos, d = S.random_rays(shape, 5000, seed=2, dir_kind="f32" if kind == "f32" else "i16")
    sc2 = ref.RefScene(shape, ior, tr, which="dropin")
    got = sc2.trace(pos, d, [1, 1, 1], 0, 300)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, [1, 1, 1], 300, round_mode=oracle.ROUND_DEVICE)
    assert EQ(got[0], want[0] + np.uint32(0x10000)) and EQ(got[1], want[1]) and EQ(got[2], want[2])


# ---------------------------------------------------------------------------------------------------
# the five BASELINE.json configurations at sizes the CPU oracle finishes in seconds (bench.py runs config 5 at full size)

def _run_config(vrt, oracle, ior, tr, pos, d, iterations, live=False, minb=0):
    shape = ior.shape
    sc = vrt.RaytraceScene(shape, ior, tr)
    got = sc.trace_rays(pos, d, [1, 1, 1], minb, iterations, live_translucency=live)
    vol, trc = sc._calculation_object.download_volume()
    ob = sc._calculation_object._output_sizes
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    want = oracle.trace(vol, ob, p2, d2, [1, 1, 1], iterations, translucency=trc if live else None, min_brightness=minb,
                        round_mode=oracle.ROUND_DEVICE)
    assert EQ(got[0], want[0] + np.uint32(0x10000)) and EQ(got[1], want[1]) and EQ(got[2], want[2]) and EQ(got[3], want[3])
    # and the oracle's own scene prep agrees with the GPU's
    ob2, _, planes, trc2 = oracle.prep(shape, ior, tr)
    assert EQ(oracle.fold(planes, trc2), vol) and EQ(trc2, trc)
    sc.close()
    return got, p2


def test_config2_luneburg_lens(vrt, oracle):
    """config 2 geometry at 64^3: parallel +x rays through a Luneburg-style GRIN lens focus on the far side."""
    from volumeraytracer_b200 import workloads as W
    size, R = 64, 25.0
    ior = W.ior_luneburg(size, R)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    pos, d = W.rays_parallel_x(96, 96, 8.0, 55.0, x0=2.0)
    got, _ = _run_config(vrt, oracle, ior, tr, pos, d, 1024)
    # physics sanity: rays that entered the lens cross the axis near the rim focus (x = c + R, y = z = c)
    c = (size - 1) / 2.0
    y0 = pos[:, 1] / 65536.0 - c; z0 = pos[:, 2] / 65536.0 - c
    inside = y0 ** 2 + z0 ** 2 < (0.6 * R) ** 2
    e = got[0][inside] / 65536.0; dd = got[1][inside].astype(np.float64)
    # intersect the exit ray with the plane x = c + R, going backwards along the final direction
    t = (c + R - e[:, 0]) / dd[:, 0]
    yf = e[:, 1] + t * dd[:, 1] - c; zf = e[:, 2] + t * dd[:, 2] - c
    assert np.median(np.hypot(yf, zf)) < 1.5, np.median(np.hypot(yf, zf))


def test_config3_translucency_and_min_brightness(vrt, oracle):
    """config 3 at 96^3: live translucency plane + opaque ball + min_brightness: all three exit classes occur."""
    from volumeraytracer_b200 import workloads as W
    size = 96
    ior = W.ior_sines(size, period=32.0)
    tr = W.translucency_c3(size)
    absorb = (np.uint64(0xFFFFFFFF) - tr.astype(np.uint64)) * np.uint64(8)          # rescale absorption to the shorter paths
    tr2 = (np.uint64(0xFFFFFFFF) - np.minimum(absorb, np.uint64(0xFFFFFFFF))).astype(np.uint32)
    tr2[tr == 0] = 0
    pos, d = W.rays_parallel_x(128, 128, 4.0, size - 5.0, x0=2.0)
    minb = 0x40000000
    got, _ = _run_config(vrt, oracle, ior, tr2, pos, d, 4096, live=True, minb=minb)
    eit, light = got[2], got[3]
    escaped = light >= minb
    dimmed = light < minb
    assert escaped.sum() > 100 and dimmed.sum() > 100
    # rays aimed at the opaque ball stop early with light still above the threshold
    stopped_opaque = escaped & (eit < np.percentile(eit[escaped], 5))
    assert stopped_opaque.sum() > 10
    assert len(np.unique(eit)) > 50


def test_config4_harmonic_field_random_rays(vrt, oracle):
    from volumeraytracer_b200 import workloads as W
    size = 64
    ior = W.solve_harmonic(size, inner_radius=8.0, sweeps=60)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    pos, d = W.rays_random(30000, 8.0, size - 9.0, 0x5EED0004)
    got, _ = _run_config(vrt, oracle, ior, tr, pos, d, 4096)
    assert got[2].min() >= 2 and got[2].max() < 4096 and len(np.unique(got[2])) > 100


def test_config5_roofline_workload_small(vrt, oracle):
    """config 5 at 128^3 / cap 256: every ray runs the cap, so the reported steps are exactly rays x cap."""
    from volumeraytracer_b200 import workloads as W
    size = 128
    ior = W.ior_c5(size, period=32.0)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    pos, d = W.rays_parallel_x(256, 256, 2.0, size - 3.0, x0=2.0)
    got, _ = _run_config(vrt, oracle, ior, tr, pos, d, 256)
    inner = (pos[:, 1] > 0x80000) & (pos[:, 1] < (size - 8) << 16) & (pos[:, 2] > 0x80000) & (pos[:, 2] < (size - 8) << 16)
    assert np.all(got[2][inner] == 256)


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("live", [False, True])
def test_bricked_layout_is_bit_identical(vrt, oracle, volk, live):
    """VRT_SCENE_LAYOUT_BRICK (2x2x2-voxel bricks) changes only where bytes live, never a result; odd extents are padded."""
    shape = (31, 36, 29)
    ior, tr = S.random_scene(shape, seed=19, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.004)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    if live:
        trc = trc.copy(); trc[trc != 0] -= np.uint32(1 << 24)
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 8000, seed=3, dir_kind="f32" if volk == "f32" else "i16", scale=1.2)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.25, 0.8], 300, translucency=trc if live else None, min_brightness=0x40000000,
                        round_mode=oracle.ROUND_DEVICE)
    t = vrt.TraceRaysCu(ob, planes, trc, bricked=True, keep_i16=live)     # covers widened-float and kept-int16 staging
    for refill in (0, 1, 32):
        t.set_option(vrt.VRT_OPT_REFILL, refill)
        got = t.trace_rays_cu(pos, d, [1.0, 1.25, 0.8], 0x40000000, 300, live_translucency=live)
        _assert_same(got, want[:4], "bricked refill=%d" % refill)
    back, back_tr = t.download_volume()                 # handed out in the reference's linear order
    assert EQ(back, vol) and EQ(back_tr, trc)
    with pytest.raises(vrt.VrtError):
        t.trace_rays_cu(pos, d, [1, 1, 1], 0, 300, trace_paths=True)


def test_c_abi_in_place_and_concurrent_host_threads(vrt, oracle):
    """The boundary contract: results may be written back into the start buffers (java_binding.cpp:158-160), and one scene
    may be traced from several host threads at once (the JNI binding can be entered from any Java thread)."""
    import ctypes as C
    import threading
    from volumeraytracer_b200 import _lib
    ob, planes, trc, vol, t = _mk(vrt, oracle, (36, 30, 33), 4, "f32")
    lib = vrt.lib()
    isc = np.array([1, 1, 1], np.float32)
    results = {}

    def worker(k):
        pos, d = S.random_rays(ob, 20000 + 13 * k, seed=100 + k)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
        want = oracle.trace(vol, ob, pos, d, [1, 1, 1], 300, round_mode=oracle.ROUND_DEVICE, threads=2)
        p = pos.copy(); dd = d.copy()
        n = p.shape[0]
        eit = np.empty(n, np.uint32); light = np.empty(n, np.uint32)
        ok = True
        for _ in range(3):
            p[:] = pos; dd[:] = d
            rc = lib.vrt_trace(t._h, n, p.ctypes.data, dd.ctypes.data, _lib.VRT_F32, isc.ctypes.data, 0, 300, 0,
                               p.ctypes.data, dd.ctypes.data, eit.ctypes.data, light.ctypes.data, None)     # in place
            ok = ok and rc == 0 and EQ(p, want[0]) and EQ(dd, want[1]) and EQ(eit, want[2])
        results[k] = ok

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert results == {0: True, 1: True, 2: True, 3: True}


def test_int16_scene_large_batch_matches_reference_cuda(vrt, oracle):
    """int16 scene / int16 directions (the reference's default instantiation) on a 1M-ray batch: bit-exact against the
    reference's own CUDA kernel, which also checks our whole-batch launch against its 32 768-ray chunking."""
    from oracle import ref
    if not ref.available(cuda=True):
        pytest.skip("oracle/_ref/libvrt_ref_cuda.so not built")
    from volumeraytracer_b200 import workloads as W
    size = 96
    ior = W.ior_to_u32(W.ior_luneburg(size, 36.0))
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    ob, iorlog, planes, trc = oracle.prep(ior.shape, ior, tr)
    pos, d = W.rays_parallel_x(1024, 1024, 6.0, size - 7.0, x0=2.0)
    pos, d = oracle.normalise(ior.shape, ior, pos, W.dirs_to_i16(d))
    rt = ref.RefTracer(ob, planes, trc, cuda=True)
    want = rt.trace(pos, d, [1, 1, 1], 0, 2048)
    t = vrt.TraceRaysCu(ob, planes, trc)
    got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, 2048)
    _assert_same(got, want[:4], "1M rays int16")
    rt.close()


def test_int16_scene_staging_variants_agree(vrt, oracle):
    """An int16 scene is staged as float on the device by default (VRT_SCENE_KEEP_I16 keeps 8-byte voxels): same bits out,
    and the volume is handed back as int16 either way."""
    ob, planes, trc, vol, t_wide = _mk(vrt, oracle, (30, 34, 28), 9, "i16", opaque=0.004)
    t_keep = vrt.TraceRaysCu(ob, planes, trc, keep_i16=True)
    pos, d = S.random_rays(ob, 6000, seed=3, dir_kind="i16", scale=1.2)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    want = oracle.trace(vol, ob, pos, d, [1, 1, 1], 300, trace_path=True, round_mode=oracle.ROUND_DEVICE)
    for t in (t_wide, t_keep):
        _assert_same(t.trace_rays_cu(pos, d, [1, 1, 1], 0, 300, trace_paths=True), want, "staging")
        back, _ = t.download_volume()
        assert back.dtype == np.int16 and EQ(back, vol)


def test_dropin_splits_large_batches_across_devices(vrt, oracle):
    """> 32 768 rays through the reference C++ API on the drop-in: on a multi-GPU box the shim keeps a scene per device
    and traces contiguous chunks concurrently (the reference's multi-device model, cu:676-686,820-843); on one GPU it is a
    single chunk.  Either way: same bits as the oracle."""
    from oracle import ref
    if not ref.available("dropin"):
        pytest.skip("oracle/_ref/libvrt_dropin.so not built")
    shape = (40, 36, 34)
    ior, tr = S.random_scene(shape, seed=43, kind="f32", opaque_fraction=0.005)
    pos, d = S.random_rays(shape, 100003, seed=9)
    sc = ref.RefScene(shape, ior, tr, which="dropin")
    got = sc.trace(pos, d, [1, 1, 1], 0, 200)
    ob, _, planes, trc = oracle.prep(shape, ior, tr)
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, [1, 1, 1], 200, round_mode=oracle.ROUND_DEVICE)
    assert EQ(got[0], want[0] + np.uint32(0x10000)) and EQ(got[1], want[1]) and EQ(got[2], want[2])
    # the dynamic piece queue (VRT_SPLIT_PIECES): 1.3 M rays = 3 pieces of >= 2^19 rays, also on a single device
    pos, d = S.random_rays(shape, 1300000, seed=10)
    os.environ["VRT_SPLIT_PIECES"] = "4"
    try:
        got = sc.trace(pos, d, [1, 1, 1], 0, 60)
    finally:
        del os.environ["VRT_SPLIT_PIECES"]
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, [1, 1, 1], 60, round_mode=oracle.ROUND_DEVICE)
    assert EQ(got[0], want[0] + np.uint32(0x10000)) and EQ(got[1], want[1]) and EQ(got[2], want[2])


def test_randomised_configurations(vrt, oracle):
    """Seeded fuzz over shapes (odd / tiny / elongated), element types, invscale, iteration caps, live translucency,
    paths, launch options and volume layout: the CUDA path must equal the oracle bit for bit every time."""
    rng = np.random.default_rng(20261018)
    for case in range(40):
        dim = 3 if case % 5 else 2
        shape = tuple(int(x) for x in rng.integers(5, 40, size=dim))
        volk = "f32" if rng.random() < 0.5 else "i16"
        dirk = "f32" if rng.random() < 0.5 else "i16"
        live = bool(rng.random() < 0.4)
        paths = bool(rng.random() < 0.3)
        bricked = bool(dim == 3 and not paths and rng.random() < 0.3)
        iters = int(rng.choice([1, 2, 3, 17, 64, 257, 1000]))
        n = int(rng.choice([1, 5, 32, 33, 100, 1000, 4097]))
        isc = [float(x) for x in rng.choice([0.5, 1.0, 1.0, 1.5, 2.0], size=dim)]
        ior, tr = S.random_scene(shape, seed=1000 + case, kind="f32" if volk == "f32" else "u32", opaque_fraction=float(rng.choice([0.0, 0.01, 0.2])))
        ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
        if live:
            trc = trc.copy(); trc[trc != 0] -= np.uint32(1 << int(rng.integers(16, 27)))
        vol = oracle.fold(planes, trc)
        pos, d = S.random_rays(ob, n, seed=2000 + case, dir_kind=dirk, scale=float(rng.choice([0.3, 1.0, 1.7])))
        pos = pos - np.uint32(0x10000) + np.uint32(int(rng.integers(0, 0x10000)))
        if n > 4:                       # a few rays that start outside / on the boundary / with zero direction
            pos[0, 0] = np.uint32((ob[0] - 1) << 16); pos[1, 0] = np.uint32(0xFFFF0000); d[2] = 0
        minb = int(rng.choice([0, 0x40000000, 0xF0000000]))
        want = oracle.trace(vol, ob, pos, d, isc, iters, translucency=trc if live else None, min_brightness=minb, trace_path=paths,
                            round_mode=oracle.ROUND_DEVICE)
        t = vrt.TraceRaysCu(ob, planes, trc, bricked=bricked, keep_i16=bool(rng.random() < 0.5))
        if dim == 3:
            t.set_option(vrt.VRT_OPT_KERNEL, int(rng.choice([0, 1, 2, 3])))
            t.set_option(vrt.VRT_OPT_REFILL, int(rng.choice([0, 1, 7, 32])))
            t.set_option(vrt.VRT_OPT_STEPS_PER_POLL, int(rng.choice([1, 3, 32, 500])))
            t.set_option(vrt.VRT_OPT_BLOCK_THREADS, int(rng.choice([32, 64, 128, 256])))
            t.set_option(vrt.VRT_OPT_CHUNK_RAYS, int(rng.choice([0, 0, 7, 1000])))
        got = t.trace_rays_cu(pos, d, isc, minb, iters, trace_paths=paths, live_translucency=live)
        _assert_same(got, want, "case %d shape %s vol %s dir %s live %s paths %s brick %s iters %d n %d" % (case, shape, volk, dirk, live, paths, bricked, iters, n))
        t.close()


def test_short_division_sequence_is_exact_over_its_whole_range(vrt):
    """The fast loop divides 0x42000000p0f by |dir|^2 with a reciprocal + 5 FMA sequence instead of div.rn.f32 (cu:346) and, with
    unit invscale, rounds the step to an integer by adding 1.5 * 2^23 instead of cvt.rni (cu:347); the device self-test compares
    each shortcut with the instruction it replaces for every float it is used for (1.6e9 + 2.5e9 values)."""
    import ctypes as C
    bad = C.c_uint64(12345)
    assert vrt.lib().vrt_selftest_division(0, C.byref(bad)) == 0
    assert bad.value == 0


def test_channel3_edge_values_and_degenerate_directions(vrt, oracle):
    """The default kernel skips channel 3 in cells whose 8 corners all carry its sign bit and re-derives the reason for
    leaving the step loop afterwards.  Volumes whose channel 3 mixes negative values, -0.0, +0.0, denormal / tiny positives and
    NaN in small blobs (so clear cells, opaque cells and mixed cells all occur), gradients with infinities, and rays with zero,
    huge, denormal, infinite and NaN direction components must all come out exactly as the oracle computes them."""
    rng = np.random.default_rng(77)
    shape = (23, 19, 21)
    nvox = int(np.prod(shape))
    vol = np.zeros((nvox, 4), np.float32)
    vol[:, :3] = rng.normal(0, 3000.0, size=(nvox, 3)).astype(np.float32)      # bends a unit direction (65536 internally) noticeably per step
    choices = np.array([-32768.0, -1.0, -0.0, 0.0, 1e-40, 1e-30, 5.0, 32767.0, np.nan], np.float32)
    blob = rng.integers(0, len(choices), size=tuple((np.array(shape) + 3) // 4))
    blob[rng.random(blob.shape) < 0.7] = 0                                      # mostly clear
    c3 = choices[np.kron(blob, np.ones((4, 4, 4), np.int64))[:shape[0], :shape[1], :shape[2]]]
    c3.view(np.uint32)[np.isnan(c3) & (rng.random(shape) < 0.5)] |= np.uint32(0x80000000)   # NaNs of either sign
    vol[:, 3] = c3.reshape(-1)
    vol[rng.integers(0, nvox, 5), rng.integers(0, 3, 5)] = np.inf
    tr = (np.uint32(0xFFFFFFFF) - rng.integers(0, 1 << 25, nvox).astype(np.uint32)).astype(np.uint32)   # only read by the live runs below
    n = 6000
    pos = (rng.random((n, 3)) * (np.array(shape) - 1.0) * 65536.0).astype(np.uint32)
    d = rng.normal(0, 1.2, size=(n, 3)).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-45, 1e-38, 1e30, 3e38, np.inf, -np.inf, np.nan], np.float32)
    rows = rng.integers(0, n, 600)
    d[rows, rng.integers(0, 3, 600)] = special[rng.integers(0, len(special), 600)]
    d[rows[:60]] = 0.0
    for isc, iters in (([1.0, 1.0, 1.0], 300), ([0.5, 2.0, 1.25], 77), ([-1.0, 2.0, -0.5], 60), ([0.0, 0.0, 0.0], 9), ([1e20, 1.0, 1.0], 9),
                       ([1e-20, 1.0, 3e4], 40), ([float('nan'), 1.0, 1.0], 5), ([float('inf'), 1.0, 1.0], 5)):
        want = oracle.trace(vol, shape, pos, d, isc, iters, round_mode=oracle.ROUND_DEVICE)
        t = vrt.TraceRaysCu.from_interleaved(shape, vol, tr)
        for kernel, refill, poll in ((0, 32, 128), (3, 1, 1), (3, 0, 7), (2, 32, 32), (1, 8, 500), (6, 32, 32)):
            t.set_option(vrt.VRT_OPT_KERNEL, kernel); t.set_option(vrt.VRT_OPT_REFILL, refill); t.set_option(vrt.VRT_OPT_STEPS_PER_POLL, poll)
            got = t.trace_rays_cu(pos, d, isc, 0, iters)
            for g, w, nme in zip(got[:3], want[:3], ("end_position", "end_direction", "end_iteration")):
                g = np.ascontiguousarray(g).reshape(-1).view(np.uint32).copy(); w = np.ascontiguousarray(w).reshape(-1).view(np.uint32).copy()
                if nme == "end_direction":      # NaN payloads are unspecified (x86 keeps an operand's payload, the GPU returns the canonical NaN)
                    g[np.isnan(g.view(np.float32))] = 0x7FC00000; w[np.isnan(w.view(np.float32))] = 0x7FC00000
                bad = np.flatnonzero(g != w)
                assert bad.size == 0, "kernel %d refill %d poll %d: %s differs at %s: %s vs %s" % (kernel, refill, poll, nme, bad[:5], g[bad[:5]], w[bad[:5]])
        # the same with the translucency plane live (fast loop of the unit-invscale kernel / generic loop otherwise)
        wantl = oracle.trace(vol, shape, pos, d, isc, iters, translucency=tr, min_brightness=0xC0000000, round_mode=oracle.ROUND_DEVICE)
        for kernel, refill, poll in ((0, 32, 128), (3, 1, 1), (2, 0, 9)):
            t.set_option(vrt.VRT_OPT_KERNEL, kernel); t.set_option(vrt.VRT_OPT_REFILL, refill); t.set_option(vrt.VRT_OPT_STEPS_PER_POLL, poll)
            got = t.trace_rays_cu(pos, d, isc, 0xC0000000, iters, live_translucency=True)
            for g, w, nme in zip(got[:4], wantl[:4], ("end_position", "end_direction", "end_iteration", "remaining_light")):
                g = np.ascontiguousarray(g).reshape(-1).view(np.uint32).copy(); w = np.ascontiguousarray(w).reshape(-1).view(np.uint32).copy()
                if nme == "end_direction":
                    g[np.isnan(g.view(np.float32))] = 0x7FC00000; w[np.isnan(w.view(np.float32))] = 0x7FC00000
                bad = np.flatnonzero(g != w)
                assert bad.size == 0, "live kernel %d refill %d poll %d: %s differs at %s: %s vs %s" % (kernel, refill, poll, nme, bad[:5], g[bad[:5]], w[bad[:5]])
        t.close()


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("live", [False, True])
def test_empty_space_fast_path_is_bit_identical(vrt, oracle, volk, live):
    """KVER 6: in cells whose 8 corners have zero gradient the step is re-used instead of recomputed.  A lens in air (flat
    outside, curved inside), rays with -0.0 and 0.0 direction components, live translucency, and a volume with -0.0 gradients
    (which must NOT count as flat) all have to come out exactly as the oracle computes them."""
    from volumeraytracer_b200 import workloads as W
    size = 48
    ior = W.ior_luneburg(size, 14.0)
    if volk == "i16":
        ior = W.ior_to_u32(ior)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    ob, iorlog, planes, trc = oracle.prep(ior.shape, ior, tr)
    if live:
        trc = trc.copy().reshape(ob); trc[:, :, ::3] -= np.uint32(1 << 22); trc[30:34, 20:26, 20:26] = 0; trc = trc.reshape(-1)
    if volk == "f32":
        planes[1][::7] = np.where(planes[1][::7] == 0, np.float32(-0.0), planes[1][::7])      # -0 is not +0: not flat
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 12000, seed=5, dir_kind="f32" if volk == "f32" else "i16", scale=1.0)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    d[:3000, 1] = 0; d[1000:2000, 2] = 0                       # axis-aligned rays keep exact zeros in the direction
    if volk == "f32":
        d[:500, 1] = np.float32(-0.0)
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.0, 1.0], 700, translucency=trc if live else None, min_brightness=0x40000000,
                        round_mode=oracle.ROUND_DEVICE)
    t = vrt.TraceRaysCu(ob, planes, trc)
    assert t.get_option(vrt.VRT_OPT_KERNEL) == 0
    for kver in (0, 6, 3):
        for refill in (0, 1, 32):
            t.set_option(vrt.VRT_OPT_KERNEL, kver); t.set_option(vrt.VRT_OPT_REFILL, refill)
            got = t.trace_rays_cu(pos, d, [1.0, 1.0, 1.0], 0x40000000, 700, live_translucency=live)
            _assert_same(got, want[:4], "kver %d refill %d" % (kver, refill))
    flat_voxels = np.mean((vol[:, 0] == 0) & (vol[:, 1] == 0) & (vol[:, 2] == 0) & (vol[:, 3] <= 0))
    assert flat_voxels > 0.3
    assert abs(t.get_option(vrt.VRT_INFO_EMPTY_PERMILLE) - 1000 * flat_voxels) <= 1


def test_reference_python_binding_on_the_dropin(vrt, oracle, tmp_path, monkeypatch):
    """The reference's OWN pybind11 module (src/python_binding.cpp, unmodified, built by `make dropin_py`) linked against
    our TraceRaysCu<> drop-in: `cuda_raytrace.cuda_raytrace(...)` -- the reference's Python entry point -- runs on the B200
    marcher and returns what the oracle predicts.  (Tuple slots 2/3 carry end_iteration / remaining_light: the binding's own
    slot mix-up, python_binding.cpp:38-45.)"""
    import glob
    import importlib.util
    import os
    from oracle import ref
    so = glob.glob(os.path.join(os.path.dirname(ref.LIB_PATH), "cuda_raytrace*.so"))
    if not so or not ref.available("dropin"):
        pytest.skip("oracle/_ref/cuda_raytrace*.so not built")
    spec = importlib.util.spec_from_file_location("cuda_raytrace", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.chdir(tmp_path)                       # the binding writes debug_raytrace_instance into the CWD on every call
    inp = S.scaling_test_inputs("u32")
    res = mod.cuda_raytrace(inp["bounds"], inp["ior"].tolist(), inp["translucency"].tolist(), inp["pos"].ravel().tolist(),
                            inp["dir"].ravel().tolist(), inp["invscale"], 0, inp["iterations"], False)
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, inp["invscale"], inp["iterations"], round_mode=oracle.ROUND_DEVICE)
    assert np.array_equal(np.array(res[0], np.uint32).reshape(-1, 3), want[0] + np.uint32(0x10000))
    assert np.array_equal(np.array(res[1], np.int16).reshape(-1, 3), want[1])
    assert list(res[2]) == want[2].tolist() == [46734, 46623]
    assert list(res[3]) == [0xFFFFFFFF, 0xFFFFFFFF]
    # a larger call: 3000 random rays
    shape = (24, 20, 28)
    ior, tr = S.random_scene(shape, seed=13, kind="u32", opaque_fraction=0.01)
    pos, d = S.random_rays(shape, 3000, seed=5, dir_kind="i16")
    res = mod.cuda_raytrace(list(shape), ior.ravel().tolist(), tr.ravel().tolist(), pos.ravel().tolist(), d.ravel().tolist(), [1.0, 0.75, 1.5], 0, 400, False)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    p2, d2 = oracle.normalise(shape, ior, pos, d)
    want = oracle.trace(oracle.fold(planes, trc), ob, p2, d2, [1.0, 0.75, 1.5], 400, round_mode=oracle.ROUND_DEVICE)
    assert np.array_equal(np.array(res[0], np.uint32).reshape(-1, 3), want[0] + np.uint32(0x10000))
    assert np.array_equal(np.array(res[1], np.int16).reshape(-1, 3), want[1]) and list(res[2]) == want[2].tolist()


@pytest.mark.parametrize("volk,dirk,live", [("f32", "f32", False), ("i16", "i16", True), ("f32", "i16", True), ("i16", "f32", False)])
def test_region_mode_is_bit_identical(vrt, oracle, volk, dirk, live):
    """VRT_OPT_REGION_LOG2: rays sorted by region and marched region by region, suspended and resumed across rounds --
    every step is still the reference's step, so the outputs equal the oracle's whatever the region size / round count."""
    shape = (70, 45, 81)
    ior, tr = S.random_scene(shape, seed=23, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.002)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    if live:
        trc = trc.copy(); trc[trc != 0] -= np.uint32(1 << 22)
    vol = oracle.fold(planes, trc)
    pos, d = S.random_rays(ob, 30000, seed=4, dir_kind=dirk, scale=1.0)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    pos[0] = [0xFFFF0000, 0x20000, 0x20000]; pos[1, 2] = np.uint32((ob[2] - 1) << 16)            # start outside / on the far face
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.0, 1.0], 900, translucency=trc if live else None, min_brightness=0x20000000,
                        round_mode=oracle.ROUND_DEVICE)
    t = vrt.TraceRaysCu(ob, planes, trc, keep_i16=(dirk == "f32"))
    for log2, rounds, refill in ((5, 24, 1), (5, 2, 32), (6, 1, 8), (7, 24, 32)):
        t.set_option(vrt.VRT_OPT_REGION_LOG2, log2); t.set_option(vrt.VRT_OPT_REGION_ROUNDS, rounds); t.set_option(vrt.VRT_OPT_REFILL, refill)
        got = t.trace_rays_cu(pos, d, [1.0, 1.0, 1.0], 0x20000000, 900, live_translucency=live)
        _assert_same(got, want[:4], "region log2=%d rounds=%d refill=%d" % (log2, rounds, refill))
    assert len(np.unique(want[2])) > 50
    t.close()
    tb = vrt.TraceRaysCu(ob, planes, trc, bricked=True)          # region mode over the 2x2x2-brick layout
    tb.set_option(vrt.VRT_OPT_REGION_LOG2, 5)
    _assert_same(tb.trace_rays_cu(pos, d, [1.0, 1.0, 1.0], 0x20000000, 900, live_translucency=live), want[:4], "region mode, bricked")
    tb.close()
