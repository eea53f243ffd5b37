"""Round-2 GPU tests (pytest -m gpu, through the C ABI):
  * VRT_TRACE_ROUND_HOST -- the marcher rounding like the reference's CPU build (tuple_math.h:270-278): bit-exact against the
    oracle's ROUND_HOST mode, against the UNMODIFIED reference CPU build where it is present, and against the reference's
    scaling_test known answers (cuda_volume_raytracer_test.h:4-75; SURVEY.md section 4: 46734/46623 and 46718/46656 steps);
  * multi-GPU replication in the C++ product path: vrt_scene_replicate (NVLink peer chain, one process) and vrt_comm_* /
    vrt_scene_broadcast (NCCL, one process per GPU) give bit-identical scenes and results;
  * the iteration-cap flag (vrt_trace_cap_hit), vrt_scene_info / vrt_scene_storage_info, the normalise edge case."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu
EQ = np.array_equal
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vrt():
    import volumeraytracer_b200 as v
    v.lib()
    return v


def _mk(vrt, oracle, shape, seed, volk, opaque=0.01, absorb=False, **kw):
    ior, tr = S.random_scene(shape, seed=seed, kind="f32" if volk == "f32" else "u32", opaque_fraction=opaque)
    ob, iorlog, planes, trc = oracle.prep(shape, ior, tr)
    if absorb:
        trc = trc.copy()
        trc[trc != 0] -= np.uint32(1 << 24)
    vol = oracle.fold(planes, trc)
    tracer = vrt.TraceRaysCu(ob, planes, trc, **kw)
    return ob, planes, trc, vol, tracer


def _assert_same(got, want, what=""):
    names = ["end_position", "end_direction", "end_iteration", "remaining_light", "path"]
    for g, w, nme in zip(got, want, names):
        if w is None:
            continue
        assert EQ(g, w), "%s %s differs (%d of %d)" % (what, nme, int(np.sum(g != w)), g.size)


def _gpu_count(vrt):
    import ctypes as C
    n = C.c_int(0)
    vrt.lib().vrt_device_count(C.byref(n))
    return n.value


# ---------------------------------------------------------------------------------------------------
# VRT_TRACE_ROUND_HOST

@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
@pytest.mark.parametrize("live", [False, True])
def test_round_host_all_type_combinations_3d(vrt, oracle, volk, dirk, live):
    ob, planes, trc, vol, t = _mk(vrt, oracle, (30, 26, 34), 41, volk, absorb=live)
    pos, d = S.random_rays(ob, 12000, seed=9, dir_kind=dirk, scale=1.1)
    pos = pos - np.uint32(0x10000) + np.uint32(0x3000)
    isc = [1.0, 0.9, 1.2]
    minb = 0x40000000 if live else 0
    want = oracle.trace(vol, ob, pos, d, isc, 400, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_HOST)
    dev = oracle.trace(vol, ob, pos, d, isc, 400, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_DEVICE)
    assert not EQ(want[0], dev[0]), "the two rounding modes should differ somewhere on 12000 rays"
    for refill in (32, 0, 1):
        t.set_option(vrt.VRT_OPT_REFILL, refill)
        got = t.trace_rays_cu(pos, d, isc, minb, 400, live_translucency=live, round_host=True)
        _assert_same(got, want[:4], "ROUND_HOST %s/%s live=%s refill=%d" % (volk, dirk, live, refill))
    wantp = oracle.trace(vol, ob, pos[:400], d[:400], isc, 90, translucency=trc if live else None, min_brightness=minb, trace_path=True,
                         round_mode=oracle.ROUND_HOST)
    gotp = t.trace_rays_cu(pos[:400], d[:400], isc, minb, 90, trace_paths=True, live_translucency=live, round_host=True)
    _assert_same(gotp, wantp, "ROUND_HOST paths %s/%s" % (volk, dirk))
    # the default (device rounding) is untouched by the flag's existence
    _assert_same(t.trace_rays_cu(pos, d, isc, minb, 400, live_translucency=live), dev[:4], "default rounding")
    t.close()


@pytest.mark.parametrize("volk", ["f32", "i16"])
@pytest.mark.parametrize("dirk", ["f32", "i16"])
def test_round_host_2d(vrt, oracle, volk, dirk):
    """2-D: the CPU build also contracts the lerps differently per channel (oracle sample2, read from g++'s output)."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (70, 55), 43, volk)
    pos, d = S.random_rays(ob, 9000, seed=11, dir_kind=dirk)
    pos = pos - np.uint32(0x10000) + np.uint32(0x5000)
    isc = [1.0, 1.1]
    want = oracle.trace(vol, ob, pos, d, isc, 500, round_mode=oracle.ROUND_HOST)
    got = t.trace_rays_cu(pos, d, isc, 0, 500, round_host=True)
    _assert_same(got, want[:4], "ROUND_HOST 2-D %s/%s" % (volk, dirk))
    wantp = oracle.trace(vol, ob, pos[:300], d[:300], isc, 60, trace_path=True, round_mode=oracle.ROUND_HOST)
    _assert_same(t.trace_rays_cu(pos[:300], d[:300], isc, 0, 60, trace_paths=True, round_host=True), wantp, "ROUND_HOST 2-D paths")
    t.close()


@pytest.mark.parametrize("kind", ["u32", "f32"])
def test_round_host_scaling_test_known_answers(vrt, oracle, kind):
    """The reference's scaling_test: with host rounding the GPU reproduces the reference CPU build's known answers EXACTLY
    (step counts 46734/46623 resp. 46718/46656, end positions, int16 end directions)."""
    inp = S.scaling_test_inputs(kind)
    ob, iorlog, planes, trc = oracle.prep(inp["bounds"], inp["ior"], inp["translucency"])
    vol = oracle.fold(planes, trc)
    p2, d2 = oracle.normalise(inp["bounds"], inp["ior"], inp["pos"], inp["dir"])
    t = vrt.TraceRaysCu(ob, planes, trc)
    got = t.trace_rays_cu(p2, d2, inp["invscale"], 0, inp["iterations"], round_host=True)
    want = oracle.trace(vol, ob, p2, d2, inp["invscale"], inp["iterations"], round_mode=oracle.ROUND_HOST)
    _assert_same(got, want[:4], "scaling_test ROUND_HOST")
    known = S.SCALING_KNOWN[kind]
    assert got[2].tolist() == known["eit"]
    assert ((got[0].ravel() + np.uint32(0x10000)) == np.array(known["epos"], dtype=np.uint32)).all()
    if kind == "u32":
        assert got[1].ravel().tolist() == known["edir"]
    else:
        assert np.allclose(got[1].ravel(), known["edir"], rtol=0, atol=2e-4)
    t.close()


def _config_inputs(name):
    from volumeraytracer_b200 import workloads as W
    if name == "c2":
        size = 64
        ior = W.ior_luneburg(size, 25.0); tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
        pos, d = W.rays_parallel_x(96, 96, 8.0, 55.0, x0=2.0)
        return ior, tr, pos, d, 1024, False, 0
    if name == "c3":
        size = 96
        ior = W.ior_sines(size, period=32.0); tr = W.translucency_c3(size)
        absorb = (np.uint64(0xFFFFFFFF) - tr.astype(np.uint64)) * np.uint64(8)
        tr2 = (np.uint64(0xFFFFFFFF) - np.minimum(absorb, np.uint64(0xFFFFFFFF))).astype(np.uint32)
        tr2[tr == 0] = 0
        pos, d = W.rays_parallel_x(128, 128, 4.0, size - 5.0, x0=2.0)
        return ior, tr2, pos, d, 4096, True, 0x40000000
    size = 64
    ior = W.solve_harmonic(size, inner_radius=8.0, sweeps=60); tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    pos, d = W.rays_random(30000, 8.0, size - 9.0, 0x5EED0004)
    return ior, tr, pos, d, 4096, False, 0


@pytest.mark.parametrize("name", ["c2", "c3", "c4"])
def test_round_host_configs_equal_reference_cpu(vrt, oracle, name):
    """Configs 2-4 (test sizes): with VRT_TRACE_ROUND_HOST the GPU equals the oracle's ROUND_HOST mode bit for bit -- zero
    step-count mismatches -- and, where oracle/_ref is present, the UNMODIFIED reference CPU marcher itself."""
    from oracle import ref
    ior, tr, pos, d, iters, live, minb = _config_inputs(name)
    sc = vrt.RaytraceScene(ior.shape, ior, tr)
    co = sc._calculation_object
    vol, trc = co.download_volume()
    ob = co._output_sizes
    p2, d2 = oracle.normalise(ior.shape, ior, pos, d)
    got = co.trace_rays_cu(p2, d2, [1, 1, 1], minb, iters, live_translucency=live, round_host=True)
    want = oracle.trace(vol, ob, p2, d2, [1, 1, 1], iters, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_HOST)
    _assert_same(got, want[:4], "%s ROUND_HOST vs oracle" % name)
    if ref.available():
        cpu = ref.trace_live(vol, trc if live else None, ob, [1, 1, 1], p2, d2, iters, minb)
        _assert_same(got, cpu[:4], "%s ROUND_HOST vs the reference CPU build" % name)
    sc.close()


def test_round_host_is_rejected_on_brick_scenes(vrt, oracle):
    ob, planes, trc, vol, t = _mk(vrt, oracle, (20, 20, 20), 3, "f32", bricked=True)
    pos, d = S.random_rays(ob, 100, seed=1)
    with pytest.raises(vrt.VrtError) as e:
        t.trace_rays_cu(pos - np.uint32(0x10000), d, [1, 1, 1], 0, 50, round_host=True)
    assert e.value.code == 4
    t.close()


# ---------------------------------------------------------------------------------------------------
# replication

def _trace_both(vrt, a, b, ob, live):
    pos, d = S.random_rays(ob, 20000, seed=77)
    pos = pos - np.uint32(0x10000) + np.uint32(0x1234)
    ra = a.trace_rays_cu(pos, d, [1.0, 1.1, 0.9], 0x40000000 if live else 0, 300, live_translucency=live)
    rb = b.trace_rays_cu(pos, d, [1.0, 1.1, 0.9], 0x40000000 if live else 0, 300, live_translucency=live)
    _assert_same(rb, ra[:4], "replica vs source")


@pytest.mark.parametrize("kw", [{}, {"bricked": True}, {"keep_i16": True}])
@pytest.mark.parametrize("volk", ["f32", "i16"])
def test_scene_replicate_is_bit_identical(vrt, oracle, volk, kw):
    """vrt_scene_replicate: the replica (same device on a 1-GPU box; device 1, and a 0 -> 1 -> 0 chain, when there are two) holds
    the same bits (vrt_scene_download on both) and traces the same results; storage description and options carry over."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (33, 29, 37), 51, volk, absorb=True, **kw)
    t.set_option(vrt.VRT_OPT_STEPS_PER_POLL, 64)
    ndev = _gpu_count(vrt)
    targets = [0] if ndev < 2 else [1, 0]
    reps, secs = t.replicate(targets)
    assert len(reps) == len(targets) and secs > 0
    v0, tr0 = t.download_volume()
    assert EQ(np.asarray(v0).reshape(-1), np.asarray(vol).reshape(-1)) and EQ(tr0, trc)
    for r, dev in zip(reps, targets):
        assert r.device == dev and r._output_sizes == t._output_sizes and r.diff_dtype == t.diff_dtype
        assert r.storage_info()[:3] == t.storage_info()[:3]
        assert r.get_option(vrt.VRT_OPT_STEPS_PER_POLL) == 64
        v1, tr1 = r.download_volume()
        assert EQ(np.asarray(v1).reshape(-1), np.asarray(v0).reshape(-1)) and EQ(tr1, tr0)
        _trace_both(vrt, t, r, ob, live=True)
    t.close()                                       # replicas own their buffers: they outlive the source
    _trace_both(vrt, reps[0], reps[-1], ob, live=False)
    for r in reps:
        r.close()


def test_replicated_from_ior_scene_keeps_the_normalise_step(vrt, oracle):
    """A scene made by the GPU scene prep keeps ior for the ray normalisation (f2); the replica gets a copy of it."""
    import torch
    shape = (40, 36, 44)
    ior, tr = S.random_scene(shape, seed=5, kind="f32")
    src = vrt.TraceRaysCu.from_ior(shape, ior, tr)
    ndev = _gpu_count(vrt)
    reps, _ = src.replicate([ndev - 1])
    pos, d = S.random_rays(shape, 5000, seed=2)
    want = oracle.normalise(shape, ior, pos, d)
    dev = torch.device("cuda", reps[0].device)
    tp = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); td = torch.from_numpy(d.reshape(-1)).to(dev)
    reps[0].normalise_rays_device(tp, td)
    assert EQ(tp.cpu().numpy().view(np.uint32).reshape(-1, 3), want[0]) and EQ(td.cpu().numpy().reshape(-1, 3), want[1])
    src.close(); reps[0].close()


def test_nccl_comm_single_rank(vrt, oracle):
    """world = 1: exercises the dlopen of libnccl, communicator set-up/warm-up and the broadcast code path on one GPU."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (24, 22, 26), 61, "f32")
    comm = vrt.Comm(0, 0, 1, lambda uid: uid)
    s2, secs = comm.broadcast_scene(t, root=0)
    assert s2 is t and secs >= 0
    comm.close()
    t.close()


def test_nccl_scene_broadcast_two_processes(vrt, oracle, tmp_path):
    """One process per GPU (the bench's launch model): rank 0 builds the scene, vrt_scene_broadcast replicates it with
    in-place ncclBroadcasts; both ranks download identical bits and trace identical results."""
    if _gpu_count(vrt) < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ROOT, "tests", "_nccl_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    a = np.load(os.path.join(str(tmp_path), "rank0.npz")); b = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    for k in ("vol", "tr", "epos", "edir", "eit", "light"):
        assert EQ(a[k], b[k]), k
    assert float(b["secs"]) > 0


# ---------------------------------------------------------------------------------------------------
# small ABI additions

def test_cap_flag_replaces_the_host_scan(vrt, oracle):
    """cu:507-515 warns when any end_iteration == iterations; the marcher reports that through a flag."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (40, 40, 40), 7, "f32", opaque=0.0)
    pos, d = S.random_rays(ob, 300000, seed=3)
    pos = pos - np.uint32(0x10000)
    for n in (1000, 300000):                                 # small-batch path and the chunk pipeline
        got = t.trace_rays_cu(pos[:n], d[:n], [1, 1, 1], 0, 5)               # nobody gets out in 5 steps
        assert (got[2] == 5).any() and t.cap_hit() == 1
        got = t.trace_rays_cu(pos[:n], d[:n], [1, 1, 1], 0, 100000)          # everybody gets out
        assert not (got[2] == 100000).any() and t.cap_hit() == 0
    t.set_option(vrt.VRT_OPT_REFILL, 0)
    got = t.trace_rays_cu(pos, d, [1, 1, 1], 0, 5)
    assert t.cap_hit() == 1
    t.close()
    ob, planes, trc, vol, t2 = _mk(vrt, oracle, (60, 50), 7, "f32", opaque=0.0)
    pos, d = S.random_rays(ob, 50000, seed=3)
    t2.trace_rays_cu(pos - np.uint32(0x10000), d, [1, 1], 0, 4)
    assert t2.cap_hit() == 1
    t2.close()


def test_scene_info_hides_non_reference_storage(vrt, oracle):
    """vrt_scene_info must not hand out a staging pointer that does not hold the reference layout / element type."""
    ob, planes, trc, vol, wide = _mk(vrt, oracle, (20, 22, 24), 9, "i16")               # int16 scene, widened to float on the device
    assert wide.volume_ptr is None and wide.volume_bytes == vol.size * 2
    dt, flags, nbytes, ptr = wide.storage_info()
    assert dt == vrt.VRT_F32 and nbytes == vol.size * 4 and ptr
    wide.close()
    ob, planes, trc, vol, kept = _mk(vrt, oracle, (20, 22, 24), 9, "i16", keep_i16=True)
    assert kept.volume_ptr and kept.storage_info()[0] == vrt.VRT_I16 and kept.storage_info()[2] == vol.size * 2
    kept.close()
    ob, planes, trc, vol, br = _mk(vrt, oracle, (20, 22, 24), 9, "f32", bricked=True)
    assert br.volume_ptr is None and br.storage_info()[1] & 2
    br.close()
    ob, planes, trc, vol, lin = _mk(vrt, oracle, (20, 22, 24), 9, "f32")
    assert lin.volume_ptr and lin.storage_info()[2] == vol.size * 4 == lin.volume_bytes
    lin.close()


@pytest.mark.parametrize("kind", ["f32", "u32"])
def test_normalise_accepts_rays_in_the_last_half_voxel(vrt, oracle, kind):
    """a ray at bounds*0x10000 - 2 passes the reference's range check (image_util.cpp:686) and then has its upper
    interpolation corner outside the volume; the GPU step clamps that corner instead of reading past the buffer, and the marcher
    retires the ray on its first bounds test (end_iteration 1)."""
    import torch
    shape = (18, 16, 20)
    ior, tr = S.random_scene(shape, seed=4, kind=kind)
    sc = vrt.RaytraceScene(shape, ior, tr)
    pos = np.array([[s * 0x10000 - 2 for s in shape], [shape[0] * 0x10000 - 2, 0x20000, 0x20000], [0x20000, 0x20000, shape[2] * 0x10000 - 2]], np.uint32)
    d = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    if kind == "u32":
        d = (d * 256).astype(np.int16)
    for _ in range(3):                                      # a fault would be sticky: repeat and then use the context again
        got = sc.trace_rays(pos, d, [1, 1, 1], 0, 100)
        assert got[2].tolist() == [1, 1, 1]
        assert EQ(got[0], pos)                              # not moved: -0x10000 in, +0x10000 out
    torch.cuda.synchronize()
    pos2, d2 = S.random_rays(shape, 2000, seed=8, dir_kind="f32" if kind == "f32" else "i16")
    sc.trace_rays(pos2, d2, [1, 1, 1], 0, 100)
    sc.close()


@pytest.mark.parametrize("volk,dirk,live", [("f32", "f32", False), ("i16", "i16", False), ("f32", "i16", True), ("i16", "f32", True)])
def test_all_clear_scene_uses_the_kernel_without_the_cell_test(vrt, oracle, volk, dirk, live):
    """A scene in which no voxel can make a sample opaque (channel 3 negative everywhere) is marched by KVER 11, the default kernel
    without the per-cell clear test; one possibly opaque voxel (channel 3 == +0: tr = 0x7FFFFFFF) switches it off.  Same bits as the
    kernel with the test and as the oracle."""
    for opaque_voxels in (0, 1):
        ior, tr = S.random_scene((40, 36, 44), seed=31, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.0)
        ob, iorlog, planes, trc = oracle.prep((40, 36, 44), ior, tr)
        trc = trc.copy()
        if live:
            trc[trc != 0] -= np.uint32(1 << 24)
        if opaque_voxels:
            trc.reshape(-1)[trc.size // 2 + 7] = np.uint32(0x7FFFFFFF)          # (0x7FFFFFFF - tr) / 0x10000 == 0: sign bit clear
        vol = oracle.fold(planes, trc)
        t = vrt.TraceRaysCu(ob, planes, trc, keep_i16=(volk == "i16" and live))
        assert t.get_option(vrt.VRT_INFO_ALL_CLEAR) == (0 if opaque_voxels else 1)
        pos, d = S.random_rays(ob, 30000, seed=9, dir_kind=dirk, scale=1.1)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4321)
        isc = [1.0, 1.0, 1.0]
        minb = 0x40000000 if live else 0
        want = oracle.trace(vol, ob, pos, d, isc, 500, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_DEVICE)
        for allclear in (1, 0):
            t.set_option(vrt.VRT_OPT_ALL_CLEAR_KERNEL, allclear)
            got = t.trace_rays_cu(pos, d, isc, minb, 500, live_translucency=live)
            _assert_same(got, want[:4], "all-clear kernel option %d, %d opaque voxel(s), %s/%s" % (allclear, opaque_voxels, volk, dirk))
        t.close()


def test_host_call_chunk_schedule_on_a_large_batch(vrt, oracle):
    """vrt_trace cuts a large batch into whole waves of the persistent grid with a small first and last chunk (head and tail of the
    copy pipeline).  5 M + 17 rays is past the point where that schedule starts: every ray must come back, in place, with the bits
    of the device-resident call and of the oracle."""
    import torch
    ob, planes, trc, vol, t = _mk(vrt, oracle, (40, 36, 44), 23, "f32")
    n = 5_000_017
    pos, d = S.random_rays(ob, 4096, seed=3, dir_kind="f32", scale=1.1)
    reps = -(-n // pos.shape[0])
    pos = np.ascontiguousarray(np.tile(pos - np.uint32(0x10000) + np.uint32(0x777), (reps, 1))[:n])
    d = np.ascontiguousarray(np.tile(d, (reps, 1))[:n])
    pos[:, 2] += (np.arange(n, dtype=np.uint32) % np.uint32(977))            # not a periodic batch: a misplaced chunk would show
    isc = [1.0, 1.0, 1.0]
    dev = torch.device("cuda", 0)
    tp = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); td = torch.from_numpy(d.reshape(-1)).to(dev)
    want = [o.cpu().numpy() for o in t.trace_device(tp, td, isc, 0, 40)]
    hp = pos.reshape(-1).copy(); hd = d.reshape(-1).copy(); hi = np.empty(n, np.uint32); hl = np.empty(n, np.uint32)
    t.trace_host_buffers(hp, hd, isc, 0, 40, hp, hd, hi, hl)                 # in place, pageable
    assert EQ(hp, want[0].view(np.uint32)) and EQ(hd, want[1]) and EQ(hi, want[2].view(np.uint32))
    sel = np.arange(0, n, 997)
    ref = oracle.trace(vol, ob, pos[sel], d[sel], isc, 40, round_mode=oracle.ROUND_DEVICE)
    assert EQ(hp.reshape(-1, 3)[sel], ref[0]) and EQ(hi[sel], ref[2])
    t.close()


# ---------------------------------------------------------------------------------------------------
# wavefront marcher (incoherent batches) and the device-side coherence probe

@pytest.mark.parametrize("volk,dirk,live", [("f32", "f32", False), ("i16", "i16", True), ("f32", "i16", True), ("i16", "f32", False)])
def test_wavefront_mode_is_bit_identical(vrt, oracle, volk, dirk, live):
    """VRT_OPT_WAVE_LOG2: rays bucketed by brick inside one cooperative launch, marched brick by brick with CTA-wide compaction,
    suspended and resumed across rounds -- every output bit equals the single-launch marcher's and the oracle's."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (44, 38, 50), 91, volk, absorb=live, keep_i16=(volk == "i16" and live))
    pos, d = S.random_rays(ob, 70000, seed=13, dir_kind=dirk, scale=1.2)
    pos = pos - np.uint32(0x10000) + np.uint32(0x2222)
    pos[:50] = np.uint32(0xFFF00000)                              # rays that start outside the volume
    isc = [1.0, 1.2, 0.85]
    minb = 0x40000000 if live else 0
    want = oracle.trace(vol, ob, pos, d, isc, 700, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_DEVICE)
    for k, margin, check, tail in ((3, 2, 16, 20), (4, 0, 5, 0), (3, 1, 64, 500), (5, 3, 16, 1000), (3, 8, 1, 20)):
        t.set_option(vrt.VRT_OPT_WAVE_LOG2, k); t.set_option(vrt.VRT_OPT_WAVE_MARGIN, margin)
        t.set_option(vrt.VRT_OPT_WAVE_CHECK, check); t.set_option(vrt.VRT_OPT_WAVE_TAIL_PERMILLE, tail)
        got = t.trace_rays_cu(pos, d, isc, minb, 700, live_translucency=live)
        _assert_same(got, want[:4], "wavefront k=%d margin=%d check=%d tail=%d %s/%s" % (k, margin, check, tail, volk, dirk))
        assert t.get_option(vrt.VRT_INFO_WAVE_ROUNDS) >= 1
    # the cap: every ray that is still inside after `iterations` steps reports `iterations`, and the flag says so
    t.set_option(vrt.VRT_OPT_WAVE_LOG2, 3); t.set_option(vrt.VRT_OPT_WAVE_MARGIN, 2); t.set_option(vrt.VRT_OPT_WAVE_CHECK, 16)
    want5 = oracle.trace(vol, ob, pos, d, isc, 5, translucency=trc if live else None, min_brightness=minb, round_mode=oracle.ROUND_DEVICE)
    got5 = t.trace_rays_cu(pos, d, isc, minb, 5, live_translucency=live)
    _assert_same(got5, want5[:4], "wavefront cap")
    assert t.cap_hit() == 1
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,volk", [((30, 28, 33), "f32"), ((21, 70, 9), "f32"), ((30, 28, 33), "i16"), ((26, 24, 40), "i16")])
def test_wavefront_odd_extents_and_extreme_directions(vrt, oracle, shape, volk):
    """Odd extents, thin volumes, and directions outside the range of the exact short division / add-a-constant rounding (zero, tiny,
    huge, NaN, inf: the generic div.rn.f32 / cvt.rni path of the wavefront marcher)."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, shape, 17, volk, keep_i16=(volk == "i16"))
    pos, d = S.random_rays(ob, 30000, seed=5, dir_kind="f32", scale=1.1)
    pos = pos - np.uint32(0x10000) + np.uint32(0x1234)
    d[:200] = 0.0
    d[200:400] *= np.float32(1e-30)
    d[400:600] *= np.float32(1e25)
    d[600:620] = np.float32(np.nan)
    d[620:640, 1] = np.float32(np.inf)
    want = oracle.trace(vol, ob, pos, d, [1.0, 1.0, 1.0], 300, round_mode=oracle.ROUND_DEVICE)
    for isc in ([1.0, 1.0, 1.0], [3.0, 0.5, 1e-3]):
        want = oracle.trace(vol, ob, pos, d, isc, 300, round_mode=oracle.ROUND_DEVICE)
        for k in (3, 4):
            t.set_option(vrt.VRT_OPT_WAVE_LOG2, k); t.set_option(vrt.VRT_OPT_WAVE_MARGIN, 2)
            got = list(t.trace_rays_cu(pos, d, isc, 0, 300))
            w = [x.copy() for x in want[:4]]
            # NaN payloads are unspecified (x86 keeps an operand's payload, the GPU returns the canonical NaN)
            gd = got[1].view(np.uint32).copy(); wd = w[1].view(np.uint32).copy()
            gd[np.isnan(got[1])] = 0x7FC00000; wd[np.isnan(w[1])] = 0x7FC00000
            got[1], w[1] = gd, wd
            _assert_same(got, w, "wavefront odd extents k=%d isc=%s %s %s" % (k, isc, shape, volk))
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("volk,dirk,keep", [("f32", "f32", False), ("f32", "i16", False), ("i16", "f32", True), ("i16", "i16", True), ("i16", "f32", False)])
def test_wavefront_all_clear_variant_is_bit_identical(vrt, oracle, volk, dirk, keep):
    """A scene without any possibly opaque voxel (VRT_INFO_ALL_CLEAR) is marched by the wavefront kernel's variant that keeps no
    channel 3 in its cell cache (4 resident CTAs per SM): same bits as the oracle, as the generic wavefront kernel
    (VRT_OPT_ALL_CLEAR_KERNEL 0) and as the single launch -- incl. rays outside the volume, degenerate directions, the cap flag."""
    ob, planes, trc, vol, t = _mk(vrt, oracle, (41, 36, 52), 23, volk, opaque=0.0, keep_i16=keep)
    assert t.get_option(vrt.VRT_INFO_ALL_CLEAR) == 1
    pos, d = S.random_rays(ob, 60000, seed=29, dir_kind=dirk, scale=1.15)
    pos = pos - np.uint32(0x10000) + np.uint32(0x0F0F)
    pos[:40] = np.uint32(0xFFF00000)                              # rays that start outside the volume
    if dirk == "f32":
        d[100:200] = 0.0
        d[200:300] *= np.float32(1e-30)
        d[300:400] *= np.float32(1e25)
    for isc, iters in (([1.0, 1.0, 1.0], 600), ([0.9, 1.3, 1.0], 600), ([1.0, 1.0, 1.0], 7)):
        want = oracle.trace(vol, ob, pos, d, isc, iters, round_mode=oracle.ROUND_DEVICE)
        for k, margin, check, allclear in ((3, 2, 16, 1), (4, 0, 5, 1), (5, 8, 16, 1), (3, 2, 16, 0), (-1, 2, 16, 1)):
            t.set_option(vrt.VRT_OPT_WAVE_LOG2, k); t.set_option(vrt.VRT_OPT_WAVE_MARGIN, margin)
            t.set_option(vrt.VRT_OPT_WAVE_CHECK, check); t.set_option(vrt.VRT_OPT_ALL_CLEAR_KERNEL, allclear)
            got = t.trace_rays_cu(pos, d, isc, 0, iters)
            _assert_same(got, want[:4], "all-clear wavefront k=%d margin=%d check=%d allclear=%d isc=%s %s/%s" % (k, margin, check, allclear, isc, volk, dirk))
            if k > 0:
                assert t.get_option(vrt.VRT_INFO_WAVE_ROUNDS) >= 1
        if iters == 7:
            assert t.cap_hit() == 1
    t.close()


def test_device_probe_picks_the_marcher_without_a_sync(vrt, oracle):
    """vrt_trace_device on a large volume: a probe KERNEL looks at the device-resident rays and gates the single-launch and the
    wavefront marcher; a shuffled batch runs the wavefront kernel (rounds > 0), a coherent bundle does not (rounds == 0); the
    results are the same bits either way and equal the oracle's on a subsample."""
    import torch
    from volumeraytracer_b200 import workloads as W
    size = 200                                                   # 198^3 x 16 B = 124 MB: does not fit the 96 MB the probe asks for
    ior = W.ior_sines(size, base=1.3, amp=0.08, period=40.0)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    sc = vrt.TraceRaysCu.from_ior(ior.shape, ior, tr)
    dev = torch.device("cuda", 0)
    n = 1 << 18
    pos, d = W.rays_random(n, 4.0, size - 5.0, 0xABCDE)
    tp = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); td = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tp, td)
    out = [o.cpu().numpy() for o in sc.trace_device(tp, td, [1, 1, 1], 0, 600)]
    assert sc.get_option(vrt.VRT_INFO_WAVE_ROUNDS) > 1, "the probe should have chosen the wavefront marcher for a shuffled batch"
    sc.set_option(vrt.VRT_OPT_WAVE_LOG2, -1)
    ref_out = [o.cpu().numpy() for o in sc.trace_device(tp, td, [1, 1, 1], 0, 600)]
    for a, b in zip(out, ref_out):
        assert EQ(a, b)
    vol, trc = sc.download_volume()
    sel = np.arange(0, n, 64)
    p_h = tp.cpu().numpy().view(np.uint32).reshape(-1, 3)[sel]; d_h = td.cpu().numpy().reshape(-1, 3)[sel]
    want = oracle.trace(vol, sc._output_sizes, p_h, d_h, [1, 1, 1], 600, round_mode=oracle.ROUND_DEVICE)
    assert EQ(out[0].view(np.uint32).reshape(-1, 3)[sel], want[0]) and EQ(out[1].reshape(-1, 3)[sel], want[1]) and EQ(out[2].view(np.uint32)[sel], want[2])
    # a coherent bundle of the same size: the gate keeps the single-launch marcher
    sc.set_option(vrt.VRT_OPT_WAVE_LOG2, 0)
    pos2, d2 = W.rays_parallel_x(512, 512, 4.0, size - 5.0, x0=2.0)
    tp2 = torch.from_numpy(pos2.view(np.int32).reshape(-1)).to(dev); td2 = torch.from_numpy(d2.reshape(-1)).to(dev)
    sc.normalise_rays_device(tp2, td2)
    out2 = [o.cpu().numpy() for o in sc.trace_device(tp2, td2, [1, 1, 1], 0, 600)]
    assert sc.get_option(vrt.VRT_INFO_WAVE_ROUNDS) == 0
    sel2 = np.arange(0, pos2.shape[0], 64)
    p_h = tp2.cpu().numpy().view(np.uint32).reshape(-1, 3)[sel2]; d_h = td2.cpu().numpy().reshape(-1, 3)[sel2]
    want2 = oracle.trace(vol, sc._output_sizes, p_h, d_h, [1, 1, 1], 600, round_mode=oracle.ROUND_DEVICE)
    assert EQ(out2[0].view(np.uint32).reshape(-1, 3)[sel2], want2[0]) and EQ(out2[2].view(np.uint32)[sel2], want2[2])
    # and the host call makes the same choice with its host-side probe
    got_h = sc.trace_rays_cu(tp.cpu().numpy().view(np.uint32).reshape(-1, 3), td.cpu().numpy().reshape(-1, 3), [1, 1, 1], 0, 600)
    assert sc.get_option(vrt.VRT_INFO_WAVE_ROUNDS) > 1
    assert EQ(got_h[0].reshape(-1), out[0].view(np.uint32)) and EQ(got_h[2], out[2].view(np.uint32))
    sc.close()
