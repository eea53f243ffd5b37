"""GPU exploration script (not part of the product): times kernel variants on the BASELINE configurations.
usage: python tools/sweep.py [c5|c4|c2|ref|l2 ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volumeraytracer_b200 as vrt
from volumeraytracer_b200 import workloads as W

dev = torch.device("cuda", 0)
BRICK = bool(int(os.environ.get("SWEEP_BRICK", "0")))
KEEP16 = bool(int(os.environ.get("SWEEP_KEEP_I16", "0")))
TEX = bool(int(os.environ.get("SWEEP_TEXTURE", "0")))
PAIR = bool(int(os.environ.get("SWEEP_PAIR", "0")))


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best, out


KV_OPTS = {"wave": "VRT_OPT_WAVE_LOG2", "margin": "VRT_OPT_WAVE_MARGIN", "check": "VRT_OPT_WAVE_CHECK", "tail": "VRT_OPT_WAVE_TAIL_PERMILLE",
           "wctas": "VRT_OPT_WAVE_CTAS_PER_SM", "wrefill": "VRT_OPT_WAVE_REFILL", "allclear": "VRT_OPT_ALL_CLEAR_KERNEL", "region": "VRT_OPT_REGION_LOG2", "rounds": "VRT_OPT_REGION_ROUNDS", "kver": "VRT_OPT_KERNEL",
           "block": "VRT_OPT_BLOCK_THREADS", "refill": "VRT_OPT_REFILL", "poll": "VRT_OPT_STEPS_PER_POLL", "ctas": "VRT_OPT_MAX_CTAS_PER_SM"}
KV_DEFAULTS = {"wave": -1, "margin": 8, "check": 16, "tail": 20, "wctas": 0, "wrefill": 8, "allclear": 1, "region": 0, "rounds": 12, "kver": 0, "block": 128, "refill": 32, "poll": 128, "ctas": 0}


def run_kv_variants(name, co, tpos, tdir, iterations, variants, live=False):
    """variants: dicts of option short names (KV_OPTS); unspecified options take KV_DEFAULTS"""
    first = None
    for var in variants:
        opts = dict(KV_DEFAULTS); opts.update(var)
        if opts["region"] > 0 and "wave" not in var:
            opts["wave"] = 0                      # VRT_OPT_REGION_LOG2 is an alias that only counts while VRT_OPT_WAVE_LOG2 is 0
        for k, v in opts.items():
            co.set_option(getattr(vrt, KV_OPTS[k]), v)
        t, out = timed(lambda: co.trace_device(tpos, tdir, [1, 1, 1], 0x40000000 if live else 0, iterations, live_translucency=live))
        steps = int(out[2].to(torch.int64).sum().item())
        same = None
        if first is None:
            first = [o.clone() for o in out]
        else:
            same = bool(all(torch.equal(a, b) for a, b in zip(out, first)))
        rounds = co.get_option(vrt.VRT_INFO_WAVE_ROUNDS) if opts["wave"] >= 0 else None
        print(json.dumps(dict(cfg=name, **var, sec=round(t, 4), steps=steps, grays=round(steps / t / 1e9, 2), same_bits_as_first=same, wave_rounds=rounds)), flush=True)


def run_variants(name, co, tpos, tdir, iterations, variants, live=False):
    if variants and isinstance(variants[0], dict):
        return run_kv_variants(name, co, tpos, tdir, iterations, variants, live)
    res = []
    for var in variants:
        kver, block, refill, poll = var[:4]
        ctas = var[4] if len(var) > 4 else 0
        region = var[5] if len(var) > 5 else 0
        co.set_option(vrt.VRT_OPT_WAVE_LOG2, 0 if region > 0 else -1)
        co.set_option(vrt.VRT_OPT_REGION_LOG2, region)
        co.set_option(vrt.VRT_OPT_REGION_ROUNDS, var[6] if len(var) > 6 else 24)
        co.set_option(vrt.VRT_OPT_KERNEL, kver); co.set_option(vrt.VRT_OPT_BLOCK_THREADS, block)
        co.set_option(vrt.VRT_OPT_REFILL, refill); co.set_option(vrt.VRT_OPT_STEPS_PER_POLL, poll)
        co.set_option(vrt.VRT_OPT_MAX_CTAS_PER_SM, ctas)
        t, out = timed(lambda: co.trace_device(tpos, tdir, [1, 1, 1], 0x40000000 if live else 0, iterations, live_translucency=live))
        steps = int(out[2].to(torch.int64).sum().item())
        print(json.dumps(dict(cfg=name + ("_brick" if BRICK else "") + ("_tex" if TEX else "") + ("_pair" if PAIR else ""), kver=kver, block=block, refill=refill, poll=poll, ctas=ctas, region=region, sec=round(t, 4), steps=steps,
                              grays=round(steps / t / 1e9, 2))), flush=True)
        res.append((steps / t / 1e9, kver, block, refill, poll))
    return res


def _env_variants():
    v = os.environ.get("SWEEP_VARIANTS")
    if not v:
        return None
    items = v.replace(";", ":").split(":")
    if "=" in v:            # key=value form: "wave=4,margin=2:wave=-1"
        return [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in item.split(",") if kv) for item in items]
    return [tuple(int(x) for x in item.split(",")) for item in items]


VARIANTS = _env_variants() or [(1, 128, 0, 8), (2, 128, 0, 8), (3, 128, 0, 8), (3, 256, 0, 8), (3, 64, 0, 8), (3, 128, 32, 8), (3, 128, 16, 8),
            (3, 128, 8, 8), (3, 128, 1, 8), (3, 128, 16, 2), (3, 128, 16, 32), (3, 256, 16, 8), (2, 128, 16, 8), (1, 128, 16, 8)]


def cfg_c5(size=1024, nray=4096, iterations=2048):
    ior = W.ior_c5_torch(size, dev); tr = W.clear_translucency_torch((size,) * 3, dev)
    t0 = time.time(); sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr, bricked=BRICK, texture=TEX, paired=PAIR); torch.cuda.synchronize()
    print("c5 scene prep %.2fs, volume %.2f GB" % (time.time() - t0, sc.volume_bytes / 1e9), flush=True)
    del tr
    pos, d = W.rays_parallel_x(nray, nray, 2.0, size - 3.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c5_%d" % size, sc, tpos, tdir, iterations, VARIANTS)
    sc.close()


def cfg_c5i(size=1024, nray=4096, iterations=2048):
    """config 5 with the int16 scene / int16 directions instantiation (64 B per ray-step, 8.5 GB volume)"""
    ior = W.ior_c5_torch(size, dev)
    ior_u = torch.floor(ior.double() * 65536.0 + 0.5).to(torch.int64)
    ior_u = torch.where(ior_u >= (1 << 31), ior_u - (1 << 32), ior_u).to(torch.int32)
    del ior
    tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior_u, tr, bricked=BRICK, keep_i16=KEEP16, paired=PAIR); torch.cuda.synchronize()
    print("c5 int16 scene, volume %.2f GB" % (sc.volume_bytes / 1e9), flush=True)
    del tr
    pos, d = W.rays_parallel_x(nray, nray, 2.0, size - 3.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(W.dirs_to_i16(d).reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c5i_%d" % size, sc, tpos, tdir, iterations, VARIANTS)
    sc.close()


def cfg_c4(size=512, nrays=8 << 20, iterations=4096):
    nrays = int(os.environ.get("SWEEP_C4_RAYS", nrays))          # a shard of the batch, as one GPU of a multi-GPU run sees it
    ior = W.solve_harmonic_torch(size, dev, inner_radius=64.0 * size / 512.0, sweeps=300)
    tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr, bricked=BRICK, texture=TEX, paired=PAIR); torch.cuda.synchronize()
    pos, d = W.rays_random(nrays, 8.0, size - 9.0, 0x5EED0004)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c4_%d_%d" % (size, nrays), sc, tpos, tdir, iterations, VARIANTS)
    sc.close()


def cfg_c2h(size=256, nray=1024, iterations=4096):
    """config 2 through the HOST call with pageable buffers, for several chunk sizes (VRT_OPT_CHUNK_RAYS; 0 = the library's choice)"""
    ior = W.ior_luneburg_torch(size, dev, 100.0 * size / 256.0); tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr); torch.cuda.synchronize()
    pos, d = W.rays_parallel_x(nray, nray, 30.0 * size / 256.0, 225.0 * size / 256.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    p_h = tpos.cpu().numpy().view(np.uint32); d_h = tdir.cpu().numpy()
    n = p_h.size // 3
    epos = np.zeros_like(p_h); edir = np.zeros_like(d_h); eit = np.zeros(n, np.uint32); light = np.zeros(n, np.uint32)
    for chunk in (0, 0, 65536, 131072, 262144, 524288, 1048576, 0):
        sc.set_option(vrt.VRT_OPT_CHUNK_RAYS, chunk)
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter(); sc.trace_host_buffers(p_h, d_h, [1, 1, 1], 0, iterations, epos, edir, eit, light); best = min(best, time.perf_counter() - t0)
        steps = int(eit.astype(np.int64).sum())
        print(json.dumps(dict(cfg="c2_hostcall_pageable", chunk=chunk, sec=round(best, 5), grays=round(steps / best / 1e9, 2))), flush=True)
    sc.close()


def cfg_c5h(size=1024, nray=4096, iterations=2048):
    """config 5 through the HOST call with PAGEABLE buffers (numpy arrays, like the std::vectors of the reference API)"""
    ior = W.ior_c5_torch(size, dev); tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr); torch.cuda.synchronize()
    del tr, ior
    pos, d = W.rays_parallel_x(nray, nray, 2.0, size - 3.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    p_h = tpos.cpu().numpy().view(np.uint32); d_h = tdir.cpu().numpy()
    n = p_h.size // 3
    epos = np.empty_like(p_h); edir = np.empty_like(d_h); eit = np.empty(n, np.uint32); light = np.empty(n, np.uint32)
    for rep in range(4):
        t0 = time.perf_counter(); sc.trace_host_buffers(p_h, d_h, [1, 1, 1], 0, iterations, epos, edir, eit, light); dt = time.perf_counter() - t0
        steps = int(eit.astype(np.int64).sum())
        print(json.dumps(dict(cfg="c5_hostcall_pageable", sec=round(dt, 4), grays=round(steps / dt / 1e9, 2))), flush=True)
    sc.close()


def cfg_c4h(size=512, nrays=8 << 20, iterations=4096):
    """config 4 through the HOST call (vrt_trace): default options (coherence probe -> region mode) vs region mode disabled"""
    ior = W.solve_harmonic_torch(size, dev, inner_radius=64.0 * size / 512.0, sweeps=300)
    tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr); torch.cuda.synchronize()
    pos, d = W.rays_random(nrays, 8.0, size - 9.0, 0x5EED0004)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    p_h = tpos.cpu().numpy().view(np.uint32).reshape(-1, 3); d_h = tdir.cpu().numpy().reshape(-1, 3)
    n = p_h.shape[0]
    outs = [np.zeros_like(p_h), np.zeros_like(d_h), np.zeros(n, np.uint32), np.zeros(n, np.uint32)]       # touched pages, like the caller's pre-sized vectors
    ref_out = None
    combos = [("auto", 0, 0), ("off", -1, 0), ("auto", 0, 0), ("off", -1, 0)]
    for c in os.environ.get("SWEEP_C4H", "").split(":"):
        if c: combos.append(("forced",) + tuple(int(v) for v in c.split(",")))
    for name, region, chunk in combos:
        sc.set_option(vrt.VRT_OPT_REGION_LOG2, region); sc.set_option(vrt.VRT_OPT_CHUNK_RAYS, chunk)
        t0 = time.perf_counter(); sc.trace_host_buffers(p_h, d_h, [1, 1, 1], 0, iterations, *outs); dt = time.perf_counter() - t0
        steps = int(outs[2].astype(np.int64).sum())
        same = True if ref_out is None else bool(all(np.array_equal(a, b) for a, b in zip(outs, ref_out)))
        ref_out = ref_out or [o.copy() for o in outs]
        print(json.dumps(dict(cfg="c4_hostcall", region=name, log2=region, chunk=chunk, sec=round(dt, 4), grays=round(steps / dt / 1e9, 2), same_bits=same)), flush=True)
    sc.close()


def cfg_c3(size=512, nray=2048, iterations=4096):
    ior = W.ior_sines_torch(size, dev); tr = W.translucency_c3_torch(size, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr, bricked=BRICK, texture=TEX, paired=PAIR); torch.cuda.synchronize()
    pos, d = W.rays_parallel_x(nray, nray, 4.0, size - 5.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c3_%d_live" % size, sc, tpos, tdir, iterations, VARIANTS, live=True)
    sc.close()


def cfg_c1(size=64, nray=64, iterations=1024, reps=64):
    """config 1 geometry (constant index: the whole volume is empty space), replicated to 262 144 rays"""
    ior = torch.full((size,) * 3, 1.0, dtype=torch.float32, device=dev); tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr); torch.cuda.synchronize()
    pos, d = W.rays_parallel_x(nray * 8, nray * 8, 1.5, 61.5, x0=1.5)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c1_%d" % size, sc, tpos, tdir, iterations, VARIANTS)
    sc.close()


def cfg_c2(size=256, nray=1024, iterations=4096, with_ref=False):
    ior = W.ior_luneburg_torch(size, dev); tr = W.clear_translucency_torch((size,) * 3, dev)
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr, bricked=BRICK, texture=TEX, paired=PAIR); torch.cuda.synchronize()
    pos, d = W.rays_parallel_x(nray, nray, 30.0, 225.0, x0=2.0)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    run_variants("c2_%d" % size, sc, tpos, tdir, iterations, VARIANTS)
    if with_ref:
        # the reference's own CUDA kernel (sm_100 build of the unmodified source) on the same rays, host buffers
        from oracle import ref
        if ref.available(cuda=True):
            vol, trc = sc.download_volume()
            planes = [np.ascontiguousarray(vol[:, a]) for a in range(3)]
            rt = ref.RefTracer(sc._output_sizes, planes, trc, cuda=True)
            p_h = tpos.cpu().numpy().view(np.uint32).reshape(-1, 3); d_h = tdir.cpu().numpy().reshape(-1, 3)
            for rep in range(2):
                t0 = time.time(); out = rt.trace(p_h, d_h, [1, 1, 1], 0, iterations); dt = time.time() - t0
                steps = int(out[2].astype(np.int64).sum())
                print(json.dumps(dict(cfg="c2_refcuda_hostcall", sec=round(dt, 3), steps=steps, grays=round(steps / dt / 1e9, 3))), flush=True)
            # ours through the host-buffer call for the same comparison
            sc.set_option(vrt.VRT_OPT_KERNEL, 3); sc.set_option(vrt.VRT_OPT_REFILL, 16)
            for rep in range(2):
                t0 = time.time(); mine = sc.trace_rays_cu(p_h, d_h, [1, 1, 1], 0, iterations); dt = time.time() - t0
                print(json.dumps(dict(cfg="c2_ours_hostcall", sec=round(dt, 3), grays=round(steps / dt / 1e9, 3),
                                      bit_exact_vs_refcuda=bool(all(np.array_equal(a, b) for a, b in zip(mine[:4], out[:4]))))), flush=True)
    sc.close()


def cfg_latency():
    """small batches through the host call: the reference's own perf instrument is 2000 rays (performance_test.h:9-86)"""
    from oracle import ref
    size = 64
    ior = W.ior_sines(size, period=32.0); tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    sc = vrt.RaytraceScene(ior.shape, ior, tr)
    co = sc._calculation_object
    vol, trc = co.download_volume()
    planes = [np.ascontiguousarray(vol[:, a]) for a in range(3)]
    rt = ref.RefTracer(co._output_sizes, planes, trc, cuda=True) if ref.available(cuda=True) else None
    for n in (256, 2048, 16384, 131072):
        pos, d = W.rays_random(n, 4.0, size - 5.0, 7)
        pos = pos - np.uint32(0x10000)
        for name, fn in (("ours", lambda: co.trace_rays_cu(pos, d, [1, 1, 1], 0, 1024)), ("refcuda", (lambda: rt.trace(pos, d, [1, 1, 1], 0, 1024)) if rt else None)):
            if fn is None:
                continue
            fn(); ts = []
            for _ in range(5):
                t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
            print(json.dumps(dict(cfg="latency", impl=name, rays=n, ms=round(min(ts) * 1e3, 3), steps=int(out[2].astype(np.int64).sum()))), flush=True)


def cfg_l2():
    import ctypes as C
    for mb in (16, 32, 48, 64, 256, 4096):
        for sec in (32, 16):
            g = C.c_double()
            rc = vrt.lib().vrt_measure_gather_bandwidth(0, mb << 20, sec, 3, C.byref(g))
            print(json.dumps(dict(cfg="gather", mb=mb, sector=sec, gbs=round(g.value, 1), rc=rc)), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["l2", "c2", "c5", "c4", "c3"]
    print(torch.cuda.get_device_name(0), "cpus", os.cpu_count(), flush=True)
    for w in which:
        {"c2h": cfg_c2h, "c5h": cfg_c5h, "c4h": cfg_c4h, "c1": cfg_c1, "c5": cfg_c5, "c5i": cfg_c5i, "c4": cfg_c4, "c3": cfg_c3, "c2": cfg_c2, "l2": cfg_l2, "latency": cfg_latency}[w]()
