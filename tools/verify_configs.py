"""Runs BASELINE.json configs 2, 3 and 4 at FULL size on the GPU (config 5 is bench.py; config 1 is a GPU test), checks a
subsample of each against the CPU oracle bit for bit, compares with the reference's CPU rounding, times the reference CPU
marcher on the same subsample, and prints one JSON line per config.   python tools/verify_configs.py > profiles/rNN_configs.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volumeraytracer_b200 as vrt
from volumeraytracer_b200 import workloads as W
from oracle import oracle as orc, ref

dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best, out


def run(name, ior, tr, pos, d, iterations, live, minb, stride, note):
    size = ior.shape[0]
    sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr)
    tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
    sc.normalise_rays_device(tpos, tdir)
    sc.trace_device(tpos, tdir, [1, 1, 1], minb, iterations, live_translucency=live)          # warm-up
    t, out = timed(lambda: sc.trace_device(tpos, tdir, [1, 1, 1], minb, iterations, live_translucency=live))
    epos, edir, eit, light = [o.cpu().numpy() for o in out]
    steps = int(eit.view(np.uint32).astype(np.int64).sum())
    vol, trc = sc.download_volume()
    ob = sc._output_sizes
    sel = np.arange(0, pos.shape[0], stride)
    p_s = tpos.cpu().numpy().view(np.uint32).reshape(-1, 3)[sel]; d_s = tdir.cpu().numpy().reshape(-1, 3)[sel]
    want = orc.trace(vol, ob, p_s, d_s, [1, 1, 1], iterations, translucency=trc if live else None, min_brightness=minb, round_mode=orc.ROUND_DEVICE)
    g = (epos.view(np.uint32).reshape(-1, 3)[sel], edir.reshape(-1, 3)[sel], eit.view(np.uint32)[sel], light.view(np.uint32)[sel])
    exact = all(np.array_equal(a, b) for a, b in zip(g, want[:4]))
    host = orc.trace(vol, ob, p_s, d_s, [1, 1, 1], iterations, translucency=trc if live else None, min_brightness=minb, round_mode=orc.ROUND_HOST)
    same = g[2] == host[2]
    dp = float(np.abs(g[0].astype(np.int32) - host[0].astype(np.int32))[same].max() / 65536.0)
    a, b = g[1].astype(np.float64)[same], host[1].astype(np.float64)[same]
    ang = float(np.arccos(np.clip(np.sum(a * b, 1) / np.maximum(np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1), 1e-300), -1, 1)).max())
    cpu = None
    if ref.available():
        t0 = time.perf_counter()
        res = ref.trace_live(vol, trc if live else None, ob, [1, 1, 1], p_s, d_s, iterations, minb, threads=len(os.sched_getaffinity(0)))
        dt = time.perf_counter() - t0
        cpu = {"g_ray_steps_per_s": int(res[2].astype(np.int64).sum()) / dt / 1e9, "threads": len(os.sched_getaffinity(0)), "rays": int(sel.size),
               "bit_exact_vs_oracle_host_mode": bool(all(np.array_equal(x, y) for x, y in zip(res[:4], host[:4])))}
    classes = {"left_volume_or_cap": int(np.sum(light.view(np.uint32) >= minb)) if live else None,
               "below_min_brightness": int(np.sum(light.view(np.uint32) < minb)) if live else None}
    print(json.dumps({"config": name, "note": note, "rays": int(pos.shape[0]), "iterations_cap": iterations, "ray_steps": steps, "seconds": t,
                      "g_ray_steps_per_s": steps / t / 1e9, "mean_steps_per_ray": steps / pos.shape[0],
                      "parity": {"subsample_rays": int(sel.size), "bit_exact_vs_oracle_device_mode": bool(exact),
                                 "vs_reference_cpu_rounding": {"max_pos_err_voxel": dp, "max_dir_err_rad": ang, "step_count_mismatches": int(np.sum(~same))}},
                      "reference_cpu": cpu, "exit_classes": classes}), flush=True)
    sc.close()


def main():
    size = 256
    pos, d = W.rays_parallel_x(1024, 1024, 30.0, 225.0, x0=2.0)
    run("2: 256^3 Luneburg lens, 1M parallel rays", W.ior_luneburg_torch(size, dev), W.clear_translucency_torch((size,) * 3, dev), pos, d, 4096, False, 0, 64,
        "float scene, shipped translucency behaviour")
    size = 512
    pos, d = W.rays_parallel_x(2048, 2048, 4.0, size - 5.0, x0=2.0)
    run("3: 512^3 index + translucency, 4M rays, min-brightness termination", W.ior_sines_torch(size, dev), W.translucency_c3_torch(size, dev), pos, d, 4096, True,
        0x40000000, 256, "float scene, LIVE translucency plane + min_brightness 0x40000000 + opaque ball")
    pos, d = W.rays_random(8 << 20, 8.0, size - 9.0, 0x5EED0004)
    run("4: 512^3 harmonic field, 8M randomly directed rays", W.solve_harmonic_torch(size, dev, inner_radius=64.0, sweeps=300),
        W.clear_translucency_torch((size,) * 3, dev), pos, d, 4096, False, 0, 512, "float scene; incoherent batch (default launch parameters, linear layout)")


if __name__ == "__main__":
    main()
