#!/usr/bin/env python
"""bench.py -- ray-steps/s of the marcher hot path on BASELINE.json's roofline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config 5 of BASELINE.json): 1024^3 index volume
n = 1.2 + 0.05 sin(2 pi x/256) cos(2 pi y/256) cos(2 pi z/256) -> 1022^3 x float4 gradient volume (17.08 GB) built on
the GPU by the scene-prep kernels; 4096^2 = 16 777 216 parallel +x rays per GPU, cap 2048 steps (every ray runs the
cap, so one pass = 34.36 G ray-steps per GPU).  A "step" of this benchmark is one such pass.

  value     whole-job G ray-steps/s with volume and ray buffers resident in HBM (vrt_trace_device), CUDA events on the
            launching stream, max over ranks.  ray-steps = sum of the end_iteration output (reference: cu:953-956).
  e2e       the same pass through the reference-facing host call (vrt_trace: pinned HOST ray buffers in, HOST results
            out, copies inside the timed region).
  roofline  algorithmic gather bytes (128 B per ray-step: 8 corners x 4 channels x fp32, cu:140-143) / launch time,
            against the measured HBM copy bandwidth of MEASURED_PEAKS.json; `roofline_l2` does the same against a
            random 32-byte-sector gather bandwidth measured on this GPU over an L2-resident buffer.
  cpu_baseline  the UNMODIFIED reference CPU marcher (oracle/_ref, trace_rays_cpu cu:376-394, all host threads) on a
            1/16 strided subsample of the same rays against the same volume bits (downloaded from the GPU).
Multi-GPU (weak scaling): the staged volume is built on rank 0 and replicated with one NCCL broadcast; rank 0 prepares
the global N x 16M-ray batch and scatters contiguous chunks; each rank marches its chunk; no per-step collective.

--impl reference: the reference's own CPU implementation of the path (oracle/_ref) on the host cores, on a bounded
sample of the same workload (a y/z window of the ray grid with the matching analytic sub-volume, built on the CPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_F32 = 128.0          # algorithmic gather bytes per ray-step, float scene
SIZE = 1024
RAY_SIDE = 4096
ITERATIONS = 2048


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class quiet_stdout:
    """The reference prints to std::cout ("Warning, maximum iterations hitted", cu:512-515); keep bench output to one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null); os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(",") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                out["samples"] = len(sm)
                out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 9)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[5 + i].strip().lower().startswith("active") for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU marcher on a bounded sample

def c5_sample_on_cpu(window, margin=24):
    """A y/z window of the config-5 ray grid and the analytic sub-volume it travels through, built on the CPU.
    Returns (volume interleaved, bounds, pos, dir) in the sub-volume's cropped coordinates."""
    from oracle import oracle as orc
    from volumeraytracer_b200 import workloads as W
    lo, hi = 2.0, SIZE - 3.0
    pitch = (hi - lo) / (RAY_SIDE - 1)
    j0 = RAY_SIDE // 2 - window // 2                       # window centred on the volume
    y0, y1 = lo + pitch * j0, lo + pitch * (j0 + window - 1)
    ymin = int(np.floor(y0)) - margin
    ymax = int(np.ceil(y1)) + margin
    xmax = min(SIZE, 2 + int(ITERATIONS * 0.2578 / 1.15) + margin)
    k = np.float32(2.0 * np.pi / 256.0)
    tx = k * np.arange(0, xmax, dtype=np.float32)
    ty = k * np.arange(ymin, ymax, dtype=np.float32)
    ior = (np.float32(1.2) + np.float32(0.05) * np.sin(tx)[:, None, None] * np.cos(ty)[None, :, None] * np.cos(ty)[None, None, :]).astype(np.float32)
    tr = np.full(ior.shape, 0xFFFFFFFF, np.uint32)
    ob, iorlog, planes, trc = orc.prep(ior.shape, ior, tr)
    vol = orc.fold(planes, trc)
    pos, d = W.rays_parallel_x(window, window, y0 - ymin, y1 - ymin, x0=2.0)
    pos, d = orc.normalise(ior.shape, ior, pos, d)
    return vol, ob, pos, d


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle import ref, oracle as orc
    t_setup = time.time()
    window = 512 if args.ref_window is None else args.ref_window            # 512^2 = 262 144 rays x 2048 steps per bench step
    vol, ob, pos, d = c5_sample_on_cpu(window)
    use_ref = ref.available()
    threads = len(os.sched_getaffinity(0))      # explicit: torchrun sets OMP_NUM_THREADS=1, the harness passes num_threads itself

    def one_pass():
        if use_ref:
            return ref.trace_live(vol, None, ob, [1, 1, 1], pos, d, ITERATIONS, 0, threads=threads)
        return orc.trace(vol, ob, pos, d, [1, 1, 1], ITERATIONS, round_mode=orc.ROUND_HOST, threads=threads)

    with quiet_stdout():
        for _ in range(args.warmup):
            out = one_pass()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = one_pass()
        dt = time.perf_counter() - t0
    steps = int(out[2].astype(np.int64).sum())
    value = steps * args.steps / dt / 1e9
    sample = "%d x %d ray window (1/%d of one GPU's 4096^2 rays) through the matching %dx%dx%d analytic sub-volume, cap %d" % (
        window, window, (RAY_SIDE // window) ** 2, ob[0], ob[1], ob[2], ITERATIONS)
    line = {
        "impl": "reference", "metric": "ray-steps/sec", "value": value, "unit": "G ray-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "G ray-steps/s", "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": "G ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "setup_s": round(time.time() - t_setup - dt, 1),
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "config 5: 1024^3 index volume (1022^3 float4 gradient volume, 17.08 GB), 4096^2 parallel +x rays per GPU x cap 2048 steps",
            "volume": "n=1.2+0.05 sin(2pi x/256) cos(2pi y/256) cos(2pi z/256)", "rays_per_gpu": RAY_SIDE * RAY_SIDE, "iterations": ITERATIONS,
            "scene": "float scene / float dirs, invscale 1, shipped translucency behaviour (compiled out, cu:785)",
            "parallelism": "rays%d (contiguous ray chunks, volume replicated by one NCCL broadcast, no per-step collective)" % n_gpus,
            "l2": "inputs larger than L2 (17 GB volume, 0.4 GB ray buffers); no explicit flush"}


# ---------------------------------------------------------------------------------------------------
# our arm

def run_ours(args, rank, local_rank, world):
    import torch
    import volumeraytracer_b200 as vrt
    from volumeraytracer_b200 import workloads as W
    import torch.distributed as dist
    from volumeraytracer_b200 import dist as vd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    size, side, iters = args.size, args.ray_side, args.iterations
    n_local = side * side
    nvox = (size - 2) ** 3

    # ---- scene: built once on rank 0 (GPU scene prep), replicated by ONE NCCL broadcast over NVLink --------------
    t0 = time.time()
    scene0 = None
    if rank == 0:
        ior = W.ior_c5_torch(size, dev)
        tr = W.clear_translucency_torch((size,) * 3, dev)
        scene0 = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr)
        del tr
    if world > 1:
        vol_t = torch.empty(nvox * 4, dtype=torch.float32, device=dev)
        if rank == 0:
            scene0.export_device(vol_t)
        torch.cuda.synchronize()
        tb = time.time()
        vd.broadcast_volume(vol_t, src=0)                     # ONE collective, at scene creation (NCCL over NVLink)
        torch.cuda.synchronize()
        bcast_s = time.time() - tb
        scene = vrt.TraceRaysCu.from_device([size - 2] * 3, vol_t, None, borrow=True)
    else:
        scene, bcast_s = scene0, 0.0

    # ---- rays: rank 0 prepares the global batch (world x side^2 rays over [2, size-3]^2) and scatters contiguous chunks
    pos_t = torch.empty(n_local * 3, dtype=torch.int32, device=dev)
    dir_t = torch.empty(n_local * 3, dtype=torch.float32, device=dev)
    if rank == 0:
        chunks_p, chunks_d = [], []
        lo, hi = 2.0, size - 3.0
        rows = world * side
        for r in range(world):
            y0 = lo + (hi - lo) * (r * side) / (rows - 1)
            y1 = lo + (hi - lo) * (r * side + side - 1) / (rows - 1)
            p, d = W.rays_parallel_x(side, side, y0, y1, x0=2.0, lo_z=lo, hi_z=hi)
            tp = torch.from_numpy(p.view(np.int32).reshape(-1)).to(dev)
            td = torch.from_numpy(d.reshape(-1)).to(dev)
            scene0.normalise_rays_device(tp, td)                  # f2 on the GPU (needs ior, which only rank 0 keeps)
            chunks_p.append(tp); chunks_d.append(td)
        if world > 1:
            dist.scatter(pos_t, chunks_p, src=0); dist.scatter(dir_t, chunks_d, src=0)
        else:
            pos_t, dir_t = chunks_p[0], chunks_d[0]
        del chunks_p, chunks_d
    else:
        dist.scatter(pos_t, None, src=0); dist.scatter(dir_t, None, src=0)
    if world > 1 and rank == 0:
        ior = None
    torch.cuda.synchronize()
    setup_s = time.time() - t0

    epos = torch.empty_like(pos_t); edir = torch.empty_like(dir_t)
    eit = torch.empty(n_local, dtype=torch.int32, device=dev); light = torch.empty(n_local, dtype=torch.int32, device=dev)
    isc = [1.0, 1.0, 1.0]
    stream = torch.cuda.current_stream(dev)

    def one_pass():
        scene.trace_device(pos_t, dir_t, isc, 0, iters, epos=epos, edir=edir, eit=eit, light=light, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = vrt.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for k in range(args.steps):
        one_pass()
        evs[k + 1].record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = vrt.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    steps_local = int(eit.to(torch.int64).sum().item())

    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    s_all = torch.tensor([steps_local], dtype=torch.int64, device=dev)
    l_all = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX); dist.all_reduce(s_all, op=dist.ReduceOp.SUM); dist.all_reduce(l_all, op=dist.ReduceOp.SUM)
    max_ms = float(t_ms.item()); steps_global = int(s_all.item())
    value = steps_global * args.steps / (max_ms * 1e-3) / 1e9

    # ---- e2e: the reference-facing host call, pinned host buffers, copies inside the timed region -------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_pos = torch.empty(n_local * 3, dtype=torch.int32).pin_memory(); h_dir = torch.empty(n_local * 3, dtype=torch.float32).pin_memory()
    h_pos.copy_(pos_t.cpu()); h_dir.copy_(dir_t.cpu())
    h_epos = torch.empty_like(h_pos).pin_memory(); h_edir = torch.empty_like(h_dir).pin_memory()
    h_eit = torch.empty(n_local, dtype=torch.int32).pin_memory(); h_light = torch.empty(n_local, dtype=torch.int32).pin_memory()
    scene.trace_host_buffers(h_pos, h_dir, isc, 0, iters, h_epos, h_edir, h_eit, h_light)      # warm-up
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.trace_host_buffers(h_pos, h_dir, isc, 0, iters, h_epos, h_edir, h_eit, h_light)
    e2e_s = time.perf_counter() - t1
    barrier()
    e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_value = steps_global * e2e_steps / float(e_t.item()) / 1e9
    e2e_ok = bool(torch.equal(h_eit, eit.cpu()) and torch.equal(h_epos, epos.cpu()))

    if rank == 0:
        peaks, peak_src = measured_peaks()
        launch_s = float(np.mean(per_launch_ms)) * 1e-3
        achieved = B_ALG_F32 * steps_local / launch_s / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_c5_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": peak_src, "bytes_per_ray_step": B_ALG_F32, "ray_steps_per_launch": steps_local,
                "launch_ms": launch_s * 1e3,
                "note": "coherent bundle: the 8-corner gathers are served from registers/L1 (cell cache), so the algorithmic gather "
                        "rate exceeds DRAM bandwidth; see roofline_l2 and profiles/ for the counters"}
        # what actually bounds the kernel on this coherent workload is instruction issue: report that
        # utilisation too, from the committed ncu instruction count per warp-step and the clock sampled in this run
        roof_issue = None
        try:
            wi = json.load(open(prof)).get("warp_instructions_per_warp_step")
            sm_hz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
            peak_issue = 148 * 4 * sm_hz * 1e6                      # one warp-instruction per scheduler per clock, 4 schedulers per SM
            ach_issue = wi * (steps_local / 32.0) / launch_s
            roof_issue = {"bound": "issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s", "frac": ach_issue / peak_issue,
                          "warp_instructions_per_warp_step": wi, "source": "profiles/ncu_c5_summary.json (ncu smsp__inst_executed.sum)"}
        except Exception:
            roof_issue = None
        g = C.c_double(0.0)
        roof_l2 = None
        if vrt.lib().vrt_measure_gather_bandwidth(local_rank, 32 << 20, 32, 3, C.byref(g)) == 0 and g.value > 0:
            roof_l2 = {"bound": "l2-gather", "achieved": achieved, "peak": g.value, "unit": "GB/s", "frac": achieved / g.value,
                       "how": "random 32-byte-sector gather over a 32 MiB (L2-resident) buffer, measured in this run"}
        line = {
            "metric": "ray-steps/sec", "value": value, "unit": "G ray-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "G ray-steps/s", "h2d_bytes_per_step": n_local * 24 * world, "d2h_bytes_per_step": n_local * 32 * world,
                    "steps": e2e_steps, "matches_device_run": e2e_ok, "call": "vrt_trace (C ABI, pinned host buffers)"},
            "gpu_launches": int(l_all.item()), "roofline": roof, "roofline_l2": roof_l2, "roofline_issue": roof_issue,
            "ray_steps_per_pass": steps_global, "setup_s": round(setup_s, 2), "nccl_broadcast_s": round(bcast_s, 3),
            "kernel": {"variant": scene.get_option(vrt.VRT_OPT_KERNEL), "block": scene.get_option(vrt.VRT_OPT_BLOCK_THREADS),
                       "refill": scene.get_option(vrt.VRT_OPT_REFILL), "steps_per_poll": scene.get_option(vrt.VRT_OPT_STEPS_PER_POLL)},
        }
        if world == 1 and not args.no_cpu_baseline:
            line.update(cpu_baseline_and_parity(scene, pos_t, dir_t, epos, edir, eit, side, iters))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_and_parity(scene, pos_t, dir_t, epos, edir, eit, side, iters):
    """Rank 0, N=1: (a) the reference CPU marcher timed on a strided 1/16 subsample against the same volume bits,
    (b) bit-exact parity of the GPU result against the CPU oracle (device rounding) on a 1/256 subsample."""
    from oracle import ref, oracle as orc
    out = {}
    vol, _ = scene.download_volume()
    ob = scene._output_sizes
    P = pos_t.cpu().numpy().view(np.uint32).reshape(side, side, 3)
    D = dir_t.cpu().numpy().reshape(side, side, 3)
    # (b) parity, 1/256 subsample (every 16th ray in y and z)
    p_s = np.ascontiguousarray(P[::16, ::16].reshape(-1, 3)); d_s = np.ascontiguousarray(D[::16, ::16].reshape(-1, 3))
    want = orc.trace(vol, ob, p_s, d_s, [1, 1, 1], iters, round_mode=orc.ROUND_DEVICE)
    g_pos = epos.cpu().numpy().view(np.uint32).reshape(side, side, 3)[::16, ::16].reshape(-1, 3)
    g_dir = edir.cpu().numpy().reshape(side, side, 3)[::16, ::16].reshape(-1, 3)
    g_it = eit.cpu().numpy().view(np.uint32).reshape(side, side)[::16, ::16].reshape(-1)
    out["parity"] = {"rays_checked": int(p_s.shape[0]), "oracle": "oracle/vrt_oracle.c ROUND_DEVICE",
                     "bit_exact": bool(np.array_equal(g_pos, want[0]) and np.array_equal(g_dir, want[1]) and np.array_equal(g_it, want[2]))}
    # (a) cpu baseline, 1/16 subsample (every 4th ray in y and z): same volume, same coverage of it as the full batch
    p_c = np.ascontiguousarray(P[::4, ::4].reshape(-1, 3)); d_c = np.ascontiguousarray(D[::4, ::4].reshape(-1, 3))
    use_ref = ref.available()
    threads = len(os.sched_getaffinity(0))
    with quiet_stdout():
        t0 = time.perf_counter()
        if use_ref:
            res = ref.trace_live(vol, None, ob, [1, 1, 1], p_c, d_c, iters, 0, threads=threads)
        else:
            res = orc.trace(vol, ob, p_c, d_c, [1, 1, 1], iters, round_mode=orc.ROUND_HOST, threads=threads)
        dt = time.perf_counter() - t0
    steps = int(res[2].astype(np.int64).sum())
    out["cpu_baseline"] = {"value": steps / dt / 1e9, "unit": "G ray-steps/s", "cores": threads, "kind": "reference" if use_ref else "port",
                           "sample": "every 4th ray in y and z of the 4096^2 batch (1 048 576 rays x cap 2048) against the full 17 GB volume "
                                     "downloaded from the GPU; one pass, %.1f s" % dt,
                           "impl": "unmodified reference trace_rays_cpu (cu:376-394) from oracle/_ref" if use_ref else "oracle/vrt_oracle.c"}
    # tolerance vs the reference's CPU semantics on the parity subsample (host rounding): north_star 1e-3 voxel / 1e-5 rad
    host = orc.trace(vol, ob, p_s, d_s, [1, 1, 1], iters, round_mode=orc.ROUND_HOST)
    dp = float(np.abs(g_pos.astype(np.int64) - host[0].astype(np.int64)).max() / 65536.0)
    a, b = g_dir.astype(np.float64), host[1].astype(np.float64)
    ang = float(np.arccos(np.clip(np.sum(a * b, 1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1)), -1, 1)).max())
    out["parity"].update({"vs_reference_cpu_rounding": {"max_pos_err_voxel": dp, "max_dir_err_rad": ang,
                                                         "step_count_mismatches": int(np.sum(g_it != host[2]))}})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--ray-side", type=int, default=RAY_SIDE)
    ap.add_argument("--iterations", type=int, default=ITERATIONS)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-window", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun but asked for N GPUs: re-launch ourselves under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
