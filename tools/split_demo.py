"""Multi-device split of the TraceRaysCu<> drop-in on a SKEWED batch (config-3-like: absorption grows with y and the rays are
ordered y-major, so a static split hands one device the long rays): static split vs VRT_SPLIT_PIECES=k, through the reference's
TraceRaysCu<float>::trace_rays_cu on the drop-in (oracle/_ref/libvrt_dropin.so).  Needs >= 2 visible GPUs to show a difference.
    python tools/split_demo.py [size] [rays_per_axis]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volumeraytracer_b200 as vrt
from volumeraytracer_b200 import workloads as W
from oracle import ref

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nray = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda", 0)
# scene prep and ray normalisation on the GPU (bit-compatible with the reference's host code), then the reference's class
prep = vrt.TraceRaysCu.from_ior((size,) * 3, W.ior_sines_torch(size, dev), W.translucency_c3_torch(size, dev))
vol, trc = prep.download_volume()
pos, d = W.rays_parallel_x(nray, nray, 4.0 * size / 512.0, size - 1 - 4.0 * size / 512.0, x0=2.0)
tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
prep.normalise_rays_device(tpos, tdir)
p = tpos.cpu().numpy().view(np.uint32).reshape(-1, 3); dn = tdir.cpu().numpy().reshape(-1, 3)
prep.close(); del tpos, tdir
torch.cuda.empty_cache()
os.environ["VRT_LIVE_TRANSLUCENCY"] = "1"
tracer = ref.RefTracer([size - 2] * 3, [np.ascontiguousarray(vol[:, k]) for k in range(3)], trc, cuda="dropin")
base = None
for pieces in (1, 1, 2, 4, 8, 1):
    os.environ["VRT_SPLIT_PIECES"] = str(pieces)
    t0 = time.perf_counter(); out = tracer.trace(p, dn, [1, 1, 1], 0x40000000, 4096); dt = time.perf_counter() - t0
    steps = int(out[2].astype(np.int64).sum())
    same = True if base is None else bool(all(np.array_equal(a, b) for a, b in zip(out[:4], base[:4])))
    base = base or out
    print(json.dumps(dict(cfg="split_demo", gpus=torch.cuda.device_count(), size=size, rays=int(p.shape[0]), pieces_per_device=pieces, sec=round(dt, 4),
                          ray_steps=steps, grays=round(steps / dt / 1e9, 2), same_bits=same)), flush=True)
