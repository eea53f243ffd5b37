"""One config-4 trace (512^3 harmonic field, 8M random rays) with the options given as key=value (tools/sweep.py's short names), for ncu:
    ncu --set full -k regex:wave -c 1 -o out python tools/profile_c4.py wave=4 margin=2
prints the ray-step count (needed by tools/ncu_summary.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volumeraytracer_b200 as vrt
from volumeraytracer_b200 import workloads as W
import sweep

dev = torch.device("cuda", 0)
size = int(os.environ.get("C4_SIZE", 512))
nrays = int(os.environ.get("C4_RAYS", 8 << 20))
opts = dict(sweep.KV_DEFAULTS)
opts.update(dict((a.split("=")[0], int(a.split("=")[1])) for a in sys.argv[1:]))
ior = W.solve_harmonic_torch(size, dev, inner_radius=64.0 * size / 512.0, sweeps=300)
tr = W.clear_translucency_torch((size,) * 3, dev)
sc = vrt.TraceRaysCu.from_ior((size,) * 3, ior, tr)
pos, d = W.rays_random(nrays, 8.0, size - 9.0, 0x5EED0004)
tpos = torch.from_numpy(pos.view(np.int32).reshape(-1)).to(dev); tdir = torch.from_numpy(d.reshape(-1)).to(dev)
sc.normalise_rays_device(tpos, tdir)
for k, v in opts.items():
    sc.set_option(getattr(vrt, sweep.KV_OPTS[k]), v)
out = sc.trace_device(tpos, tdir, [1, 1, 1], 0, 4096)
torch.cuda.synchronize()
print("ray_steps", int(out[2].to(torch.int64).sum().item()))
