"""Seeded small scenes shared by the parity tests (oracle vs reference, CUDA vs oracle, golden fixtures)."""
import numpy as np

from volumeraytracer_b200 import workloads as W


def scaling_test_inputs(kind):
    """Inputs of the reference's scaling_test (cuda_volume_raytracer_test.h:4-46): 1000x10x10 volume, n ramps
    1 -> 2 along axis 0, two rays +x / -x, invscale 2, iterations 1e6."""
    b = [1000, 10, 10]
    npx, nl = 1000 * 10 * 10, 100
    if kind == "f32":
        ior = np.zeros(npx, np.float32)
        ior[:10 * nl] = 1.0
        ior[-10 * nl:] = 2.0
        for i in range(10, 990):
            ior[i * nl:(i + 1) * nl] = np.float32(1) + np.float32(1) * np.float32(i) / np.float32(979)
        d = np.array([16, 0, 0, -16, 0, 0], np.float32)
    else:
        ior = np.zeros(npx, np.uint32)
        ior[:10 * nl] = 0x10000
        ior[-10 * nl:] = 0x20000
        for i in range(10, 990):
            ior[i * nl:(i + 1) * nl] = 0x10000 + (0x10000 * i) // 979
        d = np.array([0x1000, 0, 0, -0x1000, 0, 0], np.int16)
    tr = np.full(npx, 0xFFFFFFFF, np.uint32)
    pos = np.array([0x10000, 0x40000, 0x40000, 0x10000 * 1000 - 0x30000, 0x40000, 0x40000], np.uint32)
    return dict(bounds=b, ior=ior, translucency=tr, pos=pos.reshape(2, 3), dir=d.reshape(2, 3),
                invscale=[2.0, 2.0, 2.0], iterations=1000000, min_brightness=0)


# known answers of scaling_test from the unmodified reference CPU build
SCALING_KNOWN = {
    "u32": dict(epos=[65405418, 262144, 262144, 65388, 262144, 262144], edir=[8201, 0, 0, -4092, 0, 0], eit=[46734, 46623]),
    "f32": dict(epos=[65405502, 262144, 262144, 63693, 262144, 262144], edir=[32.0002, 8e-8, 0, -15.9999, 8e-8, 0], eit=[46718, 46656]),
}


def random_scene(shape, seed, kind="f32", opaque_fraction=0.002):
    """ior (+translucency) of a small seeded scene; kind 'f32' or 'u32'."""
    ior = W.ior_random_smooth(shape, seed)
    tr = W.translucency_random(shape, seed ^ 0xABCDEF, opaque_fraction=opaque_fraction)
    if kind == "u32":
        ior = W.ior_to_u32(ior)
    return ior, tr


def random_rays(shape, n, seed, dir_kind="f32", scale=1.0):
    dim = len(shape)
    lo, hi = 1.25, min(shape) - 2.25
    pos, d = W.rays_random(n, lo, hi, seed, dim=dim)
    # per-axis upper bound differs: rescale each axis into its own range
    u = W.uniform01(seed ^ 0x1234, n * dim).reshape(n, dim)
    for a in range(dim):
        pos[:, a] = W.to_fixed(1.25 + (shape[a] - 3.5) * u[:, a])
    d = (d * scale).astype(np.float32)
    if dir_kind == "i16":
        d = W.dirs_to_i16(d)
    return pos, d
