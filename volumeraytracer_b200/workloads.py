"""Synthetic inputs for the five BASELINE.json configurations plus small seeded
random scenes for the parity tests.  Pure numpy: used by tests/ and bench.py on the host; bench.py has
torch twins of the large analytic fields so that 512^3 / 1024^3 volumes are produced on the GPU.

Conventions (reference: image_util.cpp:686, types.h:5): volumes are ior float32 (or uint32 16.16) arrays
with axis 0 slowest; ray positions are uint32 16.16 in API coordinates (original-volume voxels, must
satisfy 1 <= p < bound-1); directions are float32 (or int16 with unit 0x100).
"""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(seed, n):
    """n outputs of SplitMix64 as uint64 (vectorised; identical to the scalar C++ definition)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed, n):
    """float64 in [0,1) from the top 53 bits."""
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def to_fixed(x):
    """voxel coordinate (float) -> uint32 16.16."""
    return np.floor(np.asarray(x, dtype=np.float64) * 65536.0 + 0.5).astype(np.uint32)


def grid_coords(shape):
    return np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")


# ---------------------------------------------------------------------------------------------------
# index fields

def ior_constant(size, n=1.0):
    return np.full((size,) * 3, n, dtype=np.float32)


def ior_luneburg(size=256, radius=100.0):
    """C2: Luneburg-style spherical GRIN lens, centre (size-1)/2, n = sqrt(2 - r^2/R^2) inside R, 1 outside."""
    c = (size - 1) / 2.0
    x, y, z = grid_coords((size,) * 3)
    r2 = (x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2
    n = np.sqrt(np.maximum(2.0 - r2 / np.float32(radius * radius), 1.0))
    return n.astype(np.float32)


def ior_sines(size=512, base=1.3, amp=0.1, period=128.0):
    """C3: n = base + amp sin(2 pi x/P) sin(2 pi y/P) sin(2 pi z/P)."""
    k = np.float32(2.0 * np.pi / period)
    s = np.sin(k * np.arange(size, dtype=np.float32))
    return (np.float32(base) + np.float32(amp) * s[:, None, None] * s[None, :, None] * s[None, None, :]).astype(np.float32)


def translucency_c3(size=512):
    """C3 translucency plane: absorption grows with y, modulated along x, plus an opaque ball r=40 at the centre."""
    x = np.arange(size, dtype=np.float64)
    ax = 1.0 + np.cos(2.0 * np.pi * x / 64.0)
    ay = x / (size - 1.0)
    absorb = np.floor((1 << 21) * ay[None, :, None] * ax[:, None, None]).astype(np.uint64)
    tr = (np.uint64(0xFFFFFFFF) - absorb).astype(np.uint32)
    tr = np.broadcast_to(tr, (size, size, size)).copy()
    c = (size - 1) / 2.0
    xx, yy, zz = grid_coords((size,) * 3)
    ball = (xx - c) ** 2 + (yy - c) ** 2 + (zz - c) ** 2 < np.float32((40.0 * size / 512.0) ** 2)
    tr[ball] = 0
    return tr


def ior_c5(size=1024, base=1.2, amp=0.05, period=256.0):
    """C5: n = base + amp sin(2 pi x/P) cos(2 pi y/P) cos(2 pi z/P)."""
    k = np.float32(2.0 * np.pi / period)
    t = k * np.arange(size, dtype=np.float32)
    s, c = np.sin(t), np.cos(t)
    return (np.float32(base) + np.float32(amp) * s[:, None, None] * c[None, :, None] * c[None, None, :]).astype(np.float32)


def solve_harmonic(size, inner_radius, inner_value=1.6, face_value=1.0, sweeps=300, omega=1.0):
    """C4 field: damped-Jacobi relaxation of Laplace's equation (restated from the behaviour of
    solve_harmonic.cpp:17-119 with derrivative_divisor == 0, max_error == 0): faces fixed at face_value, a
    centred ball fixed at inner_value, everything else starts at face_value; `sweeps` Jacobi sweeps with the
    6-neighbour mean.  Only determinism and smoothness matter for the benchmark, not convergence."""
    v = np.full((size,) * 3, face_value, dtype=np.float32)
    c = (size - 1) / 2.0
    x, y, z = grid_coords((size,) * 3)
    ball = (x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2 < np.float32(inner_radius ** 2)
    fixed = ball.copy()
    fixed[0, :, :] = fixed[-1, :, :] = True
    fixed[:, 0, :] = fixed[:, -1, :] = True
    fixed[:, :, 0] = fixed[:, :, -1] = True
    v[ball] = inner_value
    w = np.float32(omega)
    for _ in range(sweeps):
        m = np.zeros_like(v)
        m[1:-1, 1:-1, 1:-1] = (v[:-2, 1:-1, 1:-1] + v[2:, 1:-1, 1:-1] + v[1:-1, :-2, 1:-1] + v[1:-1, 2:, 1:-1]
                               + v[1:-1, 1:-1, :-2] + v[1:-1, 1:-1, 2:]) * np.float32(1.0 / 6.0)
        v = np.where(fixed, v, v + w * (m - v)).astype(np.float32)
    return v


def ior_random_smooth(shape, seed, lo=1.0, hi=1.6, waves=4):
    """Seeded smooth random field for parity tests: a few random plane waves."""
    dim = len(shape)
    u = uniform01(seed, waves * (dim + 2))
    coords = grid_coords(shape)
    acc = np.zeros(shape, dtype=np.float64)
    for w in range(waves):
        ph = 2 * np.pi * u[w * (dim + 2)]
        amp = 0.5 + u[w * (dim + 2) + 1]
        arg = ph
        for d in range(dim):
            arg = arg + coords[d] * (2 * np.pi * (u[w * (dim + 2) + 2 + d] - 0.5) * 0.25)
        acc += amp * np.sin(arg)
    acc = (acc - acc.min()) / max(acc.max() - acc.min(), 1e-9)
    return (lo + (hi - lo) * acc).astype(np.float32)


def translucency_random(shape, seed, opaque_fraction=0.002, absorb_bits=22):
    """Seeded random translucency plane: mostly clear-ish with random absorption < 2^absorb_bits and a
    sprinkling of fully opaque voxels (tr = 0)."""
    n = int(np.prod(shape))
    r = splitmix64(seed, n)
    absorb = (r & np.uint64((1 << absorb_bits) - 1)).astype(np.uint64)
    tr = (np.uint64(0xFFFFFFFF) - absorb).astype(np.uint32)
    opaque = (r >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24)) < opaque_fraction
    tr[opaque] = 0
    return tr.reshape(shape)


def ior_to_u32(ior_f32):
    return np.floor(np.asarray(ior_f32, dtype=np.float64) * 65536.0 + 0.5).astype(np.uint32)


# ---------------------------------------------------------------------------------------------------
# rays

def rays_parallel_x(ny, nz, lo, hi, x0=2.0, lo_z=None, hi_z=None, rows=None):
    """ny*nz parallel +x rays on a uniform grid over [lo,hi]^2 (C2, C3, C5), z fastest.  rows=(j0, j1) returns only the
    y-rows j0 <= j < j1 of that grid (a contiguous index range of the batch: one rank's shard), with identical bits."""
    lo_z = lo if lo_z is None else lo_z
    hi_z = hi if hi_z is None else hi_z
    ys = lo + (hi - lo) * (np.arange(ny, dtype=np.float64) / max(ny - 1, 1))
    zs = lo_z + (hi_z - lo_z) * (np.arange(nz, dtype=np.float64) / max(nz - 1, 1))
    if rows is not None:
        ys = ys[rows[0]:rows[1]]
        ny = ys.size
    pos = np.empty((ny, nz, 3), dtype=np.uint32)
    pos[..., 0] = to_fixed(x0)
    pos[..., 1] = to_fixed(ys)[:, None]
    pos[..., 2] = to_fixed(zs)[None, :]
    d = np.zeros((ny * nz, 3), dtype=np.float32)
    d[:, 0] = 1.0
    return pos.reshape(-1, 3), d


def rays_random(n, lo, hi, seed, dim=3):
    """n rays, start uniform in [lo,hi]^dim, direction uniform on the unit sphere / circle (C4)."""
    u = uniform01(seed, n * (dim + 2)).reshape(n, dim + 2)
    pos = to_fixed(lo + (hi - lo) * u[:, :dim])
    if dim == 3:
        cz = 2.0 * u[:, 3] - 1.0
        ph = 2.0 * np.pi * u[:, 4]
        s = np.sqrt(np.maximum(1.0 - cz * cz, 0.0))
        d = np.stack([s * np.cos(ph), s * np.sin(ph), cz], axis=1)
    else:
        ph = 2.0 * np.pi * u[:, 2]
        d = np.stack([np.cos(ph), np.sin(ph)], axis=1)
    return pos.astype(np.uint32), d.astype(np.float32)


def dirs_to_i16(d):
    return np.clip(np.rint(np.asarray(d, dtype=np.float64) * 256.0), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------------------
# torch twins of the large analytic fields (produced on the GPU: 512^3 / 1024^3 are too slow in numpy).
# Values can differ from the numpy versions in the last float32 ulp (different sin/cos); whoever needs
# identical bits on the CPU side (oracle, cpu_baseline) downloads the staged volume from the device.

def ior_c5_torch(size, device, base=1.2, amp=0.05, period=256.0):
    import torch
    t = torch.arange(size, dtype=torch.float32, device=device) * float(2.0 * np.pi / period)
    s, c = torch.sin(t), torch.cos(t)
    return (base + amp * s[:, None, None] * c[None, :, None] * c[None, None, :]).contiguous()


def ior_sines_torch(size, device, base=1.3, amp=0.1, period=128.0):
    import torch
    s = torch.sin(torch.arange(size, dtype=torch.float32, device=device) * float(2.0 * np.pi / period))
    return (base + amp * s[:, None, None] * s[None, :, None] * s[None, None, :]).contiguous()


def ior_luneburg_torch(size, device, radius=100.0):
    import torch
    c = (size - 1) / 2.0
    g = torch.arange(size, dtype=torch.float32, device=device) - c
    r2 = (g * g)[:, None, None] + (g * g)[None, :, None] + (g * g)[None, None, :]
    return torch.sqrt(torch.clamp(2.0 - r2 / float(radius * radius), min=1.0)).contiguous()


def translucency_c3_torch(size, device):
    """uint32 plane held in an int32 tensor (bit pattern)."""
    import torch
    x = torch.arange(size, dtype=torch.float64, device=device)
    ax = 1.0 + torch.cos(2.0 * np.pi * x / 64.0)
    ay = x / (size - 1.0)
    absorb = torch.floor((1 << 21) * ay[None, :, None] * ax[:, None, None]).to(torch.int64)
    tr = (0xFFFFFFFF - absorb).expand(size, size, size).clone()
    c = (size - 1) / 2.0
    g = torch.arange(size, dtype=torch.float32, device=device) - c
    r2 = (g * g)[:, None, None] + (g * g)[None, :, None] + (g * g)[None, None, :]
    tr[r2 < (40.0 * size / 512.0) ** 2] = 0
    return torch.where(tr >= (1 << 31), tr - (1 << 32), tr).to(torch.int32).contiguous()


def solve_harmonic_torch(size, device, inner_radius, inner_value=1.6, face_value=1.0, sweeps=300):
    import torch
    v = torch.full((size,) * 3, face_value, dtype=torch.float32, device=device)
    c = (size - 1) / 2.0
    g = torch.arange(size, dtype=torch.float32, device=device) - c
    r2 = (g * g)[:, None, None] + (g * g)[None, :, None] + (g * g)[None, None, :]
    ball = r2 < float(inner_radius ** 2)
    v[ball] = inner_value
    free = ~ball[1:-1, 1:-1, 1:-1]
    for _ in range(sweeps):
        m = (v[:-2, 1:-1, 1:-1] + v[2:, 1:-1, 1:-1] + v[1:-1, :-2, 1:-1] + v[1:-1, 2:, 1:-1]
             + v[1:-1, 1:-1, :-2] + v[1:-1, 1:-1, 2:]) * (1.0 / 6.0)
        inner = v[1:-1, 1:-1, 1:-1]
        v = v.clone()
        v[1:-1, 1:-1, 1:-1] = torch.where(free, m, inner)
    return v.contiguous()


def clear_translucency_torch(shape, device):
    import torch
    return torch.full(tuple(shape), -1, dtype=torch.int32, device=device)
