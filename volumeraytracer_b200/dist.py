"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box; the same
code runs on gloo/CPU tensors in the tests).

Rays are independent, so the path shards with no per-step collective:
  * the staged volume is built once (rank `src`) and replicated with ONE broadcast -- the B200 replacement for the
    reference's per-device host upload (cuda_volume_raytracer.cu:676-686);
  * the ray batch is split into `world` contiguous chunks (the reference hands out 32 768-ray chunks dynamically,
    cu:820-821); every rank marches its chunk and results land at the rays' original indices.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def chunk_bounds(n_rays, world, rank):
    """[lo, hi) of rank's contiguous chunk; chunk sizes differ by at most one ray."""
    lo = n_rays * rank // world
    hi = n_rays * (rank + 1) // world
    return lo, hi


def broadcast_volume(volume, src=0, group=None):
    """Replicate the staged interleaved volume (any dtype, flat) from `src` into every rank's buffer."""
    dist.broadcast(volume, src=src, group=group)
    return volume


def scatter_rays(n_rays, dim, pos, d, dir_dtype, device, src=0, group=None):
    """Rank `src` holds the whole batch (pos [n*dim] int32 bit patterns, d [n*dim]); every rank receives its contiguous
    chunk.  Chunks are padded to a common length for the collective and trimmed on arrival."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [chunk_bounds(n_rays, world, r)[1] - chunk_bounds(n_rays, world, r)[0] for r in range(world)]
    m = max(sizes) * dim
    my_pos = torch.empty(m, dtype=torch.int32, device=device)
    my_dir = torch.empty(m, dtype=dir_dtype, device=device)
    if rank == src:
        lp, ld = [], []
        for r in range(world):
            lo, hi = chunk_bounds(n_rays, world, r)
            cp = torch.zeros(m, dtype=torch.int32, device=device); cp[:(hi - lo) * dim] = pos[lo * dim:hi * dim]
            cd = torch.zeros(m, dtype=dir_dtype, device=device); cd[:(hi - lo) * dim] = d[lo * dim:hi * dim]
            lp.append(cp); ld.append(cd)
        dist.scatter(my_pos, lp, src=src, group=group)
        dist.scatter(my_dir, ld, src=src, group=group)
    else:
        dist.scatter(my_pos, None, src=src, group=group)
        dist.scatter(my_dir, None, src=src, group=group)
    k = sizes[rank] * dim
    return my_pos[:k].contiguous(), my_dir[:k].contiguous()


def gather_results(n_rays, per_ray, local, dst=0, group=None):
    """Inverse of scatter_rays for one output array with `per_ray` elements per ray: rank `dst` gets the results at the
    rays' original indices (others get None)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [chunk_bounds(n_rays, world, r)[1] - chunk_bounds(n_rays, world, r)[0] for r in range(world)]
    m = max(sizes) * per_ray
    buf = torch.zeros(m, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        return torch.cat([parts[r][:sizes[r] * per_ray] for r in range(world)])
    dist.gather(buf, None, dst=dst, group=group)
    return None


class SharedBatch:
    """ONE ray batch in ONE set of host arrays shared by all ranks of the node (a file per array under /dev/shm, mapped by
    every rank).  This is how one-process-per-GPU ranks serve the reference's calling convention -- one caller-owned batch,
    results written in place at the rays' original indices (cuda_volume_raytracer.cu:804-821 splits one batch over the
    devices the same way) -- without a gather: rank r traces `slices(r)` of the shared arrays through vrt_trace and the
    results are where the caller expects them.  Arrays: pos [n*dim] u32, dir [n*dim] f32|i16, epos, edir, eit [n], light [n]."""

    FIELDS = ("pos", "dir", "epos", "edir", "eit", "light")

    def __init__(self, tag, n_rays, dim, dir_dtype, create, directory="/dev/shm"):
        self.n, self.dim, self.dir_dtype = int(n_rays), int(dim), np.dtype(dir_dtype)
        self.paths = {f: os.path.join(directory, "%s_%s.bin" % (tag, f)) for f in self.FIELDS}
        shapes = {"pos": (self.n * dim, np.uint32), "dir": (self.n * dim, self.dir_dtype), "epos": (self.n * dim, np.uint32),
                  "edir": (self.n * dim, self.dir_dtype), "eit": (self.n, np.uint32), "light": (self.n, np.uint32)}
        self.arrays = {}
        for f, (count, dt) in shapes.items():
            if create:
                with open(self.paths[f], "wb") as fh:
                    fh.truncate(max(count, 1) * np.dtype(dt).itemsize)
            self.arrays[f] = np.memmap(self.paths[f], dtype=dt, mode="r+", shape=(count,))
        self.owner = bool(create)

    def slices(self, world, rank):
        """rank's views (pos, dir, epos, edir, eit, light) of the shared arrays: its contiguous chunk, no copy."""
        lo, hi = chunk_bounds(self.n, world, rank)
        d = self.dim
        a = self.arrays
        return (a["pos"][lo * d:hi * d], a["dir"][lo * d:hi * d], a["epos"][lo * d:hi * d], a["edir"][lo * d:hi * d], a["eit"][lo:hi], a["light"][lo:hi])

    def close(self):
        self.arrays = {}
        if self.owner:
            for p in self.paths.values():
                try:
                    os.unlink(p)
                except OSError:
                    pass
