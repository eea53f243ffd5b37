"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box; the same
code runs on gloo/CPU tensors in the tests).

Rays are independent, so the path shards with no per-step collective:
  * the staged volume is built once (rank `src`) and replicated with ONE broadcast -- the B200 replacement for the
    reference's per-device host upload (cuda_volume_raytracer.cu:676-686);
  * the ray batch is split into `world` contiguous chunks (the reference hands out 32 768-ray chunks dynamically,
    cu:820-821); every rank marches its chunk and results land at the rays' original indices.
"""
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
