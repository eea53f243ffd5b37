"""world_size-2 gloo test of the multi-GPU host logic on CPU tensors: volume broadcast, contiguous ray sharding,
results back at the rays' original indices.  The per-rank "march" is the CPU oracle here (this is a test of the
sharding plumbing, not of the kernel)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_rays, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from tests import scenes as S
    from volumeraytracer_b200 import dist as vd
    shape = (22, 20, 24)
    ob = [s - 2 for s in shape]
    nvox = int(np.prod(ob))
    vol_t = torch.zeros(nvox * 4, dtype=torch.float32)
    pos_t = d_t = None
    if rank == 0:
        ior, tr = S.random_scene(shape, seed=7, kind="f32")
        _, _, planes, trc = orc.prep(shape, ior, tr)
        vol = orc.fold(planes, trc)
        vol_t.copy_(torch.from_numpy(vol.reshape(-1)))
        pos, d = S.random_rays(ob, n_rays, seed=3)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
        pos_t = torch.from_numpy(pos.view(np.int32).reshape(-1)); d_t = torch.from_numpy(d.reshape(-1))
    vd.broadcast_volume(vol_t, src=0)                                   # one collective at scene creation
    my_pos, my_dir = vd.scatter_rays(n_rays, 3, pos_t, d_t, torch.float32, torch.device("cpu"), src=0)
    lo, hi = vd.chunk_bounds(n_rays, world, rank)
    assert my_pos.numel() == (hi - lo) * 3
    ep, ed, ei, li, _ = orc.trace(vol_t.numpy().reshape(-1, 4), ob, my_pos.numpy().view(np.uint32), my_dir.numpy(), [1, 1, 1], 200)
    g_pos = vd.gather_results(n_rays, 3, torch.from_numpy(ep.view(np.int32).reshape(-1)))
    g_it = vd.gather_results(n_rays, 1, torch.from_numpy(ei.view(np.int32)))
    if rank == 0:
        want = orc.trace(vol, ob, pos, d, [1, 1, 1], 200)
        ok = np.array_equal(g_pos.numpy().view(np.uint32).reshape(-1, 3), want[0]) and np.array_equal(g_it.numpy().view(np.uint32), want[2])
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [1001, 64])
def test_two_rank_sharding_gloo(tmp_path, n_rays):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), n_rays, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_chunk_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from volumeraytracer_b200 import dist as vd
    for n in (0, 1, 7, 16777216, 1001):
        for w in (1, 2, 4, 8):
            b = [vd.chunk_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _shared_worker(rank, world, port, n_rays, tag, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from tests import scenes as S
    from volumeraytracer_b200 import dist as vd
    shape = (22, 20, 24)
    ob = [s - 2 for s in shape]
    ior, tr = S.random_scene(shape, seed=7, kind="f32")          # every rank can build the tiny scene itself; what is tested is the batch
    _, _, planes, trc = orc.prep(shape, ior, tr)
    vol = orc.fold(planes, trc)
    batch = vd.SharedBatch(tag, n_rays, 3, np.float32, create=(rank == 0)) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        batch = vd.SharedBatch(tag, n_rays, 3, np.float32, create=False)
    pos, d = S.random_rays(ob, n_rays, seed=3)
    pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
    if rank == 0:                                                 # ONE caller-owned batch
        batch.arrays["pos"][:] = pos.reshape(-1); batch.arrays["dir"][:] = d.reshape(-1)
    dist.barrier()
    p_s, d_s, ep, ed, ei, li = batch.slices(world, rank)          # this rank's chunk, in place
    res = orc.trace(vol, ob, np.asarray(p_s), np.asarray(d_s), [1, 1, 1], 200)
    ep[:] = res[0].reshape(-1); ed[:] = res[1].reshape(-1); ei[:] = res[2]; li[:] = res[3]
    dist.barrier()
    if rank == 0:
        want = orc.trace(vol, ob, pos, d, [1, 1, 1], 200)
        a = batch.arrays
        ok = (np.array_equal(a["epos"].reshape(-1, 3), want[0]) and np.array_equal(a["edir"].reshape(-1, 3), want[1])
              and np.array_equal(a["eit"], want[2]) and np.array_equal(a["light"], want[3]))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    batch.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [1001, 3])
def test_one_shared_batch_sharded_in_place_gloo(tmp_path, n_rays):
    """Strong-scaling plumbing of bench.py: ONE batch in shared host arrays, each rank traces its contiguous chunk in place."""
    out = str(tmp_path / "result.txt")
    tag = "vrt_test_%d_%d" % (os.getpid(), n_rays)
    mp.spawn(_shared_worker, args=(2, _free_port(), n_rays, tag, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
    assert not [f for f in os.listdir("/dev/shm") if f.startswith(tag)]


def test_ray_grid_rows_are_a_contiguous_slice():
    sys.path.insert(0, ROOT)
    from volumeraytracer_b200 import workloads as W
    full_p, full_d = W.rays_parallel_x(37, 29, 2.0, 60.0, x0=2.0)
    for j0, j1 in ((0, 10), (10, 37), (36, 37)):
        p, d = W.rays_parallel_x(37, 29, 2.0, 60.0, x0=2.0, rows=(j0, j1))
        assert np.array_equal(p, full_p[j0 * 29:j1 * 29]) and np.array_equal(d, full_d[j0 * 29:j1 * 29])
