"""Child process of tests/test_round2.py::test_nccl_scene_broadcast_two_processes: rank <r> of <world>, one GPU each.
usage: python tests/_nccl_worker.py RANK WORLD OUTDIR"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import volumeraytracer_b200 as vrt
    from oracle import oracle as orc
    from tests import scenes as S

    uid_file = os.path.join(outdir, "uid.bin")

    def exchange(uid):                       # the unique id travels through a file: the C ABI does not care how
        if uid is not None:
            with open(uid_file + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(uid_file + ".tmp", uid_file)
            return uid
        for _ in range(600):
            if os.path.exists(uid_file):
                return open(uid_file, "rb").read()
            time.sleep(0.1)
        raise RuntimeError("no unique id")

    comm = vrt.Comm(rank, rank, world, exchange)
    scene = None
    shape = (37, 41, 35)
    if rank == 0:
        ior, tr = S.random_scene(shape, seed=71, kind="u32", opaque_fraction=0.01)      # int16 scene: the staged copy is widened
        scene = vrt.TraceRaysCu.from_ior(shape, ior, tr)
        scene.set_option(vrt.VRT_OPT_STEPS_PER_POLL, 96)
    mine, secs = comm.broadcast_scene(scene, root=0)
    assert mine.device == rank and mine.get_option(vrt.VRT_OPT_STEPS_PER_POLL) == 96
    ob = mine._output_sizes
    pos, d = S.random_rays(ob, 30000, seed=5, dir_kind="i16")
    pos = pos - np.uint32(0x10000) + np.uint32(0x777)
    epos, edir, eit, light, _ = mine.trace_rays_cu(pos, d, [1.0, 1.0, 1.0], 0x40000000, 400, live_translucency=True)
    vol, trc = mine.download_volume()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), vol=vol, tr=trc, epos=epos, edir=edir, eit=eit, light=light, secs=secs)
    # the normalise step works on the replica too (ior travels with the scene)
    import torch
    dev = torch.device("cuda", rank)
    p_api, d_api = S.random_rays(shape, 1000, seed=9, dir_kind="i16")
    tp = torch.from_numpy(p_api.view(np.int32).reshape(-1)).to(dev); td = torch.from_numpy(d_api.reshape(-1)).to(dev)
    mine.normalise_rays_device(tp, td)
    comm.close()
    mine.close()


if __name__ == "__main__":
    main()
