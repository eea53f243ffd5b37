"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libvrt_ref.so, built from
/root/reference/src by oracle/Makefile) on small seeded inputs.  Run in the build container (where /root/reference
exists):   python tests/golden/make_golden.py
The fixtures let tests/test_golden.py pin the oracle where the reference cannot be built (e.g. the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref            # noqa: E402
from tests import scenes as S     # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def scaling():
    """reference scaling_test (cuda_volume_raytracer_test.h:4-75), both instantiations; path subsampled every 97th entry."""
    for kind in ("u32", "f32"):
        inp = S.scaling_test_inputs(kind)
        sc = ref.RefScene(inp["bounds"], inp["ior"], inp["translucency"])
        ep, ed, ei, li, pa = sc.trace(inp["pos"], inp["dir"], inp["invscale"], 0, inp["iterations"], trace_path=True)
        save("scaling_test_" + kind, end_position=ep, end_direction=ed, end_iteration=ei, remaining_light=li,
             path_every_97=pa[:, ::97, :].copy(), diff0_x_profile=sc.diff(0).reshape(sc.diff_bounds.astype(int))[:, 3, 3].copy())


def api_level():
    """RaytraceScene end to end on seeded random scenes, 3-D and 2-D, float and int16 instantiations."""
    for kind, dirk in (("f32", "f32"), ("u32", "i16")):
        for shape in ((24, 20, 28), (33, 17)):
            ior, tr = S.random_scene(shape, seed=11 + len(shape), kind=kind, opaque_fraction=0.01)
            pos, d = S.random_rays(shape, 1500, seed=5, dir_kind=dirk)
            sc = ref.RefScene(shape, ior, tr)
            isc = [1.0, 0.75, 1.5][:len(shape)]
            ep, ed, ei, li, pa = sc.trace(pos, d, isc, 0, 400, trace_path=False)
            save("api_%s_%dd" % (kind, len(shape)), end_position=ep, end_direction=ed, end_iteration=ei, remaining_light=li,
                 volume=sc.interleaved())


def live():
    """the live-translucency instantiation of trace_rays_cpu (cu:337-341)."""
    from oracle import oracle as orc
    for volk, dirk in (("f32", "f32"), ("i16", "i16")):
        shape = (20, 26, 22)
        ior, tr = S.random_scene(shape, seed=21, kind="f32" if volk == "f32" else "u32", opaque_fraction=0.003)
        sc = ref.RefScene(shape, ior, tr)
        trc = sc.translucency_cropped()
        trc[trc != 0] -= np.uint32(1 << 24)
        vol = sc.interleaved()
        ob = [int(b) for b in sc.diff_bounds]
        pos, d = S.random_rays(ob, 1500, seed=10, dir_kind=dirk, scale=1.1)
        pos = pos - np.uint32(0x10000) + np.uint32(0x4000)
        ep, ed, ei, li, pa = ref.trace_live(vol, trc, ob, [1, 1, 1], pos, d, 300, 0x40000000)
        save("live_%s" % volk, end_position=ep, end_direction=ed, end_iteration=ei, remaining_light=li, volume=vol,
             translucency=trc, start_position=pos, start_direction=d)


if __name__ == "__main__":
    scaling(); api_level(); live()
