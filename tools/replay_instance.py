"""Replay a reference `debug_*_instance` dump on the B200 marcher -- the counterpart of the reference's raytrace_test CLI
(src/raytrace_test.cpp:33-114).   usage: replay_instance.py scene.bin rays.bin   |   replay_instance.py instance.bin"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import volumeraytracer_b200 as vrt                      # noqa: E402
from volumeraytracer_b200 import instance_io as io      # noqa: E402


def main(argv):
    if len(argv) >= 3:
        scene = io.read_scene(open(argv[1], "rb")); rays = io.read_rays(open(argv[2], "rb"))
    else:
        inst = io.read_instance(open(argv[1], "rb")); scene = rays = inst
    sc = vrt.RaytraceScene([int(b) for b in scene["bound_vec"]], scene["ior"], scene["translucency"])
    ep, ed, ei, li, path = sc.trace_rays(rays["start_position"], rays["start_direction"], rays["invscale"], rays["minimum_brightness"],
                                         rays["iterations"], trace_path=rays["trace_path"])
    print("begin", " ".join(str(int(v)) for v in np.asarray(rays["start_position"]).ravel()))
    print("end", " ".join(str(int(v)) for v in ep.ravel()))
    print("iterations", " ".join(str(int(v)) for v in ei))


if __name__ == "__main__":
    main(sys.argv)
