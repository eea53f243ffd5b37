"""Memory-safety evidence without compute-sanitizer (it is closed on the GPU pool): the parity tests run once more, in a child process, on
volumeraytracer_b200/libvrt_b200_check.so -- the same sources compiled with -DVRT_CHECK, where every computed index of a gather, of a
ray-buffer access, of a path write and of the wavefront marcher's state arrays is tested against the size of what it indexes and
violations are counted on the device (csrc/vrt_march.cuh, VRT_CHK; a deliberately failing check is the negative control).  The count must be 0 after the fuzz, edge-value, type-combination,
2-D, path, brick-layout, host-rounding and wavefront tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "volumeraytracer_b200", "libvrt_b200_check.so")

CHILD = r"""
import ctypes as C, sys
import pytest
rc = pytest.main(["tests/test_cuda_parity.py", "tests/test_round2.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                  "-k", "fuzz or edge or type_combinations or two_dimensional or step_count or bricked or empty_space or round_host_all or round_host_2d "
                        "or wavefront or probe or normalise or chunk_schedule or in_place or concurrent"])
import volumeraytracer_b200 as vrt
lib = vrt.lib()
n = C.c_uint64(0)
lib.vrt_check_violations.restype = C.c_int
assert lib.vrt_check_violations(0, C.byref(n)) == 0
before = n.value
assert lib.vrt_check_selftest(0) == 0 and lib.vrt_check_violations(0, C.byref(n)) == 0      # negative control: one failing check is counted
print("CONTROL %d" % (n.value - before))
n.value = before
print("PYTEST_RC %d" % int(rc))
print("LAUNCHES %d" % vrt.launch_count())
print("VIOLATIONS %d" % n.value)
"""


def test_parity_tests_on_the_bounds_checked_build_count_no_violation():
    if not os.path.exists(CHECK_LIB):
        pytest.skip("libvrt_b200_check.so not built (make -C volumeraytracer_b200/csrc check)")
    env = dict(os.environ, VRT_B200_LIB=CHECK_LIB)
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    out = r.stdout
    assert r.returncode == 0, out[-3000:] + r.stderr[-2000:]
    assert "PYTEST_RC 0" in out, out[-3000:]
    launches = int(out.split("LAUNCHES ")[1].split()[0])
    assert launches > 100, "the checked library was not the one under test"
    assert "CONTROL 1" in out, "the violation counter does not count: " + out[-500:]
    assert "VIOLATIONS 0" in out, out[-500:]
