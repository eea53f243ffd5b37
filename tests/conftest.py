import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The libraries are built by __graft_entry__.build(); if a checkout arrives without them (they are git-ignored),
    build what this machine can build (needs nvcc / gcc, no GPU)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "volumeraytracer_b200", "libvrt_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "volumeraytracer_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "libvrt_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def refimpl():
    """The unmodified reference CPU build (oracle/_ref/libvrt_ref.so); skipped where it was never built."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libvrt_ref.so not built (needs /root/reference at build time)")
    ref.lib()
    return ref
