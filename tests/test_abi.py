"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/vrt_b200.h declares, fails
loudly (no fallback) without a CUDA device, and the product never touches oracle/."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vrt_b200.h")).read()
    return sorted(set(re.findall(r"VRT_API\s+[\w\s\*]+?\b(vrt_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import volumeraytracer_b200 as vrt
    from volumeraytracer_b200 import _lib
    lib = vrt.lib()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libvrt_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared, "python binding list and header disagree"
    assert b"sm_100a" in lib.vrt_version()


def test_argument_validation_needs_no_gpu():
    import volumeraytracer_b200 as vrt
    from volumeraytracer_b200 import _lib
    lib = vrt.lib()
    h = C.c_void_p()
    b = np.array([4, 4, 4, 4], dtype=np.uint64)
    assert lib.vrt_scene_create(C.byref(h), 0, 4, b.ctypes.data_as(C.c_void_p), 0, None, None, 0) == _lib.VRT_ERR_INVALID
    assert lib.vrt_trace(None, 0, None, None, 0, None, 0, 0, 0, None, None, None, None, None) == _lib.VRT_ERR_INVALID
    assert b"scene is null" in lib.vrt_last_error()
    assert lib.vrt_scene_destroy(None) == 0
    assert lib.vrt_scene_set_option(None, 0, 0) == _lib.VRT_ERR_INVALID
    g = C.c_double()
    assert lib.vrt_measure_gather_bandwidth(0, 16, 32, 1, C.byref(g)) == _lib.VRT_ERR_INVALID


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import volumeraytracer_b200 as vrt
    planes = [np.zeros(64, np.float32)] * 3
    with pytest.raises(vrt.VrtError):
        vrt.TraceRaysCu([4, 4, 4], planes, np.zeros(64, np.uint32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "volumeraytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert not re.search(r'#include\s*[<"].*vrt_oracle', src), f
                assert "libvrt_oracle" not in src and "libvrt_ref" not in src, f


def test_error_behaviour_mirrors_the_reference():
    """Argument errors surface as exceptions with the reference's wording where it has one (cu:768-771,
    image_util.cpp:508-515,558,738-741) -- checked without a GPU because validation happens before any CUDA call."""
    import volumeraytracer_b200 as vrt
    with pytest.raises(vrt.VrtError, match="Illegal dimension"):
        vrt.TraceRaysCu([4, 4, 4, 4], [np.zeros(256, np.float32)] * 4, np.zeros(256, np.uint32))
    with pytest.raises(ValueError):
        vrt.TraceRaysCu([4, 4, 4], [np.zeros(64, np.float32)] * 2, np.zeros(64, np.uint32))
    with pytest.raises(ValueError, match="imagesizes"):
        vrt.TraceRaysCu([4, 4, 4], [np.zeros(63, np.float32)] * 3, np.zeros(64, np.uint32))
    with pytest.raises(vrt.VrtError, match="65535"):
        vrt.TraceRaysCu([70000, 2, 2], [np.zeros(280000, np.float32)] * 3, np.zeros(280000, np.uint32))
    with pytest.raises(vrt.VrtError, match="dimension is zero"):
        vrt.RaytraceScene([], np.zeros(0, np.float32), np.zeros(0, np.uint32))
    with pytest.raises(vrt.VrtError, match="Illegal dimension"):
        vrt.RaytraceScene([5, 5, 5, 5], np.ones(625, np.float32), np.zeros(625, np.uint32))
    with pytest.raises(TypeError):
        vrt.TraceRaysCu([4, 4, 4], [np.zeros(64, np.float64)] * 3, np.zeros(64, np.uint32))


def test_python_constants_match_the_header():
    """Every enumerator of include/vrt_b200.h that the ctypes module mirrors has the same value there."""
    import re
    from volumeraytracer_b200 import _lib
    text = open(os.path.join(ROOT, "include", "vrt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    seen = 0
    for name, expr in re.findall(r"\b(VRT_[A-Z0-9_]+)\s*=\s*([^,}\n]+)", text):
        expr = expr.strip().replace("u", "")
        value = eval(expr, {"__builtins__": {}})           # "1 << 3", "-2", "100"
        if hasattr(_lib, name):
            assert getattr(_lib, name) == value, "%s: header %r, _lib.py %r" % (name, value, getattr(_lib, name))
            seen += 1
    assert seen >= 20
