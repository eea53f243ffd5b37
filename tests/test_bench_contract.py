"""The bench.py contract that can be checked without a GPU: the reference arm runs on the host cores and prints ONE JSON
line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-window", "32"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray-steps/sec" and d["unit"] == "G ray-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_sass_profile_matches_the_shipped_binary():
    """bench.py's issue roofline multiplies block execution counts by the SASS lengths of the marcher's blocks.  The committed
    profiles/r02_sass_blocks.json (and the listing next to it) must be the ones of the library that is built from this tree: the tool
    is re-run on libvrt_b200.so and the opcode-stream hashes compared, so a kernel change without refreshed profiles fails here."""
    import shutil
    import pytest
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("CUDA binary tools not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_blocks
    committed = json.load(open(os.path.join(ROOT, "profiles", "r02_sass_blocks.json")))
    live = sass_blocks.analyse(sass_blocks.disassemble(os.path.join(ROOT, "volumeraytracer_b200", "libvrt_b200.so"), committed["kernel"]))
    assert live["sass_sha1"] == committed["sass_sha1"], "profiles/r02_sass_blocks.json is stale: re-run tools/sass_blocks.py"
    assert live["blocks"] == committed["blocks"]
    # the fast loop of the all-clear kernel: 64 instructions per step, 27 per cell reload
    assert committed["blocks"]["fast_step"] <= 64 and committed["blocks"]["reload"] <= 27
