"""CPU: the bench.py contract that can be checked without a GPU — the reference arm prints ONE JSON line on stdout
with the keys the driver reads, the product arm refuses to run without a CUDA device (no CPU fallback), and the
algorithmic-byte accounting matches DESIGN.md §4."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    pytest.importorskip("cv2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "6", "--steps", "1",
                        "--warmup", "1", "--features", "80"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "radar frames/sec polar->pose" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--frames", "4", "--steps", "1", "--warmup", "1",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]


def test_algorithmic_bytes_match_design():
    sys.path.insert(0, ROOT)
    import bench
    sb = bench.stage_bytes({"azimuths": 400, "range_bins": 1997, "n": 1996}, 256, 255, 255 * 200, False, 4, fused=True)
    assert sb["scan_to_l0l1"] == 256 * (400 * 1997 + 1996 * 1996 + 998 * 998) == 256 * 5778820
    assert sb["polar2cart"] == 256 * 2 * 400 * 1997
    assert sb["pyr_down"] == 256 * (998 * 998 + 499 * 499 + 499 * 499 + 250 * 250)
