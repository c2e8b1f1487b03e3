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


@pytest.mark.gpu
def test_product_arm_line_has_the_contract_keys():
    """Structure of the one JSON line of the GPU arm at a toy size (no timing thresholds): the driver's keys, the
    roofline / e2e objects, and the step = passes x device batch bookkeeping."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--frames", "32", "--steps", "2", "--warmup", "3",
                        "--passes-per-step", "2", "--legs", "none", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "radar frames/sec polar->pose" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert d["passes_per_step"] == 2 and abs(d["ms_per_step"] - 2 * d["ms_per_pass"]) <= 1e-9 * d["ms_per_step"]
    cfg = d["config"]
    assert cfg["frames_per_step_per_gpu"] == 64 and cfg["pairs_per_step_per_gpu"] == 62 and cfg["frames_per_pass"] == 32 and "workload" in cfg
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "frames/s" and e["d2h_bytes_per_step"] == 2 * 31 * 120
    assert e["h2d_bytes_per_step"] >= 2 * 32 * 400 * (11 + 1997)            # every scan of both passes is uploaded
    ro = d["roofline"]
    assert ro["bound"] == "hbm" and ro["unit"] == "GB/s" and ro["peak"] > 0 and abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-12
    assert ro["alg_bytes_per_launch"] == 32 * (400 * 1997 + 1996 * 1996 + 998 * 998) or ro["kernel"] != "scan_to_l0l1"
    # interleave, scan, 2 pyrDown, klt, compact, adjacency, maxclique, viable, adjseq, clique, kabsch, finish per pass (no MDS)
    assert d["gpu_launches"] == 2 * 2 * 13
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
