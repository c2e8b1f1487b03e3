#!/usr/bin/env python
"""KLT tracks/s on BASELINE.json configs[4] (2000^2 grid, 10 000 features, 4 pyramid levels), next to cv2 on one core.

    python tools/klt_stress.py [--features 10000] [--reps 10]

One JSON line: GPU tracks/s through rf_klt (host points in, host results out: H2D + kernel + D2H per call), the
kernel-only rate from CUDA events, and cv2.calcOpticalFlowPyrLK tracks/s on the host."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--features", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import cv2
    from radarslampy_b200 import _ffi
    from test_gpu_stress import N, _scene
    cfg = _ffi.default_config()
    cfg.range_bins, cfg.downsample = N + 1, 2
    cfg.raw_width = cfg.meta_bytes + cfg.range_bins
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 2, 1, 2048
    fe = _ffi.RadarFE(cfg, device=0)
    a, b = _scene(5), _scene(5, shift=(1.7, -0.9))
    fa, fb = fe.frame_from_cart(a), fe.frame_from_cart(b)
    cand, n = fe.detect(fa, -0.01, cap=200000)
    t0 = time.perf_counter()
    sel = fe.ssc(cand, args.features, 0.1, N, N)
    t_ssc = time.perf_counter() - t0
    pts = np.ascontiguousarray(cand[sel][:, [1, 0]], np.float32)
    fe.klt(fa, fb, pts, apply_err_gate=True)
    t0 = time.perf_counter()
    fe.timer_start()
    for _ in range(args.reps):
        nxt, st, err = fe.klt(fa, fb, pts, apply_err_gate=True)
    ms_dev = fe.timer_stop_ms() / args.reps
    dt = (time.perf_counter() - t0) / args.reps
    u8a, u8b = fa.download(1), fb.download(1)
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    cv2.calcOpticalFlowPyrLK(u8a, u8b, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    dt_cv1 = time.perf_counter() - t0
    cv2.setNumThreads(0)
    t0 = time.perf_counter()
    cv2.calcOpticalFlowPyrLK(u8a, u8b, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    dt_cvN = time.perf_counter() - t0
    K = len(pts)
    print(json.dumps({"metric": "KLT tracks/s (dense stress: 2000^2, 4 levels)", "features": K, "candidates": int(n),
                      "gpu_tracks_per_s_e2e": K / dt, "gpu_tracks_per_s_device": K / (ms_dev * 1e-3), "gpu_ms_per_call": 1e3 * dt,
                      "ssc_ms": 1e3 * t_ssc, "cv2_tracks_per_s_1thread": K / dt_cv1, "cv2_tracks_per_s_all_threads": K / dt_cvN,
                      "cv2_threads": cv2.getNumThreads(), "tracked_fraction": float(st.mean())}))
    fe.close()


if __name__ == "__main__":
    main()
