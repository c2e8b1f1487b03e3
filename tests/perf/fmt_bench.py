#!/usr/bin/env python
"""FMT rotation prior throughput on the resident scans of a batch, next to the CPU path.

    python tests/perf/fmt_bench.py [--frames 256] [--reps 5]

Prints one JSON line: pairs/s on the GPU (rf_batch_fmt over the batch's u8 scans, synchronous call incl. the
D2H of the results), pairs/s of the reference's cv2 calls (oracle/ref_pipeline.rotation_fmt) on one core, and
the largest |angle| difference between the two on the sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-pairs", type=int, default=16)
    args = ap.parse_args()
    from radarslampy_b200 import _ffi, synthetic as S
    from oracle import ref_pipeline as P
    world = S.World(seed=1234)
    raw, poses = S.make_sequence(args.frames, res_m=0.0432, world=world, first=0)
    F = args.frames
    cfg = _ffi.default_config()
    cfg.max_frames, cfg.max_pairs, cfg.max_features = F, F - 1, 64
    fe = _ffi.RadarFE(cfg, device=0)
    b = fe.new_batch()
    pairs = np.stack([np.arange(F - 1), np.arange(1, F)], 1).astype(np.int32)
    b.upload(raw, pairs, np.zeros((F - 1, 64, 2), np.float32), np.zeros(F - 1, np.int32))
    clip = int(87.5 / 0.0864)
    ang, sc, resp, sh = b.fmt_rotation(10, clip)      # warm-up (scratch allocation)
    n0 = fe.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ang, sc, resp, sh = b.fmt_rotation(10, clip)
    dt = (time.perf_counter() - t0) / args.reps
    launches = (fe.launch_count() - n0) // args.reps
    n_cpu = min(args.cpu_pairs, F - 1)
    polar = [P.extract_polar(raw[i]) for i in range(n_cpu + 1)]
    t0 = time.perf_counter()
    cpu = [P.rotation_fmt(polar[i], polar[i + 1])[0] for i in range(n_cpu)]
    dt_cpu = (time.perf_counter() - t0) / n_cpu
    print(json.dumps({"metric": "FMT rotation prior pairs/s", "gpu_pairs_per_s": (F - 1) / dt, "gpu_ms_per_batch": 1e3 * dt,
                      "frames": F, "launches_per_batch": int(launches), "cpu_pairs_per_s_1core": 1.0 / dt_cpu,
                      "max_abs_angle_diff_rad": float(np.abs(np.array(cpu) - ang[:n_cpu]).max()),
                      "median_angle_rad": float(np.median(ang)), "expected_rad_per_frame": 0.025}))
    b.close()
    fe.close()


if __name__ == "__main__":
    main()
