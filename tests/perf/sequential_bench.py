#!/usr/bin/env python
"""Single-sequence (replica) mode: the reference's SYSTEM loop run frame by frame through the drop-in modules
(radarslampy_b200.odometry.run_odometry: every call NumPy in / NumPy out, f32 Cartesian image returned to the host
each frame, FMT prior, point cloud, keyframes, MDS) next to the same loop made of the reference's library calls on
the CPU (oracle/ref_system.py; + FMT and point-cloud extraction per frame, as RawROAMSystem.run does).

    python tests/perf/sequential_bench.py [--frames 24]

One JSON line with frames/s of both and the largest per-frame pose difference.  This is BASELINE configs[0]/[1]
as a latency-bound sequential run (pose chaining does not shard); the batch path in bench.py is the throughput mode."""
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
    ap.add_argument("--frames", type=int, default=24)
    args = ap.parse_args()
    from oracle import ref_pipeline as P
    from oracle import ref_system
    from radarslampy_b200 import odometry, synthetic as S
    from radarslampy_b200.getFeatures import appendNewFeatures
    raw, _ = S.make_sequence(args.frames, res_m=0.0432, world=S.World(seed=4321))
    odometry.run_odometry(raw[:3])                                    # warm-up: handles, scratch, pinned buffers
    t0 = time.perf_counter()
    got = odometry.run_odometry(raw)
    dt_gpu = time.perf_counter() - t0
    det_calls = []

    def detect(cart, old):                                           # the shared detector (parity unpinned), not timed on the CPU side
        t = time.perf_counter()
        out = appendNewFeatures(cart, old)[0]
        det_calls.append(time.perf_counter() - t)
        return out

    t0 = time.perf_counter()
    want = ref_system.run_odometry(raw, detect)
    dt_cpu = time.perf_counter() - t0 - sum(det_calls)
    # what RawROAMSystem.run adds per frame on the CPU: the FMT prior and the point cloud of every keyframe update
    polar = [P.extract_polar(r) for r in raw[:5]]
    t0 = time.perf_counter()
    for i in range(4):
        P.rotation_fmt(polar[i], polar[i + 1])
    dt_fmt = (time.perf_counter() - t0) / 4
    n = args.frames - 1
    d = got["traj"].poses - want["poses"]
    print(json.dumps({"metric": "sequential odometry frames/s (system loop through the drop-in API)", "frames": args.frames,
                      "gpu_frames_per_s": n / dt_gpu, "gpu_ms_per_frame": 1e3 * dt_gpu / n,
                      "cpu_loop_frames_per_s": n / dt_cpu, "cpu_ms_per_frame": 1e3 * dt_cpu / n,
                      "cpu_fmt_ms_per_frame_extra": 1e3 * dt_fmt, "host_cores": os.cpu_count(),
                      "max_pose_diff_m": float(np.abs(d[:, :2]).max()), "max_pose_diff_rad": float(np.abs(d[:, 2]).max()),
                      "tracked_equal": bool(np.array_equal(got["n_tracked"], want["n_tracked"])),
                      "keyframes": len(got["map"].keyframes)}))


if __name__ == "__main__":
    main()
