#!/usr/bin/env python
"""Scan-file ingest rate: rf_ingest_png on all host cores vs cv2.imread one file at a time.

    python tools/ingest_bench.py [--files 256]

Writes synthetic Oxford-shaped scans as PNG to a temp directory, then prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=256)
    args = ap.parse_args()
    import cv2
    from radarslampy_b200 import _ffi, synthetic as S
    raw, _ = S.make_sequence(16, res_m=0.0438, world=S.World(seed=1))
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(args.files):
            p = os.path.join(d, f"{1547131046353776 + 250000 * i}.png")
            cv2.imwrite(p, raw[i % len(raw)])
            paths.append(p)
        size = np.mean([os.path.getsize(p) for p in paths])
        buf = _ffi.pinned_empty((args.files, raw.shape[1], raw.shape[2]), np.uint8)
        _ffi.ingest_png(paths[:8], out=buf)
        t0 = time.perf_counter()
        out = _ffi.ingest_png(paths, threads=0, out=buf)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        _ffi.ingest_png(paths[:32], threads=1, out=buf)
        dt1 = (time.perf_counter() - t0) / 32
        cv2.setNumThreads(1)
        t0 = time.perf_counter()
        ref = [cv2.imread(p, cv2.IMREAD_GRAYSCALE) for p in paths[:32]]
        dtc = (time.perf_counter() - t0) / 32
        ok = all(np.array_equal(ref[i], raw[i % len(raw)]) for i in range(32)) and np.array_equal(out[40], raw[40 % len(raw)])
    print(json.dumps({"metric": "scan files decoded per second", "files": args.files, "mean_file_bytes": float(size),
                      "host_cores": os.cpu_count(), "ingest_all_cores_fps": args.files / dt, "ingest_1_thread_fps": 1 / dt1,
                      "cv2_imread_fps": 1 / dtc, "identical": bool(ok)}))


if __name__ == "__main__":
    main()
