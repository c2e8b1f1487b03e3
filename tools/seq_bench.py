#!/usr/bin/env python
"""Chained odometry on the device (rf_seq): frames/s for S sequences in lock step, and single-sequence latency.

    python tools/seq_bench.py [--seq 256] [--runners 4] [--steps 16] [--res 0.0438] [--mds 1] [--graph 1]

Layout: one long synthetic drive of S + steps frames is uploaded once; sequence s starts at frame s and step t reads
frame s + t (arena stride 1), so every sequence is a genuine contiguous drive and the scans are resident in HBM."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cached_sequence(n_frames, res, distort, seed=1234):
    from radarslampy_b200 import synthetic as S
    d = os.path.join("/tmp", "radarfe_cache")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"seq_{seed}_{n_frames}_{res}_{int(distort)}.npy")
    if os.path.exists(p):
        return np.load(p)
    raw, _ = S.make_sequence(n_frames, res_m=res, world=S.World(seed=seed), distort=distort)
    np.save(p, raw)
    return raw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=256)
    ap.add_argument("--runners", type=int, default=4)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=float, default=0.0438)
    ap.add_argument("--mds", type=int, default=1)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--detector", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--retrack", type=int, default=None, help="ablation: rf_config.retrack_threshold (-1 = never re-detect)")
    ap.add_argument("--node-limit", type=int, default=None, help="ablation: rf_config.clique_node_limit")
    a = ap.parse_args()
    from radarslampy_b200 import _ffi
    S, NG, T = a.seq, a.runners, a.warmup + a.steps
    per = S // NG
    S = per * NG          # the sequences that actually run
    raw = cached_sequence(S + T + 1, a.res, bool(a.mds))
    rb = int(87.5 / a.res)
    cfg = _ffi.default_config()
    cfg.range_bins, cfg.cart_res_m, cfg.dist_thr_px = rb, 2 * a.res, 0.5 / (2 * a.res)
    cfg.max_features, cfg.max_pairs, cfg.max_frames = 320, 2, 2
    if a.retrack is not None:
        cfg.retrack_threshold = a.retrack
    if a.node_limit is not None:
        cfg.clique_node_limit = a.node_limit
    fe = _ffi.RadarFE(cfg, device=0)
    runners = [fe.new_sequences(per, per + T + 1, detector_mode=a.detector) for _ in range(NG)]
    for g, r in enumerate(runners):
        r.upload(0, raw[g * per:g * per + per + T + 1])
    for r in runners:
        r.reset(0, 1)
    fe.sync()
    for t in range(1, a.warmup + 1):
        for r in runners:
            r.step(t, 1, with_mds=bool(a.mds), graph=bool(a.graph))
    fe.sync()
    n0 = fe.launch_count()
    fe.timer_start()
    for t in range(a.warmup + 1, T + 1):
        for r in runners:
            r.step(t, 1, with_mds=bool(a.mds), graph=bool(a.graph))
    ms = fe.timer_stop_ms()
    launches = fe.launch_count() - n0
    recs = np.concatenate([np.stack([r.results(t) for t in range(a.warmup + 1, T + 1)], 1) for r in runners])   # [S, steps]
    out = {"sequences": S, "runners": NG, "steps": a.steps, "ms_per_step": ms / a.steps, "frames_per_s": S * a.steps / (ms * 1e-3),
           "launches_per_step": launches / a.steps, "graph": bool(a.graph), "mds": bool(a.mds), "detector_mode": a.detector,
           "retrack_fraction": float(recs["retrack"].mean()), "median_tracked": float(np.median(recs["n_tracked"])),
           "median_features_in": float(np.median(recs["n_features_in"])), "median_candidates": float(np.median(recs["n_candidates"][recs["retrack"] > 0])) if recs["retrack"].any() else 0.0,
           "status_nonzero": int((recs["status"] != 0).sum()), "median_step_m": float(np.median(np.hypot(recs["h"][..., 0], recs["h"][..., 1]))),
           "median_dtheta": float(np.median(np.arctan2(recs["R"][..., 2], recs["R"][..., 0])))}
    for r in runners:
        r.close()
    # single-sequence latency: one runner, one sequence, graph replay, no host sync between steps
    one = fe.new_sequences(1, T + 2, detector_mode=a.detector)
    one.upload(0, raw[:T + 2])
    one.reset(0, 1)
    for t in range(1, a.warmup + 1):
        one.step(t, 1, with_mds=bool(a.mds), graph=True)
    fe.sync()
    fe.timer_start()
    for t in range(a.warmup + 1, T + 1):
        one.step(t, 1, with_mds=bool(a.mds), graph=True)
    ms1 = fe.timer_stop_ms()
    r1 = np.stack([one.results(t) for t in range(a.warmup + 1, T + 1)], 1)
    out["single_sequence"] = {"ms_per_frame": ms1 / a.steps, "launches_per_step": one.launches_per_step,
                              "retrack_fraction": float(r1["retrack"].mean())}
    one.close()
    fe.close()
    s = json.dumps(out)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
