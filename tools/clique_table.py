#!/usr/bin/env python
"""Clique-search time on REAL graphs: the 10 data/tiny pairs (tests/golden/tiny_stages.npz klt_good_*) and the
reference's outlier_test.npz fixture (139 nodes, 8 097 maximal cliques), GPU (rf_reject_outliers, device time by CUDA
events incl. the adjacency kernel) beside networkx on one host core (outlierRejection.py:62-78).
    python tools/clique_table.py [--out profiles/r02_clique_table.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nx_first_max(prev, new, thr):
    import networkx as nx
    from scipy.spatial.distance import cdist
    t0 = time.perf_counter()
    m = (np.abs(cdist(prev, prev) - cdist(new, new)) <= thr).astype(np.int8)
    G = nx.Graph(m)
    best, n = [], 0
    for c in nx.find_cliques(G):
        n += 1
        if len(c) > len(best):
            best = c
    dt = time.perf_counter() - t0
    return best, n, dt, float((m.sum() - len(m)) / max(1, len(m) * (len(m) - 1)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from radarslampy_b200 import _ffi
    fe = _ffi.RadarFE(device=0)
    st = np.load(os.path.join(ROOT, "tests", "golden", "tiny_stages.npz"))
    fx = np.load(os.path.join(ROOT, "tests", "golden", "clique_fixture.npz"))
    cases = [(f"tiny pair {i}", st[f"klt_good_old_{i}"], st[f"klt_good_new_{i}"], st[f"rej_mask_{i}"]) for i in range(10)]
    cases.append(("outlier_test.npz", fx["prev"], fx["new"], fx["mask"]))
    thr = 0.5 / 0.0864
    rows = []
    for name, prev, new, want in cases:
        fe.reject_outliers(prev, new)                       # warm-up
        best_ms = 1e9
        for _ in range(5):
            fe.timer_start()
            mask, n_in, nodes = fe.reject_outliers(prev, new)
            best_ms = min(best_ms, fe.timer_stop_ms())
        clique, n_cliques, cpu_s, dens = nx_first_max(prev, new, thr)
        ok = bool(np.array_equal(mask, want.astype(bool)))
        rows.append({"case": name, "nodes": int(len(prev)), "edge_density": round(dens, 3), "maximal_cliques": n_cliques,
                     "clique_size": int(n_in), "gpu_ms": round(best_ms, 4), "gpu_descents": int(nodes),
                     "networkx_ms": round(cpu_s * 1e3, 2), "mask_exact": ok})
        print(rows[-1])
    out = {"rows": rows, "worst_tiny_ms": max(r["gpu_ms"] for r in rows[:10]), "fixture_ms": rows[-1]["gpu_ms"],
           "all_exact": all(r["mask_exact"] for r in rows),
           "note": "gpu_ms = CUDA events around rf_reject_outliers (H2D of the points, k_adjacency, k_maxclique, k_clique, D2H of the mask)"}
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    fe.close()


if __name__ == "__main__":
    main()
