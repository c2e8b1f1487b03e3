"""CPU tests: pin the oracle (oracle/restate.py + oracle/c/oracle_c.c) against
  (1) the goldens generated from the UNMODIFIED reference (tests/golden/*.npz, made by
      oracle/gen_golden.py in the authoring container), and
  (2) the live third-party libraries the reference calls (cv2, scipy, networkx).
Nothing here touches a GPU or /root/reference."""
import hashlib

import numpy as np
import pytest

from oracle import restate as R

RES = 0.0864


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tiny(golden):
    g = golden["tiny_frames"]
    out = []
    for i in range(2):
        raw = g[f"raw_{i}"]
        polar = R.extract_polar(raw, 2025)
        cart = R.warp_polar(polar)
        out.append((raw, polar, cart, R.to_u8(cart)))
    return out


# ---- a1/a2/a3 ------------------------------------------------------------------------
def test_extract_and_warp_match_reference_sha(golden, tiny):
    g = golden["tiny_frames"]
    for i, (raw, polar, cart, u8) in enumerate(tiny):
        assert polar.shape == (400, 2025) and polar.dtype == np.float32
        assert sha(polar) == str(g[f"polar_sha256_{i}"])
        assert cart.shape == (2024, 2024)
        assert sha(cart) == str(g[f"cart_sha256_{i}"])        # bit-exact with the reference's cv2.warpPolar
        assert sha(u8) == str(g[f"u8_sha256_{i}"])


def test_warp_matches_live_cv2_on_random_scan():
    import cv2
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, (400, 3779), dtype=np.uint8)
    polar = R.extract_polar(raw, 2025)
    assert np.array_equal(polar, raw[:, 11:11 + 2025].astype(np.float32) / 255.)
    ref = cv2.warpPolar(polar, (2024, 2024), (1012.0, 1012.0), 1012,
                        cv2.WARP_POLAR_LINEAR + cv2.WARP_INVERSE_MAP + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS)
    assert np.array_equal(R.warp_polar(polar), ref)


def test_warp_other_geometry_matches_cv2():
    """BASELINE config resolution 0.0438 m/bin -> 1997 bins, R=998."""
    import cv2
    rng = np.random.default_rng(8)
    polar = (rng.integers(0, 256, (400, 1997)).astype(np.float32) / 255.)
    ref = cv2.warpPolar(polar, (1996, 1996), (998.0, 998.0), 998,
                        cv2.WARP_POLAR_LINEAR + cv2.WARP_INVERSE_MAP + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS)
    assert np.array_equal(R.warp_polar(polar), ref)


def test_metadata_decode(golden):
    raw = golden["tiny_frames"]["raw_0"]
    ts, az, valid = R.extract_metadata(raw)
    assert ts.shape == (400, 1) and ts.dtype == np.int64 and np.all(np.diff(ts.ravel()) > 0)
    assert az.dtype == np.float32 and 0 <= az.min() and az.max() < 2 * np.pi + 1e-6
    assert valid.dtype == bool


# ---- a4 ------------------------------------------------------------------------------
def test_pyramid_and_scharr_match_cv2(tiny):
    import cv2
    u8 = tiny[0][3]
    _, pyr = cv2.buildOpticalFlowPyramid(u8, (15, 15), 3, withDerivatives=False, pyrBorder=cv2.BORDER_REFLECT_101,
                                         derivBorder=cv2.BORDER_CONSTANT, tryReuseInputImage=False)
    lvl = u8
    for l in range(4):
        pad = pyr[l]
        inner = pad[15:-15, 15:-15] if pad.shape[0] > lvl.shape[0] else pad
        assert np.array_equal(inner, lvl), f"level {l}"
        lvl = R.pyr_down(lvl)
    sch = R.scharr(u8[:300, :400].copy())
    assert np.array_equal(sch[..., 0], cv2.Scharr(u8[:300, :400], cv2.CV_16S, 1, 0))
    assert np.array_equal(sch[..., 1], cv2.Scharr(u8[:300, :400], cv2.CV_16S, 0, 1))


def test_pyr_lk_matches_cv2_and_reference_goldens(golden, tiny):
    import cv2
    st = golden["tiny_stages"]
    u0, u1 = tiny[0][3], tiny[1][3]
    pts = st["feat_in_0"]
    nxt, status, err = R.pyr_lk(u0, u1, pts)
    c_nxt, c_st, c_err = cv2.calcOpticalFlowPyrLK(u0, u1, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    assert np.array_equal(status, c_st.ravel())
    g = status.astype(bool)
    assert np.abs(nxt[g] - c_nxt[g]).max() <= 2e-3
    assert np.abs(err[g] - c_err.ravel()[g]).max() <= 5e-3   # err moves with the sub-1e-3 px position noise
    # reference gating: status &= err < 10  (getTransformKLT.py:365)
    gated = (status.astype(bool) & (err < 10)).astype(np.uint8)
    assert np.array_equal(gated.reshape(-1, 1), st["klt_status_0"])
    assert np.abs(nxt[gated.astype(bool)] - st["klt_good_new_0"]).max() <= 2e-3


def test_pyr_lk_border_points(tiny):
    import cv2
    u0, u1 = tiny[0][3], tiny[1][3]
    rng = np.random.default_rng(0)
    pts = np.vstack([rng.uniform(-3, 12, (40, 2)), rng.uniform(2010, 2026, (40, 2)), rng.uniform(0, 2024, (100, 2)),
                     [[0, 0], [2023, 2023], [3.5, 1000], [2020.2, 5.1], [1012, 1012]]]).astype(np.float32)
    nxt, status, err = R.pyr_lk(u0, u1, pts)
    c_nxt, c_st, _ = cv2.calcOpticalFlowPyrLK(u0, u1, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    assert np.array_equal(status, c_st.ravel())
    assert np.abs(nxt - c_nxt).max() <= 2e-3


# ---- a6 ------------------------------------------------------------------------------
def test_adjacency_matches_scipy(golden):
    from scipy.spatial.distance import cdist
    g = golden["clique_fixture"]
    thr = float(g["thr"])
    assert thr == 0.5 / 0.0864
    ref = (np.abs(cdist(g["prev"], g["prev"]) - cdist(g["new"], g["new"])) <= thr).astype(np.uint8)
    assert np.array_equal(R.consistency_adjacency(g["prev"], g["new"], thr), ref)


def _nx_first_max(adj):
    import networkx as nx
    G = nx.Graph(adj.astype(np.int8))
    best, n, ys = [], 0, []
    for c in nx.find_cliques(G):
        n += 1
        ys.append(list(c))
        if len(c) > len(best):
            best = list(c)
    return best, n, ys


def test_clique_order_matches_live_networkx():
    rng = np.random.default_rng(5)
    for K, p in [(1, .5), (2, .5), (3, .9), (5, .5), (9, .3), (17, .5), (19, .6), (33, .4), (40, .7), (64, .5), (77, .35),
                 (100, .3), (129, .25), (150, .3), (200, .2), (90, .05)]:
        M = rng.random((K, K)) < p
        M = np.triu(M, 1)
        M = (M | M.T)
        np.fill_diagonal(M, True)
        adj = M.astype(np.uint8)
        best, n, ys = _nx_first_max(adj)
        total = sum(len(y) + 1 for y in ys)
        clique, ny, oys = R.first_max_clique(adj, want_yields=total + 8)
        assert ny == n and oys == ys, f"K={K} p={p}: enumeration order differs from networkx"
        assert clique == best
        pruned, _ = R.first_max_clique_pruned(adj)
        assert sorted(pruned) == sorted(best)


def test_reject_outliers_reference_fixture(golden):
    g = golden["clique_fixture"]
    _, _, mask = R.reject_outliers(g["prev"], g["new"], float(g["thr"]))
    assert mask.sum() == 67 and np.array_equal(mask, g["mask"])


def test_reject_outliers_tiny_goldens(golden):
    st = golden["tiny_stages"]
    for i in range(int(st["n_pairs"])):
        _, _, mask = R.reject_outliers(st[f"klt_good_old_{i}"], st[f"klt_good_new_{i}"], 0.5 / 0.0864)
        assert np.array_equal(mask, st[f"rej_mask_{i}"]), f"pair {i}"


# ---- a7 / a8 -------------------------------------------------------------------------
def test_kabsch_vs_reference(golden):
    st = golden["tiny_stages"]
    for i in range(int(st["n_pairs"])):
        Rm, h = R.kabsch(st[f"svd_src_{i}"], st[f"svd_tgt_{i}"])
        assert np.abs(h - st[f"svd_h_{i}"]).max() * RES <= 1e-4
        assert abs(np.arctan2(Rm[1, 0], Rm[0, 0]) - np.arctan2(st[f"svd_R_{i}"][1, 0], st[f"svd_R_{i}"][0, 0])) <= 1e-5
    g = golden["kabsch_fixture"]
    Rm, h = R.kabsch(g["src"], g["tgt"])
    assert np.abs(h - g["h"]).max() * RES <= 1e-4


def test_mds_vs_reference(golden):
    st = golden["tiny_stages"]
    for i in range(0, int(st["n_pairs"]), 3):
        x = R.mds_solve(st[f"mds_Twj0_{i}"], st[f"mds_pw_{i}"], st[f"mds_pjt_{i}"], st[f"mds_Twj_{i}"])
        ref = st[f"mds_x_{i}"]
        assert np.abs(x[3:5] - ref[3:5]).max() <= 1e-4 and abs(x[5] - ref[5]) <= 1e-5


# ---- a9 / a12 ------------------------------------------------------------------------
def test_ssc_matches_reference(golden):
    g = golden["ssc_fixture"]
    for j in range(int(g["n"])):
        kp, ref = g[f"kp_{j}"], g[f"sel_{j}"]
        k, W, H = (int(v) for v in g[f"cfg_{j}"])
        sel = R.ssc(kp, k, 0.1, W, H)
        assert np.array_equal(kp[sel], ref), f"fixture {j}"


def test_polar_peaks_match_reference_and_scipy(golden, tiny):
    from scipy.signal import find_peaks
    polar = tiny[0][1]
    got = R.polar_peaks(polar)
    assert np.array_equal(got, golden["peaks_fixture"]["peaks"].astype(np.int64))
    row = np.array([0, 1, 1, 1, 0, 2, 3, 3, 2, 5, 5, 1, 4], np.float32)
    assert find_peaks(row)[0].tolist() == [2, 6, 9]


# ---- a10 detector (structure tensor) -----------------------------------------------------
def test_corner_response_and_selection_match_live_cv2(tiny):
    """oracle restatement of cv2.cornerMinEigenVal (<= 2e-6 abs) and of cv2.goodFeaturesToTrack's
    candidate rule + ordering (exact, given cv2's own response map)."""
    import cv2
    cart = tiny[0][2]
    ref = cv2.cornerMinEigenVal(cart, 3, ksize=3)
    assert np.abs(R.corner_min_eig(cart) - ref).max() <= 2e-6
    pts = cv2.goodFeaturesToTrack(cart, 0, 0.01, 0, blockSize=3, useHarrisDetector=False).reshape(-1, 2)
    sel = R.nms_select(ref, float(ref.max()) * 0.01)
    assert len(sel) == len(pts) > 1000
    assert np.array_equal(sel[:, 1], pts[:, 0]) and np.array_equal(sel[:, 0], pts[:, 1])


# ---- the toolchain the goldens were recorded with ------------------------------------------
def test_golden_toolchain_versions():
    """Clique order is CPython-set-layout and networkx specific (oracle/c/oracle_c.c restates CPython 3.12's
    setobject.c and networkx 3.x's find_cliques): the goldens and the live-library pins of this file are only
    comparable under the interpreter / wheels that recorded them (tests/golden/VERSIONS.json)."""
    import json
    import os
    from oracle import toolchain_versions as T
    assert os.path.exists(T.PATH), "tests/golden/VERSIONS.json missing: run python -m oracle.toolchain_versions"
    rec, cur = json.load(open(T.PATH)), T.current()
    mm = lambda v: tuple(v.split(".")[:2])
    assert mm(rec["cpython"]) == mm(cur["cpython"]), f"goldens recorded under CPython {rec['cpython']}, running {cur['cpython']}: set iteration order may differ"
    assert mm(rec["networkx"])[0] == mm(cur["networkx"])[0], f"goldens recorded with networkx {rec['networkx']}, running {cur['networkx']}"
    for k in ("scipy", "opencv", "numpy"):
        assert mm(rec[k])[0] == mm(cur[k])[0], f"goldens recorded with {k} {rec[k]}, running {cur[k]}"
