"""GPU parity: pyramidal LK (a4) + err gating (a5) against the reference-generated goldens
(getTrackedPointsKLT on data/tiny), live cv2 and the C restatement.
Tolerance (north_star): positions within 0.02 px, identical status."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_PX = 0.02


def _frames(fe, golden, idxs):
    g = golden["tiny_frames"]
    return [fe.polar_to_cart(raw=g[f"raw_{i}"], want_host=False)[0] for i in idxs]


@pytest.mark.parametrize("pair", [0, 1])
def test_klt_matches_reference_goldens(fe, golden, pair):
    st = golden["tiny_stages"]
    f0, f1 = _frames(fe, golden, [pair, pair + 1])
    pts = st[f"feat_in_{pair}"]
    nxt, status, err = fe.klt(f0, f1, pts, apply_err_gate=True)
    ref_status = st[f"klt_status_{pair}"]
    assert np.array_equal(status, ref_status)
    good = status.ravel().astype(bool)
    assert np.abs(nxt[good] - st[f"klt_good_new_{pair}"]).max() <= TOL_PX
    assert np.array_equal(pts[good], st[f"klt_good_old_{pair}"])
    assert np.abs(nxt[~good] - st[f"klt_bad_new_{pair}"]).max() <= TOL_PX


def test_klt_matches_cv2_and_oracle_including_borders(fe, golden):
    import cv2
    from oracle import restate as R
    f0, f1 = _frames(fe, golden, [0, 1])
    u0, u1 = f0.download(1), f1.download(1)
    rng = np.random.default_rng(0)
    gf = cv2.goodFeaturesToTrack(u0, 400, 0.01, 10).reshape(-1, 2)
    pts = np.vstack([gf, rng.uniform(0, 2024, (300, 2)), rng.uniform(-3, 12, (40, 2)), rng.uniform(2010, 2026, (40, 2)),
                     [[0, 0], [2023, 2023], [3.5, 1000], [2020.2, 5.1], [1012, 1012]]]).astype(np.float32)
    nxt, status, err = fe.klt(f0, f1, pts, apply_err_gate=False)
    c_nxt, c_st, c_err = cv2.calcOpticalFlowPyrLK(u0, u1, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    o_nxt, o_st, o_err = R.pyr_lk(u0, u1, pts)
    assert np.array_equal(status.ravel(), c_st.ravel())
    assert np.array_equal(status.ravel(), o_st)
    g = status.ravel().astype(bool)
    assert g.sum() > 300
    assert np.abs(nxt[g] - c_nxt[g]).max() <= TOL_PX
    assert np.abs(nxt[g] - o_nxt[g]).max() <= 2e-3          # same integer arithmetic, only the f32 sums differ
    assert np.abs(err.ravel()[g] - c_err.ravel()[g]).max() <= 1e-3
    assert np.abs(nxt[~g] - c_nxt[~g]).max() <= TOL_PX       # lost points keep the propagated position


def test_klt_identity_and_shift(fe):
    """Property at full size: tracking a frame against itself returns the input points."""
    rng = np.random.default_rng(3)
    from scipy.ndimage import gaussian_filter
    img = gaussian_filter(rng.random((2024, 2024)).astype(np.float32), 2.0)
    img = (img - img.min()) / (img.max() - img.min())
    f = fe.frame_from_cart(img)
    pts = rng.uniform(30, 1990, (1000, 2)).astype(np.float32)
    nxt, status, err = fe.klt(f, f, pts, apply_err_gate=True)
    g = status.ravel().astype(bool)
    assert g.mean() > 0.9
    assert np.abs(nxt[g] - pts[g]).max() <= 1e-3
    assert err.ravel()[g].max() == 0.0


def test_klt_small_images_every_window_near_a_border():
    """500 x 500 frames (pyramid 500 / 250 / 125 / 63): at the upper levels almost every 16 x 16 / 18 x 18 window touches
    the image border, so the staging switches between the aligned-word fast path and the REFLECT_101 path all the time.
    Status identical and positions within 2e-3 px of the C restatement (same integer arithmetic)."""
    from oracle import restate as R
    from radarslampy_b200 import _ffi
    from scipy.ndimage import gaussian_filter, shift as nd_shift
    rng = np.random.default_rng(21)
    cfg = _ffi.default_config()
    cfg.azimuths, cfg.range_bins, cfg.raw_width = 200, 501, 11 + 501
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 2, 1, 64
    fe = _ffi.RadarFE(cfg)
    try:
        a = gaussian_filter(rng.random((500, 500)).astype(np.float32), 1.5)
        a = ((a - a.min()) / (a.max() - a.min())).astype(np.float32)
        b = np.clip(nd_shift(a, (0.8, -1.3), order=1, mode="nearest"), 0, 1).astype(np.float32)
        fa, fb = fe.frame_from_cart(a), fe.frame_from_cart(b)
        pts = np.vstack([rng.uniform(-2, 502, (400, 2)), rng.uniform(0, 20, (60, 2)), rng.uniform(480, 500, (60, 2)),
                         [[0, 0], [499, 499], [7.5, 250.25], [492.4, 8.6], [250, 250]]]).astype(np.float32)
        nxt, st, err = fe.klt(fa, fb, pts, apply_err_gate=False)
        o_nxt, o_st, o_err = R.pyr_lk(fa.download(1), fb.download(1), pts)
        assert np.array_equal(st.ravel(), o_st)
        g = o_st.astype(bool)
        assert g.sum() > 300
        assert np.abs(nxt[g] - o_nxt[g]).max() <= 2e-3
        assert np.abs(err.ravel()[g] - o_err[g]).max() <= 1e-3
        if (~g).any():
            assert np.abs(nxt[~g] - o_nxt[~g]).max() <= TOL_PX
    finally:
        fe.close()
