"""GPU parity of feature selection: SSC (ANMS.py:5-102) bit-exact against the reference-generated
goldens and the C oracle; threshold + 3x3 NMS selection bit-exact given a response map; the
structure-tensor response against live cv2.cornerMinEigenVal (tolerance 2e-6 absolute: cv2's SIMD
path contracts differently in the last bit); polar peak extraction (getPointCloud.py:11-54) exact."""
import numpy as np
import pytest

from oracle import restate as R

pytestmark = pytest.mark.gpu


# ---- a9 SSC ----------------------------------------------------------------------------
def test_ssc_matches_reference_goldens(fe, golden):
    g = golden["ssc_fixture"]
    for j in range(int(g["n"])):
        kp, want = g[f"kp_{j}"], g[f"sel_{j}"]
        k, W, H = (int(v) for v in g[f"cfg_{j}"])
        sel = fe.ssc(kp, k, 0.1, W, H)
        assert np.array_equal(kp[sel], want), f"fixture {j}"


@pytest.mark.parametrize("n,k,W,H,seed", [(0, 200, 2024, 2024, 0), (5, 200, 2024, 2024, 1), (150, 200, 640, 480, 2),
                                          (3000, 200, 2024, 2024, 3), (60000, 10000, 2000, 2000, 4),
                                          (40000, 500, 2024, 1000, 5), (2000, 50, 300, 200, 6)])
def test_ssc_matches_oracle_random(fe, n, k, W, H, seed):
    rng = np.random.default_rng(seed)
    kp = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), rng.choice([0.01, 5.005, 10.0], n)], axis=1).astype(np.float64)
    if n and seed % 2:
        kp[:, :2] += rng.uniform(0, 0.999, (n, 2))          # sub-pixel keypoints
    assert np.array_equal(fe.ssc(kp, k, 0.1, W, H), R.ssc(kp, k, 0.1, W, H))


def test_ssc_raster_ordered_input(fe):
    """worst case for the serial part: spatially sorted keypoints (every selection is a dependent step)"""
    ys, xs = np.mgrid[0:500:3, 0:500:3]
    kp = np.stack([ys.ravel(), xs.ravel(), np.ones(ys.size)], axis=1).astype(np.float64)
    assert np.array_equal(fe.ssc(kp, 300, 0.1, 500, 500), R.ssc(kp, 300, 0.1, 500, 500))


def test_ssc_rejects_what_the_reference_raises_on(fe):
    kp = np.array([[10.0, 5000.0, 1.0]])
    with pytest.raises(ValueError):
        fe.ssc(kp, 200, 0.1, 2024, 2024)          # IndexError in the reference
    with pytest.raises(ValueError):
        fe.ssc(np.zeros((4, 3)), 1, 0.1, 100, 100)   # ZeroDivisionError in the reference


# ---- a10 detector --------------------------------------------------------------------
@pytest.fixture(scope="module")
def cart0(fe, golden):
    raw = golden["tiny_frames"]["raw_0"]
    frame, cart = fe.polar_to_cart(raw=raw)
    return frame, cart


def test_corner_response_matches_cv2(fe, cart0):
    import cv2
    frame, cart = cart0
    got = fe.corner_response(frame)
    ref = cv2.cornerMinEigenVal(cart, 3, ksize=3)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() <= 2e-6
    assert np.abs(got - R.corner_min_eig(cart)).max() <= 2e-6


def test_nms_selection_bit_exact_given_reference_response(fe, cart0):
    """north_star: feature selection bit-exact when given the reference's corner responses."""
    import cv2
    _, cart = cart0
    ref = cv2.cornerMinEigenVal(cart, 3, ksize=3)
    thr = float(np.float32(ref.max() * 0.01))
    got, n = fe.nms_select(ref, thr)
    want = R.nms_select(ref, thr)
    assert n == len(want) and np.array_equal(got, want)
    # and that rule IS cv2.goodFeaturesToTrack's (minDistance = 0 keeps every candidate, in order)
    pts = cv2.goodFeaturesToTrack(cart, 0, 0.01, 0, blockSize=3, useHarrisDetector=False).reshape(-1, 2)
    thr_cv = float(ref.max()) * 0.01
    got_cv, _ = fe.nms_select(ref, thr_cv)
    if len(got_cv) == len(pts):                   # cv thresholds in double; equal unless a value sits on the edge
        assert np.array_equal(got_cv[:, 1], pts[:, 0]) and np.array_equal(got_cv[:, 0], pts[:, 1])


def test_nms_selection_plateaus_and_borders(fe):
    rng = np.random.default_rng(3)
    resp = rng.integers(0, 4, (64, 96)).astype(np.float32)          # many exact ties and plateaus
    resp[0, :] = 9; resp[:, 0] = 9; resp[-1, :] = 9; resp[:, -1] = 9   # border pixels never qualify
    got, n = fe.nms_select(resp, 0.5)
    want = R.nms_select(resp, 0.5)
    assert n == len(want) and np.array_equal(got, want)
    assert fe.nms_select(resp, 100.0)[1] == 0


def test_detect_end_to_end_consistent(fe, cart0):
    frame, cart = cart0
    resp = fe.corner_response(frame)
    cand, n = fe.detect(frame, -0.01, cap=100000)
    want = R.nms_select(resp, float(np.float32(float(resp.max()) * 0.01)))
    assert n == len(want) and np.array_equal(cand, want)
    capped, n2 = fe.detect(frame, -0.01, cap=100)
    assert n2 == n and np.array_equal(capped, want[:100])


# ---- a12 polar peaks -----------------------------------------------------------------
def test_polar_peaks_match_reference_golden(fe, golden):
    raw = golden["tiny_frames"]["raw_0"]
    polar = R.extract_polar(raw, 2025)
    got = fe.polar_peaks(polar)
    assert got.dtype == np.int64
    assert np.array_equal(got, golden["peaks_fixture"]["peaks"].astype(np.int64))


@pytest.mark.parametrize("A,W,seed", [(400, 2025, 0), (7, 3, 1), (16, 3768, 2), (3, 50, 3)])
def test_polar_peaks_match_oracle_random(fe, A, W, seed):
    rng = np.random.default_rng(seed)
    polar = (rng.integers(0, 12, (A, W)).astype(np.float32) / 255.)   # few levels -> many plateaus
    polar[0, :] = 0.5                                                   # a flat row has no peaks at all
    assert np.array_equal(fe.polar_peaks(polar), R.polar_peaks(polar))


def test_polar_peaks_known_answer(fe):
    row = np.array([[0, 1, 1, 1, 0, 2, 3, 3, 2, 5, 5, 1, 4]], np.float32)   # SURVEY §8 a12 probe: peaks 2, 6, 9
    got = fe.polar_peaks(row)
    want = R.polar_peaks(row)
    assert np.array_equal(got, want) and set(got[:, 1]) <= {2, 6, 9}
