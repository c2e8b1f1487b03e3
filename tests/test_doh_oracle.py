"""CPU: the restatement of the reference's detector (oracle/doh_restate.py = skimage 0.19.2 blob_doh, PARITY UNPINNED —
scikit-image is not in the reference tree nor in this image) checked where it CAN be checked here: against the SciPy
routines blob_doh bottoms out in (scipy.ndimage.maximum_filter incl. its NaN behaviour, cKDTree pair queries), against
closed-form properties of the box filters, and for the blob-recovery property the detector exists for."""
import numpy as np
import pytest

from oracle import doh_restate as D


def _blob_image(n, blobs, dtype=np.float32):
    yy, xx = np.mgrid[:n, :n]
    img = np.zeros((n, n))
    for r, c, s in blobs:
        img += np.exp(-((yy - r) ** 2 + (xx - c) ** 2) / (2 * s * s))
    return img.astype(dtype)


def test_ring_buffer_maximum_filter_matches_scipy_including_nan():
    from scipy import ndimage as ndi
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 7))
        line = rng.normal(size=n)
        line[rng.random(n) < 0.3] = np.nan
        want = ndi.maximum_filter1d(line, 3, mode="nearest")
        ext = [line[0]] + list(line) + [line[-1]]
        got = np.array(D.ring_max_1d(ext, 3))
        assert np.array_equal(got, want, equal_nan=True), (line, got, want)


@pytest.mark.parametrize("params", [(0.01, 10, 3), (1, 6, 4), (0.01, 4, 2), (2, 2, 1)])
def test_peak_mask_restatement_matches_scipy_maximum_filter(params):
    rng = np.random.default_rng(3)
    import cv2
    img = cv2.GaussianBlur(rng.random((70, 90)).astype(np.float32), (0, 0), 2.5)
    cube = D.hessian_cube(img, *params)
    a = D.peak_local_max_3d(cube, 1e-8, use_scipy=True)
    b = D.peak_local_max_3d(cube, 1e-8, use_scipy=False)
    assert np.array_equal(a, b)
    assert len(a) > 3 or params == (0.01, 4, 2)           # [NaN, M1] filters to [NaN, NaN]: no blob at all with two scales
    if params[0] == 0.01 and len(a):
        # the zero-size box: an all-NaN first plane, and under mode='nearest' blobs only from the LAST scale
        assert np.isnan(cube[..., 0]).all() and np.isfinite(cube[..., 1:]).all()
        assert set(a[:, 2]) == {params[2] - 1}


def test_box_geometry_follows_c_integer_division():
    assert D.box_geometry(0.01)[:4] == (0, 0, 0, 0) and np.isinf(D.box_geometry(0.01)[4])      # (0 - 1) / 2 == 0 in C
    assert D.box_geometry(5.005) == (15, 5, 7, 15, 1.0 / 15 / 15)
    assert D.box_geometry(10.0) == (30, 10, 14, 30, 1.0 / 30 / 30)
    assert list(D.sigma_list(0.01, 10, 3)) == [0.01, 5.005, 10.0]


def test_hessian_det_of_a_constant_image_is_zero_inside_and_integral_is_exact():
    img = np.full((80, 80), 0.5, np.float32)
    ii = D.integral_image(img)
    assert ii[-1, -1] == 0.5 * 80 * 80 and ii[3, 7] == 0.5 * 4 * 8
    det = D.hessian_det(ii, 5.005)
    assert np.all(det[20:60, 20:60] == 0.0)              # mid - 3 * side cancels exactly on a constant image


def test_blobs_are_recovered_at_their_scale():
    truth = [(40, 40, 4.0), (40, 120, 6.0), (110, 50, 8.0), (115, 118, 5.0)]
    img = _blob_image(160, truth)
    blobs = D.blob_doh(img, min_sigma=2, max_sigma=12, num_sigma=11, threshold=0.002)
    assert len(blobs) == len(truth)
    sig = {}
    for r, c, s in truth:
        d = np.hypot(blobs[:, 0] - r, blobs[:, 1] - c)
        assert d.min() <= 2.0
        sig[s] = blobs[d.argmin(), 2]
    assert sig[4.0] < sig[5.0] < sig[6.0] <= sig[8.0]     # the detected scale grows with the blob


def test_prune_orders_agree_without_chains_and_differ_only_on_chains():
    # isolated overlapping pairs: both orders keep the same blobs
    lm = np.array([[10, 10, 5.0], [12, 11, 5.0], [80, 80, 5.0], [200, 40, 10.0], [203, 41, 5.0]])
    a, b = D.prune_blobs(lm, 0.5, "sorted"), D.prune_blobs(lm, 0.5, "scipy")
    assert np.array_equal(a, b)
    assert a.tolist() == [[12, 11, 5.0], [80, 80, 5.0], [200, 40, 10.0]]   # equal sigma: the first of the pair goes; else the smaller
    # a chain A~B~C (A, C apart): ascending pair order removes A then B
    chain = np.array([[10, 10, 5.0], [10, 15, 5.0], [10, 20, 5.0]])
    assert D._blob_overlap(chain[0], chain[1]) > 0.5 and D._blob_overlap(chain[0], chain[2]) <= 0.5
    assert D.prune_blobs(chain, 0.5, "sorted").tolist() == [[10, 20, 5.0]]


def test_detector_on_a_real_scan(golden):
    """tests/golden/tiny_frames.npz raw scan 0 (data/tiny): the restated blob_doh with the reference's parameters."""
    from oracle import ref_pipeline as P
    raw = golden["tiny_frames"]["raw_0"]
    cart = P.polar_to_cart(P.extract_polar(raw))
    cube = D.hessian_cube(cart[700:1300, 700:1300], 0.01, 10, 3)
    lm = D.peak_local_max_3d(cube, 0.0005)
    assert len(lm) > 5 and set(lm[:, 2]) == {2}
    blobs = D.blob_doh(cart[700:1300, 700:1300], 0.01, 10, 3, 0.0005)
    assert 0 < len(blobs) <= len(lm) and np.all(blobs[:, 2] == 10.0)
