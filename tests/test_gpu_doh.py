"""GPU: the reference's feature detector on the device (rf_detect_doh, csrc/k_doh.cu = skimage.feature.blob_doh as
getFeatures.getBlobsFromCart calls it, getFeatures.py:13-18,22-53) against its CPU restatement
(oracle/doh_restate.py): integral image and Hessian-determinant planes BIT FOR BIT, the blob list identical, on real
scans (data/tiny) with the reference's parameters and on random images with other scale lists.

PARITY WITH THE REAL PACKAGE IS UNPINNED (scikit-image 0.19.2 is neither in the reference tree nor in this image); what
is pinned is the restatement against SciPy (tests/test_doh_oracle.py) and the device against the restatement (here)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REF_PARAMS = (0.01, 10, 3, 0.0005)        # getFeatures.DEFAULT_FEATURE_PARAMS


@pytest.fixture(scope="module")
def tiny(fe, golden):
    g = golden["tiny_frames"]
    out = []
    for i in range(2):
        frame, cart = fe.polar_to_cart(raw=g[f"raw_{i}"])
        out.append((frame, np.array(cart)))
    return out


@pytest.fixture(scope="module")
def small():
    from radarslampy_b200 import _ffi
    cfg = _ffi.default_config()
    cfg.range_bins = 301
    cfg.raw_width = cfg.meta_bytes + cfg.range_bins
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 2, 1, 64
    eng = _ffi.RadarFE(cfg, device=0)
    assert eng.n == 300
    yield eng
    eng.close()


def _smooth(seed, n, sigma=3.0):
    import cv2
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.random((n, n)).astype(np.float32), (0, 0), sigma)
    return np.ascontiguousarray((img - img.min()) / (img.max() - img.min()), np.float32)


def test_integral_image_and_planes_bit_exact_on_a_real_scan(fe, tiny):
    from oracle import doh_restate as D
    frame, cart = tiny[0]
    ii = fe.doh_response(frame, 0.01, 10, 3, -1)
    want_ii = D.integral_image(cart.astype(np.float64))
    assert np.array_equal(ii, want_ii)
    p0 = fe.doh_response(frame, 0.01, 10, 3, 0)
    assert np.isnan(p0).all()                              # box size int(3 * 0.01) == 0: w_i = inf, dxy = -0.0 * inf
    for k, sigma in ((1, 5.005), (2, 10.0)):
        got = fe.doh_response(frame, 0.01, 10, 3, k)
        want = D.hessian_det(want_ii, sigma)
        assert np.array_equal(got, want), (k, np.abs(got - want).max())


@pytest.mark.parametrize("i", [0, 1])
def test_blobs_equal_the_restatement_on_real_scans(fe, tiny, i):
    from oracle import doh_restate as D
    frame, cart = tiny[i]
    got = fe.detect_doh(frame, *REF_PARAMS)
    want = D.blob_doh(cart.astype(np.double), *REF_PARAMS)
    assert len(want) > 100 and got.shape == want.shape and np.array_equal(got, want)
    assert np.all(got[:, 2] == 10.0)                       # the NaN plane: every blob comes from the last scale
    # rf_detect(mode 1) is the same call with the handle's configured parameters
    cand, n = fe.detect(frame, -1.0, cap=8192, mode=1)
    assert n == len(want) and np.array_equal(cand, want)


@pytest.mark.parametrize("params", [(1, 6, 4, 1e-6), (2, 12, 6, 1e-6), (0.5, 3, 8, 1e-7), (4, 4, 1, 1e-6), (0.01, 10, 3, 1e-7), (1, 30, 10, 1e-6)])
def test_other_scale_lists_on_random_images(small, params):
    from oracle import doh_restate as D
    for seed in (1, 2):
        img = _smooth(seed, small.n, sigma=2.0 + seed)
        frame = small.frame_from_cart(img)
        ii = D.integral_image(img.astype(np.float64))
        for k, s in enumerate(D.sigma_list(*params[:3])):
            got = small.doh_response(frame, *params[:3], k)
            assert np.array_equal(got, D.hessian_det(ii, s), equal_nan=True), (params, k)
        got = small.detect_doh(frame, *params)
        want = D.blob_doh(img.astype(np.double), *params)
        assert got.shape == want.shape and np.array_equal(got, want), (params, seed, len(got), len(want))
        frame.close()


def test_blobs_are_recovered_at_their_scale(small):
    truth = [(60, 60, 4.0), (60, 220, 6.0), (210, 70, 8.0), (215, 218, 5.0)]
    yy, xx = np.mgrid[:small.n, :small.n]
    img = np.zeros((small.n, small.n))
    for r, c, s in truth:
        img += np.exp(-((yy - r) ** 2 + (xx - c) ** 2) / (2 * s * s))
    frame = small.frame_from_cart(img.astype(np.float32))
    blobs = small.detect_doh(frame, 2, 12, 11, 0.002)
    assert len(blobs) == len(truth)
    sig = {}
    for r, c, s in truth:
        d = np.hypot(blobs[:, 0] - r, blobs[:, 1] - c)
        assert d.min() <= 2.0
        sig[s] = blobs[d.argmin(), 2]
    assert sig[4.0] < sig[5.0] < sig[6.0] <= sig[8.0]


def test_dropin_getfeatures_is_the_reference_chain(fe, tiny):
    """getFeatures.appendNewFeatures with the reference's defaults = blob_doh -> argsort(sigma) -> ssc -> fliplr ->
    vstack + order-preserving unique (getFeatures.py:66-118)."""
    from radarslampy_b200 import getFeatures
    from oracle import doh_restate as D
    frame, cart = tiny[0]
    old = np.array([[100.5, 200.25], [300.0, 400.0]], np.float32)
    got, thr = getFeatures.appendNewFeatures(cart, old)
    want = D.get_features(cart, old)
    assert thr == 80 and got.dtype == np.float32 and np.array_equal(got, want)
    assert 180 + 2 <= len(got) <= 220 + 2
    blobs = getFeatures.getBlobsFromCart(cart, **getFeatures.DEFAULT_FEATURE_PARAMS)
    assert np.array_equal(blobs, D.blob_doh(cart.astype(np.double), *REF_PARAMS))
    with pytest.raises(NotImplementedError):
        getFeatures.getBlobsFromCart(cart, method="log")


def test_device_chain_with_the_reference_detector(fe):
    """rf_seq with detector_mode = 1: every (re-)detection is blob_doh -> adaptiveNMS -> append on the device, checked
    against the CPU loop (oracle/ref_system.py) driving the restated detector."""
    from radarslampy_b200 import odometry, synthetic as S
    from oracle import doh_restate as D, ref_system
    raw, _ = S.make_sequence(4, res_m=0.0432, world=S.World(seed=4321), first=0, distort=True)
    got = odometry.run_odometry_device([raw], with_mds=True, detector_mode=1, fe=fe)
    want = ref_system.run_odometry(raw, lambda cart, old: D.get_features(cart, old))
    assert np.all(got["status"] == 0)
    assert got["n_features_in"][0].tolist() == want["n_features_in"].tolist()
    assert got["n_tracked"][0].tolist() == want["n_tracked"].tolist()
    assert got["retrack"][0].tolist() == want["retrack"].tolist()
    d = got["poses"][0] - want["poses"]
    assert np.abs(d[:, :2]).max() <= 3e-4 and np.abs(d[:, 2]).max() <= 3e-5
