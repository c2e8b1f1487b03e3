"""GPU parity of the drop-in modules (radarslampy_b200/{parseData,getTransformKLT,outlierRejection,
motionDistortion,Tracker,getFeatures,ANMS,getPointCloud}.py): called exactly as RawROAMSystem/Tracker call
the reference's modules, checked against goldens generated from the unmodified reference.
Tolerances (north_star): images bit-exact, status identical, tracks <= 0.02 px, pose <= 1e-4 m / 1e-5 rad."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_PX, TOL_M, TOL_RAD = 0.02, 1e-4, 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def carts(golden):
    from radarslampy_b200 import parseData
    fr = golden["tiny_frames"]
    out = []
    for i in range(3):
        polar, az, rres, ares, valid, ts = parseData.extractDataFromRadarImage(fr[f"raw_{i}"])
        assert polar.shape == (400, 2025) and polar.dtype == np.float32 and sha(polar) == str(fr[f"polar_sha256_{i}"])
        assert az.shape == (400, 1) and az.dtype == np.float32 and valid.dtype == bool and ts.dtype == np.int64
        assert rres == 0.0432
        cart = parseData.convertPolarImageToCartesian(polar)
        assert cart.shape == (2024, 2024) and cart.dtype == np.float32
        assert sha(cart) == str(fr[f"cart_sha256_{i}"])
        assert sha(parseData.convertRawScanToCartesian(fr[f"raw_{i}"])) == str(fr[f"cart_sha256_{i}"])
        out.append((polar, cart))
    return out


def test_metadata_matches_reference_decode(golden):
    from oracle import restate as R
    from radarslampy_b200 import parseData
    raw = golden["tiny_frames"]["raw_1"]
    _, az, _, ares, valid, ts = parseData.extractDataFromRadarImage(raw)
    ts0, az0, v0 = R.extract_metadata(raw)
    assert np.array_equal(ts, ts0) and np.array_equal(az, az0) and np.array_equal(valid, v0)
    assert ares == az0[1] - az0[0]


def test_klt_module(carts, golden):
    from radarslampy_b200 import getTransformKLT as G
    st = golden["tiny_stages"]
    for p in range(2):
        good_new, good_old, bad_new, bad_old, status = G.getTrackedPointsKLT(carts[p][1], carts[p + 1][1], st[f"feat_in_{p}"])
        assert status.shape == st[f"klt_status_{p}"].shape and status.dtype == np.uint8
        assert np.array_equal(status, st[f"klt_status_{p}"])
        assert np.array_equal(good_old, st[f"klt_good_old_{p}"]) and np.array_equal(bad_old, st[f"klt_bad_old_{p}"])
        assert np.abs(good_new - st[f"klt_good_new_{p}"]).max() <= TOL_PX
        assert good_new.dtype == np.float32 and bad_new.shape == st[f"klt_bad_new_{p}"].shape


def test_klt_module_accepts_plain_numpy_images(carts, golden):
    """a caller-made copy (no device frame attached) takes the upload path and gives the same answer"""
    from radarslampy_b200 import getTransformKLT as G
    st = golden["tiny_stages"]
    a, b = np.array(carts[0][1]), np.array(carts[1][1])
    assert type(a) is np.ndarray and a.flags.writeable
    good_new, _, _, _, status = G.getTrackedPointsKLT(a, b, st["feat_in_0"])
    assert np.array_equal(status, st["klt_status_0"])
    assert np.abs(good_new - st["klt_good_new_0"]).max() <= TOL_PX


def test_reject_and_transform_modules(golden):
    from radarslampy_b200 import getTransformKLT as G
    from radarslampy_b200 import outlierRejection as O
    st = golden["tiny_stages"]
    assert O.DIST_THRESHOLD_PX == 0.5 / 0.0864
    for p in range(int(st["n_pairs"])):
        old, new = st[f"klt_good_old_{p}"], st[f"klt_good_new_{p}"]
        po, pn, mask = O.rejectOutliers(old, new)
        assert mask.dtype == bool and np.array_equal(mask, st[f"rej_mask_{p}"])
        assert np.array_equal(po, st[f"svd_src_{p}"]) and np.array_equal(pn, st[f"svd_tgt_{p}"])
        Rm, h = G.calculateTransformSVD(po, pn)
        assert Rm.shape == (2, 2) and h.shape == (2, 1) and Rm.dtype == np.float64
        assert np.abs(h - st[f"svd_h_{p}"]).max() * 0.0864 <= TOL_M
        assert abs(np.arctan2(Rm[1, 0], Rm[0, 0]) - np.arctan2(st[f"svd_R_{p}"][1, 0], st[f"svd_R_{p}"][0, 0])) <= TOL_RAD
    c = golden["clique_fixture"]
    assert np.array_equal(O.rejectOutliers(c["prev"], c["new"])[2], c["mask"])


def test_tracker_module(carts, golden):
    from radarslampy_b200.Tracker import Tracker
    st = golden["tiny_stages"]
    t = Tracker("tiny", ["a", "b", "c"], {}, {"rejectOutliers": True, "useFMT": False})
    for p in range(2):
        good_old, good_new, ang, corr = t.track(carts[p][1], carts[p + 1][1], carts[p][0], carts[p + 1][0], st[f"feat_in_{p}"], p + 1)
        want = st[f"klt_status_{p}"].copy()
        want[want.flatten().astype(bool)] &= st[f"rej_mask_{p}"][:, None].astype(np.uint8)
        assert np.array_equal(corr, want)
        assert np.array_equal(good_old, st[f"svd_src_{p}"])
        assert np.abs(good_new - st[f"svd_tgt_{p}"]).max() <= TOL_PX
        Rm, h = t.getTransform(good_old, good_new, pixel=False)
        assert np.abs(h - st[f"svd_h_{p}"] * 0.0864).max() <= TOL_M
        assert abs(np.arctan2(Rm[1, 0], Rm[0, 0]) - np.arctan2(st[f"svd_R_{p}"][1, 0], st[f"svd_R_{p}"][0, 0])) <= TOL_RAD


def test_motion_distortion_module(golden):
    from radarslampy_b200.motionDistortion import MotionDistortionSolver
    st = golden["tiny_stages"]
    cov_p = np.diag([4, 4])
    cov_v = np.diag([1, 1, (5 * np.pi / 180) ** 2])                 # RawROAMSystem.py:135-139
    mds = MotionDistortionSolver(cov_p, cov_v)
    for p in range(int(st["n_pairs"])):
        mds.update_problem(st[f"mds_Twj0_{p}"], st[f"mds_pw_{p}"], st[f"mds_pjt_{p}"], st[f"mds_Twj_{p}"])
        x = mds.optimize_library()
        ref = st[f"mds_x_{p}"]
        assert x.shape == (6,)
        assert np.abs(x[3:5] - ref[3:5]).max() <= TOL_M and abs(x[5] - ref[5]) <= TOL_RAD
        assert np.allclose(mds.T_wj0_inv, np.linalg.inv(st[f"mds_Twj0_{p}"]))
    # undistort: closed form of motionDistortion.py:127-153
    v = np.array([3.0, -1.0, 0.2])
    pts = np.random.default_rng(0).uniform(-80, 80, (50, 2))
    u = MotionDistortionSolver.undistort(v, pts)
    dT = MotionDistortionSolver.compute_time_deltas(0.25, pts)
    th = v[2] * dT
    wx = np.cos(th) * pts[:, 0] - np.sin(th) * pts[:, 1] + v[0] * dT
    wy = np.sin(th) * pts[:, 0] + np.cos(th) * pts[:, 1] + v[1] * dT
    assert u.shape == (50, 3) and np.allclose(u[:, 0], wx, atol=1e-12) and np.allclose(u[:, 1], wy, atol=1e-12)
    assert np.all(u[:, 2] == 1)


def test_features_modules(carts, golden):
    from radarslampy_b200 import ANMS, getFeatures, getPointCloud
    g = golden["ssc_fixture"]
    k, W, H = (int(v) for v in g["cfg_0"])
    assert np.array_equal(ANMS.ssc(g["kp_0"], k, 0.1, W, H), g["sel_0"])
    peaks = getPointCloud.getPointCloudPolarInd(carts[0][0])
    assert np.array_equal(peaks, golden["peaks_fixture"]["peaks"].astype(int))
    # detector + ANMS + append: 200 +- 10 % well-spread features in (x, y) order, no duplicates, f32
    cart = carts[0][1]
    coord, radii = getFeatures.getFeatures(cart)
    assert 180 <= coord.shape[0] <= 220 and coord.shape[1] == 2 and radii.shape == (coord.shape[0],)
    assert 0 <= coord.min() and coord.max() < 2024
    old = golden["tiny_stages"]["feat_in_4"]
    feats, thr = getFeatures.appendNewFeatures(cart, old)
    assert thr == 80 and feats.dtype == np.float32
    assert np.array_equal(feats[:len(old)], old)
    assert len(np.unique(feats, axis=0)) == len(feats)
    # and the features are trackable: a good part survives KLT + err gating on the next frame (sigma = 10 blobs of the
    # reference's detector on a real scan: 84 of 191; the structure-tensor corners: more than half)
    from radarslampy_b200 import getTransformKLT as G
    good_new, *_ = G.getTrackedPointsKLT(cart, carts[1][1], coord)
    assert good_new.shape[0] >= 0.3 * coord.shape[0]
    blobs = getFeatures.adaptiveNMS(cart, getFeatures.getBlobsFromCart(cart, method="mineig"))
    good_new, *_ = G.getTrackedPointsKLT(cart, carts[1][1], np.fliplr(blobs[:, :2]))
    assert good_new.shape[0] >= 0.5 * blobs.shape[0]


def test_cartesian_to_polar_forward_maps_match_cv2(carts):
    """parseData.convertCartesianImageToPolar (parseData.py:69-97): cv2.warpPolar forward map, linear and semi-log,
    default and explicit output sizes — the same bits as the cv2 wheel of this image."""
    import cv2
    from radarslampy_b200 import parseData
    cart = np.ascontiguousarray(carts[0][1][::4, ::4])                   # 506 x 506
    n = cart.shape[0]
    for log_mode in (False, True):
        flags = (cv2.WARP_POLAR_LOG if log_mode else cv2.WARP_POLAR_LINEAR) + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS
        for shape_hw in (None, (360, 200)):
            size = None if shape_hw is None else (shape_hw[1], shape_hw[0])
            want = cv2.warpPolar(cart, size, (n / 2, n / 2), n / 2, flags)
            got = parseData.convertCartesianImageToPolar(cart, logPolarMode=log_mode, shapeHW=shape_hw)
            assert got.shape == want.shape and got.dtype == np.float32
            assert np.array_equal(got, want), (log_mode, shape_hw, np.abs(got - want).max())
    with pytest.raises(AssertionError):
        parseData.convertCartesianImageToPolar(cart[:, :-2])


def test_polar_to_cartesian_log_mode_matches_cv2_within_its_logf(carts):
    """convertPolarImageToCartesian(logPolarMode=True) (parseData.py:131-133).  cv2 takes the radius through its f32
    cv::log (a vendor routine: <= 1 ulp from logf on ~2 % of arguments), so a few sample positions land on the
    neighbouring 1/32-pixel step: all but a handful of pixels are identical, the rest differ by a fraction of a grey level."""
    import cv2
    from radarslampy_b200 import parseData
    polar = carts[0][0]
    got = parseData.convertPolarImageToCartesian(polar, logPolarMode=True)
    w, h = polar.shape
    R = h // 2
    flags = cv2.WARP_POLAR_LINEAR + cv2.WARP_INVERSE_MAP + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS + cv2.WARP_POLAR_LOG
    want = cv2.warpPolar(polar, (2 * R, 2 * R), (R, R), R, flags)
    assert got.shape == want.shape == (2024, 2024) and got.dtype == np.float32
    differ = got != want
    assert differ.mean() < 5e-3, differ.mean()
    assert np.abs(got - want).max() <= 0.1


def test_undistort_with_explicit_times_and_solver_debug_methods(golden):
    from radarslampy_b200.motionDistortion import MotionDistortionSolver
    rng = np.random.default_rng(4)
    pts = rng.uniform(-60, 60, (40, 2))
    v = np.array([6.0, 0.3, -0.05])
    times = rng.uniform(-0.12, 0.12, 40)
    u = MotionDistortionSolver.undistort(v, pts, times=times)
    th = v[2] * times
    wx = np.cos(th) * pts[:, 0] - np.sin(th) * pts[:, 1] + v[0] * times
    wy = np.sin(th) * pts[:, 0] + np.cos(th) * pts[:, 1] + v[1] * times
    assert u.shape == (40, 3) and np.allclose(u[:, 0], wx, atol=1e-12) and np.allclose(u[:, 1], wy, atol=1e-12)
    # error_vector is the residual the device solve minimises: 0.5 |e|^2 at its solution equals the returned cost
    st = golden["tiny_stages"]
    s = MotionDistortionSolver(np.diag([4, 4]), np.diag([1, 1, (5 * np.pi / 180) ** 2]))
    s.update_problem(st["mds_Twj0_3"], st["mds_pw_3"], st["mds_pjt_3"], st["mds_Twj_3"])
    x = s.optimize_library()
    e = s.error_vector(x)
    assert e.shape == (2 * len(st["mds_pw_3"]) + 3,) and abs(0.5 * e @ e - s.cost) <= 1e-9 * max(1.0, s.cost)
    assert s.jacobian_vector(x).shape == (e.size, 6) and s.optimize() is None
    assert s.expected_observed_pts(st["mds_Twj_3"]).shape == (3, len(st["mds_pw_3"]))


def test_driver_plot_calls_run_under_a_headless_matplotlib(carts, golden, tmp_path):
    """RawROAMSystem.plot (RawROAMSystem.py:335-405) calls tracker.plot -> visualize_transform and map.plot every third
    frame: they exist and run.  matplotlib is absent from this image; the inert shim stands in for it."""
    import os
    import sys
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
    had = "matplotlib" in sys.modules
    sys.path.insert(0, shims)
    try:
        from radarslampy_b200 import FMT, parseData
        from radarslampy_b200.Mapping import Keyframe, Map
        from radarslampy_b200.Tracker import Tracker
        from radarslampy_b200.trajectoryPlotting import Trajectory, plotGtAndEstTrajectory
        st = golden["tiny_stages"]
        tracker = Tracker("tiny", [None] * 3, {"imgSave": str(tmp_path)}, {"rejectOutliers": True})
        est = Trajectory([0, 250000, 500000, 750000, 1000000], np.cumsum(np.full((5, 3), 0.1), axis=0))
        mp = Map("tiny", est, [None] * 3, {})
        mp.addKeyframe(Keyframe(np.zeros(3), (st["feat_in_0"] - 1012) * 0.0864, carts[0][0], np.zeros(3)))
        for seq in (1, 2):
            good_old, good_new, ang, corr = tracker.track(carts[seq - 1][1], carts[seq][1], carts[seq - 1][0], carts[seq][0], st["feat_in_0"], seq)
            tracker.plot(carts[seq - 1][1], carts[seq][1], good_old, good_new, seq, save=True, show=False)
            mp.plot(None, show=False)
        est.plot(savePath=False)
        plotGtAndEstTrajectory(est, est, title="t", info="i", savePath=None)
        FMT.plotCartPolar(carts[0][0], carts[1][0], carts[0][1], carts[1][1])
        img = np.zeros((8, 8, 3), np.uint8)
        assert parseData.drawCVPoint(img, (3, 5))[5, 3].tolist() == [0, 0, 255] and img.sum() == 255
    finally:
        sys.path.remove(shims)
        if not had:
            for k in [k for k in sys.modules if k == "matplotlib" or k.startswith("matplotlib.")]:
                del sys.modules[k]
