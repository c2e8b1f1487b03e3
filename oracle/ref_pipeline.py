"""TEST INFRASTRUCTURE ONLY — the reference's per-pair hot path restated as the sequence of
THIRD-PARTY calls the reference itself makes (OpenCV, SciPy, networkx, NumPy), with the same
arguments, dtypes and order.  This is the CPU baseline `bench.py` times (cpu_baseline /
--impl reference, kind "port") and a second oracle for the GPU parity tests: where
oracle/restate.py re-derives the arithmetic in plain C, this module runs the very library
code the reference bottoms out in (/root/reference itself cannot travel to the GPU box; the
libraries are part of the image).

Pinned by tests/test_ref_pipeline.py against the goldens generated from the unmodified
reference (tests/golden/tiny_stages.npz).

Each function cites the reference file:line it follows.  The product never imports this."""
import numpy as np


def extract_polar(raw_u8, range_res_m=0.0432, max_range_m=87.5):
    """parseData.py:17-53 extractDataFromRadarImage -> range_azimuth_data f32 [A, W]."""
    power = raw_u8[:, 11:].astype(np.float32) / 255.
    return power[:, :int(max_range_m / range_res_m)]


def polar_to_cart(polar_f32, downsample=2):
    """parseData.py:100-135 convertPolarImageToCartesian (default, linear mode)."""
    import cv2
    w, h = polar_f32.shape
    max_radius = h // downsample if downsample > 1 else h
    size = (max_radius * 2, max_radius * 2)
    center = tuple(np.array(size) / 2)
    flags = cv2.WARP_POLAR_LINEAR + cv2.WARP_INVERSE_MAP + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS
    return cv2.warpPolar(polar_f32, size, center, max_radius, flags)


def tracked_points_klt(src_cart, tgt_cart, pts_xy, err_threshold=10):
    """getTransformKLT.py:317-381 getTrackedPointsKLT without the re-detection branch
    (:347-352 — detector parity is unpinned, callers hand in >= 80 features)."""
    import cv2
    pts = np.ascontiguousarray(np.asarray(pts_xy)[:, :2]).astype(np.float32)
    src_u8 = (src_cart * 255).astype(np.uint8)
    tgt_u8 = (tgt_cart * 255).astype(np.uint8)
    nxt, status, err = cv2.calcOpticalFlowPyrLK(src_u8, tgt_u8, pts, None, winSize=(15, 15), maxLevel=3,
                                                criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    status &= (err < err_threshold)
    good = (status == 1).flatten()
    return nxt[good, :], pts[good, :], nxt[~good, :], pts[~good, :], status


def reject_outliers(prev_xy, new_xy, thr_px=0.5 / (0.0432 * 2)):
    """outlierRejection.py:16-95: cdist -> thresholded graph -> first largest clique
    nx.find_cliques yields."""
    import networkx as nx
    from scipy.spatial.distance import cdist
    K = prev_xy.shape[0]
    mask = np.zeros(K, dtype=bool)
    if K == 0:
        return prev_xy, new_xy, mask
    diff = np.abs(cdist(prev_xy, prev_xy, metric="euclidean") - cdist(new_xy, new_xy, metric="euclidean"))
    G = nx.Graph((diff <= thr_px).astype(np.int8))
    best, best_n = [], 0
    for clique in nx.find_cliques(G):
        if len(clique) > best_n:
            best_n, best = len(clique), clique
    mask[np.array(best)] = True
    return prev_xy[mask], new_xy[mask], mask


def transform_svd(src_xy, tgt_xy):
    """getTransformKLT.py:129-162 calculateTransformSVD: src = R tgt + h (dtype follows the
    inputs: float32 points give a float32 SVD, as in the reference)."""
    x1m = np.mean(tgt_xy, axis=0, keepdims=True)
    x0m = np.mean(src_xy, axis=0, keepdims=True)
    Cm = (src_xy - x0m).T @ (tgt_xy - x1m)
    U, _, Vt = np.linalg.svd(Cm)
    fix = np.eye(2)
    fix[-1, -1] = np.linalg.det(U @ Vt)
    R = U @ fix @ Vt
    h = x0m - (R @ x1m.T).T
    return R, h.T


def pose_matrix(x, y, th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s, x], [s, c, y], [0, 0, 1.0]])


def mds_solve(T_wj0, p_w, p_jt, T_wj, period=0.25):
    """motionDistortion.py:80-99, 295-325: least_squares(error_vector, x0, jac='2-point', method='lm')."""
    from . import restate
    return restate.mds_solve(T_wj0, p_w, p_jt, T_wj, period=period)


def track_pair(raw_prev, raw_next, feats_xy, prev_pose=None, with_mds=False, range_res_m=0.0432, carts=None):
    """One frame pair, polar -> pose: RawROAMSystem.py:164-212 with the previous frame as the
    keyframe (p_w = prev_pose o centred good_old).  Returns a dict mirroring rf_pair_result.
    `carts` lets a sequence caller reuse the previous frame's Cartesian image exactly like the
    reference loop does (RawROAMSystem.py:297)."""
    res_cart = range_res_m * 2
    if carts is None:
        c0 = polar_to_cart(extract_polar(raw_prev, range_res_m))
        c1 = polar_to_cart(extract_polar(raw_next, range_res_m))
    else:
        c0, c1 = carts
    good_new, good_old, _, _, status = tracked_points_klt(c0, c1, feats_xy)
    p_old, p_new, mask = reject_outliers(good_old, good_new, 0.5 / res_cart)
    corr = status.copy()
    rows = np.arange(corr.shape[0])
    corr[rows[corr.flatten().astype(bool)]] &= mask[:, np.newaxis]      # Tracker.py:102-104
    if p_old.shape[0]:
        R, h = transform_svd(p_old, p_new)
    else:
        R, h = np.eye(2), np.zeros((2, 1))
    h = h * res_cart                                                   # Tracker.py:125-126
    out = dict(R=np.asarray(R, np.float64), h=np.asarray(h, np.float64).ravel(), n_good=int(good_old.shape[0]),
               n_inliers=int(mask.sum()), corr_status=corr.ravel(), mask=mask, good_old=good_old, good_new=good_new)
    if with_mds:
        pose = np.zeros(3) if prev_pose is None else np.asarray(prev_pose, np.float64)
        T0 = pose_matrix(*pose)
        center = c0.shape[0] / 2
        old_m = (p_old.astype(np.float64) - center) * res_cart
        p_w = (T0[:2, :2] @ old_m.T).T + T0[:2, 2]
        p_jt = (p_new.astype(np.float64) - center) * res_cart       # RawROAMSystem.py:198-199
        T_wj = T0 @ np.block([[out["R"], out["h"].reshape(2, 1)], [np.zeros((2,)), 1]])   # :201
        out["mds_x"] = mds_solve(T0, p_w, p_jt, T_wj)
    return out


def rotation_fmt(src_polar, tgt_polar, downsample=10, max_range_clip_m=87.5, range_res_cart_m=0.0432 * 2):
    """FMT.py:36-90 getRotationUsingFMT as the cv2 calls the reference makes (resize, warpPolar x 2,
    createHanningWindow, phaseCorrelate) -> (angle rad, scale, response)."""
    import cv2
    if max_range_clip_m > 0:
        clip = int(max_range_clip_m / range_res_cart_m)
        src_polar, tgt_polar = src_polar[:, :clip], tgt_polar[:, :clip]
    H, W = src_polar.shape

    def log_polar(p):
        small = cv2.resize(p, (int(W // downsample), H))
        cart = polar_to_cart(small, downsample=1)                               # parseData.py:148-151
        h, w = cart.shape
        return cv2.warpPolar(cart, None, (h / 2, w / 2), w / 2, cv2.WARP_POLAR_LOG + cv2.INTER_LINEAR + cv2.WARP_FILL_OUTLIERS)

    a, b = log_polar(src_polar), log_polar(tgt_polar)
    h_lp, w_lp = a.shape
    (scale, angle), response = cv2.phaseCorrelate(a, b, cv2.createHanningWindow((w_lp, h_lp), cv2.CV_32F))
    sz = max(h_lp, w_lp)
    angle = (-float(angle) * 2 * np.pi / sz + np.pi) % (2 * np.pi) - np.pi        # utils.normalize_angles
    return angle, np.exp(np.log(h_lp / 2) / sz) ** scale, response
