"""TEST INFRASTRUCTURE ONLY — the reference's system loop (RawROAMSystem.run, RawROAMSystem.py:104-300) over
in-memory scans as the sequence of THIRD-PARTY calls the reference makes (oracle/ref_pipeline.py: cv2, scipy,
networkx, numpy), with the keyframe bookkeeping of Mapping.py restated in NumPy.  The detector is a callback
(`detect(cart_f32, old_xy) -> f32 [K, 2]`, the role of getFeatures.appendNewFeatures): detector parity is unpinned
(scikit-image absent), so tests hand the same detector to both sides.  The product never imports this."""
import numpy as np

from . import ref_pipeline as P

N_FEATURES_BEFORE_RETRACK = 60
ROT_THRESHOLD, TRANS_THRESHOLD_SQ = 0.2, 4.0


def _pose_T(p):
    c, s = np.cos(p[2]), np.sin(p[2])
    return np.array([[c, -s, p[0]], [s, c, p[1]], [0, 0, 1.0]])


def _undistort(v, pts, period=0.25):
    """motionDistortion.py:107-153."""
    dT = period * np.arctan2(-pts[:, 1], -pts[:, 0]) / (2 * np.pi)
    th = v[2] * dT
    x = np.cos(th) * pts[:, 0] - np.sin(th) * pts[:, 1] + v[0] * dT
    y = np.sin(th) * pts[:, 0] + np.cos(th) * pts[:, 1] + v[1] * dT
    return np.stack([x, y], 1)


class _KF:
    def __init__(self, pose, local, v):
        self.update(pose, local, v)

    def update(self, pose, local, v):
        self.pose = np.asarray(pose, np.float64)
        self.pruned_und = _undistort(v, np.asarray(local, np.float64))

    def prune(self, corr):
        self.pruned_und = self.pruned_und[corr.flatten().astype(bool)]

    def global_pts(self):
        T = _pose_T(self.pose)
        return (T[:2, :2] @ self.pruned_und.T + T[:2, 2:]).T


def run_odometry(raw_scans, detect, init_pose=(0.0, 0.0, 0.0), with_mds=True, range_res_m=0.0432):
    """with_mds=False: the pose chain without motion compensation, T_wj = prev_pose @ [R, h] (RawROAMSystem.py:201,
    BASELINE configs[1]); the keyframe points are then undistorted with zero velocity."""
    scans = list(raw_scans)
    RES = 2 * range_res_m                                            # RANGE_RESOLUTION_CART_M (parseData.py:13-15)
    thr_px = 0.5 / RES                                               # outlierRejection.py:29-31
    init_pose = np.asarray(init_pose, np.float64)
    prev_pose = _pose_T(init_pose)
    prev_cart = P.polar_to_cart(P.extract_polar(scans[0], range_res_m))
    blob = detect(prev_cart, np.empty((0, 2), np.float32))
    center = np.array(prev_cart.shape) / 2
    metric = (blob - center) * RES
    old_kf = _KF(init_pose, metric, np.zeros(3))
    kfs = [old_kf]
    possible = _KF(init_pose, metric, np.zeros(3))
    out = {"R": [], "h": [], "mds_x": [], "n_tracked": [], "n_features_in": [], "poses": [init_pose], "retrack": []}
    for k in range(1, len(scans)):
        curr_cart = P.polar_to_cart(P.extract_polar(scans[k], range_res_m))
        out["n_features_in"].append(len(blob))
        g_new, g_old, b_new, b_old, status = P.tracked_points_klt(prev_cart, curr_cart, blob)
        good_old, good_new, mask = P.reject_outliers(g_old, g_new, thr_px)
        corr = status.copy()
        corr[np.arange(len(corr))[corr.flatten().astype(bool)]] &= mask[:, np.newaxis].astype(corr.dtype)
        old_kf.prune(corr)
        R, h = P.transform_svd(good_old, good_new)
        h = h * RES
        p_w = old_kf.global_pts()
        centered_new = (good_new - center) * RES
        T_wj = prev_pose @ np.block([[R, h], [np.zeros((2,)), 1]])
        if with_mds:
            sol = P.mds_solve(prev_pose, p_w, centered_new, T_wj)
        else:
            sol = np.array([0.0, 0.0, 0.0, T_wj[0, 2], T_wj[1, 2], np.arctan2(T_wj[1, 0], T_wj[0, 0])])
        pose_vector = sol[3:]
        rel = np.linalg.inv(prev_pose) @ _pose_T(pose_vector)
        velocity = sol[:3]
        out["poses"].append(pose_vector)
        possible.update(pose_vector, centered_new, velocity)
        n_feat = good_new.shape[0]
        retrack = n_feat <= N_FEATURES_BEFORE_RETRACK
        d = kfs[-1].pose - possible.pose
        if retrack or abs(d[2]) >= ROT_THRESHOLD or (d[:2] ** 2).sum() >= TRANS_THRESHOLD_SQ:
            kfs.append(possible)
            old_kf = possible
            if retrack:
                good_new = detect(curr_cart, good_new)
                centered_new = (good_new - center) * RES
                old_kf.update(pose_vector, centered_new, velocity)
            possible = _KF(pose_vector, centered_new, velocity)
        out["R"].append(rel[:2, :2]); out["h"].append(rel[:2, 2:]); out["mds_x"].append(sol)
        out["n_tracked"].append(n_feat); out["retrack"].append(bool(retrack))
        blob = good_new.copy()
        prev_cart = curr_cart
        prev_pose = _pose_T(pose_vector)
    res = {k: np.array(v) for k, v in out.items()}
    res["n_keyframes"] = len(kfs)
    res["features"] = blob
    return res
