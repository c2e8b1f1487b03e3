"""The reference's system loop (RawROAMSystem.run, RawROAMSystem.py:104-300) over in-memory scans, built only
from the drop-in modules: scan decode -> Cartesian -> detect/append features -> Tracker.track (FMT prior, KLT,
outlier rejection) -> rigid transform -> motion-distortion solve -> trajectory / keyframes.  Plotting, file
access and ground-truth lookup are the caller's business; everything numeric follows the reference line by line
(cited below), so a sequence can be run on a machine that has no checkout of the reference."""
import numpy as np

from . import Mapping
from .getFeatures import N_FEATURES_BEFORE_RETRACK, appendNewFeatures
from .Mapping import Keyframe, Map
from .motionDistortion import MotionDistortionSolver
from .parseData import RANGE_RESOLUTION_CART_M, convertRawScanToCartesian, extractDataFromRadarImage
from .Tracker import Tracker
from .trajectoryPlotting import Trajectory, convertPoseToTransform


def run_odometry(raw_scans, timestamps=None, init_pose=(0.0, 0.0, 0.0), param_flags=None, use_fmt_prior=True, append_features=None):
    """raw_scans: iterable of u8 [A, 11 + bins] scans.  Returns a dict with the estimated Trajectory (`traj`), the
    per-frame relative transforms (`R` [P,2,2], `h` [P,2,1] metres), MDS solutions (`mds_x` [P,6]), per-frame feature
    counts and the Map of keyframes.  append_features(cart, old_xy) -> (f32 [K, 2], threshold) replaces
    getFeatures.appendNewFeatures (the reference's determinant-of-Hessian detector) when given."""
    append = append_features or appendNewFeatures
    flags = {"rejectOutliers": True, "useFMT": False}
    flags.update(param_flags or {})
    scans = list(raw_scans)
    n = len(scans)
    timestamps = np.arange(n) * 250000 if timestamps is None else np.asarray(timestamps)
    tracker = Tracker("memory", [None] * n, {}, flags)
    init_pose = np.asarray(init_pose, np.float64)
    est = Trajectory([timestamps[0]], [init_pose])
    mp = Map("memory", est, [None] * n, {})
    tracker.initTraj(est, None)
    cov_p = np.diag([4, 4])                                          # RawROAMSystem.py:133-139
    cov_v = np.diag([1, 1, (5 * np.pi / 180) ** 2])
    mds = MotionDistortionSolver(cov_p, cov_v)
    prev_pose = convertPoseToTransform(init_pose)
    prev_polar = extractDataFromRadarImage(scans[0])[0]              # RawROAMSystem.py:145-146
    prev_cart = convertRawScanToCartesian(scans[0])                  # = convertPolarImageToCartesian(prev_polar), bit for bit, without re-uploading the f32 polar image
    blob, _ = append(prev_cart, np.empty((0, 2)))         # RawROAMSystem.py:149-150
    center = Mapping.cartCenter(prev_polar)
    metric = (blob - center) * RANGE_RESOLUTION_CART_M               # RawROAMSystem.py:153
    zero_v = np.zeros((3,))
    old_kf = Keyframe(init_pose, metric, prev_polar, zero_v)
    mp.addKeyframe(old_kf)
    possible_kf = Keyframe(init_pose, metric, prev_polar, zero_v)
    out = {"R": [], "h": [], "mds_x": [], "n_tracked": [], "n_features_in": [], "fmt_angle": [], "retrack": []}
    for k in range(1, n):
        curr_polar = extractDataFromRadarImage(scans[k])[0]
        curr_cart = convertRawScanToCartesian(scans[k])
        out["n_features_in"].append(len(blob))
        if use_fmt_prior:
            # NB the reference never advances prevImgPolar (RawROAMSystem.py:296-298): the prior is always taken
            # against the first scan; it is returned but not used (Tracker.py:66-72)
            good_old, good_new, ang, corr = tracker.track(prev_cart, curr_cart, prev_polar, curr_polar, blob, k)
        else:
            from .getTransformKLT import getTrackedPointsKLT
            from .outlierRejection import rejectOutliers
            g_new, g_old, b_new, b_old, corr = getTrackedPointsKLT(prev_cart, curr_cart, blob)
            good_old, good_new, mask = rejectOutliers(g_old, g_new)
            corr[np.arange(len(corr))[corr.flatten().astype(bool)]] &= mask[:, np.newaxis]
            ang = 0.0
        old_kf.pruneFeaturePoints(corr)                              # RawROAMSystem.py:185
        R, h = tracker.getTransform(good_old, good_new, pixel=False)
        p_w = old_kf.getPrunedFeaturesGlobalPosition()               # RawROAMSystem.py:194
        centered_new = (good_new - center) * RANGE_RESOLUTION_CART_M
        T_wj = prev_pose @ np.block([[R, h], [np.zeros((2,)), 1]])   # RawROAMSystem.py:201
        mds.update_problem(prev_pose, p_w, centered_new, T_wj, False)
        sol = mds.optimize_library()
        pose_vector = sol[3:]                                        # RawROAMSystem.py:211-214
        rel = mds.T_wj0_inv @ convertPoseToTransform(pose_vector)
        velocity = sol[:3]
        est.appendAbsoluteTransform(timestamps[k], pose_vector)      # RawROAMSystem.py:237
        latest = pose_vector
        possible_kf.updateInfo(latest, centered_new, curr_polar, velocity)
        n_feat = good_new.shape[0]
        retrack = n_feat <= N_FEATURES_BEFORE_RETRACK                # RawROAMSystem.py:250-251
        if retrack or mp.isGoodKeyframe(possible_kf):
            mp.addKeyframe(possible_kf)
            old_kf = possible_kf
            if retrack:
                good_new, _ = append(curr_cart, good_new)
                centered_new = (good_new - center) * RANGE_RESOLUTION_CART_M
                old_kf.updateInfo(latest, centered_new, curr_polar, velocity)
            possible_kf = Keyframe(latest, centered_new, curr_polar, velocity)
        out["R"].append(rel[:2, :2]); out["h"].append(rel[:2, 2:]); out["mds_x"].append(sol)
        out["n_tracked"].append(n_feat); out["fmt_angle"].append(ang); out["retrack"].append(bool(retrack))
        blob = good_new.copy()                                       # RawROAMSystem.py:296-298
        prev_cart = curr_cart
        prev_pose = convertPoseToTransform(latest)
    res = {k: np.array(v) for k, v in out.items()}
    res["traj"], res["map"] = est, mp
    return res


def run_odometry_device(sequences, init_pose=None, with_mds=True, graph=True, detector_mode=1, device=None, fe=None):
    """The same loop for several independent sequences at once, entirely on the device (rf_seq, csrc/k_seq.cu):
    features, keyframe state and poses never leave HBM between frames, re-detection (response, NMS, SSC bisection,
    append) runs on the device for exactly the sequences that need it, and a step is one CUDA-graph launch.

    detector_mode 1 (default) is the reference's detector, determinant-of-Hessian blobs (getFeatures.py:13-18; csrc/k_doh.cu);
    0 is the structure-tensor minimum-eigenvalue response (BASELINE.json north_star; csrc/k_detect.cu).
    sequences: [S][T] raw scans (u8 [A, 11 + bins] each; every sequence has T frames).  with_mds=False gives the pose
    chain without motion compensation (T_wj = prev_pose @ [R, h], RawROAMSystem.py:201; BASELINE configs[1]).
    Returns per-sequence arrays: poses [S, T, 3], R [S, T-1, 2, 2], h [S, T-1, 2, 1], mds_x [S, T-1, 6], n_tracked,
    n_features_in, retrack [S, T-1], n_keyframes [S], and the raw per-step records (`steps`)."""
    from . import _engine, _ffi
    seqs = [list(s) for s in sequences]
    S, T = len(seqs), len(seqs[0])
    if any(len(s) != T for s in seqs):
        raise ValueError("every sequence needs the same number of frames")
    if fe is None:
        if device is not None:
            _engine.set_device(device)
        fe = _engine.engine(raw_width=seqs[0][0].shape[1])
    runner = fe.new_sequences(S, S * T, detector_mode=detector_mode)
    try:
        # arena layout: frame t of sequence s at index t * S + s  (one slot of S fresh scans per step)
        for t in range(T):
            runner.upload(t * S, np.stack([seqs[s][t] for s in range(S)]))
        runner.reset(0, 1, init_pose)
        for t in range(1, T):
            runner.step(t * S, 1, with_mds=with_mds, graph=graph)
        steps = [runner.results(t) for t in range(0, T)]
        runner.sync()
    finally:
        runner.close()
    rec = np.stack(steps, axis=1)                                   # [S, T]
    return {"poses": rec["pose"], "R": rec["R"][:, 1:].reshape(S, T - 1, 2, 2), "h": rec["h"][:, 1:].reshape(S, T - 1, 2, 1),
            "mds_x": rec["mds_x"][:, 1:], "kab_R": rec["kab_R"][:, 1:].reshape(S, T - 1, 2, 2), "kab_h": rec["kab_h"][:, 1:],
            "n_tracked": rec["n_tracked"][:, 1:], "n_features_in": rec["n_features_in"][:, 1:],
            "n_features_out": rec["n_features_out"], "retrack": rec["retrack"][:, 1:].astype(bool),
            "n_keyframes": rec["n_keyframes"][:, -1], "status": rec["status"], "steps": rec}
