"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, imported through oracle/shims) in the authoring container.

    python -m oracle.gen_golden            # writes tests/golden/tiny_stages.npz, tiny_frames.npz, ...

What is recorded (all produced by the reference's own functions, intercepted at the
call boundaries that SURVEY.md §8(b) names; nothing is recomputed by this repo's code):

  tiny_frames.npz   raw uint8 [400,3779] scans 0..2 of data/tiny (the parity input) and
                    SHA-256 of the reference's Cartesian f32 / u8 images for them.
  tiny_stages.npz   for every frame pair of RawROAMSystem("tiny").run():
                      feat_in_i      features handed to Tracker.track          (K,2) f32
                      klt_*_i        getTrackedPointsKLT outputs
                      rej_mask_i     rejectOutliers mask over klt_good rows
                      svd_R_i/h_i    calculateTransformSVD (pixel units)
                      mds_*_i        MotionDistortionSolver inputs and solution
                      pose_i         pose appended to the trajectory
  clique_fixture.npz  outlier_test.npz correspondences + the reference mask
  kabsch_fixture.npz  the 95 real pairs of archive/testTransformKLT2.py + reference R,h
  ssc_fixture.npz     seeded keypoints + ANMS.ssc selections
  peaks_fixture.npz   getPointCloudPolarInd on tiny frame 0

The detector is a stand-in (scikit-image is not installed; detector parity is
unpinned): cv2.goodFeaturesToTrack → (row, col, sigma) fed to the reference's own
adaptiveNMS/ssc.  Features are therefore golden *inputs*, not golden outputs.
"""
import glob
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def standin_blob_doh(image, **kw):
    import cv2
    u8 = (np.asarray(image) * 255).astype(np.uint8)
    pts = cv2.goodFeaturesToTrack(u8, 3000, 0.01, 5)
    pts = pts.reshape(-1, 2)
    sig = np.where(np.arange(len(pts)) % 2 == 0, 5.005, 10.0)
    return np.column_stack([pts[:, 1], pts[:, 0], sig]).astype(np.float64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import as ri
    ri.activate()
    import cv2
    import skimage.feature as skf
    skf.set_blob_doh(standin_blob_doh)
    os.makedirs(GOLD, exist_ok=True)

    # ---------------- raw frames + image checksums -------------------------------
    pd = ri.ref_module("parseData")
    paths = sorted(glob.glob(os.path.join(ri.REFERENCE_ROOT, "data", "tiny", "radar", "*.png")))
    frames = {}
    for i in range(3):
        raw = cv2.imread(paths[i], cv2.IMREAD_GRAYSCALE)
        polar = pd.extractDataFromRadarImage(raw)[0]
        cart = pd.convertPolarImageToCartesian(polar)
        frames[f"raw_{i}"] = raw
        frames[f"cart_sha256_{i}"] = np.array(sha(cart))
        frames[f"u8_sha256_{i}"] = np.array(sha((cart * 255).astype(np.uint8)))
        frames[f"polar_sha256_{i}"] = np.array(sha(polar))
    np.savez_compressed(os.path.join(GOLD, "tiny_frames.npz"), **frames)

    # ---------------- full system run with interception ---------------------------
    rec = {}
    state = {"pair": 0}
    import Tracker as T
    import getTransformKLT as gk
    import motionDistortion as md
    import RawROAMSystem as rs

    orig_klt, orig_rej, orig_svd = T.getTrackedPointsKLT, T.rejectOutliers, T.calculateTransformSVD

    def klt(src, tgt, feat):
        i = state["pair"]
        out = orig_klt(src, tgt, feat)
        rec[f"feat_in_{i}"] = np.asarray(feat).copy()
        for name, v in zip(("good_new", "good_old", "bad_new", "bad_old", "status"), out):
            rec[f"klt_{name}_{i}"] = np.asarray(v).copy()
        return out

    def rej(a, b):
        i = state["pair"]
        out = orig_rej(a, b)
        rec[f"rej_mask_{i}"] = out[2].copy()
        return out

    def svd(a, b):
        i = state["pair"]
        R, h = orig_svd(a, b)
        rec[f"svd_src_{i}"], rec[f"svd_tgt_{i}"] = np.asarray(a).copy(), np.asarray(b).copy()
        rec[f"svd_R_{i}"], rec[f"svd_h_{i}"] = R.copy(), h.copy()
        return R, h

    T.getTrackedPointsKLT, T.rejectOutliers, T.calculateTransformSVD = klt, rej, svd
    orig_up, orig_opt = md.MotionDistortionSolver.update_problem, md.MotionDistortionSolver.optimize_library

    def up(self, T_wj0, p_w, p_jt, T_wj, debug=False):
        i = state["pair"]
        rec[f"mds_Twj0_{i}"], rec[f"mds_pw_{i}"] = np.array(T_wj0), np.array(p_w)
        rec[f"mds_pjt_{i}"], rec[f"mds_Twj_{i}"] = np.array(p_jt), np.array(T_wj)
        return orig_up(self, T_wj0, p_w, p_jt, T_wj, debug)

    def opt(self):
        i = state["pair"]
        x = orig_opt(self)
        rec[f"mds_x_{i}"] = x.copy()
        state["pair"] += 1
        return x

    md.MotionDistortionSolver.update_problem, md.MotionDistortionSolver.optimize_library = up, opt
    rs.RawROAMSystem.plot = lambda *a, **k: None

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(ri.REFERENCE_ROOT, "data"), os.path.join(tmp, "data"))
        os.chdir(tmp)
        try:
            with ri.quiet():
                system = rs.RawROAMSystem("tiny", paramFlags={"rejectOutliers": True, "useFMT": False})
                system.run(0, -1)
        finally:
            os.chdir(cwd)
    poses = np.array(system.estTraj.poses)
    rec["poses"] = poses
    rec["n_pairs"] = np.array(state["pair"])
    np.savez_compressed(os.path.join(GOLD, "tiny_stages.npz"), **rec)
    print("pairs recorded:", state["pair"], "poses", poses.shape)
    T.getTrackedPointsKLT, T.rejectOutliers, T.calculateTransformSVD = orig_klt, orig_rej, orig_svd

    # ---------------- clique fixture (reference's own outlier_test.npz) -----------
    orj = ri.ref_module("outlierRejection")
    d = np.load(os.path.join(ri.REFERENCE_ROOT, "outlier_test.npz"))
    with ri.quiet():
        _, _, mask = orj.rejectOutliers(d["prev_coord"], d["new_coord"])
    np.savez_compressed(os.path.join(GOLD, "clique_fixture.npz"), prev=d["prev_coord"], new=d["new_coord"],
                        mask=mask, thr=np.array(orj.DIST_THRESHOLD_PX))

    # ---------------- Kabsch fixture (archive/testTransformKLT2.py:15-190) --------
    src_txt = open(os.path.join(ri.REFERENCE_ROOT, "archive", "testTransformKLT2.py")).read()
    ns = {"np": np}
    start = src_txt.index("srcCoord")
    end = src_txt.index("])", src_txt.index("targetCoord")) + 2
    exec(src_txt[start:end], ns)  # two np.array literals only
    s = np.asarray(ns["srcCoord"], np.float32)
    t = np.asarray(ns["targetCoord"], np.float32)
    R, h = gk.calculateTransformSVD(s, t)
    np.savez_compressed(os.path.join(GOLD, "kabsch_fixture.npz"), src=s, tgt=t, R=R, h=h)

    # ---------------- SSC fixture -------------------------------------------------
    anms = ri.ref_module("ANMS")
    rng = np.random.default_rng(20240607)
    ssc = {}
    for j, (n, k, W, H) in enumerate([(3000, 200, 2024, 2024), (700, 200, 2024, 2024), (12000, 200, 2024, 2024),
                                       (5000, 1000, 2000, 2000), (250, 200, 2024, 2024)]):
        kp = np.column_stack([rng.uniform(0, H, n), rng.uniform(0, W, n), rng.choice([0.01, 5.005, 10.0], n)])
        kp = kp[np.argsort(kp[:, 2], kind="stable")]
        sel = anms.ssc(kp, k, 0.1, W, H)
        ssc[f"kp_{j}"], ssc[f"sel_{j}"], ssc[f"cfg_{j}"] = kp, sel, np.array([k, W, H])
    ssc["n"] = np.array(5)
    np.savez_compressed(os.path.join(GOLD, "ssc_fixture.npz"), **ssc)

    # ---------------- peaks fixture -----------------------------------------------
    gpc = ri.ref_module("getPointCloud")
    polar = pd.extractDataFromRadarImage(frames["raw_0"])[0]
    peaks = gpc.getPointCloudPolarInd(polar)
    np.savez_compressed(os.path.join(GOLD, "peaks_fixture.npz"), peaks=peaks.astype(np.int32))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
