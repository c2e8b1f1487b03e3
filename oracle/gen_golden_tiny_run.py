"""TEST INFRASTRUCTURE ONLY — the whole of data/tiny through the UNMODIFIED reference driver.

    python -m oracle.gen_golden_tiny_run

1. copies the 11 raw scans of /root/reference/data/tiny (radar/*.png + radar.timestamps; input DATA of the
   reference, 5.3 MB) to tests/golden/tiny_scans/, so that the GPU box — which has no /root/reference — can run the
   reference's own configs[0] sequence end to end;
2. runs RawROAMSystem("tiny").run() of the unmodified reference (oracle/ref_import.py: inert matplotlib / tkinter
   shims) with `skimage.feature.blob_doh` bound to the restatement oracle/doh_restate.blob_doh (scikit-image is not
   installable here: detector parity is UNPINNED, everything downstream of the detector is the reference's own code
   and third-party calls), intercepting the stage boundaries exactly like oracle/gen_golden.py, and writes
   tests/golden/tiny_run_doh.npz:
     feat_in_i, klt_status_i, klt_good_new_i, rej_mask_i, svd_R_i / svd_h_i, mds_x_i   for every pair i = 0..9
     poses [11, 3], n_pairs, first_features (appendNewFeatures on frame 0), retrack_i (re-detection after pair i)

np.argsort(blobs[:, 2]) inside the reference's adaptiveNMS is unstable; on this image's NumPy (2.3.5, AVX-512) it is
the identity for equal sigmas, which is what the product's stable partition assumes (recorded as argsort_identity)."""
import glob
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, ROOT)
    from oracle import doh_restate as D
    from oracle import ref_import as ri
    ri.activate()
    src = os.path.join(ri.REFERENCE_ROOT, "data", "tiny")
    dst = os.path.join(GOLD, "tiny_scans")
    os.makedirs(os.path.join(dst, "radar"), exist_ok=True)
    for p in sorted(glob.glob(os.path.join(src, "radar", "*.png"))):
        shutil.copyfile(p, os.path.join(dst, "radar", os.path.basename(p)))
    shutil.copyfile(os.path.join(src, "radar.timestamps"), os.path.join(dst, "radar.timestamps"))

    import skimage.feature as skf
    skf.set_blob_doh(lambda image, **kw: D.blob_doh(image, **kw))
    rec, state = {}, {"pair": 0}
    import Tracker as T
    import getFeatures as gf
    import motionDistortion as md
    import RawROAMSystem as rs
    orig_klt, orig_rej, orig_svd = T.getTrackedPointsKLT, T.rejectOutliers, T.calculateTransformSVD
    orig_append = rs.appendNewFeatures

    def klt(a, b, feat):
        i = state["pair"]
        out = orig_klt(a, b, feat)
        rec[f"feat_in_{i}"] = np.asarray(feat).copy()
        rec[f"klt_good_new_{i}"], rec[f"klt_status_{i}"] = np.asarray(out[0]).copy(), np.asarray(out[4]).copy()
        return out

    def rej(a, b):
        out = orig_rej(a, b)
        rec[f"rej_mask_{state['pair']}"] = out[2].copy()
        return out

    def svd(a, b):
        R, h = orig_svd(a, b)
        rec[f"svd_R_{state['pair']}"], rec[f"svd_h_{state['pair']}"] = R.copy(), h.copy()
        return R, h

    def append(img, old):
        out = orig_append(img, old)
        if len(old) == 0 and "first_features" not in rec:
            rec["first_features"] = out[0].copy()
        else:
            rec[f"retrack_{state['pair'] - 1}"] = np.array(1)
            rec[f"feat_after_retrack_{state['pair'] - 1}"] = out[0].copy()
        return out

    T.getTrackedPointsKLT, T.rejectOutliers, T.calculateTransformSVD = klt, rej, svd
    rs.appendNewFeatures = append
    orig_opt = md.MotionDistortionSolver.optimize_library

    def opt(self):
        x = orig_opt(self)
        rec[f"mds_x_{state['pair']}"] = x.copy()
        state["pair"] += 1
        return x

    md.MotionDistortionSolver.optimize_library = opt
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
    rec["poses"] = np.array(system.estTraj.poses)
    rec["n_pairs"] = np.array(state["pair"])
    rec["argsort_identity"] = np.array(bool(np.array_equal(np.argsort(np.full(343, 10.0)), np.arange(343))))
    np.savez_compressed(os.path.join(GOLD, "tiny_run_doh.npz"), **rec)
    print("pairs:", state["pair"], "poses", rec["poses"].shape, "retracks", sorted(k for k in rec if k.startswith("retrack_")),
          "features per pair", [len(rec[f"feat_in_{i}"]) for i in range(state["pair"])])


if __name__ == "__main__":
    main()
