"""TEST INFRASTRUCTURE ONLY — CPU feature detector handed to oracle/ref_system.run_odometry (the role of
getFeatures.appendNewFeatures, getFeatures.py:98-118) when the whole loop has to run on the host: the CPU baseline of
bench.py's `chained` leg and the CPU-only tests.  The reference detects with skimage.feature.blob_doh, which is not
installed here (parity unpinned); this stand-in is the structure-tensor response named by BASELINE.json's
north_star through the library the reference would use for it (cv2.cornerMinEigenVal), followed by exactly the
selection chain of the product: 1 % quality level + 3x3 NMS in cv2.goodFeaturesToTrack order (oracle.restate.nms_select),
the strongest 65 536 candidates, the reference's own SSC (oracle.restate.ssc, pinned to ANMS.py), np.fliplr, vstack +
order-preserving np.unique (getFeatures.py:88,108-112).  It is cheaper than blob_doh on a 2024^2 image, so a CPU rate
measured with it flatters the CPU.  The product never imports this."""
import numpy as np

from . import restate as R


def select_from_response(resp, old_xy, quality=0.01, num_ret=200, tol=0.1, max_candidates=65536):
    """Selection + append given a response map (north_star: selection is bit-exact given the same responses)."""
    thr = float(np.float32(float(resp.max()) * quality))
    cand = R.nms_select(resp, thr)[:max_candidates]
    kp = np.column_stack([cand[:, 0], cand[:, 1], np.ones(len(cand))])
    sel = R.ssc(kp, num_ret, tol, resp.shape[1], resp.shape[0])
    new = np.fliplr(kp[sel][:, :2])
    pts = np.vstack((np.asarray(old_xy, np.float64).reshape(-1, 2), new))
    _, idx = np.unique(pts, axis=0, return_index=True)
    return np.ascontiguousarray(pts[np.sort(idx)]).astype(np.float32)


def detect_min_eig(cart_f32, old_xy):
    import cv2
    resp = cv2.cornerMinEigenVal(np.ascontiguousarray(cart_f32, np.float32), 3, ksize=3)
    return select_from_response(resp, old_xy)
