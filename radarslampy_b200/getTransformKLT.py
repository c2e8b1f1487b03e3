"""Drop-in for the hot-path functions of the reference's getTransformKLT.py: pyramidal Lucas-Kanade
tracking (rf_klt, csrc/k_klt.cu) and the 2-D Kabsch rigid transform (rf_kabsch, csrc/k_solve.cu)."""
import numpy as np

from . import _engine
from .getFeatures import appendNewFeatures

N_FEATURES_BEFORE_RETRACK = 60          # getTransformKLT.py:18
ERR_THRESHOLD = 10                      # getTransformKLT.py:84
LK_PARAMS = dict(maxLevel=3, criteria=(3, 10, 0.03))   # getTransformKLT.py:77-81 (EPS | COUNT, 10, 0.03)


def calculateTransformSVD(srcCoords: np.ndarray, targetCoords: np.ndarray):
    """getTransformKLT.py:129-162 -> (R f64 [2,2], h f64 [2,1]) with  src = R @ target + h."""
    return _engine.engine().kabsch(srcCoords, targetCoords)


def getTrackedPointsKLT(srcImg: np.ndarray, targetImg: np.ndarray, blobCoordSrc: np.ndarray):
    """getTransformKLT.py:317-381 -> (good_new, good_old, bad_new, bad_old, correspondenceStatus u8 [K,1])."""
    featurePtSrc = np.ascontiguousarray(np.asarray(blobCoordSrc)[:, :2]).astype(np.float32)
    global N_FEATURES_BEFORE_RETRACK
    if featurePtSrc.shape[0] < N_FEATURES_BEFORE_RETRACK:
        featurePtSrc, N_FEATURES_BEFORE_RETRACK = appendNewFeatures(srcImg, featurePtSrc)
    fe, prev = _engine.frame_of(srcImg)
    _, nxt = _engine.frame_of(targetImg, fe)
    # status already carries  status &= (err < ERR_THRESHOLD)  (getTransformKLT.py:365)
    nextPts, status, _ = fe.klt(prev, nxt, featurePtSrc, apply_err_gate=True)
    good = (status == 1).flatten()
    return nextPts[good, :], featurePtSrc[good, :], nextPts[~good, :], featurePtSrc[~good, :], status


def visualize_transform(*args, **kwargs):
    """getTransformKLT.py:165-230 is a matplotlib debugging plot (imported, never called on the hot path by
    Tracker.py:9); it is outside this front end."""
    raise NotImplementedError("visualize_transform is a plotting helper of the reference and is not part of the drop-in")
