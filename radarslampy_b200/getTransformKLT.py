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


def visualize_transform(prevImg: np.ndarray, currImg: np.ndarray, prevFeatureCoord: np.ndarray, newFeatureCoord: np.ndarray,
                        alpha: float = 1, extraLabel: str = "", show: bool = False) -> None:
    """getTransformKLT.py:20-73: the tracking overlay Tracker.plot draws (matplotlib is imported only here; plotting is
    not part of the hot path and touches no device state)."""
    from matplotlib import pyplot as plt
    if currImg is not None:
        plt.imshow(currImg)
    if newFeatureCoord is not None or alpha == 0:
        plt.scatter(newFeatureCoord[:, 0], newFeatureCoord[:, 1], marker='+', color='red', alpha=alpha,
                    label=f'Tracked Features{extraLabel}')
    if prevFeatureCoord is not None or alpha == 0:
        plt.scatter(prevFeatureCoord[:, 0], prevFeatureCoord[:, 1], marker='.', color='yellow', alpha=alpha,
                    label=f'Previous Features{extraLabel}')
    plt.legend()
    plt.axis("off")
    plt.tight_layout()
    if show:
        plt.show()
