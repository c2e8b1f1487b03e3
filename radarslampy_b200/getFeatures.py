"""Drop-in for the reference's getFeatures.py: feature detection + adaptive non-maximal suppression.

method="doh" (the reference's default, getFeatures.py:13-18) is skimage.feature.blob_doh on the GPU (rf_detect_doh,
csrc/k_doh.cu): float64 integral image, box-filter Hessian determinant per sigma, 3x3x3 peak_local_max, overlap
pruning — the published scikit-image 0.19.2 algorithm as restated in oracle/doh_restate.py.  scikit-image is not part
of the reference tree and not installable here, so parity with the real package is UNPINNED (INTEGRATION.md §3 lists
the three implementation-defined points).  method="mineig" is the Sobel / structure-tensor minimum-eigenvalue response
named by BASELINE.json's north_star (rf_detect mode 0, pinned to cv2.cornerMinEigenVal).  "dog" / "log" (which the
reference can route to skimage but never uses) raise NotImplementedError.  Everything after the detector —
adaptiveNMS -> ANMS.ssc, the (x, y) flip, the order-preserving de-duplication in appendNewFeatures — follows
getFeatures.py:66-118."""
import numpy as np

from . import _engine
from .ANMS import ssc

DEFAULT_FEATURE_PARAMS = dict(min_sigma=0.01, max_sigma=10, num_sigma=3, threshold=.0005, method="doh")   # getFeatures.py:13-18
DETECTOR_QUALITY_LEVEL = 0.01            # method="mineig": candidates above this fraction of the strongest response
MAX_CANDIDATES = 65536

PERCENT_FEATURE_LOSS_THRESHOLD = 0.75
N_FEATURES_BEFORE_RETRACK = 60          # getFeatures.py:57


def getBlobsFromCart(cartImage: np.ndarray, min_sigma: int = 1, max_sigma: int = 30, num_sigma: int = 10,
                     threshold=0.01, method="doh") -> np.ndarray:
    """getFeatures.py:22-53 -> [K, 3] rows [r, c, sigma]."""
    if method not in ("doh", "dog", "log", "mineig"):
        raise NotImplementedError(f"{method} not implemented! Use one of ('doh', 'dog', 'log')")
    if method in ("dog", "log"):
        raise NotImplementedError(f"method={method!r}: only the determinant-of-Hessian detector the reference uses ('doh') and the "
                                  "structure-tensor response ('mineig') are built")
    fe, frame = _engine.frame_of(cartImage)
    if method == "doh":
        return fe.detect_doh(frame, min_sigma, max_sigma, num_sigma, threshold)
    # strongest first; sigma is the constant 1.0 (a 3x3 structure tensor has one scale), so adaptiveNMS' argsort keeps the order
    cand, _ = fe.detect(frame, -DETECTOR_QUALITY_LEVEL, cap=MAX_CANDIDATES)
    blobs = cand.copy()
    blobs[:, 2] = 1.0
    return blobs


def calculateFeatureLossThreshold(nInitialFeatures):
    return 80                            # getFeatures.py:61-64


def adaptiveNMS(img, blobs, ret_points=200, tolerance=0.1):
    """getFeatures.py:66-72.  The sort is stable here (the reference's quicksort leaves ties unspecified)."""
    H, W = img.shape
    sort_ind = np.argsort(blobs[:, 2], kind="stable")
    keypoints = blobs[sort_ind, :]
    return ssc(keypoints, ret_points, tolerance, W, H)


def getFeatures(img, feature_params: dict = DEFAULT_FEATURE_PARAMS):
    """getFeatures.py:74-95 -> (blobCoord [K, 2] (x, y), blobRadii [K])."""
    blobs = getBlobsFromCart(img, **feature_params)
    blobs = adaptiveNMS(img, blobs)
    return np.fliplr(blobs[:, :2]), blobs[:, 2]


def appendNewFeatures(srcImg, oldFeaturesCoord):
    """getFeatures.py:98-118 -> (f32 [K', 2], retrack threshold)."""
    newFeatureCoord, _ = getFeatures(srcImg)
    featurePtSrc = np.vstack((oldFeaturesCoord, newFeatureCoord))
    _, idx = np.unique(featurePtSrc, axis=0, return_index=True)
    featurePtSrc = np.ascontiguousarray(featurePtSrc[np.sort(idx)]).astype(np.float32)
    return featurePtSrc, calculateFeatureLossThreshold(featurePtSrc.shape[0])
