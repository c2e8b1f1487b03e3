"""Drop-in for the reference's ANMS.py: Suppression via Square Covering on the GPU (rf_ssc,
csrc/k_features.cu), bit-exact with ANMS.py:5-102."""
import numpy as np

from . import _engine


def ssc(keypoints, num_ret_points, tolerance, cols, rows):
    """keypoints [n, 3] (row, col, sigma) in priority order -> selected rows [m, 3] in selection order."""
    kp = np.asarray(keypoints, dtype=np.float64).reshape(-1, 3)
    if num_ret_points == 1:
        raise ZeroDivisionError("float division by zero")        # ANMS.py:19
    sel = _engine.engine().ssc(kp, num_ret_points, tolerance, cols, rows)
    out = np.empty((0, 3))
    return np.vstack((out, kp[sel])) if len(sel) else out
