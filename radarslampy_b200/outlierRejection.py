"""Drop-in for the reference's outlierRejection.py: distance-consistency graph and the first
largest maximal clique in networkx order (rf_reject_outliers, csrc/k_clique.cu)."""
import numpy as np

from . import _engine
from .parseData import RANGE_RESOLUTION_CART_M

DIST_THRESHOLD_M = 0.5                                             # outlierRejection.py:10
DIST_THRESHOLD_PX = DIST_THRESHOLD_M / RANGE_RESOLUTION_CART_M     # outlierRejection.py:11
DISTSQ_THRESHOLD_PX = DIST_THRESHOLD_PX * DIST_THRESHOLD_PX


def rejectOutliers(prev_coord: np.ndarray, new_coord: np.ndarray):
    """outlierRejection.py:16-95 -> (pruned_prev_coord, pruned_new_coord, pruning_mask bool [K])."""
    assert prev_coord.shape == new_coord.shape, "Coordinates should be the same shape"
    fe = _engine.engine()
    if fe.cfg.dist_thr_px != DIST_THRESHOLD_PX:
        raise RuntimeError("engine threshold differs from DIST_THRESHOLD_PX")
    mask, _, _ = fe.reject_outliers(prev_coord, new_coord)
    return prev_coord[mask], new_coord[mask], mask
