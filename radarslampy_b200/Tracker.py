"""Drop-in for the reference's Tracker.py: per-pair front end (KLT -> outlier rejection ->
corrStatus fix-up) and the metric rigid transform.  Tracker.py:16-127."""
from typing import Tuple

import numpy as np

from .FMT import getRotationUsingFMT
from .getTransformKLT import calculateTransformSVD, getTrackedPointsKLT, visualize_transform
from .outlierRejection import rejectOutliers
from .parseData import RANGE_RESOLUTION_CART_M


class Tracker():

    def __init__(self, sequenceName: str, imgPathArr: list, filePaths: dict, paramFlags: dict) -> None:
        self.sequenceName = sequenceName
        self.imgPathArr = imgPathArr
        self.sequenceSize = len(self.imgPathArr)
        self.filePaths = filePaths
        self.paramFlags = paramFlags
        self.estTraj = None
        self.gtTraj = None

    def initTraj(self, estTraj, gtTraj=None):
        self.estTraj = estTraj
        self.gtTraj = gtTraj

    def track(self, prevImgCart: np.ndarray, currImgCart: np.ndarray, prevImgPolar: np.ndarray,
              currImgPolar: np.ndarray, featureCoord: np.ndarray, seqInd: int
              ) -> Tuple[np.ndarray, np.ndarray, float, np.ndarray]:
        """Tracker.py:35-106 -> (good_old, good_new, angleRotRad, corrStatus u8 [K,1]).
        angleRotRad is the FMT rotation prior (Tracker.py:62-63, rf_fmt_rotation); like the reference, the
        images are not pre-rotated by it (the `useFMT` branch is a no-op there, Tracker.py:66-72)."""
        angleRotRad, _, _ = getRotationUsingFMT(prevImgPolar, currImgPolar)
        good_new, good_old, bad_new, bad_old, corrStatus = getTrackedPointsKLT(prevImgCart, currImgCart, featureCoord)
        nFeatures = good_new.shape[0] + bad_new.shape[0]
        if self.paramFlags.get("rejectOutliers", True):
            good_old, good_new, pruning_mask = rejectOutliers(good_old, good_new)
        else:
            pruning_mask = np.ones(good_old.shape[0], dtype=bool)   # the reference leaves this unbound (NameError)
        rng = np.arange(nFeatures)
        corrStatus[rng[corrStatus.flatten().astype(bool)]] &= pruning_mask[:, np.newaxis]   # Tracker.py:103-104
        return good_old, good_new, angleRotRad, corrStatus

    def getTransform(self, srcCoord: np.ndarray, targetCoord: np.ndarray, pixel: bool) -> Tuple[np.ndarray, np.ndarray]:
        """Tracker.py:108-127: (R, h) with src = R @ target + h; h in metres unless pixel."""
        R, h = calculateTransformSVD(srcCoord, targetCoord)
        if not pixel:
            h *= RANGE_RESOLUTION_CART_M
        return R, h

    def plot(self, prevImg, currImg, good_old, good_new, seqInd, save=True, show=False):
        """Tracker.py:129-150: tracking overlay of one frame (RawROAMSystem.plot calls it every third frame,
        RawROAMSystem.py:371-377); saved to filePaths["imgSave"] as the reference does."""
        import os
        from matplotlib import pyplot as plt
        visualize_transform(prevImg, currImg, good_old, good_new, show=False)
        plt.title(f"Tracking on Image {seqInd:04d}")
        if save:
            plt.savefig(os.path.join(self.filePaths["imgSave"], f"{seqInd:04d}.jpg"))
        if show:
            plt.pause(0.01)
