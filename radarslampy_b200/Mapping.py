"""Drop-in for the reference's Mapping.py: keyframe bookkeeping between the KLT front end and
the motion-distortion solve.  Same class / method names and semantics as Mapping.py:20-174; the per-keyframe
array work runs on the device through the other drop-ins (rf_polar_peaks for the point cloud,
rf_mds_undistort for the undistorted feature points); plotting is out of scope."""
import numpy as np

from . import parseData
from .getPointCloud import getPointCloudPolarInd
from .motionDistortion import MotionDistortionSolver

ROT_THRESHOLD = 0.2                                       # Mapping.py:13-15
TRANS_THRESHOLD = 2.0
TRANS_THRESHOLD_SQ = TRANS_THRESHOLD * TRANS_THRESHOLD

RADAR_CART_CENTER = None


def getRotationMatrix(th, degrees=False):
    """utils.py:36-43."""
    if degrees:
        th = np.deg2rad(th)
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s], [s, c]])


def cartCenter(radarPolarImg: np.ndarray) -> np.ndarray:
    """Centre of the Cartesian image of a polar scan, in pixels (Mapping.py:52-56): the reference converts the whole
    image to learn its shape; the shape follows from the scan geometry (parseData.py:117-123)."""
    r = radarPolarImg.shape[1] // parseData.DOWNSAMPLE_FACTOR if parseData.DOWNSAMPLE_FACTOR > 1 else radarPolarImg.shape[1]
    return np.array([2 * r, 2 * r]) / 2


class Keyframe():
    """Mapping.py:20-125."""

    def __init__(self, globalPose: np.ndarray, featurePointsLocal: np.ndarray, radarPolarImg: np.ndarray,
                 velocity: np.ndarray) -> None:
        self.updateInfo(globalPose, featurePointsLocal, radarPolarImg, velocity)

    def updateInfo(self, globalPose: np.ndarray, featurePointsLocal: np.ndarray, radarPolarImg: np.ndarray,
                   velocity: np.ndarray) -> None:
        self.pose = globalPose
        self.radarPolarImg = radarPolarImg
        global RADAR_CART_CENTER
        if RADAR_CART_CENTER is None:
            RADAR_CART_CENTER = cartCenter(radarPolarImg)
        self.featurePointsLocal = featurePointsLocal
        self.prunedFeaturePoints = self.featurePointsLocal
        self.pointCloud = getPointCloudPolarInd(radarPolarImg)
        self.velocity = velocity
        self.featurePointsLocalUndistorted = MotionDistortionSolver.undistort(velocity, featurePointsLocal)[:, :2]
        self.prunedUndistortedLocals = self.featurePointsLocalUndistorted

    def copyFromOtherKeyframe(self, keyframe) -> None:
        # the reference omits the velocity argument here (Mapping.py:69-71, a TypeError if ever called)
        self.updateInfo(keyframe.pose, keyframe.featurePointsLocal, keyframe.radarPolarImg, keyframe.velocity)

    def _sensor_to_global(self, pts_m: np.ndarray) -> np.ndarray:
        """Rigid transform of metric sensor-frame points by this keyframe's pose (x, y, theta)."""
        c, s = np.cos(self.pose[2]), np.sin(self.pose[2])
        pts_m = np.asarray(pts_m, np.float64)
        return np.column_stack((c * pts_m[:, 0] - s * pts_m[:, 1] + self.pose[0], s * pts_m[:, 0] + c * pts_m[:, 1] + self.pose[1]))

    def convertFeaturesLocalToGlobal(self, featurePointsLocal: np.ndarray) -> np.ndarray:
        """Mapping.py:73-99: pixels -> metres about the image centre -> global frame of this keyframe's pose."""
        return self._sensor_to_global((np.asarray(featurePointsLocal, np.float64) - RADAR_CART_CENTER) * parseData.RANGE_RESOLUTION_CART_M)

    def getPrunedFeaturesGlobalPosition(self) -> np.ndarray:
        """Mapping.py:101-120: the pruned, undistorted local points (already metres) in the global frame."""
        return self._sensor_to_global(self.prunedUndistortedLocals)

    def pruneFeaturePoints(self, corrStatus: np.ndarray) -> None:
        keep = np.asarray(corrStatus).flatten().astype(bool)
        self.prunedFeaturePoints = self.prunedFeaturePoints[keep]
        self.prunedUndistortedLocals = self.prunedUndistortedLocals[keep]


class Map():
    """Mapping.py:128-174."""

    def __init__(self, sequenceName: str, estTraj, imgPathArr: list, filePaths: dict) -> None:
        self.sequenceName = sequenceName
        self.imgPathArr = imgPathArr
        self.sequenceSize = len(self.imgPathArr)
        self.filePaths = filePaths
        self.estTraj = estTraj
        self.mapPoints = []
        self.keyframes = []

    def updateInternalTraj(self, traj):
        self.estTraj = traj

    def isGoodKeyframe(self, keyframe: Keyframe) -> bool:
        srcPose, targetPose = self.keyframes[-1].pose, keyframe.pose
        if np.abs(srcPose[2] - targetPose[2]) >= ROT_THRESHOLD:
            return True
        return bool(((srcPose[0:2] - targetPose[0:2]) ** 2).sum() >= TRANS_THRESHOLD_SQ)

    def addKeyframe(self, keyframe: Keyframe) -> None:
        self.keyframes.append(keyframe)

    def plot(self, fig=None, subsampleFactor: int = 5, show: bool = False) -> None:
        """Mapping.py:176-199 (matplotlib is imported only here; plotting is not part of the hot path)."""
        from matplotlib import pyplot as plt
        pts = self.getMapPointsGlobal()
        plt.scatter(pts[::subsampleFactor, 0], pts[::subsampleFactor, 1], marker='+', color='g', alpha=.8, label='Map Points')
        if show:
            plt.show()

    def getMapPointsGlobal(self) -> np.ndarray:
        """The point set Map.plot scatters (Mapping.py:187-191): pruned global feature positions of every keyframe."""
        pts = [kf.getPrunedFeaturesGlobalPosition() for kf in self.keyframes]
        return np.vstack(pts) if pts else np.empty((0, 2))
