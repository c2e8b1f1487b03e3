"""Drop-in for the reference's motionDistortion.py: the 6-parameter motion-distortion nonlinear
least squares (rf_mds_solve / rf_mds_undistort, csrc/k_solve.cu).  Class and method names follow
motionDistortion.py:38-325."""
import numpy as np

from . import _engine

RADAR_SCAN_FREQUENCY = 4    # motionDistortion.py:36
VERBOSE = False


def _homogenize(points):
    points = np.asarray(points)
    if points.shape[1] == 3:
        return points
    return np.hstack((points, np.ones((points.shape[0], 1))))


class MotionDistortionSolver():
    def __init__(self, sigma_p, sigma_v, frequency=RADAR_SCAN_FREQUENCY):
        self.total_scan_time = 1 / frequency
        self.sigma_p = np.diag(sigma_p)      # motionDistortion.py:75-76: only the diagonals are used
        self.sigma_v = np.diag(sigma_v)

    def update_problem(self, T_wj0, p_w, p_jt, T_wj, debug=False):
        """motionDistortion.py:80-99."""
        assert (p_w.shape == p_jt.shape)
        self.T_wj0 = np.asarray(T_wj0, dtype=np.float64)
        self.T_wj0_inv = np.linalg.inv(self.T_wj0)
        self.p_w = _homogenize(p_w)
        self.p_jt = _homogenize(p_jt)
        self.T_wj_initial = np.asarray(T_wj, dtype=np.float64)
        self.debug = debug
        self.v_j_initial = self.infer_velocity(self.T_wj0_inv @ self.T_wj_initial)
        self.dT = MotionDistortionSolver.compute_time_deltas(self.total_scan_time, self.p_jt)
        sigma_vector = np.concatenate((np.tile(self.sigma_p, self.p_jt.shape[0]), self.sigma_v))
        self.info_vector = 1 / sigma_vector

    def infer_velocity(self, transform):
        dx, dy = transform[0, 2], transform[1, 2]
        dtheta = np.arctan2(transform[1, 0], transform[0, 0])
        return np.array([dx, dy, dtheta]) / self.total_scan_time

    @staticmethod
    def compute_time_deltas(period, points):
        """motionDistortion.py:107-124."""
        points = np.asarray(points)
        return period * np.arctan2(-points[:, 1], -points[:, 0]) / (2 * np.pi)

    @staticmethod
    def undistort(v_j, points, period=1 / RADAR_SCAN_FREQUENCY, times=None):
        """motionDistortion.py:127-153 -> [N, 3] homogeneous undistorted points."""
        points = _homogenize(points)
        if times is not None:
            raise NotImplementedError("explicit `times` is never passed by the reference (Mapping.py:65)")
        assert (period > 0)
        xy = _engine.engine().mds_undistort(np.asarray(v_j, dtype=np.float64), points[:, :2], period)
        return np.hstack((xy, np.ones((xy.shape[0], 1))))

    def optimize_library(self):
        """motionDistortion.py:295-325 -> [vx, vy, vtheta, x, y, theta]."""
        x, self.num_evals, self.cost = _engine.engine().mds_solve(
            self.T_wj0, self.p_w, self.p_jt, self.T_wj_initial, sigma_p=self.sigma_p, sigma_v=self.sigma_v,
            period=self.total_scan_time)
        return x
