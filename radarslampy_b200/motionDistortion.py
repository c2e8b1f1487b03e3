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
        if times is None:
            assert (period > 0)
        xy = _engine.engine().mds_undistort(np.asarray(v_j, dtype=np.float64), points[:, :2], period, times=times)
        return np.hstack((xy, np.ones((xy.shape[0], 1))))

    # ---- residual / Jacobian as plain NumPy (motionDistortion.py:154-263): debugging aids of the reference, never
    # called by its live path (optimize_library is); kept so that scripts which print them keep working
    def expected_observed_pts(self, T_wj):
        return np.linalg.inv(T_wj) @ self.p_w.T

    def _pose_matrix(self, params):
        c, s_ = np.cos(params[5]), np.sin(params[5])
        return np.array([[c, -s_, params[3]], [s_, c, params[4]], [0, 0, 1]])

    def error_vector(self, params):
        return self.info_vector * self.error(np.asarray(params, np.float64)[:3], self._pose_matrix(params))

    def error(self, v_j, T_wj):
        """log(e^2 / 2 + 1) of the point mismatch (x, y interleaved per point), then N * (v - v_prior)."""
        v_j = np.asarray(v_j, np.float64)
        und = MotionDistortionSolver.undistort(v_j, self.p_jt, times=self.dT)
        mismatch = self.expected_observed_pts(T_wj) - und.T
        e_p_i = np.log(np.square(mismatch[:2, :]) / 2 + 1)
        rel = self.T_wj0_inv @ T_wj
        prior = np.array([rel[0, 2], rel[1, 2], np.arctan2(rel[1, 0], rel[0, 0])]) / self.total_scan_time
        v_diff = v_j - prior
        v_diff[2] = (v_diff[2] + np.pi) % (2 * np.pi) - np.pi
        return np.hstack((e_p_i.flatten(order='F'), v_diff * e_p_i.shape[1]))

    def jacobian_vector(self, params):
        params = np.asarray(params, np.float64)
        return np.expand_dims(self.info_vector, axis=1) * self.jacobian(params[:3], self._pose_matrix(params))

    def jacobian(self, v_j, T_wj):
        """The reference's hand-written Jacobian [J_v | J_p] (motionDistortion.py:216-263).  It is NOT the derivative of
        error() (it differentiates e^2 / 2 + 1's argument, not the log, in places) and the reference never uses it;
        reproduced term by term."""
        v_j = np.asarray(v_j, np.float64)
        N = self.p_w.shape[0]
        und = MotionDistortionSolver.undistort(v_j, self.p_jt, times=self.dT)
        e = (self.expected_observed_pts(T_wj) - und.T)[:2]
        cd = e / (np.square(e) / 2 + 1)                                   # 2 x N
        c0, s0, c1, s1 = self.T_wj0[0, 0], self.T_wj0[1, 0], T_wj[0, 0], T_wj[1, 0]
        Tx, Ty = T_wj[0, 2], T_wj[1, 2]
        pwx, pwy = self.p_w[:, 0], self.p_w[:, 1]
        one = np.ones(N)
        Jp = np.array([[-c1 * one, -s1 * one, -pwx * s1 + pwy * c1 - c1 * Ty + s1 * Tx],
                       [s1 * one, -c1 * one, -pwx * c1 - pwy * s1 + s1 * Ty + c1 * Tx]]) * cd[:, None, :]
        Jp = Jp.transpose(2, 0, 1).reshape(2 * N, 3)
        Jp = np.vstack((Jp, np.array([[c0, s0, 0], [-s0, c0, 0], [0, 0, 1]]) / self.total_scan_time * N))
        x, y = self.p_jt[:, 0], self.p_jt[:, 1]
        th = v_j[2] * self.dT
        zero = np.zeros(N)
        Jv = np.array([[-self.dT, zero, self.dT * (np.sin(th) * x + np.cos(th) * y)],
                       [zero, -self.dT, self.dT * (-np.cos(th) * x + np.sin(th) * y)]]) * cd[:, None, :]
        Jv = np.vstack((Jv.transpose(2, 0, 1).reshape(2 * N, 3), np.eye(3) * N))
        return np.hstack((Jv, Jp))

    def optimize(self, max_iters=20):
        """motionDistortion.py:265-289: a stub in the reference (returns None)."""
        return None

    def optimize_library(self):
        """motionDistortion.py:295-325 -> [vx, vy, vtheta, x, y, theta]."""
        x, self.num_evals, self.cost = _engine.engine().mds_solve(
            self.T_wj0, self.p_w, self.p_jt, self.T_wj_initial, sigma_p=self.sigma_p, sigma_v=self.sigma_v,
            period=self.total_scan_time)
        return x
