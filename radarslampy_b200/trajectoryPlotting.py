"""Drop-in for the numeric half of the reference's trajectoryPlotting.py: trajectory
bookkeeping, ground-truth integration, cubic pose interpolation and RMSE.  Same names, arguments and
return values as trajectoryPlotting.py:11-123,183-236; the plot functions import matplotlib lazily.

`Trajectory.from_relative_transforms` chains a whole batch of per-pair (R, h) results on the device
(rf_chain_poses, csrc/k_traj.cu) — the concatenation step of the multi-GPU gather."""
import csv

import numpy as np

from . import _engine


def normalize_angles(th):
    """utils.py:29-33."""
    return (th + np.pi) % (2 * np.pi) - np.pi


def convertPoseToTransform(poses):
    """utils.py:46-73: (3,) or (N, 3) -> (3, 3) or (N, 3, 3)."""
    poses = np.asarray(poses, np.float64)
    single = poses.ndim == 1
    p = np.atleast_2d(poses)
    T = np.zeros((p.shape[0], 3, 3))
    c, s = np.cos(p[:, 2]), np.sin(p[:, 2])
    T[:, 0, 0], T[:, 0, 1], T[:, 1, 0], T[:, 1, 1] = c, -s, s, c
    T[:, 0, 2], T[:, 1, 2], T[:, 2, 2] = p[:, 0], p[:, 1], 1
    return T[0] if single else T


def convertTransformToPose(pose_transforms):
    """utils.py:76-94."""
    T = np.asarray(pose_transforms, np.float64)
    single = T.ndim == 2
    T = T.reshape(-1, 3, 3)
    out = np.stack([T[:, 0, 2], T[:, 1, 2], np.arctan2(T[:, 1, 0], T[:, 0, 0])], axis=1)
    return out[0] if single else out


def _cubic_not_a_knot(x, y, xq):
    """scipy.interpolate.interp1d(x, y, kind='cubic', bounds_error=False)(xq): the C2 cubic spline with not-a-knot
    end conditions (scipy builds the same function as a B-spline); NaN outside [x0, xn]."""
    x, y, xq = np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(xq, np.float64)
    n = x.size
    if n < 4:
        raise ValueError("cubic interpolation needs at least 4 points")
    dx = np.diff(x)
    if np.any(dx <= 0):
        raise ValueError("timestamps must be strictly increasing")
    slope = np.diff(y) / dx
    lo, di, up, b = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)   # sub-, main, super-diagonal, rhs
    di[1:-1] = 2 * (dx[:-1] + dx[1:])
    up[1:-1] = dx[:-1]
    lo[1:-1] = dx[1:]
    b[1:-1] = 3 * (dx[1:] * slope[:-1] + dx[:-1] * slope[1:])
    d = x[2] - x[0]
    di[0], up[0] = dx[1], d
    b[0] = ((dx[0] + 2 * d) * dx[1] * slope[0] + dx[0] ** 2 * slope[1]) / d
    d = x[-1] - x[-3]
    di[-1], lo[-1] = dx[-2], d
    b[-1] = (dx[-1] ** 2 * slope[-2] + (2 * d + dx[-1]) * dx[-2] * slope[-1]) / d
    for i in range(1, n):                      # Thomas algorithm
        w = lo[i] / di[i - 1]
        di[i] -= w * up[i - 1]
        b[i] -= w * b[i - 1]
    s = np.zeros(n)
    s[-1] = b[-1] / di[-1]
    for i in range(n - 2, -1, -1):
        s[i] = (b[i] - up[i] * s[i + 1]) / di[i]
    k = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, n - 2)
    t = xq - x[k]
    c2 = (3 * slope[k] - 2 * s[k] - s[k + 1]) / dx[k]
    c3 = (s[k] + s[k + 1] - 2 * slope[k]) / dx[k] ** 2
    out = y[k] + t * (s[k] + t * (c2 + t * c3))
    out = np.where((xq < x[0]) | (xq > x[-1]), np.nan, out)
    return out


class Trajectory():
    """trajectoryPlotting.py:11-101."""

    def __init__(self, timestamps, poses):
        self.timestamps = np.array(timestamps)
        self.poses = np.array(poses)
        self.pose_transform = convertPoseToTransform(self.poses[-1])

    @classmethod
    def from_relative_transforms(cls, timestamps, R, h, start_pose=(0.0, 0.0, 0.0)):
        """Whole trajectory from P relative transforms at once (device prefix product): equivalent to P calls of
        appendRelativeTransform starting from start_pose.  timestamps: P + 1 values."""
        poses = _engine.engine().chain_poses(R, h, start_pose=start_pose, left_multiply=True)
        return cls(np.asarray(timestamps), poses)

    def getGroundTruthDeltasAtTime(self, time):
        return self.gt_deltas[time]

    def appendRelativeDeltas(self, time, d_xyth):
        dx, dy, dth = d_xyth
        self.timestamps = np.append(self.timestamps, time)
        x, y, th = self.poses[-1]
        x += dx * np.cos(th) - dy * np.sin(th)
        y += dx * np.sin(th) + dy * np.cos(th)
        th += dth
        self.poses = np.vstack((self.poses, [x, y, th]))

    def appendRelativeTransform(self, time, R, h):
        self.timestamps = np.append(self.timestamps, time)
        A = np.block([[np.asarray(R, np.float64).reshape(2, 2), np.asarray(h, np.float64).reshape(2, 1)], [np.zeros((1, 2)), 1]])
        self.pose_transform = A @ self.pose_transform
        self.poses = np.vstack((self.poses, convertTransformToPose(self.pose_transform)))

    def appendAbsoluteTransform(self, time, pose):
        self.timestamps = np.append(self.timestamps, time)
        self.poses = np.vstack((self.poses, pose))

    def plot(self, title='My Trajectory', savePath=False):
        """trajectoryPlotting.py:103-113."""
        from matplotlib import pyplot as plt
        plt.clf()
        plt.plot(self.poses[:, 0], self.poses[:, 1], 'b-')
        plt.xlabel('x [m]')
        plt.ylabel('y [m]')
        plt.grid(True)
        plt.axis('square')
        plt.title(title)
        if savePath:
            plt.tight_layout()
            plt.savefig(savePath)

    def getPoseAtTimes(self, times):
        """Cubic interpolation of x, y, theta at `times`; nearest recorded pose if there are too few points."""
        scalar = np.ndim(times) == 0
        tq = np.atleast_1d(np.asarray(times, np.float64))
        try:
            ts = np.asarray(self.timestamps, np.float64)
            poses = np.vstack([_cubic_not_a_knot(ts, self.poses[:, c], tq) for c in range(3)]).T
        except ValueError:
            poses = np.zeros((len(tq), 3))
            for i, t in enumerate(tq):
                poses[i, :] = self.poses[np.argmin(np.abs(self.timestamps - t))]
        if poses.shape[0] == 1 and scalar and isinstance(times, int):
            poses = poses[0, :]
        return poses


def plotGtAndEstTrajectory(gtTraj, estTraj, title='GT and EST Trajectories', info=None, savePath=None, arrow=False):
    """trajectoryPlotting.py:125-180 (RawROAMSystem.plotTraj draws this; matplotlib is imported only here)."""
    from matplotlib import pyplot as plt
    if savePath is not None:
        plt.clf()
    t0, t1 = estTraj.timestamps[0], estTraj.timestamps[-1]
    timestamps = [t for t in gtTraj.timestamps if t0 <= t <= t1]
    gtPoses, estPoses = gtTraj.getPoseAtTimes(timestamps), estTraj.getPoseAtTimes(timestamps)
    if arrow:
        for poses, c, label in ((gtPoses, 'b', "Ground Truth"), (estPoses, 'r', "Estimated")):   # utils.quiver
            plt.quiver(poses[:, 0], poses[:, 1], np.cos(poses[:, 2]), np.sin(poses[:, 2]), color=c, width=0.02, scale=10, alpha=.5, label=label)
    else:
        plt.plot(gtPoses[:, 0], gtPoses[:, 1], 'b-', label='Ground Truth')
        plt.plot(estPoses[:, 0], estPoses[:, 1], 'r-', label='Estimated')
    if info is not None:
        plt.text(0.01, 0.99, info, horizontalalignment='left', verticalalignment='top', transform=plt.gca().transAxes, fontsize='small')
    plt.xlabel('x [m]')
    plt.ylabel('y [m]')
    plt.grid(True)
    plt.legend()
    plt.axis('square')
    plt.title(f'{title}: RMSE={computePosesRMSE(gtPoses, estPoses):.2f}')
    if savePath:
        plt.tight_layout()
        plt.savefig(savePath)


def computePosesRMSE(gtPoses, estPoses):
    """trajectoryPlotting.py:116-123."""
    euclidean_err = np.linalg.norm(np.asarray(gtPoses)[:, :-1] - np.asarray(estPoses)[:, :-1], axis=-1)
    return np.sqrt(np.mean(euclidean_err ** 2))


def evaluateTrajectoryRMSE(gtTraj, estTraj):
    """The number plotGtAndEstTrajectory puts in its title (trajectoryPlotting.py:139-150,175): RMSE over the
    ground-truth timestamps inside the estimated trajectory's time span."""
    t0, t1 = estTraj.timestamps[0], estTraj.timestamps[-1]
    ts = [t for t in gtTraj.timestamps if t0 <= t <= t1]
    return computePosesRMSE(gtTraj.getPoseAtTimes(ts), estTraj.getPoseAtTimes(ts))


def getGroundTruthTrajectory(gtPath):
    """trajectoryPlotting.py:183-213: integrate radar_odometry.csv deltas into a Trajectory."""
    with open(gtPath) as gt_file:
        gt_reader = csv.reader(gt_file)
        next(gt_file)
        gt_timestamps, gt_poses, d_xyths = [], [], {}
        x, y, th = 0, 0, 0
        for row in gt_reader:
            timestamp = int(row[9])
            gt_timestamps.append(timestamp)
            dx, dy, dth = float(row[2]), float(row[3]), float(row[7])
            x += dx * np.cos(th) + dy * -np.sin(th)
            y += dx * np.sin(th) + dy * np.cos(th)
            th = normalize_angles(th + dth)
            gt_poses.append([x, y, th])
            d_xyths[timestamp] = [dx, dy, dth]
    gt_traj = Trajectory(np.array(gt_timestamps), np.array(gt_poses))
    gt_traj.gt_deltas = d_xyths
    return gt_traj


def getGroundTruthTrajectoryGPS(gtPath):
    """trajectoryPlotting.py:216-236: gps.csv -> Trajectory of (x, y, 0) at the source timestamps."""
    stamps, poses = [], []
    with open(gtPath) as f:
        next(f)                                      # header
        for row in csv.reader(f):
            stamps.append(int(row[0]))
            poses.append([float(row[2]), float(row[3]), 0])
    return Trajectory(np.array(stamps), np.array(poses))
