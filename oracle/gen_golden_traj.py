"""TEST INFRASTRUCTURE ONLY — generates tests/golden/traj_fixture.npz by running the UNMODIFIED reference's
trajectoryPlotting.Trajectory / getGroundTruthTrajectory / computePosesRMSE (imported from /root/reference
through oracle/shims) on data/tiny's ground truth and on seeded relative transforms.

    python -m oracle.gen_golden_traj
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import as ri
    ri.activate()
    with ri.quiet():
        import trajectoryPlotting as TP
    gt_path = os.path.join(ri.REFERENCE_ROOT, "data", "tiny", "gt", "radar_odometry.csv")
    rows = [l.strip().split(",") for l in open(gt_path)][1:]
    gt = TP.getGroundTruthTrajectory(gt_path)
    # seeded relative transforms chained by the reference's appendRelativeTransform / appendRelativeDeltas
    rng = np.random.default_rng(7)
    P = 40
    th = rng.normal(0.02, 0.01, P)
    R = np.stack([[[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]] for t in th])
    h = rng.normal([2.0, 0.1], 0.2, (P, 2))
    start = np.array([1.5, -2.0, 0.3])
    ts = gt.timestamps[0] + np.arange(P + 1) * 250000
    tr = TP.Trajectory([ts[0]], [start])
    td = TP.Trajectory([ts[0]], [start])
    for k in range(P):
        tr.appendRelativeTransform(ts[k + 1], R[k], h[k].reshape(2, 1))
        td.appendRelativeDeltas(ts[k + 1], [h[k, 0], h[k, 1], th[k]])
    tq = np.linspace(ts[0], ts[-1], 57)
    np.savez_compressed(os.path.join(GOLD, "traj_fixture.npz"),
                        gt_csv_cols=np.array([[r[2], r[3], r[7]] for r in rows], np.float64),
                        gt_csv_ts=np.array([int(r[9]) for r in rows], np.int64),
                        gt_timestamps=gt.timestamps, gt_poses=gt.poses,
                        gt_query=gt.timestamps[:-1] + 1000, gt_interp=gt.getPoseAtTimes(gt.timestamps[:-1] + 1000),
                        R=R, h=h, start=start, ts=ts, chain_left=tr.poses, chain_deltas=td.poses,
                        tq=tq, chain_interp=tr.getPoseAtTimes(tq),
                        rmse=np.float64(TP.computePosesRMSE(tr.getPoseAtTimes(tq), td.getPoseAtTimes(tq))))
    print("traj_fixture.npz:", gt.poses.shape, tr.poses[-1], td.poses[-1])


if __name__ == "__main__":
    main()
