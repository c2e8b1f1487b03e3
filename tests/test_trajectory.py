"""CPU: the trajectory drop-in (radarslampy_b200/trajectoryPlotting.py) against goldens recorded
from the unmodified reference (tests/golden/traj_fixture.npz, oracle/gen_golden_traj.py) and live scipy."""
import os

import numpy as np
import pytest

from radarslampy_b200 import trajectoryPlotting as TP


def test_cubic_interpolation_matches_scipy():
    interp = pytest.importorskip("scipy.interpolate")
    rng = np.random.default_rng(0)
    for n in (4, 5, 11, 400):
        x = np.cumsum(rng.uniform(0.5, 2.0, n))
        y = rng.normal(size=n).cumsum()
        xq = np.concatenate([rng.uniform(x[0], x[-1], 200), x, [x[0] - 1.0, x[-1] + 1.0]])
        want = interp.interp1d(x, y, kind="cubic", bounds_error=False)(xq)
        got = TP._cubic_not_a_knot(x, y, xq)
        assert np.array_equal(np.isnan(want), np.isnan(got))
        ok = ~np.isnan(want)
        assert np.abs(want[ok] - got[ok]).max() <= 1e-9 * max(1.0, np.abs(y).max())
    with pytest.raises(ValueError):
        TP._cubic_not_a_knot([0, 1, 2], [0, 1, 0], [0.5])


def test_trajectory_matches_reference_goldens(golden):
    g = golden["traj_fixture"]
    P = g["R"].shape[0]
    tr = TP.Trajectory([g["ts"][0]], [g["start"]])
    td = TP.Trajectory([g["ts"][0]], [g["start"]])
    for k in range(P):
        tr.appendRelativeTransform(g["ts"][k + 1], g["R"][k], g["h"][k].reshape(2, 1))
        td.appendRelativeDeltas(g["ts"][k + 1], [g["h"][k, 0], g["h"][k, 1], np.arctan2(g["R"][k][1, 0], g["R"][k][0, 0])])
    assert np.abs(tr.poses - g["chain_left"]).max() <= 1e-10 and np.abs(td.poses - g["chain_deltas"]).max() <= 1e-10
    assert np.abs(tr.getPoseAtTimes(g["tq"]) - g["chain_interp"]).max() <= 1e-8
    rmse = TP.computePosesRMSE(tr.getPoseAtTimes(g["tq"]), td.getPoseAtTimes(g["tq"]))
    assert abs(rmse - float(g["rmse"])) <= 1e-8
    assert abs(TP.evaluateTrajectoryRMSE(tr, tr)) <= 1e-12
    # too few points: nearest recorded pose, like the reference's except branch
    short = TP.Trajectory(g["ts"][:2], g["chain_left"][:2])
    assert np.array_equal(short.getPoseAtTimes([g["ts"][1] - 1])[0], g["chain_left"][1])


def test_ground_truth_integration(golden, tmp_path):
    g = golden["traj_fixture"]
    path = os.path.join(tmp_path, "radar_odometry.csv")
    with open(path, "w") as f:
        f.write("source_timestamp,destination_timestamp,x,y,z,roll,pitch,yaw,source_radar_timestamp,destination_radar_timestamp\n")
        for (dx, dy, dth), ts in zip(g["gt_csv_cols"], g["gt_csv_ts"]):
            f.write(f"0,0,{float(dx)!r},{float(dy)!r},0,0,0,{float(dth)!r},0,{int(ts)}\n")
    gt = TP.getGroundTruthTrajectory(path)
    assert np.array_equal(gt.timestamps, g["gt_timestamps"])
    assert np.abs(gt.poses - g["gt_poses"]).max() <= 1e-12
    assert gt.getGroundTruthDeltasAtTime(int(g["gt_csv_ts"][3])) == [float(v) for v in g["gt_csv_cols"][3]]
    got = gt.getPoseAtTimes(g["gt_query"])
    assert np.abs(got - g["gt_interp"]).max() <= 1e-7


def test_pose_transform_round_trip():
    p = np.array([[1.0, -2.0, 0.3], [0.0, 0.0, -3.0]])
    T = TP.convertPoseToTransform(p)
    assert T.shape == (2, 3, 3) and np.allclose(TP.convertTransformToPose(T), p)
    assert TP.convertPoseToTransform(p[0]).shape == (3, 3) and np.allclose(TP.convertTransformToPose(T[0]), p[0])
