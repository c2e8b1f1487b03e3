"""GPU: rf_chain_poses (SE(2) prefix product, csrc/k_traj.cu) against the sequential chain the reference runs
(trajectoryPlotting.py:27-60) and its goldens."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def seq_chain(R, h, start, left):
    c, s = np.cos(start[2]), np.sin(start[2])
    T = np.array([[c, -s, start[0]], [s, c, start[1]], [0, 0, 1]])
    out = [start]
    for k in range(len(R)):
        A = np.eye(3); A[:2, :2] = R[k]; A[:2, 2] = h[k]
        T = A @ T if left else T @ A
        out.append([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])])
    return np.array(out, np.float64)


@pytest.mark.parametrize("P", [0, 1, 2, 255, 256, 257, 4096])
@pytest.mark.parametrize("left", [True, False])
def test_chain_matches_sequential(fe, P, left):
    rng = np.random.default_rng(P + 17 * left)
    th = rng.normal(0.01, 0.02, P)
    R = np.stack([[[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]] for t in th]) if P else np.zeros((0, 2, 2))
    h = rng.normal([1.0, 0.0], 0.3, (P, 2))
    start = np.array([3.0, -1.0, 0.7])
    got = fe.chain_poses(R, h, start_pose=start, left_multiply=left)
    want = seq_chain(R, h, start, left)
    assert got.shape == (P + 1, 3)
    d = got - want
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() <= 1e-8 * max(1.0, np.abs(want[:, :2]).max())


def test_chain_matches_reference_goldens(fe, golden):
    from radarslampy_b200 import trajectoryPlotting as TP
    g = golden["traj_fixture"]
    got = fe.chain_poses(g["R"], g["h"], start_pose=g["start"], left_multiply=True)
    assert np.abs(got - g["chain_left"]).max() <= 1e-9
    got = fe.chain_poses(g["R"], g["h"], start_pose=g["start"], left_multiply=False)
    assert np.abs(got - g["chain_deltas"]).max() <= 1e-9
    tr = TP.Trajectory.from_relative_transforms(g["ts"], g["R"], g["h"], start_pose=g["start"])
    assert np.abs(tr.poses - g["chain_left"]).max() <= 1e-9
    assert np.abs(tr.getPoseAtTimes(g["tq"]) - g["chain_interp"]).max() <= 1e-7
