"""GPU: chained odometry on the device (rf_seq, csrc/k_seq.cu) — the reference's system loop
(RawROAMSystem.run, RawROAMSystem.py:141-300; Mapping.py:37-66,97-125,149-174; getFeatures.py:66-118) for several
sequences in lock step, with features / keyframes / poses resident in HBM, on-device re-detection (response, NMS,
candidate sort, SSC bisection, append) and one CUDA graph per step — against the same loop made of the reference's
third-party calls on the CPU (oracle/ref_system.py).

The oracle side detects with the ORACLE's selection chain (oracle.restate.nms_select -> strongest 65536 -> the
reference-pinned ssc -> fliplr -> vstack + np.unique) on the response map (north_star: feature selection bit-exact
when given the same corner responses); the response itself is pinned to cv2.cornerMinEigenVal in
test_gpu_features.py.  Tolerances (north_star): per-frame pose within 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_M, TOL_RAD = 1e-4, 1e-5


def oracle_detector(fe):
    from oracle import restate as R

    def detect(cart, old):
        frame = fe.frame_from_cart(np.ascontiguousarray(cart, np.float32))
        resp = fe.corner_response(frame)
        frame.close()
        thr = float(np.float32(float(resp.max()) * 0.01))
        cand = R.nms_select(resp, thr)[:65536]
        kp = np.column_stack([cand[:, 0], cand[:, 1], np.ones(len(cand))])
        sel = R.ssc(kp, 200, 0.1, cart.shape[1], cart.shape[0])
        new = np.fliplr(kp[sel][:, :2])                               # getFeatures.py:88
        pts = np.vstack((np.asarray(old, np.float64).reshape(-1, 2), new))
        _, idx = np.unique(pts, axis=0, return_index=True)            # getFeatures.py:110-112
        return np.ascontiguousarray(pts[np.sort(idx)]).astype(np.float32)
    return detect


@pytest.fixture(scope="module")
def sequences():
    from radarslampy_b200 import synthetic as S
    out = []
    for seed, first, distort in ((4321, 0, False), (99, 3, False), (4321, 40, True)):
        world = S.World(seed=seed)
        raw, poses = S.make_sequence(6, res_m=0.0432, world=world, first=first, distort=distort)
        out.append(raw)
    return out


@pytest.mark.parametrize("with_mds", [True, False], ids=["mds", "no-motion-compensation"])
def test_device_chain_matches_cpu_reference_loop(fe, sequences, with_mds):
    from radarslampy_b200 import odometry
    from oracle import ref_system
    T = len(sequences[0])
    init = np.array([[0.0, 0.0, 0.0], [1.5, -2.0, 0.3], [10.0, 5.0, -1.0]])
    got = odometry.run_odometry_device(sequences, init_pose=init, with_mds=with_mds, graph=True, detector_mode=0, fe=fe)
    assert np.all(got["status"] == 0)
    det = oracle_detector(fe)
    for s, raw in enumerate(sequences):
        want = ref_system.run_odometry(raw, det, init_pose=init[s], with_mds=with_mds)
        assert got["n_features_in"][s].tolist() == want["n_features_in"].tolist(), s
        assert got["n_tracked"][s].tolist() == want["n_tracked"].tolist(), s
        assert got["retrack"][s].tolist() == want["retrack"].tolist(), s
        assert int(got["n_keyframes"][s]) == want["n_keyframes"], s
        for k in range(T - 1):
            dth = np.arctan2(got["R"][s, k][1, 0], got["R"][s, k][0, 0]) - np.arctan2(want["R"][k][1, 0], want["R"][k][0, 0])
            assert abs(dth) <= TOL_RAD, (s, k, dth)
            assert np.abs(got["h"][s, k] - want["h"][k]).max() <= TOL_M, (s, k, got["h"][s, k].ravel(), want["h"][k].ravel())
        d = got["poses"][s] - want["poses"]
        assert np.abs(d[:, :2]).max() <= (T - 1) * TOL_M and np.abs(d[:, 2]).max() <= (T - 1) * TOL_RAD


def test_device_chain_features_bit_exact(fe, sequences):
    """the carried feature set (blobCoord) after the last frame, incl. every on-device re-detection, equals the oracle's"""
    from oracle import ref_system
    S, T = len(sequences), len(sequences[0])
    runner = fe.new_sequences(S, S * T)
    for t in range(T):
        runner.upload(t * S, np.stack([sequences[s][t] for s in range(S)]))
    runner.reset(0, 1)
    feats0, counts0 = runner.features()
    for t in range(1, T):
        runner.step(t * S, 1, with_mds=True, graph=False)
    feats, counts = runner.features()
    runner.close()
    det = oracle_detector(fe)
    for s, raw in enumerate(sequences):
        first = det(fe.polar_to_cart(raw=raw[0])[1], np.empty((0, 2), np.float32))
        assert counts0[s] == len(first) and np.array_equal(feats0[s, :counts0[s]], first), s
        want = ref_system.run_odometry(raw, det)
        assert counts[s] == len(want["features"]), s
        assert np.abs(feats[s, :counts[s]] - want["features"]).max() <= 0.02, s     # KLT positions: 0.02 px


def test_graph_and_eager_steps_are_identical(fe, sequences):
    from radarslampy_b200 import odometry
    a = odometry.run_odometry_device(sequences[:2], with_mds=True, graph=True, detector_mode=0, fe=fe)
    b = odometry.run_odometry_device(sequences[:2], with_mds=True, graph=False, detector_mode=0, fe=fe)
    assert a["steps"].tobytes() == b["steps"].tobytes()


def test_device_chain_matches_dropin_loop(fe, sequences):
    """rf_seq against the per-frame drop-in loop (odometry.run_odometry: NumPy in / NumPy out through the C ABI), both
    with the reference's detector (determinant of Hessian, detector_mode 1 = the default of both)"""
    from radarslampy_b200 import odometry
    got = odometry.run_odometry_device(sequences[:1], with_mds=True, fe=fe)
    want = odometry.run_odometry(sequences[0], use_fmt_prior=False)
    assert got["n_tracked"][0].tolist() == want["n_tracked"].tolist()
    assert got["retrack"][0].tolist() == want["retrack"].tolist()
    assert np.abs(got["poses"][0] - want["traj"].poses).max() <= 5e-4
