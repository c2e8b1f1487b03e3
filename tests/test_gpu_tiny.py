"""GPU: the reference's own configs[0] — RawROAMSystem.py on data/tiny, all 11 scans / 10 pairs (RawROAMSystem.py:162-298)
— against goldens recorded from the UNMODIFIED reference driver (oracle/gen_golden_tiny_run.py; its detector call is
bound to the restated blob_doh, everything else is the reference's code and its third-party calls).
  * every pair through the pair-batch path (rf_track_batch) with the golden features: KLT status, tracked positions,
    clique mask and Kabsch transform of every pair;
  * the whole chained run on the device (rf_seq, detector_mode 1) and through the per-frame drop-in loop: identical
    feature sets at every (re-)detection, identical tracked / inlier counts and re-detection frames, poses within the
    north_star tolerance per chained frame."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SCANS = os.path.join(HERE, "golden", "tiny_scans")
TOL_PX, TOL_M, TOL_RAD = 0.02, 1e-4, 1e-5


@pytest.fixture(scope="module")
def scans(golden):
    from radarslampy_b200 import parseData
    paths = parseData.getRadarImgPaths(os.path.join(SCANS, "radar"), os.path.join(SCANS, "radar.timestamps"))
    assert len(paths) == 11 and paths == sorted(glob.glob(os.path.join(SCANS, "radar", "*.png")))
    raw = parseData.readRadarScans(paths)
    assert raw.shape == (11, 400, 3779) and raw.dtype == np.uint8
    fr = golden["tiny_frames"]
    for i in range(3):                                   # the decoder against the scans cv2.imread gave the generator
        assert np.array_equal(raw[i], fr[f"raw_{i}"])
    return raw


def test_every_pair_through_the_batch_path(fe, scans, golden):
    from radarslampy_b200 import _ffi
    g = golden["tiny_run_doh"]
    P = int(g["n_pairs"])
    assert P == 10
    K = fe.cfg.max_features
    cfg = _ffi.default_config()
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 11, 10, 320
    eng = _ffi.RadarFE(cfg, device=0)
    K = cfg.max_features
    feats = np.zeros((P, K, 2), np.float32)
    counts = np.zeros(P, np.int32)
    for i in range(P):
        f = g[f"feat_in_{i}"]
        counts[i] = len(f)
        feats[i, :len(f)] = f
    pair_idx = np.stack([np.arange(P), np.arange(1, P + 1)], 1).astype(np.int32)
    b = eng.new_batch()
    res, nxt, corr = b.track(scans, pair_idx, feats, counts, prev_pose=g["poses"][:P], with_mds=False)
    st, err = b.klt_status()
    for i in range(P):
        n = counts[i]
        want_st = g[f"klt_status_{i}"].ravel()
        assert np.array_equal(st[i, :n], want_st), i
        good = want_st.astype(bool)
        assert np.abs(nxt[i, :n][good] - g[f"klt_good_new_{i}"]).max() <= TOL_PX, i
        mask = g[f"rej_mask_{i}"]
        want_corr = want_st.copy()
        want_corr[np.flatnonzero(good)] &= mask.astype(want_corr.dtype)      # Tracker.py:103-104
        assert np.array_equal(corr[i, :n], want_corr), i
        assert res["n_good"][i] == good.sum() and res["n_inliers"][i] == mask.sum() and res["status"][i] == 0
        R, h = res["R"][i].reshape(2, 2), res["h"][i]
        dth = np.arctan2(R[1, 0], R[0, 0]) - np.arctan2(g[f"svd_R_{i}"][1, 0], g[f"svd_R_{i}"][0, 0])
        assert abs(dth) <= TOL_RAD, (i, dth)
        assert np.abs(h - g[f"svd_h_{i}"].ravel() * 0.0864).max() <= TOL_M, (i, h, g[f"svd_h_{i}"].ravel() * 0.0864)
    b.close()
    eng.close()


def _check_chain(got_poses, n_in, n_tracked, retrack, g, label):
    P = int(g["n_pairs"])
    assert list(n_in) == [len(g[f"feat_in_{i}"]) for i in range(P)], label
    assert list(n_tracked) == [int(g[f"rej_mask_{i}"].sum()) for i in range(P)], label
    assert [bool(r) for r in retrack] == [f"retrack_{i}" in g.files for i in range(P)], label
    d = got_poses - g["poses"]
    dth = np.abs((d[:, 2] + np.pi) % (2 * np.pi) - np.pi)
    k = np.maximum(np.arange(P + 1), 1)
    assert np.all(np.abs(d[:, :2]).max(axis=1) <= k * TOL_M), (label, np.abs(d[:, :2]).max(axis=1))
    assert np.all(dth <= k * TOL_RAD), (label, dth)


def test_whole_sequence_chained_on_the_device(fe, scans, golden):
    from radarslampy_b200 import odometry
    g = golden["tiny_run_doh"]
    assert bool(g["argsort_identity"])
    runner = fe.new_sequences(1, 11, detector_mode=1)
    runner.upload(0, scans)
    runner.reset(0, 1)
    f0, c0 = runner.features()
    assert c0[0] == len(g["first_features"]) and np.array_equal(f0[0, :c0[0]], g["first_features"])   # detection bit-exact
    runner.close()
    got = odometry.run_odometry_device([scans], init_pose=g["poses"][:1], with_mds=True, detector_mode=1, fe=fe)   # the driver starts from the GT pose (RawROAMSystem.py:122-128)
    assert np.all(got["status"] == 0)
    _check_chain(got["poses"][0], got["n_features_in"][0], got["n_tracked"][0], got["retrack"][0], g, "rf_seq")


def test_whole_sequence_through_the_dropin_loop(scans, golden):
    """RawROAMSystem.run's call sequence over the drop-in modules (NumPy in / NumPy out, FMT prior included)."""
    from radarslampy_b200 import odometry
    g = golden["tiny_run_doh"]
    got = odometry.run_odometry(scans, init_pose=g["poses"][0])
    _check_chain(got["traj"].poses, got["n_features_in"], got["n_tracked"], got["retrack"], g, "drop-in loop")
