"""GPU parity of the fused batch path (rf_track_batch: Tracker.track + getTransform + MDS for
many independent pairs) against the reference-generated goldens and against the reference's
own library calls (oracle/ref_pipeline.py) on seeded synthetic Oxford-shaped scans.
Tolerances (north_star): identical status / clique, tracks <= 0.02 px, pose <= 1e-4 m, 1e-5 rad."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_PX, TOL_M, TOL_RAD = 0.02, 1e-4, 1e-5
RES = 0.0864


def _engine(range_bins=2025, res=0.0432, max_pairs=16, max_frames=16, max_features=256, f32=0):
    from radarslampy_b200 import _ffi
    cfg = _ffi.default_config()
    cfg.range_bins = range_bins
    cfg.cart_res_m = 2 * res
    cfg.dist_thr_px = 0.5 / (2 * res)
    cfg.max_pairs, cfg.max_frames, cfg.max_features = max_pairs, max_frames, max_features
    cfg.write_cart_f32 = f32
    return _ffi.RadarFE(cfg)


def _pad(feat_list, kmax):
    feats = np.zeros((len(feat_list), kmax, 2), np.float32)
    counts = np.zeros(len(feat_list), np.int32)
    for i, f in enumerate(feat_list):
        counts[i] = len(f)
        feats[i, :len(f)] = f
    return feats, counts


def _theta(R):
    R = np.asarray(R).reshape(2, 2)
    return np.arctan2(R[1, 0], R[0, 0])


def test_batch_matches_reference_goldens(golden):
    fr, st = golden["tiny_frames"], golden["tiny_stages"]
    fe = _engine(f32=1)
    try:
        b = fe.new_batch()
        raw = np.stack([fr[f"raw_{i}"] for i in range(3)])
        feats, counts = _pad([st["feat_in_0"], st["feat_in_1"]], 256)
        T0 = st["mds_Twj0_0"]
        prev_pose = np.array([[T0[0, 2], T0[1, 2], np.arctan2(T0[1, 0], T0[0, 0])], [0, 0, 0]])
        res, nxt, corr = b.track(raw, [[0, 1], [1, 2]], feats, counts, prev_pose=prev_pose, with_mds=True)
        klt_st, err = b.klt_status()
        for p in range(2):
            K = counts[p]
            ref_status = st[f"klt_status_{p}"].ravel()
            assert np.array_equal(klt_st[p, :K], ref_status)
            good = ref_status.astype(bool)
            assert np.abs(nxt[p, :K][good] - st[f"klt_good_new_{p}"]).max() <= TOL_PX
            want = ref_status.copy()
            want[good] &= st[f"rej_mask_{p}"].astype(np.uint8)
            assert np.array_equal(corr[p, :K], want)                       # Tracker.py:102-104
            assert not corr[p, K:].any()
            r = res[p]
            assert (r["n_features"], r["n_good"], r["n_inliers"]) == (K, good.sum(), st[f"rej_mask_{p}"].sum())
            assert r["status"] == 0
            assert np.abs(r["h"] - st[f"svd_h_{p}"].ravel() * RES).max() <= TOL_M
            assert abs(_theta(r["R"]) - _theta(st[f"svd_R_{p}"])) <= TOL_RAD
        # pair 0: previous frame == keyframe with zero velocity -> the fused MDS problem is the reference's
        x, ref = res[0]["mds_x"], st["mds_x_0"]
        assert np.abs(x[3:5] - ref[3:5]).max() <= TOL_M and abs(x[5] - ref[5]) <= TOL_RAD
        # frames converted inside the batch are bit-exact too
        import hashlib
        assert hashlib.sha256(b.frame(1, 0).tobytes()).hexdigest() == str(fr["cart_sha256_1"])
        assert hashlib.sha256(b.frame(2, 1).tobytes()).hexdigest() == str(fr["u8_sha256_2"])
        b.close()
    finally:
        fe.close()


def test_batch_synthetic_sequence_vs_reference_libraries():
    """BASELINE config 2/3 shape (0.0438 m/bin -> 1997 bins, 1996^2 image): every pair of a seeded
    synthetic sequence against the reference's own cv2/scipy/networkx/numpy calls."""
    from oracle import ref_pipeline as P
    from radarslampy_b200 import synthetic as S
    res_m, n = 0.0438, 6
    rb = int(87.5 / res_m)
    world = S.World()
    raw, poses = S.make_sequence(n, res_m=res_m, world=world)
    pair_idx, feats, counts = S.sequence_pairs(n, world, poses, res_m, rb, k=200, max_features=256)
    fe = _engine(range_bins=rb, res=res_m)
    try:
        b = fe.new_batch()
        res, nxt, corr = b.track(raw, pair_idx, feats, counts, prev_pose=poses[:-1], with_mds=True)
        klt_st, _ = b.klt_status()
        assert np.array_equal(b.frame(0, 1), (P.polar_to_cart(P.extract_polar(raw[0], res_m)) * 255).astype(np.uint8))
        for p in range(n - 1):
            K = counts[p]
            o = P.track_pair(raw[p], raw[p + 1], feats[p, :K], prev_pose=poses[p], with_mds=True, range_res_m=res_m)
            assert np.array_equal(corr[p, :K], o["corr_status"]), f"pair {p}: inlier set differs"
            assert res[p]["n_good"] == o["n_good"] and res[p]["n_inliers"] == o["n_inliers"]
            g = klt_st[p, :K].astype(bool)
            assert np.abs(nxt[p, :K][g] - o["good_new"]).max() <= TOL_PX
            assert np.abs(res[p]["h"] - o["h"]).max() <= TOL_M
            assert abs(_theta(res[p]["R"]) - _theta(o["R"])) <= TOL_RAD
            assert np.abs(res[p]["mds_x"][3:5] - o["mds_x"][3:5]).max() <= TOL_M
            assert abs(res[p]["mds_x"][5] - o["mds_x"][5]) <= TOL_RAD
            # and the estimate is physically right: 2.5 m / 0.025 rad per frame
            assert abs(_theta(res[p]["R"]) - 0.025) < 2e-3
        b.close()
    finally:
        fe.close()


def test_batch_edge_cases():
    from radarslampy_b200 import synthetic as S
    fe = _engine(max_pairs=4, max_frames=4, max_features=64)
    try:
        b = fe.new_batch()
        rng = np.random.default_rng(1)
        raw = rng.integers(0, 256, (2, 400, 3779), dtype=np.uint8)
        feats = np.zeros((3, 64, 2), np.float32)
        feats[1, :5] = [[-100, -100], [5000, 5000], [10, 10], [1000, 1000], [2023, 2023]]
        feats[2, :64] = rng.uniform(0, 2024, (64, 2))
        counts = np.array([0, 5, 64], np.int32)
        res, nxt, corr = b.track(raw, [[0, 1], [0, 0], [1, 1]], feats, counts, with_mds=True)
        assert res[0]["n_good"] == 0 and res[0]["n_inliers"] == 0
        assert np.array_equal(np.asarray(res[0]["R"]), [1, 0, 0, 1]) and not np.asarray(res[0]["h"]).any()
        assert res[1]["n_good"] <= 3                       # the two far-outside points are lost (status 0)
        # identical frames: every tracked point stays put -> identity transform, all good points inliers
        assert res[2]["n_inliers"] == res[2]["n_good"] > 0
        assert abs(_theta(res[2]["R"])) < 1e-6 and np.abs(res[2]["h"]).max() < 1e-3
        assert np.all(np.isfinite(res["mds_x"]))
        # capacity and argument errors are reported, not crashed on
        with pytest.raises(ValueError):
            b.track(raw, [[0, 2]], feats[:1], counts[:1])
        with pytest.raises(ValueError):
            b.track(np.zeros((5, 400, 3779), np.uint8), [[0, 1]], feats[:1], counts[:1])
        with pytest.raises(ValueError):
            b.track(raw, [[0, 1]], feats[:1], np.array([65], np.int32))
        # empty batch
        res, _, _ = b.track(raw, np.zeros((0, 2), np.int32), np.zeros((0, 64, 2), np.float32), np.zeros(0, np.int32))
        assert len(res) == 0
        b.close()
    finally:
        fe.close()


def test_batch_profiling_and_launch_count():
    from radarslampy_b200 import synthetic as S
    fe = _engine(max_pairs=4, max_frames=4, max_features=64)
    try:
        b = fe.new_batch()
        rng = np.random.default_rng(2)
        raw = rng.integers(0, 256, (2, 400, 3779), dtype=np.uint8)
        feats = rng.uniform(100, 1900, (1, 64, 2)).astype(np.float32)
        b.upload(raw, [[0, 1]], feats, [64])
        b.set_profiling(True)
        n0 = fe.launch_count()
        for _ in range(3):
            b.run_async(with_mds=True)
        fe.sync()
        assert fe.launch_count() - n0 == 3 * 14      # interleave, scan->L0+L1, 2 pyrDown levels, klt, compact, adjacency, maxclique, viable set, adjacency slot orders, clique, kabsch, mds, finish
        ms, runs = b.stage_times()
        assert runs == 3 and all(v >= 0 for v in ms.values()) and ms["polar2cart"] > 0 and ms["klt"] > 0
        b.close()
    finally:
        fe.close()


@pytest.mark.parametrize("range_bins,res,n_frames", [(2025, 0.0432, 3), (1997, 0.0438, 9), (2025, 0.0432, 18)])
def test_fused_image_path_all_levels_bit_exact(range_bins, res, n_frames):
    """The batch image path (frame interleave -> scan -> level 0 + 1 in one kernel -> warp-tile pyrDown) must give
    the u8 image and EVERY pyramid level bit-identical to the oracle's warpPolar / truncation / pyrDown chain, for
    frame counts that are not a multiple of the 16-frame interleave group and for both scan geometries.  The kernel
    takes an integer shortcut wherever the fixed-point bilinear sum is not a multiple of 1024 and cv2's f32 chain
    elsewhere: the saturated / constant / real scans below are the cases where the second path dominates."""
    from oracle import restate as R
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, (n_frames, 400, 3779), dtype=np.uint8)
    raw[1, :, 11:] = 255                                    # saturated scan: every blend must give exactly 255
    raw[2, ::2, 11:] = 0                                    # azimuth comb: wrap-around rows and borders
    if n_frames > 4:
        raw[4, :, 11:] = rng.integers(0, 256, (400, 1), dtype=np.uint8)      # constant along range: equal tap pairs
        raw[5, :, 11:] = rng.integers(0, 4, (400, 3768), dtype=np.uint8)     # tiny values: many zero sums
    if n_frames > 17:
        tiny = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_frames.npz"))
        raw[3], raw[16], raw[17] = tiny["raw_0"], tiny["raw_1"], 77          # real Oxford scans; a constant scan
    fe = _engine(range_bins=range_bins, res=res, max_pairs=2, max_frames=n_frames, max_features=64, f32=0)
    try:
        b = fe.new_batch()
        b.upload(raw, np.zeros((0, 2), np.int32), np.zeros((0, 64, 2), np.float32), np.zeros(0, np.int32))
        b.run_async()
        fe.sync()
        for f in sorted({0, 1, 2, n_frames - 1} | ({3, 4, 5} if n_frames > 5 else set()) | ({16} if n_frames > 16 else set())):
            want = R.to_u8(R.warp_polar(R.extract_polar(raw[f], range_bins)))
            for lvl in range(4):
                got = b.frame(f, 1 + lvl)
                assert got.shape == want.shape, (f, lvl)
                assert np.array_equal(got, want), f"frame {f} level {lvl} differs"
                want = R.pyr_down(want)
        b.close()
    finally:
        fe.close()


def test_two_batches_pipelined_equal_serial():
    """Two batches alternating on one handle (copy / image+KLT / tail streams overlapping) must return exactly what
    each returns on its own, run after run (no cross-batch races, deterministic reductions)."""
    from radarslampy_b200 import synthetic as S
    res_m, n = 0.0438, 5
    rb = int(87.5 / res_m)
    seqs = []
    for seed in (1, 2):
        world = S.World(seed=seed)
        raw, poses = S.make_sequence(n, res_m=res_m, world=world)
        seqs.append((raw, poses) + tuple(S.sequence_pairs(n, world, poses, res_m, rb, k=150, max_features=192)))
    fe = _engine(range_bins=rb, res=res_m, max_pairs=n - 1, max_frames=n, max_features=192)
    try:
        serial = []
        for raw, poses, pi, feats, counts in seqs:
            b = fe.new_batch()
            r, nx, c = b.track(raw, pi, feats, counts, prev_pose=poses[:-1], with_mds=True)
            serial.append((r.copy(), nx.copy(), c.copy()))
            b.close()
        bs = [fe.new_batch(), fe.new_batch()]
        outs = [b.alloc_outputs(pinned=True) for b in bs]
        for it in range(6):
            k = it % 2
            raw, poses, pi, feats, counts = seqs[k]
            bs[k].upload(raw, pi, feats, counts, prev_pose=poses[:-1], sync=False)
            bs[k].run_async(with_mds=True)
            bs[k].download(outs[k], sync=False)
            if it:
                j = 1 - k
                bs[j].wait()
                for name in ("R", "h", "mds_x", "n_good", "n_inliers", "clique_nodes", "status"):
                    assert np.array_equal(outs[j][0][:n - 1][name], serial[j][0][name]), (it, name)
                assert np.array_equal(outs[j][1][:n - 1], serial[j][1]) and np.array_equal(outs[j][2][:n - 1], serial[j][2])
        fe.sync()
        for b in bs:
            b.close()
    finally:
        fe.close()


def test_track_pair_equals_batch_path(golden):
    """rf_track_pair (two separately held scans, handle-owned two-frame batch) returns what rf_track_batch returns."""
    fr, st = golden["tiny_frames"], golden["tiny_stages"]
    fe = _engine(max_pairs=2, max_frames=3, max_features=256)
    try:
        pts = st["feat_in_0"]
        K = pts.shape[0]
        raw = np.stack([fr["raw_0"], fr["raw_1"]])
        feats = np.zeros((1, 256, 2), np.float32)
        feats[0, :K] = pts
        b = fe.new_batch()
        res_b, nxt_b, corr_b = b.track(raw, [[0, 1]], feats, [K], with_mds=True)
        for _ in range(2):                                   # the handle-owned batch is reused call after call
            res, nxt, corr = fe.track_pair(fr["raw_0"], fr["raw_1"], pts, with_mds=True)
            for name in ("R", "h", "mds_x", "n_features", "n_good", "n_inliers", "status", "clique_nodes"):
                assert np.array_equal(res[name], res_b[0][name]), name
            assert np.array_equal(nxt, nxt_b[0, :K]) and np.array_equal(corr, corr_b[0, :K])
        want = st["klt_status_0"].copy()
        want[want.flatten().astype(bool)] &= st["rej_mask_0"][:, None].astype(np.uint8)
        assert np.array_equal(corr, want.ravel())
        res0, nxt0, corr0 = fe.track_pair(fr["raw_0"], fr["raw_1"], np.zeros((0, 2), np.float32))
        assert int(res0["n_features"]) == 0 and nxt0.shape == (0, 2)
        with pytest.raises((ValueError, RuntimeError)):
            fe.track_pair(fr["raw_0"], fr["raw_1"], np.zeros((300, 2), np.float32))
        b.close()
    finally:
        fe.close()


@pytest.mark.parametrize("A,bins,n_frames", [(200, 501, 5), (360, 1205, 17)])
def test_other_scan_geometries_bit_exact(A, bins, n_frames):
    """Nothing is specialised to 400 x 2025: other azimuth counts / range clips (small, odd pyramid sizes, tiles that are
    all border) give the oracle's images at every level through the batch path and the single-frame path."""
    from oracle import restate as R
    from radarslampy_b200 import _ffi
    rng = np.random.default_rng(5)
    cfg = _ffi.default_config()
    cfg.azimuths, cfg.range_bins, cfg.raw_width = A, bins, 11 + bins + 37
    cfg.max_frames, cfg.max_pairs, cfg.max_features, cfg.write_cart_f32 = n_frames, 2, 64, 0
    raw = rng.integers(0, 256, (n_frames, A, cfg.raw_width), dtype=np.uint8)
    fe = _ffi.RadarFE(cfg)
    try:
        n = 2 * (bins // 2)
        assert fe.n == n
        b = fe.new_batch()
        b.upload(raw, np.zeros((0, 2), np.int32), np.zeros((0, 64, 2), np.float32), np.zeros(0, np.int32))
        b.run_async()
        fe.sync()
        for f in (0, n_frames - 1):
            polar = (raw[f][:, 11:11 + bins].astype(np.float32) / np.float32(255))
            want = R.to_u8(R.warp_polar(polar))
            frame, cart = fe.polar_to_cart(raw=raw[f])
            assert np.array_equal(cart, R.warp_polar(polar))
            for lvl in range(4):
                got = b.frame(f, 1 + lvl)
                assert got.shape == want.shape and np.array_equal(got, want), (f, lvl)
                assert np.array_equal(frame.download(1 + lvl), want), (f, lvl, "single-frame path")
                want = R.pyr_down(want)
            frame.close()
        b.close()
    finally:
        fe.close()


def test_plain_load_pyramid_fallback_bit_exact():
    """RADARFE_NO_TMA=1 selects k_pyr_down_w (aligned-word loads) for the upper levels instead of the TMA-staged
    kernel; the switch is read once per process, so the same bit-exact level tests run in a child process."""
    import subprocess
    import sys
    env = dict(os.environ, RADARFE_NO_TMA="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_batch.py"), "-m", "gpu", "-q", "-x", "-k",
                        "fused_image_path_all_levels_bit_exact or other_scan_geometries_bit_exact"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
