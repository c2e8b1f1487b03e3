"""GPU: BASELINE.json configs[4], the dense stress case — 2000 x 2000 Cartesian grid, 10 000 ANMS features per
frame from >= 40 k detector candidates, 4-level pyramidal LK on all of them — against the CPU oracle.
Bars: SSC selection bit-exact given the same candidates; KLT status identical, positions <= 0.02 px."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_PX = 0.02
N = 2000


def _scene(seed, shift=(0.0, 0.0)):
    from radarslampy_b200.synthetic import dense_scene
    return dense_scene(seed, shift, n=N)


@pytest.fixture(scope="module")
def stress():
    from radarslampy_b200 import _ffi
    cfg = _ffi.default_config()
    cfg.range_bins, cfg.downsample = N + 1, 2            # polar 400 x 2001 used bins -> R = 1000, 2000^2 grid
    cfg.raw_width = cfg.meta_bytes + cfg.range_bins
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 2, 1, 2048
    eng = _ffi.RadarFE(cfg, device=0)
    assert eng.n == N
    a, b = _scene(5), _scene(5, shift=(1.7, -0.9))
    yield eng, a, b
    eng.close()


def test_dense_stress_config(stress):
    from oracle import restate as R
    eng, a, b = stress
    fa, fb = eng.frame_from_cart(a), eng.frame_from_cart(b)
    # a10 + a9: >= 40 k candidates -> 10 000 spatially spread features, selection identical to the oracle's ssc
    cand, n = eng.detect(fa, -0.01, cap=200000)
    assert n == len(cand) and n >= 40000
    sel = eng.ssc(cand, 10000, 0.1, N, N)
    assert 9000 <= len(sel) <= 11000
    assert np.array_equal(sel, R.ssc(cand, 10000, 0.1, N, N))
    pts = np.ascontiguousarray(cand[sel][:, [1, 0]], np.float32)      # (x, y)
    # a4 / a5 on all of them, 4 levels (2000 / 1000 / 500 / 250)
    nxt, st, err = eng.klt(fa, fb, pts, apply_err_gate=False)
    u8a, u8b = fa.download(1), fb.download(1)
    assert np.array_equal(u8a, R.to_u8(a))
    o_nxt, o_st, o_err = R.pyr_lk(u8a, u8b, pts)
    assert np.array_equal(st.ravel(), o_st)
    g = o_st.astype(bool)
    assert g.mean() > 0.9
    assert np.abs(nxt[g] - o_nxt[g]).max() <= TOL_PX
    flow = np.median(nxt[g] - pts[g], axis=0)
    assert abs(flow[0] - 1.7) < 0.1 and abs(flow[1] + 0.9) < 0.1
