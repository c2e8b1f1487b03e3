"""GPU parity of the Fourier-Mellin rotation prior through the C ABI:
rf_fmt_log_polar bit-exact, rf_phase_correlate / rf_fmt_rotation within the tolerance below against the
goldens recorded from the unmodified reference (FMT.getRotationUsingFMT on data/tiny) and the CPU oracle."""
import numpy as np
import pytest

from oracle import fmt_restate as F
from oracle import restate as R

pytestmark = pytest.mark.gpu

TOL_SHIFT_PX = 2e-3                              # f32 direct DFT on the GPU vs cv2's f32 FFT / the f64 oracle
TOL_ANGLE_RAD = TOL_SHIFT_PX * 2 * np.pi / 317   # 4e-5 rad: angle = -dy 2 pi / 317 (FMT.py:82)


@pytest.fixture(scope="module")
def polar(golden):
    fr = golden["tiny_frames"]
    return [R.extract_polar(fr[f"raw_{i}"], 2025) for i in range(3)]


def test_log_polar_bit_exact(fe, polar, golden):
    lp = fe.fmt_log_polar(polar[0], downsample=10, clip_px=1012)
    assert lp.shape == (317, 101) and lp.dtype == np.float32
    assert np.array_equal(lp, golden["fmt_fixture"]["logpolar_0"])
    # no clipping / down-sampling = parseData.convertPolarImgToLogPolar of a small polar image
    small = F.resize_cols_linear(polar[1][:, :1012], 101)
    assert np.array_equal(fe.fmt_log_polar(small, downsample=1, clip_px=0), F.polar_to_log_polar(small))


def test_phase_correlate(fe, polar):
    rng = np.random.default_rng(3)
    base = rng.random((96, 80)).astype(np.float32)
    k = np.array([0.25, 0.5, 0.25])            # mild smoothing: keeps the spectrum above the f32 noise floor everywhere
    img = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, base)
    img = np.apply_along_axis(lambda c: np.convolve(c, k, "same"), 0, img).astype(np.float32)
    for shift in ((5, -3), (0, 0), (-11, 7)):
        sh = np.roll(img, shift, axis=(0, 1))
        (dx, dy), resp = fe.phase_correlate(img, sh)
        (ox, oy), oresp = F.phase_correlate(img, sh, F.hanning_window(96, 80))
        assert abs(dx - ox) <= TOL_SHIFT_PX and abs(dy - oy) <= TOL_SHIFT_PX and abs(resp - oresp) <= 1e-4
    # odd sizes are zero-padded to the optimal DFT size (97 -> 100, 61 -> 64)
    a, b = img[:, :61].copy(), np.roll(img, (2, 1), axis=(0, 1))[:, :61].copy()
    a, b = np.vstack([a, a[:1]]), np.vstack([b, b[:1]])
    (dx, dy), resp = fe.phase_correlate(a, b)
    (ox, oy), oresp = F.phase_correlate(a, b, F.hanning_window(97, 61))
    assert abs(dx - ox) <= TOL_SHIFT_PX and abs(dy - oy) <= TOL_SHIFT_PX


def test_fmt_rotation_vs_reference_goldens(fe, polar, golden):
    g = golden["fmt_fixture"]
    ang, sc, resp, sh = fe.fmt_rotation(np.stack(polar), g["pairs"], downsample=10, clip_px=1012)
    assert np.abs(sh - g["fmt_shift"]).max() <= TOL_SHIFT_PX
    assert np.abs(ang - g["fmt_angle"]).max() <= TOL_ANGLE_RAD
    assert np.abs(sc - g["fmt_scale"]).max() <= 1e-4 and np.abs(resp - g["fmt_response"]).max() <= 1e-4
    rolled = np.stack([polar[0]] + [np.roll(polar[0], int(k), axis=0) for k in g["roll_k"]])
    pairs = [[0, i + 1] for i in range(len(g["roll_k"]))]
    ang, _, _, _ = fe.fmt_rotation(rolled, pairs, downsample=10, clip_px=1012)
    assert np.abs(ang - g["roll_angle"]).max() <= TOL_ANGLE_RAD
    # identical images: zero shift, response of the centred peak
    ang, sc, resp, sh = fe.fmt_rotation(np.stack(polar[:1]), [[0, 0]], downsample=10, clip_px=1012)
    assert abs(ang[0]) <= 1e-6 and abs(sc[0] - 1) <= 1e-6 and np.abs(sh).max() <= 1e-4


def test_fmt_dropin_module(polar, golden):
    from radarslampy_b200 import FMT, parseData
    g = golden["fmt_fixture"]
    a, s, r = FMT.getRotationUsingFMT(polar[0], polar[1])
    assert isinstance(a, float) and abs(a - g["fmt_angle"][0]) <= TOL_ANGLE_RAD and abs(s - g["fmt_scale"][0]) <= 1e-4
    assert abs(r - g["fmt_response"][0]) <= 1e-4
    small = F.resize_cols_linear(polar[0][:, :1012], 101)
    lp = parseData.convertPolarImgToLogPolar(small)
    assert np.array_equal(lp, g["logpolar_0"])
    (dx, dy), _ = FMT.getTranslationUsingPhaseCorrelation(lp, np.roll(lp, 4, axis=0))
    assert abs(dy + 4) < 0.3 or abs(dy - 4) < 0.3


def test_batch_fmt_from_resident_scans(golden):
    """rf_batch_fmt reads the u8 scans a batch already holds; same numbers as the f32-polar entry point."""
    from radarslampy_b200 import _ffi
    fr, g = golden["tiny_frames"], golden["fmt_fixture"]
    cfg = _ffi.default_config()
    cfg.max_pairs, cfg.max_frames, cfg.max_features = 4, 3, 64
    eng = _ffi.RadarFE(cfg, device=0)
    try:
        b = eng.new_batch()
        raw = np.stack([fr[f"raw_{i}"] for i in range(3)])
        pairs = g["pairs"]
        feats = np.zeros((len(pairs), 64, 2), np.float32)
        b.upload(raw, pairs, feats, np.zeros(len(pairs), np.int32))
        ang, sc, resp, sh = b.fmt_rotation(downsample=10, clip_px=1012)
        assert np.abs(sh - g["fmt_shift"]).max() <= TOL_SHIFT_PX and np.abs(ang - g["fmt_angle"]).max() <= TOL_ANGLE_RAD
        assert np.abs(resp - g["fmt_response"]).max() <= 1e-4
        b.close()
    finally:
        eng.close()


def test_bad_arguments(fe, polar):
    with pytest.raises((ValueError, RuntimeError)):
        fe.fmt_rotation(np.stack(polar[:2]), [[0, 2]])
    with pytest.raises((ValueError, RuntimeError)):
        fe.phase_correlate(np.zeros((4, 4), np.float32), np.zeros((4, 5), np.float32))


def test_fmt_other_geometry(fe):
    """200 azimuths x 400 bins, down-sampling 5, no clipping: log-polar 251 x 80, DFT 256 x 80 (column split 16 x 16)."""
    rng = np.random.default_rng(8)
    base = rng.random((200, 400)).astype(np.float32)
    k = np.array([0.25, 0.5, 0.25])
    img = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, base).astype(np.float32)
    rolled = np.roll(img, 7, axis=0)
    lp = [F.polar_to_log_polar(F.resize_cols_linear(p, 80)) for p in (img, rolled)]
    assert lp[0].shape == (251, 80)
    assert np.array_equal(fe.fmt_log_polar(img, downsample=5, clip_px=0), lp[0])
    (ox, oy), oresp = F.phase_correlate(lp[0], lp[1], F.hanning_window(251, 80))
    ang, sc, resp, sh = fe.fmt_rotation([img, rolled], [[0, 1]], downsample=5, clip_px=0)
    assert abs(sh[0, 0] - ox) <= TOL_SHIFT_PX and abs(sh[0, 1] - oy) <= TOL_SHIFT_PX and abs(resp[0] - oresp) <= 1e-4
    want = (-oy * 2 * np.pi / 251 + np.pi) % (2 * np.pi) - np.pi
    assert abs(ang[0] - want) <= TOL_ANGLE_RAD
