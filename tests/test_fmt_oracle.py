"""CPU: the Fourier-Mellin restatement (oracle/fmt_restate.py) against the goldens recorded from the
unmodified reference (tests/golden/fmt_fixture.npz, oracle/gen_golden_fmt.py) and against live cv2."""
import numpy as np
import pytest

from oracle import fmt_restate as F
from oracle import restate as R

TOL_SHIFT_PX = 1e-5      # oracle (f64 DFT) vs cv2.phaseCorrelate (f32 FFT); measured 2.5e-7
TOL_ANGLE_RAD = 1e-6


@pytest.fixture(scope="module")
def polar(golden):
    fr = golden["tiny_frames"]
    return [R.extract_polar(fr[f"raw_{i}"], 2025) for i in range(3)]


def test_fmt_oracle_vs_reference_goldens(polar, golden):
    g = golden["fmt_fixture"]
    for (i, j), ang, sc, resp in zip(g["pairs"], g["fmt_angle"], g["fmt_scale"], g["fmt_response"]):
        a, s, r = F.rotation_fmt(polar[i], polar[j])
        assert abs(a - ang) <= TOL_ANGLE_RAD and abs(s - sc) <= 1e-6 and abs(r - resp) <= 1e-5
    for k, ang in zip(g["roll_k"], g["roll_angle"]):
        a, _, _ = F.rotation_fmt(polar[0], np.roll(polar[0], int(k), axis=0))
        assert abs(a - ang) <= TOL_ANGLE_RAD


def test_fmt_log_polar_bit_exact(polar, golden):
    g = golden["fmt_fixture"]
    lp = F.polar_to_log_polar(F.resize_cols_linear(polar[0][:, :1012], 101))
    assert lp.shape == (317, 101) and np.array_equal(lp, g["logpolar_0"])


def test_fmt_stages_vs_live_cv2(polar, golden):
    cv2 = pytest.importorskip("cv2")
    a = polar[0][:, :1012]
    assert np.array_equal(F.resize_cols_linear(a, 101), cv2.resize(a, (101, 400)))          # IPP linear resize
    assert np.array_equal(F.hanning_window(317, 101), cv2.createHanningWindow((101, 317), cv2.CV_32F))
    assert F.optimal_dft_size(317) == cv2.getOptimalDFTSize(317) == 320 and F.optimal_dft_size(101) == cv2.getOptimalDFTSize(101) == 108
    lp = [F.polar_to_log_polar(F.resize_cols_linear(p[:, :1012], 101)) for p in polar[:2]]
    w = cv2.createHanningWindow((101, 317), cv2.CV_32F)
    (dx, dy), resp = cv2.phaseCorrelate(lp[0], lp[1], w)
    (ox, oy), oresp = F.phase_correlate(lp[0], lp[1], w)
    assert abs(dx - ox) <= TOL_SHIFT_PX and abs(dy - oy) <= TOL_SHIFT_PX and abs(resp - oresp) <= 1e-5
    # a pure circular shift is recovered to a fraction of a pixel.  (Broadband image: where the spectrum falls
    # below cv2's f32 FFT noise floor the normalised cross-power amplifies that noise, and no restatement that is
    # not a bit-exact copy of cv2's FFT schedule can follow it — sigma = 2 blur: 7e-4 px; here: < 1e-6 px.)
    rng = np.random.default_rng(3)
    img = cv2.GaussianBlur(rng.random((96, 80)).astype(np.float32), (0, 0), 0.6)
    sh = np.roll(img, (5, -3), axis=(0, 1))
    (ox, oy), _ = F.phase_correlate(img, sh, F.hanning_window(96, 80))
    (cx, cy), _ = cv2.phaseCorrelate(img, sh, cv2.createHanningWindow((80, 96), cv2.CV_32F))
    assert abs(ox - cx) <= TOL_SHIFT_PX and abs(oy - cy) <= TOL_SHIFT_PX and abs(ox + 3) < 0.2 and abs(oy - 5) < 0.2
