"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Fourier-Mellin rotation prior.

Follows, stage by stage, what the reference's FMT.getRotationUsingFMT (FMT.py:36-90) makes
OpenCV 4.13 compute; every function cites the call site it restates.  The algorithms live in
OpenCV (not under /root/reference): cv::resize INTER_LINEAR, cv::warpPolar (inverse linear and
forward semi-log), cv::createHanningWindow and cv::phaseCorrelate.  The restatement is plain
NumPy (f32 where OpenCV computes in f32, f64 where it computes in f64); the 2-D DFT is evaluated
in f64 — OpenCV's own f32 FFT cannot be reproduced bit for bit, so FMT parity is a tolerance on
the sub-pixel shift, written in the tests (<= 2e-3 px, i.e. <= 4e-5 rad).

Pinned by tests/test_oracle_pins.py::test_fmt_* against live cv2 and by
tests/golden/fmt_fixture.npz (outputs of the unmodified reference on data/tiny).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import numpy as np

from . import restate as R

FMT_DOWNSAMPLE_FACTOR = 10      # FMT.py:10
FMT_RANGE_CLIP_M = 87.5         # FMT.py:11
RANGE_RESOLUTION_CART_M = 0.0432 * 2   # parseData.py:9-11


def resize_cols_linear(img: np.ndarray, new_w: int) -> np.ndarray:
    """cv2.resize(img, (new_w, H)) with INTER_LINEAR when only the width changes (FMT.py:63-65).
    The cv2 4.13 wheel routes f32 INTER_LINEAR through Intel IPP (ippicv 2022.2): sample position
    (dx + 0.5) * scale - 0.5 in double, its fraction cast to f32, then fma(fl(s1 - s0), frac, s0) — bit-exact
    against live cv2 here (tests/test_oracle_pins.py::test_fmt_resize)."""
    img = np.ascontiguousarray(img, np.float32)
    H, W = img.shape
    scale = W / new_w
    f = (np.arange(new_w, dtype=np.float64) + 0.5) * scale - 0.5
    sx = np.floor(f).astype(np.int64)
    fr = (f - sx).astype(np.float32)
    lo = sx < 0
    fr[lo] = 0.0
    sx[lo] = 0
    hi = sx >= W - 1
    fr[hi] = 0.0
    sx[hi] = W - 1
    sx1 = np.minimum(sx + 1, W - 1)
    s0, s1 = img[:, sx], img[:, sx1]
    d = (s1 - s0).astype(np.float32)
    # one rounding: the f64 product of two f32 is exact and the sum is rounded once (= fmaf)
    return (s0.astype(np.float64) + d.astype(np.float64) * fr.astype(np.float64)[None, :]).astype(np.float32)


def warp_polar_forward(cart: np.ndarray, log_mode: bool = True) -> np.ndarray:
    """parseData.convertCartesianImageToPolar(cart, logPolarMode, shapeHW=None) (parseData.py:69-97):
    cv2.warpPolar(cart, None, (h/2, w/2), w/2, WARP_POLAR_LOG|LINEAR + INTER_LINEAR + WARP_FILL_OUTLIERS).
    dsize = (round(maxRadius), round(maxRadius * pi)); map built in double, stored f32; cv::remap bilinear with
    5-bit fixed-point coordinates, BORDER_CONSTANT 0."""
    cart = np.ascontiguousarray(cart, np.float32)
    h, w = cart.shape
    cx, cy = np.float32(h / 2), np.float32(w / 2)
    max_radius = w / 2
    dw, dh = int(np.rint(max_radius)), int(np.rint(max_radius * np.pi))
    k_angle = 2 * np.pi / dh
    rho = np.arange(dw, dtype=np.float64)
    if log_mode:
        k_mag = np.log(max_radius) / dw
        rhos = (np.exp(rho * k_mag) - 1.0).astype(np.float32)
    else:
        rhos = (rho * (max_radius / dw)).astype(np.float32)
    phi = np.arange(dh, dtype=np.float64) * k_angle
    mx = (rhos[None, :].astype(np.float64) * np.cos(phi)[:, None] + np.float64(cx)).astype(np.float32)
    my = (rhos[None, :].astype(np.float64) * np.sin(phi)[:, None] + np.float64(cy)).astype(np.float32)
    return remap_bilinear(cart, mx, my)


def remap_bilinear(src: np.ndarray, mx: np.ndarray, my: np.ndarray) -> np.ndarray:
    """cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with f32 maps: coordinates rounded to 1/32 px
    (cvRound(v * 32)), weights (1-fy)(1-fx) ... as f32 products of multiples of 1/32, accumulation
    ((s00*w00 + s01*w01) + s10*w10) + s11*w11 in f32."""
    h, w = src.shape
    sx = np.rint(mx.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(my.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = ((sx & 31).astype(np.float32) * np.float32(1 / 32)).astype(np.float32)
    fy = ((sy & 31).astype(np.float32) * np.float32(1 / 32)).astype(np.float32)
    one = np.float32(1)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0)).astype(np.float32)

    w00 = ((one - fy) * (one - fx)).astype(np.float32)
    w01 = ((one - fy) * fx).astype(np.float32)
    w10 = (fy * (one - fx)).astype(np.float32)
    w11 = (fy * fx).astype(np.float32)
    acc = (tap(iy, ix) * w00).astype(np.float32)
    acc = (acc + (tap(iy, ix + 1) * w01).astype(np.float32)).astype(np.float32)
    acc = (acc + (tap(iy + 1, ix) * w10).astype(np.float32)).astype(np.float32)
    acc = (acc + (tap(iy + 1, ix + 1) * w11).astype(np.float32)).astype(np.float32)
    return acc


def polar_to_log_polar(polar_ds: np.ndarray) -> np.ndarray:
    """parseData.convertPolarImgToLogPolar (parseData.py:138-157): inverse linear warp without
    down-sampling, then forward semi-log warp."""
    cart = R.warp_polar(polar_ds, downsample=1)
    return warp_polar_forward(cart, log_mode=True)


def hanning_window(rows: int, cols: int) -> np.ndarray:
    """cv2.createHanningWindow((cols, rows), CV_32F) (FMT.py:30): sqrt of the separable Hann product,
    both factors and their product in double, cast to f32, f32 sqrt (bit-exact vs live cv2)."""
    c0, c1 = 2 * np.pi / (cols - 1), 2 * np.pi / (rows - 1)
    wc = 0.5 * (1.0 - np.cos(c0 * np.arange(cols)))
    wr = 0.5 * (1.0 - np.cos(c1 * np.arange(rows)))
    return np.sqrt((wr[:, None] * wc[None, :]).astype(np.float32)).astype(np.float32)


def optimal_dft_size(n: int) -> int:
    """cv::getOptimalDFTSize: the smallest 2^a 3^b 5^c >= n."""
    best = None
    p2 = 1
    while p2 < 2 * n:
        p3 = p2
        while p3 < 2 * n:
            p5 = p3
            while p5 < 2 * n:
                if p5 >= n and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def phase_correlate(a: np.ndarray, b: np.ndarray, window: np.ndarray):
    """cv2.phaseCorrelate(a, b, window) (FMT.py:31-32) -> ((dx, dy), response).
    Zero-pad to the optimal DFT size, window, R = F1 conj(F2) / (|F1 conj(F2)| + eps-guarded), inverse DFT
    (unnormalised), fftshift, arg-max (first in row-major order), 5x5 weighted centroid, response / (M N)."""
    rows, cols = a.shape
    M, N = optimal_dft_size(rows), optimal_dft_size(cols)
    pa = np.zeros((M, N), np.float32)
    pb = np.zeros((M, N), np.float32)
    pa[:rows, :cols] = (window * a).astype(np.float32)
    pb[:rows, :cols] = (window * b).astype(np.float32)
    F1 = np.fft.fft2(pa.astype(np.float64))
    F2 = np.fft.fft2(pb.astype(np.float64))
    P = F1 * np.conj(F2)
    mag = np.abs(P)
    eps = np.finfo(np.float32).eps
    # divSpectrums with a real-valued denominator (the magnitude): P * mag / (mag^2 + eps)
    C = P * mag / (mag * mag + eps)
    c = np.real(np.fft.ifft2(C)) * (M * N)
    c = np.fft.fftshift(c).astype(np.float32)
    peak = int(np.argmax(c))
    py, px = divmod(peak, N)
    minr, maxr = max(py - 2, 0), min(py + 2, M - 1)
    minc, maxc = max(px - 2, 0), min(px + 2, N - 1)
    box = c[minr:maxr + 1, minc:maxc + 1].astype(np.float64)
    ys, xs = np.mgrid[minr:maxr + 1, minc:maxc + 1]
    s = box.sum()
    tx, ty = (xs * box).sum() / (s + np.finfo(np.float64).eps), (ys * box).sum() / (s + np.finfo(np.float64).eps)
    return (N / 2.0 - tx, M / 2.0 - ty), s / (M * N)


def normalize_angle(a: float) -> float:
    """utils.normalize_angles (utils.py): wrap to [-pi, pi)."""
    return float((a + np.pi) % (2 * np.pi) - np.pi)


def rotation_fmt(src_polar: np.ndarray, tgt_polar: np.ndarray, downsample: int = FMT_DOWNSAMPLE_FACTOR,
                 max_range_clip_m: float = FMT_RANGE_CLIP_M):
    """FMT.getRotationUsingFMT (FMT.py:36-90) -> (angle rad, scale, response)."""
    if max_range_clip_m > 0:
        clip = int(max_range_clip_m / RANGE_RESOLUTION_CART_M)
        src_polar, tgt_polar = src_polar[:, :clip], tgt_polar[:, :clip]
    H, W = src_polar.shape
    new_w = int(W // downsample)
    lp_a = polar_to_log_polar(resize_cols_linear(src_polar, new_w))
    lp_b = polar_to_log_polar(resize_cols_linear(tgt_polar, new_w))
    h_lp, w_lp = lp_a.shape
    (dx, dy), response = phase_correlate(lp_a, lp_b, hanning_window(h_lp, w_lp))
    sz = max(h_lp, w_lp)
    angle = normalize_angle(-float(dy) * 2 * np.pi / sz)
    log_base = np.exp(np.log(h_lp / 2) / sz)
    return angle, float(log_base ** dx), float(response)
