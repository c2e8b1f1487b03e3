// k_fmt.cu — Fourier-Mellin rotation prior, for a set of frames / pairs at once.
//
// Replaces (reference file:line):
//   FMT.py:36-90      getRotationUsingFMT: clip, cv2.resize, log-polar resampling, phase correlation
//   FMT.py:13-33      getTranslationUsingPhaseCorrelation: cv2.createHanningWindow + cv2.phaseCorrelate
//   parseData.py:138-157  convertPolarImgToLogPolar = convertPolarImageToCartesian(downsampleFactor=1)
//                         followed by convertCartesianImageToPolar(logPolarMode=True)   (parseData.py:69-135)
//
// Stages (one launch each for all frames / pairs of the call):
//   k_fmt_resize     cv::resize INTER_LINEAR along the range axis (the IPP arithmetic of the cv2 wheel:
//                    fma(fl(s1 - s0), frac, s0), frac = f32 of the double sample position's fraction)
//   k_fmt_cart       cv::warpPolar inverse linear map, 2Wd x 2Wd (same roundings as k_build_map + k_polar2cart)
//   k_fmt_logpolar   cv::warpPolar forward semi-log map (double exp/cos/sin -> f32 map -> 5-bit fixed point ->
//                    f32 bilinear, BORDER_CONSTANT 0), times the Hann window, zero-padded to the DFT size
//   k_fmt_dft_rows / k_fmt_dft_cols   2-D real-to-half-complex DFT: direct row sums, Cooley-Tukey split columns, shared-memory
//                    twiddle tables (sizes are 2^a 3^b 5^c, e.g. 320 x 108: direct sums in f32, tables from f64)
//   k_fmt_cross      R = F1 conj(F2) |F1 conj(F2)| / (|.|^2 + FLT_EPSILON)   (mulSpectrums / magSpectrums / divSpectrums)
//   k_fmt_idft_cols / k_fmt_idft_rows  unnormalised inverse, Hermitian rows -> real
//   k_fmt_peak       fftshift, first arg-max in row-major order, 5x5 weighted centroid in double, response / (M N)
// OpenCV evaluates its FFT in f32 with its own radix schedule, which cannot be reproduced bit for bit; FMT
// parity is therefore a tolerance on the sub-pixel shift (tests: <= 2e-3 px vs live cv2 / the f64 oracle).
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace {

struct FmtDims {
    int A, W, clip, Wd;       // azimuths, input range bins, clipped bins, down-sampled bins
    int n;                    // Cartesian size 2 Wd
    int w_lp, h_lp;           // log-polar size: round(Wd), round(Wd pi)
    int M, N, Nh;             // DFT size (rows, cols) and half-spectrum columns N/2 + 1
};

int optimal_dft_size(int n) {   // cv::getOptimalDFTSize: smallest 2^a 3^b 5^c >= n
    int best = 0;
    for (long p2 = 1; p2 < 2L * n; p2 *= 2)
        for (long p3 = p2; p3 < 2L * n; p3 *= 3)
            for (long p5 = p3; p5 < 2L * n; p5 *= 5)
                if (p5 >= n && (best == 0 || p5 < best)) best = (int)p5;
    return best;
}

}  // namespace

// ------------------------------------------------------------------------------------
// cv2.resize(polar[:, :clip], (Wd, A)) — horizontal linear interpolation only
// ------------------------------------------------------------------------------------
template <typename T> struct FmtSample;
template <> struct FmtSample<float> { static __device__ __forceinline__ float get(const float* p) { return __ldg(p); } };
// resident raw scans: parseData.py:43 power.astype(f32) / 255.
template <> struct FmtSample<uint8_t> { static __device__ __forceinline__ float get(const uint8_t* p) { return __fdiv_rn((float)__ldg(p), 255.0f); } };

template <typename T>
__global__ void __launch_bounds__(128)
k_fmt_resize(const T* __restrict__ polar, int A, size_t pitch, int clip, int Wd, float* __restrict__ out) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y, f = blockIdx.z;
    if (dx >= Wd) return;
    const double scale = (double)clip / (double)Wd;
    const double pos = ((double)dx + 0.5) * scale - 0.5;
    int sx = (int)floor(pos);
    float fr = (float)(pos - (double)sx);
    if (sx < 0) { sx = 0; fr = 0.0f; }
    if (sx >= clip - 1) { sx = clip - 1; fr = 0.0f; }
    const T* row = polar + ((size_t)f * A + a) * pitch;
    const float s0 = FmtSample<T>::get(row + sx), s1 = FmtSample<T>::get(row + min(sx + 1, clip - 1));
    out[((size_t)f * A + a) * Wd + dx] = __fmaf_rn(__fsub_rn(s1, s0), fr, s0);
}

// ------------------------------------------------------------------------------------
// inverse linear warpPolar of the [A][Wd] image, no down-sampling: Cartesian n x n, n = 2 Wd
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fmt_cart(const float* __restrict__ src, int A, int Wd, float* __restrict__ cart) {
    const int n = 2 * Wd;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), f = blockIdx.z;
    if (x >= n || y >= n) return;
    const float c0 = (float)Wd;
    const double Kangle = (2.0 * M_PI) / A, Kmag = (double)Wd / (double)Wd;
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float dx = __fsub_rn((float)x, c0), dy = __fsub_rn((float)y, c0);
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float ax = fabsf(dx), ay = fabsf(dy);
    const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), (float)DBL_EPSILON));
    const float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(p7, c2, p5), c2, p3), c2, p1), c);
    if (ax < ay) a = __fsub_rn(90.0f, a);
    if (dx < 0.f) a = __fsub_rn(180.0f, a);
    if (dy < 0.f) a = __fsub_rn(360.0f, a);
    const float ang = __fmul_rn(a, (float)(M_PI / 180.0));
    const float mapx = __double2float_rn(__ddiv_rn((double)mag, Kmag));
    const float mapy = __fadd_rn(__double2float_rn(__ddiv_rn((double)ang, Kangle)), 1.0f);
    const int sx = __float2int_rn(__fmul_rn(mapx, 32.0f)), sy = __float2int_rn(__fmul_rn(mapy, 32.0f));
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    const float* img = src + (size_t)f * A * Wd;
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = iy + (t >> 1), xx = ix + (t & 1);
        v[t] = 0.0f;
        if (xx >= 0 && xx < Wd && yy >= 0 && yy < A + 2) {   // rows of cv2's wrapped source: azimuth (yy - 1) mod A
            int ar = yy - 1;
            ar = ar < 0 ? ar + A : (ar >= A ? ar - A : ar);
            v[t] = __ldg(img + (size_t)ar * Wd + xx);
        }
    }
    float acc = __fmul_rn(v[0], __fmul_rn(gy, gx));
    acc = __fadd_rn(acc, __fmul_rn(v[1], __fmul_rn(gy, fx)));
    acc = __fadd_rn(acc, __fmul_rn(v[2], __fmul_rn(fy, gx)));
    acc = __fadd_rn(acc, __fmul_rn(v[3], __fmul_rn(fy, fx)));
    cart[((size_t)f * n + y) * n + x] = acc;
}

// ------------------------------------------------------------------------------------
// forward warpPolar of an n x n image, linear or semi-log radius (no window): same map arithmetic as k_fmt_logpolar
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_cart_to_polar(const float* __restrict__ cart, int n, int w_out, int h_out, int semilog, float* __restrict__ out) {
    const int rho = blockIdx.x * blockDim.x + threadIdx.x, phi = blockIdx.y;
    if (rho >= w_out || phi >= h_out) return;
    const double max_radius = (double)n / 2.0;
    const float ctr = (float)(n / 2.0);
    const double Kangle = (2.0 * M_PI) / h_out;
    const double Kmag = semilog ? log(max_radius) / w_out : max_radius / w_out;
    const float r = semilog ? (float)(exp(rho * Kmag) - 1.0) : (float)(rho * Kmag);
    const double ang = Kangle * phi;
    const float mx = (float)__dadd_rn(__dmul_rn((double)r, cos(ang)), (double)ctr);
    const float my = (float)__dadd_rn(__dmul_rn((double)r, sin(ang)), (double)ctr);
    const int sx = __float2int_rn(__fmul_rn(mx, 32.0f)), sy = __float2int_rn(__fmul_rn(my, 32.0f));
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = iy + (t >> 1), xx = ix + (t & 1);
        v[t] = (xx >= 0 && xx < n && yy >= 0 && yy < n) ? __ldg(cart + (size_t)yy * n + xx) : 0.0f;
    }
    float acc = __fmul_rn(v[0], __fmul_rn(gy, gx));
    acc = __fadd_rn(acc, __fmul_rn(v[1], __fmul_rn(gy, fx)));
    acc = __fadd_rn(acc, __fmul_rn(v[2], __fmul_rn(fy, gx)));
    acc = __fadd_rn(acc, __fmul_rn(v[3], __fmul_rn(fy, fx)));
    out[(size_t)phi * w_out + rho] = acc;
}

// ------------------------------------------------------------------------------------
// forward semi-log warpPolar (+ Hann window, zero padding to M x N)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_fmt_logpolar(const float* __restrict__ cart, int n, int w_lp, int h_lp, float* __restrict__ lp,
               float* __restrict__ padded, int M, int N) {
    const int rho = blockIdx.x * blockDim.x + threadIdx.x, phi = blockIdx.y, f = blockIdx.z;
    if (rho >= N) return;
    float* prow = padded + ((size_t)f * M + phi) * N;
    if (phi >= h_lp || rho >= w_lp) { prow[rho] = 0.0f; return; }
    const double max_radius = (double)n / 2.0;                        // parseData.py:85-86
    const float ctr = (float)(n / 2.0);                               // Point2f center
    const double Kangle = (2.0 * M_PI) / h_lp, Kmag = log(max_radius) / w_lp;
    const float r = (float)(exp(rho * Kmag) - 1.0);
    const double ang = Kangle * phi;
    const float mx = (float)__dadd_rn(__dmul_rn((double)r, cos(ang)), (double)ctr);   // no contraction: cv2 rounds the product
    const float my = (float)__dadd_rn(__dmul_rn((double)r, sin(ang)), (double)ctr);
    const int sx = __float2int_rn(__fmul_rn(mx, 32.0f)), sy = __float2int_rn(__fmul_rn(my, 32.0f));
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    const float* img = cart + (size_t)f * n * n;
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = iy + (t >> 1), xx = ix + (t & 1);
        v[t] = (xx >= 0 && xx < n && yy >= 0 && yy < n) ? __ldg(img + (size_t)yy * n + xx) : 0.0f;   // BORDER_CONSTANT 0
    }
    float acc = __fmul_rn(v[0], __fmul_rn(gy, gx));
    acc = __fadd_rn(acc, __fmul_rn(v[1], __fmul_rn(gy, fx)));
    acc = __fadd_rn(acc, __fmul_rn(v[2], __fmul_rn(fy, gx)));
    acc = __fadd_rn(acc, __fmul_rn(v[3], __fmul_rn(fy, fx)));
    if (lp) lp[((size_t)f * h_lp + phi) * w_lp + rho] = acc;
    // cv::createHanningWindow: sqrt((float)(wr * wc)), both factors in double
    const double wc = 0.5 * (1.0 - cos((2.0 * M_PI / (w_lp - 1)) * rho)), wr = 0.5 * (1.0 - cos((2.0 * M_PI / (h_lp - 1)) * phi));
    prow[rho] = __fmul_rn(__fsqrt_rn((float)(wr * wc)), acc);
}

// window + zero padding of caller-supplied images (rf_phase_correlate)
__global__ void __launch_bounds__(128)
k_fmt_window(const float* __restrict__ img, int rows, int cols, float* __restrict__ padded, int M, int N) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= N) return;
    float v = 0.0f;
    if (y < rows && x < cols) {
        const double wc = 0.5 * (1.0 - cos((2.0 * M_PI / (cols - 1)) * x)), wr = 0.5 * (1.0 - cos((2.0 * M_PI / (rows - 1)) * y));
        v = __fmul_rn(__fsqrt_rn((float)(wr * wc)), __ldg(img + ((size_t)f * rows + y) * cols + x));
    }
    padded[((size_t)f * M + y) * N + x] = v;
}

// ------------------------------------------------------------------------------------
// DFT passes.  Spectra are stored column-major ([k][m]) so that the column passes stream.
// ------------------------------------------------------------------------------------
__global__ void k_fmt_twiddles(float2* __restrict__ tw, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const double a = -2.0 * M_PI * (double)i / (double)len;
    tw[i] = make_float2((float)cos(a), (float)sin(a));
}

// rows: Yt[f][k][m] = sum_{n < cols} x[f][m][n] w_N^{k n},  k = 0 .. N/2.   grid (rows, F), block 64
__global__ void __launch_bounds__(64)
k_fmt_dft_rows(const float* __restrict__ x, int M, int N, int rows, int cols, const float2* __restrict__ twN,
               float2* __restrict__ Yt) {
    extern __shared__ float sm[];
    float* xs = sm;                                      // [N]
    float2* tw = reinterpret_cast<float2*>(sm + N);      // [N]  (N is even: 8-byte aligned)
    const int m = blockIdx.x, f = blockIdx.y, Nh = N / 2 + 1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        xs[i] = i < cols ? x[((size_t)f * M + m) * N + i] : 0.0f;
        tw[i] = twN[i];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Nh; k += blockDim.x) {
        float re = 0.0f, im = 0.0f;
        int idx = 0;
        for (int n = 0; n < cols; ++n) {
            const float2 w = tw[idx];
            re = fmaf(xs[n], w.x, re);
            im = fmaf(xs[n], w.y, im);
            idx += k;
            if (idx >= N) idx -= N;
        }
        Yt[((size_t)f * Nh + k) * M + m] = make_float2(re, im);
    }
}

// columns: Zt[f][k][j] = sum_{m < rows} Yt[f][k][m] w_M^{sign j m}.   grid (Nh, F), block <= 256 (loops over outputs)
// One Cooley-Tukey split M = L1 * L2 (L1 ~ sqrt(M); 320 = 16 * 20): with m = L2 m1 + m2 and j = j1 + L1 j2,
//   Z[j1 + L1 j2] = sum_{m2} w_M^{m2 j1} ( sum_{m1} y[L2 m1 + m2] w_M^{L2 m1 j1} ) w_M^{L1 m2 j2}
// i.e. L2 DFTs of length L1, a twiddle, L1 DFTs of length L2: M (L1 + L2) complex MACs per column instead of M^2
// (8.9x fewer at M = 320).  L1 = 1 degenerates to the direct sum (prime lengths).  All twiddles come from the one
// table w_M^i (f32 from f64), so every product uses an exactly tabulated root of unity.
template <int SIGN>
__global__ void __launch_bounds__(256)
k_fmt_dft_cols(const float2* __restrict__ Yt, int M, int Nh, int rows, int L1, const float2* __restrict__ twM,
               float2* __restrict__ Zt) {
    extern __shared__ float sm[];
    float2* ys = reinterpret_cast<float2*>(sm);   // [M] input column (zero beyond `rows`)
    float2* tw = ys + M;                          // [M] w_M^{sign i}
    float2* as = tw + M;                          // [M] stage-1 results A[j1][m2] at j1 * L2 + m2
    const int k = blockIdx.x, f = blockIdx.y, L2 = M / L1;
    const float2* col = Yt + ((size_t)f * Nh + k) * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        ys[i] = i < rows ? col[i] : make_float2(0.0f, 0.0f);
        float2 w = twM[i];
        if (SIGN > 0) w.y = -w.y;
        tw[i] = w;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < M; t += blockDim.x) {       // A[j1][m2] = w^{m2 j1} sum_{m1} y[L2 m1 + m2] w^{L2 m1 j1}
        const int j1 = t / L2, m2 = t - j1 * L2;
        float re = 0.0f, im = 0.0f;
        int idx = 0;
        const int step = (L2 * j1) % M;
        for (int m1 = 0; m1 < L1; ++m1) {
            const float2 w = tw[idx], y = ys[L2 * m1 + m2];
            re = fmaf(y.x, w.x, fmaf(-y.y, w.y, re));
            im = fmaf(y.x, w.y, fmaf(y.y, w.x, im));
            idx += step;
            if (idx >= M) idx -= M;
        }
        const float2 w = tw[(m2 * j1) % M];
        as[t] = make_float2(re * w.x - im * w.y, re * w.y + im * w.x);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < M; t += blockDim.x) {       // Z[j1 + L1 j2] = sum_{m2} A[j1][m2] w^{L1 m2 j2}
        const int j2 = t / L1, j1 = t - j2 * L1;
        float re = 0.0f, im = 0.0f;
        int idx = 0;
        const int step = (L1 * j2) % M;
        const float2* arow = as + j1 * L2;
        for (int m2 = 0; m2 < L2; ++m2) {
            const float2 w = tw[idx], y = arow[m2];
            re = fmaf(y.x, w.x, fmaf(-y.y, w.y, re));
            im = fmaf(y.x, w.y, fmaf(y.y, w.x, im));
            idx += step;
            if (idx >= M) idx -= M;
        }
        Zt[((size_t)f * Nh + k) * M + j1 + L1 * j2] = make_float2(re, im);
    }
}

// cross-power spectrum of pair p = (fa, fb): cv::mulSpectrums(conjB) -> magSpectrums -> divSpectrums
__global__ void __launch_bounds__(256)
k_fmt_cross(const float2* __restrict__ Zt, const int32_t* __restrict__ pair_idx, size_t plane, float2* __restrict__ Ct) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (i >= plane) return;
    const float2 a = Zt[(size_t)pair_idx[2 * p] * plane + i], b = Zt[(size_t)pair_idx[2 * p + 1] * plane + i];
    const float pr = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));      // a conj(b), f32 like mulSpectrums
    const float pi = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
    const float mag = (float)sqrt((double)pr * pr + (double)pi * pi);          // magSpectrums
    const double den = (double)mag * mag + (double)FLT_EPSILON;                // divSpectrums by (mag, 0)
    Ct[(size_t)p * plane + i] = make_float2((float)((double)pr * mag / den), (float)((double)pi * mag / den));
}

// Hermitian rows -> real: c[p][m][n] = D[0] + (-1)^n D[N/2] + 2 sum_{0<k<N/2} Re(D[k] w_N^{-k n}).  grid (M, P), block 128
__global__ void __launch_bounds__(128)
k_fmt_idft_rows(const float2* __restrict__ Dt, int M, int N, const float2* __restrict__ twN, float* __restrict__ c) {
    extern __shared__ float sm[];
    float2* ds = reinterpret_cast<float2*>(sm);   // [Nh]
    const int Nh = N / 2 + 1;
    float2* tw = ds + Nh;                         // [N]
    const int m = blockIdx.x, p = blockIdx.y;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (i < Nh) ds[i] = Dt[((size_t)p * Nh + i) * M + m];
        const float2 w = twN[i];
        tw[i] = make_float2(w.x, -w.y);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = 0.0f;
        int idx = n;                               // k n mod N for k = 1
        for (int k = 1; k < Nh - 1; ++k) {
            const float2 w = tw[idx], d = ds[k];
            acc = fmaf(d.x, w.x, fmaf(-d.y, w.y, acc));
            idx += n;
            if (idx >= N) idx -= N;
        }
        const float edge = ds[0].x + ((n & 1) ? -ds[Nh - 1].x : ds[Nh - 1].x);
        c[((size_t)p * M + m) * N + n] = fmaf(2.0f, acc, edge);
    }
}

// fftshift + first arg-max + 5x5 weighted centroid (cv::phaseCorrelate tail).  grid P, block 256
__global__ void __launch_bounds__(256)
k_fmt_peak(const float* __restrict__ c, int M, int N, double* __restrict__ out /*[P][3]: dx, dy, response*/) {
    __shared__ float s_val[256];
    __shared__ int s_idx[256];
    const int p = blockIdx.x;
    const float* img = c + (size_t)p * M * N;
    const int hM = M / 2, hN = N / 2;
    // shifted(ys, xs) = c[(ys + M - hM') % M][...]: cv's fftShift swaps quadrants (even sizes) / rotates (odd sizes)
    float best = -FLT_MAX;
    int best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
        const int ys = i / N, xs = i - ys * N;
        const int m = (ys + (M - hM)) % M, n = (xs + (N - hN)) % N;
        const float v = img[(size_t)m * N + n];
        if (v > best) { best = v; best_i = i; }     // i ascends per thread: the first maximum is kept
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float v = s_val[threadIdx.x + s];
            const int ii = s_idx[threadIdx.x + s];
            if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && ii < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = ii; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int py = s_idx[0] / N, px = s_idx[0] - py * N;
        const int minr = max(py - 2, 0), maxr = min(py + 2, M - 1), minc = max(px - 2, 0), maxc = min(px + 2, N - 1);
        double cx = 0.0, cy = 0.0, sum = 0.0;
        for (int y = minr; y <= maxr; ++y)
            for (int x = minc; x <= maxc; ++x) {
                const double v = (double)img[(size_t)((y + (M - hM)) % M) * N + (x + (N - hN)) % N];
                cx += (double)x * v; cy += (double)y * v; sum += v;
            }
        const double resp = sum;
        sum += DBL_EPSILON;
        out[3 * p + 0] = (double)N / 2.0 - cx / sum;
        out[3 * p + 1] = (double)M / 2.0 - cy / sum;
        out[3 * p + 2] = resp / ((double)M * (double)N);
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
namespace {

size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// phase correlation of P pairs over F padded windowed images already at d_pad [F][M][N]; results to d_out [P][3]
int fmt_correlate(rf_handle* h, char* ws, const float* d_pad, int F, const int32_t* d_pairs, int P, int M, int N, int rows,
                  int cols, double* d_out) {
    const int Nh = N / 2 + 1;
    const size_t plane = (size_t)Nh * M;
    float2* twN = (float2*)ws;                        ws += al256((size_t)N * 8);
    float2* twM = (float2*)ws;                        ws += al256((size_t)M * 8);
    float2* Yt = (float2*)ws;                         ws += al256((size_t)F * plane * 8);
    float2* Zt = (float2*)ws;                         ws += al256((size_t)F * plane * 8);
    float2* Ct = (float2*)ws;                         ws += al256((size_t)P * plane * 8);
    float2* Dt = (float2*)ws;                         ws += al256((size_t)P * plane * 8);
    float* c = (float*)ws;
    k_fmt_twiddles<<<(N + 127) / 128, 128, 0, h->stream>>>(twN, N);
    RF_CHECK_LAUNCH(h);
    k_fmt_twiddles<<<(M + 127) / 128, 128, 0, h->stream>>>(twM, M);
    RF_CHECK_LAUNCH(h);
    k_fmt_dft_rows<<<dim3(rows, F), 64, (size_t)N * 12, h->stream>>>(d_pad, M, N, rows, cols, twN, Yt);
    RF_CHECK_LAUNCH(h);
    const int passes = (M + 255) / 256, ct = ((M + passes - 1) / passes + 31) / 32 * 32;   // threads per column block
    int L1 = 1;                                    // the divisor of M closest to sqrt(M) from below (320 -> 16)
    for (int d = 1; d * d <= M; ++d) if (M % d == 0) L1 = d;
    k_fmt_dft_cols<-1><<<dim3(Nh, F), ct, (size_t)M * 24, h->stream>>>(Yt, M, Nh, rows, L1, twM, Zt);
    RF_CHECK_LAUNCH(h);
    k_fmt_cross<<<dim3((unsigned)((plane + 255) / 256), P), 256, 0, h->stream>>>(Zt, d_pairs, plane, Ct);
    RF_CHECK_LAUNCH(h);
    k_fmt_dft_cols<1><<<dim3(Nh, P), ct, (size_t)M * 24, h->stream>>>(Ct, M, Nh, M, L1, twM, Dt);
    RF_CHECK_LAUNCH(h);
    k_fmt_idft_rows<<<dim3(M, P), 128, (size_t)(Nh + N) * 8, h->stream>>>(Dt, M, N, twN, c);
    RF_CHECK_LAUNCH(h);
    k_fmt_peak<<<P, 256, 0, h->stream>>>(c, M, N, d_out);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

size_t fmt_correlate_ws(int F, int P, int M, int N) {
    const size_t plane = (size_t)(N / 2 + 1) * M;
    return al256((size_t)N * 8) + al256((size_t)M * 8) + 2 * al256((size_t)F * plane * 8) + 2 * al256((size_t)P * plane * 8) +
           al256((size_t)P * M * N * 4);
}

int fmt_dims(rf_handle* h, int A, int W, int downsample, int clip_px, FmtDims* d) {
    if (A < 2 || W < 4 || downsample < 1) return rf_fail(h, RF_E_BADARG, "rf_fmt: bad image size");
    d->A = A; d->W = W;
    d->clip = (clip_px > 0 && clip_px < W) ? clip_px : W;           // FMT.py:55-58
    d->Wd = d->clip / downsample;                                   // FMT.py:62
    if (d->Wd < 4) return rf_fail(h, RF_E_BADARG, "rf_fmt: down-sampled width %d too small", d->Wd);
    d->n = 2 * d->Wd;
    d->w_lp = (int)lrint((double)d->n / 2.0);                       // cvRound(maxRadius)
    d->h_lp = (int)lrint((double)d->n / 2.0 * M_PI);                // cvRound(maxRadius * CV_PI)
    d->M = optimal_dft_size(d->h_lp);
    d->N = optimal_dft_size(d->w_lp);
    if (d->N & 1) d->N = optimal_dft_size(d->N + 1);                // keep the Nyquist column (sizes here are even anyway)
    d->Nh = d->N / 2 + 1;
    if ((size_t)d->M * 24 > 48 * 1024) return rf_fail(h, RF_E_CAPACITY, "rf_fmt: DFT size %d exceeds shared memory", d->M);
    return RF_OK;
}

}  // namespace

// FMT rotation of P pairs over F scans already resident on the device (u8 power bins, row pitch `pitch`):
// results [P][3] (dx, dy, response) left at *d_out_p inside the handle scratch; everything on h->stream.
int rf_fmt_resident_u8(rf_handle* h, const uint8_t* d_raw, int F, int A, int W, size_t pitch, const int32_t* d_pairs, int P,
                       int downsample, int clip_px, double** d_out_p, int* sz_out, double* log_base_out) {
    FmtDims d;
    int rc = fmt_dims(h, A, W, downsample, clip_px, &d);
    if (rc) return rc;
    const size_t b_rs = al256((size_t)F * A * d.Wd * 4), b_cart = al256((size_t)F * d.n * d.n * 4);
    const size_t b_pad = al256((size_t)F * d.M * d.N * 4), b_out = al256((size_t)P * 24);
    rc = rf_ensure_scratch(h, b_rs + b_cart + b_pad + b_out + fmt_correlate_ws(F, P, d.M, d.N));
    if (rc) return rc;
    char* ws = (char*)h->d_scratch;
    float* d_rs = (float*)ws;         ws += b_rs;
    float* d_cart = (float*)ws;       ws += b_cart;
    float* d_pad = (float*)ws;        ws += b_pad;
    double* d_out = (double*)ws;      ws += b_out;
    k_fmt_resize<uint8_t><<<dim3((d.Wd + 127) / 128, A, F), 128, 0, h->stream>>>(d_raw, A, pitch, d.clip, d.Wd, d_rs);
    RF_CHECK_LAUNCH(h);
    k_fmt_cart<<<dim3((d.n + 31) / 32, (d.n + 7) / 8, F), 256, 0, h->stream>>>(d_rs, A, d.Wd, d_cart);
    RF_CHECK_LAUNCH(h);
    k_fmt_logpolar<<<dim3((d.N + 127) / 128, d.M, F), 128, 0, h->stream>>>(d_cart, d.n, d.w_lp, d.h_lp, nullptr, d_pad, d.M, d.N);
    RF_CHECK_LAUNCH(h);
    rc = fmt_correlate(h, ws, d_pad, F, d_pairs, P, d.M, d.N, d.h_lp, d.w_lp, d_out);
    if (rc) return rc;
    *d_out_p = d_out;
    *sz_out = d.h_lp > d.w_lp ? d.h_lp : d.w_lp;
    *log_base_out = exp(log((double)d.h_lp / 2.0) / *sz_out);
    return RF_OK;
}

// FMT.py:79-87: shift -> (normalised angle, scale)
void rf_fmt_finish(const double* out3, int P, int sz, double log_base, double* angle_rad, double* scale, double* response,
                   double* shift_xy) {
    for (int p = 0; p < P; ++p) {
        const double dx = out3[3 * p], dy = out3[3 * p + 1];
        double ang = fmod(-dy * 2.0 * M_PI / sz + M_PI, 2.0 * M_PI);    // FMT.py:82, utils.normalize_angles
        if (ang < 0) ang += 2.0 * M_PI;
        angle_rad[p] = ang - M_PI;
        if (scale) scale[p] = pow(log_base, dx);                         // FMT.py:87
        if (response) response[p] = out3[3 * p + 2];
        if (shift_xy) { shift_xy[2 * p] = dx; shift_xy[2 * p + 1] = dy; }
    }
}

extern "C" {

// polar [n_frames][A][W] f32 (host) -> per pair: angle (rad), scale, response, shift (dx, dy)
static int fmt_rotation_impl(rf_handle* h, const float* polar, const float* const* frames, int n_frames, int A, int W,
                             const int32_t* pair_idx, int n_pairs, int downsample, int clip_px, double* angle_rad, double* scale,
                             double* response, double* shift_xy) {
    if (!h || (!polar && !frames) || !pair_idx || n_frames < 1 || n_pairs < 1 || !angle_rad)
        return rf_fail(h, RF_E_BADARG, "rf_fmt_rotation: bad argument");
    if (frames)
        for (int i = 0; i < n_frames; ++i)
            if (!frames[i]) return rf_fail(h, RF_E_BADARG, "rf_fmt_rotation: null frame pointer");
    for (int i = 0; i < 2 * n_pairs; ++i)
        if (pair_idx[i] < 0 || pair_idx[i] >= n_frames) return rf_fail(h, RF_E_BADARG, "rf_fmt_rotation: pair index out of range");
    FmtDims d;
    int rc = fmt_dims(h, A, W, downsample, clip_px, &d);
    if (rc) return rc;
    const int F = n_frames, P = n_pairs;
    const size_t b_polar = al256((size_t)F * A * W * 4), b_rs = al256((size_t)F * A * d.Wd * 4), b_cart = al256((size_t)F * d.n * d.n * 4);
    const size_t b_pad = al256((size_t)F * d.M * d.N * 4), b_pairs = al256((size_t)P * 8), b_out = al256((size_t)P * 24);
    rc = rf_ensure_scratch(h, b_polar + b_rs + b_cart + b_pad + b_pairs + b_out + fmt_correlate_ws(F, P, d.M, d.N));
    if (rc) return rc;
    char* ws = (char*)h->d_scratch;
    float* d_polar = (float*)ws;      ws += b_polar;
    float* d_rs = (float*)ws;         ws += b_rs;
    float* d_cart = (float*)ws;       ws += b_cart;
    float* d_pad = (float*)ws;        ws += b_pad;
    int32_t* d_pairs = (int32_t*)ws;  ws += b_pairs;
    double* d_out = (double*)ws;      ws += b_out;
    if (polar) RF_CUDA(h, cudaMemcpyAsync(d_polar, polar, (size_t)F * A * W * 4, cudaMemcpyHostToDevice, h->stream));
    else
        for (int i = 0; i < F; ++i)   // separately allocated images (the reference's call passes two arrays): no host-side stacking
            RF_CUDA(h, cudaMemcpyAsync(d_polar + (size_t)i * A * W, frames[i], (size_t)A * W * 4, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(d_pairs, pair_idx, (size_t)P * 8, cudaMemcpyHostToDevice, h->stream));
    k_fmt_resize<float><<<dim3((d.Wd + 127) / 128, A, F), 128, 0, h->stream>>>(d_polar, A, (size_t)W, d.clip, d.Wd, d_rs);
    RF_CHECK_LAUNCH(h);
    k_fmt_cart<<<dim3((d.n + 31) / 32, (d.n + 7) / 8, F), 256, 0, h->stream>>>(d_rs, A, d.Wd, d_cart);
    RF_CHECK_LAUNCH(h);
    k_fmt_logpolar<<<dim3((d.N + 127) / 128, d.M, F), 128, 0, h->stream>>>(d_cart, d.n, d.w_lp, d.h_lp, nullptr, d_pad, d.M, d.N);
    RF_CHECK_LAUNCH(h);
    rc = fmt_correlate(h, ws, d_pad, F, d_pairs, P, d.M, d.N, d.h_lp, d.w_lp, d_out);
    if (rc) return rc;
    std::vector<double> out((size_t)P * 3);
    RF_CUDA(h, cudaMemcpyAsync(out.data(), d_out, (size_t)P * 24, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    const int sz = d.h_lp > d.w_lp ? d.h_lp : d.w_lp;                    // FMT.py:79-80
    rf_fmt_finish(out.data(), P, sz, exp(log((double)d.h_lp / 2.0) / sz), angle_rad, scale, response, shift_xy);
    return RF_OK;
}

int rf_fmt_rotation(rf_handle* h, const float* polar, int n_frames, int A, int W, const int32_t* pair_idx, int n_pairs,
                    int downsample, int clip_px, double* angle_rad, double* scale, double* response, double* shift_xy) {
    RfDeviceGuard rf_guard_(h);
    return fmt_rotation_impl(h, polar, nullptr, n_frames, A, W, pair_idx, n_pairs, downsample, clip_px, angle_rad, scale, response,
                             shift_xy);
}

// the same with one pointer per image ([A, W] f32 each) instead of one stacked array
int rf_fmt_rotation_frames(rf_handle* h, const float* const* frames, int n_frames, int A, int W, const int32_t* pair_idx,
                           int n_pairs, int downsample, int clip_px, double* angle_rad, double* scale, double* response,
                           double* shift_xy) {
    RfDeviceGuard rf_guard_(h);
    return fmt_rotation_impl(h, nullptr, frames, n_frames, A, W, pair_idx, n_pairs, downsample, clip_px, angle_rad, scale, response,
                             shift_xy);
}

// parseData.convertPolarImgToLogPolar(cv2.resize(polar[:, :clip_px], (clip_px // downsample, A)))
int rf_fmt_log_polar(rf_handle* h, const float* polar, int A, int W, int downsample, int clip_px, float* out, int64_t out_cap,
                     int* h_lp, int* w_lp) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !polar || !h_lp || !w_lp) return rf_fail(h, RF_E_BADARG, "rf_fmt_log_polar: bad argument");
    FmtDims d;
    int rc = fmt_dims(h, A, W, downsample, clip_px, &d);
    if (rc) return rc;
    *h_lp = d.h_lp; *w_lp = d.w_lp;
    if (!out) return RF_OK;                                              // size query
    if (out_cap < (int64_t)d.h_lp * d.w_lp) return rf_fail(h, RF_E_CAPACITY, "rf_fmt_log_polar: output buffer too small");
    const size_t b_polar = al256((size_t)A * W * 4), b_rs = al256((size_t)A * d.Wd * 4), b_cart = al256((size_t)d.n * d.n * 4);
    const size_t b_pad = al256((size_t)d.M * d.N * 4), b_lp = al256((size_t)d.h_lp * d.w_lp * 4);
    rc = rf_ensure_scratch(h, b_polar + b_rs + b_cart + b_pad + b_lp);
    if (rc) return rc;
    char* ws = (char*)h->d_scratch;
    float* d_polar = (float*)ws;      ws += b_polar;
    float* d_rs = (float*)ws;         ws += b_rs;
    float* d_cart = (float*)ws;       ws += b_cart;
    float* d_pad = (float*)ws;        ws += b_pad;
    float* d_lp = (float*)ws;
    RF_CUDA(h, cudaMemcpyAsync(d_polar, polar, (size_t)A * W * 4, cudaMemcpyHostToDevice, h->stream));
    k_fmt_resize<float><<<dim3((d.Wd + 127) / 128, A, 1), 128, 0, h->stream>>>(d_polar, A, (size_t)W, d.clip, d.Wd, d_rs);
    RF_CHECK_LAUNCH(h);
    k_fmt_cart<<<dim3((d.n + 31) / 32, (d.n + 7) / 8, 1), 256, 0, h->stream>>>(d_rs, A, d.Wd, d_cart);
    RF_CHECK_LAUNCH(h);
    k_fmt_logpolar<<<dim3((d.N + 127) / 128, d.M, 1), 128, 0, h->stream>>>(d_cart, d.n, d.w_lp, d.h_lp, d_lp, d_pad, d.M, d.N);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(out, d_lp, (size_t)d.h_lp * d.w_lp * 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

// parseData.convertCartesianImageToPolar(imgCart, logPolarMode, shapeHW)          parseData.py:69-97
// cv::warpPolar forward map (linear or semi-log), INTER_LINEAR | WARP_FILL_OUTLIERS, centre (n / 2, n / 2), maxRadius n / 2;
// dsize (0, 0) selects cv's default (round(maxRadius), round(maxRadius * pi)).
int rf_cart_to_polar(rf_handle* h, const float* cart, int n, int log_mode, int rows_out, int cols_out, float* out, int64_t out_cap,
                     int* rows, int* cols) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !cart || n < 2 || !rows || !cols) return rf_fail(h, RF_E_BADARG, "rf_cart_to_polar: bad argument");
    const double max_radius = (double)n / 2.0;
    const int w = cols_out > 0 ? cols_out : (int)nearbyint(max_radius);
    const int hgt = rows_out > 0 ? rows_out : (int)nearbyint(max_radius * M_PI);
    *rows = hgt; *cols = w;
    if (!out) return RF_OK;
    if (out_cap < (int64_t)hgt * w) return rf_fail(h, RF_E_CAPACITY, "rf_cart_to_polar: output buffer too small");
    const size_t b_cart = al256((size_t)n * n * 4), b_out = al256((size_t)hgt * w * 4);
    int rc = rf_ensure_scratch(h, b_cart + b_out);
    if (rc) return rc;
    float* d_cart = (float*)h->d_scratch;
    float* d_out = (float*)((char*)h->d_scratch + b_cart);
    RF_CUDA(h, cudaMemcpyAsync(d_cart, cart, (size_t)n * n * 4, cudaMemcpyHostToDevice, h->stream));
    k_cart_to_polar<<<dim3((w + 127) / 128, hgt, 1), 128, 0, h->stream>>>(d_cart, n, w, hgt, log_mode ? 1 : 0, d_out);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)hgt * w * 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

// cv2.phaseCorrelate(a, b, cv2.createHanningWindow((cols, rows), CV_32F)) -> (dx, dy), response   (FMT.py:13-33)
int rf_phase_correlate(rf_handle* h, const float* a, const float* b, int rows, int cols, double* dx, double* dy, double* response) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !a || !b || rows < 2 || cols < 2 || !dx || !dy) return rf_fail(h, RF_E_BADARG, "rf_phase_correlate: bad argument");
    const int M = optimal_dft_size(rows);
    int N = optimal_dft_size(cols);
    if (N & 1) N = optimal_dft_size(N + 1);
    if ((size_t)M * 24 > 48 * 1024 || (size_t)N * 12 > 48 * 1024) return rf_fail(h, RF_E_CAPACITY, "rf_phase_correlate: image too large");
    const size_t b_img = al256((size_t)2 * rows * cols * 4), b_pad = al256((size_t)2 * M * N * 4);
    int rc = rf_ensure_scratch(h, b_img + b_pad + 512 + fmt_correlate_ws(2, 1, M, N));
    if (rc) return rc;
    char* ws = (char*)h->d_scratch;
    float* d_img = (float*)ws;        ws += b_img;
    float* d_pad = (float*)ws;        ws += b_pad;
    int32_t* d_pairs = (int32_t*)ws;  ws += 256;
    double* d_out = (double*)ws;      ws += 256;
    const int32_t pr[2] = {0, 1};
    RF_CUDA(h, cudaMemcpyAsync(d_img, a, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(d_img + (size_t)rows * cols, b, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(d_pairs, pr, 8, cudaMemcpyHostToDevice, h->stream));
    k_fmt_window<<<dim3((N + 127) / 128, M, 2), 128, 0, h->stream>>>(d_img, rows, cols, d_pad, M, N);
    RF_CHECK_LAUNCH(h);
    rc = fmt_correlate(h, ws, d_pad, 2, d_pairs, 1, M, N, rows, cols, d_out);
    if (rc) return rc;
    double out[3];
    RF_CUDA(h, cudaMemcpyAsync(out, d_out, 24, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    *dx = out[0]; *dy = out[1];
    if (response) *response = out[2];
    return RF_OK;
}

}  // extern "C"
