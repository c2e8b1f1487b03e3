// k_doh.cu — the reference's feature detector: determinant-of-Hessian blobs.
//
// Replaces getFeatures.getBlobsFromCart (getFeatures.py:13-18,22-53), i.e.
//   skimage.feature.blob_doh(img.astype(np.double), min_sigma=.01, max_sigma=10, num_sigma=3, threshold=.0005)
// following the published algorithm of scikit-image 0.19.2 as restated in oracle/doh_restate.py (PARITY UNPINNED: the
// package is absent from the reference tree and from this image; the restatement's header lists the three
// implementation-defined points — NaN plane of a zero-size box, order of equal responses, pruning order):
//   integral_image                      float64 cumsum down the columns, then along the rows   -> k_doh_colsum, k_doh_rowsum
//   _hessian_matrix_det (box filters)   per sigma: dxx, dyy, dxy from clipped-window sums       -> doh_det (device function)
//   peak_local_max(cube, 3x3x3)         == maximum filter (SciPy ring buffer along the scales)  -> k_doh_peaks
//   argsort(-response), _prune_blobs    sort, overlap pruning in ascending pair order           -> k_doh_finish
//   adaptiveNMS' argsort(sigma)         stable partition by scale                               -> k_doh_finish (ws.rc)
//
// Arithmetic is float64 in the operand order of the Cython / Python sources with explicit round-to-nearest
// intrinsics (no fused multiply-add), so the response planes equal the restatement bit for bit; the sequential
// cumulative sums are kept sequential along the summed axis and parallel across the other one.
#include <math.h>

#include "detect.cuh"

#define DOH_MAX_SIGMA 16
#define DOH_CAP 8192                 // peaks per frame the finishing kernel sorts and prunes (real scans: a few hundred)

struct DohScale { int nan, l, b, w; double w_i, sigma; };
struct DohParams { int ns; DohScale s[DOH_MAX_SIGMA]; };

// blob_doh's sigma list (np.linspace) and _hessian_matrix_det's box geometry (C integer division: cdivision=True)
static int doh_make_params(rf_handle* h, double min_sigma, double max_sigma, int num_sigma, DohParams* P) {
    if (num_sigma < 1 || num_sigma > DOH_MAX_SIGMA)
        return rf_fail(h, RF_E_BADARG, "DoH: num_sigma %d outside 1..%d", num_sigma, DOH_MAX_SIGMA);
    if (!(min_sigma >= 0) || !(max_sigma >= min_sigma)) return rf_fail(h, RF_E_BADARG, "DoH: bad sigma range [%g, %g]", min_sigma, max_sigma);
    P->ns = num_sigma;
    const double step = num_sigma > 1 ? (max_sigma - min_sigma) / (double)(num_sigma - 1) : 0.0;
    for (int k = 0; k < num_sigma; ++k) {
        // np.linspace: start + arange(num) * step, last sample forced to `stop`
        double s = min_sigma + (double)k * step;
        if (k == num_sigma - 1 && num_sigma > 1) s = max_sigma;
        const int size = (int)(3.0 * s);
        DohScale& d = P->s[k];
        d.sigma = s;
        d.nan = size == 0;                          // w_i = 1.0 / 0 / 0 = inf, dxy = -0.0 * inf = NaN on the whole plane
        d.l = size / 3; d.b = (size - 1) / 2; d.w = size;
        d.w_i = size ? 1.0 / (double)size / (double)size : 0.0;
    }
    return RF_OK;
}

// ---- integral image ----------------------------------------------------------------------------------
// S = img.astype(f64).cumsum(axis=0): one thread per column, rows in order (sequential float64 additions)
__global__ void __launch_bounds__(128)
k_doh_colsum(const float* __restrict__ img_base, size_t img_stride, int n, double* __restrict__ ii_base, size_t ii_stride,
             const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float* __restrict__ img = img_base + (size_t)p * img_stride + c;
    double* __restrict__ ii = ii_base + (size_t)p * ii_stride + c;
    double acc = 0.0;
    int r = 0;
    for (; r + 8 <= n; r += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(img + (size_t)(r + k) * n);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc = __dadd_rn(acc, (double)v[k]); ii[(size_t)(r + k) * n] = acc; }
    }
    for (; r < n; ++r) { acc = __dadd_rn(acc, (double)__ldg(img + (size_t)r * n)); ii[(size_t)r * n] = acc; }
}

// S = S.cumsum(axis=1), in place: a CTA owns 32 rows and walks them 32 columns at a time through shared memory;
// one thread per row adds sequentially and carries the running sum
__global__ void __launch_bounds__(256)
k_doh_rowsum(double* __restrict__ ii_base, size_t ii_stride, int n, const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    __shared__ double tile[32][33];
    double* __restrict__ ii = ii_base + (size_t)p * ii_stride;
    const int r0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 warps: warp ty loads rows ty, ty + 8, ...
    double carry = 0.0;
    for (int c0 = 0; c0 < n; c0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            tile[ty + 8 * k][tx] = (r < n && c < n) ? ii[(size_t)r * n + c] : 0.0;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lim = min(32, n - c0);
            for (int k = 0; k < lim; ++k) { carry = __dadd_rn(carry, tile[threadIdx.x][k]); tile[threadIdx.x][k] = carry; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            if (r < n && c < n) ii[(size_t)r * n + c] = tile[ty + 8 * k][tx];
        }
        __syncthreads();
    }
}

// ---- _hessian_det_appx.pyx ---------------------------------------------------------------------------
__device__ __forceinline__ double doh_integ(const double* __restrict__ ii, int n, int r, int c, int rl, int cl) {
    r = min(max(r, 0), n - 1); c = min(max(c, 0), n - 1);
    const int r2 = min(max(r + rl, 0), n - 1), c2 = min(max(c + cl, 0), n - 1);     // from the CLIPPED corner, as the source does
    const double ans = __dsub_rn(__dsub_rn(__dadd_rn(__ldg(ii + (size_t)r * n + c), __ldg(ii + (size_t)r2 * n + c2)),
                                           __ldg(ii + (size_t)r * n + c2)), __ldg(ii + (size_t)r2 * n + c));
    return ans > 0.0 ? ans : 0.0;                   // max(0, ans)
}

__device__ double doh_det(const double* __restrict__ ii, int n, int r, int c, const DohScale& s) {
    if (s.nan) return __longlong_as_double(0x7ff8000000000000LL);
    const int l = s.l, b = s.b, w = s.w;
    const double tl = doh_integ(ii, n, r - l, c - l, l, l);
    const double br = doh_integ(ii, n, r + 1, c + 1, l, l);
    const double bl = doh_integ(ii, n, r - l, c + 1, l, l);
    const double tr = doh_integ(ii, n, r + 1, c - l, l, l);
    double dxy = __dsub_rn(__dsub_rn(__dadd_rn(bl, tr), tl), br);
    dxy = __dmul_rn(-dxy, s.w_i);
    double mid = doh_integ(ii, n, r - l + 1, c - b, 2 * l - 1, w);
    double side = doh_integ(ii, n, r - l + 1, c - l / 2, 2 * l - 1, l);
    double dxx = __dsub_rn(mid, __dmul_rn(3.0, side));
    dxx = __dmul_rn(-dxx, s.w_i);
    mid = doh_integ(ii, n, r - b, c - l + 1, w, 2 * l - 1);
    side = doh_integ(ii, n, r - l / 2, c - l + 1, l, 2 * l - 1);
    double dyy = __dsub_rn(mid, __dmul_rn(3.0, side));
    dyy = __dmul_rn(-dyy, s.w_i);
    return __dsub_rn(__dmul_rn(dxx, dyy), __dmul_rn(0.81, __dmul_rn(dxy, dxy)));
}

// one response plane to global memory (test hook rf_doh_response)
__global__ void __launch_bounds__(256) k_doh_plane(const double* __restrict__ ii, int n, DohScale s, double* __restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r < n && c < n) out[(size_t)r * n + c] = doh_det(ii, n, r, c, s);
}

// ---- peak_local_max ------------------------------------------------------------------------------------
// scipy.ndimage.maximum_filter1d (size 3, mode 'nearest') along a line of NS values, NI_MinOrMaximumFilter1D's ring
// buffer: every comparison with a NaN is false, which is what decides the result on a NaN plane.
template <int NS>
__device__ __forceinline__ void ring_max3(const double (&M)[NS], double (&out)[NS]) {
    double line[NS + 2];
    line[0] = M[0];
#pragma unroll
    for (int k = 0; k < NS; ++k) line[k + 1] = M[k];
    line[NS + 1] = M[NS - 1];
    double rv[3]; int rd[3];
    int mp = 0, last = 0, o = 0;
    rv[0] = line[0]; rd[0] = 3; rv[1] = rv[2] = 0.0; rd[1] = rd[2] = 0;
#pragma unroll
    for (int ll = 1; ll < NS + 2; ++ll) {
        const double val = line[ll];
        if (rd[mp] == ll) mp = mp == 2 ? 0 : mp + 1;
        if (val >= rv[mp]) { rv[mp] = val; rd[mp] = ll + 3; last = mp; }
        else {
            while (rv[last] <= val) last = last == 0 ? 2 : last - 1;
            last = last == 2 ? 0 : last + 1;
            rv[last] = val; rd[last] = ll + 3;
        }
        if (ll >= 2) out[o++] = rv[mp];
    }
}

#define DP_T 32                        // outputs per tile side
template <int NS>
__global__ void __launch_bounds__(256)
k_doh_peaks(const double* __restrict__ ii_base, size_t ii_stride, int n, DohParams P, double thr, double* __restrict__ vals_base,
            uint32_t* __restrict__ idx_base, unsigned* __restrict__ count, const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const double* __restrict__ ii = ii_base + (size_t)p * ii_stride;
    double* __restrict__ vals = vals_base + (size_t)p * DOH_CAP;
    uint32_t* __restrict__ idx = idx_base + (size_t)p * DOH_CAP;
    __shared__ double det[DP_T + 2][DP_T + 3];
    const int tiles_x = (n + DP_T - 1) / DP_T, ntiles = tiles_x * tiles_x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int ox = (t % tiles_x) * DP_T, oy = (t / tiles_x) * DP_T;
        double v[4][NS], M[4][NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (P.s[k].nan) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { v[q][k] = qnan; M[q][k] = qnan; }
                continue;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < (DP_T + 2) * (DP_T + 2); i += 256) {
                const int rr = i / (DP_T + 2), cc = i - rr * (DP_T + 2);
                // mode 'nearest': positions outside the image repeat the border value
                const int r = min(max(oy - 1 + rr, 0), n - 1), c = min(max(ox - 1 + cc, 0), n - 1);
                det[rr][cc] = doh_det(ii, n, r, c, P.s[k]);
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rr = ty + 8 * q + 1, cc = tx + 1;
                double m = det[rr][cc];
                v[q][k] = m;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) m = fmax(m, det[rr + dy][cc + dx]);
                M[q][k] = m;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = oy + ty + 8 * q, c = ox + tx;
            if (r >= n || c >= n) continue;
            bool any = false;
#pragma unroll
            for (int k = 0; k < NS; ++k) any |= v[q][k] > thr;
            if (!any) continue;
            double fm[NS];
            ring_max3<NS>(M[q], fm);
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                if (v[q][k] == fm[k] && v[q][k] > thr) {
                    const unsigned slot = atomicAdd(count + p, 1u);
                    if (slot < DOH_CAP) { vals[slot] = v[q][k]; idx[slot] = ((uint32_t)r * (uint32_t)n + (uint32_t)c) * NS + k; }
                }
            }
        }
    }
}

// ---- _prune_blobs ------------------------------------------------------------------------------------
__device__ double doh_disk_overlap(double d, double r1, double r2) {
    const double d2 = __dmul_rn(d, d), r12 = __dmul_rn(r1, r1), r22 = __dmul_rn(r2, r2);
    double ratio1 = __ddiv_rn(__dsub_rn(__dadd_rn(d2, r12), r22), __dmul_rn(__dmul_rn(2.0, d), r1));
    ratio1 = fmin(fmax(ratio1, -1.0), 1.0);
    const double acos1 = acos(ratio1);
    double ratio2 = __ddiv_rn(__dsub_rn(__dadd_rn(d2, r22), r12), __dmul_rn(__dmul_rn(2.0, d), r2));
    ratio2 = fmin(fmax(ratio2, -1.0), 1.0);
    const double acos2 = acos(ratio2);
    const double a = __dadd_rn(__dadd_rn(-d, r2), r1);
    const double b = __dadd_rn(__dsub_rn(d, r2), r1);
    const double c = __dsub_rn(__dadd_rn(d, r2), r1);
    const double dd = __dadd_rn(__dadd_rn(d, r2), r1);
    const double prod = __dmul_rn(__dmul_rn(__dmul_rn(a, b), c), dd);
    const double area = __dsub_rn(__dadd_rn(__dmul_rn(r12, acos1), __dmul_rn(r22, acos2)), __dmul_rn(0.5, sqrt(fabs(prod))));
    const double rm = fmin(r1, r2);
    return __ddiv_rn(area, __dmul_rn(3.141592653589793, __dmul_rn(rm, rm)));
}

// _blob_overlap of two live 2-D blobs (row, col, sigma > 0)
__device__ double doh_blob_overlap(double y1, double x1, double s1, double y2, double x2, double s2) {
    const double root = 1.4142135623730951;      // math.sqrt(2)
    double r1, r2, ms;
    if (s1 > s2) { ms = s1; r1 = 1.0; r2 = __ddiv_rn(s2, s1); }
    else { ms = s2; r2 = 1.0; r1 = __ddiv_rn(s1, s2); }
    const double den = __dmul_rn(ms, root);
    const double dy = __dsub_rn(__ddiv_rn(y2, den), __ddiv_rn(y1, den)), dx = __dsub_rn(__ddiv_rn(x2, den), __ddiv_rn(x1, den));
    const double d = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)));
    if (d > __dadd_rn(r1, r2)) return 0.0;
    if (d <= fabs(__dsub_rn(r1, r2))) return 1.0;
    return doh_disk_overlap(d, r1, r2);
}

struct DohFinishArgs {
    int n, S;
    DohParams P;
    double overlap;
    double* vals; uint32_t* idx; unsigned* count;       // [S][DOH_CAP], [S]
    double* blobs; int32_t* n_blobs;                     // [S][DOH_CAP][3] (row, col, sigma) in blob_doh's order, [S]
    double2* rc; unsigned ssc_cap; int32_t* n_kp;        // keypoints in adaptiveNMS order (stable argsort by sigma) for the SSC
    int32_t* status;
    const int32_t* flags;
};

#define DF_THREADS 1024
// dynamic shared memory: sv[DOH_CAP] f64 | si[DOH_CAP] u32 | alive[DOH_CAP] u8 | mark[DOH_CAP] u8
#define DF_SMEM (DOH_CAP * 8 + DOH_CAP * 4 + DOH_CAP * 2)
__global__ void __launch_bounds__(DF_THREADS) k_doh_finish(const DohFinishArgs a) {
    const int p = blockIdx.x;
    if (a.flags && !a.flags[p]) return;
    extern __shared__ unsigned char dsm[];
    double* sv = (double*)dsm;
    uint32_t* si = (uint32_t*)(sv + DOH_CAP);
    uint8_t* alive = (uint8_t*)(si + DOH_CAP);
    uint8_t* mark = alive + DOH_CAP;
    __shared__ int s_kill, s_cnt[DOH_MAX_SIGMA + 1];
    const unsigned tid = threadIdx.x;
    const unsigned total = a.count[p];
    const unsigned N = total < DOH_CAP ? total : DOH_CAP;
    const int NS = a.P.ns, n = a.n;
    unsigned np2 = 2;
    while (np2 < N) np2 <<= 1;
    const double* gv = a.vals + (size_t)p * DOH_CAP;
    const uint32_t* gi = a.idx + (size_t)p * DOH_CAP;
    for (unsigned i = tid; i < np2; i += DF_THREADS) {
        sv[i] = i < N ? gv[i] : -INFINITY;
        si[i] = i < N ? gi[i] : 0xffffffffu;
    }
    __syncthreads();
    // np.argsort(-intensities, kind stable over np.nonzero's C order): response descending, flat index ascending
    for (unsigned k = 2; k <= np2; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = tid; t < (np2 >> 1); t += DF_THREADS) {
                const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), q = i | j;
                const double x = sv[i], y = sv[q];
                const uint32_t xi = si[i], yi = si[q];
                const bool x_first = x > y || (x == y && xi < yi);
                if (x_first != ((i & k) == 0)) { sv[i] = y; sv[q] = x; si[i] = yi; si[q] = xi; }
            }
            __syncthreads();
        }
    for (unsigned i = tid; i < N; i += DF_THREADS) { alive[i] = 1; mark[i] = 0; }
    if (tid == 0) s_kill = 0x7fffffff;
    __syncthreads();
    // _prune_blobs in ascending (i, j) pair order.  A pair acts only while both blobs are alive (a zeroed sigma can
    // neither win a comparison nor lose again), and then it is judged on the original sigmas: the larger sigma
    // survives, on a tie blob j (the weaker response) survives.
    double smax = 0.0;
    for (int k = 0; k < NS; ++k) smax = fmax(smax, a.P.s[k].sigma);
    const double reach = 2.0 * smax * 1.4142135623730951 + 1.0;
    for (unsigned i = 0; i + 1 < N; ++i) {
        if (!alive[i]) continue;                          // uniform: written before the last barrier
        const uint32_t ei = si[i];
        const int ki = ei % NS; const uint32_t pi = ei / NS;
        const double yi = (double)(pi / n), xi = (double)(pi % n), sgi = a.P.s[ki].sigma;
        for (unsigned j = i + 1 + tid; j < N; j += DF_THREADS) {
            if (!alive[j]) continue;
            const uint32_t ej = si[j];
            const int kj = ej % NS; const uint32_t pj = ej / NS;
            const double yj = (double)(pj / n), xj = (double)(pj % n);
            if (fabs(yj - yi) > reach || fabs(xj - xi) > reach) continue;
            const double sgj = a.P.s[kj].sigma;
            if (doh_blob_overlap(yi, xi, sgi, yj, xj, sgj) > a.overlap) {
                if (sgi > sgj) mark[j] = 1;               // blob2[-1] = 0, if pair (i, j) is reached with i still alive
                else atomicMin(&s_kill, (int)j);          // blob1[-1] = 0 at the first such j
            }
        }
        __syncthreads();
        const int kill = s_kill;
        for (unsigned j = i + 1 + tid; j < N; j += DF_THREADS)
            if (mark[j]) { mark[j] = 0; if ((int)j < kill) alive[j] = 0; }
        __syncthreads();
        if (tid == 0) { if (kill != 0x7fffffff) alive[i] = 0; s_kill = 0x7fffffff; }
        __syncthreads();
    }
    // survivors in order (blob_doh's output), and their stable partition by scale (adaptiveNMS' argsort(blobs[:, 2]))
    if (tid <= DOH_MAX_SIGMA) s_cnt[tid] = 0;
    __syncthreads();
    // counts per scale, by one warp-strided pass with atomics (order does not matter for the counts)
    for (unsigned i = tid; i < N; i += DF_THREADS) if (alive[i]) atomicAdd(&s_cnt[si[i] % NS], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < NS; ++k) { const int c = s_cnt[k]; s_cnt[k] = acc; acc += c; }
        s_cnt[DOH_MAX_SIGMA] = acc;
    }
    __syncthreads();
    const int M = s_cnt[DOH_MAX_SIGMA];
    // ranks: a thread-0..31 warp walks the list in order (N is a few hundred; at most DOH_CAP)
    if (tid < 32) {
        int base_all = 0, base_k[DOH_MAX_SIGMA];
        for (int k = 0; k < DOH_MAX_SIGMA; ++k) base_k[k] = 0;
        double* blobs = a.blobs + (size_t)p * DOH_CAP * 3;
        double2* rc = a.rc + (size_t)p * a.ssc_cap;
        for (unsigned i0 = 0; i0 < N; i0 += 32) {
            const unsigned i = i0 + tid;
            const bool ok = i < N && alive[i];
            const uint32_t e = ok ? si[i] : 0u;
            const int k = (int)(e % NS); const uint32_t px = e / NS;
            const unsigned bm = __ballot_sync(0xffffffffu, ok);
            const unsigned before = bm & ((1u << tid) - 1u);
            int rank_k = 0;
            for (int kk = 0; kk < NS; ++kk) {
                const unsigned bk = __ballot_sync(0xffffffffu, ok && k == kk);
                if (ok && k == kk) rank_k = base_k[kk] + __popc(bk & ((1u << tid) - 1u));
                base_k[kk] += __popc(bk);
            }
            if (ok) {
                const int m = base_all + __popc(before);
                const double y = (double)(px / n), x = (double)(px % n);
                blobs[3 * m] = y; blobs[3 * m + 1] = x; blobs[3 * m + 2] = a.P.s[k].sigma;
                const unsigned pos = (unsigned)(s_cnt[k] + rank_k);
                if (a.rc && pos < a.ssc_cap) rc[pos] = make_double2(y, x);
            }
            base_all += __popc(bm);
        }
    }
    if (tid == 0) {
        a.n_blobs[p] = M;
        if (a.n_kp) a.n_kp[p] = (unsigned)M < a.ssc_cap ? M : (int)a.ssc_cap;
        if (a.status) a.status[p] = (total > DOH_CAP || (a.rc && (unsigned)M > a.ssc_cap)) ? RF_E_CAPACITY : RF_OK;
    }
}

// ---- workspace + launchers -----------------------------------------------------------------------------
struct DohWs { double* ii; size_t ii_stride; double* vals; uint32_t* idx; unsigned* count; double* blobs; int32_t* n_blobs; };

static inline size_t a256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t rf_doh_ws_bytes(const rf_handle* h, int S) {
    const size_t n2 = (size_t)h->n * h->n;
    return a256(S * n2 * 8) + a256((size_t)S * DOH_CAP * 8) + a256((size_t)S * DOH_CAP * 4) + a256((size_t)S * 4) +
           a256((size_t)S * DOH_CAP * 24) + a256((size_t)S * 4);
}

static DohWs doh_carve(const rf_handle* h, void* base, int S, int n) {
    DohWs w;
    char* p = (char*)base;
    const size_t n2 = (size_t)n * n;
    w.ii = (double*)p; w.ii_stride = n2; p += a256(S * n2 * 8);
    w.vals = (double*)p; p += a256((size_t)S * DOH_CAP * 8);
    w.idx = (uint32_t*)p; p += a256((size_t)S * DOH_CAP * 4);
    w.count = (unsigned*)p; p += a256((size_t)S * 4);
    w.blobs = (double*)p; p += a256((size_t)S * DOH_CAP * 24);
    w.n_blobs = (int32_t*)p;
    return w;
}

static int doh_prepare(rf_handle* h) {
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_doh_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM));
        attr_set[h->device & 63] = true;
    }
    return RF_OK;
}
int rf_doh_prepare(rf_handle* h) { return doh_prepare(h); }

static int doh_launch_integral(rf_handle* h, const DohWs& w, const float* d_cart, size_t cart_stride, int n, int S, const int32_t* d_flags) {
    dim3 g1((n + 127) / 128, S);
    k_doh_colsum<<<g1, 128, 0, h->stream>>>(d_cart, cart_stride, n, w.ii, w.ii_stride, d_flags);
    RF_CHECK_LAUNCH(h);
    dim3 g2((n + 31) / 32, S);
    k_doh_rowsum<<<g2, 256, 0, h->stream>>>(w.ii, w.ii_stride, n, d_flags);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

template <int NS>
static void doh_launch_peaks_ns(rf_handle* h, dim3 g, const DohWs& w, int n, const DohParams& P, double thr, const int32_t* d_flags) {
    k_doh_peaks<NS><<<g, 256, 0, h->stream>>>(w.ii, w.ii_stride, n, P, thr, w.vals, w.idx, w.count, d_flags);
}

// integral image -> peaks -> sort / prune -> blobs (+ SSC keypoints in ws_det->rc / n_kp when ws_det is given)
static int doh_launch_all(rf_handle* h, const DohWs& w, const float* d_cart, size_t cart_stride, int n, int S, const DohParams& P,
                          double thr, double overlap, const DetectWs* det, const int32_t* d_flags) {
    int rc = doh_prepare(h);
    if (rc) return rc;
    if ((rc = doh_launch_integral(h, w, d_cart, cart_stride, n, S, d_flags))) return rc;
    RF_CUDA(h, cudaMemsetAsync(w.count, 0, (size_t)S * 4, h->stream));
    const int tiles = ((n + DP_T - 1) / DP_T) * ((n + DP_T - 1) / DP_T);
    dim3 g(rf_tile_workers(h, tiles, S), S);
    switch (P.ns) {
        case 1: doh_launch_peaks_ns<1>(h, g, w, n, P, thr, d_flags); break;
        case 2: doh_launch_peaks_ns<2>(h, g, w, n, P, thr, d_flags); break;
        case 3: doh_launch_peaks_ns<3>(h, g, w, n, P, thr, d_flags); break;
        case 4: doh_launch_peaks_ns<4>(h, g, w, n, P, thr, d_flags); break;
        case 5: doh_launch_peaks_ns<5>(h, g, w, n, P, thr, d_flags); break;
        case 6: doh_launch_peaks_ns<6>(h, g, w, n, P, thr, d_flags); break;
        case 7: doh_launch_peaks_ns<7>(h, g, w, n, P, thr, d_flags); break;
        case 8: doh_launch_peaks_ns<8>(h, g, w, n, P, thr, d_flags); break;
        case 9: doh_launch_peaks_ns<9>(h, g, w, n, P, thr, d_flags); break;
        case 10: doh_launch_peaks_ns<10>(h, g, w, n, P, thr, d_flags); break;
        case 11: doh_launch_peaks_ns<11>(h, g, w, n, P, thr, d_flags); break;
        case 12: doh_launch_peaks_ns<12>(h, g, w, n, P, thr, d_flags); break;
        case 13: doh_launch_peaks_ns<13>(h, g, w, n, P, thr, d_flags); break;
        case 14: doh_launch_peaks_ns<14>(h, g, w, n, P, thr, d_flags); break;
        case 15: doh_launch_peaks_ns<15>(h, g, w, n, P, thr, d_flags); break;
        default: doh_launch_peaks_ns<16>(h, g, w, n, P, thr, d_flags); break;
    }
    RF_CHECK_LAUNCH(h);
    DohFinishArgs a;
    a.n = n; a.S = S; a.P = P; a.overlap = overlap; a.vals = w.vals; a.idx = w.idx; a.count = w.count;
    a.blobs = w.blobs; a.n_blobs = w.n_blobs;
    a.rc = det ? det->rc : nullptr; a.ssc_cap = det ? det->ssc_cap : 0; a.n_kp = det ? det->n_kp : nullptr;
    a.status = det ? det->status : nullptr; a.flags = d_flags;
    k_doh_finish<<<S, DF_THREADS, DF_SMEM, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// lock-step runner (k_seq.cu): blobs of every flagged sequence -> det.rc / det.n_kp (adaptiveNMS order), det.count = peaks
int rf_launch_doh_keypoints(rf_handle* h, const DetectWs& det, const float* d_cart, size_t cart_stride, int n, void* d_doh_ws,
                            const int32_t* d_flags) {
    DohParams P;
    const rf_config& c = h->cfg;
    int rc = doh_make_params(h, c.doh_min_sigma, c.doh_max_sigma, c.doh_num_sigma, &P);
    if (rc) return rc;
    DohWs w = doh_carve(h, d_doh_ws, det.S, n);
    w.count = det.count;                 // n_candidates of the step record
    return doh_launch_all(h, w, d_cart, cart_stride, n, det.S, P, c.doh_threshold, 0.5, &det, d_flags);
}

extern "C" {

int rf_detect_doh(rf_handle* h, const rf_frame* f, double min_sigma, double max_sigma, int num_sigma, double threshold,
                  double overlap, double* out, int cap, int* n_out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !f || !n_out || cap < 0 || (cap > 0 && !out)) return rf_fail(h, RF_E_BADARG, "rf_detect_doh: bad argument");
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_detect_doh: frame has no f32 plane");
    cudaSetDevice(h->device);
    DohParams P;
    int rc = doh_make_params(h, min_sigma, max_sigma, num_sigma, &P);
    if (rc) return rc;
    if ((rc = rf_ensure_scratch(h, rf_doh_ws_bytes(h, 1)))) return rc;
    DohWs w = doh_carve(h, h->d_scratch, 1, h->n);
    if ((rc = doh_launch_all(h, w, f->fs.cart, 0, h->n, 1, P, threshold, overlap, nullptr, nullptr))) return rc;
    int32_t m = 0; unsigned peaks = 0;
    RF_CUDA(h, cudaMemcpyAsync(&m, w.n_blobs, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&peaks, w.count, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (peaks > DOH_CAP) return rf_fail(h, RF_E_CAPACITY, "rf_detect_doh: %u peaks exceed the capacity %d (raise the threshold)", peaks, DOH_CAP);
    *n_out = m;
    const int take = m < cap ? m : cap;
    if (take) {
        RF_CUDA(h, cudaMemcpyAsync(out, w.blobs, (size_t)take * 24, cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RF_OK;
}

// Test hook: the float64 integral image (sigma_index < 0) or the determinant-of-Hessian plane of scale `sigma_index`
int rf_doh_response(rf_handle* h, const rf_frame* f, double min_sigma, double max_sigma, int num_sigma, int sigma_index, double* out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !f || !out) return rf_fail(h, RF_E_BADARG, "rf_doh_response: null argument");
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_doh_response: frame has no f32 plane");
    cudaSetDevice(h->device);
    DohParams P;
    int rc = doh_make_params(h, min_sigma, max_sigma, num_sigma, &P);
    if (rc) return rc;
    if (sigma_index >= num_sigma) return rf_fail(h, RF_E_BADARG, "rf_doh_response: sigma_index %d >= num_sigma %d", sigma_index, num_sigma);
    const int n = h->n;
    const size_t n2 = (size_t)n * n;
    if ((rc = rf_ensure_scratch(h, rf_doh_ws_bytes(h, 1) + a256(n2 * 8)))) return rc;
    DohWs w = doh_carve(h, h->d_scratch, 1, n);
    double* d_plane = (double*)((char*)h->d_scratch + rf_doh_ws_bytes(h, 1));
    if ((rc = doh_launch_integral(h, w, f->fs.cart, 0, n, 1, nullptr))) return rc;
    const double* src = w.ii;
    if (sigma_index >= 0) {
        dim3 g((n + 31) / 32, (n + 7) / 8);
        k_doh_plane<<<g, 256, 0, h->stream>>>(w.ii, n, P.s[sigma_index], d_plane);
        RF_CHECK_LAUNCH(h);
        src = d_plane;
    }
    RF_CUDA(h, cudaMemcpyAsync(out, src, n2 * 8, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

}  // extern "C"

// rf_detect(mode 1): the handle's configured detector parameters (getFeatures.DEFAULT_FEATURE_PARAMS); a negative
// threshold argument selects the configured one
int rf_doh_detect_host(rf_handle* h, const rf_frame* f, float threshold, double* out, int cap, int* n_out) {
    const rf_config& c = h->cfg;
    return rf_detect_doh(h, f, c.doh_min_sigma, c.doh_max_sigma, c.doh_num_sigma, threshold < 0 ? c.doh_threshold : (double)threshold, 0.5,
                         out, cap, n_out);
}
