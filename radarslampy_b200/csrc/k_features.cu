// k_features.cu — feature selection: corner response, 3x3 non-maximum suppression,
// exact SSC adaptive NMS, and the per-azimuth polar peak extractor.
//
// Replaces (reference file:line):
//   ANMS.py:5-102            ssc (Suppression via Square Covering)        -> k_ssc_pass (+ host binary search)
//   getFeatures.py:22-53     getBlobsFromCart (detector front half)        -> k_min_eig + k_nms_select
//   getPointCloud.py:11-54   getPointCloudPolarInd                         -> k_peaks_rows + k_peaks_gather
//
// The reference's detector is skimage.feature.blob_doh, which is not available to pin against
// (parity unpinned).  Mode 0 here is the Sobel / structure-tensor minimum
// eigenvalue named by BASELINE.json's north_star, with cv2.cornerMinEigenVal(blockSize=3,
// ksize=3, BORDER_REFLECT_101) semantics on the f32 Cartesian image; candidate selection
// (threshold + 3x3 NMS + ordering) and SSC are bit-exact against the oracle.
#include <cub/device/device_radix_sort.cuh>
#include <math.h>

#include "common.cuh"

// =====================================================================================
// a9  SSC.  One pass of the greedy square covering for one width: keypoints are visited
// in input order; a keypoint whose cell is uncovered is selected and covers the cells
// within +-reach.  The cover grid is a bitset (shared memory when it fits, else global).
// One warp walks the list 32 keypoints at a time: coverage only ever grows, so every
// covered keypoint of the group is rejected at once and only SELECTED keypoints cost a
// serial step.
// =====================================================================================
struct SscPass {
    const double* kp;   // [n][3] (row, col, sigma)
    int n;
    double c;           // cell size = width / 2
    int reach;          // floor(width / c)
    int ncr, ncc;       // floor(rows / c), floor(cols / c)
    uint32_t* gcov;     // global bitset (pre-cleared) or nullptr -> shared memory
    unsigned words;
    int32_t* sel;       // [n] selected indices in selection order
    int32_t* nres;      // [1]
};

__global__ void __launch_bounds__(256) k_ssc_pass(const SscPass a) {
    extern __shared__ uint32_t s_cov[];
    uint32_t* cov = a.gcov ? a.gcov : s_cov;
    if (!a.gcov) {
        for (unsigned i = threadIdx.x; i < a.words; i += blockDim.x) s_cov[i] = 0u;
        __syncthreads();
    }
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const long long stride = (long long)a.ncc + 1;
    int nres = 0;
    for (int base = 0; base < a.n; base += 32) {
        const int i = base + lane;
        const bool valid = i < a.n;
        int row = 0, col = 0;
        if (valid) {
            row = (int)floor(a.kp[3 * (size_t)i] / a.c);       // ANMS.py:52-59
            col = (int)floor(a.kp[3 * (size_t)i + 1] / a.c);
        }
        const long long bit = (long long)row * stride + col;
        unsigned pending = __ballot_sync(0xffffffffu, valid);
        while (pending) {
            bool unc = false;
            if ((pending >> lane) & 1u) unc = !((cov[bit >> 5] >> (bit & 31)) & 1u);
            const unsigned um = __ballot_sync(0xffffffffu, unc);
            if (!um) break;
            const int j = __ffs(um) - 1;
            const int r = __shfl_sync(0xffffffffu, row, j), cc = __shfl_sync(0xffffffffu, col, j);
            if (lane == 0) a.sel[nres] = base + j;
            ++nres;
            const int r0 = max(r - a.reach, 0), r1 = min(r + a.reach, a.ncr);     // ANMS.py:63-82
            const int c0 = max(cc - a.reach, 0), c1 = min(cc + a.reach, a.ncc);
            for (int rr = r0 + lane; rr <= r1; rr += 32) {
                const long long b0 = (long long)rr * stride + c0, b1 = (long long)rr * stride + c1;
                for (long long w = b0 >> 5; w <= (b1 >> 5); ++w) {
                    const int lo = (w == (b0 >> 5)) ? (int)(b0 & 31) : 0;
                    const int hi = (w == (b1 >> 5)) ? (int)(b1 & 31) : 31;
                    const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
                    atomicOr(&cov[w], m);
                }
            }
            __syncwarp();
            pending = um & ~((2u << j) - 1u);
        }
    }
    if (lane == 0) *a.nres = nres;
}

// Same pass without a grid, for cell sizes so small that the cover grid would not fit in memory (the
// reference would build a list of lists of that size).  A keypoint is covered iff an earlier SELECTED keypoint
// lies within +-reach cells in both axes (the clamping at the grid edges never changes that test).  Cell
// coordinates are kept as doubles (integer-valued, exact).
struct SscSparse {
    const double* kp; int n; double c; double reach;
    double* srow; double* scol;   // [n] cell coordinates of the selected keypoints
    int32_t* sel; int32_t* nres;
};
__global__ void __launch_bounds__(32) k_ssc_pass_sparse(const SscSparse a) {
    const int lane = threadIdx.x;
    int nres = 0;
    for (int i = 0; i < a.n; ++i) {
        const double row = floor(a.kp[3 * (size_t)i] / a.c), col = floor(a.kp[3 * (size_t)i + 1] / a.c);
        bool cov = false;
        for (int j = lane; j < nres && !cov; j += 32)
            cov = fabs(a.srow[j] - row) <= a.reach && fabs(a.scol[j] - col) <= a.reach;
        if (!__any_sync(0xffffffffu, cov)) {
            if (lane == 0) { a.srow[nres] = row; a.scol[nres] = col; a.sel[nres] = i; }
            ++nres;
            __syncwarp();
        }
    }
    if (lane == 0) *a.nres = nres;
}

static double py_round_half_even(double x) { return nearbyint(x); }  // Python round() on a float

#define SSC_SMEM_MAX (200 * 1024)

extern "C" int rf_ssc(rf_handle* h, const double* kp, int n, int num_ret, double tol, int cols, int rows,
                      int32_t* sel_idx, int* m) {
    if (!h || !m || n < 0 || (n > 0 && (!kp || !sel_idx)) || cols <= 0 || rows <= 0)
        return rf_fail(h, RF_E_BADARG, "rf_ssc: bad argument");
    if (num_ret == 1) return rf_fail(h, RF_E_BADARG, "rf_ssc: num_ret_points == 1 divides by zero (ANMS.py:19-22)");
    *m = 0;
    if (n == 0) return RF_OK;      // nothing can be selected (the reference still walks its binary search)
    for (int i = 0; i < n; ++i) {  // the reference indexes covered_vec[row][col]: anything outside raises there
        const double r = kp[3 * (size_t)i], c = kp[3 * (size_t)i + 1];
        if (!(r >= 0 && r <= rows && c >= 0 && c <= cols))
            return rf_fail(h, RF_E_BADARG, "rf_ssc: keypoint %d (%g, %g) lies outside the %d x %d image", i, r, c, rows, cols);
    }
    // ANMS.py:6-35 — closed-form upper bound of the search range (Python ints/floats restated in double)
    const double exp1 = (double)rows + cols + 2.0 * num_ret;
    const double exp2 = 4.0 * cols + 4.0 * num_ret + 4.0 * rows * (double)num_ret + (double)rows * rows +
                        (double)cols * cols - 2.0 * rows * (double)cols + 4.0 * (double)rows * cols * num_ret;
    if (exp2 < 0) return rf_fail(h, RF_E_BADARG, "rf_ssc: math domain error (ANMS.py:17)");
    const double exp3 = sqrt(exp2), exp4 = num_ret - 1;
    const double sol1 = -py_round_half_even((exp1 + exp3) / exp4), sol2 = -py_round_half_even((exp1 - exp3) / exp4);
    double high = sol1 > sol2 ? sol1 : sol2;
    if (num_ret <= 0 || (double)n / num_ret < 0) return rf_fail(h, RF_E_BADARG, "rf_ssc: num_ret_points must be positive");
    double low = floor(sqrt((double)n / num_ret));
    const double kmin = py_round_half_even(num_ret - num_ret * tol), kmax = py_round_half_even(num_ret + num_ret * tol);
    double prev_width = -1;

    const size_t kp_bytes = ((size_t)n * 3 * sizeof(double) + 255) & ~(size_t)255;
    const size_t sel_bytes = ((size_t)(n > 0 ? n : 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    size_t need = kp_bytes + sel_bytes + 256;
    int rc = rf_ensure_scratch(h, need);
    if (rc) return rc;
    static bool attr_set[64] = {};   // per device: the attribute belongs to the device's context
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_ssc_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, SSC_SMEM_MAX));
        attr_set[h->device & 63] = true;
    }
    if (n) RF_CUDA(h, cudaMemcpyAsync(h->d_scratch, kp, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int nres = 0;           // len(result) of the most recent pass
    bool have = false;      // a pass has run (its result sits in the device sel buffer)
    for (;;) {
        const double width = low + (high - low) / 2;                    // ANMS.py:38
        if (width == prev_width || low > high) break;                   // -> result of the previous pass
        const double c = width / 2;
        if (!(c != 0)) return rf_fail(h, RF_E_BADARG, "rf_ssc: zero cell size (the reference divides by zero, ANMS.py:45)");
        const double fcc = floor(cols / c), fcr = floor(rows / c);
        if (!(fcc >= 0 && fcr >= 0)) return rf_fail(h, RF_E_BADARG, "rf_ssc: negative cell size (ANMS.py:45-47)");
        if ((fcc + 1) * (fcr + 1) > 1073741824.0) {   // > 2^30 cells: grid-free pass
            const size_t co_bytes = ((size_t)n * sizeof(double) + 255) & ~(size_t)255;
            const uint64_t before = h->scratch_gen;
            if ((rc = rf_ensure_scratch(h, need + 2 * co_bytes))) return rc;
            if (h->scratch_gen != before)
                RF_CUDA(h, cudaMemcpyAsync(h->d_scratch, kp, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            char* base = (char*)h->d_scratch;
            SscSparse sp;
            sp.kp = (const double*)base; sp.n = n; sp.c = c; sp.reach = floor(width / c);
            sp.sel = (int32_t*)(base + kp_bytes); sp.nres = (int32_t*)(base + kp_bytes + sel_bytes);
            sp.srow = (double*)(base + need); sp.scol = (double*)(base + need + co_bytes);
            k_ssc_pass_sparse<<<1, 32, 0, h->stream>>>(sp);
            RF_CHECK_LAUNCH(h);
            RF_CUDA(h, cudaMemcpyAsync(&nres, sp.nres, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            RF_CUDA(h, cudaStreamSynchronize(h->stream));
            have = true;
            if (kmin <= nres && nres <= kmax) break;
            if (nres < kmin) high = width - 1; else low = width + 1;
            prev_width = width;
            continue;
        }
        SscPass a;
        a.n = n; a.c = c; a.reach = (int)floor(width / c); a.ncr = (int)fcr; a.ncc = (int)fcc;
        const unsigned long long bits = (unsigned long long)(a.ncr + 1) * (unsigned long long)(a.ncc + 1);
        a.words = (unsigned)((bits + 31) / 32);
        const size_t cov_bytes = (size_t)a.words * 4;
        size_t smem = 0;
        if (cov_bytes <= SSC_SMEM_MAX) {
            smem = cov_bytes;
        } else {
            const uint64_t before = h->scratch_gen;
            if ((rc = rf_ensure_scratch(h, need + cov_bytes))) return rc;
            if (h->scratch_gen != before) {   // the arena moved: stage the keypoints again (earlier results are obsolete)
                if (n) RF_CUDA(h, cudaMemcpyAsync(h->d_scratch, kp, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            }
        }
        char* base = (char*)h->d_scratch;
        a.kp = (const double*)base;
        a.sel = (int32_t*)(base + kp_bytes);
        a.nres = (int32_t*)(base + kp_bytes + sel_bytes);
        a.gcov = nullptr;
        if (!smem) {
            a.gcov = (uint32_t*)(base + need);
            RF_CUDA(h, cudaMemsetAsync(a.gcov, 0, cov_bytes, h->stream));
        }
        k_ssc_pass<<<1, 256, smem, h->stream>>>(a);
        RF_CHECK_LAUNCH(h);
        RF_CUDA(h, cudaMemcpyAsync(&nres, a.nres, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
        have = true;
        if (kmin <= nres && nres <= kmax) break;                        // ANMS.py:89-91
        if (nres < kmin) high = width - 1; else low = width + 1;        // ANMS.py:92-95
        prev_width = width;
    }
    if (have && nres > 0) {
        RF_CUDA(h, cudaMemcpyAsync(sel_idx, (char*)h->d_scratch + kp_bytes, (size_t)nres * sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    *m = have ? nres : 0;
    return RF_OK;
}

// =====================================================================================
// a10 corner response: minimum eigenvalue of the 3x3-summed structure tensor of the
// scaled 3x3 Sobel gradients (cv::cornerMinEigenVal, f32 input: scale = 1 / (4 * 3)).
// Block = 32 x 8 threads -> 32 x 32 outputs; the image tile (+2 halo) and the gradient
// products (+1 halo, evaluated at REFLECT_101 positions of the gradient image like
// cv::boxFilter does) are staged in shared memory.
// =====================================================================================
#define ME_TW 32
#define ME_TH 32
__global__ void __launch_bounds__(256)
k_min_eig(const float* __restrict__ img, int n, float k0, float k1, float* __restrict__ resp) {
    __shared__ float tile[ME_TH + 4][ME_TW + 4 + 1];
    __shared__ float gxx[ME_TH + 2][ME_TW + 2 + 1], gxy[ME_TH + 2][ME_TW + 2 + 1], gyy[ME_TH + 2][ME_TW + 2 + 1];
    const int ox = blockIdx.x * ME_TW, oy = blockIdx.y * ME_TH;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    // tile position (r, c) holds img at clamp-reflected (oy - 2 + r, ox - 2 + c); positions are looked up
    // through tpos() so gradients at reflected locations read the taps a full-image Sobel would read
    for (int i = tid; i < (ME_TH + 4) * (ME_TW + 4); i += 256) {
        const int r = i / (ME_TW + 4), c = i - r * (ME_TW + 4);
        int y = oy - 2 + r, x = ox - 2 + c;
        y = min(max(y, -2), n + 1); x = min(max(x, -2), n + 1);
        tile[r][c] = __ldg(img + (size_t)reflect101(y, n) * n + reflect101(x, n));
    }
    __syncthreads();
    for (int i = tid; i < (ME_TH + 2) * (ME_TW + 2); i += 256) {
        const int r = i / (ME_TW + 2), c = i - r * (ME_TW + 2);
        // gradient-image position, reflected into the image (cv::boxFilter border on the gradient products)
        int y = oy - 1 + r, x = ox - 1 + c;
        y = min(max(y, -1), n); x = min(max(x, -1), n);
        const int qy = reflect101(y, n), qx = reflect101(x, n);
        // taps of the full-image Sobel at (qy, qx): neighbours at reflect101(q +- 1)
        const int ym = reflect101(qy - 1, n), yp = reflect101(qy + 1, n);
        const int xm = reflect101(qx - 1, n), xp = reflect101(qx + 1, n);
        // every index is within [oy-2, oy+TH+1] x [ox-2, ox+TW+1] *after reflection* only when the reflected
        // coordinate stays inside the tile; near borders it does (|q - p| <= 2), else q == p
#define T(yy, xx) tile[(yy) - (oy - 2)][(xx) - (ox - 2)]
        const float a00 = T(ym, xm), a01 = T(ym, qx), a02 = T(ym, xp);
        const float a10 = T(qy, xm), a12 = T(qy, xp);
        const float a20 = T(yp, xm), a21 = T(yp, qx), a22 = T(yp, xp);
#undef T
        // Dx: row filter [-1 0 1], column filter [k1 k0 k1] (the smoothing taps carry the scale)
        const float rx0 = __fsub_rn(a02, a00), rx1 = __fsub_rn(a12, a10), rx2 = __fsub_rn(a22, a20);
        const float dx = __fmaf_rn(k1, __fadd_rn(rx0, rx2), __fmul_rn(k0, rx1));
        // Dy: row filter [k1 k0 k1], column filter [-1 0 1]
        const float ry0 = __fmaf_rn(k1, __fadd_rn(a00, a02), __fmul_rn(k0, a01));
        const float ry2 = __fmaf_rn(k1, __fadd_rn(a20, a22), __fmul_rn(k0, a21));
        const float dy = __fsub_rn(ry2, ry0);
        gxx[r][c] = __fmul_rn(dx, dx); gxy[r][c] = __fmul_rn(dx, dy); gyy[r][c] = __fmul_rn(dy, dy);
    }
    __syncthreads();
    for (int k = 0; k < ME_TH / 8; ++k) {
        const int r = threadIdx.y + 8 * k, c = threadIdx.x;
        const int y = oy + r, x = ox + c;
        if (y >= n || x >= n) continue;
        double sxx = 0, sxy = 0, syy = 0;  // cv::boxFilter accumulates f32 planes in double
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                sxx += (double)gxx[r + dy][c + dx]; sxy += (double)gxy[r + dy][c + dx]; syy += (double)gyy[r + dy][c + dx];
            }
        const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, cc = __fmul_rn((float)syy, 0.5f);
        const float d = __fsub_rn(a, cc);
        resp[(size_t)y * n + x] = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fmaf_rn(d, d, __fmul_rn(b, b))));
    }
}

// threshold + 3x3 non-maximum suppression (cv::goodFeaturesToTrack's rule: interior pixels with
// resp > thr and resp == max of the 3x3 neighbourhood).  Emits sortable 64-bit keys:
// high word = ~bits(resp) (so ascending key = descending response), low word = ~pixel index
// (ties: descending index, the order cv::goodFeaturesToTrack's pointer comparison produces).
__global__ void __launch_bounds__(256)
k_nms_select(const float* __restrict__ resp, int rows, int cols, float thr, unsigned long long* __restrict__ keys,
             unsigned cap, unsigned* __restrict__ count) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < 1 || y < 1 || x >= cols - 1 || y >= rows - 1) return;
    const float v = __ldg(resp + (size_t)y * cols + x);
    if (!(v > thr)) return;
    float mx = v;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) mx = fmaxf(mx, __ldg(resp + (size_t)(y + dy) * cols + (x + dx)));
    if (v != mx) return;
    const unsigned slot = atomicAdd(count, 1u);
    if (slot < cap) {
        const unsigned idx = (unsigned)y * (unsigned)cols + (unsigned)x;
        keys[slot] = ((unsigned long long)(~__float_as_uint(v)) << 32) | (unsigned long long)(~idx);
    }
}

__global__ void __launch_bounds__(256)
k_keys_to_rows(const unsigned long long* __restrict__ keys, unsigned n, int cols, double* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const unsigned idx = ~(unsigned)(k & 0xffffffffull);
    const float v = __uint_as_float(~(unsigned)(k >> 32));
    out[3 * (size_t)i] = (double)(idx / (unsigned)cols);
    out[3 * (size_t)i + 1] = (double)(idx % (unsigned)cols);
    out[3 * (size_t)i + 2] = (double)v;
}

// maximum of a non-negative response map (bit pattern order == value order for floats >= 0)
__global__ void __launch_bounds__(256) k_max_resp(const float* __restrict__ resp, size_t count, unsigned* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, __ldg(resp + i));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

static int launch_min_eig(rf_handle* h, const float* d_img, int n, float* d_resp) {
    const double scale = 1.0 / ((double)(1 << 2) * 3.0);  // ksize 3, blockSize 3, f32 input
    dim3 blk(32, 8), grd((n + ME_TW - 1) / ME_TW, (n + ME_TH - 1) / ME_TH);
    k_min_eig<<<grd, blk, 0, h->stream>>>(d_img, n, (float)(2.0 * scale), (float)(1.0 * scale), d_resp);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// select + sort on a device response map; out (host) receives min(*n, cap) rows
static int select_sorted(rf_handle* h, const float* d_resp, int rows, int cols, float thr, char* ws, size_t ws_bytes,
                         unsigned key_cap, double* out, int cap, int* n) {
    // workspace layout: count | keys_in[key_cap] | keys_out[key_cap] | rows[cap*3] | cub temp
    unsigned* d_count = (unsigned*)ws;
    unsigned long long* k_in = (unsigned long long*)(ws + 256);
    unsigned long long* k_out = k_in + key_cap;
    double* d_rows = (double*)(k_out + key_cap);
    char* d_tmp = (char*)(d_rows + (size_t)cap * 3);
    size_t tmp_bytes = ws_bytes - (size_t)(d_tmp - ws);
    RF_CUDA(h, cudaMemsetAsync(d_count, 0, sizeof(unsigned), h->stream));
    dim3 blk(32, 8), grd((cols + 31) / 32, (rows + 7) / 8);
    k_nms_select<<<grd, blk, 0, h->stream>>>(d_resp, rows, cols, thr, k_in, key_cap, d_count);
    RF_CHECK_LAUNCH(h);
    unsigned cnt = 0;
    RF_CUDA(h, cudaMemcpyAsync(&cnt, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (cnt > key_cap)
        return rf_fail(h, RF_E_CAPACITY, "rf_detect: %u candidates exceed the key capacity %u (raise the threshold)", cnt, key_cap);
    *n = (int)cnt;
    if (!cnt) return RF_OK;
    RF_CUDA(h, cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, k_in, k_out, (int)cnt, 0, 64, h->stream));
    h->launches += 1;
    const unsigned take = cnt < (unsigned)cap ? cnt : (unsigned)cap;
    if (take) {
        k_keys_to_rows<<<(take + 255) / 256, 256, 0, h->stream>>>(k_out, take, cols, d_rows);
        RF_CHECK_LAUNCH(h);
        RF_CUDA(h, cudaMemcpyAsync(out, d_rows, (size_t)take * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

static int detect_ws(rf_handle* h, size_t resp_bytes, unsigned key_cap, int cap, size_t* ws_off, size_t* ws_bytes) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)key_cap, 0, 64,
                                   h->stream);
    const size_t ws = 256 + (size_t)key_cap * 16 + (size_t)cap * 24 + tmp + 256;
    *ws_off = (resp_bytes + 255) & ~(size_t)255;
    *ws_bytes = ws;
    return rf_ensure_scratch(h, *ws_off + ws);
}

extern "C" {

int rf_corner_response(rf_handle* h, const rf_frame* f, int mode, float* resp) {
    if (!h || !f || !resp) return rf_fail(h, RF_E_BADARG, "rf_corner_response: null argument");
    if (mode != 0) return rf_fail(h, RF_E_BADARG, "rf_corner_response: mode %d is not available (0 = structure-tensor min eigenvalue)", mode);
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_corner_response: frame has no f32 plane");
    const size_t bytes = (size_t)h->n * h->n * sizeof(float);
    int rc = rf_ensure_scratch(h, bytes);
    if (rc) return rc;
    if ((rc = launch_min_eig(h, f->fs.cart, h->n, (float*)h->d_scratch))) return rc;
    RF_CUDA(h, cudaMemcpyAsync(resp, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

int rf_detect(rf_handle* h, const rf_frame* f, int mode, float threshold, double* out, int cap, int* n) {
    if (!h || !f || !n || cap < 0 || (cap > 0 && !out)) return rf_fail(h, RF_E_BADARG, "rf_detect: bad argument");
    if (mode != 0) return rf_fail(h, RF_E_BADARG, "rf_detect: mode %d is not available (0 = structure-tensor min eigenvalue)", mode);
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_detect: frame has no f32 plane");
    const size_t bytes = (size_t)h->n * h->n * sizeof(float);
    const unsigned key_cap = (unsigned)((size_t)h->n * h->n / 4 + 1024);   // 3x3 maxima cannot be denser than 1 in 4
    size_t off, ws;
    int rc = detect_ws(h, bytes, key_cap, cap, &off, &ws);
    if (rc) return rc;
    float* d_resp = (float*)h->d_scratch;
    if ((rc = launch_min_eig(h, f->fs.cart, h->n, d_resp))) return rc;
    if (threshold < 0) {   // relative: -threshold is cv2.goodFeaturesToTrack's qualityLevel (fraction of the maximum)
        unsigned* d_max = (unsigned*)((char*)h->d_scratch + off);
        unsigned bits = 0;
        RF_CUDA(h, cudaMemsetAsync(d_max, 0, sizeof(unsigned), h->stream));
        k_max_resp<<<h->sm_count * 4, 256, 0, h->stream>>>(d_resp, (size_t)h->n * h->n, d_max);
        RF_CHECK_LAUNCH(h);
        RF_CUDA(h, cudaMemcpyAsync(&bits, d_max, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
        float mx; memcpy(&mx, &bits, 4);
        threshold = (float)((double)mx * (double)(-threshold));   // cv: threshold(eig, eig, maxVal * qualityLevel, ...)
    }
    return select_sorted(h, d_resp, h->n, h->n, threshold, (char*)h->d_scratch + off, ws, key_cap, out, cap, n);
}

int rf_nms_select(rf_handle* h, const float* resp, int rows, int cols, float threshold, double* out, int cap, int* n) {
    if (!h || !resp || !n || rows < 3 || cols < 3 || cap < 0 || (cap > 0 && !out))
        return rf_fail(h, RF_E_BADARG, "rf_nms_select: bad argument");
    const size_t bytes = (size_t)rows * cols * sizeof(float);
    const unsigned key_cap = (unsigned)((size_t)rows * cols + 1024);        // plateaus may keep every pixel
    size_t off, ws;
    int rc = detect_ws(h, bytes, key_cap, cap, &off, &ws);
    if (rc) return rc;
    RF_CUDA(h, cudaMemcpyAsync(h->d_scratch, resp, bytes, cudaMemcpyHostToDevice, h->stream));
    return select_sorted(h, (const float*)h->d_scratch, rows, cols, threshold, (char*)h->d_scratch + off, ws, key_cap, out, cap, n);
}

}  // extern "C"

// =====================================================================================
// a12 getPointCloudPolarInd: one CTA per azimuth.  scipy.signal.find_peaks without
// constraints = strict local maxima with plateaus reported at their middle sample; then
// keep peaks >= mean + std of the peak heights, with NumPy's f32 pairwise summation.
// =====================================================================================
__device__ float np_pairwise_sum_f32(const float* a, int n) {
    // numpy/core/src/umath/loops_utils.h.src  pairwise sum, restated (see oracle/c/oracle_c.c)
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum_f32(a, n2), np_pairwise_sum_f32(a + n2, n - n2));
}

__device__ __forceinline__ int block_excl_scan_256(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane < 8) s_warp[lane] = x;
    }
    __syncthreads();
    const int base = w ? s_warp[w - 1] : 0;
    *total = s_warp[7];
    __syncthreads();
    return base + inc - v;
}

// peak at plateau [i, ahead-1] if x[i-1] < x[i] and the first differing sample to the right is lower
__device__ __forceinline__ int peak_at(const float* x, int i, int imax) {
    if (!(x[i - 1] < x[i])) return -1;
    int ahead = i + 1;
    while (ahead < imax && x[ahead] == x[i]) ++ahead;
    return (x[ahead] < x[i]) ? (i + ahead - 1) / 2 : -1;
}

__global__ void __launch_bounds__(256)
k_peaks_rows(const float* __restrict__ polar, int W, int32_t* __restrict__ row_peaks, int row_cap,
             int32_t* __restrict__ row_count) {
    extern __shared__ float s_row[];           // [W] samples, then [W/2+1] peak heights, then [W/2+1] peak positions
    float* x = s_row;
    float* hv = s_row + W;
    int* pk = (int*)(hv + (W / 2 + 1));
    __shared__ int s_warp[8];
    __shared__ float s_thr;
    const int a = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < W; i += 256) x[i] = __ldg(polar + (size_t)a * W + i);
    __syncthreads();
    const int imax = W - 1;
    const int chunk = (W + 255) / 256;
    const int i0 = max(tid * chunk, 1), i1 = min((tid + 1) * chunk, imax);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += peak_at(x, i, imax) >= 0;
    int np_;
    int off = block_excl_scan_256(cnt, s_warp, &np_);
    for (int i = i0; i < i1; ++i) {
        const int p = peak_at(x, i, imax);
        if (p >= 0) { pk[off] = p; hv[off] = x[p]; ++off; }
    }
    __syncthreads();
    if (np_ == 0) { if (tid == 0) row_count[a] = 0; return; }
    if (tid == 0) {
        const float mean = __fdiv_rn(np_pairwise_sum_f32(hv, np_), (float)np_);    // np.mean (f32)
        s_thr = mean;
    }
    __syncthreads();
    const float mean = s_thr;
    __syncthreads();
    for (int k = tid; k < np_; k += 256) { const float d = __fsub_rn(hv[k], mean); hv[k] = __fmul_rn(d, d); }
    __syncthreads();
    if (tid == 0) {
        const float var = __fdiv_rn(np_pairwise_sum_f32(hv, np_), (float)np_);     // np.std (population)
        s_thr = __fadd_rn(mean, __fsqrt_rn(var));                                  // getPointCloud.py:44
    }
    __syncthreads();
    const float thr = s_thr;
    const int pchunk = (np_ + 255) / 256;
    const int k0 = min(tid * pchunk, np_), k1 = min((tid + 1) * pchunk, np_);
    int keep = 0;
    for (int k = k0; k < k1; ++k) keep += x[pk[k]] >= thr;
    int total;
    int o = block_excl_scan_256(keep, s_warp, &total);
    for (int k = k0; k < k1; ++k)
        if (x[pk[k]] >= thr) { if (o < row_cap) row_peaks[(size_t)a * row_cap + o] = pk[k]; ++o; }
    if (tid == 0) row_count[a] = total;
}

// azimuth-major concatenation (np.vstack per azimuth, getPointCloud.py:50-52)
__global__ void __launch_bounds__(256)
k_peaks_gather(const int32_t* __restrict__ row_peaks, int row_cap, const int32_t* __restrict__ row_count, int A,
               int64_t* __restrict__ out, long long cap, long long* __restrict__ n_out) {
    __shared__ int s_warp[8];
    __shared__ long long s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int a0 = 0; a0 < A; a0 += 256) {
        const int a = a0 + threadIdx.x;
        const int c = a < A ? row_count[a] : 0;
        int total;
        const int off = block_excl_scan_256(c, s_warp, &total);
        const long long base = s_base + off;
        for (int k = 0; k < c; ++k) {
            const long long j = base + k;
            if (j < cap) { out[2 * j] = a; out[2 * j + 1] = row_peaks[(size_t)a * row_cap + k]; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_base;
}

extern "C" int rf_polar_peaks(rf_handle* h, const float* polar, int A, int W, int64_t* out, int64_t cap, int64_t* n) {
    if (!h || !polar || !n || A < 0 || W < 0 || cap < 0 || (cap > 0 && !out))
        return rf_fail(h, RF_E_BADARG, "rf_polar_peaks: bad argument");
    *n = 0;
    if (A == 0 || W < 3) return RF_OK;     // find_peaks needs three samples
    const int row_cap = W / 2 + 1;
    const size_t smem = (size_t)W * 4 + (size_t)row_cap * 8;
    if (smem > 200 * 1024) return rf_fail(h, RF_E_CAPACITY, "rf_polar_peaks: %d range bins exceed the shared-memory row buffer", W);
    const size_t polar_b = ((size_t)A * W * 4 + 255) & ~(size_t)255;
    const size_t rows_b = ((size_t)A * row_cap * 4 + 255) & ~(size_t)255;
    const size_t cnt_b = ((size_t)A * 4 + 255) & ~(size_t)255;
    const size_t take = (size_t)(cap < (int64_t)A * row_cap ? cap : (int64_t)A * row_cap);
    const size_t out_b = (take * 16 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, polar_b + rows_b + cnt_b + out_b + 256);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    float* d_polar = (float*)base;
    int32_t* d_rows = (int32_t*)(base + polar_b);
    int32_t* d_cnt = (int32_t*)(base + polar_b + rows_b);
    int64_t* d_out = (int64_t*)(base + polar_b + rows_b + cnt_b);
    long long* d_n = (long long*)(base + polar_b + rows_b + cnt_b + out_b);
    static bool attr_set[64] = {};   // per device: the attribute belongs to the device's context
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_peaks_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[h->device & 63] = true;
    }
    RF_CUDA(h, cudaMemcpyAsync(d_polar, polar, (size_t)A * W * 4, cudaMemcpyHostToDevice, h->stream));
    k_peaks_rows<<<A, 256, smem, h->stream>>>(d_polar, W, d_rows, row_cap, d_cnt);
    RF_CHECK_LAUNCH(h);
    k_peaks_gather<<<1, 256, 0, h->stream>>>(d_rows, row_cap, d_cnt, A, d_out, (long long)take, d_n);
    RF_CHECK_LAUNCH(h);
    long long total = 0;
    RF_CUDA(h, cudaMemcpyAsync(&total, d_n, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    *n = total;
    const size_t got = (size_t)(total < (long long)take ? total : (long long)take);
    if (got) {
        RF_CUDA(h, cudaMemcpyAsync(out, d_out, got * 16, cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RF_OK;
}
