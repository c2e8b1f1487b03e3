// k_features.cu — the per-azimuth polar peak extractor.
//
// Replaces (reference file:line):
//   getPointCloud.py:11-54   getPointCloudPolarInd                         -> k_peaks_rows + k_peaks_gather
// (feature detection, candidate selection and SSC live in k_detect.cu / k_doh.cu)
#include <math.h>

#include "common.cuh"

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
