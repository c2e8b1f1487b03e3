// k_detect.cu — feature detection and adaptive non-maximal suppression, entirely on the device.
//
// Replaces (reference file:line):
//   getFeatures.py:22-53     getBlobsFromCart (detector front half)        -> k_min_eig / k_doh_* + k_nms_select + k_sort_tiles / k_sort_global / k_sort_merge
//   getFeatures.py:66-72     adaptiveNMS (argsort + ssc)                   -> k_ssc_prepare_* + k_ssc_bisect
//   ANMS.py:5-102            ssc (Suppression via Square Covering)         -> k_ssc_bisect (the WHOLE binary search)
//
// Every kernel takes a batch of independent problems (one frame each) and an optional per-problem flag
// array, so the lock-step sequence runner (k_seq.cu) can re-detect on exactly the sequences that need
// it inside one CUDA graph, with no host round trip: the candidate count, the sort size and every
// bisection probe stay on the device.
//
// SSC.  The reference walks the keypoints in priority order; a keypoint whose cell is not covered
// yet is selected and covers the cells within +-reach.  Coverage is symmetric in cell coordinates
// (j covers i  <=>  |cell_i - cell_j| <= reach in both axes), so the greedy result is the unique
// "priority-greedy maximal independent set" and can be computed in parallel rounds: a live keypoint
// that has the highest priority among ALL live keypoints within +-reach cells is selected; live
// keypoints within reach of a keypoint selected this round die.  Induction over the priority order
// shows this yields exactly the sequential selection (a keypoint only dies next to a selected keypoint
// of higher priority, and a keypoint is only selected once every higher-priority neighbour is dead).
// Per cell only the best live keypoint matters, so a round is: atomicMin of the keypoint index into
// its cell, then a 5x5 cell scan.  The cover grid lives in shared memory for the widths the search
// visits on real images (<= 40 960 cells), else in global memory; for degenerate widths (grid larger than
// the workspace: only reachable with fewer keypoints than requested) a grid-free pass compares against
// the selected list.
#include <math.h>

#include "detect.cuh"

#define FULLM 0xffffffffu
#define CELL_EMPTY 0xffffffffu
#define SSC_THREADS 1024
#define SSC_SMEM_CELLS 40960            // 160 KB of shared cover grid
#define SSC_COV_WORDS 8192               // coverage bits for up to 262 144 cells (cells_cap of the workspaces is 1 << 18)
#define SORT_TILE 4096                  // u64 keys sorted inside shared memory (32 KB)

// =====================================================================================
// a10 corner response: minimum eigenvalue of the 3x3-summed structure tensor of the
// scaled 3x3 Sobel gradients (cv::cornerMinEigenVal, f32 input: scale = 1 / (4 * 3)).
// Block = 32 x 8 threads -> 32 x 32 outputs; the image tile (+2 halo) and the gradient
// products (+1 halo, evaluated at REFLECT_101 positions of the gradient image like
// cv::boxFilter does) are staged in shared memory.  blockIdx.z = problem.
// =====================================================================================
// One warp owns a strip of ME_COLS output columns and walks ME_ROWS output rows down it.  Lane l stands on gradient
// column x0 - 1 + l (lanes 1 .. 30 also own an output column); per gradient row it reads its 3 x 3 image taps straight
// from global memory (each image row is touched by three consecutive gradient rows and three neighbouring lanes: L1
// hits), forms the three gradient products, widens them once, takes the horizontal 3-sums from its neighbours with
// shuffles and keeps the last three rows of sums in registers for the vertical 3-sum.  Nothing is staged in shared
// memory: the tiled version of this kernel (36 x 36 image tile, three product planes, three f64 planes of row sums) spent
// 74 % of the L1 data pipe on shared-memory wavefronts and 188 instructions per pixel, half of them index arithmetic.
// Border rule as before: a gradient position outside the image is the gradient AT its REFLECT_101 position
// (cv::boxFilter's border on the product planes), whose taps are again reflected (cv::Sobel's border); the box sums are
// f64, rows first, then columns (cv::boxFilter's order).
#define ME_COLS 30
#define ME_ROWS 64
#define ME_WARPS 4

// scaled Sobel products at one gradient position from its eight neighbours
__device__ __forceinline__ void me_products(float a00, float a01, float a02, float a10, float a12, float a20, float a21, float a22, float k0,
                                            float k1, float& pxx, float& pxy, float& pyy) {
    // Dx: row filter [-1 0 1], column filter [k1 k0 k1] (the smoothing taps carry the scale)
    const float rx0 = __fsub_rn(a02, a00), rx1 = __fsub_rn(a12, a10), rx2 = __fsub_rn(a22, a20);
    const float dx = __fmaf_rn(k1, __fadd_rn(rx0, rx2), __fmul_rn(k0, rx1));
    // Dy: row filter [k1 k0 k1], column filter [-1 0 1]
    const float ry0 = __fmaf_rn(k1, __fadd_rn(a00, a02), __fmul_rn(k0, a01));
    const float ry2 = __fmaf_rn(k1, __fadd_rn(a20, a22), __fmul_rn(k0, a21));
    const float dy = __fsub_rn(ry2, ry0);
    pxx = __fmul_rn(dx, dx); pxy = __fmul_rn(dx, dy); pyy = __fmul_rn(dy, dy);
}

__global__ void __launch_bounds__(32 * ME_WARPS)
k_min_eig(const float* __restrict__ img_base, size_t img_stride, int n, float k0, float k1, float* __restrict__ resp_base,
          size_t resp_stride, const int32_t* __restrict__ flags, unsigned* __restrict__ maxbits) {
    // grid = (strip workers, problems): a CTA walks the strips of its problem, so the un-flagged problems of a lock-step
    // batch cost gridDim.x empty CTAs each instead of one per strip
    if (flags && !flags[blockIdx.y]) return;
    const float* __restrict__ img = img_base + (size_t)blockIdx.y * img_stride;
    float* __restrict__ resp = resp_base + (size_t)blockIdx.y * resp_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strips_x = (n + ME_COLS - 1) / ME_COLS, nstrips = strips_x * ((n + ME_ROWS - 1) / ME_ROWS);
    float vmax = 0.f;    // maximum of the responses this lane wrote (k_max_resp's rule: the maximum of max(resp, 0))
    for (int t = blockIdx.x * ME_WARPS + warp; t < nstrips; t += gridDim.x * ME_WARPS) {
        const int sy = t / strips_x, sx = t - sy * strips_x;
        const int x0 = sx * ME_COLS, y0 = sy * ME_ROWS;
        // this lane's gradient column and the columns of its three taps
        const int gx = min(max(x0 - 1 + lane, -1), n);
        const int qx = reflect101(gx, n);
        const int xm = reflect101(qx - 1, n), xp = reflect101(qx + 1, n);
        const int x = x0 + lane - 1;                                    // output column of lanes 1 .. ME_COLS
        const bool x_ok = lane >= 1 && lane <= ME_COLS && x < n;
        const int rows = min(ME_ROWS, n - y0);
        double hxx0 = 0, hxy0 = 0, hyy0 = 0, hxx1 = 0, hxy1 = 0, hyy1 = 0;   // row sums of gradient rows gy - 2, gy - 1
#pragma unroll 2
        for (int r = 0; r < rows + 2; ++r) {
            const int gy = min(max(y0 - 1 + r, -1), n);                 // <= n: y0 + rows <= n
            const int qy = reflect101(gy, n);
            const float* __restrict__ rm = img + (size_t)reflect101(qy - 1, n) * n;
            const float* __restrict__ rq = img + (size_t)qy * n;
            const float* __restrict__ rp = img + (size_t)reflect101(qy + 1, n) * n;
            float pxx, pxy, pyy;
            me_products(__ldg(rm + xm), __ldg(rm + qx), __ldg(rm + xp), __ldg(rq + xm), __ldg(rq + xp), __ldg(rp + xm), __ldg(rp + qx),
                        __ldg(rp + xp), k0, k1, pxx, pxy, pyy);
            const double dxx = (double)pxx, dxy = (double)pxy, dyy = (double)pyy;
            // horizontal 3-sum centred on this lane's gradient column: (left + centre) + right
            const double hxx2 = (__shfl_up_sync(FULLM, dxx, 1) + dxx) + __shfl_down_sync(FULLM, dxx, 1);
            const double hxy2 = (__shfl_up_sync(FULLM, dxy, 1) + dxy) + __shfl_down_sync(FULLM, dxy, 1);
            const double hyy2 = (__shfl_up_sync(FULLM, dyy, 1) + dyy) + __shfl_down_sync(FULLM, dyy, 1);
            if (r >= 2 && x_ok) {
                const double sxx = (hxx0 + hxx1) + hxx2, sxy = (hxy0 + hxy1) + hxy2, syy = (hyy0 + hyy1) + hyy2;
                const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, cc = __fmul_rn((float)syy, 0.5f);
                const float d = __fsub_rn(a, cc);
                const float v = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fmaf_rn(d, d, __fmul_rn(b, b))));
                resp[(size_t)(y0 + r - 2) * n + x] = v;
                vmax = fmaxf(vmax, v);
            }
            hxx0 = hxx1; hxy0 = hxy1; hyy0 = hyy1;
            hxx1 = hxx2; hxy1 = hxy2; hyy1 = hyy2;
        }
    }
    if (maxbits) {   // bit pattern order == value order for floats >= 0
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULLM, vmax, d));
        if (lane == 0) atomicMax(maxbits + blockIdx.y, __float_as_uint(vmax));
    }
}

// maximum of a non-negative response map (bit pattern order == value order for floats >= 0); blockIdx.y = problem
__global__ void __launch_bounds__(256)
k_max_resp(const float* __restrict__ resp_base, size_t resp_stride, size_t count, unsigned* __restrict__ out,
           const int32_t* __restrict__ flags) {
    if (flags && !flags[blockIdx.y]) return;
    const float* __restrict__ resp = resp_base + (size_t)blockIdx.y * resp_stride;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, __ldg(resp + i));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULLM, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out + blockIdx.y, __float_as_uint(m));
}

__global__ void k_clear_u32(unsigned* __restrict__ a, unsigned* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { if (a) a[i] = 0u; if (b) b[i] = 0u; }
}

// threshold + 3x3 non-maximum suppression (cv::goodFeaturesToTrack's rule: interior pixels with
// resp > thr and resp == max of the 3x3 neighbourhood).  Emits sortable 64-bit keys:
// high word = ~bits(resp) (so ascending key = descending response), low word = ~pixel index
// (ties: descending index, the order cv::goodFeaturesToTrack's pointer comparison produces).
// thr_rel != 0: threshold = (float)((double)max * thr_rel), cv's maxVal * qualityLevel.
// Warp strips like k_min_eig: lane l stands on column x0 - 1 + l, reads ONE response per row, takes the horizontal
// 3-maximum from its neighbours by shuffle and keeps the last rows in registers for the vertical one (the per-pixel
// version read 10 responses for every pixel above the threshold: 25 us per 1996^2 frame).
#define NMS_COLS 30
#define NMS_ROWS 64
#define NMS_WARPS 4
// Candidates are collected per warp in shared memory and appended to the problem's key list 32 or more at a time: one
// atomicAdd per batch instead of one per candidate.  The counters of all problems share a few cache lines, i.e. one L2
// slice, and with 130 k candidates per frame the per-candidate atomics of a lock-step batch serialised there (2.0 ms per
// 75 re-detecting frames, five times the kernel's memory time).
#define NMS_BUF 64
__global__ void __launch_bounds__(32 * NMS_WARPS)
k_nms_select(const float* __restrict__ resp_base, size_t resp_stride, int rows, int cols, float thr_abs, double thr_rel,
             const unsigned* __restrict__ maxbits, unsigned long long* __restrict__ keys_base, unsigned cap,
             unsigned* __restrict__ count, const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;      // grid = (strip workers, problems), see k_min_eig
    if (flags && !flags[p]) return;
    __shared__ unsigned long long s_keys[NMS_WARPS][NMS_BUF];
    const float* __restrict__ resp = resp_base + (size_t)p * resp_stride;
    const float thr = thr_rel != 0.0 ? (float)((double)__uint_as_float(maxbits[p]) * thr_rel) : thr_abs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* __restrict__ buf = s_keys[warp];
    unsigned long long* __restrict__ keys = keys_base + (size_t)p * cap;
    const int strips_x = (cols + NMS_COLS - 1) / NMS_COLS, nstrips = strips_x * ((rows + NMS_ROWS - 1) / NMS_ROWS);
    const float NEG = -3.402823466e+38f;
    int nbuf = 0;                               // warp-uniform fill of buf
    auto flush = [&]() {                        // append buf[0 .. nbuf) to the problem's list (slots past `cap` are dropped, still counted)
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(count + p, (unsigned)nbuf);
        base = __shfl_sync(FULLM, base, 0);
        for (int i = lane; i < nbuf; i += 32)
            if (base + i < cap) keys[base + i] = buf[i];
        nbuf = 0;
        __syncwarp();
    };
    for (int t = blockIdx.x * NMS_WARPS + warp; t < nstrips; t += gridDim.x * NMS_WARPS) {
        const int sy = t / strips_x, sx = t - sy * strips_x;
        const int x0 = sx * NMS_COLS, y0 = sy * NMS_ROWS;
        const int x = x0 - 1 + lane;                                  // this lane's column
        const bool x_in = x >= 0 && x < cols;
        const bool x_out = lane >= 1 && lane <= NMS_COLS && x >= 1 && x < cols - 1;   // interior output column of this strip
        const int nr = min(NMS_ROWS, rows - y0);
        float h0 = NEG, h1 = NEG, vc = NEG;     // horizontal maxima of rows y - 2, y - 1; the centre value of row y - 1
#pragma unroll 4
        for (int r = 0; r < nr + 2; ++r) {
            const int y = y0 - 1 + r;
            const float v = (x_in && y >= 0 && y < rows) ? __ldg(resp + (size_t)y * cols + x) : NEG;
            const float h2 = fmaxf(fmaxf(__shfl_up_sync(FULLM, v, 1), v), __shfl_down_sync(FULLM, v, 1));
            // centre row y - 1 (an output row of this strip for r >= 2), interior rows only
            const int yc = y - 1;
            const bool emit = r >= 2 && x_out && yc >= 1 && yc < rows - 1 && vc > thr && vc == fmaxf(fmaxf(h0, h1), h2);
            const unsigned bm = __ballot_sync(FULLM, emit);
            if (bm) {
                if (emit) {
                    const unsigned idx = (unsigned)yc * (unsigned)cols + (unsigned)x;
                    buf[nbuf + __popc(bm & ((1u << lane) - 1u))] = ((unsigned long long)(~__float_as_uint(vc)) << 32) | (unsigned long long)(~idx);
                }
                nbuf += __popc(bm);
                __syncwarp();
                if (nbuf > NMS_BUF - 32) flush();
            }
            h0 = h1; h1 = h2; vc = v;
        }
    }
    if (nbuf) flush();
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

// =====================================================================================
// Ascending sort of each problem's keys: a bitonic network spread over the machine.  The size follows the problem's
// own candidate count (device memory), so the launches are static and graph-capturable: every launch is sized for the
// key capacity and CTAs beyond a problem's padded size exit.  Keys are unique (the low word is the pixel index), so the
// result does not depend on the network.
//   k_sort_tiles    every tile of SORT_TILE keys fully sorted in shared memory (alternating direction = the network
//                   up to stage k = SORT_TILE), padding with ~0 up to the next power of two
//   k_sort_global   one stage (k, j >= SORT_TILE) of the network, one compare-exchange per thread
//   k_sort_merge    the stages j < SORT_TILE of step k, per tile in shared memory
// (Round 2's first version ran the whole network in ONE CTA per problem: 2.4 ms for 131 072 keys, the longest link of
// the chained step's latency chain.)
// =====================================================================================
__device__ __forceinline__ void cswap(unsigned long long& a, unsigned long long& b, bool up) {
    if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

__device__ __forceinline__ unsigned sort_np2(unsigned n) {
    unsigned np2 = 2;
    while (np2 < n) np2 <<= 1;
    return np2;
}

__global__ void __launch_bounds__(1024)
k_sort_tiles(unsigned long long* __restrict__ keys_base, unsigned cap, const unsigned* __restrict__ count,
             const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const unsigned n = min(count[p], cap);
    if (n < 2) return;
    const unsigned np2 = sort_np2(n);                 // <= cap (cap is a power of two)
    const unsigned T = np2 < SORT_TILE ? np2 : SORT_TILE;
    const unsigned base = blockIdx.x * SORT_TILE;
    if (base >= np2) return;
    __shared__ unsigned long long s[SORT_TILE];
    unsigned long long* __restrict__ a = keys_base + (size_t)p * cap;
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < T; i += 1024) s[i] = base + i < n ? a[base + i] : ~0ull;
    __syncthreads();
    for (unsigned k = 2; k <= T; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = tid; t < (T >> 1); t += 1024) {
                const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                cswap(s[i], s[i | j], ((base + i) & k) == 0);
            }
            __syncthreads();
        }
    for (unsigned i = tid; i < T; i += 1024) a[base + i] = s[i];
}

__global__ void __launch_bounds__(1024)
k_sort_global(unsigned long long* __restrict__ keys_base, unsigned cap, const unsigned* __restrict__ count, unsigned k, unsigned j,
              const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const unsigned n = min(count[p], cap);
    if (n < 2) return;
    const unsigned np2 = sort_np2(n);
    if (k > np2) return;
    const unsigned t = blockIdx.x * 1024 + threadIdx.x;
    if (t >= (np2 >> 1)) return;
    unsigned long long* __restrict__ a = keys_base + (size_t)p * cap;
    const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const unsigned long long x = a[i], y = a[i | j];
    if ((x > y) == ((i & k) == 0)) { a[i] = y; a[i | j] = x; }
}

__global__ void __launch_bounds__(1024)
k_sort_merge(unsigned long long* __restrict__ keys_base, unsigned cap, const unsigned* __restrict__ count, unsigned k,
             const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const unsigned n = min(count[p], cap);
    if (n < 2) return;
    const unsigned np2 = sort_np2(n);
    const unsigned base = blockIdx.x * SORT_TILE;
    if (k > np2 || base >= np2) return;
    __shared__ unsigned long long s[SORT_TILE];
    unsigned long long* __restrict__ a = keys_base + (size_t)p * cap;
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < SORT_TILE; i += 1024) s[i] = a[base + i];
    __syncthreads();
    for (unsigned j = SORT_TILE >> 1; j > 0; j >>= 1) {
        for (unsigned t = tid; t < (SORT_TILE >> 1); t += 1024) {
            const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            cswap(s[i], s[i | j], ((base + i) & k) == 0);
        }
        __syncthreads();
    }
    for (unsigned i = tid; i < SORT_TILE; i += 1024) a[base + i] = s[i];
}

// the whole network for S problems of capacity `cap` (a power of two)
static int launch_sort_keys(rf_handle* h, unsigned long long* keys, unsigned cap, const unsigned* count, int S, const int32_t* d_flags) {
    const unsigned tiles = cap > SORT_TILE ? cap / SORT_TILE : 1;
    k_sort_tiles<<<dim3(tiles, S), 1024, 0, h->stream>>>(keys, cap, count, d_flags);
    RF_CHECK_LAUNCH(h);
    for (unsigned k = 2 * SORT_TILE; k <= cap; k <<= 1) {
        for (unsigned j = k >> 1; j >= SORT_TILE; j >>= 1) {
            k_sort_global<<<dim3((cap / 2 + 1023) / 1024, S), 1024, 0, h->stream>>>(keys, cap, count, k, j, d_flags);
            RF_CHECK_LAUNCH(h);
        }
        k_sort_merge<<<dim3(tiles, S), 1024, 0, h->stream>>>(keys, cap, count, k, d_flags);
        RF_CHECK_LAUNCH(h);
    }
    return RF_OK;
}

// sorted keys -> keypoints (row, col) for SSC: the strongest min(count, max_kp) candidates
// (getFeatures.getBlobsFromCart caps at MAX_CANDIDATES; adaptiveNMS's stable argsort of a constant sigma keeps the order)
__global__ void __launch_bounds__(256)
k_ssc_prepare_keys(const unsigned long long* __restrict__ keys_base, unsigned key_cap, const unsigned* __restrict__ count,
                   int cols, unsigned max_kp, double2* __restrict__ rc_base, unsigned ssc_cap, int32_t* __restrict__ n_kp,
                   int32_t* __restrict__ status, const int32_t* __restrict__ flags) {
    const int p = blockIdx.y;
    if (flags && !flags[p]) return;
    const unsigned cnt = count[p];
    const unsigned n = min(min(cnt, key_cap), min(max_kp, ssc_cap));
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { n_kp[p] = (int)n; status[p] = cnt > key_cap ? RF_E_CAPACITY : RF_OK; }
    if (i >= n) return;
    const unsigned idx = ~(unsigned)(keys_base[(size_t)p * key_cap + i] & 0xffffffffull);
    rc_base[(size_t)p * ssc_cap + i] = make_double2((double)(idx / (unsigned)cols), (double)(idx % (unsigned)cols));
}

// caller-supplied keypoints [n][3] (row, col, sigma) -> rc
__global__ void __launch_bounds__(256) k_ssc_prepare_kp(const double* __restrict__ kp, int n, double2* __restrict__ rc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rc[i] = make_double2(kp[3 * (size_t)i], kp[3 * (size_t)i + 1]);
}

// =====================================================================================
// a9  ANMS.ssc: the whole binary search for one problem per CTA.
// =====================================================================================
struct SscArgs {
    const int32_t* flags;
    const double2* rc; const int32_t* n_kp; unsigned ssc_cap;
    int num_ret; double tol; int cols, rows;
    uint32_t* cell; uint32_t* alive; uint32_t* selmask; uint32_t* grid; unsigned cells_cap;
    int32_t* sel_idx; int32_t* m; int32_t* status;
};

__global__ void __launch_bounds__(SSC_THREADS) k_ssc_bisect(const SscArgs a) {
    extern __shared__ uint32_t s_grid[];           // SSC_SMEM_CELLS words
    // cells within reach of a keypoint selected so far in this probe, one bit per cell: a selected keypoint marks its
    // (2 reach + 1)^2 cells once, and a live keypoint then dies on ONE bit test instead of scanning that neighbourhood
    // for selected keypoints (25 grid reads + 25 mask reads per live keypoint and round, the bulk of this kernel)
    __shared__ uint32_t s_cov[SSC_COV_WORDS];
    __shared__ int s_w[32];
    __shared__ int s_nsel, s_next;
    const int p = blockIdx.x;
    if (a.flags && !a.flags[p]) return;
    const int tid = threadIdx.x;
    const int n = a.n_kp[p];
    const double2* __restrict__ rc = a.rc + (size_t)p * a.ssc_cap;
    uint32_t* __restrict__ cell = a.cell + (size_t)p * a.ssc_cap;
    uint32_t* alive0 = a.alive + (size_t)p * 2 * a.ssc_cap;
    uint32_t* alive1 = alive0 + a.ssc_cap;
    const unsigned mwords = (a.ssc_cap + 31) / 32;
    uint32_t* mask0 = a.selmask + (size_t)p * 2 * mwords;
    uint32_t* mask1 = mask0 + mwords;
    uint32_t* ggrid = a.grid + (size_t)p * a.cells_cap;
    int32_t* __restrict__ sel_idx = a.sel_idx + (size_t)p * a.ssc_cap;
    int status = a.status[p];                      // RF_E_CAPACITY from the candidate stage is kept
    if (n <= 0) { if (tid == 0) a.m[p] = 0; return; }   // nothing can be selected (the reference still walks its search)

    // ANMS.py:6-35 — closed-form upper bound of the search range (Python ints / floats restated in double)
    const double rows = a.rows, cols = a.cols, k = a.num_ret;
    const double exp1 = rows + cols + 2.0 * k;
    const double exp2 = 4.0 * cols + 4.0 * k + 4.0 * rows * k + rows * rows + cols * cols - 2.0 * rows * cols + 4.0 * rows * cols * k;
    const double exp3 = sqrt(exp2), exp4 = k - 1.0;
    const double sol1 = -nearbyint((exp1 + exp3) / exp4), sol2 = -nearbyint((exp1 - exp3) / exp4);   // Python round(): half to even
    double high = sol1 > sol2 ? sol1 : sol2;
    double low = floor(sqrt((double)n / k));
    const double kmin = nearbyint(k - k * a.tol), kmax = nearbyint(k + k * a.tol);
    double prev_width = -1.0;
    bool have = false;
    int last = 1, nres = 0;                        // mask holding the most recent pass; its selection count
    const int nwords = (n + 31) >> 5;

    for (;;) {
        const double width = low + (high - low) / 2;                  // ANMS.py:38
        if (width == prev_width || low > high) break;                 // -> result of the previous pass (:39-43)
        const double c = width / 2;
        if (!(c != 0.0)) { status = RF_E_BADARG; have = false; break; }        // the reference divides by zero (:45)
        const double fcc = floor(cols / c), fcr = floor(rows / c);
        if (!(fcc >= 0.0 && fcr >= 0.0)) { status = RF_E_BADARG; have = false; break; }
        const int cur = last ^ 1;
        uint32_t* mask = cur ? mask1 : mask0;
        for (int w = tid; w < nwords; w += SSC_THREADS) mask[w] = 0u;
        if (tid == 0) { s_nsel = 0; s_next = 0; }
        const double reach_d = floor(width / c);                      // 2 (:63-82)
        const double cells_d = (fcc + 1.0) * (fcr + 1.0);
        if (cells_d <= (double)a.cells_cap || cells_d <= (double)SSC_SMEM_CELLS) {
            const int reach = (int)reach_d, ncr = (int)fcr, ncc = (int)fcc, stride = ncc + 1;
            const unsigned cells = (unsigned)(ncr + 1) * (unsigned)stride;
            const bool in_smem = cells <= SSC_SMEM_CELLS;
            uint32_t* g = in_smem ? s_grid : ggrid;
            if (in_smem) for (unsigned i = tid; i < cells; i += SSC_THREADS) s_grid[i] = CELL_EMPTY;
            const bool use_cov = cells <= 32u * SSC_COV_WORDS;
            if (use_cov) for (unsigned i = tid; i < (cells + 31) / 32; i += SSC_THREADS) s_cov[i] = 0u;
            // (the passes over the live list are bound by the latency of their global loads on this one SM: four
            // independent elements per trip keep four loads in flight per thread)
            for (int i0 = tid; i0 < n; i0 += 4 * SSC_THREADS) {
                double2 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + u * SSC_THREADS; q[u] = i < n ? rc[i] : make_double2(0.0, 0.0); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * SSC_THREADS;
                    if (i < n) cell[i] = (unsigned)((int)floor(q[u].x / c) * stride + (int)floor(q[u].y / c));     // ANMS.py:52-59
                }
            }
            __syncthreads();
            int nal = n;
            bool first = true;
            uint32_t* src = alive0; uint32_t* dst = alive1;
            while (nal > 0) {
                // A: best live keypoint of every cell
                for (int t0 = tid; t0 < nal; t0 += 4 * SSC_THREADS) {
                    unsigned ii[4], cc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int t = t0 + u * SSC_THREADS; ii[u] = t < nal ? (first ? (unsigned)t : src[t]) : 0u; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cc[u] = (t0 + u * SSC_THREADS < nal) ? cell[ii[u]] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t0 + u * SSC_THREADS < nal) atomicMin(&g[cc[u]], ii[u]);
                }
                __syncthreads();
                // B: a keypoint that is the best of its whole neighbourhood is selected
                for (int tb = tid; tb < nal; tb += 4 * SSC_THREADS) {
                  unsigned bi[4], bc[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) { const int t = tb + u * SSC_THREADS; bi[u] = t < nal ? (first ? (unsigned)t : src[t]) : 0u; }
#pragma unroll
                  for (int u = 0; u < 4; ++u) bc[u] = (tb + u * SSC_THREADS < nal) ? cell[bi[u]] : 0u;
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    if (tb + u * SSC_THREADS >= nal) continue;
                    const unsigned i = bi[u];
                    const unsigned ci = bc[u];
                    if (g[ci] != i) continue;
                    const int r = (int)(ci / (unsigned)stride), cc = (int)(ci - (unsigned)r * (unsigned)stride);
                    const int r0 = max(r - reach, 0), r1 = min(r + reach, ncr), c0 = max(cc - reach, 0), c1 = min(cc + reach, ncc);
                    bool best = true;
                    for (int rr = r0; rr <= r1 && best; ++rr)
                        for (int c2 = c0; c2 <= c1; ++c2)
                            if (g[rr * stride + c2] < i) { best = false; break; }
                    if (best) {
                        atomicOr(&mask[i >> 5], 1u << (i & 31)); atomicAdd(&s_nsel, 1);
                        if (use_cov)
                            for (int rr = r0; rr <= r1; ++rr)
                                for (int c2 = c0; c2 <= c1; ++c2) { const unsigned cj = (unsigned)(rr * stride + c2); atomicOr(&s_cov[cj >> 5], 1u << (cj & 31)); }
                    }
                  }
                }
                __syncthreads();
                // C: survivors = live keypoints not within reach of a keypoint selected this round
                // (uniform trip count: every lane of a warp takes part in the ballot)
                for (int tq = 0; tq < nal; tq += 4 * SSC_THREADS) {
                  // the four elements' index / cell loads are issued together (uniform trip count: every lane of a warp
                  // takes part in the ballots below)
                  unsigned qi[4], qc[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) { const int t = tq + u * SSC_THREADS + tid; qi[u] = t < nal ? (first ? (unsigned)t : src[t]) : 0u; }
#pragma unroll
                  for (int u = 0; u < 4; ++u) qc[u] = (tq + u * SSC_THREADS + tid < nal) ? cell[qi[u]] : 0u;
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int t0 = tq + u * SSC_THREADS;
                    if (t0 >= nal) break;
                    const int t = t0 + tid;
                    unsigned i = 0;
                    bool keep = false;
                    if (t < nal) {
                        i = qi[u];
                        keep = !((mask[i >> 5] >> (i & 31)) & 1u);
                        if (keep && use_cov) {
                            const unsigned ci = qc[u];
                            keep = !((s_cov[ci >> 5] >> (ci & 31)) & 1u);
                        } else if (keep) {
                            const unsigned ci = qc[u];
                            const int r = (int)(ci / (unsigned)stride), cc = (int)(ci - (unsigned)r * (unsigned)stride);
                            const int r0 = max(r - reach, 0), r1 = min(r + reach, ncr), c0 = max(cc - reach, 0), c1 = min(cc + reach, ncc);
                            for (int rr = r0; rr <= r1 && keep; ++rr)
                                for (int c2 = c0; c2 <= c1; ++c2) {
                                    const unsigned v = g[rr * stride + c2];
                                    if (v != CELL_EMPTY && ((mask[v >> 5] >> (v & 31)) & 1u)) { keep = false; break; }
                                }
                        }
                    }
                    // warp-aggregated append (the order inside the live list is irrelevant)
                    const unsigned bm = __ballot_sync(FULLM, keep);
                    const int lane = tid & 31;
                    int basei = 0;
                    if (bm && lane == 0) basei = atomicAdd(&s_next, __popc(bm));
                    basei = __shfl_sync(FULLM, basei, 0);
                    if (keep) dst[basei + __popc(bm & ((1u << lane) - 1u))] = i;
                  }
                }
                __syncthreads();
                // D: leave the grid empty again
                for (int t0 = tid; t0 < nal; t0 += 4 * SSC_THREADS) {
                    unsigned ii[4], cc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int t = t0 + u * SSC_THREADS; ii[u] = t < nal ? (first ? (unsigned)t : src[t]) : 0u; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cc[u] = (t0 + u * SSC_THREADS < nal) ? cell[ii[u]] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (t0 + u * SSC_THREADS < nal) g[cc[u]] = CELL_EMPTY;
                }
                nal = s_next;
                __syncthreads();
                if (tid == 0) s_next = 0;
                first = false;
                uint32_t* tsw = src; src = dst; dst = tsw;
                __syncthreads();
            }
            nres = s_nsel;
        } else {
            // grid-free pass: covered iff an earlier SELECTED keypoint lies within +-reach cells (the clamping at the
            // grid edges never changes that test).  Cell coordinates stay doubles (integer-valued, exact).
            uint32_t* sl = alive0;
            int nsel = 0;
            __syncthreads();
            for (int i = 0; i < n; ++i) {
                const double2 q = rc[i];
                const double row = floor(q.x / c), col = floor(q.y / c);
                bool cov = false;
                for (int j = tid; j < nsel && !cov; j += SSC_THREADS) {
                    const double2 s = rc[sl[j]];
                    cov = fabs(floor(s.x / c) - row) <= reach_d && fabs(floor(s.y / c) - col) <= reach_d;
                }
                if (!__syncthreads_or(cov)) {
                    if (tid == 0) { sl[nsel] = (unsigned)i; mask[i >> 5] |= 1u << (i & 31); }
                    ++nsel;
                }
                __syncthreads();
            }
            nres = nsel;
        }
        last = cur;
        have = true;
        if (kmin <= (double)nres && (double)nres <= kmax) break;               // ANMS.py:89-91
        if ((double)nres < kmin) high = width - 1; else low = width + 1;       // ANMS.py:92-95
        prev_width = width;
        __syncthreads();
    }
    __syncthreads();
    int m = 0;
    if (have) {
        // ordered compaction of the selected bits: ascending index == selection order
        const uint32_t* mask = last ? mask1 : mask0;
        const int wpt = (nwords + SSC_THREADS - 1) / SSC_THREADS;
        const int w0 = min(tid * wpt, nwords), w1 = min(w0 + wpt, nwords);
        int c = 0;
        for (int w = w0; w < w1; ++w) c += __popc(mask[w]);
        // block exclusive scan
        const int lane = tid & 31, wi = tid >> 5;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLM, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) s_w[wi] = inc;
        __syncthreads();
        if (wi == 0) {
            int x = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLM, x, o); if (lane >= o) x += t; }
            s_w[lane] = x;
        }
        __syncthreads();
        int off = (wi ? s_w[wi - 1] : 0) + inc - c;
        m = s_w[31];
        for (int w = w0; w < w1; ++w) {
            uint32_t v = mask[w];
            while (v) { const int b = __ffs(v) - 1; sel_idx[off++] = w * 32 + b; v &= v - 1; }
        }
    }
    if (tid == 0) { a.m[p] = m; a.status[p] = status; }
}

// =====================================================================================
// host side
// =====================================================================================
static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static void ws_layout(DetectWs& ws, char* base, bool with_resp, size_t* total) {
    char* p = base;
    auto take = [&](size_t bytes) { char* r = p; p += al256(bytes); return r; };
    const size_t S = ws.S;
    ws.resp_stride = (size_t)ws.rows * ws.cols;
    ws.resp = with_resp ? (float*)take(S * ws.resp_stride * sizeof(float)) : nullptr;
    ws.maxbits = (unsigned*)take(S * 4);
    ws.count = (unsigned*)take(S * 4);
    ws.keys = (unsigned long long*)take(S * (size_t)ws.key_cap * 8);
    ws.rc = (double2*)take(S * (size_t)ws.ssc_cap * 16);
    ws.n_kp = (int32_t*)take(S * 4);
    ws.cell = (uint32_t*)take(S * (size_t)ws.ssc_cap * 4);
    ws.alive = (uint32_t*)take(S * 2 * (size_t)ws.ssc_cap * 4);
    ws.selmask = (uint32_t*)take(S * 2 * (size_t)((ws.ssc_cap + 31) / 32) * 4);
    ws.grid = (uint32_t*)take(S * (size_t)ws.cells_cap * 4);
    ws.sel_idx = (int32_t*)take(S * (size_t)ws.ssc_cap * 4);
    ws.m = (int32_t*)take(S * 4);
    ws.status = (int32_t*)take(S * 4);
    *total = (size_t)(p - base);
}

size_t rf_detect_ws_bytes(int S, int rows, int cols, unsigned key_cap, unsigned ssc_cap, unsigned cells_cap, bool with_resp) {
    DetectWs ws; memset(&ws, 0, sizeof(ws));
    ws.S = S; ws.rows = rows; ws.cols = cols; ws.key_cap = key_cap; ws.ssc_cap = ssc_cap; ws.cells_cap = cells_cap;
    size_t total = 0;
    ws_layout(ws, nullptr, with_resp, &total);
    return total + 256;
}

DetectWs rf_detect_ws_carve(void* base, int S, int rows, int cols, unsigned key_cap, unsigned ssc_cap, unsigned cells_cap, bool with_resp) {
    DetectWs ws; memset(&ws, 0, sizeof(ws));
    ws.S = S; ws.rows = rows; ws.cols = cols; ws.key_cap = key_cap; ws.ssc_cap = ssc_cap; ws.cells_cap = cells_cap;
    size_t total = 0;
    ws_layout(ws, (char*)base, with_resp, &total);
    ws.base = base; ws.bytes = total;
    return ws;
}

int rf_detect_ws_init(rf_handle* h, const DetectWs& ws) {
    RF_CUDA(h, cudaMemsetAsync(ws.grid, 0xFF, (size_t)ws.S * ws.cells_cap * 4, h->stream));
    RF_CUDA(h, cudaMemsetAsync(ws.status, 0, (size_t)ws.S * 4, h->stream));
    RF_CUDA(h, cudaMemsetAsync(ws.m, 0, (size_t)ws.S * 4, h->stream));
    RF_CUDA(h, cudaMemsetAsync(ws.n_kp, 0, (size_t)ws.S * 4, h->stream));
    RF_CUDA(h, cudaMemsetAsync(ws.count, 0, (size_t)ws.S * 4, h->stream));
    return RF_OK;
}

int rf_detect_prepare(rf_handle* h) {
    static bool attr_set[64] = {};   // per device: the attribute belongs to the device's context
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_ssc_bisect, cudaFuncAttributeMaxDynamicSharedMemorySize, SSC_SMEM_CELLS * 4));
        attr_set[h->device & 63] = true;
    }
    return RF_OK;
}

int rf_launch_min_eig(rf_handle* h, const float* d_img, size_t img_stride, int n, float* d_resp, size_t resp_stride, int S,
                      const int32_t* d_flags, unsigned* d_maxbits) {
    const double scale = 1.0 / ((double)(1 << 2) * 3.0);  // ksize 3, blockSize 3, f32 input
    const int nstrips = ((n + ME_COLS - 1) / ME_COLS) * ((n + ME_ROWS - 1) / ME_ROWS);
    dim3 grd(rf_tile_workers(h, (nstrips + ME_WARPS - 1) / ME_WARPS, S), S);
    k_min_eig<<<grd, 32 * ME_WARPS, 0, h->stream>>>(d_img, img_stride, n, (float)(2.0 * scale), (float)(1.0 * scale), d_resp, resp_stride, d_flags,
                                                    d_maxbits);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_detect_clear(rf_handle* h, const DetectWs& ws) {
    k_clear_u32<<<(ws.S + 255) / 256, 256, 0, h->stream>>>(ws.count, ws.maxbits, ws.S);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_select_sorted(rf_handle* h, const DetectWs& ws, const float* d_resp, size_t resp_stride, float threshold,
                            const int32_t* d_flags, bool have_max) {
    const int S = ws.S;
    if (!have_max) {
        int rc = rf_launch_detect_clear(h, ws);
        if (rc) return rc;
    }
    double rel = 0.0;
    if (threshold < 0) {   // relative: -threshold is cv2.goodFeaturesToTrack's qualityLevel (fraction of the maximum)
        if (!have_max) {
            dim3 g(h->sm_count * 2 > 64 ? 64 : h->sm_count * 2, S);
            k_max_resp<<<g, 256, 0, h->stream>>>(d_resp, resp_stride, (size_t)ws.rows * ws.cols, ws.maxbits, d_flags);
            RF_CHECK_LAUNCH(h);
        }
        rel = (double)(-threshold);
    }
    const int nstrips = ((ws.cols + NMS_COLS - 1) / NMS_COLS) * ((ws.rows + NMS_ROWS - 1) / NMS_ROWS);
    dim3 grd(rf_tile_workers(h, (nstrips + NMS_WARPS - 1) / NMS_WARPS, S), S);
    k_nms_select<<<grd, 32 * NMS_WARPS, 0, h->stream>>>(d_resp, resp_stride, ws.rows, ws.cols, threshold, rel, ws.maxbits, ws.keys, ws.key_cap,
                                                        ws.count, d_flags);
    RF_CHECK_LAUNCH(h);
    return launch_sort_keys(h, ws.keys, ws.key_cap, ws.count, S, d_flags);
}

int rf_launch_ssc(rf_handle* h, const DetectWs& ws, int num_ret, double tol, int cols, int rows, const int32_t* d_flags) {
    int rc = rf_detect_prepare(h);
    if (rc) return rc;
    SscArgs a;
    a.flags = d_flags; a.rc = ws.rc; a.n_kp = ws.n_kp; a.ssc_cap = ws.ssc_cap; a.num_ret = num_ret; a.tol = tol; a.cols = cols;
    a.rows = rows; a.cell = ws.cell; a.alive = ws.alive; a.selmask = ws.selmask; a.grid = ws.grid; a.cells_cap = ws.cells_cap;
    a.sel_idx = ws.sel_idx; a.m = ws.m; a.status = ws.status;
    k_ssc_bisect<<<ws.S, SSC_THREADS, SSC_SMEM_CELLS * 4, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_ssc_from_keys(rf_handle* h, const DetectWs& ws, int num_ret, double tol, const int32_t* d_flags) {
    const unsigned max_kp = ws.ssc_cap < RF_SSC_MAX_CANDIDATES ? ws.ssc_cap : RF_SSC_MAX_CANDIDATES;
    dim3 g((max_kp + 255) / 256, ws.S);
    k_ssc_prepare_keys<<<g, 256, 0, h->stream>>>(ws.keys, ws.key_cap, ws.count, ws.cols, max_kp, ws.rc, ws.ssc_cap, ws.n_kp,
                                                 ws.status, d_flags);
    RF_CHECK_LAUNCH(h);
    // getFeatures.adaptiveNMS: ssc(keypoints, ret_points, tolerance, W, H)
    return rf_launch_ssc(h, ws, num_ret, tol, ws.cols, ws.rows, d_flags);
}

static unsigned pow2_at_least(size_t v) { unsigned p = 2; while (p < v) p <<= 1; return p; }

extern "C" {

int rf_ssc(rf_handle* h, const double* kp, int n, int num_ret, double tol, int cols, int rows, int32_t* sel_idx, int* m) {
    RfDeviceGuard rf_guard_(h);
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
    const double exp2 = 4.0 * cols + 4.0 * num_ret + 4.0 * rows * (double)num_ret + (double)rows * rows +
                        (double)cols * cols - 2.0 * rows * (double)cols + 4.0 * (double)rows * cols * num_ret;
    if (exp2 < 0) return rf_fail(h, RF_E_BADARG, "rf_ssc: math domain error (ANMS.py:17)");
    if (num_ret <= 0) return rf_fail(h, RF_E_BADARG, "rf_ssc: num_ret_points must be positive");
    const unsigned cap = (unsigned)((n + 31) & ~31);
    const unsigned cells_cap = 1u << 20;
    const size_t kp_bytes = al256((size_t)n * 3 * sizeof(double));
    const size_t wsb = rf_detect_ws_bytes(1, rows, cols, 2, cap, cells_cap, false);
    int rc = rf_ensure_scratch(h, kp_bytes + wsb);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    DetectWs ws = rf_detect_ws_carve(base + kp_bytes, 1, rows, cols, 2, cap, cells_cap, false);
    if ((rc = rf_detect_ws_init(h, ws))) return rc;
    RF_CUDA(h, cudaMemcpyAsync(base, kp, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(ws.n_kp, &n, 4, cudaMemcpyHostToDevice, h->stream));
    k_ssc_prepare_kp<<<(n + 255) / 256, 256, 0, h->stream>>>((const double*)base, n, ws.rc);
    RF_CHECK_LAUNCH(h);
    if ((rc = rf_launch_ssc(h, ws, num_ret, tol, cols, rows, nullptr))) return rc;
    int32_t out[2] = {0, 0};
    RF_CUDA(h, cudaMemcpyAsync(&out[0], ws.m, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&out[1], ws.status, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(sel_idx, ws.sel_idx, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out[1] != RF_OK) return rf_fail(h, out[1], "rf_ssc: zero or negative cell size (the reference divides by zero, ANMS.py:45)");
    *m = out[0];
    return RF_OK;
}

}  // extern "C"

// select + sort on a device response map (one problem); out (host) receives min(*n, cap) rows
static int select_sorted_to_host(rf_handle* h, const float* d_resp, int rows, int cols, float thr, char* ws_base, unsigned key_cap,
                                 double* out, int cap, int* n) {
    DetectWs ws = rf_detect_ws_carve(ws_base, 1, rows, cols, key_cap, 32, 32, false);
    int rc = rf_launch_select_sorted(h, ws, d_resp, (size_t)rows * cols, thr, nullptr);
    if (rc) return rc;
    unsigned cnt = 0;
    RF_CUDA(h, cudaMemcpyAsync(&cnt, ws.count, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (cnt > key_cap)
        return rf_fail(h, RF_E_CAPACITY, "rf_detect: %u candidates exceed the key capacity %u (raise the threshold)", cnt, key_cap);
    *n = (int)cnt;
    const unsigned take = cnt < (unsigned)cap ? cnt : (unsigned)cap;
    if (take) {
        double* d_rows = (double*)(ws_base + al256(ws.bytes));
        k_keys_to_rows<<<(take + 255) / 256, 256, 0, h->stream>>>(ws.keys, take, cols, d_rows);
        RF_CHECK_LAUNCH(h);
        RF_CUDA(h, cudaMemcpyAsync(out, d_rows, (size_t)take * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RF_OK;
}

static int detect_scratch(rf_handle* h, size_t resp_bytes, int rows, int cols, unsigned key_cap, int cap, size_t* ws_off) {
    *ws_off = al256(resp_bytes);
    return rf_ensure_scratch(h, *ws_off + al256(rf_detect_ws_bytes(1, rows, cols, key_cap, 32, 32, false)) + (size_t)cap * 24 + 512);
}

// k_doh.cu: determinant-of-Hessian response planes (mode 1)
int rf_doh_detect_host(rf_handle* h, const rf_frame* f, float threshold, double* out, int cap, int* n_out);

extern "C" {

int rf_corner_response(rf_handle* h, const rf_frame* f, int mode, float* resp) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !f || !resp) return rf_fail(h, RF_E_BADARG, "rf_corner_response: null argument");
    if (mode != 0) return rf_fail(h, RF_E_BADARG, "rf_corner_response: mode %d has no single response plane (0 = structure-tensor min eigenvalue; DoH planes: rf_doh_response)", mode);
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_corner_response: frame has no f32 plane");
    const size_t bytes = (size_t)h->n * h->n * sizeof(float);
    int rc = rf_ensure_scratch(h, bytes);
    if (rc) return rc;
    if ((rc = rf_launch_min_eig(h, f->fs.cart, 0, h->n, (float*)h->d_scratch, 0, 1, nullptr))) return rc;
    RF_CUDA(h, cudaMemcpyAsync(resp, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

int rf_detect(rf_handle* h, const rf_frame* f, int mode, float threshold, double* out, int cap, int* n) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !f || !n || cap < 0 || (cap > 0 && !out)) return rf_fail(h, RF_E_BADARG, "rf_detect: bad argument");
    if (mode != 0 && mode != 1) return rf_fail(h, RF_E_BADARG, "rf_detect: mode %d is not available (0 = structure-tensor min eigenvalue, 1 = determinant of Hessian)", mode);
    if (!f->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_detect: frame has no f32 plane");
    if (mode == 1) return rf_doh_detect_host(h, f, threshold, out, cap, n);
    const size_t bytes = (size_t)h->n * h->n * sizeof(float);
    const unsigned key_cap = pow2_at_least((size_t)h->n * h->n / 4 + 1024);   // 3x3 maxima cannot be denser than 1 in 4
    size_t off;
    int rc = detect_scratch(h, bytes, h->n, h->n, key_cap, cap, &off);
    if (rc) return rc;
    float* d_resp = (float*)h->d_scratch;
    if ((rc = rf_launch_min_eig(h, f->fs.cart, 0, h->n, d_resp, 0, 1, nullptr))) return rc;
    return select_sorted_to_host(h, d_resp, h->n, h->n, threshold, (char*)h->d_scratch + off, key_cap, out, cap, n);
}

int rf_nms_select(rf_handle* h, const float* resp, int rows, int cols, float threshold, double* out, int cap, int* n) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !resp || !n || rows < 3 || cols < 3 || cap < 0 || (cap > 0 && !out))
        return rf_fail(h, RF_E_BADARG, "rf_nms_select: bad argument");
    const size_t bytes = (size_t)rows * cols * sizeof(float);
    const unsigned key_cap = pow2_at_least((size_t)rows * cols + 1024);        // plateaus may keep every pixel
    size_t off;
    int rc = detect_scratch(h, bytes, rows, cols, key_cap, cap, &off);
    if (rc) return rc;
    RF_CUDA(h, cudaMemcpyAsync(h->d_scratch, resp, bytes, cudaMemcpyHostToDevice, h->stream));
    return select_sorted_to_host(h, (const float*)h->d_scratch, rows, cols, threshold, (char*)h->d_scratch + off, key_cap, out, cap, n);
}

}  // extern "C"
