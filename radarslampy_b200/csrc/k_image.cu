// k_image.cu — scan decode, polar->Cartesian remap, u8 conversion and the LK image pyramid.
//
// Replaces (reference file:line):
//   parseData.py:17-53    extractDataFromRadarImage        -> k_extract
//   parseData.py:100-135  convertPolarImageToCartesian     -> k_build_map (once) + k_polar2cart
//   getTransformKLT.py:356-357  (img*255).astype(uint8)    -> fused into k_polar2cart / k_cart_to_u8
//   cv2.buildOpticalFlowPyramid inside calcOpticalFlowPyrLK (getTransformKLT.py:359-360) -> k_pyr_down
//
// Arithmetic contract (bit-exact with cv2 4.13 on AVX2/AVX-512 hosts, see oracle/c/oracle_c.c):
//   every rounding step of cv::warpPolar is reproduced with explicit _rn intrinsics so that
//   the compiler can neither contract nor reassociate; the only FMAs are the three Horner
//   steps of cv::fastAtan32f.  The inverse map depends on geometry only, so it is evaluated
//   ONCE per handle into a packed table (sx:17 | sy:14 bits) that stays L2-resident; the
//   per-frame kernel is then a pure gather + 7 flops per pixel, bound by HBM writes.
#include <float.h>
#include <math.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------
// geometry table
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_map(uint32_t* __restrict__ map, int n, float cx, float cy,
                                                   double Kmag, double Kangle, int semilog) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= n || y >= n) return;
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float dx = __fsub_rn((float)x, cx), dy = __fsub_rn((float)y, cy);
    float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    // WARP_POLAR_LOG: bufp += 1.f; cv::log(bufp, bufp).  cv::log's f32 kernel is a vendor routine that differs from a
    // correctly rounded logf by at most one ulp on ~2 % of its arguments, so this mode is a tolerance match (parseData.py:131-133)
    if (semilog) mag = logf(__fadd_rn(mag, 1.0f));
    float ax = fabsf(dx), ay = fabsf(dy);
    float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    float c = __fdiv_rn(mn, __fadd_rn(mx, (float)DBL_EPSILON));
    float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(p7, c2, p5), c2, p3), c2, p1), c);
    if (ax < ay) a = __fsub_rn(90.0f, a);
    if (dx < 0.f) a = __fsub_rn(180.0f, a);
    if (dy < 0.f) a = __fsub_rn(360.0f, a);
    float ang = __fmul_rn(a, (float)(M_PI / 180.0));
    float mapx = __double2float_rn(__ddiv_rn((double)mag, Kmag));
    float mapy = __fadd_rn(__double2float_rn(__ddiv_rn((double)ang, Kangle)), 1.0f);
    int sx = __float2int_rn(__fmul_rn(mapx, 32.0f));
    int sy = __float2int_rn(__fmul_rn(mapy, 32.0f));
    sx = min(sx, (1 << 17) - 1);  // beyond every legal range bin -> all four taps are outliers
    sy = min(max(sy, 0), (1 << 14) - 1);
    map[(size_t)y * n + x] = (uint32_t)sx | ((uint32_t)sy << 17);
}

// ------------------------------------------------------------------------------------
// per-frame remap.  One thread = 4 consecutive destination pixels (one 128-bit map load,
// one 128-bit f32 store, one 32-bit u8 store).  Block = 16 x 16 threads = 64 x 16 pixels,
// a compact footprint in the polar scan so the byte taps hit L1.
// ------------------------------------------------------------------------------------
template <typename TapT>
struct Tap;
template <>
struct Tap<uint8_t> {
    static __device__ __forceinline__ float get(const uint8_t* p, const float* lut) { return lut[__ldg(p)]; }
};
template <>
struct Tap<float> {
    static __device__ __forceinline__ float get(const float* p, const float*) { return __ldg(p); }
};

template <typename TapT, bool WRITE_F32>
__global__ void __launch_bounds__(256)
k_polar2cart(const TapT* __restrict__ src, size_t src_frame_stride, int row_pitch, int col0, int A, int W,
             const uint32_t* __restrict__ map, int n, float* __restrict__ cart, size_t cart_stride,
             uint8_t* __restrict__ l0, size_t l0_stride, int l0_pitch, int first) {
    __shared__ float lut[256];
    const int tid = threadIdx.y * 16 + threadIdx.x;
    lut[tid] = __fdiv_rn((float)tid, 255.0f);  // parseData.py:43  u8 -> f32 / 255. (IEEE division)
    __syncthreads();
    const int x4 = blockIdx.x * 16 + threadIdx.x;
    const int y = blockIdx.y * 16 + threadIdx.y;
    const int n4 = n >> 2;
    if (x4 >= n4 || y >= n) return;
    const int frame = first + blockIdx.z;
    const TapT* s = src + (size_t)blockIdx.z * src_frame_stride + col0;
    const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(map) + (size_t)y * n4 + x4);
    const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int sx = mm[k] & 0x1FFFF, sy = mm[k] >> 17;
        const int ix = sx >> 5, iy = sy >> 5;
        const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
        const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
        const float w00 = __fmul_rn(gy, gx), w01 = __fmul_rn(gy, fx);
        const float w10 = __fmul_rn(fy, gx), w11 = __fmul_rn(fy, fx);
        // wrapped source rows: row r of cv2's bordered image is azimuth (r - 1) mod A
        int r0 = iy - 1; r0 = r0 < 0 ? r0 + A : (r0 >= A ? r0 - A : r0);
        int r1 = iy;     r1 = r1 >= A ? r1 - A : r1;
        const bool y0ok = iy < A + 2, y1ok = iy + 1 < A + 2;
        const bool x0ok = ix < W, x1ok = ix + 1 < W;
        const TapT* p0 = s + (size_t)r0 * row_pitch + ix;
        const TapT* p1 = s + (size_t)r1 * row_pitch + ix;
        const float v00 = (y0ok && x0ok) ? Tap<TapT>::get(p0, lut) : 0.0f;
        const float v01 = (y0ok && x1ok) ? Tap<TapT>::get(p0 + 1, lut) : 0.0f;
        const float v10 = (y1ok && x0ok) ? Tap<TapT>::get(p1, lut) : 0.0f;
        const float v11 = (y1ok && x1ok) ? Tap<TapT>::get(p1 + 1, lut) : 0.0f;
        float acc = __fmul_rn(v00, w00);
        acc = __fadd_rn(acc, __fmul_rn(v01, w01));
        acc = __fadd_rn(acc, __fmul_rn(v10, w10));
        acc = __fadd_rn(acc, __fmul_rn(v11, w11));
        o[k] = acc;
    }
    const size_t pix = (size_t)y * n + (size_t)x4 * 4;
    if (WRITE_F32) {
        float4 v = make_float4(o[0], o[1], o[2], o[3]);
        __stcs(reinterpret_cast<float4*>(cart + (size_t)frame * cart_stride + pix), v);
    }
    // getTransformKLT.py:356  (img * 255).astype(np.uint8): f32 multiply, truncate toward zero
    uchar4 u;
    u.x = (unsigned char)__float2int_rz(__fmul_rn(o[0], 255.0f));
    u.y = (unsigned char)__float2int_rz(__fmul_rn(o[1], 255.0f));
    u.z = (unsigned char)__float2int_rz(__fmul_rn(o[2], 255.0f));
    u.w = (unsigned char)__float2int_rz(__fmul_rn(o[3], 255.0f));
    *reinterpret_cast<uchar4*>(l0 + (size_t)frame * l0_stride + (size_t)y * l0_pitch + (size_t)x4 * 4) = u;
}

// f32 Cartesian image supplied by the caller -> u8 level 0
__global__ void __launch_bounds__(256) k_cart_to_u8(const float* __restrict__ cart, uint8_t* __restrict__ l0, size_t count4, int n4, int pitch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    const size_t y = i / n4, x4 = i - y * n4;
    float4 v = __ldg(reinterpret_cast<const float4*>(cart) + i);
    uchar4 u;
    u.x = (unsigned char)__float2int_rz(__fmul_rn(v.x, 255.0f));
    u.y = (unsigned char)__float2int_rz(__fmul_rn(v.y, 255.0f));
    u.z = (unsigned char)__float2int_rz(__fmul_rn(v.z, 255.0f));
    u.w = (unsigned char)__float2int_rz(__fmul_rn(v.w, 255.0f));
    *reinterpret_cast<uchar4*>(l0 + y * pitch + 4 * x4) = u;
}

// ------------------------------------------------------------------------------------
// cv::pyrDown (5x5 binomial, BORDER_REFLECT_101, (sum + 128) >> 8), one level.
// Block = 256 threads -> 64 x 16 output pixels; the (131 x 35) input tile is staged in
// shared memory once, filtered horizontally at even columns, then vertically.
// ------------------------------------------------------------------------------------
#define PD_TW 64
#define PD_TH 16
#define PD_IW (2 * PD_TW + 3)
#define PD_IH (2 * PD_TH + 3)
__global__ void __launch_bounds__(256)
k_pyr_down(const uint8_t* __restrict__ src, size_t src_stride, int sp, int sw, int sh, uint8_t* __restrict__ dst,
           size_t dst_stride, int dp, int dw, int dh, int first) {
    __shared__ uint8_t tile[PD_IH][PD_IW + 1];
    __shared__ uint16_t hsum[PD_IH][PD_TW];
    const int frame = first + blockIdx.z;
    const uint8_t* s = src + (size_t)frame * src_stride;
    uint8_t* d = dst + (size_t)frame * dst_stride;
    const int ox = blockIdx.x * PD_TW, oy = blockIdx.y * PD_TH;
    const int ix0 = 2 * ox - 2, iy0 = 2 * oy - 2;
    const int tid = threadIdx.x;
    for (int i = tid; i < PD_IH * PD_IW; i += 256) {
        int r = i / PD_IW, c = i - r * PD_IW;
        int yy = reflect101(iy0 + r, sh), xx = reflect101(ix0 + c, sw);
        // rows/cols past the last needed tap may reflect twice on tiny levels; clamp defensively
        yy = min(max(yy, 0), sh - 1); xx = min(max(xx, 0), sw - 1);
        tile[r][c] = __ldg(s + (size_t)yy * sp + xx);
    }
    __syncthreads();
    for (int i = tid; i < PD_IH * PD_TW; i += 256) {
        int r = i / PD_TW, c = i - r * PD_TW;
        const uint8_t* t = &tile[r][2 * c];
        hsum[r][c] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
    }
    __syncthreads();
    for (int i = tid; i < PD_TH * PD_TW; i += 256) {
        int r = i / PD_TW, c = i - r * PD_TW;
        int X = ox + c, Y = oy + r;
        if (X < dw && Y < dh) {
            int v = hsum[2 * r][c] + 4 * hsum[2 * r + 1][c] + 6 * hsum[2 * r + 2][c] + 4 * hsum[2 * r + 3][c] +
                    hsum[2 * r + 4][c];
            d[(size_t)Y * dp + X] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// parseData.py:39-43: power bins -> f32 / 255.  (metadata columns are decoded on the host side of the ABI:
// 11 bytes per azimuth, no arithmetic worth a kernel.)
__global__ void __launch_bounds__(256)
k_extract(const uint8_t* __restrict__ raw, int A, int raw_width, int meta, int W, float* __restrict__ polar) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int a = blockIdx.y;
    if (r >= W) return;
    polar[(size_t)a * W + r] = __fdiv_rn((float)__ldg(raw + (size_t)a * raw_width + meta + r), 255.0f);
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
int rf_frameset_alloc(rf_handle* h, FrameSet* fs, int count, bool with_f32) {
    memset(fs, 0, sizeof(*fs));
    fs->count = count;
    int w = h->n, hh = h->n;
    int nl = 0;
    for (int l = 0; l <= h->cfg.klt_max_level && l < RF_MAX_LEVELS; ++l) {
        if (l > 0) {
            w = (w + 1) / 2; hh = (hh + 1) / 2;
            if (w <= h->cfg.klt_win || hh <= h->cfg.klt_win) break;  // cv::buildOpticalFlowPyramid stop rule
        }
        fs->w[l] = w; fs->h[l] = hh;
        // 16-byte multiple (TMA strides); never a multiple of 256: the rows of a KLT window or pyrDown tile would all fall
        // into the same L1 / L2 sets (measured: k_klt 0.645 -> 0.70 ms with pitches 512 / 256 on levels 2 / 3)
        fs->pitch[l] = (w + 15) & ~15;
        if (fs->pitch[l] % 256 == 0) fs->pitch[l] += 16;
        fs->lvl_stride[l] = ((size_t)fs->pitch[l] * hh + 255) & ~(size_t)255;
        // + 256: the warp-tile pyrDown reads aligned words that may run a few bytes past the last row
        RF_CUDA(h, cudaMalloc(&fs->lvl[l], fs->lvl_stride[l] * count + 256));
        nl = l + 1;
    }
    fs->n_levels = nl;
    if (with_f32) {
        fs->cart_stride = (size_t)h->n * h->n;
        RF_CUDA(h, cudaMalloc(&fs->cart, fs->cart_stride * sizeof(float) * count));
    }
    return RF_OK;
}

void rf_frameset_free(FrameSet* fs) {
    for (int l = 0; l < RF_MAX_LEVELS; ++l)
        if (fs->lvl[l]) cudaFree(fs->lvl[l]);
    if (fs->cart) cudaFree(fs->cart);
    memset(fs, 0, sizeof(*fs));
}

int rf_launch_build_map(rf_handle* h) {
    const int n = h->n;
    const double Kangle = (2.0 * M_PI) / h->cfg.azimuths;          // CV_2PI / ssize.height
    const double Kmag = (double)h->R / (double)h->cfg.range_bins;  // maxRadius / ssize.width
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8);
    k_build_map<<<grd, blk, 0, h->stream>>>(h->map, n, (float)h->R, (float)h->R, Kmag, Kangle, 0);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

template <typename TapT>
static int launch_p2c(rf_handle* h, const TapT* d_src, size_t src_frame_stride, int row_pitch, int col0,
                      const FrameSet& dst, int first, int n_frames, bool write_f32) {
    const int n = h->n;
    if (n % 4) return rf_fail(h, RF_E_BADARG, "cartesian size %d is not a multiple of 4", n);
    if (write_f32 && !dst.cart) return rf_fail(h, RF_E_BADARG, "frame set has no f32 plane");
    dim3 blk(16, 16), grd((n / 4 + 15) / 16, (n + 15) / 16, n_frames);
    if (write_f32)
        k_polar2cart<TapT, true><<<grd, blk, 0, h->stream>>>(d_src, src_frame_stride, row_pitch, col0, h->cfg.azimuths,
                                                             h->cfg.range_bins, h->map, n, dst.cart, dst.cart_stride,
                                                             dst.lvl[0], dst.lvl_stride[0], dst.pitch[0], first);
    else
        k_polar2cart<TapT, false><<<grd, blk, 0, h->stream>>>(d_src, src_frame_stride, row_pitch, col0, h->cfg.azimuths,
                                                              h->cfg.range_bins, h->map, n, nullptr, 0, dst.lvl[0],
                                                              dst.lvl_stride[0], dst.pitch[0], first);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_polar2cart_u8(rf_handle* h, const uint8_t* d_src, size_t src_frame_stride, int row_pitch, int col0,
                            const FrameSet& dst, int first, int n_frames, bool write_f32) {
    return launch_p2c<uint8_t>(h, d_src, src_frame_stride, row_pitch, col0, dst, first, n_frames, write_f32);
}

int rf_launch_polar2cart_f32(rf_handle* h, const float* d_src, int row_pitch, const FrameSet& dst) {
    return launch_p2c<float>(h, d_src, 0, row_pitch, 0, dst, 0, 1, true);
}

int rf_launch_cart_to_u8(rf_handle* h, const FrameSet& fs) {
    size_t count4 = (size_t)h->n * h->n / 4;
    k_cart_to_u8<<<(unsigned)((count4 + 255) / 256), 256, 0, h->stream>>>(fs.cart, fs.lvl[0], count4, h->n / 4, fs.pitch[0]);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_pyramid(rf_handle* h, const FrameSet& fs, int first, int n_frames) {
    for (int l = 1; l < fs.n_levels; ++l) {
        dim3 grd((fs.w[l] + PD_TW - 1) / PD_TW, (fs.h[l] + PD_TH - 1) / PD_TH, n_frames);
        k_pyr_down<<<grd, 256, 0, h->stream>>>(fs.lvl[l - 1], fs.lvl_stride[l - 1], fs.pitch[l - 1], fs.w[l - 1], fs.h[l - 1], fs.lvl[l],
                                               fs.lvl_stride[l], fs.pitch[l], fs.w[l], fs.h[l], first);
        RF_CHECK_LAUNCH(h);
    }
    return RF_OK;
}

int rf_launch_extract(rf_handle* h, const uint8_t* d_raw, float* d_polar) {
    const int W = h->cfg.range_bins;
    dim3 grd((W + 255) / 256, h->cfg.azimuths);
    k_extract<<<grd, 256, 0, h->stream>>>(d_raw, h->cfg.azimuths, h->cfg.raw_width, h->cfg.meta_bytes, W, d_polar);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}


// convertPolarImageToCartesian(imgPolar, logPolarMode=True)      parseData.py:100-135 (flags += cv2.WARP_POLAR_LOG)
// cv::warpPolar's inverse SEMI-LOG map: rho = log(|p - centre| + 1) * W / log(maxRadius).  The map is built into
// scratch for the call (the reference never takes this branch on its live path) and sampled by the same remap kernel.
extern "C" int rf_polar_to_cart_log(rf_handle* h, const float* polar, float* cart_out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !polar || !cart_out) return rf_fail(h, RF_E_BADARG, "rf_polar_to_cart_log: null argument");
    const rf_config& c = h->cfg;
    const int n = h->n;
    if (n % 4) return rf_fail(h, RF_E_BADARG, "cartesian size %d is not a multiple of 4", n);
    const size_t n2 = (size_t)n * n;
    const size_t map_b = (n2 * 4 + 255) & ~(size_t)255, cart_b = (n2 * 4 + 255) & ~(size_t)255, u8_b = (n2 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, map_b + cart_b + u8_b);
    if (rc) return rc;
    uint32_t* d_map = (uint32_t*)h->d_scratch;
    float* d_cart = (float*)((char*)h->d_scratch + map_b);
    uint8_t* d_u8 = (uint8_t*)((char*)h->d_scratch + map_b + cart_b);
    const double Kangle = (2.0 * M_PI) / c.azimuths;
    const double Kmag = log((double)h->R) / (double)c.range_bins;     // std::log(maxRadius) / ssize.width
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8);
    k_build_map<<<grd, blk, 0, h->stream>>>(d_map, n, (float)h->R, (float)h->R, Kmag, Kangle, 1);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(h->d_polar, polar, (size_t)c.azimuths * c.range_bins * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    dim3 blk2(16, 16), grd2((n / 4 + 15) / 16, (n + 15) / 16, 1);
    k_polar2cart<float, true><<<grd2, blk2, 0, h->stream>>>(h->d_polar, 0, c.range_bins, 0, c.azimuths, c.range_bins, d_map, n, d_cart, 0,
                                                            d_u8, 0, n, 0);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(cart_out, d_cart, n2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}
