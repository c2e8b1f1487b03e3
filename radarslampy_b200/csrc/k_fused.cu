// k_fused.cu — the batch image path: raw scans -> u8 Cartesian level 0 + LK pyramid, for many
// frames at once.  Bit-identical to k_polar2cart + k_pyr_down (k_image.cu), restructured around
// what bounds that pair on B200: issue slots, not HBM.
//
// Replaces (reference file:line), for a whole batch of frames:
//   parseData.py:17-53,100-135   extractDataFromRadarImage + convertPolarImageToCartesian (cv2.warpPolar)
//   getTransformKLT.py:356-357   (img * 255).astype(np.uint8)
//   cv2.buildOpticalFlowPyramid  (pyrDown 5x5, REFLECT_101) inside calcOpticalFlowPyrLK (getTransformKLT.py:359)
//
//   k_interleave     raw [F][A][pitch] u8 -> [F/4][A][Wp] words, byte f of a word = frame 4g+f.
//                    One 32-bit gather then serves a bilinear tap of FOUR frames.
//   k_build_map2     per-pixel geometry record (8 B): tap word offset + 5-bit fractions + tap
//                    validity, derived once per handle from the fixed-point inverse map.
//   k_scan_to_l0l1   CTA = 128x32 level-0 pixels (+ pyrDown halo) x 8 frames.  Each thread decodes
//                    the geometry of a pixel ONCE and applies it to 8 frames; u8 -> f32/255 goes
//                    through a bank-replicated shared-memory table (exact IEEE quotient, no
//                    conflicts); the level-0 tile never leaves shared memory before level 1 is
//                    built from it (dp4a horizontal taps, packed-u16 vertical taps).
//   k_pyr_down_w     the same warp-tile pyrDown for levels >= 2, input from global memory.
#include "common.cuh"

#define FT_TW1 64
#define FT_TH1 16
#define FT_RW (2 * FT_TW1 + 4)   // 132 region columns: level-0 x in [2*ox1 - 2, 2*ox1 + 130)
#define FT_RH (2 * FT_TH1 + 3)   // 35 region rows:     level-0 y in [2*oy1 - 2, 2*oy1 + 33)
#define FT_RWW (FT_RW / 4)       // 33 words per region row
#define FT_FR 8                  // frames per CTA (two interleave groups)
#define FT_LUT_WORDS (256 * 32)
#define FT_TILE_WORDS (FT_RH * FT_RWW)
#define FT_HS_WORDS (FT_RH * 32)
#define FT_SMEM_BYTES ((FT_LUT_WORDS + FT_FR * FT_TILE_WORDS + 8 * FT_HS_WORDS) * 4)

__device__ __forceinline__ int reflect101_safe(int p, int len) {
    p = min(max(p, -(len - 1)), 2 * len - 2);
    return reflect101(p, len);
}

// ------------------------------------------------------------------------------------
// frame interleave: 4 frames -> one word per polar sample
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_interleave(const uint8_t* __restrict__ raw, size_t frame_stride, int pitch, int A, int W, int n_frames,
             uint32_t* __restrict__ out, int Wp) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 samples
    const int a = blockIdx.y, g = blockIdx.z;
    if (x4 * 4 >= Wp) return;
    uint32_t w[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const int fr = 4 * g + f;
        // raw rows are 16-byte aligned and hold only the power bins; columns >= W are padding
        w[f] = (fr < n_frames && x4 * 4 < pitch)
                   ? __ldg(reinterpret_cast<const uint32_t*>(raw + (size_t)fr * frame_stride + (size_t)a * pitch) + x4) : 0u;
    }
    // 4x4 byte transpose: out[k] = (w0.bk, w1.bk, w2.bk, w3.bk)
    const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[0], w[1], 0x7362);
    const uint32_t t2 = __byte_perm(w[2], w[3], 0x5140), t3 = __byte_perm(w[2], w[3], 0x7362);
    uint4 o;
    o.x = __byte_perm(t0, t2, 0x5410); o.y = __byte_perm(t0, t2, 0x7632);
    o.z = __byte_perm(t1, t3, 0x5410); o.w = __byte_perm(t1, t3, 0x7632);
    // samples beyond the used range read as 0 (they only ever meet zero weights, keep them defined)
    const int x = 4 * x4;
    if (x + 0 >= W) o.x = 0; if (x + 1 >= W) o.y = 0; if (x + 2 >= W) o.z = 0; if (x + 3 >= W) o.w = 0;
    reinterpret_cast<uint4*>(out + ((size_t)g * A + a) * Wp)[x4] = o;
}

// ------------------------------------------------------------------------------------
// geometry records
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_build_map2(const uint32_t* __restrict__ map, size_t count, int A, int W, int Wp, uint2* __restrict__ map2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t m = map[i];
    const int sx = m & 0x1FFFF, sy = m >> 17;
    const int ix = sx >> 5, iy = sy >> 5;
    int r0 = iy - 1; r0 = r0 < 0 ? r0 + A : (r0 >= A ? r0 - A : r0);
    int r1 = iy;     r1 = r1 >= A ? r1 - A : r1;
    const bool y0ok = iy < A + 2, y1ok = iy + 1 < A + 2;
    const bool x0ok = ix < W, x1ok = ix + 1 < W;
    const unsigned flags = (unsigned)(y0ok && x0ok) | ((unsigned)(y0ok && x1ok) << 1) | ((unsigned)(y1ok && x0ok) << 2) |
                           ((unsigned)(y1ok && x1ok) << 3);
    const int ixc = min(ix, Wp - 2);
    uint2 o;
    o.x = (unsigned)r0 * (unsigned)Wp + (unsigned)ixc;
    o.y = (unsigned)(sx & 31) | ((unsigned)(sy & 31) << 5) | (flags << 10) | ((unsigned)(r1 != r0 + 1) << 14);
    map2[i] = o;
}

// ------------------------------------------------------------------------------------
// warp-level pyrDown of one tile held in shared memory.
//   tile : [FT_RH][FT_RWW] words = 35 x 132 source bytes; byte (i, j) = source (2*oy1 - 2 + i, 2*ox1 - 2 + j)
//   hs   : [FT_RH][32] words scratch (two u16 horizontal sums per word)
// Writes the 16 x 64 destination tile at (oy1, ox1).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_pyr_tile(const uint32_t* __restrict__ tile, uint32_t* __restrict__ hs,
                                              uint8_t* __restrict__ dst, int dw, int dh, int ox1, int oy1, int lane) {
#pragma unroll 5
    for (int r = 0; r < FT_RH; ++r) {
        const uint32_t w0 = tile[r * FT_RWW + lane], w1 = tile[r * FT_RWW + lane + 1];
        const uint32_t he = __dp4a(w1, 0x00000001u, __dp4a(w0, 0x04060401u, 0u));   // columns 4k .. 4k+4
        const uint32_t ho = __dp4a(w1, 0x00010406u, __dp4a(w0, 0x04010000u, 0u));   // columns 4k+2 .. 4k+6
        hs[r * 32 + lane] = he | (ho << 16);
    }
    __syncwarp();
    const int x = ox1 + 2 * lane;
    const bool even_pitch = (dw & 1) == 0;
#pragma unroll 4
    for (int r1 = 0; r1 < FT_TH1; ++r1) {
        const uint32_t* p = hs + (2 * r1) * 32 + lane;
        // both halves stay below 2^16 (255 * 256 + 128), so the packed sum never carries across
        const uint32_t s = p[0] + 4u * p[32] + 6u * p[64] + 4u * p[96] + p[128] + 0x00800080u;
        const int y = oy1 + r1;
        if (y < dh && x < dw) {
            const uint32_t lo = (s >> 8) & 0xFFu, hi = s >> 24;
            uint8_t* q = dst + (size_t)y * dw + x;
            if (even_pitch) *reinterpret_cast<uint16_t*>(q) = (uint16_t)(lo | (hi << 8));
            else { q[0] = (uint8_t)lo; if (x + 1 < dw) q[1] = (uint8_t)hi; }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------
// raw (interleaved) -> level 0 + level 1
// ------------------------------------------------------------------------------------
struct FusedArgs {
    const uint32_t* rawi; size_t group_stride;   // words per interleave group plane (A * Wp)
    int Wp, A;
    const uint2* map2; int n;
    uint8_t* l0; size_t l0_stride;
    uint8_t* l1; size_t l1_stride; int w1, h1;
    int n_frames;
};

__device__ __forceinline__ float lut_tap(const float* __restrict__ lut_lane, uint32_t t, int f) {
    return lut_lane[((t >> (8 * f)) & 0xFFu) << 5];
}

__global__ void __launch_bounds__(256, 2) k_scan_to_l0l1(const FusedArgs a) {
    extern __shared__ uint32_t smem[];
    float* lut = reinterpret_cast<float*>(smem);             // [256][32]: entry b replicated once per bank
    uint32_t* tiles = smem + FT_LUT_WORDS;                   // [FT_FR][FT_RH][FT_RWW]
    uint32_t* hs_all = tiles + FT_FR * FT_TILE_WORDS;        // [8 warps][FT_RH][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < FT_LUT_WORDS; i += 256) lut[i] = __fdiv_rn((float)(i >> 5), 255.0f);   // parseData.py:43
    __syncthreads();
    const float* lut_lane = lut + lane;
    const int ox1 = blockIdx.x * FT_TW1, oy1 = blockIdx.y * FT_TH1;
    const int f0 = blockIdx.z * FT_FR;
    const uint32_t* __restrict__ g0 = a.rawi + (size_t)(2 * blockIdx.z) * a.group_stride;
    const uint32_t* __restrict__ g1 = g0 + a.group_stride;
    const unsigned AWp = (unsigned)a.A * (unsigned)a.Wp;
    const int n = a.n;

    for (int item = tid; item < FT_RH * FT_RWW; item += 256) {
        const int ry = item / FT_RWW, rx4 = item - ry * FT_RWW;
        const int gy = reflect101_safe(2 * oy1 - 2 + ry, n);
        uint32_t word[FT_FR];
#pragma unroll
        for (int f = 0; f < FT_FR; ++f) word[f] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int gx = reflect101_safe(2 * ox1 - 2 + 4 * rx4 + k, n);
            const uint2 m = __ldg(a.map2 + (size_t)gy * n + gx);
            const int fx = m.y & 31, fy = (m.y >> 5) & 31;
            const unsigned fl = m.y >> 10;
            const unsigned o0 = m.x, o1 = m.x + (unsigned)a.Wp - ((fl & 16u) ? AWp : 0u);
            // (1 - fy)(1 - fx) etc. are exact multiples of 2^-10, as cv2 computes them
            const float w00 = (fl & 1u) ? (float)((32 - fy) * (32 - fx)) * 0.0009765625f : 0.0f;
            const float w01 = (fl & 2u) ? (float)((32 - fy) * fx) * 0.0009765625f : 0.0f;
            const float w10 = (fl & 4u) ? (float)(fy * (32 - fx)) * 0.0009765625f : 0.0f;
            const float w11 = (fl & 8u) ? (float)(fy * fx) * 0.0009765625f : 0.0f;
            const uint32_t t00[2] = {__ldg(g0 + o0), __ldg(g1 + o0)}, t01[2] = {__ldg(g0 + o0 + 1), __ldg(g1 + o0 + 1)};
            const uint32_t t10[2] = {__ldg(g0 + o1), __ldg(g1 + o1)}, t11[2] = {__ldg(g0 + o1 + 1), __ldg(g1 + o1 + 1)};
#pragma unroll
            for (int f = 0; f < FT_FR; ++f) {
                const int g = f >> 2, b = f & 3;
                float acc = __fmul_rn(lut_tap(lut_lane, t00[g], b), w00);
                acc = __fadd_rn(acc, __fmul_rn(lut_tap(lut_lane, t01[g], b), w01));
                acc = __fadd_rn(acc, __fmul_rn(lut_tap(lut_lane, t10[g], b), w10));
                acc = __fadd_rn(acc, __fmul_rn(lut_tap(lut_lane, t11[g], b), w11));
                // (img * 255).astype(uint8): f32 product, truncation; 2^23 + x rounded toward zero keeps floor(x) in the low byte
                const uint32_t u = __float_as_uint(__fadd_rz(__fmul_rn(acc, 255.0f), 8388608.0f));
                word[f] = __byte_perm(word[f], u, k == 0 ? 0x3214 : k == 1 ? 0x3240 : k == 2 ? 0x3410 : 0x4210);
            }
        }
#pragma unroll
        for (int f = 0; f < FT_FR; ++f) tiles[f * FT_TILE_WORDS + item] = word[f];
    }
    __syncthreads();

    // one warp per frame from here on
    const int frame = f0 + warp;
    if (frame >= a.n_frames) return;
    const uint32_t* tile = tiles + warp * FT_TILE_WORDS;
    {   // level 0: region rows 2..33, region bytes 2..129 -> 32 aligned words per row
        uint8_t* l0 = a.l0 + (size_t)frame * a.l0_stride;
        const int x = 2 * ox1 + 4 * lane;
#pragma unroll 4
        for (int r = 0; r < 2 * FT_TH1; ++r) {
            const int y = 2 * oy1 + r;
            const uint32_t lo = tile[(r + 2) * FT_RWW + lane], hi = tile[(r + 2) * FT_RWW + lane + 1];
            if (y < n && x < n) *reinterpret_cast<uint32_t*>(l0 + (size_t)y * n + x) = __funnelshift_r(lo, hi, 16);
        }
    }
    warp_pyr_tile(tile, hs_all + warp * FT_HS_WORDS, a.l1 + (size_t)frame * a.l1_stride, a.w1, a.h1, ox1, oy1, lane);
}

// ------------------------------------------------------------------------------------
// pyrDown for the higher levels: each warp loads its 35 x 132 source tile from global memory
// (REFLECT_101), then the same shared-memory pass.  4 warps per CTA, one tile each.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pyr_down_w(const uint8_t* __restrict__ src, size_t src_stride, int sw, int sh, uint8_t* __restrict__ dst,
             size_t dst_stride, int dw, int dh, int tiles_x, int tiles_per_frame, int n_tiles) {
    __shared__ uint32_t s_tile[4][FT_TILE_WORDS];
    __shared__ uint32_t s_hs[4][FT_HS_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * 4 + warp;
    if (t >= n_tiles) return;
    const int frame = t / tiles_per_frame, tt = t - frame * tiles_per_frame;
    const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
    const int ox1 = tx * FT_TW1, oy1 = ty * FT_TH1;
    const uint8_t* __restrict__ s = src + (size_t)frame * src_stride;
    uint32_t* tw = s_tile[warp];
    const int x0 = 2 * ox1 - 2;
    // Region word k (k = lane, plus word 32 on lane 0) holds source bytes x0 + 4k .. x0 + 4k + 3 of the row.
    // Words that lie fully inside the row are cut out of two ALIGNED 32-bit loads with a funnel shift (rows are
    // only byte-aligned: widths are odd at the higher levels); the few words that straddle an image edge fall
    // back to REFLECT_101 byte loads.  Aligned loads may run up to 7 bytes past the row end: into the next row,
    // or into the padding every level allocation carries (rf_frameset_alloc).
    const int xk = x0 + 4 * lane, xk32 = x0 + 128;
    const bool in_k = xk >= 0 && xk + 3 < sw, in_32 = xk32 + 3 < sw;
#pragma unroll 5
    for (int r = 0; r < FT_RH; ++r) {
        const uint8_t* row = s + (size_t)reflect101_safe(2 * oy1 - 2 + r, sh) * sw;
        const uint8_t* p = row + x0;
        const unsigned m = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3u);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(p - m);
        // (the left-edge tile starts 2 bytes before the row: never read in front of the frame)
        const uint32_t a = (reinterpret_cast<const uint8_t*>(ap + lane) >= s) ? __ldg(ap + lane) : 0u;
        const uint32_t b = __ldg(ap + 32 + (lane & 1));                     // words 32, 33 (x0 + 128 > 0 always)
        uint32_t nxt = __shfl_down_sync(0xffffffffu, a, 1);
        const uint32_t w32 = __shfl_sync(0xffffffffu, b, 0), w33 = __shfl_sync(0xffffffffu, b, 1);
        if (lane == 31) nxt = w32;
        uint32_t v = __funnelshift_r(a, nxt, 8 * m);
        if (!in_k) {
            v = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) v |= (uint32_t)__ldg(row + reflect101_safe(xk + q, sw)) << (8 * q);
        }
        tw[r * FT_RWW + lane] = v;
        if (lane == 0) {
            uint32_t v2 = __funnelshift_r(w32, w33, 8 * m);
            if (!in_32) {
                v2 = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) v2 |= (uint32_t)__ldg(row + reflect101_safe(xk32 + q, sw)) << (8 * q);
            }
            tw[r * FT_RWW + 32] = v2;
        }
    }
    __syncwarp();
    warp_pyr_tile(s_tile[warp], s_hs[warp], dst + (size_t)frame * dst_stride, dw, dh, ox1, oy1, lane);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
int rf_fused_wp(const rf_handle* h) { return ((h->cfg.range_bins + 1) + 3) & ~3; }   // >= W + 1, multiple of 4

int rf_launch_build_map2(rf_handle* h) {
    const size_t count = (size_t)h->n * h->n;
    k_build_map2<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->map, count, h->cfg.azimuths, h->cfg.range_bins,
                                                                          rf_fused_wp(h), h->map2);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// d_raw: frames of [A][pitch] power bins (no metadata), pitch a multiple of 16 and >= Wp
int rf_launch_interleave(rf_handle* h, const uint8_t* d_raw, size_t frame_stride, int pitch, int n_frames, uint32_t* d_out) {
    const int Wp = rf_fused_wp(h);
    const int groups = 2 * ((n_frames + FT_FR - 1) / FT_FR);   // zero-filled up to a multiple of 8 frames
    dim3 grd((Wp / 4 + 255) / 256, h->cfg.azimuths, groups);
    k_interleave<<<grd, 256, 0, h->stream>>>(d_raw, frame_stride, pitch, h->cfg.azimuths, h->cfg.range_bins, n_frames, d_out, Wp);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

size_t rf_interleave_words(const rf_handle* h, int max_frames) {
    return (size_t)(2 * ((max_frames + FT_FR - 1) / FT_FR)) * h->cfg.azimuths * rf_fused_wp(h);
}

int rf_launch_scan_to_l0l1(rf_handle* h, const uint32_t* d_rawi, const FrameSet& fs, int n_frames) {
    if (h->n % 4) return rf_fail(h, RF_E_BADARG, "cartesian size %d is not a multiple of 4", h->n);
    if (fs.n_levels < 2) return rf_fail(h, RF_E_BADARG, "fused image path needs at least two pyramid levels");
    static bool attr_set = false;
    if (!attr_set) {
        RF_CUDA(h, cudaFuncSetAttribute(k_scan_to_l0l1, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
        attr_set = true;
    }
    FusedArgs a;
    a.rawi = d_rawi; a.Wp = rf_fused_wp(h); a.A = h->cfg.azimuths; a.group_stride = (size_t)a.A * a.Wp;
    a.map2 = h->map2; a.n = h->n;
    a.l0 = fs.lvl[0]; a.l0_stride = fs.lvl_stride[0];
    a.l1 = fs.lvl[1]; a.l1_stride = fs.lvl_stride[1]; a.w1 = fs.w[1]; a.h1 = fs.h[1];
    a.n_frames = n_frames;
    dim3 grd((fs.w[1] + FT_TW1 - 1) / FT_TW1, (fs.h[1] + FT_TH1 - 1) / FT_TH1, (n_frames + FT_FR - 1) / FT_FR);
    k_scan_to_l0l1<<<grd, 256, FT_SMEM_BYTES, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// levels first_level .. n_levels-1 from their predecessors
int rf_launch_pyr_levels(rf_handle* h, const FrameSet& fs, int first_level, int n_frames) {
    for (int l = first_level; l < fs.n_levels; ++l) {
        const int tiles_x = (fs.w[l] + FT_TW1 - 1) / FT_TW1, tiles_y = (fs.h[l] + FT_TH1 - 1) / FT_TH1;
        const int n_tiles = tiles_x * tiles_y * n_frames;
        k_pyr_down_w<<<(n_tiles + 3) / 4, 128, 0, h->stream>>>(fs.lvl[l - 1], fs.lvl_stride[l - 1], fs.w[l - 1], fs.h[l - 1],
                                                               fs.lvl[l], fs.lvl_stride[l], fs.w[l], fs.h[l], tiles_x,
                                                               tiles_x * tiles_y, n_tiles);
        RF_CHECK_LAUNCH(h);
    }
    return RF_OK;
}
