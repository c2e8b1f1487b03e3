// k_fused.cu — the batch image path: raw scans -> u8 Cartesian level 0 + LK pyramid, for many
// frames at once.  Bit-identical to k_polar2cart + k_pyr_down (k_image.cu), restructured around
// what bounds that pair on B200: issue slots and L1 wavefronts, not HBM.
//
// Replaces (reference file:line), for a whole batch of frames:
//   parseData.py:17-53,100-135   extractDataFromRadarImage + convertPolarImageToCartesian (cv2.warpPolar)
//   getTransformKLT.py:356-357   (img * 255).astype(np.uint8)
//   cv2.buildOpticalFlowPyramid  (pyrDown 5x5, REFLECT_101) inside calcOpticalFlowPyrLK (getTransformKLT.py:359)
//
//   k_interleave16   raw [F][A][pitch] u8 -> [F/16][A][Wp] uint4, byte f of a sample = frame 16g+f.
//                    One 128-bit gather then serves a bilinear tap of SIXTEEN frames.
//   k_build_map2     per-pixel geometry record (8 B): tap sample offset + 5-bit fractions + tap
//                    validity, derived once per handle from the fixed-point inverse map.
//   k_scan16_to_l0l1 CTA = 128x16 level-0 pixels (+ pyrDown halo) x 16 frames, one pixel per lane
//                    (neighbouring lanes gather neighbouring samples).  The u8 result is
//                    floor(V / 1024), V = sum of tap byte x 10-bit integer weight, whenever V is not
//                    a multiple of 1024: the f32 chain cv2 evaluates is then within 1.1e-4 of V/1024
//                    and cannot cross an integer (see `exactness` below).  V is formed for two frames
//                    per instruction in packed 16-bit lanes (horizontal) + one dp2a per frame
//                    (vertical).  The 1-3 % of (pixel, frame) pairs with V = 0 mod 1024 re-run cv2's
//                    exact f32 chain.  The level-0 tile never leaves shared memory before level 1 is
//                    built from it (dp4a horizontal taps, packed-u16 vertical taps).
//   k_pyr_down_w     the same warp-tile pyrDown for levels >= 2, input from global memory.
//
// exactness.  cv2: out = ((S00*w00 + S01*w01) + S10*w10) + S11*w11 with S = fl(b / 255), w = m / 1024
// (m = (32 - fy | fy)(32 - fx | fx), exact), every product and sum rounded to f32, then
// u8 = trunc(fl(out * 255)).  Nine roundings of relative size 2^-24 on values <= 1, times 255:
// |fl(out*255) - V/1024| < 1.1e-4 < 1/1024.  V/1024 has a fractional part that is a multiple of 1/1024,
// so unless that part is 0 the truncation of both is the same integer.  (tests: bit-exact against the
// plain-C restatement of cv2 on real Oxford scans, where 2.8 % of the pixels take the exact path.)
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

#define FT_TW1 64
// Tile height (level-1 rows per CTA) and resident CTAs per SM of the scan kernel.  Measured on B200 (256 frames):
//   TH1 = 16, 3 CTAs/SM (74 KB tiles, 72 regs): scan 1.158 ms, upper levels 0.417 ms, 100.0 k frames/s
//   TH1 =  8, 4 CTAs/SM (40 KB tiles, 64 regs): scan 1.106 ms, upper levels 0.384 ms, 105.0 k frames/s  <- default
//   TH1 =  8, 5 CTAs/SM (48 regs, spills):       scan 1.274 ms                          96.8 k frames/s
// The shorter tile pays 19/16 instead of 35/32 halo rows but lifts occupancy from 37 % to 50 %.
#ifndef FT_TH1
#define FT_TH1 8
#endif
#ifndef FT_SCAN_MIN_BLOCKS
#define FT_SCAN_MIN_BLOCKS 4
#endif
#define FT_RW (2 * FT_TW1 + 4)   // 132 region columns: level-0 x in [2*ox1 - 2, 2*ox1 + 130)
#define FT_RH (2 * FT_TH1 + 3)   // 19 region rows:     level-0 y in [2*oy1 - 2, 2*oy1 + 2*FT_TH1 + 1)
#define FT_RWW (FT_RW / 4)       // 33 words per region row (plain byte tiles of k_pyr_down_w)
#define FT_ITEMS (FT_RH * FT_RW) // region pixels of one CTA
#define FT_FR 16                 // frames per CTA (one interleave group)
#define FT_PAIRS (FT_FR / 2)
// Shared tile of the scan kernel: [FT_PAIRS][FT_ITEMS] u16, byte f & 1 of element (f >> 1, item) = frame f of region pixel item.
// The blend stores one u16 per frame pair and pixel (neighbouring lanes -> neighbouring elements: one wavefront); the
// pyramid pass reads four pixels of a pair as one 64-bit word and splits the two frames with a PRMT each.
#define FT_TILE_BYTES (FT_PAIRS * FT_ITEMS * 2)
#define FT_LIST_CAP 64           // deferred (pixel, frame flags) entries per warp
#define FT_LIST_WORDS 7          // item, geometry words x / y, four words of per-frame flag bytes
#define FT_SMEM_BYTES (FT_TILE_BYTES + 8 * FT_LIST_WORDS * FT_LIST_CAP * 4 + 1024)   // + fl(b / 255) table
#define FT_OFF_BITS 22           // sample offsets inside one interleave group (A * Wp < 4 M)
#define FT_OFF_MASK ((1u << FT_OFF_BITS) - 1u)

__device__ __forceinline__ int reflect101_safe(int p, int len) {
    p = min(max(p, -(len - 1)), 2 * len - 2);
    return reflect101(p, len);
}

// ------------------------------------------------------------------------------------
// frame interleave: 16 frames -> one uint4 per polar sample
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_interleave16(const uint8_t* __restrict__ raw, size_t frame_stride, int pitch, int A, int W, int n_frames,
               uint4* __restrict__ out, int Wp, const int32_t* __restrict__ frame_sel) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 samples
    const int a = blockIdx.y, g = blockIdx.z;
    if (x4 * 4 >= Wp) return;
    const int sel_base = frame_sel ? frame_sel[0] : 0, sel_stride = frame_sel ? frame_sel[1] : 1;
    uint32_t o[4][4];   // o[k][c] = sample 4*x4 + k, frames 4c .. 4c+3
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int fr = FT_FR * g + 4 * c + f;
            // raw rows are 16-byte aligned and hold only the power bins; columns >= W are padding
            w[f] = (fr < n_frames && x4 * 4 < pitch)
                       ? __ldg(reinterpret_cast<const uint32_t*>(raw + (size_t)(sel_base + fr * sel_stride) * frame_stride + (size_t)a * pitch) + x4) : 0u;
        }
        // 4x4 byte transpose: t[k] = (w0.bk, w1.bk, w2.bk, w3.bk)
        const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[0], w[1], 0x7362);
        const uint32_t t2 = __byte_perm(w[2], w[3], 0x5140), t3 = __byte_perm(w[2], w[3], 0x7362);
        const uint32_t s0 = __byte_perm(t0, t2, 0x5410), s1 = __byte_perm(t0, t2, 0x7632);
        const uint32_t s2 = __byte_perm(t1, t3, 0x5410), s3 = __byte_perm(t1, t3, 0x7632);
        o[0][c] = s0; o[1][c] = s1; o[2][c] = s2; o[3][c] = s3;
    }
    // samples beyond the used range read as 0 (they only ever meet zero weights, keep them defined)
    const int x = 4 * x4;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* q = out + ((size_t)g * A + a) * Wp + x;
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = (x + k >= W) ? z : make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
}

// ------------------------------------------------------------------------------------
// geometry records (16 B per Cartesian pixel, built once per handle from the fixed-point inverse map)
//   x = off0 | fx << 22 | fy << 27     sample offset of tap (iy, ix) inside an interleave group, 5-bit fractions
//   y = off1 | valid << 22 | ident << 26    offset of tap (iy + 1, ix); tap validity (bit 0: 00, 1: 01, 2: 10, 3: 11)
//   z = 64 * (m00 | m01 << 16), w = 64 * (m10 | m11 << 16)    m = (32 - fy | fy)(32 - fx | fx), 0 for a tap outside the scan
// ident marks fx = fy = 0 with tap 00 valid: its weight 64 * 1024 does not fit 16 bits (and every V is a multiple of 1024
// there), so all sixteen frames of such a pixel go through the exact chain.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_build_map2(const uint32_t* __restrict__ map, size_t count, int A, int W, int Wp, uint4* __restrict__ geo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t m = map[i];
    const int sx = m & 0x1FFFF, sy = m >> 17;
    const int ix = sx >> 5, iy = sy >> 5;
    int r0 = iy - 1; r0 = r0 < 0 ? r0 + A : (r0 >= A ? r0 - A : r0);
    int r1 = iy;     r1 = r1 >= A ? r1 - A : r1;
    const bool y0ok = iy < A + 2, y1ok = iy + 1 < A + 2;
    const bool x0ok = ix < W, x1ok = ix + 1 < W;
    const unsigned valid = (unsigned)(y0ok && x0ok) | ((unsigned)(y0ok && x1ok) << 1) | ((unsigned)(y1ok && x0ok) << 2) |
                           ((unsigned)(y1ok && x1ok) << 3);
    const int ixc = min(ix, Wp - 2);
    const unsigned fx = sx & 31, fy = sy & 31;
    const unsigned ident = (fx == 0u && fy == 0u && (valid & 1u)) ? 1u : 0u;
    unsigned m00 = (32u - fy) * (32u - fx), m01 = (32u - fy) * fx, m10 = fy * (32u - fx), m11 = fy * fx;
    if (!(valid & 1u) || ident) m00 = 0u;
    if (!(valid & 2u)) m01 = 0u;
    if (!(valid & 4u)) m10 = 0u;
    if (!(valid & 8u)) m11 = 0u;
    uint4 o;
    // a pixel with no valid tap points at sample 0: its loads are one broadcast line and meet zero weights
    o.x = (valid ? (unsigned)r0 * (unsigned)Wp + (unsigned)ixc : 0u) | (fx << FT_OFF_BITS) | (fy << (FT_OFF_BITS + 5));
    o.y = (valid ? (unsigned)r1 * (unsigned)Wp + (unsigned)ixc : 0u) | (valid << FT_OFF_BITS) | (ident << (FT_OFF_BITS + 4));
    o.z = 64u * (m00 | (m01 << 16));
    o.w = 64u * (m10 | (m11 << 16));
    geo[i] = o;
}

// ------------------------------------------------------------------------------------
// One row of the warp-level pyrDown: lane k owns destination columns 2k, 2k + 1 of the tile; w0 / w1 are source bytes
// 4k .. 4k+3 / 4k+4 .. 4k+7 of the row.  Returns the horizontal 1-4-6-4-1 sums of the two columns as two u16.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pyr_hsum(uint32_t w0, uint32_t w1) {
    const uint32_t he = __dp4a(w1, 0x00000001u, __dp4a(w0, 0x04060401u, 0u));   // columns 4k .. 4k+4
    const uint32_t ho = __dp4a(w1, 0x00010406u, __dp4a(w0, 0x04010000u, 0u));   // columns 4k+2 .. 4k+6
    return he | (ho << 16);
}
// vertical 1-4-6-4-1 of five packed row sums, + 128 >> 8, as two bytes (both halves stay below 2^16: 255 * 256 + 128)
__device__ __forceinline__ uint32_t pyr_vsum(uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t h4) {
    const uint32_t s = h0 + 4u * h1 + 6u * h2 + 4u * h3 + h4 + 0x00800080u;
    return __byte_perm(s, 0u, 0x4431);   // (s >> 8) & 0xFF | (s >> 24) << 8
}

// ------------------------------------------------------------------------------------
// warp-level pyrDown of one plain byte tile held in shared memory (k_pyr_down_w).
//   tile : [rows][tile_ww] words; byte (i, j) = source (2*oy1 - 2 + i, 2*ox1 - 2 + j)
// Writes the TH1 x 64 destination tile at (oy1, ox1); destination rows are `dp` bytes apart (a multiple of 16, so the
// two-byte store of an odd last column lands in row padding).
// ------------------------------------------------------------------------------------
template <bool INTERIOR, int TH1>
__device__ __forceinline__ void warp_pyr_tile_impl(const uint32_t* __restrict__ tile, int tile_ww, uint8_t* __restrict__ dst, int dp, int dw,
                                                   int dh, int ox1, int oy1, int lane) {
    const int x = ox1 + 2 * lane;
    const bool x_ok = INTERIOR || x < dw;
    const int rows1 = INTERIOR ? TH1 : min(TH1, dh - oy1);
    uint8_t* q1 = dst + (size_t)oy1 * dp + x;
    const uint32_t* tp = tile + lane;
    uint32_t h[5];
#pragma unroll
    for (int r = 0; r < 2 * TH1 + 3; ++r) {
        h[r % 5] = pyr_hsum(tp[r * tile_ww], tp[r * tile_ww + 1]);
        if (r >= 4 && (r & 1) == 0) {
            const uint32_t px2 = pyr_vsum(h[(r - 4) % 5], h[(r - 3) % 5], h[(r - 2) % 5], h[(r - 1) % 5], h[r % 5]);
            if (INTERIOR || ((r - 4) / 2 < rows1 && x_ok)) *reinterpret_cast<uint16_t*>(q1) = (uint16_t)px2;
            q1 += dp;
        }
    }
}

template <int TH1>
__device__ __forceinline__ void warp_pyr_tile(const uint32_t* __restrict__ tile, int tile_ww, uint8_t* __restrict__ dst, int dp, int dw, int dh,
                                              int ox1, int oy1, int lane) {
    if (ox1 + FT_TW1 <= dw && oy1 + TH1 <= dh) warp_pyr_tile_impl<true, TH1>(tile, tile_ww, dst, dp, dw, dh, ox1, oy1, lane);
    else warp_pyr_tile_impl<false, TH1>(tile, tile_ww, dst, dp, dw, dh, ox1, oy1, lane);
}

// ------------------------------------------------------------------------------------
// raw (interleaved) -> level 0 + level 1
// ------------------------------------------------------------------------------------
struct FusedArgs {
    const uint4* rawi; size_t group_stride;   // samples per interleave group plane (A * Wp)
    int Wp, A;
    const uint4* geo; int n;
    uint8_t* l0; size_t l0_stride; int p0;            // level 0: frame stride, row pitch
    uint8_t* l1; size_t l1_stride; int p1, w1, h1;    // level 1
    int n_frames;
};

__device__ float g_lut255[256];   // fl(b / 255), parseData.py:43 (filled by k_build_lut at handle creation)

__global__ void k_build_lut() {
    g_lut255[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
}

// the two frames of one pair of the scan tile, pyrDown'ed and stored together (level 0 rows 2 .. 2*FT_TH1+1 of the tile as
// 32 aligned words per row and frame; level 1 as two bytes per lane, row and frame)
template <bool INTERIOR>
__device__ __forceinline__ void warp_pyr_pair_impl(const uint16_t* __restrict__ tp, const FusedArgs& a, int frame, int nb, int ox1, int oy1,
                                                   int lane) {
    const int x = ox1 + 2 * lane;
    const bool x_ok = INTERIOR || x < a.w1, x0_ok = INTERIOR || 2 * ox1 + 4 * lane < a.n;
    const int rows0 = INTERIOR ? 2 * FT_TH1 : min(2 * FT_TH1, a.n - 2 * oy1);   // level-0 rows of this tile inside the image
    const int rows1 = INTERIOR ? FT_TH1 : min(FT_TH1, a.h1 - oy1);
    uint8_t* q0 = a.l0 + (size_t)frame * a.l0_stride + (size_t)(2 * oy1) * a.p0 + 2 * ox1 + 4 * lane;
    uint8_t* q1 = a.l1 + (size_t)frame * a.l1_stride + (size_t)oy1 * a.p1 + x;
    const bool two = nb > 1;
    const uint2* rp = reinterpret_cast<const uint2*>(tp) + lane;   // four pixels x two frames per 64-bit word
    uint32_t hA[5], hB[5];
#pragma unroll
    for (int r = 0; r < FT_RH; ++r) {
        const uint2 u = rp[r * FT_RWW], v = rp[r * FT_RWW + 1];
        const uint32_t a0 = __byte_perm(u.x, u.y, 0x6420), b0 = __byte_perm(u.x, u.y, 0x7531);
        const uint32_t a1 = __byte_perm(v.x, v.y, 0x6420), b1 = __byte_perm(v.x, v.y, 0x7531);
        if (r >= 2 && r < 2 + 2 * FT_TH1) {
            if (INTERIOR || (r - 2 < rows0 && x0_ok)) {
                *reinterpret_cast<uint32_t*>(q0) = __funnelshift_r(a0, a1, 16);
                if (two) *reinterpret_cast<uint32_t*>(q0 + a.l0_stride) = __funnelshift_r(b0, b1, 16);
            }
            q0 += a.p0;
        }
        hA[r % 5] = pyr_hsum(a0, a1);
        hB[r % 5] = pyr_hsum(b0, b1);
        if (r >= 4 && (r & 1) == 0) {
            const uint32_t pa = pyr_vsum(hA[(r - 4) % 5], hA[(r - 3) % 5], hA[(r - 2) % 5], hA[(r - 1) % 5], hA[r % 5]);
            const uint32_t pb = pyr_vsum(hB[(r - 4) % 5], hB[(r - 3) % 5], hB[(r - 2) % 5], hB[(r - 1) % 5], hB[r % 5]);
            if (INTERIOR || ((r - 4) / 2 < rows1 && x_ok)) {
                *reinterpret_cast<uint16_t*>(q1) = (uint16_t)pa;
                if (two) *reinterpret_cast<uint16_t*>(q1 + a.l1_stride) = (uint16_t)pb;
            }
            q1 += a.p1;
        }
    }
}

// the four taps (16 frames each) of one pixel
struct Taps { uint4 t00, t01, t10, t11; };
__device__ __forceinline__ Taps load_taps(const uint4* __restrict__ src, uint4 m) {
    const uint4* p0 = src + (m.x & FT_OFF_MASK);
    const uint4* p1 = src + (m.y & FT_OFF_MASK);
    Taps t;
    t.t00 = __ldg(p0); t.t01 = __ldg(p0 + 1); t.t10 = __ldg(p1); t.t11 = __ldg(p1 + 1);
    return t;
}

// 4-bit "byte is non-zero" mask of a word whose bytes are 0 or a single bit
__device__ __forceinline__ uint32_t nz_nibble(uint32_t a) {
    const uint32_t m = (((a + 0x7F7F7F7Fu) | a) & 0x80808080u) >> 7;   // bit 8i = byte i != 0
    return ((m * 0x00204081u) >> 21) & 15u;                             // bit 8i -> bit 21 + i (no two terms collide)
}

// byte f (0 .. 15) of a 16-frame tap
__device__ __forceinline__ uint32_t tap_byte(const uint4& t, uint32_t sel, bool hi) {
    return (hi ? __byte_perm(t.z, t.w, sel) : __byte_perm(t.x, t.y, sel)) & 0xFFu;
}

// cv2's f32 chain for the flagged frames of one pixel (deferred list entry `slot` of this warp).  The entry carries the
// pixel's geometry words, so nothing but the four taps is fetched again (one round trip for all sixteen frames: the
// shared-memory carve-out leaves too little L1 for the lines to have survived), and fl(b / 255) comes from shared memory.
__device__ __forceinline__ void exact_item(const uint4* __restrict__ src, const uint32_t* __restrict__ list, int slot,
                                           uint8_t* __restrict__ tile_bytes, const float* __restrict__ lut) {
    const int item = (int)list[slot];
    const uint32_t mx = list[FT_LIST_CAP + slot], my = list[2 * FT_LIST_CAP + slot];
    const uint4* p0 = src + (mx & FT_OFF_MASK);
    const uint4* p1 = src + (my & FT_OFF_MASK);
    const uint4 t00 = __ldg(p0), t01 = __ldg(p0 + 1), t10 = __ldg(p1), t11 = __ldg(p1 + 1);
    uint32_t mask = nz_nibble(list[3 * FT_LIST_CAP + slot]) | (nz_nibble(list[4 * FT_LIST_CAP + slot]) << 4) |
                    (nz_nibble(list[5 * FT_LIST_CAP + slot]) << 8) | (nz_nibble(list[6 * FT_LIST_CAP + slot]) << 12);
    const unsigned fx = (mx >> FT_OFF_BITS) & 31u, fy = mx >> (FT_OFF_BITS + 5), fl = (my >> FT_OFF_BITS) & 15u;
    // (1 - fy)(1 - fx) etc. are exact multiples of 2^-10, as cv2 computes them; a tap outside the source has weight 0
    const float w00 = (fl & 1u) ? (float)((32u - fy) * (32u - fx)) * 0.0009765625f : 0.0f;
    const float w01 = (fl & 2u) ? (float)((32u - fy) * fx) * 0.0009765625f : 0.0f;
    const float w10 = (fl & 4u) ? (float)(fy * (32u - fx)) * 0.0009765625f : 0.0f;
    const float w11 = (fl & 8u) ? (float)(fy * fx) * 0.0009765625f : 0.0f;
    uint8_t* out = tile_bytes + 2 * item;
    while (mask) {
        const int f = __ffs(mask) - 1;
        mask &= mask - 1u;
        const uint32_t sel = f & 7;
        const bool hi = (f & 8) != 0;
        float acc = __fmul_rn(lut[tap_byte(t00, sel, hi)], w00);
        acc = __fadd_rn(acc, __fmul_rn(lut[tap_byte(t01, sel, hi)], w01));
        acc = __fadd_rn(acc, __fmul_rn(lut[tap_byte(t10, sel, hi)], w10));
        acc = __fadd_rn(acc, __fmul_rn(lut[tap_byte(t11, sel, hi)], w11));
        // (img * 255).astype(uint8): f32 product, truncation; 2^23 + x rounded toward zero keeps floor(x) in the low byte
        out[(f >> 1) * (2 * FT_ITEMS) + (f & 1)] = (uint8_t)__float_as_uint(__fadd_rz(__fmul_rn(acc, 255.0f), 8388608.0f));
    }
}

// v = 64 V (V the 10-bit fixed-point bilinear sum, v < 2^24) for the four frames of one tap word:
//   c0r0, c1r0 / c0r1, c1r1   left, right column taps of the upper / lower source row (byte f = frame 4q + f)
// One PRMT pairs the two taps of a row for two frames, one dp2a (u16 weights x u8 samples) per frame and row
// accumulates them.  The u8 result floor(V / 1024) is byte 2 of v.  V is a non-zero multiple of 1024 exactly when the
// lowest set bit of v is bit 16 or above: byte 2 of v & -v is then non-zero, and those bytes (one per frame) are what
// the deferred list keeps.
#define FT_QUAD(q, c0r0, c0r1, c1r0, c1r1)                                                                        \
    {                                                                                                             \
        const uint32_t ta = __byte_perm((c0r0), (c1r0), 0x5140), tb4 = __byte_perm((c0r0), (c1r0), 0x7362);       \
        const uint32_t ba = __byte_perm((c0r1), (c1r1), 0x5140), bb = __byte_perm((c0r1), (c1r1), 0x7362);        \
        const uint32_t v0 = __dp2a_lo(wbot, ba, __dp2a_lo(wtop, ta, 0u)), v1 = __dp2a_hi(wbot, ba, __dp2a_hi(wtop, ta, 0u));   \
        const uint32_t v2 = __dp2a_lo(wbot, bb, __dp2a_lo(wtop, tb4, 0u)), v3 = __dp2a_hi(wbot, bb, __dp2a_hi(wtop, tb4, 0u)); \
        if (st) {                                                                                                 \
            tp[(2 * (q)) * FT_ITEMS] = (uint16_t)__byte_perm(v0, v1, 0x3362);                                     \
            tp[(2 * (q) + 1) * FT_ITEMS] = (uint16_t)__byte_perm(v2, v3, 0x3362);                                 \
        }                                                                                                         \
        const uint32_t l0 = v0 & (0u - v0), l1 = v1 & (0u - v1), l2 = v2 & (0u - v2), l3 = v3 & (0u - v3);        \
        acc[q] = __byte_perm(__byte_perm(l0, l1, 0x3362), __byte_perm(l2, l3, 0x3362), 0x5410);                   \
    }

template <bool INTERIOR>
__device__ __forceinline__ void scan_blend(const FusedArgs& a, const uint4* __restrict__ src, uint16_t* __restrict__ tile, uint32_t* __restrict__ list,
                                           const float* __restrict__ lut, int ox1, int oy1, int tid, int lane) {
    uint8_t* tile_bytes = reinterpret_cast<uint8_t*>(tile);
    int cnt = 0;                                             // warp-uniform fill of `list`
    const unsigned lt = (1u << lane) - 1u;
    // Two-deep software pipeline over the CTA's FT_RH x 132 region pixels (one per thread and step): while pixel `it`
    // is blended, the taps of pixel it + 1 and the geometry record of pixel it + 2 are in flight, so neither of
    // the two dependent gathers (L2-resident record -> scan samples) is waited for.
    constexpr int NIT = (FT_ITEMS + 255) / 256;
    // region coordinates of the next geometry fetch; they advance by 256 items = one row + 124 columns.
    int gry = tid / FT_RW, grx = tid - gry * FT_RW;
    int gitem = tid;
    const uint4* gp = a.geo + ((long long)(2 * oy1 - 2 + gry) * a.n + (2 * ox1 - 2 + grx));   // INTERIOR: walks the table directly
    auto next_geom = [&](int& item) -> uint4 {
        uint4 m = make_uint4(0u, 0u, 0u, 0u);                // no valid tap, zero weights, sample 0
        item = -1;
        if (gitem < FT_ITEMS) {
            item = gitem;
            if (INTERIOR) m = __ldg(gp);
            else {
                const int gy = reflect101_safe(2 * oy1 - 2 + gry, a.n), gx = reflect101_safe(2 * ox1 - 2 + grx, a.n);
                m = __ldg(a.geo + (unsigned)(gy * a.n + gx));
            }
        }
        gitem += 256; grx += 256 - FT_RW; gry += 1; gp += a.n + 256 - FT_RW;
        if (grx >= FT_RW) { grx -= FT_RW; gry += 1; gp += a.n - FT_RW; }
        return m;
    };
    // blend one region pixel for the 16 frames of the group, then queue its flagged frames
    auto step = [&](const int item, const uint4 m, const Taps& T) {
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        const bool st = item >= 0;
        uint16_t* tp = tile + (st ? item : 0);
        // beyond the last range bin (image corners, WARP_FILL_OUTLIERS) whole warps have no valid tap: store zeros
        if (__any_sync(0xffffffffu, (m.y >> FT_OFF_BITS) & 15u)) {
            const uint32_t wtop = m.z, wbot = m.w;
            FT_QUAD(0, T.t00.x, T.t10.x, T.t01.x, T.t11.x)
            FT_QUAD(1, T.t00.y, T.t10.y, T.t01.y, T.t11.y)
            FT_QUAD(2, T.t00.z, T.t10.z, T.t01.z, T.t11.z)
            FT_QUAD(3, T.t00.w, T.t10.w, T.t01.w, T.t11.w)
            if (m.y & (1u << (FT_OFF_BITS + 4))) acc[0] = acc[1] = acc[2] = acc[3] = 0x01010101u;   // ident: all frames exact
        } else if (st) {
#pragma unroll
            for (int p = 0; p < FT_PAIRS; ++p) tp[p * FT_ITEMS] = 0;
        }
        // Defer the flagged frames: a lane appends ONE entry (pixel, flag bytes) to the warp's list, and whenever 32
        // are waiting the whole warp redoes them with the f32 chain, one pixel per lane.
        const bool flagged = (acc[0] | acc[1] | acc[2] | acc[3]) != 0u;
        const unsigned pending = __ballot_sync(0xffffffffu, flagged);
        if (pending) {
            if (flagged) {
                const int slot = cnt + __popc(pending & lt);
                list[slot] = (uint32_t)item;
                list[FT_LIST_CAP + slot] = m.x; list[2 * FT_LIST_CAP + slot] = m.y;
#pragma unroll
                for (int q = 0; q < 4; ++q) list[(q + 3) * FT_LIST_CAP + slot] = acc[q];
            }
            cnt += __popc(pending);
            __syncwarp();
            if (cnt >= 32) {
                cnt -= 32;
                exact_item(src, list, cnt + lane, tile_bytes, lut);
                __syncwarp();
            }
        }
    };
    int iA, iB, iC, iD;
    uint4 mA = next_geom(iA);
    Taps TA = load_taps(src, mA);
    uint4 mB = next_geom(iB);
#pragma unroll 1
    for (int it = 0; it + 1 < NIT; it += 2) {
        const Taps TB = load_taps(src, mB);
        const uint4 mC = next_geom(iC);
        step(iA, mA, TA);
        TA = load_taps(src, mC);
        const uint4 mD = next_geom(iD);
        step(iB, mB, TB);
        mA = mC; mB = mD; iA = iC; iB = iD;
    }
    if (NIT & 1) step(iA, mA, TA);
    if (lane < cnt) exact_item(src, list, lane, tile_bytes, lut);
}

__global__ void __launch_bounds__(256, FT_SCAN_MIN_BLOCKS) k_scan16_to_l0l1(const FusedArgs a) {
    extern __shared__ uint4 smem4[];
    uint16_t* tile = reinterpret_cast<uint16_t*>(smem4);                                           // [FT_PAIRS][FT_ITEMS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* list = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(smem4) + FT_TILE_BYTES) + warp * FT_LIST_WORDS * FT_LIST_CAP;
    float* lut = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem4) + FT_TILE_BYTES + 8 * FT_LIST_WORDS * FT_LIST_CAP * 4);
    lut[tid] = g_lut255[tid];   // fl(b / 255) for the exact chain
    __syncthreads();
    // frame group fastest: the CTAs resident at any time share one neighbourhood of the geometry table
    const int ox1 = blockIdx.y * FT_TW1, oy1 = blockIdx.z * FT_TH1;
    const int f0 = blockIdx.x * FT_FR;
    const uint4* src = a.rawi + (size_t)blockIdx.x * a.group_stride;
    asm volatile("" : "+l"(src));   // keep the group base in a register pair: tap addresses are then one IMAD.WIDE each
    // the whole region (tile + pyrDown halo) inside the image: no REFLECT_101 on the record fetch, no bounds on the stores
    const bool interior = 2 * ox1 - 2 >= 0 && 2 * oy1 - 2 >= 0 && 2 * ox1 - 2 + FT_RW <= a.n && 2 * oy1 - 2 + FT_RH <= a.n &&
                          ox1 + FT_TW1 <= a.w1 && oy1 + FT_TH1 <= a.h1;
    if (interior) scan_blend<true>(a, src, tile, list, lut, ox1, oy1, tid, lane);
    else scan_blend<false>(a, src, tile, list, lut, ox1, oy1, tid, lane);
    __syncthreads();

    // one warp per frame pair from here on: level-0 store + level 1 from the shared-memory tile
    const int frame = f0 + 2 * warp;
    if (frame < a.n_frames) {
        const int nb = min(2, a.n_frames - frame);
        if (interior) warp_pyr_pair_impl<true>(tile + warp * FT_ITEMS, a, frame, nb, ox1, oy1, lane);
        else warp_pyr_pair_impl<false>(tile + warp * FT_ITEMS, a, frame, nb, ox1, oy1, lane);
    }
}

// ------------------------------------------------------------------------------------
// pyrDown for the higher levels: each warp loads its FT_RH x 132 source tile from global memory
// (REFLECT_101), then the same shared-memory pass.  4 warps per CTA, one tile each.
// ------------------------------------------------------------------------------------
// tile height of the upper-level pyrDown kernel (destination rows per warp tile).  Measured (256 frames, levels 2 + 3):
// 16 rows 0.417 ms, 8 rows 0.383 ms, 4 rows 0.406 ms (0.369 ms with 12 CTAs/SM and spills): 8 is the default
#ifndef FT_PYR_TH1
#define FT_PYR_TH1 8
#endif
#define FT_PYR_RH (2 * FT_PYR_TH1 + 3)
#define FT_PYR_TILE_WORDS (FT_PYR_RH * FT_RWW)
#ifndef FT_PYR_MIN_BLOCKS
#define FT_PYR_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, FT_PYR_MIN_BLOCKS)
k_pyr_down_w(const uint8_t* __restrict__ src, size_t src_stride, int sp, int sw, int sh, uint8_t* __restrict__ dst,
             size_t dst_stride, int dp, int dw, int dh, int tiles_x, int tiles_per_frame, int n_tiles) {
    __shared__ uint32_t s_tile[4][FT_PYR_TILE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * 4 + warp;
    if (t >= n_tiles) return;
    const int frame = t / tiles_per_frame, tt = t - frame * tiles_per_frame;
    const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
    const int ox1 = tx * FT_TW1, oy1 = ty * FT_PYR_TH1;
    const uint8_t* __restrict__ s = src + (size_t)frame * src_stride;
    uint32_t* tw = s_tile[warp];
    const int x0 = 2 * ox1 - 2;
    // Region word k (k = lane, plus word 32 on lane 0) holds source bytes x0 + 4k .. x0 + 4k + 3 of the row.
    // Words that lie fully inside the row are cut out of two ALIGNED 32-bit loads with a funnel shift (rows are
    // only byte-aligned: widths are odd at the higher levels); the few words that straddle an image edge fall
    // back to REFLECT_101 byte loads.  Aligned loads may run up to 7 bytes past the row end: into the next row,
    // or into the padding every level allocation carries (rf_frameset_alloc).
    const int xk = x0 + 4 * lane, xk32 = x0 + 128;
    if (x0 >= 0 && x0 + FT_RW + 4 <= sw && 2 * oy1 - 2 >= 0 && 2 * oy1 - 2 + FT_PYR_RH <= sh) {
        // interior tile: no reflection, no edge words; the row pointer advances by the pitch.  Rows are only
        // byte-aligned (odd widths at the higher levels), so each region word is cut out of two aligned loads.
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(s) & 3u);
        const uint32_t* __restrict__ s4 = reinterpret_cast<const uint32_t*>(s - mis) + lane;   // aligned frame base
        unsigned off = mis + (unsigned)(2 * oy1 - 2) * (unsigned)sp + (unsigned)x0;             // byte offset of the row's region
#pragma unroll 5
        for (int r = 0; r < FT_PYR_RH; ++r, off += sp) {
            const uint32_t* ap = s4 + (off >> 2);
            const unsigned sh8 = 8u * (off & 3u);
            tw[r * FT_RWW + lane] = __funnelshift_r(__ldg(ap), __ldg(ap + 1), sh8);
            if (lane == 0) tw[r * FT_RWW + 32] = __funnelshift_r(__ldg(ap + 32), __ldg(ap + 33), sh8);
        }
    } else {
    const bool in_k = xk >= 0 && xk + 3 < sw, in_32 = xk32 + 3 < sw;
#pragma unroll 5
    for (int r = 0; r < FT_PYR_RH; ++r) {
        const uint8_t* row = s + (size_t)reflect101_safe(2 * oy1 - 2 + r, sh) * sp;
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
    }
    __syncwarp();
    warp_pyr_tile<FT_PYR_TH1>(s_tile[warp], FT_RWW, dst + (size_t)frame * dst_stride, dp, dw, dh, ox1, oy1, lane);
}

// ------------------------------------------------------------------------------------
// pyrDown for the higher levels, source tiles staged by TMA.
// Every pyramid level has a 16-byte row pitch, so a level is one 3-D tensor map (x, y, frame) and a warp's
// FT_TMA_RH x 132-byte source region arrives as ONE cp.async.bulk.tensor box: no address arithmetic, no re-alignment
// shifts, no per-row loads.  The innermost box coordinate must be a multiple of 16 bytes (measured with
// tools/probe/tma_probe.cu: an unaligned x raises an illegal-instruction fault; y and the frame index are free), so the
// box starts 16 bytes left of the tile, at (2*ox1 - 16, 2*oy1 - 2), and is 160 bytes wide; the 1-4-6-4-1 taps are taken
// from three aligned words per lane instead of two.  Bytes outside the image arrive as zeros; the (at most two) columns /
// rows REFLECT_101 needs on a border tile are then copied inside the staged tile.  Each warp walks its tiles with two
// buffers: the box of tile i + 1 is in flight while tile i is filtered.
// ------------------------------------------------------------------------------------
#ifndef FT_TMA_TH1
#define FT_TMA_TH1 16                                         // destination rows per warp tile (measured in-step, 256 frames: 8 rows 0.192 ms, 16 rows 0.136 ms)
#endif
#define FT_TMA_RH (2 * FT_TMA_TH1 + 3)
#define FT_TMA_ROW 160                                        // box width in bytes: source columns 128 tx - 16 .. 128 tx + 143
#define FT_TMA_X0 14                                          // byte of a staged row that holds region column 0 (source 2*ox1 - 2)
#define FT_TMA_WW (FT_TMA_ROW / 4)
#define FT_TMA_BOX_BYTES (FT_TMA_RH * FT_TMA_ROW)             // 2736
#define FT_TMA_BUF_BYTES ((FT_TMA_BOX_BYTES + 127) & ~127)    // buffers start on 128-byte boundaries
#define FT_TMA_WARPS 4
#ifndef FT_PYR_TMA_CTAS_PER_SM
#define FT_PYR_TMA_CTAS_PER_SM 8
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// pyrDown of one TMA-staged tile: rows of FT_TMA_WW words, word 3 + 4c/4 ... holds source columns 2*ox1 - 4 + ...;
// lane k owns destination columns ox1 + 2k, 2k + 1 and reads the aligned words A | B | C = source columns
// 2*ox1 + 4k - 4 .. + 4k + 7.
template <bool INTERIOR>
__device__ __forceinline__ void warp_pyr_staged_impl(const uint32_t* __restrict__ tile, uint8_t* __restrict__ dst, int dp, int dw, int dh,
                                                     int ox1, int oy1, int lane) {
    const int x = ox1 + 2 * lane;
    const bool x_ok = INTERIOR || x < dw;
    const int rows1 = INTERIOR ? FT_TMA_TH1 : min(FT_TMA_TH1, dh - oy1);
    uint8_t* q1 = dst + (size_t)oy1 * dp + x;
    const uint32_t* tp = tile + 3 + lane;
    uint32_t h[5];
#pragma unroll
    for (int r = 0; r < FT_TMA_RH; ++r) {
        const uint32_t A = tp[r * FT_TMA_WW], B = tp[r * FT_TMA_WW + 1], C = tp[r * FT_TMA_WW + 2];
        const uint32_t he = __dp4a(B, 0x00010406u, __dp4a(A, 0x04010000u, 0u));   // columns 4k - 2 .. 4k + 2
        const uint32_t ho = __dp4a(C, 0x00000001u, __dp4a(B, 0x04060401u, 0u));   // columns 4k .. 4k + 4
        h[r % 5] = he | (ho << 16);
        if (r >= 4 && (r & 1) == 0) {
            const uint32_t px2 = pyr_vsum(h[(r - 4) % 5], h[(r - 3) % 5], h[(r - 2) % 5], h[(r - 1) % 5], h[r % 5]);
            if (INTERIOR || ((r - 4) / 2 < rows1 && x_ok)) *reinterpret_cast<uint16_t*>(q1) = (uint16_t)px2;
            q1 += dp;
        }
    }
}

__global__ void __launch_bounds__(32 * FT_TMA_WARPS)
k_pyr_down_tma(const __grid_constant__ CUtensorMap src_map, int sw, int sh, uint8_t* __restrict__ dst, size_t dst_stride, int dp, int dw,
               int dh, int tiles_x, int tiles_y, int n_tiles) {
    __shared__ __align__(128) uint8_t s_buf[FT_TMA_WARPS][2][FT_TMA_BUF_BYTES];
    __shared__ __align__(8) uint64_t s_bar[FT_TMA_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Tiles are dealt round-robin (tile t, t + stride, ...: at any instant the machine works on one compact band of the
    // level, which keeps DRAM pages and shared halos hot; contiguous runs per warp measured 30 % slower); the
    // (frame, ty, tx) coordinates advance by the decomposed stride, without divisions.
    const int stride = gridDim.x * FT_TMA_WARPS;
    int t = blockIdx.x * FT_TMA_WARPS + warp;
    const int t_end = n_tiles;
    if (t >= t_end) return;
    const int tpf = tiles_x * tiles_y;
    int frame = t / tpf, tt = t - frame * tpf;
    int ty = tt / tiles_x, tx = tt - ty * tiles_x;
    const int dframe = stride / tpf, dtt = stride - dframe * tpf, dty = dtt / tiles_x, dtx = dtt - dty * tiles_x;
    auto advance = [&](int& f, int& y, int& x) {
        x += dtx; if (x >= tiles_x) { x -= tiles_x; ++y; }
        y += dty; if (y >= tiles_y) { y -= tiles_y; ++f; }
        f += dframe;
    };
    int nframe = frame, nty = ty, ntx = tx;   // coordinates of the next tile to request
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]), bar1 = smem_u32(&s_bar[warp][1]);
    const uint32_t buf0 = smem_u32(&s_buf[warp][0][0]), buf1 = smem_u32(&s_buf[warp][1][0]);
    auto issue = [&](int b) {   // lane 0 only: request tile (nframe, nty, ntx) into buffer b, then step to the one after
        const uint32_t bar = b ? bar1 : bar0;
        mbar_expect_tx(bar, FT_TMA_BOX_BYTES);
        tma_load_3d(b ? buf1 : buf0, &src_map, bar, 2 * ntx * FT_TW1 - 16, 2 * nty * FT_TMA_TH1 - 2, nframe);
        advance(nframe, nty, ntx);
    };
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(0);
    }
    __syncwarp();
    for (int i = 0; t < t_end; t += stride, ++i) {
        const int b = i & 1;
        if (lane == 0 && t + stride < t_end) {
            // buffer b ^ 1 was read (and maybe patched) by this warp in the previous round: order those generic-proxy
            // accesses before the async-proxy write that refills it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(b ^ 1);
        }
        const uint32_t bar = b ? bar1 : bar0, parity = (uint32_t)(i >> 1) & 1u;
        int spins = 0;
        while (!mbar_try_wait(bar, parity))
            if (++spins > (1 << 22)) __trap();   // a lost transaction must not hang the device
        const int ox1 = tx * FT_TW1, oy1 = ty * FT_TMA_TH1;
        const int x0 = 2 * ox1 - 2, y0 = 2 * oy1 - 2;
        uint32_t* tw = reinterpret_cast<uint32_t*>(&s_buf[warp][b][0]);
        uint8_t* tb = &s_buf[warp][b][0];
        const int jr = sw - x0, rb = sh - y0;   // first region column / row outside the image
        if (x0 < 0 || y0 < 0 || jr <= FT_RW - 1 || rb <= FT_TMA_RH - 1) {
            // REFLECT_101 inside the staged tile: columns first (on the rows that exist), then whole rows
            for (int r = lane; r < FT_TMA_RH; r += 32) {
                uint8_t* row = tb + r * FT_TMA_ROW + FT_TMA_X0;   // region column 0
                if (x0 < 0) { row[0] = row[4]; row[1] = row[3]; }
                if (jr <= FT_RW - 1) {
                    row[jr] = row[jr - 2];
                    if (jr + 1 <= FT_RW - 1 && jr >= 3) row[jr + 1] = row[jr - 3];
                }
            }
            __syncwarp();
            for (int wd = lane; wd < FT_TMA_WW; wd += 32) {
                if (y0 < 0) { tw[wd] = tw[4 * FT_TMA_WW + wd]; tw[FT_TMA_WW + wd] = tw[3 * FT_TMA_WW + wd]; }
                if (rb <= FT_TMA_RH - 1 && rb >= 2) {
                    tw[rb * FT_TMA_WW + wd] = tw[(rb - 2) * FT_TMA_WW + wd];
                    if (rb + 1 <= FT_TMA_RH - 1 && rb >= 3) tw[(rb + 1) * FT_TMA_WW + wd] = tw[(rb - 3) * FT_TMA_WW + wd];
                }
            }
            __syncwarp();
        }
        if (ox1 + FT_TW1 <= dw && oy1 + FT_TMA_TH1 <= dh) warp_pyr_staged_impl<true>(tw, dst + (size_t)frame * dst_stride, dp, dw, dh, ox1, oy1, lane);
        else warp_pyr_staged_impl<false>(tw, dst + (size_t)frame * dst_stride, dp, dw, dh, ox1, oy1, lane);
        __syncwarp();
        advance(frame, ty, tx);
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
int rf_fused_wp(const rf_handle* h) { return ((h->cfg.range_bins + 1) + 3) & ~3; }   // >= W + 1, multiple of 4

int rf_launch_build_map2(rf_handle* h) {
    const size_t count = (size_t)h->n * h->n;
    if ((size_t)h->cfg.azimuths * rf_fused_wp(h) > (size_t)FT_OFF_MASK) return RF_OK;   // no fused path for this geometry (rf_launch_scan_to_l0l1 refuses)
    k_build_map2<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->map, count, h->cfg.azimuths, h->cfg.range_bins,
                                                                          rf_fused_wp(h), h->map2);
    RF_CHECK_LAUNCH(h);
    k_build_lut<<<1, 256, 0, h->stream>>>();
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// d_raw: frames of [A][pitch] power bins (no metadata), pitch a multiple of 16 and >= Wp
int rf_launch_interleave(rf_handle* h, const uint8_t* d_raw, size_t frame_stride, int pitch, int n_frames, uint32_t* d_out,
                         const int32_t* d_frame_sel) {
    const int Wp = rf_fused_wp(h);
    const int groups = (n_frames + FT_FR - 1) / FT_FR;   // zero-filled up to a multiple of 16 frames
    dim3 grd((Wp / 4 + 255) / 256, h->cfg.azimuths, groups);
    k_interleave16<<<grd, 256, 0, h->stream>>>(d_raw, frame_stride, pitch, h->cfg.azimuths, h->cfg.range_bins, n_frames,
                                               reinterpret_cast<uint4*>(d_out), Wp, d_frame_sel);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

size_t rf_interleave_words(const rf_handle* h, int max_frames) {
    return (size_t)((max_frames + FT_FR - 1) / FT_FR) * h->cfg.azimuths * rf_fused_wp(h) * 4;
}

int rf_launch_scan_to_l0l1(rf_handle* h, const uint32_t* d_rawi, const FrameSet& fs, int n_frames) {
    if (h->n % 4) return rf_fail(h, RF_E_BADARG, "cartesian size %d is not a multiple of 4", h->n);
    if (fs.n_levels < 2) return rf_fail(h, RF_E_BADARG, "fused image path needs at least two pyramid levels");
    if ((size_t)h->cfg.azimuths * rf_fused_wp(h) > (size_t)FT_OFF_MASK)
        return rf_fail(h, RF_E_BADARG, "fused image path: %d azimuths x %d bins exceed the %d-bit sample offsets of the geometry records",
                       h->cfg.azimuths, rf_fused_wp(h), FT_OFF_BITS);
    static bool attr_set[64] = {};   // per device: the attribute belongs to the device's context
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_scan16_to_l0l1, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
        attr_set[h->device & 63] = true;
    }
    FusedArgs a;
    a.rawi = reinterpret_cast<const uint4*>(d_rawi); a.Wp = rf_fused_wp(h); a.A = h->cfg.azimuths; a.group_stride = (size_t)a.A * a.Wp;
    a.geo = h->map2; a.n = h->n;
    a.l0 = fs.lvl[0]; a.l0_stride = fs.lvl_stride[0]; a.p0 = fs.pitch[0];
    a.l1 = fs.lvl[1]; a.l1_stride = fs.lvl_stride[1]; a.p1 = fs.pitch[1]; a.w1 = fs.w[1]; a.h1 = fs.h[1];
    a.n_frames = n_frames;
    dim3 grd((n_frames + FT_FR - 1) / FT_FR, (fs.w[1] + FT_TW1 - 1) / FT_TW1, (fs.h[1] + FT_TH1 - 1) / FT_TH1);
    k_scan16_to_l0l1<<<grd, 256, FT_SMEM_BYTES, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// tensor map of one pyramid level: (x, y, frame) u8, box FT_TMA_ROW x FT_PYR_RH x 1, zero fill outside the image.
// Encoded once per distinct (base, geometry) and kept for the life of the process: a map is a pure function of its key.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static const CUtensorMap* level_tmap(rf_handle* h, const FrameSet& fs, int l) {
    typedef std::tuple<const void*, int, int, int, size_t, int> Key;
    static std::map<Key, CUtensorMap> cache;
    static std::mutex mu;
    static PFN_encodeTiled encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            rf_fail(h, RF_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
            return nullptr;
        }
        encode = (PFN_encodeTiled)fn;
    }
    const Key key(fs.lvl[l], fs.w[l], fs.h[l], fs.pitch[l], fs.lvl_stride[l], fs.count);
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)fs.w[l], (cuuint64_t)fs.h[l], (cuuint64_t)fs.count};
    const cuuint64_t strides[2] = {(cuuint64_t)fs.pitch[l], (cuuint64_t)fs.lvl_stride[l]};
    const cuuint32_t box[3] = {FT_TMA_ROW, FT_TMA_RH, 1}, estr[3] = {1, 1, 1};
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, fs.lvl[l], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rf_fail(h, RF_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for level %d: %d x %d, pitch %d", (int)r, l, fs.w[l], fs.h[l], fs.pitch[l]);
        return nullptr;
    }
    return &cache.emplace(key, tm).first->second;
}

// levels first_level .. n_levels-1 from their predecessors
int rf_launch_pyr_levels(rf_handle* h, const FrameSet& fs, int first_level, int n_frames) {
    static const bool no_tma = getenv("RADARFE_NO_TMA") != nullptr;   // diagnostic: the plain-load kernel
    for (int l = first_level; l < fs.n_levels; ++l) {
        const bool tma = !no_tma && fs.w[l - 1] >= 8 && fs.h[l - 1] >= 8;   // the in-tile border fix-up needs the reflected columns / rows inside the box
        const int th1 = tma ? FT_TMA_TH1 : FT_PYR_TH1;
        const int tiles_x = (fs.w[l] + FT_TW1 - 1) / FT_TW1, tiles_y = (fs.h[l] + th1 - 1) / th1;
        const int n_tiles = tiles_x * tiles_y * n_frames;
        if (n_tiles == 0) continue;
        if (tma) {
            const CUtensorMap* tm = level_tmap(h, fs, l - 1);
            if (!tm) return RF_E_CUDA;
            int ctas = (n_tiles + FT_TMA_WARPS - 1) / FT_TMA_WARPS;
            // persistent: as many CTAs as are resident at once, every warp walks its tiles with a prefetch in flight
            static int per_sm = 0;
            if (!per_sm) {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pyr_down_tma, 32 * FT_TMA_WARPS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
                if (per_sm > FT_PYR_TMA_CTAS_PER_SM) per_sm = FT_PYR_TMA_CTAS_PER_SM;
            }
            const int cap = h->sm_count * per_sm;
            if (ctas > cap) ctas = cap;
            k_pyr_down_tma<<<ctas, 32 * FT_TMA_WARPS, 0, h->stream>>>(*tm, fs.w[l - 1], fs.h[l - 1], fs.lvl[l], fs.lvl_stride[l], fs.pitch[l],
                                                                      fs.w[l], fs.h[l], tiles_x, tiles_y, n_tiles);
        } else {
            k_pyr_down_w<<<(n_tiles + 3) / 4, 128, 0, h->stream>>>(fs.lvl[l - 1], fs.lvl_stride[l - 1], fs.pitch[l - 1], fs.w[l - 1], fs.h[l - 1],
                                                                   fs.lvl[l], fs.lvl_stride[l], fs.pitch[l], fs.w[l], fs.h[l], tiles_x,
                                                                   tiles_x * tiles_y, n_tiles);
        }
        RF_CHECK_LAUNCH(h);
    }
    return RF_OK;
}
