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
#define FT_RWW (FT_RW / 4)       // 33 words per region row
#define FT_FR 16                 // frames per CTA (one interleave group)
#define FT_TILE_WORDS (FT_RH * FT_RWW)
#define FT_TILE_BYTES (FT_TILE_WORDS * 4)
#define FT_SMEM_BYTES ((FT_FR * FT_TILE_WORDS + 8 * 64) * 4)   // tiles + 8 per-warp deferred lists

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
// warp-level pyrDown of one tile held in shared memory (optionally also the level-0 store of it).
//   tile : [FT_RH][FT_RWW] words = 19 x 132 source bytes; byte (i, j) = source (2*oy1 - 2 + i, 2*ox1 - 2 + j)
// Lane k owns destination columns ox1 + 2k, 2k + 1: their horizontal 1-4-6-4-1 sums (two u16 per word) of
// the last five source rows roll through registers, every second row emits one destination row.
// Writes the FT_TH1 x 64 destination tile at (oy1, ox1); with L0, also source rows 2..2*FT_TH1+1 / bytes 2..129 of the
// tile to the level-0 image (n x n) as 32 aligned words per row.
// ------------------------------------------------------------------------------------
// INTERIOR: the whole tile lies inside both images and the destination pitch is even — no per-row or per-lane
// bounds checks; addresses advance by pointer increments either way.
template <bool L0, bool INTERIOR, int TH1>
__device__ __forceinline__ void warp_pyr_tile_impl(const uint32_t* __restrict__ tile, uint8_t* __restrict__ dst, int dw, int dh,
                                                   int ox1, int oy1, int lane, uint8_t* __restrict__ l0, int n) {
    const int x = ox1 + 2 * lane;
    const bool even_pitch = INTERIOR || (dw & 1) == 0;
    const bool x_ok = INTERIOR || x < dw, x0_ok = INTERIOR || 2 * ox1 + 4 * lane < n;
    const int rows0 = INTERIOR ? 2 * TH1 : min(2 * TH1, n - 2 * oy1);   // level-0 rows of this tile inside the image
    const int rows1 = INTERIOR ? TH1 : min(TH1, dh - oy1);
    uint8_t* q0 = L0 ? l0 + (size_t)(2 * oy1) * n + 2 * ox1 + 4 * lane : nullptr;
    uint8_t* q1 = dst + (size_t)oy1 * dw + x;
    const uint32_t* tp = tile + lane;
    uint32_t h[5];
#pragma unroll
    for (int r = 0; r < 2 * TH1 + 3; ++r) {
        const uint32_t w0 = tp[r * FT_RWW], w1 = tp[r * FT_RWW + 1];
        if (L0 && r >= 2 && r < 2 + 2 * TH1) {
            if (INTERIOR || (r - 2 < rows0 && x0_ok)) *reinterpret_cast<uint32_t*>(q0) = __funnelshift_r(w0, w1, 16);
            q0 += n;
        }
        const uint32_t he = __dp4a(w1, 0x00000001u, __dp4a(w0, 0x04060401u, 0u));   // columns 4k .. 4k+4
        const uint32_t ho = __dp4a(w1, 0x00010406u, __dp4a(w0, 0x04010000u, 0u));   // columns 4k+2 .. 4k+6
        h[r % 5] = he | (ho << 16);
        if (r >= 4 && (r & 1) == 0) {
            // both halves stay below 2^16 (255 * 256 + 128), so the packed sum never carries across
            const uint32_t s = h[(r - 4) % 5] + 4u * h[(r - 3) % 5] + 6u * h[(r - 2) % 5] + 4u * h[(r - 1) % 5] + h[r % 5] + 0x00800080u;
            if (INTERIOR || ((r - 4) / 2 < rows1 && x_ok)) {
                const uint32_t px2 = __byte_perm(s, 0u, 0x4431);   // (s >> 8) & 0xFF | (s >> 24) << 8
                if (even_pitch) *reinterpret_cast<uint16_t*>(q1) = (uint16_t)px2;
                else { q1[0] = (uint8_t)px2; if (x + 1 < dw) q1[1] = (uint8_t)(px2 >> 8); }
            }
            q1 += dw;
        }
    }
}

template <bool L0, int TH1>
__device__ __forceinline__ void warp_pyr_tile(const uint32_t* __restrict__ tile, uint8_t* __restrict__ dst, int dw, int dh,
                                              int ox1, int oy1, int lane, uint8_t* __restrict__ l0, int n) {
    const bool interior = (dw & 1) == 0 && ox1 + FT_TW1 <= dw && oy1 + TH1 <= dh &&
                          (!L0 || (2 * ox1 + 2 * FT_TW1 <= n && 2 * oy1 + 2 * TH1 <= n));
    if (interior) warp_pyr_tile_impl<L0, true, TH1>(tile, dst, dw, dh, ox1, oy1, lane, l0, n);
    else warp_pyr_tile_impl<L0, false, TH1>(tile, dst, dw, dh, ox1, oy1, lane, l0, n);
}

// ------------------------------------------------------------------------------------
// raw (interleaved) -> level 0 + level 1
// ------------------------------------------------------------------------------------
struct FusedArgs {
    const uint4* rawi; size_t group_stride;   // samples per interleave group plane (A * Wp)
    int Wp, A;
    const uint2* map2; int n;
    uint8_t* l0; size_t l0_stride;
    uint8_t* l1; size_t l1_stride; int w1, h1;
    int n_frames;
};

__device__ float g_lut255[256];   // fl(b / 255), parseData.py:43 (filled by k_build_lut at handle creation)

__device__ uint8_t g_lut_id[256];  // trunc(fl(fl(b / 255) * 255)): parseData.py:43 followed by getTransformKLT.py:356-357

__global__ void k_build_lut() {
    const float s = __fdiv_rn((float)threadIdx.x, 255.0f);
    g_lut255[threadIdx.x] = s;
    g_lut_id[threadIdx.x] = (uint8_t)__float_as_uint(__fadd_rz(__fmul_rn(s, 255.0f), 8388608.0f));
}

// V (the 10-bit fixed-point bilinear sum, <= 255 * 1024) for the four frames of one tap word.
//   c0r0, c1r0 / c0r1, c1r1   left, right column taps of the upper / lower source row (byte f = frame 4q + f)
//   wtop = w00 | w01 << 16, wbot = w10 | w11 << 16    the four integer weights m = (32 - fy | fy)(32 - fx | fx)
// One PRMT pairs the two taps of a row for two frames, one dp2a (u16 weights x u8 samples) per frame and row
// accumulates them.  The u8 result is V >> 10; V is a multiple of 1024 exactly when its low 10 bits are 0.
// Such frames (V != 0) are flagged in `need` and redone with cv2's f32 chain.
#define FT_QUAD(q, c0r0, c0r1, c1r0, c1r1)                                                                        \
    {                                                                                                             \
        const uint32_t ta = __byte_perm((c0r0), (c1r0), 0x5140), tb4 = __byte_perm((c0r0), (c1r0), 0x7362);       \
        const uint32_t ba = __byte_perm((c0r1), (c1r1), 0x5140), bb = __byte_perm((c0r1), (c1r1), 0x7362);        \
        const uint32_t v4[4] = {__dp2a_lo(wbot, ba, __dp2a_lo(wtop, ta, 0u)), __dp2a_hi(wbot, ba, __dp2a_hi(wtop, ta, 0u)), \
                                __dp2a_lo(wbot, bb, __dp2a_lo(wtop, tb4, 0u)), __dp2a_hi(wbot, bb, __dp2a_hi(wtop, tb4, 0u))}; \
        _Pragma("unroll") for (int f = 0; f < 4; ++f) {                                                           \
            if ((v4[f] & 0x3FFu) == 0u && v4[f] != 0u) need |= 1u << (4 * (q) + f);                               \
            tb[(4 * (q) + f) * FT_TILE_BYTES] = (uint8_t)(v4[f] >> 10);                                           \
        }                                                                                                         \
    }

struct PixelGeom {
    const uint4 *p0, *p1;      // sample (iy, ix) and (iy + 1, ix) of the interleave group
    unsigned fx, fy, fl;       // 5-bit fractions, tap validity (bit 0: 00, 1: 01, 2: 10, 3: 11)
};

// geometry record of region item `item` (REFLECT_101 at the image border: the pyrDown halo)
__device__ __forceinline__ uint2 load_geom(const FusedArgs& a, int item, int ox1, int oy1) {
    const int ry = item / FT_RW, rx = item - ry * FT_RW;
    const int gy = reflect101_safe(2 * oy1 - 2 + ry, a.n), gx = reflect101_safe(2 * ox1 - 2 + rx, a.n);
    return __ldg(a.map2 + (unsigned)(gy * a.n + gx));
}

__device__ __forceinline__ PixelGeom decode_geom(const FusedArgs& a, const uint4* __restrict__ src, uint2 m) {
    PixelGeom g;
    g.fx = m.y & 31u; g.fy = (m.y >> 5) & 31u; g.fl = (m.y >> 10) & 15u;
    g.p0 = src + m.x;
    g.p1 = g.p0 + a.Wp - ((m.y & 0x4000u) ? (unsigned)a.A * (unsigned)a.Wp : 0u);
    return g;
}

__device__ __forceinline__ PixelGeom pixel_geom(const FusedArgs& a, const uint4* __restrict__ src, int item, int ox1, int oy1) {
    return decode_geom(a, src, load_geom(a, item, ox1, oy1));
}

// the four taps (16 frames each) of one pixel; nothing is loaded for a pixel with no valid tap
struct Taps { uint4 t00, t01, t10, t11; };
__device__ __forceinline__ Taps load_taps(const FusedArgs& a, const uint4* __restrict__ src, uint2 m) {
    Taps t;
    t.t00 = t.t01 = t.t10 = t.t11 = make_uint4(0u, 0u, 0u, 0u);
    if ((m.y >> 10) & 15u) {
        const PixelGeom g = decode_geom(a, src, m);
        t.t00 = __ldg(g.p0); t.t01 = __ldg(g.p0 + 1); t.t10 = __ldg(g.p1); t.t11 = __ldg(g.p1 + 1);
    }
    return t;
}

// cv2's f32 chain for the flagged frames of one pixel: entry = item | frame mask << 13
__device__ __forceinline__ void exact_item(const FusedArgs& a, const uint4* __restrict__ src, uint32_t entry, int ox1, int oy1,
                                           uint8_t* __restrict__ tile_bytes) {
    const int item = entry & 0x1FFF;
    uint32_t mask = entry >> 13;
    const PixelGeom g = pixel_geom(a, src, item, ox1, oy1);
    // (1 - fy)(1 - fx) etc. are exact multiples of 2^-10, as cv2 computes them; a tap outside the source has weight 0
    const float w00 = (g.fl & 1u) ? (float)((32u - g.fy) * (32u - g.fx)) * 0.0009765625f : 0.0f;
    const float w01 = (g.fl & 2u) ? (float)((32u - g.fy) * g.fx) * 0.0009765625f : 0.0f;
    const float w10 = (g.fl & 4u) ? (float)(g.fy * (32u - g.fx)) * 0.0009765625f : 0.0f;
    const float w11 = (g.fl & 8u) ? (float)(g.fy * g.fx) * 0.0009765625f : 0.0f;
    const uint8_t* b0 = reinterpret_cast<const uint8_t*>(g.p0);
    const uint8_t* b1 = reinterpret_cast<const uint8_t*>(g.p1);
    uint8_t* out = tile_bytes + item;
    while (mask) {
        const int f = __ffs(mask) - 1;
        mask &= mask - 1u;
        float acc = __fmul_rn(g_lut255[__ldg(b0 + f)], w00);
        acc = __fadd_rn(acc, __fmul_rn(g_lut255[__ldg(b0 + f + 16)], w01));
        acc = __fadd_rn(acc, __fmul_rn(g_lut255[__ldg(b1 + f)], w10));
        acc = __fadd_rn(acc, __fmul_rn(g_lut255[__ldg(b1 + f + 16)], w11));
        // (img * 255).astype(uint8): f32 product, truncation; 2^23 + x rounded toward zero keeps floor(x) in the low byte
        out[f * FT_TILE_BYTES] = (uint8_t)__float_as_uint(__fadd_rz(__fmul_rn(acc, 255.0f), 8388608.0f));
    }
}

#define FT_LIST_CAP 64   // deferred (pixel, frame) entries per warp

__global__ void __launch_bounds__(256, FT_SCAN_MIN_BLOCKS) k_scan16_to_l0l1(const FusedArgs a) {
    extern __shared__ uint32_t smem[];
    uint32_t* tiles = smem;                                  // [FT_FR][FT_RH][FT_RWW]
    uint32_t* lists = tiles + FT_FR * FT_TILE_WORDS;         // [8 warps][FT_LIST_CAP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // frame group fastest: the CTAs resident at any time share one neighbourhood of the geometry table
    const int ox1 = blockIdx.y * FT_TW1, oy1 = blockIdx.z * FT_TH1;
    const int f0 = blockIdx.x * FT_FR;
    const uint4* __restrict__ src = a.rawi + (size_t)blockIdx.x * a.group_stride;
    uint8_t* tile_bytes = reinterpret_cast<uint8_t*>(tiles);
    uint32_t* list = lists + warp * FT_LIST_CAP;
    int cnt = 0;                                             // warp-uniform fill of `list`
    const unsigned lt = (1u << lane) - 1u;

    // Two-deep software pipeline over the CTA's FT_RH x 132 region pixels (one per thread and step): while pixel `it`
    // is blended, the taps of pixel it + 1 and the geometry record of pixel it + 2 are in flight, so neither of
    // the two dependent gathers (L2-resident record -> scan samples) is waited for.
    constexpr int NIT = (FT_RH * FT_RW + 255) / 256;
    // region coordinates of the next geometry fetch; they advance by 256 items = one row + 124 columns.
    // (A 4 x 8 pixel patch per warp instead of a 32 x 1 strip was measured: 4 % fewer L1 wavefronts, 9 % more
    // instructions from the ragged 33 x 5 patch grid, 5 % slower.)
    int gry = tid / FT_RW, grx = tid - gry * FT_RW;
    int gitem = tid;
    auto next_geom = [&](int& item) -> uint2 {
        uint2 m = make_uint2(0u, 0u);                        // fl == 0: no taps
        item = -1;
        if (gitem < FT_RH * FT_RW) {
            item = gitem;
            const int gy = reflect101_safe(2 * oy1 - 2 + gry, a.n), gx = reflect101_safe(2 * ox1 - 2 + grx, a.n);
            m = __ldg(a.map2 + (unsigned)(gy * a.n + gx));
        }
        gitem += 256; grx += 256 - FT_RW; gry += 1;
        if (grx >= FT_RW) { grx -= FT_RW; gry += 1; }
        return m;
    };
    // blend one region pixel for the 16 frames of the group, then queue its flagged frames
    auto step = [&](const int item, const uint2 m, const Taps& T) {
        uint32_t need = 0u;
        if (item >= 0) {
            const unsigned fx = m.y & 31u, fy = (m.y >> 5) & 31u, fl = (m.y >> 10) & 15u;
            uint8_t* tb = tile_bytes + item;
            if (fl == 0u) {   // beyond the last range bin (image corners): WARP_FILL_OUTLIERS
#pragma unroll
                for (int f = 0; f < FT_FR; ++f) tb[f * FT_TILE_BYTES] = 0;
            } else if ((m.y & 0x7FFu) == 0x400u) {
                // fx = fy = 0 and tap 00 valid: the weights are 1, 0, 0, 0 and cv2's chain collapses to
                // trunc(fl(fl(b / 255) * 255)) of that one sample (every such V is a multiple of 1024)
                const uint32_t q[4] = {T.t00.x, T.t00.y, T.t00.z, T.t00.w};
#pragma unroll
                for (int f = 0; f < FT_FR; ++f) tb[f * FT_TILE_BYTES] = g_lut_id[(q[f >> 2] >> (8 * (f & 3))) & 0xFFu];
            } else {
                uint4 T00 = T.t00, T01 = T.t01, T10 = T.t10, T11 = T.t11;
                if (fl != 15u) {   // a tap outside the source contributes 0
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    if (!(fl & 1u)) T00 = z; if (!(fl & 2u)) T01 = z; if (!(fl & 4u)) T10 = z; if (!(fl & 8u)) T11 = z;
                }
                const uint32_t wy0 = 32u - fy, wx0 = 32u - fx;
                const uint32_t wtop = wy0 * wx0 | (wy0 * fx) << 16, wbot = fy * wx0 | (fy * fx) << 16;
                FT_QUAD(0, T00.x, T10.x, T01.x, T11.x)
                FT_QUAD(1, T00.y, T10.y, T01.y, T11.y)
                FT_QUAD(2, T00.z, T10.z, T01.z, T11.z)
                FT_QUAD(3, T00.w, T10.w, T01.w, T11.w)
            }
        }
        // Defer the flagged frames: a lane appends ONE entry (pixel, frame mask) to the warp's list, and whenever 32
        // are waiting the whole warp redoes them with the f32 chain, one pixel per lane.
        const unsigned pending = __ballot_sync(0xffffffffu, need != 0u);
        if (pending) {
            if (need) list[cnt + __popc(pending & lt)] = (uint32_t)item | (need << 13);
            cnt += __popc(pending);
            __syncwarp();
            if (cnt >= 32) {
                cnt -= 32;
                exact_item(a, src, list[cnt + lane], ox1, oy1, tile_bytes);
                __syncwarp();
            }
        }
    };
    int iA, iB, iC, iD;
    uint2 mA = next_geom(iA);
    Taps TA = load_taps(a, src, mA);
    uint2 mB = next_geom(iB);
#pragma unroll 1
    for (int it = 0; it + 1 < NIT; it += 2) {
        const Taps TB = load_taps(a, src, mB);
        const uint2 mC = next_geom(iC);
        step(iA, mA, TA);
        TA = load_taps(a, src, mC);
        const uint2 mD = next_geom(iD);
        step(iB, mB, TB);
        mA = mC; mB = mD; iA = iC; iB = iD;
    }
    if (NIT & 1) step(iA, mA, TA);
    if (lane < cnt) exact_item(a, src, list[lane], ox1, oy1, tile_bytes);
    __syncthreads();

    // one warp per frame from here on (two frames per warp): level-0 store + level 1 from the shared-memory tile
#pragma unroll 1
    for (int fw = warp; fw < FT_FR; fw += 8) {
        const int frame = f0 + fw;
        if (frame >= a.n_frames) break;
        warp_pyr_tile<true, FT_TH1>(tiles + fw * FT_TILE_WORDS, a.l1 + (size_t)frame * a.l1_stride, a.w1, a.h1, ox1, oy1, lane,
                            a.l0 + (size_t)frame * a.l0_stride, a.n);
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
k_pyr_down_w(const uint8_t* __restrict__ src, size_t src_stride, int sw, int sh, uint8_t* __restrict__ dst,
             size_t dst_stride, int dw, int dh, int tiles_x, int tiles_per_frame, int n_tiles) {
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
        unsigned off = mis + (unsigned)(2 * oy1 - 2) * (unsigned)sw + (unsigned)x0;             // byte offset of the row's region
#pragma unroll 5
        for (int r = 0; r < FT_PYR_RH; ++r, off += sw) {
            const uint32_t* ap = s4 + (off >> 2);
            const unsigned sh8 = 8u * (off & 3u);
            tw[r * FT_RWW + lane] = __funnelshift_r(__ldg(ap), __ldg(ap + 1), sh8);
            if (lane == 0) tw[r * FT_RWW + 32] = __funnelshift_r(__ldg(ap + 32), __ldg(ap + 33), sh8);
        }
    } else {
    const bool in_k = xk >= 0 && xk + 3 < sw, in_32 = xk32 + 3 < sw;
#pragma unroll 5
    for (int r = 0; r < FT_PYR_RH; ++r) {
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
    }
    __syncwarp();
    warp_pyr_tile<false, FT_PYR_TH1>(s_tile[warp], dst + (size_t)frame * dst_stride, dw, dh, ox1, oy1, lane, nullptr, 0);
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
    static bool attr_set[64] = {};   // per device: the attribute belongs to the device's context
    if (!attr_set[h->device & 63]) {
        RF_CUDA(h, cudaFuncSetAttribute(k_scan16_to_l0l1, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
        attr_set[h->device & 63] = true;
    }
    FusedArgs a;
    a.rawi = reinterpret_cast<const uint4*>(d_rawi); a.Wp = rf_fused_wp(h); a.A = h->cfg.azimuths; a.group_stride = (size_t)a.A * a.Wp;
    a.map2 = h->map2; a.n = h->n;
    a.l0 = fs.lvl[0]; a.l0_stride = fs.lvl_stride[0];
    a.l1 = fs.lvl[1]; a.l1_stride = fs.lvl_stride[1]; a.w1 = fs.w[1]; a.h1 = fs.h[1];
    a.n_frames = n_frames;
    dim3 grd((n_frames + FT_FR - 1) / FT_FR, (fs.w[1] + FT_TW1 - 1) / FT_TW1, (fs.h[1] + FT_TH1 - 1) / FT_TH1);
    k_scan16_to_l0l1<<<grd, 256, FT_SMEM_BYTES, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// levels first_level .. n_levels-1 from their predecessors
int rf_launch_pyr_levels(rf_handle* h, const FrameSet& fs, int first_level, int n_frames) {
    for (int l = first_level; l < fs.n_levels; ++l) {
        const int tiles_x = (fs.w[l] + FT_TW1 - 1) / FT_TW1, tiles_y = (fs.h[l] + FT_PYR_TH1 - 1) / FT_PYR_TH1;
        const int n_tiles = tiles_x * tiles_y * n_frames;
        k_pyr_down_w<<<(n_tiles + 3) / 4, 128, 0, h->stream>>>(fs.lvl[l - 1], fs.lvl_stride[l - 1], fs.w[l - 1], fs.h[l - 1],
                                                               fs.lvl[l], fs.lvl_stride[l], fs.w[l], fs.h[l], tiles_x,
                                                               tiles_x * tiles_y, n_tiles);
        RF_CHECK_LAUNCH(h);
    }
    return RF_OK;
}
