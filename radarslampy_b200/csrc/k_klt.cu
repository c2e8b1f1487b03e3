// k_klt.cu — pyramidal Lucas-Kanade tracking with OpenCV's fixed-point semantics.
//
// Replaces cv2.calcOpticalFlowPyrLK(winSize=(15,15), maxLevel=3, criteria=(EPS|COUNT,10,0.03))
// at getTransformKLT.py:359-360 and the err gating at getTransformKLT.py:364-365.
// Restated algorithm: oracle/c/oracle_c.c::lk_level (SURVEY.md Spec S-a4).
//
// Mapping: ONE WARP tracks ONE feature through all pyramid levels.  Per level the warp
//   1. stages the 18x18 u8 neighbourhood of the previous image in shared memory
//      (REFLECT_101 padding exactly like cv's padded pyramid),
//   2. evaluates the Scharr derivative on the fly for the 16x16 positions the window can
//      touch (cv computes a whole-image derivative; outside the image it is 0),
//   3. builds the Q5 image patch and Q0 derivative patches of the 15x15 window in
//      REGISTERS (8 window pixels per lane) and reduces the 2x2 normal matrix,
//   4. iterates: stages the 16x16 u8 window of the next image, forms the mismatch vector
//      with exact integer products, reduces with REDUX (split hi/lo 16 bits, exact),
//      and solves the 2x2 system in f32.
// All integer sums are exact; cv2 accumulates the same products in f32, so positions agree
// to ~5e-4 px (tolerance in north_star: 0.02 px) and status/err gating is identical.
#include <limits.h>

#include "common.cuh"

#ifndef KLT_WARPS
#define KLT_WARPS 2        // warps (features) per CTA.  2 x 8 resident CTAs: 0.637 ms in-step / 0.539 alone per 51 k tracks; 4 x 4: 0.649 / 0.551;
#endif                     // 8 x 2: 0.734 / 0.822.  Small CTAs fill the register space one-warp clique CTAs of other batches leave on an SM
#define KLT_WIN 15
#define KLT_NPIX 225
#define KLT_PER_LANE 8
#ifndef KLT_MIN_BLOCKS
#define KLT_MIN_BLOCKS 8   // resident CTAs per SM the register allocation is capped for (16 warps x 128 registers); 20 warps (96 regs) and 24 (80 regs, spills) measured equal
#endif

struct KltArgs {
    FrameSet prev, next;
    const int32_t* pair_idx;  // [P][2] frame indices (prev set, next set) or nullptr -> (0,0)
    const float* pts;         // [P][Kmax][2]
    const int32_t* counts;    // [P] or nullptr -> Kmax
    int Kmax, P;
    float* next_xy;           // [P][Kmax][2]
    uint8_t* status;          // [P][Kmax]
    float* err;               // [P][Kmax]
    int max_iters;
    double eps2;
    float min_eig, err_thr;
    int gate;
};

struct __align__(16) WarpSmem {
    uint8_t I[18][20];   // previous-image neighbourhood, origin (ip.x-1, ip.y-1); rows are 5 aligned words
    uint8_t pad_[8];     // keeps D and J 16-byte aligned (vector stores in the staging fast paths)
    short2 D[16][16];    // Scharr (dx, dy) at (ip.x + c, ip.y + r)
    uint16_t J[16][16];  // next-image window as horizontal PAIRS: J[y][x] = pixel (y, x) | pixel (y, x + 1) << 8, origin (in.x, in.y);
                         // a bilinear sample is then two 16-bit loads and two dp2a (signed Q14 weights x unsigned bytes)
                         // instead of four byte loads and four multiplies; column 15 pairs with a byte that is never sampled
};

// exact warp sum of int32 values whose total may exceed 32 bits
__device__ __forceinline__ long long warp_sum_exact(int v) {
    int lo = v & 0xFFFF, hi = v >> 16;  // v == (hi << 16) + lo, lo in [0, 65535]
    int slo = __reduce_add_sync(0xffffffffu, lo);
    int shi = __reduce_add_sync(0xffffffffu, hi);
    return ((long long)shi << 16) + (long long)slo;
}

__device__ __forceinline__ void q14_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, b);
    w00 = cv_round(__fmul_rn(__fmul_rn(ia, ib), 16384.f));
    w01 = cv_round(__fmul_rn(__fmul_rn(a, ib), 16384.f));
    w10 = cv_round(__fmul_rn(__fmul_rn(ia, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// Stage the 16x16 next-image window at (ox, oy).  Interior windows (the common case) are read as aligned 32-bit
// words: lane = 2 * row + half moves 8 consecutive bytes (three aligned loads, two funnel shifts, one 64-bit
// shared store).  The third word may reach 3 bytes past the window: into the same row, the next row, or the slack
// every level allocation carries.  Windows that touch the border take the REFLECT_101 byte path.
// the eight pixel pairs starting at bytes 0 .. 7 of (lo | hi << 32), `nxt` holding byte 8: one 128-bit shared store
__device__ __forceinline__ uint4 j_pairs(uint32_t lo, uint32_t hi, uint32_t nxt) {
    return make_uint4(__byte_perm(lo, hi, 0x2110), __byte_perm(lo, hi, 0x4332), __byte_perm(hi, nxt, 0x2110), __byte_perm(hi, nxt, 0x4332));
}

__device__ __forceinline__ void load_J(WarpSmem& s, const uint8_t* __restrict__ J, int pitch, int w, int h, int ox, int oy, int lane) {
    if (ox >= 0 && oy >= 0 && ox + 16 <= w && oy + 16 <= h) {
        const int r = lane >> 1, c0 = (lane & 1) * 8;
        const uint8_t* p = J + (size_t)(oy + r) * pitch + ox + c0;
        const unsigned m = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3u);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(p - m);
        const uint32_t w0 = __ldg(ap), w1 = __ldg(ap + 1), w2 = __ldg(ap + 2);
        *reinterpret_cast<uint4*>(&s.J[r][c0]) = j_pairs(__funnelshift_r(w0, w1, 8 * m), __funnelshift_r(w1, w2, 8 * m), __funnelshift_r(w2, 0u, 8 * m));
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = lane + 32 * k, r = idx >> 4, c = idx & 15;
            const uint8_t* row = J + (size_t)reflect101(oy + r, h) * pitch;
            // (the partner of column 15 is never sampled; any in-row byte will do)
            s.J[r][c] = (uint16_t)(__ldg(row + reflect101(ox + c, w)) | ((unsigned)__ldg(row + reflect101(min(ox + c + 1, ox + 15), w)) << 8));
        }
    }
    __syncwarp();
}

// wtop = w00 | w01 << 16, wbot = w10 | w11 << 16 as signed 16-bit halves (w11 = 2^14 - w00 - w01 - w10 can be -1)
// signed 16-bit halves of `a` x unsigned bytes 0, 1 of `b`, accumulated into c (the CUDA intrinsics only offer same-signedness forms)
__device__ __forceinline__ int dp2a_lo_s16_u8(int a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int j_sample(const WarpSmem& s, int y, int x, int wtop, int wbot) {
    const int v = dp2a_lo_s16_u8(wbot, (unsigned)s.J[y + 1][x], dp2a_lo_s16_u8(wtop, (unsigned)s.J[y][x], 0));
    return (v + (1 << 8)) >> 9;
}
__device__ __forceinline__ int pack_w(int lo, int hi) { return (lo & 0xFFFF) | (hi << 16); }

__global__ void __launch_bounds__(KLT_WARPS * 32, KLT_MIN_BLOCKS) k_klt(const KltArgs a) {
    __shared__ WarpSmem smem[KLT_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * KLT_WARPS + wib;
    const int pair = (int)(gw / a.Kmax), k = (int)(gw - (long long)pair * a.Kmax);
    if (pair >= a.P) return;
    const int count = a.counts ? a.counts[pair] : a.Kmax;
    if (k >= count) return;
    WarpSmem& s = smem[wib];
    const int fprev = a.pair_idx ? a.pair_idx[2 * pair] : 0;
    const int fnext = a.pair_idx ? a.pair_idx[2 * pair + 1] : 0;
    const size_t o = (size_t)pair * a.Kmax + k;
    const float ptx = a.pts[2 * o], pty = a.pts[2 * o + 1];
    const float half = 7.0f;
    const float FLT_SCALE = 1.f / (1 << 20);

    float nx = 0.f, ny = 0.f;  // nextPts[ptidx]
    int status = 1;
    float err = 0.f;
    const int top = a.prev.n_levels - 1;

    // per-lane window pixels (idx = lane + 32*j < 225)
    int wy[KLT_PER_LANE], wx[KLT_PER_LANE];
#pragma unroll
    for (int j = 0; j < KLT_PER_LANE; ++j) {
        const int idx = lane + 32 * j;
        wy[j] = idx / KLT_WIN; wx[j] = idx - wy[j] * KLT_WIN;
    }

    for (int l = top; l >= 0; --l) {
        const int w = a.prev.w[l], h = a.prev.h[l], pitch = a.prev.pitch[l];   // prev and next sets share one geometry
        const uint8_t* __restrict__ I = a.prev.lvl[l] + (size_t)fprev * a.prev.lvl_stride[l];
        const uint8_t* __restrict__ J = a.next.lvl[l] + (size_t)fnext * a.next.lvl_stride[l];
        const float sc = 1.f / (float)(1 << l);
        float px = __fmul_rn(ptx, sc), py = __fmul_rn(pty, sc);
        if (l == top) { nx = px; ny = py; }
        else { nx = __fmul_rn(nx, 2.f); ny = __fmul_rn(ny, 2.f); }
        px = __fsub_rn(px, half); py = __fsub_rn(py, half);
        const int ipx = cv_floor(px), ipy = cv_floor(py);
        if (ipx < -KLT_WIN || ipx >= w || ipy < -KLT_WIN || ipy >= h) {
            if (l == 0) { status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        q14_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        __syncwarp();
        // The next-image window of the FIRST iteration sits at floor(nextPt - halfWin), known now: when it and the
        // previous-image neighbourhood are both interior, their loads are issued together, so the level pays one global
        // round trip before its first iteration instead of two (this kernel is latency-bound at 25 % occupancy).
        int jx = INT_MIN, jy = INT_MIN;          // origin of the next-image window currently staged in s.J (none yet at this level)
        const int jx0 = cv_floor(__fsub_rn(nx, half)), jy0 = cv_floor(__fsub_rn(ny, half));
        const bool j_early = jx0 >= 0 && jy0 >= 0 && jx0 + 16 <= w && jy0 + 16 <= h;
        // 1. 18x18 neighbourhood (324 bytes)
        if (ipx >= 1 && ipy >= 1 && ipx + 17 <= w && ipy + 17 <= h) {
            // interior: lane r < 18 moves row r (18 bytes -> five words of the 20-byte shared row) from aligned loads
            uint32_t jw0 = 0, jw1 = 0, jw2 = 0;
            unsigned jm = 0;
            const int jr = lane >> 1, jc0 = (lane & 1) * 8;
            if (j_early) {
                const uint8_t* pj = J + (size_t)(jy0 + jr) * pitch + jx0 + jc0;
                jm = (unsigned)(reinterpret_cast<uintptr_t>(pj) & 3u);
                const uint32_t* apj = reinterpret_cast<const uint32_t*>(pj - jm);
                jw0 = __ldg(apj); jw1 = __ldg(apj + 1); jw2 = __ldg(apj + 2);
            }
            if (lane < 18) {
                const uint8_t* p = I + (size_t)(ipy - 1 + lane) * pitch + (ipx - 1);
                const unsigned m = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3u);
                const uint32_t* ap = reinterpret_cast<const uint32_t*>(p - m);
                uint32_t q[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) q[i] = __ldg(ap + i);
                uint32_t* dst = reinterpret_cast<uint32_t*>(&s.I[lane][0]);
#pragma unroll
                for (int i = 0; i < 5; ++i) dst[i] = __funnelshift_r(q[i], q[i + 1], 8 * m);
            }
            if (j_early) {
                *reinterpret_cast<uint4*>(&s.J[jr][jc0]) = j_pairs(__funnelshift_r(jw0, jw1, 8 * jm), __funnelshift_r(jw1, jw2, 8 * jm),
                                                                   __funnelshift_r(jw2, 0u, 8 * jm));
                jx = jx0; jy = jy0;
            }
        } else {
            for (int idx = lane; idx < 18 * 18; idx += 32) {
                const int r = idx / 18, c = idx - r * 18;
                s.I[r][c] = __ldg(I + (size_t)reflect101(ipy - 1 + r, h) * pitch + reflect101(ipx - 1 + c, w));
            }
        }
        __syncwarp();
        // 2. Scharr at the 16x16 positions (cv::calcSharrDeriv; zero outside the image)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = lane + 32 * j, r = idx >> 4, c = idx & 15;
            const int X = ipx + c, Y = ipy + r;
            int gx = 0, gy = 0;
            if (X >= 0 && X < w && Y >= 0 && Y < h) {
                const int p00 = s.I[r][c], p01 = s.I[r][c + 1], p02 = s.I[r][c + 2];
                const int p10 = s.I[r + 1][c], p12 = s.I[r + 1][c + 2];
                const int p20 = s.I[r + 2][c], p21 = s.I[r + 2][c + 1], p22 = s.I[r + 2][c + 2];
                gx = (3 * p02 + 10 * p12 + 3 * p22) - (3 * p00 + 10 * p10 + 3 * p20);
                gy = (3 * p20 + 10 * p21 + 3 * p22) - (3 * p00 + 10 * p01 + 3 * p02);
            }
            s.D[r][c] = make_short2((short)gx, (short)gy);
        }
        __syncwarp();
        // 3. window patches in registers + normal matrix
        int Iw[KLT_PER_LANE], Ix[KLT_PER_LANE], Iy[KLT_PER_LANE];
        int a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
        for (int j = 0; j < KLT_PER_LANE; ++j) {
            Iw[j] = Ix[j] = Iy[j] = 0;
            if (lane + 32 * j < KLT_NPIX) {
                const int y = wy[j], x = wx[j];
                const int iv = s.I[y + 1][x + 1] * w00 + s.I[y + 1][x + 2] * w01 + s.I[y + 2][x + 1] * w10 +
                               s.I[y + 2][x + 2] * w11;
                const short2 d00 = s.D[y][x], d01 = s.D[y][x + 1], d10 = s.D[y + 1][x], d11 = s.D[y + 1][x + 1];
                const int ixv = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
                const int iyv = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
                Iw[j] = (iv + (1 << 8)) >> 9;
                Ix[j] = ixv; Iy[j] = iyv;
                a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv;
            }
        }
        const float A11 = __fmul_rn((float)warp_sum_exact(a11), FLT_SCALE);
        const float A12 = __fmul_rn((float)warp_sum_exact(a12), FLT_SCALE);
        const float A22 = __fmul_rn((float)warp_sum_exact(a22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * KLT_NPIX));
        if (minEig < a.min_eig || D < 1.1920928955078125e-07f) {
            if (l == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        float qx = __fsub_rn(nx, half), qy = __fsub_rn(ny, half);  // nextPt -= halfWin
        float pdx = 0.f, pdy = 0.f;
        // 4. iterations
        for (int it = 0; it < a.max_iters; ++it) {
            const int inx = cv_floor(qx), iny = cv_floor(qy);
            if (inx < -KLT_WIN || inx >= w || iny < -KLT_WIN || iny >= h) {
                if (l == 0) status = 0;
                break;
            }
            q14_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            // sub-pixel updates usually keep the integer window origin: the staged window is still the right one
            if (inx != jx || iny != jy) {
                __syncwarp();
                load_J(s, J, pitch, w, h, inx, iny, lane);
                jx = inx; jy = iny;
            }
            int b1 = 0, b2 = 0;
            const int wtop = pack_w(w00, w01), wbot = pack_w(w10, w11);
#pragma unroll
            for (int j = 0; j < KLT_PER_LANE; ++j) {
                if (lane + 32 * j < KLT_NPIX) {
                    const int diff = j_sample(s, wy[j], wx[j], wtop, wbot) - Iw[j];
                    b1 += diff * Ix[j]; b2 += diff * Iy[j];
                }
            }
            const float B1 = __fmul_rn((float)warp_sum_exact(b1), FLT_SCALE);
            const float B2 = __fmul_rn((float)warp_sum_exact(b2), FLT_SCALE);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), D);
            qx = __fadd_rn(qx, ddx); qy = __fadd_rn(qy, ddy);
            nx = __fadd_rn(qx, half); ny = __fadd_rn(qy, half);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= a.eps2) break;
            if (it > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(ddx, 0.5f)); ny = __fsub_rn(ny, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        // 5. residual of the final position (level 0 only)
        if (status && l == 0) {
            const float ex = __fsub_rn(nx, half), ey = __fsub_rn(ny, half);
            const int iex = cv_floor(ex), iey = cv_floor(ey);
            if (iex < -KLT_WIN || iex >= w || iey < -KLT_WIN || iey >= h) {
                status = 0;
            } else {
                q14_weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
                if (iex != jx || iey != jy) {
                    __syncwarp();
                    load_J(s, J, pitch, w, h, iex, iey, lane);
                }
                int e = 0;
                const int wtop = pack_w(w00, w01), wbot = pack_w(w10, w11);
#pragma unroll
                for (int j = 0; j < KLT_PER_LANE; ++j)
                    if (lane + 32 * j < KLT_NPIX) e += abs(j_sample(s, wy[j], wx[j], wtop, wbot) - Iw[j]);
                err = __fdiv_rn((float)warp_sum_exact(e), (float)(32 * KLT_NPIX));
            }
        }
    }
    if (lane == 0) {
        a.next_xy[2 * o] = nx; a.next_xy[2 * o + 1] = ny;
        // getTransformKLT.py:365  status &= (err < ERR_THRESHOLD)
        if (a.gate && !(err < a.err_thr)) status = 0;
        a.status[o] = (uint8_t)status;
        a.err[o] = err;
    }
}

int rf_launch_klt(rf_handle* h, const FrameSet& prev, const FrameSet& next, const int32_t* d_pair_idx, const float* d_pts,
                  const int32_t* d_counts, int Kmax, int P, float* d_next, uint8_t* d_status, float* d_err, int gate) {
    KltArgs a;
    a.prev = prev; a.next = next; a.pair_idx = d_pair_idx; a.pts = d_pts; a.counts = d_counts;
    a.Kmax = Kmax; a.P = P; a.next_xy = d_next; a.status = d_status; a.err = d_err;
    int mc = h->cfg.klt_max_iters; mc = mc < 0 ? 0 : (mc > 100 ? 100 : mc);  // cv clamps the criteria
    double eps = h->cfg.klt_eps; eps = eps < 0 ? 0 : (eps > 10 ? 10 : eps);
    a.max_iters = mc; a.eps2 = eps * eps;
    a.min_eig = h->cfg.klt_min_eig; a.err_thr = h->cfg.klt_err_thr; a.gate = gate;
    long long warps = (long long)P * Kmax;
    if (warps == 0) return RF_OK;
    unsigned blocks = (unsigned)((warps + KLT_WARPS - 1) / KLT_WARPS);
    k_klt<<<blocks, KLT_WARPS * 32, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

extern "C" int rf_klt(rf_handle* h, const rf_frame* prev, const rf_frame* next, const float* pts, int K,
                      int apply_err_gate, float* next_xy, uint8_t* status, float* err) {
    if (!h || !prev || !next || !pts || !next_xy || !status || !err || K < 0)
        return rf_fail(h, RF_E_BADARG, "rf_klt: null argument");
    if (K == 0) return RF_OK;
    size_t bp = (size_t)K * 2 * sizeof(float), bs = (size_t)K, be = (size_t)K * sizeof(float);
    size_t off_next = (bp + 255) & ~(size_t)255, off_err = off_next + ((bp + 255) & ~(size_t)255);
    size_t off_st = off_err + ((be + 255) & ~(size_t)255), total = off_st + ((bs + 255) & ~(size_t)255);
    int rc = rf_ensure_scratch(h, total);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    RF_CUDA(h, cudaMemcpyAsync(base, pts, bp, cudaMemcpyHostToDevice, h->stream));
    rc = rf_launch_klt(h, prev->fs, next->fs, nullptr, (const float*)base, nullptr, K, 1, (float*)(base + off_next),
                       (uint8_t*)(base + off_st), (float*)(base + off_err), apply_err_gate);
    if (rc) return rc;
    RF_CUDA(h, cudaMemcpyAsync(next_xy, base + off_next, bp, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(err, base + off_err, be, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(status, base + off_st, bs, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}
