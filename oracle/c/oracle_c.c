/*
 * oracle_c.c — TEST INFRASTRUCTURE ONLY (never linked or called by the product).
 *
 * Plain-C CPU restatement of the arithmetic that the reference's hot path bottoms
 * out in.  The reference (Samleo8/RadarSLAMPy) is pure Python glue; the arithmetic
 * lives in third-party wheels that are NOT under /root/reference:
 *     OpenCV   4.13.0  cv2.warpPolar, cv2.calcOpticalFlowPyrLK   (unpinned upstream)
 *     SciPy    1.18.1  cdist, find_peaks                          (pinned 1.7.3 upstream)
 *     networkx 3.6.1   find_cliques over CPython 3.12 sets        (pinned 2.8 upstream)
 * Each function below restates the published algorithm of one of those calls and
 * cites the reference call site it stands in for.  tests/test_oracle_*.py pin every
 * function against the live library (and, in the authoring container, against the
 * unmodified reference imported through oracle/ref_import.py).
 *
 * Build: oracle/build_oracle.py  ->  oracle/_build/liboracle_c.so
 *        gcc -O2 -ffp-contract=off -fno-fast-math -mfma (fmaf must be a true FMA; every
 *        other expression must NOT be contracted).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

static inline int cv_round_f(float v) { return (int)lrintf(v); }   /* ties-to-even */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}

/* ------------------------------------------------------------------------------------
 * a1  parseData.py:17-53  extractDataFromRadarImage: power = raw[:, 11:].astype(f32)/255.
 *     then [:, :Wused].  True IEEE division.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_extract_polar(const uint8_t* raw, int A, int Wtot, int Wused, float* polar) {
    for (int a = 0; a < A; ++a)
        for (int r = 0; r < Wused; ++r)
            polar[(size_t)a * Wused + r] = (float)raw[(size_t)a * Wtot + 11 + r] / 255.0f;
}

/* ------------------------------------------------------------------------------------
 * a2  parseData.py:100-135  cv2.warpPolar(polar, (2R,2R), (R,R), R,
 *         WARP_POLAR_LINEAR | WARP_INVERSE_MAP | INTER_LINEAR | WARP_FILL_OUTLIERS)
 *     OpenCV: copyMakeBorder(BORDER_WRAP, 1 row top/bottom); per destination pixel
 *     cartToPolar (magnitude + fastAtan32f polynomial in degrees, FMA Horner on
 *     AVX2/AVX-512 hosts); rho = mag / Kmag, phi = ang / Kangle in double;
 *     remap INTER_LINEAR with 5-bit fixed-point coordinates, BORDER_CONSTANT(0),
 *     bilinear sum NOT contracted.
 *     polar: [A][W] f32 row-major (A azimuth rows).  cart: [2R][2R].
 * ---------------------------------------------------------------------------------- */
static inline float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y);
    float mn = ax < ay ? ax : ay, mx = ax < ay ? ay : ax;
    float c = mn / (mx + (float)DBL_EPSILON);
    float c2 = c * c;
    float a = fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
    if (ax < ay) a = 90.0f - a;
    if (x < 0) a = 180.0f - a;
    if (y < 0) a = 360.0f - a;
    return a;
}

ORC_API void orc_warp_polar_inverse(const float* polar, int A, int W, int R, float* cart) {
    const int N = 2 * R;
    const float cx = (float)R, cy = (float)R;
    const double Kangle = (2.0 * M_PI) / A;
    const double Kmag = (double)R / W;
    const float rad = (float)(M_PI / 180.0);
    const int SH = A + 2; /* wrapped source height */
    for (int y = 0; y < N; ++y) {
        float dy = (float)y - cy;
        for (int x = 0; x < N; ++x) {
            float dx = (float)x - cx;
            float mag = sqrtf(dx * dx + dy * dy);
            float ang = fast_atan2_deg(dy, dx) * rad;
            float mapx = (float)((double)mag / Kmag);
            float mapy = (float)((double)ang / Kangle) + 1.0f;
            int sx = cv_round_f(mapx * 32.0f), sy = cv_round_f(mapy * 32.0f);
            int ix = sx >> 5, iy = sy >> 5;
            float fx = (float)(sx & 31) / 32.0f, fy = (float)(sy & 31) / 32.0f;
            float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx;
            float w10 = fy * (1.0f - fx), w11 = fy * fx;
            float v[4];
            for (int t = 0; t < 4; ++t) {
                int yy = iy + (t >> 1), xx = ix + (t & 1);
                if (xx < 0 || xx >= W || yy < 0 || yy >= SH) { v[t] = 0.0f; continue; }
                int ar = yy - 1; if (ar < 0) ar += A; if (ar >= A) ar -= A;
                v[t] = polar[(size_t)ar * W + xx];
            }
            float o = v[0] * w00;
            o = o + v[1] * w01;
            o = o + v[2] * w10;
            o = o + v[3] * w11;
            cart[(size_t)y * N + x] = o;
        }
    }
}

/* a3  getTransformKLT.py:356-357  (img * 255).astype(np.uint8): f32 multiply, truncate.
 *     NumPy's f32->u8 cast of out-of-range values is UB; inputs are in [0,1]. */
ORC_API void orc_to_u8(const float* img, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)(int)(img[i] * 255.0f);
}

/* ------------------------------------------------------------------------------------
 * a4  getTransformKLT.py:359-360  cv2.calcOpticalFlowPyrLK (winSize 15x15, maxLevel 3,
 *     criteria (EPS|COUNT, 10, 0.03), flags 0, minEigThreshold 1e-4).
 *     pyrDown: separable [1 4 6 4 1], BORDER_REFLECT_101, (s + 128) >> 8.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) {
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    int* row = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; ++y) {
        for (int j = 0; j < 5; ++j) {
            int sy = reflect101(2 * y - 2 + j, h);
            for (int x = 0; x < dw; ++x) {
                int s = 0;
                for (int i = 0; i < 5; ++i) s += k[i] * src[(size_t)sy * w + reflect101(2 * x - 2 + i, w)];
                row[j * dw + x] = s;
            }
        }
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int j = 0; j < 5; ++j) s += k[j] * row[j * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
        }
    }
    free(row);
}

/* Scharr derivative at (x,y), REFLECT_101 at the image edge (cv::calcSharrDeriv). */
static inline void scharr_at(const uint8_t* I, int w, int h, int x, int y, int* gx, int* gy) {
    int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    int a = I[(size_t)ym * w + xm], b = I[(size_t)ym * w + x], c = I[(size_t)ym * w + xp];
    int d = I[(size_t)y * w + xm], f = I[(size_t)y * w + xp];
    int g = I[(size_t)yp * w + xm], hh = I[(size_t)yp * w + x], i = I[(size_t)yp * w + xp];
    *gx = (3 * c + 10 * f + 3 * i) - (3 * a + 10 * d + 3 * g);
    *gy = (3 * g + 10 * hh + 3 * i) - (3 * a + 10 * b + 3 * c);
}

ORC_API void orc_scharr(const uint8_t* I, int w, int h, int16_t* dxdy) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int gx, gy; scharr_at(I, w, h, x, y, &gx, &gy);
            dxdy[((size_t)y * w + x) * 2] = (int16_t)gx;
            dxdy[((size_t)y * w + x) * 2 + 1] = (int16_t)gy;
        }
}

/* padded image tap: REFLECT_101 within the 15-px border; padded derivative tap: 0 */
static inline int img_tap(const uint8_t* I, int w, int h, int x, int y) {
    return I[(size_t)reflect101(y, h) * w + reflect101(x, w)];
}
static inline void der_tap(const uint8_t* I, int w, int h, int x, int y, int* gx, int* gy) {
    if (x < 0 || x >= w || y < 0 || y >= h) { *gx = 0; *gy = 0; return; }
    scharr_at(I, w, h, x, y, gx, gy);
}
#define DESCALE(v, n) (((v) + (1 << ((n)-1))) >> (n))

/* One pyramid level for one point.  prev/next are that level's images. */
static void lk_level(const uint8_t* I, const uint8_t* J, int w, int h, int level, int max_level,
                     const float* pt, float* nextpt, uint8_t* status, float* err,
                     int max_count, double eps2, float min_eig_thr) {
    const int WIN = 15;
    const float half = 7.0f;
    const float FLT_SCALE = 1.0f / (1 << 20);
    float px = pt[0] * (float)(1.0 / (1 << level)), py = pt[1] * (float)(1.0 / (1 << level));
    float nx, ny;
    if (level == max_level) { nx = px; ny = py; }
    else { nx = nextpt[0] * 2.0f; ny = nextpt[1] * 2.0f; }
    nextpt[0] = nx; nextpt[1] = ny;
    px -= half; py -= half;
    int ipx = cv_floor_f(px), ipy = cv_floor_f(py);
    if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) {
        if (level == 0) { *status = 0; *err = 0; }
        return;
    }
    float a = px - ipx, b = py - ipy;
    int iw00 = cv_round_f((1.f - a) * (1.f - b) * 16384.f);
    int iw01 = cv_round_f(a * (1.f - b) * 16384.f);
    int iw10 = cv_round_f((1.f - a) * b * 16384.f);
    int iw11 = 16384 - iw00 - iw01 - iw10;
    short Iw[15 * 15], Ix[15 * 15], Iy[15 * 15];
    float A11 = 0, A12 = 0, A22 = 0;
    for (int y = 0; y < WIN; ++y)
        for (int x = 0; x < WIN; ++x) {
            int X = ipx + x, Y = ipy + y;
            int ival = DESCALE(img_tap(I, w, h, X, Y) * iw00 + img_tap(I, w, h, X + 1, Y) * iw01 +
                               img_tap(I, w, h, X, Y + 1) * iw10 + img_tap(I, w, h, X + 1, Y + 1) * iw11, 9);
            int g00x, g00y, g01x, g01y, g10x, g10y, g11x, g11y;
            der_tap(I, w, h, X, Y, &g00x, &g00y); der_tap(I, w, h, X + 1, Y, &g01x, &g01y);
            der_tap(I, w, h, X, Y + 1, &g10x, &g10y); der_tap(I, w, h, X + 1, Y + 1, &g11x, &g11y);
            int ixval = DESCALE(g00x * iw00 + g01x * iw01 + g10x * iw10 + g11x * iw11, 14);
            int iyval = DESCALE(g00y * iw00 + g01y * iw01 + g10y * iw10 + g11y * iw11, 14);
            Iw[y * WIN + x] = (short)ival; Ix[y * WIN + x] = (short)ixval; Iy[y * WIN + x] = (short)iyval;
            A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
        }
    A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * WIN * WIN);
    if (minEig < min_eig_thr || D < FLT_EPSILON) {
        if (level == 0) *status = 0;
        return;
    }
    D = 1.f / D;
    nx -= half; ny -= half;
    float pdx = 0, pdy = 0;
    for (int j = 0; j < max_count; ++j) {
        int inx = cv_floor_f(nx), iny = cv_floor_f(ny);
        if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
            if (level == 0) *status = 0;
            break;
        }
        a = nx - inx; b = ny - iny;
        iw00 = cv_round_f((1.f - a) * (1.f - b) * 16384.f);
        iw01 = cv_round_f(a * (1.f - b) * 16384.f);
        iw10 = cv_round_f((1.f - a) * b * 16384.f);
        iw11 = 16384 - iw00 - iw01 - iw10;
        float b1 = 0, b2 = 0;
        for (int y = 0; y < WIN; ++y)
            for (int x = 0; x < WIN; ++x) {
                int X = inx + x, Y = iny + y;
                int diff = DESCALE(img_tap(J, w, h, X, Y) * iw00 + img_tap(J, w, h, X + 1, Y) * iw01 +
                                   img_tap(J, w, h, X, Y + 1) * iw10 + img_tap(J, w, h, X + 1, Y + 1) * iw11, 9)
                           - Iw[y * WIN + x];
                b1 += (float)(diff * Ix[y * WIN + x]);
                b2 += (float)(diff * Iy[y * WIN + x]);
            }
        b1 *= FLT_SCALE; b2 *= FLT_SCALE;
        float ddx = (float)((A12 * b2 - A22 * b1) * D);
        float ddy = (float)((A12 * b1 - A11 * b2) * D);
        nx += ddx; ny += ddy;
        nextpt[0] = nx + half; nextpt[1] = ny + half;
        if ((double)ddx * ddx + (double)ddy * ddy <= eps2) break;
        if (j > 0 && fabs(ddx + pdx) < 0.01 && fabs(ddy + pdy) < 0.01) {
            nextpt[0] -= ddx * 0.5f; nextpt[1] -= ddy * 0.5f;
            break;
        }
        pdx = ddx; pdy = ddy;
    }
    if (*status && level == 0) {
        float qx = nextpt[0] - half, qy = nextpt[1] - half;
        int iqx = cv_floor_f(qx), iqy = cv_floor_f(qy);
        if (iqx < -WIN || iqx >= w || iqy < -WIN || iqy >= h) { *status = 0; return; }
        float aa = qx - iqx, bb = qy - iqy;
        iw00 = cv_round_f((1.f - aa) * (1.f - bb) * 16384.f);
        iw01 = cv_round_f(aa * (1.f - bb) * 16384.f);
        iw10 = cv_round_f((1.f - aa) * bb * 16384.f);
        iw11 = 16384 - iw00 - iw01 - iw10;
        float errval = 0.f;
        for (int y = 0; y < WIN; ++y)
            for (int x = 0; x < WIN; ++x) {
                int X = iqx + x, Y = iqy + y;
                int diff = DESCALE(img_tap(J, w, h, X, Y) * iw00 + img_tap(J, w, h, X + 1, Y) * iw01 +
                                   img_tap(J, w, h, X, Y + 1) * iw10 + img_tap(J, w, h, X + 1, Y + 1) * iw11, 9)
                           - Iw[y * WIN + x];
                errval += fabsf((float)diff);
            }
        *err = errval * 1.f / (32 * WIN * WIN);
    }
}

/* Full pyramidal LK.  prev/next: level-0 u8 images (w x h).  pts: K x 2 (x,y).
 * Outputs next_pts K x 2, status K, err K (err of lost points is set to 0; cv2 leaves
 * it uninitialised).  Returns the number of pyramid levels used. */
ORC_API int orc_pyr_lk(const uint8_t* prev, const uint8_t* next, int w, int h,
                       const float* pts, int K, int max_level, int max_count, double eps,
                       float* next_pts, uint8_t* status, float* err) {
    uint8_t* P[8]; uint8_t* Q[8]; int ws[8], hs[8];
    int nl = 0;
    P[0] = (uint8_t*)prev; Q[0] = (uint8_t*)next; ws[0] = w; hs[0] = h;
    for (int l = 1; l <= max_level && l < 8; ++l) {
        int nw = (ws[l - 1] + 1) / 2, nh = (hs[l - 1] + 1) / 2;
        if (nw <= 15 || nh <= 15) break;
        ws[l] = nw; hs[l] = nh;
        P[l] = (uint8_t*)malloc((size_t)nw * nh); Q[l] = (uint8_t*)malloc((size_t)nw * nh);
        orc_pyr_down(P[l - 1], ws[l - 1], hs[l - 1], P[l]);
        orc_pyr_down(Q[l - 1], ws[l - 1], hs[l - 1], Q[l]);
        nl = l;
    }
    if (max_count < 0) max_count = 0; if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0; if (eps > 10) eps = 10;
    double eps2 = eps * eps;
    for (int i = 0; i < K; ++i) { status[i] = 1; err[i] = 0.f; next_pts[2 * i] = 0; next_pts[2 * i + 1] = 0; }
    for (int l = nl; l >= 0; --l)
        for (int i = 0; i < K; ++i)
            lk_level(P[l], Q[l], ws[l], hs[l], l, nl, pts + 2 * i, next_pts + 2 * i, status + i, err + i,
                     max_count, eps2, 1e-4f);
    for (int l = 1; l <= nl; ++l) { free(P[l]); free(Q[l]); }
    return nl + 1;
}

/* ------------------------------------------------------------------------------------
 * a6  outlierRejection.py:49-58  adjacency: cdist (f64 sqrt of sum of squared f64
 *     differences of f32 inputs), |d_prev - d_new| <= thr.  adj is K x K bytes.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_consistency_adjacency(const float* prev, const float* nw, int K, double thr, uint8_t* adj) {
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            double ax = (double)prev[2 * i] - (double)prev[2 * j], ay = (double)prev[2 * i + 1] - (double)prev[2 * j + 1];
            double bx = (double)nw[2 * i] - (double)nw[2 * j], by = (double)nw[2 * i + 1] - (double)nw[2 * j + 1];
            double s1 = ax * ax; s1 = s1 + ay * ay;
            double s2 = bx * bx; s2 = s2 + by * by;
            double d = fabs(sqrt(s1) - sqrt(s2));
            adj[(size_t)i * K + j] = d <= thr;
        }
}

/* ------------------------------------------------------------------------------------
 * a6  outlierRejection.py:63-78  nx.find_cliques (networkx 3.6.1, pivoting
 *     Bron–Kerbosch) iterated in CPython-3.12 set order; keep the FIRST clique whose
 *     size is strictly greater than all earlier ones.  The set model follows CPython's
 *     Objects/setobject.c for small non-negative ints (hash(k) == k).
 * ---------------------------------------------------------------------------------- */
#define S_EMPTY (-1)
#define S_DUMMY (-2)
typedef struct { int* t; int mask, fill, used, finger; } pyset;

static void ps_init(pyset* s) { s->t = (int*)malloc(8 * sizeof(int)); for (int i = 0; i < 8; ++i) s->t[i] = S_EMPTY; s->mask = 7; s->fill = s->used = s->finger = 0; }
static void ps_free(pyset* s) { free(s->t); s->t = 0; }
static void ps_insert_clean(int* t, int mask, int key) {
    size_t perturb = (size_t)key; size_t i = (size_t)key & mask;
    for (;;) {
        if (t[i] == S_EMPTY) { t[i] = key; return; }
        if (i + 9 <= (size_t)mask) { for (int j = 1; j <= 9; ++j) if (t[i + j] == S_EMPTY) { t[i + j] = key; return; } }
        perturb >>= 5; i = (i * 5 + 1 + perturb) & mask;
    }
}
static void ps_resize(pyset* s, int minused) {
    int newsize = 8; while (newsize <= minused) newsize <<= 1;
    int* nt = (int*)malloc(sizeof(int) * newsize);
    for (int i = 0; i < newsize; ++i) nt[i] = S_EMPTY;
    for (int i = 0; i <= s->mask; ++i) if (s->t[i] >= 0) ps_insert_clean(nt, newsize - 1, s->t[i]);
    free(s->t); s->t = nt; s->mask = newsize - 1; s->fill = s->used;
}
static int ps_contains(const pyset* s, int key) {
    size_t perturb = (size_t)key; size_t i = (size_t)key & s->mask; int mask = s->mask;
    for (;;) {
        int e = s->t[i];
        if (e == key) return 1;
        if (e == S_EMPTY) return 0;
        if (i + 9 <= (size_t)mask) { for (int j = 1; j <= 9; ++j) { e = s->t[i + j]; if (e == key) return 1; if (e == S_EMPTY) return 0; } }
        perturb >>= 5; i = (i * 5 + 1 + perturb) & mask;
    }
}
static void ps_add(pyset* s, int key) {
    size_t perturb = (size_t)key; int mask = s->mask; size_t i = (size_t)key & mask; long freeslot = -1;
    for (;;) {
        int e = s->t[i];
        if (e == key) return;
        if (e == S_EMPTY) goto found_unused_or_dummy;
        if (e == S_DUMMY && freeslot < 0) freeslot = (long)i;
        if (i + 9 <= (size_t)mask) {
            for (int j = 1; j <= 9; ++j) {
                e = s->t[i + j];
                if (e == key) return;
                if (e == S_EMPTY) { i = i + j; goto found_unused_or_dummy; }
                if (e == S_DUMMY && freeslot < 0) freeslot = (long)(i + j);
            }
        }
        perturb >>= 5; i = (i * 5 + 1 + perturb) & mask;
    }
found_unused_or_dummy:
    if (freeslot >= 0) { s->t[freeslot] = key; s->used++; return; }
    s->t[i] = key; s->fill++; s->used++;
    if ((size_t)s->fill * 5 < (size_t)mask * 3) return;
    ps_resize(s, s->used > 50000 ? s->used * 2 : s->used * 4);
}
static void ps_discard(pyset* s, int key) {
    size_t perturb = (size_t)key; int mask = s->mask; size_t i = (size_t)key & mask;
    for (;;) {
        int e = s->t[i];
        if (e == key) { s->t[i] = S_DUMMY; s->used--; return; }
        if (e == S_EMPTY) return;
        if (i + 9 <= (size_t)mask) { for (int j = 1; j <= 9; ++j) { e = s->t[i + j]; if (e == key) { s->t[i + j] = S_DUMMY; s->used--; return; } if (e == S_EMPTY) return; } }
        perturb >>= 5; i = (i * 5 + 1 + perturb) & mask;
    }
}
static int ps_pop(pyset* s) {
    int i = s->finger & s->mask;
    while (s->t[i] < 0) { i++; if (i > s->mask) i = 0; }
    int key = s->t[i]; s->t[i] = S_DUMMY; s->used--; s->finger = i + 1; return key;
}
/* set_merge of `src` into the (fresh or not) set `dst` — used by copy() */
static void ps_merge(pyset* dst, const pyset* src) {
    if ((size_t)(dst->fill + src->used) * 5 >= (size_t)dst->mask * 3) ps_resize(dst, (dst->used + src->used) * 2);
    if (dst->fill == 0 && dst->mask == src->mask && src->fill == src->used) {
        memcpy(dst->t, src->t, sizeof(int) * (src->mask + 1)); dst->fill = src->fill; dst->used = src->used; return;
    }
    if (dst->fill == 0) {
        for (int i = 0; i <= src->mask; ++i) if (src->t[i] >= 0) ps_insert_clean(dst->t, dst->mask, src->t[i]);
        dst->fill = dst->used = src->used; return;
    }
    for (int i = 0; i <= src->mask; ++i) if (src->t[i] >= 0) ps_add(dst, src->t[i]);
}
static void ps_copy(pyset* dst, const pyset* src) { ps_init(dst); ps_merge(dst, src); }
/* a & b */
static void ps_and(pyset* r, const pyset* a, const pyset* b) {
    ps_init(r);
    const pyset *it = b, *other = a;                 /* set_intersection: iterate `other`=b ... */
    if (b->used > a->used) { it = a; other = b; }    /* ... unless b is larger, then swap       */
    for (int i = 0; i <= it->mask; ++i) { int k = it->t[i]; if (k >= 0 && ps_contains(other, k)) ps_add(r, k); }
}
/* a - b */
static void ps_sub(pyset* r, const pyset* a, const pyset* b) {
    if ((a->used >> 2) > b->used) {
        ps_copy(r, a);
        for (int i = 0; i <= b->mask; ++i) if (b->t[i] >= 0) ps_discard(r, b->t[i]);
        if ((r->fill - r->used) > r->mask / 4) ps_resize(r, r->used > 50000 ? r->used * 2 : r->used * 4);
        return;
    }
    ps_init(r);
    for (int i = 0; i <= a->mask; ++i) { int k = a->t[i]; if (k >= 0 && !ps_contains(b, k)) ps_add(r, k); }
}
static int ps_and_count(const pyset* a, const pyset* b) {
    const pyset *it = b, *other = a; if (b->used > a->used) { it = a; other = b; }
    int n = 0; for (int i = 0; i <= it->mask; ++i) { int k = it->t[i]; if (k >= 0 && ps_contains(other, k)) n++; }
    return n;
}
static int choose_pivot(const pyset* subg, const pyset* cand, const pyset* adj) {
    int best = -1, bestn = -1;
    for (int i = 0; i <= subg->mask; ++i) { int u = subg->t[i]; if (u < 0) continue; int n = ps_and_count(cand, &adj[u]); if (n > bestn) { bestn = n; best = u; } }
    return best;
}

typedef struct { pyset subg, cand, ext; } frame;

/* Enumerate maximal cliques in networkx order.
 *   mode 0: full enumeration; out_clique/out_size = first strictly-largest clique
 *           (outlierRejection.py:71-75); *n_yields = total number of cliques yielded.
 *   yields_buf (optional): first `yields_cap` ints of the concatenated yields,
 *           each as [size, v0, v1, ...] — used to pin the ORDER against live networkx.
 * Returns clique size. */
static unsigned long long g_order_hash;
static unsigned long long fnv_mix(unsigned long long h, int v) {
    for (int b = 0; b < 4; ++b) { h ^= (unsigned long long)((v >> (8 * b)) & 0xFF); h *= 1099511628211ULL; }
    return h;
}
/* order-sensitive FNV-1a hash over (size, members...) of every yield of the last enumeration */
ORC_API unsigned long long orc_last_order_hash(void) { return g_order_hash; }

static int first_max_clique_impl(const uint8_t* adjm, int K, int* out_clique, long* n_yields,
                                 int* yields_buf, long yields_cap, long* yields_len, int prune, long* n_descents) {
    long nd = 0;
    g_order_hash = 14695981039346656037ULL;
    if (n_yields) *n_yields = 0; if (yields_len) *yields_len = 0;
    if (K == 0) return 0;
    pyset* adj = (pyset*)malloc(sizeof(pyset) * K);
    for (int u = 0; u < K; ++u) { ps_init(&adj[u]); for (int v = 0; v < K; ++v) if (v != u && adjm[(size_t)u * K + v]) ps_add(&adj[u], v); }
    int* Q = (int*)malloc(sizeof(int) * (K + 1)); int qn = 0;
    frame* stack = (frame*)malloc(sizeof(frame) * (K + 1)); int sp = 0;
    pyset subg, cand, ext;
    ps_init(&cand); for (int u = 0; u < K; ++u) ps_add(&cand, u);
    ps_copy(&subg, &cand);
    Q[qn++] = -1;
    int u = choose_pivot(&subg, &cand, adj);
    ps_sub(&ext, &cand, &adj[u]);
    int best = 0; long ny = 0, yl = 0;
    for (;;) {
        if (ext.used) {
            int q = ps_pop(&ext);
            ps_discard(&cand, q);
            Q[qn - 1] = q;
            pyset subg_q; ps_and(&subg_q, &subg, &adj[q]);
            if (!subg_q.used) {
                ny++;
                g_order_hash = fnv_mix(g_order_hash, qn); for (int i = 0; i < qn; ++i) g_order_hash = fnv_mix(g_order_hash, Q[i]);
                if (yields_buf && yl + qn + 1 <= yields_cap) { yields_buf[yl++] = qn; for (int i = 0; i < qn; ++i) yields_buf[yl++] = Q[i]; }
                if (qn > best) { best = qn; memcpy(out_clique, Q, sizeof(int) * qn); }
                ps_free(&subg_q);
            } else {
                pyset cand_q; ps_and(&cand_q, &cand, &adj[q]);
                /* order-safe bound: a child that cannot yield a clique LARGER than the best
                 * so far can never replace it (strict '>' at outlierRejection.py:73). */
                if (cand_q.used && !(prune && qn + cand_q.used <= best)) {
                    nd++;
                    stack[sp].subg = subg; stack[sp].cand = cand; stack[sp].ext = ext; sp++;
                    Q[qn++] = -1;
                    subg = subg_q; cand = cand_q;
                    u = choose_pivot(&subg, &cand, adj);
                    ps_sub(&ext, &cand, &adj[u]);
                } else { ps_free(&subg_q); ps_free(&cand_q); }
            }
        } else {
            qn--;
            ps_free(&subg); ps_free(&cand); ps_free(&ext);
            if (sp == 0) break;
            sp--; subg = stack[sp].subg; cand = stack[sp].cand; ext = stack[sp].ext;
        }
    }
    for (int v = 0; v < K; ++v) ps_free(&adj[v]);
    free(adj); free(Q); free(stack);
    if (n_yields) *n_yields = ny; if (yields_len) *yields_len = yl;
    if (n_descents) *n_descents = nd;
    return best;
}

ORC_API int orc_first_max_clique(const uint8_t* adjm, int K, int* out_clique, long* n_yields,
                                 int* yields_buf, long yields_cap, long* yields_len) {
    return first_max_clique_impl(adjm, K, out_clique, n_yields, yields_buf, yields_cap, yields_len, 0, 0);
}

/* Same traversal with the order-safe running-best bound; returns the identical clique
 * (tests pin this against the unpruned enumeration) and reports how many child levels
 * were actually descended. */
ORC_API int orc_first_max_clique_pruned(const uint8_t* adjm, int K, int* out_clique, long* n_descents) {
    long ny, yl;
    return first_max_clique_impl(adjm, K, out_clique, &ny, 0, 0, &yl, 1, n_descents);
}

/* ------------------------------------------------------------------------------------
 * a9  ANMS.py:5-102  ssc (Suppression via Square Covering).  keypoints: n x 3 f64
 *     (row, col, sigma) in the caller's order.  Returns m and the selected indices.
 * ---------------------------------------------------------------------------------- */
static double py_round(double x) { return nearbyint(x); } /* banker's rounding, as Python round() */

ORC_API int orc_ssc(const double* kp, int n, int num_ret, double tol, int cols, int rows, int* sel) {
    double exp1 = (double)rows + cols + 2.0 * num_ret;
    double exp2 = 4.0 * cols + 4.0 * num_ret + 4.0 * rows * num_ret + (double)rows * rows + (double)cols * cols
                  - 2.0 * rows * cols + 4.0 * (double)rows * cols * num_ret;
    double exp3 = sqrt(exp2);
    double exp4 = num_ret - 1;
    double sol1 = -py_round((exp1 + exp3) / exp4), sol2 = -py_round((exp1 - exp3) / exp4);
    double high = sol1 > sol2 ? sol1 : sol2;
    double low = floor(sqrt((double)n / num_ret));
    double prev_width = -1;
    double kmin = py_round(num_ret - num_ret * tol), kmax = py_round(num_ret + num_ret * tol);
    int* result = (int*)malloc(sizeof(int) * (n > 0 ? n : 1)); int nres = 0;
    int m = 0;
    for (;;) {
        double width = low + (high - low) / 2;
        if (width == prev_width || low > high) { m = nres; memcpy(sel, result, sizeof(int) * nres); break; }
        double c = width / 2;
        int ncc = (int)floor(cols / c), ncr = (int)floor(rows / c);
        size_t stride = (size_t)ncc + 1;
        uint8_t* cov = (uint8_t*)calloc((size_t)(ncr + 1) * stride, 1);
        nres = 0;
        int reach = (int)floor(width / c);
        for (int i = 0; i < n; ++i) {
            int row = (int)floor(kp[3 * i] / c), col = (int)floor(kp[3 * i + 1] / c);
            if (!cov[(size_t)row * stride + col]) {
                result[nres++] = i;
                int r0 = row - reach >= 0 ? row - reach : 0, r1 = row + reach <= ncr ? row + reach : ncr;
                int c0 = col - reach >= 0 ? col - reach : 0, c1 = col + reach <= ncc ? col + reach : ncc;
                for (int r = r0; r <= r1; ++r) memset(cov + (size_t)r * stride + c0, 1, (size_t)(c1 - c0 + 1));
            }
        }
        free(cov);
        if (kmin <= nres && nres <= kmax) { m = nres; memcpy(sel, result, sizeof(int) * nres); break; }
        else if (nres < kmin) high = width - 1;
        else low = width + 1;
        prev_width = width;
    }
    free(result);
    return m;
}

/* ------------------------------------------------------------------------------------
 * a12 getPointCloud.py:11-54  per azimuth scipy.signal.find_peaks (strict local maxima,
 *     plateau -> middle index), keep peaks >= mean + std (population) of peak heights.
 *     np.mean/np.std of an f32 array use NumPy's pairwise f32 summation (8 partial sums,
 *     blocks of 128), restated in np_pairwise_sum_f32.
 *     out: P x 2 int64 (az, range).  Returns P.
 * ---------------------------------------------------------------------------------- */
static float np_pairwise_sum_f32(const float* a, long n) {
    if (n < 8) { float r = 0.f; for (long i = 0; i < n; ++i) r += a[i]; return r; }
    if (n <= 128) {
        float r[8]; for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i; for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2; n2 -= n2 % 8;
    return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
}

ORC_API long orc_polar_peaks(const float* polar, int A, int W, int64_t* out, long cap) {
    long P = 0;
    int* pk = (int*)malloc(sizeof(int) * (W > 0 ? W : 1));
    float* hv = (float*)malloc(sizeof(float) * (W > 0 ? W : 1));
    for (int a = 0; a < A; ++a) {
        const float* x = polar + (size_t)a * W;
        int np_ = 0;
        int i = 1, imax = W - 1;
        while (i < imax) {
            if (x[i - 1] < x[i]) {
                int ahead = i + 1;
                while (ahead < imax && x[ahead] == x[i]) ahead++;
                if (x[ahead] < x[i]) { pk[np_++] = (i + ahead - 1) / 2; i = ahead; }
            }
            i++;
        }
        if (np_ == 0) continue;
        for (int k = 0; k < np_; ++k) hv[k] = x[pk[k]];
        float mean = np_pairwise_sum_f32(hv, np_) / (float)np_;
        for (int k = 0; k < np_; ++k) { float d = hv[k] - mean; hv[k] = d * d; }
        float var = np_pairwise_sum_f32(hv, np_) / (float)np_;
        float thr = mean + sqrtf(var);
        for (int k = 0; k < np_; ++k) if (x[pk[k]] >= thr) { if (P < cap) { out[2 * P] = a; out[2 * P + 1] = pk[k]; } P++; }
    }
    free(pk); free(hv);
    return P;
}
