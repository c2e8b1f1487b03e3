// k_solve.cu — rigid-transform estimate and the motion-distortion nonlinear least squares.
//
// Replaces:
//   getTransformKLT.calculateTransformSVD (getTransformKLT.py:129-162), Tracker.getTransform
//     (Tracker.py:108-127)                                   -> k_kabsch
//   MotionDistortionSolver.update_problem / error_vector / optimize_library
//     (motionDistortion.py:80-205, 295-325; scipy least_squares 'lm')  -> k_mds
//   MotionDistortionSolver.undistort (motionDistortion.py:127-153)     -> k_undistort
//
// One warp per frame pair.  Kabsch: the reference runs in float32 because its inputs are
// float32 (sequential f32 column means, f32 centring); those two steps are reproduced
// bit-for-bit, the 2x2 cross-covariance is then accumulated exactly in f64 and the SVD is
// replaced by the closed-form 2-D optimum theta = atan2(C10 - C01, C00 + C11), which equals
// U diag(1, det) V^T whenever that is unique.
// MDS: damped Newton iteration in f64 with an analytic Jacobian derived from the residual the
// reference defines (its own jacobian() is inconsistent with error() and unused) and the exact
// curvature of the robust loss along each residual.  The solve converges to the
// minimiser, which scipy's 'lm' reaches to ~1e-5 m.
#include <math.h>

#include "common.cuh"

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <typename T>
struct KabschArgsT {
    const T* src;           // [P][Kstride][2]   x0 (old points)
    const T* tgt;           // [P][Kstride][2]   x1 (new points)
    const uint8_t* mask;    // [P][mask_stride] or nullptr (all rows up to counts[p])
    int mask_stride;
    const int32_t* counts;  // [P] rows to scan
    int Kstride, P;
    double* R;              // [P][4]
    double* h;              // [P][2]  (pixel units)
    int32_t* n_used;        // [P]
};

typedef KabschArgsT<float> KabschArgs;

// mean / centring arithmetic in the coordinate type, as NumPy does it: float32 points (what cv2 returns) keep NumPy's float32
// sequential row sums; float64 points are summed and centred in float64
__device__ __forceinline__ float kb_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double kb_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float kb_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double kb_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float kb_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double kb_div(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(128) k_kabsch(const KabschArgsT<T> a) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= a.P) return;
    const int K = a.counts[p];
    const T* s = a.src + (size_t)p * a.Kstride * 2;
    const T* t = a.tgt + (size_t)p * a.Kstride * 2;
    const uint8_t* m = a.mask ? a.mask + (size_t)p * a.mask_stride : nullptr;
    // np.mean(axis=0) of an (N,2) array: rows are added sequentially in the array's own precision
    T s0x = 0, s0y = 0, s1x = 0, s1y = 0;
    int n = 0;
    for (int i = 0; i < K; ++i) {
        if (m && !m[i]) continue;
        s0x = kb_add(s0x, s[2 * i]); s0y = kb_add(s0y, s[2 * i + 1]);
        s1x = kb_add(s1x, t[2 * i]); s1y = kb_add(s1y, t[2 * i + 1]);
        ++n;
    }
    double R00 = 1, R01 = 0, R10 = 0, R11 = 1, hx = 0, hy = 0;
    if (n > 0) {
        const T m0x = kb_div(s0x, (T)n), m0y = kb_div(s0y, (T)n);
        const T m1x = kb_div(s1x, (T)n), m1y = kb_div(s1y, (T)n);
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        for (int i = lane; i < K; i += 32) {
            if (m && !m[i]) continue;
            const double ax = (double)kb_sub(s[2 * i], m0x), ay = (double)kb_sub(s[2 * i + 1], m0y);
            const double bx = (double)kb_sub(t[2 * i], m1x), by = (double)kb_sub(t[2 * i + 1], m1y);
            c00 += ax * bx; c01 += ax * by; c10 += ay * bx; c11 += ay * by;   // C = norm_x0^T norm_x1
        }
        c00 = warp_sum_d(c00); c01 = warp_sum_d(c01); c10 = warp_sum_d(c10); c11 = warp_sum_d(c11);
        const double sn = c10 - c01, cs = c00 + c11;
        const double nr = hypot(sn, cs);
        double c = 1.0, sgn = 0.0;
        if (nr > 0.0) { c = cs / nr; sgn = sn / nr; }
        R00 = c; R01 = -sgn; R10 = sgn; R11 = c;
        hx = (double)m0x - (R00 * (double)m1x + R01 * (double)m1y);   // h = x0_mean - R x1_mean
        hy = (double)m0y - (R10 * (double)m1x + R11 * (double)m1y);
    }
    if (lane == 0) {
        double* R = a.R + (size_t)p * 4; double* h = a.h + (size_t)p * 2;
        R[0] = R00; R[1] = R01; R[2] = R10; R[3] = R11; h[0] = hx; h[1] = hy;
        if (a.n_used) a.n_used[p] = n;
    }
}

// ------------------------------------------------------------------------------------
// motion distortion
// ------------------------------------------------------------------------------------
struct MdsArgs {
    int P;
    // explicit-problem mode (rf_mds_solve): p_w / p_jt given in metres
    const double* p_w;      // [P][Nstride][2] or nullptr
    const double* p_jt;     // [P][Nstride][2] or nullptr
    const int32_t* counts;  // [P]
    int Nstride;
    const double* T_wj0;    // [P][9]
    const double* T_wj;     // [P][9]  (explicit mode) or nullptr
    // fused mode: correspondences in pixels + inlier mask + Kabsch result + previous pose
    const float* px_old; const float* px_new; const uint8_t* mask; int mask_stride; int Kstride;
    const double* kab_R; const double* kab_h; const double* prev_pose;  // [P][3] or nullptr (identity)
    // chained mode (k_seq.cu, RawROAMSystem.py:185-209): world points come from the keyframe the features belong to
    // (Keyframe.getPrunedFeaturesGlobalPosition, Mapping.py:101-120) instead of the previous frame
    const double* kf_und;      // [P][Kstride][2] undistorted keyframe-local feature points (metres), row = original feature row
    const double* kf_pose;     // [P][3]
    const int32_t* good_src;   // [P][Kstride] original feature row of every compacted correspondence
    double center, res;
    double period, sig_p0, sig_p1, sig_v0, sig_v1, sig_v2;
    int max_iters;
    double* x_out;          // [P][6]
    int32_t* iters;         // [P]
    double* cost;           // [P]
    double* scratch;        // [P][Nstride][5]: pwx, pwy, px, py, dT
};

__device__ __forceinline__ double wrap_pi(double th) {
    // utils.normalize_angles: (th + pi) % (2 pi) - pi with Python's floored modulo
    const double two_pi = 2.0 * M_PI;
    double r = fmod(th + M_PI, two_pi);
    if (r < 0) r += two_pi;
    return r - M_PI;
}

struct MdsEval { double cost; double g[6]; double H[21]; };

#define MDS_THREADS 128
#define MDS_WARPS (MDS_THREADS / 32)
#define MDS_SMEM_PTS 256     // problems with at most this many points keep them in shared memory

// residual, gradient J^T r and curvature matrix (J^T J plus the loss's own second derivative) at x, reduced over the block (every thread gets the same
// result: per-thread partial sums -> warp shuffles -> the MDS_WARPS partials added in warp order by every thread)
__device__ void mds_eval(const double* __restrict__ pts, int N, const double x[6], const double T0inv[6], double th0,
                         const MdsArgs& a, MdsEval& o, double* red, int tid) {
    const double c = cos(x[5]), s = sin(x[5]);
    const double wp0 = 1.0 / a.sig_p0, wp1 = 1.0 / a.sig_p1;
    double v[28];              // cost | g[6] | H[21]
#pragma unroll
    for (int i = 0; i < 28; ++i) v[i] = 0;
    for (int i = tid; i < N; i += MDS_THREADS) {
        const double pwx = pts[5 * i], pwy = pts[5 * i + 1], px = pts[5 * i + 2], py = pts[5 * i + 3], dT = pts[5 * i + 4];
        const double th = x[2] * dT;
        double st, ct;
        sincos(th, &st, &ct);
        const double qx = ct * px - st * py + x[0] * dT, qy = st * px + ct * py + x[1] * dT;
        const double dx = pwx - x[3], dy = pwy - x[4];
        const double mx = c * dx + s * dy, my = -s * dx + c * dy;
        const double ex = mx - qx, ey = my - qy;
        const double ux = ex * ex * 0.5 + 1.0, uy = ey * ey * 0.5 + 1.0;
        const double rx = wp0 * log(ux), ry = wp1 * log(uy);
        v[0] += rx * rx + ry * ry;
        const double iux = 1.0 / ux, iuy = 1.0 / uy;
        const double kx = wp0 * ex * iux, ky = wp1 * ey * iuy;   // d rho / d e
        // Curvature of 0.5 rho(e)^2 along e: rho'^2 + rho rho'' with rho'' = w (1 - e^2 / 2) / u^2.  The Gauss-Newton
        // matrix keeps only rho'^2 = (2/3) of it near a small residual (rho ~ w e^2 / 2), which makes every step 1.5 x
        // too long and the iteration converge linearly (error halved per step, ~45 evaluations per solve); with the
        // second term the iteration is Newton's in e and converges quadratically once the error is below the noise
        // (10-25 evaluations).  Clamped at 0 for |e| > sqrt(2), where the robust loss is not convex.
        const double hx = fmax(kx * kx + rx * wp0 * (1.0 - 0.5 * ex * ex) * iux * iux, 0.0);
        const double hy = fmax(ky * ky + ry * wp1 * (1.0 - 0.5 * ey * ey) * iuy * iuy, 0.0);
        // d e / d x
        const double dqx_dvt = dT * (-st * px - ct * py), dqy_dvt = dT * (ct * px - st * py);
        const double Jx[6] = {-dT, 0.0, -dqx_dvt, -c, -s, my};
        const double Jy[6] = {0.0, -dT, -dqy_dvt, s, -c, -mx};
        const double gx = kx * rx, gy = ky * ry;
        int idx = 7;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            v[1 + r] += Jx[r] * gx + Jy[r] * gy;
            const double ax = Jx[r] * hx, ay = Jy[r] * hy;
#pragma unroll
            for (int cc = r; cc < 6; ++cc) v[idx++] += ax * Jx[cc] + ay * Jy[cc];
        }
    }
#pragma unroll
    for (int k = 0; k < 28; ++k) v[k] = warp_sum_d(v[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 28; ++k) red[(tid >> 5) * 28 + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 28; ++k) {
        double t = red[k];
#pragma unroll
        for (int w = 1; w < MDS_WARPS; ++w) t += red[w * 28 + k];
        v[k] = t;
    }
    __syncthreads();            // `red` is free for the next evaluation
    double cost = v[0], g[6], H[21];
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = v[1 + k];
#pragma unroll
    for (int k = 0; k < 21; ++k) H[k] = v[7 + k];
    // velocity-consistency residuals (3 rows), identical on every thread
    const double tx = x[3] - T0inv[4], ty = x[4] - T0inv[5];                // t - t0
    const double relx = T0inv[0] * tx + T0inv[1] * ty, rely = T0inv[2] * tx + T0inv[3] * ty;  // R0^T (t - t0)
    const double relth = atan2(sin(x[5] - th0), cos(x[5] - th0));           // arctan2 of the relative rotation
    const double wv0 = (double)N / a.sig_v0, wv1 = (double)N / a.sig_v1, wv2 = (double)N / a.sig_v2;
    const double ip = 1.0 / a.period;
    const double r0 = wv0 * (x[0] - relx * ip), r1 = wv1 * (x[1] - rely * ip), r2 = wv2 * wrap_pi(x[2] - relth * ip);
    cost += r0 * r0 + r1 * r1 + r2 * r2;
    {
        double J0[6] = {wv0, 0, 0, -wv0 * T0inv[0] * ip, -wv0 * T0inv[1] * ip, 0};
        double J1[6] = {0, wv1, 0, -wv1 * T0inv[2] * ip, -wv1 * T0inv[3] * ip, 0};
        double J2[6] = {0, 0, wv2, 0, 0, -wv2 * ip};
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            g[r] += J0[r] * r0 + J1[r] * r1 + J2[r] * r2;
#pragma unroll
            for (int cc = r; cc < 6; ++cc) H[idx++] += J0[r] * J0[cc] + J1[r] * J1[cc] + J2[r] * J2[cc];
        }
    }
    o.cost = cost;
#pragma unroll
    for (int k = 0; k < 6; ++k) o.g[k] = g[k];
#pragma unroll
    for (int k = 0; k < 21; ++k) o.H[k] = H[k];
}

// solve (H + lam * diag(D)) dx = -g by Cholesky; returns false if not positive definite
__device__ bool solve6(const double Hu[21], const double D[6], double lam, const double g[6], double dx[6]) {
    double A[6][6];
    int idx = 0;
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) { A[r][c] = Hu[idx]; A[c][r] = Hu[idx]; ++idx; }
    for (int r = 0; r < 6; ++r) A[r][r] += lam * D[r];
    double L[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            double sum = A[i][j];
            for (int k = 0; k < j; ++k) sum -= L[i][k] * L[j][k];
            if (i == j) { if (!(sum > 0.0)) return false; L[i][i] = sqrt(sum); }
            else L[i][j] = sum / L[j][j];
        }
    double y[6];
    for (int i = 0; i < 6; ++i) { double sum = -g[i]; for (int k = 0; k < i; ++k) sum -= L[i][k] * y[k]; y[i] = sum / L[i][i]; }
    for (int i = 5; i >= 0; --i) { double sum = y[i]; for (int k = i + 1; k < 6; ++k) sum -= L[k][i] * dx[k]; dx[i] = sum / L[i][i]; }
    return true;
}

// One CTA of MDS_THREADS threads per problem: the points are spread over the threads (one or two each at the usual 60-200
// inliers), the 6x6 solve and the step control run redundantly on every thread from block-reduced sums.
__global__ void __launch_bounds__(MDS_THREADS) k_mds(const MdsArgs a) {
    __shared__ double s_pts[MDS_SMEM_PTS * 5];
    __shared__ double s_red[MDS_WARPS * 28];
    __shared__ int s_N;
    const int tid = threadIdx.x, lane = tid & 31;
    const int p = blockIdx.x;
    if (p >= a.P) return;
    double* gpts = a.scratch + (size_t)p * a.Nstride * 5;
    double* pts = gpts;
    double T0[9], Tw[9];
    int N = 0;
    if (a.p_w) {
        // explicit problem: motionDistortion.py:80-99
        N = a.counts[p];
#pragma unroll
        for (int k = 0; k < 9; ++k) { T0[k] = a.T_wj0[(size_t)p * 9 + k]; Tw[k] = a.T_wj[(size_t)p * 9 + k]; }
        const double* pw = a.p_w + (size_t)p * a.Nstride * 2;
        const double* pj = a.p_jt + (size_t)p * a.Nstride * 2;
        if (N <= MDS_SMEM_PTS) pts = s_pts;
        for (int i = tid; i < N; i += MDS_THREADS) {
            const double x = pj[2 * i], y = pj[2 * i + 1];
            pts[5 * i] = pw[2 * i]; pts[5 * i + 1] = pw[2 * i + 1]; pts[5 * i + 2] = x; pts[5 * i + 3] = y;
            pts[5 * i + 4] = a.period * atan2(-y, -x) / (2.0 * M_PI);   // compute_time_deltas :107-124
        }
    } else {
        // fused pair: RawROAMSystem.py:190-209 for a pair whose previous frame is the keyframe
        //   p_w = prev_pose o ((good_old - centre) * res),  p_jt = (good_new - centre) * res,
        //   T_wj = prev_pose @ [[R, h * res], [0, 1]]
        double x0 = 0, y0 = 0, t0 = 0;
        if (a.prev_pose) { x0 = a.prev_pose[3 * p]; y0 = a.prev_pose[3 * p + 1]; t0 = a.prev_pose[3 * p + 2]; }
        const double c0 = cos(t0), s0 = sin(t0);
        T0[0] = c0; T0[1] = -s0; T0[2] = x0; T0[3] = s0; T0[4] = c0; T0[5] = y0; T0[6] = 0; T0[7] = 0; T0[8] = 1;
        const double* R = a.kab_R + (size_t)p * 4; const double* h = a.kab_h + (size_t)p * 2;
        const double hx = h[0] * a.res, hy = h[1] * a.res;
        Tw[0] = c0 * R[0] - s0 * R[2]; Tw[1] = c0 * R[1] - s0 * R[3]; Tw[2] = c0 * hx - s0 * hy + x0;
        Tw[3] = s0 * R[0] + c0 * R[2]; Tw[4] = s0 * R[1] + c0 * R[3]; Tw[5] = s0 * hx + c0 * hy + y0;
        Tw[6] = 0; Tw[7] = 0; Tw[8] = 1;
        const int K = a.counts[p];
        const float* po = a.px_old + (size_t)p * a.Kstride * 2;
        const float* pn = a.px_new + (size_t)p * a.Kstride * 2;
        const uint8_t* m = a.mask + (size_t)p * a.mask_stride;
        double kx = 0, ky = 0, kc = 1, ks = 0;
        const double* und = nullptr; const int32_t* gsrc = nullptr;
        if (a.kf_und) {
            und = a.kf_und + (size_t)p * a.Kstride * 2; gsrc = a.good_src + (size_t)p * a.Kstride;
            kx = a.kf_pose[3 * p]; ky = a.kf_pose[3 * p + 1]; kc = cos(a.kf_pose[3 * p + 2]); ks = sin(a.kf_pose[3 * p + 2]);
        }
        // order-preserving compaction of the inliers (warp 0; straight into shared memory when they fit)
        if (K <= MDS_SMEM_PTS) pts = s_pts;
        if (tid < 32)
        for (int i0 = 0; i0 < K; i0 += 32) {
            const int i = i0 + lane;
            const bool ok = i < K && m[i];
            const unsigned bm = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int j = N + __popc(bm & ((1u << lane) - 1));
                const double ox = ((double)po[2 * i] - a.center) * a.res, oy = ((double)po[2 * i + 1] - a.center) * a.res;
                const double x = ((double)pn[2 * i] - a.center) * a.res, y = ((double)pn[2 * i + 1] - a.center) * a.res;
                if (und) {   // p_w = R(theta_kf) u + t_kf
                    const double ux = und[2 * gsrc[i]], uy = und[2 * gsrc[i] + 1];
                    pts[5 * j] = kc * ux - ks * uy + kx; pts[5 * j + 1] = ks * ux + kc * uy + ky;
                } else {
                    pts[5 * j] = c0 * ox - s0 * oy + x0; pts[5 * j + 1] = s0 * ox + c0 * oy + y0;
                }
                pts[5 * j + 2] = x; pts[5 * j + 3] = y;
                pts[5 * j + 4] = a.period * atan2(-y, -x) / (2.0 * M_PI);
            }
            N += __popc(bm);
        }
        if (tid == 0) s_N = N;
        __syncthreads();
        N = s_N;
        if (pts != s_pts && N <= MDS_SMEM_PTS) {          // K > MDS_SMEM_PTS candidates but few inliers
            for (int i = tid; i < 5 * N; i += MDS_THREADS) s_pts[i] = gpts[i];
            pts = s_pts;
        }
    }
    __syncthreads();
    // T_wj0^{-1} = [R0^T, -R0^T t0]; keep R0^T (4) and t0 (2); theta0
    const double T0inv[6] = {T0[0], T0[3], T0[1], T0[4], T0[2], T0[5]};   // R0^T row-major, then t0
    const double th0 = atan2(T0[3], T0[0]);
    // initial guess: v0 = [dx, dy, dtheta](T_wj0^-1 T_wj) / period ; pose from T_wj  (:91, 300-302)
    double x[6];
    {
        const double tx = Tw[2] - T0[2], ty = Tw[5] - T0[5];
        const double relx = T0inv[0] * tx + T0inv[1] * ty, rely = T0inv[2] * tx + T0inv[3] * ty;
        // relative rotation = R0^T Rw ; atan2(rel[1,0], rel[0,0])
        const double r00 = T0[0] * Tw[0] + T0[3] * Tw[3], r10 = T0[1] * Tw[0] + T0[4] * Tw[3];
        x[0] = relx / a.period; x[1] = rely / a.period; x[2] = atan2(r10, r00) / a.period;
        x[3] = Tw[2]; x[4] = Tw[5]; x[5] = atan2(Tw[3], Tw[0]);
    }
    MdsEval ev;
    mds_eval(pts, N, x, T0inv, th0, a, ev, s_red, tid);
    // Levenberg-Marquardt on that matrix: the damping starts small and falls by 10 per accepted step (the Kabsch start is
    // inside the basin).  The iteration ends with the first trial step -- accepted or not -- whose largest relative
    // component is below 1e-9: the accepted steps shrink quadratically by then, and a rejected step of that size means the
    // cost cannot be resolved any further in double precision (the previous rule, 1e-13, spent a dozen rejected
    // evaluations at that floor).  1e-9 of a pose component is < 1e-7 m / 1e-9 rad against north_star's 1e-4 m / 1e-5 rad.
    double lam = 1e-6;
    int it = 0;
    bool done = false;
    for (; it < a.max_iters && !done; ++it) {
        double D[6];
        for (int k = 0, idx = 0; k < 6; ++k) { D[k] = fmax(ev.H[idx], 1e-300); idx += 6 - k; }
        double gmax = 0;
        for (int k = 0; k < 6; ++k) gmax = fmax(gmax, fabs(ev.g[k]) / sqrt(D[k]));
        if (gmax < 1e-14 * fmax(1.0, sqrt(ev.cost))) break;
        bool accepted = false;
        double stepmax = 0;
        for (int tries = 0; tries < 40; ++tries) {
            double dx[6];
            if (!solve6(ev.H, D, lam, ev.g, dx)) { lam *= 10; continue; }
            double xn[6];
            stepmax = 0;
            for (int k = 0; k < 6; ++k) { xn[k] = x[k] + dx[k]; stepmax = fmax(stepmax, fabs(dx[k]) / (fabs(x[k]) + 1e-6)); }
            MdsEval en;                      // with derivatives: an accepted trial point is the next iterate
            mds_eval(pts, N, xn, T0inv, th0, a, en, s_red, tid);
            if (en.cost <= ev.cost) {
                for (int k = 0; k < 6; ++k) x[k] = xn[k];
                ev = en;
                lam = fmax(lam * 0.1, 1e-15);
                accepted = true;
                if (stepmax < 1e-9) done = true;
                break;
            }
            if (stepmax < 1e-9) { done = true; break; }
            lam *= 10;
        }
        if (!accepted) done = true;
    }
    if (tid == 0) {
        for (int k = 0; k < 6; ++k) a.x_out[(size_t)p * 6 + k] = x[k];
        if (a.iters) a.iters[p] = it;
        if (a.cost) a.cost[p] = 0.5 * ev.cost;
    }
}

__global__ void __launch_bounds__(256)
k_undistort(const double* __restrict__ pts, int N, double vx, double vy, double vth, double period, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = pts[2 * i], y = pts[2 * i + 1];
    const double t = period * atan2(-y, -x) / (2.0 * M_PI);
    const double th = vth * t, c = cos(th), s = sin(th);
    out[2 * i] = c * x - s * y + vx * t;
    out[2 * i + 1] = s * x + c * y + vy * t;
}

// explicit per-point time offsets (MotionDistortionSolver.undistort(..., times=...), motionDistortion.py:127-153)
__global__ void __launch_bounds__(256)
k_undistort_times(const double* __restrict__ pts, const double* __restrict__ times, int N, double vx, double vy, double vth,
                  double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = pts[2 * i], y = pts[2 * i + 1], t = times[i];
    const double th = vth * t, c = cos(th), s = sin(th);
    out[2 * i] = c * x - s * y + vx * t;
    out[2 * i + 1] = s * x + c * y + vy * t;
}

// ---- launchers ------------------------------------------------------------------------
int rf_launch_kabsch(rf_handle* h, const float* d_src, const float* d_tgt, const uint8_t* d_mask, int mask_stride,
                     const int32_t* d_counts, int Kstride, int P, double* d_R, double* d_h, int32_t* d_nused) {
    KabschArgs a{d_src, d_tgt, d_mask, mask_stride, d_counts, Kstride, P, d_R, d_h, d_nused};
    k_kabsch<float><<<(P + 3) / 4, 128, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

static void fill_mds_cfg(rf_handle* h, MdsArgs& a) {
    a.period = h->cfg.mds_period;
    a.sig_p0 = h->cfg.mds_sigma_p[0]; a.sig_p1 = h->cfg.mds_sigma_p[1];
    a.sig_v0 = h->cfg.mds_sigma_v[0]; a.sig_v1 = h->cfg.mds_sigma_v[1]; a.sig_v2 = h->cfg.mds_sigma_v[2];
    a.center = (double)h->R; a.res = h->cfg.cart_res_m;
    a.max_iters = 200;
}

int rf_launch_mds_fused(rf_handle* h, const float* d_old, const float* d_new, const uint8_t* d_mask, int mask_stride,
                        const int32_t* d_counts, int Kstride, int P, const double* d_R, const double* d_h,
                        const double* d_prev_pose, double* d_scratch, double* d_x, int32_t* d_iters) {
    MdsArgs a; memset(&a, 0, sizeof(a));
    fill_mds_cfg(h, a);
    a.P = P; a.counts = d_counts; a.Nstride = Kstride; a.px_old = d_old; a.px_new = d_new; a.mask = d_mask;
    a.mask_stride = mask_stride; a.Kstride = Kstride; a.kab_R = d_R; a.kab_h = d_h; a.prev_pose = d_prev_pose;
    a.x_out = d_x; a.iters = d_iters; a.cost = nullptr; a.scratch = d_scratch;
    k_mds<<<P, MDS_THREADS, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

int rf_launch_mds_chain(rf_handle* h, const float* d_old, const float* d_new, const uint8_t* d_mask, int mask_stride,
                        const int32_t* d_counts, int Kstride, int P, const double* d_R, const double* d_h,
                        const double* d_prev_pose, const double* d_kf_und, const double* d_kf_pose, const int32_t* d_good_src,
                        double* d_scratch, double* d_x, int32_t* d_iters) {
    MdsArgs a; memset(&a, 0, sizeof(a));
    fill_mds_cfg(h, a);
    a.P = P; a.counts = d_counts; a.Nstride = Kstride; a.px_old = d_old; a.px_new = d_new; a.mask = d_mask;
    a.mask_stride = mask_stride; a.Kstride = Kstride; a.kab_R = d_R; a.kab_h = d_h; a.prev_pose = d_prev_pose;
    a.kf_und = d_kf_und; a.kf_pose = d_kf_pose; a.good_src = d_good_src;
    a.x_out = d_x; a.iters = d_iters; a.cost = nullptr; a.scratch = d_scratch;
    k_mds<<<P, MDS_THREADS, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

extern "C" {

int rf_kabsch(rf_handle* h, const float* src_xy, const float* tgt_xy, int N, double R[4], double hvec[2]) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !src_xy || !tgt_xy || !R || !hvec || N < 0) return rf_fail(h, RF_E_BADARG, "rf_kabsch: bad argument");
    size_t bp = ((size_t)N * 8 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, 2 * bp + 1024);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    float* ds = (float*)base; float* dt = (float*)(base + bp);
    int32_t* dc = (int32_t*)(base + 2 * bp);
    double* dR = (double*)(base + 2 * bp + 256); double* dh = dR + 4;
    if (N) {
        RF_CUDA(h, cudaMemcpyAsync(ds, src_xy, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
        RF_CUDA(h, cudaMemcpyAsync(dt, tgt_xy, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
    }
    RF_CUDA(h, cudaMemcpyAsync(dc, &N, 4, cudaMemcpyHostToDevice, h->stream));
    rc = rf_launch_kabsch(h, ds, dt, nullptr, 0, dc, N > 0 ? N : 1, 1, dR, dh, nullptr);
    if (rc) return rc;
    double out[6];
    RF_CUDA(h, cudaMemcpyAsync(out, dR, 48, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(R, out, 32); memcpy(hvec, out + 4, 16);
    return RF_OK;
}

int rf_kabsch_f64(rf_handle* h, const double* src_xy, const double* tgt_xy, int N, double R[4], double hvec[2]) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !src_xy || !tgt_xy || !R || !hvec || N < 0) return rf_fail(h, RF_E_BADARG, "rf_kabsch_f64: bad argument");
    size_t bp = ((size_t)N * 16 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, 2 * bp + 1024);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    double* ds = (double*)base; double* dt = (double*)(base + bp);
    int32_t* dc = (int32_t*)(base + 2 * bp);
    double* dR = (double*)(base + 2 * bp + 256); double* dh = dR + 4;
    if (N) {
        RF_CUDA(h, cudaMemcpyAsync(ds, src_xy, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
        RF_CUDA(h, cudaMemcpyAsync(dt, tgt_xy, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
    }
    RF_CUDA(h, cudaMemcpyAsync(dc, &N, 4, cudaMemcpyHostToDevice, h->stream));
    KabschArgsT<double> a;
    a.src = ds; a.tgt = dt; a.mask = nullptr; a.mask_stride = 0; a.counts = dc; a.Kstride = N > 0 ? N : 1; a.P = 1;
    a.R = dR; a.h = dh; a.n_used = nullptr;
    k_kabsch<double><<<1, 128, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    double out[6];
    RF_CUDA(h, cudaMemcpyAsync(out, dR, 48, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(R, out, 32); memcpy(hvec, out + 4, 16);
    return RF_OK;
}

int rf_mds_solve(rf_handle* h, const double T_wj0[9], const double* p_w, const double* p_jt, int N, const double T_wj[9],
                 const double* sigma_p, const double* sigma_v, double period, double x_out[6], int* iters, double* cost) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !T_wj0 || !T_wj || !x_out || N < 0 || (N > 0 && (!p_w || !p_jt)))
        return rf_fail(h, RF_E_BADARG, "rf_mds_solve: bad argument");
    const int Ns = N > 0 ? N : 1;
    size_t bp = ((size_t)Ns * 16 + 255) & ~(size_t)255, bs = ((size_t)Ns * 40 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, 2 * bp + bs + 1024);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    double* dpw = (double*)base; double* dpj = (double*)(base + bp); double* dsc = (double*)(base + 2 * bp);
    char* tail = base + 2 * bp + bs;
    double* dT0 = (double*)tail; double* dTw = dT0 + 9; double* dx = dTw + 9; double* dcost = dx + 6;
    int32_t* dc = (int32_t*)(dcost + 1); int32_t* dit = dc + 1;
    if (N) {
        RF_CUDA(h, cudaMemcpyAsync(dpw, p_w, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
        RF_CUDA(h, cudaMemcpyAsync(dpj, p_jt, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
    }
    RF_CUDA(h, cudaMemcpyAsync(dT0, T_wj0, 72, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dTw, T_wj, 72, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dc, &N, 4, cudaMemcpyHostToDevice, h->stream));
    MdsArgs a; memset(&a, 0, sizeof(a));
    fill_mds_cfg(h, a);
    if (sigma_p) { a.sig_p0 = sigma_p[0]; a.sig_p1 = sigma_p[1]; }
    if (sigma_v) { a.sig_v0 = sigma_v[0]; a.sig_v1 = sigma_v[1]; a.sig_v2 = sigma_v[2]; }
    if (period > 0) a.period = period;
    a.P = 1; a.p_w = dpw; a.p_jt = dpj; a.counts = dc; a.Nstride = Ns; a.T_wj0 = dT0; a.T_wj = dTw;
    a.x_out = dx; a.iters = dit; a.cost = dcost; a.scratch = dsc;
    k_mds<<<1, MDS_THREADS, 0, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    double xo[7]; int32_t ito;
    RF_CUDA(h, cudaMemcpyAsync(xo, dx, 56, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&ito, dit, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(x_out, xo, 48);
    if (cost) *cost = xo[6];
    if (iters) *iters = ito;
    return RF_OK;
}

int rf_mds_undistort(rf_handle* h, const double v[3], const double* pts_xy, int N, double period, double* out_xy) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !v || N < 0 || (N > 0 && (!pts_xy || !out_xy))) return rf_fail(h, RF_E_BADARG, "rf_mds_undistort: bad argument");
    if (N == 0) return RF_OK;
    size_t bp = ((size_t)N * 16 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, 2 * bp);
    if (rc) return rc;
    double* din = (double*)h->d_scratch; double* dout = (double*)((char*)h->d_scratch + bp);
    RF_CUDA(h, cudaMemcpyAsync(din, pts_xy, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
    k_undistort<<<(N + 255) / 256, 256, 0, h->stream>>>(din, N, v[0], v[1], v[2], period, dout);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(out_xy, dout, (size_t)N * 16, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

int rf_mds_undistort_times(rf_handle* h, const double v[3], const double* pts_xy, const double* times, int N, double* out_xy) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !v || N < 0 || (N > 0 && (!pts_xy || !times || !out_xy))) return rf_fail(h, RF_E_BADARG, "rf_mds_undistort_times: bad argument");
    if (N == 0) return RF_OK;
    const size_t bp = ((size_t)N * 16 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, 3 * bp);
    if (rc) return rc;
    double* din = (double*)h->d_scratch; double* dout = (double*)((char*)h->d_scratch + bp); double* dt = (double*)((char*)h->d_scratch + 2 * bp);
    RF_CUDA(h, cudaMemcpyAsync(din, pts_xy, (size_t)N * 16, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dt, times, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
    k_undistort_times<<<(N + 255) / 256, 256, 0, h->stream>>>(din, dt, N, v[0], v[1], v[2], dout);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(out_xy, dout, (size_t)N * 16, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

}  // extern "C"
