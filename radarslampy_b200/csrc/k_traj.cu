// k_traj.cu — trajectory chaining: poses from per-pair relative transforms as a parallel
// prefix product over SE(2), so the concatenation of gathered multi-GPU results is one launch.
//
// Replaces (reference file:line):
//   trajectoryPlotting.py:37-60   Trajectory.appendRelativeTransform: pose_transform = A @ pose_transform
//   trajectoryPlotting.py:27-35   Trajectory.appendRelativeDeltas:    body-frame composition T = T @ A
//   utils.py:46-103               convertPoseToTransform / convertTransformToPose
#include <math.h>

#include "common.cuh"

namespace {
struct Se2 { double c, s, x, y; };   // [c -s x; s c y; 0 0 1]
__device__ __forceinline__ Se2 mul(const Se2& a, const Se2& b) {   // a @ b
    Se2 r;
    r.c = a.c * b.c - a.s * b.s;
    r.s = a.s * b.c + a.c * b.s;
    r.x = a.c * b.x - a.s * b.y + a.x;
    r.y = a.s * b.x + a.c * b.y + a.y;
    return r;
}
}  // namespace

#define TRAJ_THREADS 256

// One block.  Thread t owns the contiguous chunk [t * per, (t + 1) * per): chunk products, a shared-memory scan of
// the 256 chunk totals, then the chunk is replayed from its exclusive prefix.  left = 1: T_k = A_k ... A_1 T_0.
__global__ void __launch_bounds__(TRAJ_THREADS)
k_chain_poses(const double* __restrict__ R, const double* __restrict__ h, int P, Se2 T0, int left, double* __restrict__ poses) {
    __shared__ Se2 tot[TRAJ_THREADS];
    const int t = threadIdx.x, per = (P + TRAJ_THREADS - 1) / TRAJ_THREADS;
    const int lo = min(t * per, P), hi = min(lo + per, P);
    const Se2 id = {1.0, 0.0, 0.0, 0.0};
    auto elem = [&](int k) { Se2 a; a.c = R[4 * k]; a.s = R[4 * k + 2]; a.x = h[2 * k]; a.y = h[2 * k + 1]; return a; };
    Se2 acc = id;
    for (int k = lo; k < hi; ++k) acc = left ? mul(elem(k), acc) : mul(acc, elem(k));
    tot[t] = acc;
    __syncthreads();
    for (int off = 1; off < TRAJ_THREADS; off <<= 1) {   // inclusive Hillis-Steele scan in chunk order
        Se2 v = tot[t];
        if (t >= off) v = left ? mul(tot[t], tot[t - off]) : mul(tot[t - off], tot[t]);
        __syncthreads();
        tot[t] = v;
        __syncthreads();
    }
    Se2 cur = t ? tot[t - 1] : id;                        // product of all earlier chunks
    cur = left ? mul(cur, T0) : mul(T0, cur);
    if (t == 0) { poses[0] = T0.x; poses[1] = T0.y; poses[2] = atan2(T0.s, T0.c); }
    for (int k = lo; k < hi; ++k) {
        cur = left ? mul(elem(k), cur) : mul(cur, elem(k));
        poses[3 * (k + 1)] = cur.x; poses[3 * (k + 1) + 1] = cur.y; poses[3 * (k + 1) + 2] = atan2(cur.s, cur.c);
    }
}

// R [P,4] row-major 2x2, h [P,2] (host) -> poses [P+1,3] (x, y, theta) starting from start_pose (x, y, theta; NULL = origin)
extern "C" int rf_chain_poses(rf_handle* h, const double* R, const double* hv, int P, const double* start_pose, int left_multiply,
                              double* poses_out) {
    if (!h || P < 0 || !poses_out || (P > 0 && (!R || !hv))) return rf_fail(h, RF_E_BADARG, "rf_chain_poses: bad argument");
    Se2 T0 = {1.0, 0.0, 0.0, 0.0};
    if (start_pose) { T0.c = cos(start_pose[2]); T0.s = sin(start_pose[2]); T0.x = start_pose[0]; T0.y = start_pose[1]; }
    const size_t bR = ((size_t)P * 32 + 255) & ~(size_t)255, bh = ((size_t)P * 16 + 255) & ~(size_t)255;
    int rc = rf_ensure_scratch(h, bR + bh + (size_t)(P + 1) * 24 + 256);
    if (rc) return rc;
    char* ws = (char*)h->d_scratch;
    double *dR = (double*)ws, *dh = (double*)(ws + bR), *dp = (double*)(ws + bR + bh);
    if (P) {
        RF_CUDA(h, cudaMemcpyAsync(dR, R, (size_t)P * 32, cudaMemcpyHostToDevice, h->stream));
        RF_CUDA(h, cudaMemcpyAsync(dh, hv, (size_t)P * 16, cudaMemcpyHostToDevice, h->stream));
    }
    k_chain_poses<<<1, TRAJ_THREADS, 0, h->stream>>>(dR, dh, P, T0, left_multiply ? 1 : 0, dp);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(poses_out, dp, (size_t)(P + 1) * 24, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}
