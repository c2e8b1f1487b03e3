// k_seq.cu — chained radar odometry on the device: n_seq independent sequences advance one frame per
// step in lock step, with everything the reference's loop carries between frames resident in HBM.
//
// Replaces (reference file:line), for every sequence of the runner:
//   RawROAMSystem.py:141-160   first frame: Cartesian image, appendNewFeatures, first Keyframe     -> rf_seq_reset_async
//   RawROAMSystem.py:162-171   load + convert the next scan, Tracker.track                         -> image kernels, k_klt, k_clique
//   RawROAMSystem.py:185       old_kf.pruneFeaturePoints(corrStatus)          (Mapping.py:118-125)  -> k_seq_update
//   RawROAMSystem.py:190-214   getTransform, p_w, T_wj, MDS.update_problem / optimize_library       -> k_kabsch, k_mds (chain mode)
//   RawROAMSystem.py:237-271   trajectory update, possible_kf.updateInfo, keyframe criteria
//                              (Mapping.py:37-66,149-174), retrack -> appendNewFeatures             -> k_seq_update, detect, k_seq_append
//   RawROAMSystem.py:296-298   blobCoord / prevImgCart / prev_pose carry-over                       -> device state, pyramid ping-pong
//   getFeatures.py:98-118      appendNewFeatures (vstack + order-preserving np.unique)              -> k_seq_append
//
// Device state per sequence: blob [Kmax][2] f32 + count; prev_pose; keyframe pose; undistorted
// keyframe-local points aligned with the blob rows (the reference keeps len(prunedUndistortedLocals) ==
// len(blobCoord) as an invariant); velocity; keyframe count; re-detect flag.
// A step launches a fixed sequence of kernels whose grids do not depend on data (per-sequence work is
// predicated on device-side counts / flags), so it is captured once as a CUDA graph per pyramid parity.
#include <algorithm>
#include <map>

#include "detect.cuh"

#define SEQ_DESC_RING 64
#define SEQ_HIST 8

struct SeqDesc { int32_t base, stride, ring_slot, pad; };

struct rf_seq {
    int S, Kmax, arena_frames, detector_mode;
    int raw_pitch, raw_cols;
    size_t frame_stride;            // bytes between arena frames
    uint8_t* d_arena;
    uint32_t* d_rawi;
    FrameSet fs[2]; int cur;        // fs[cur]: pyramid of the previous frame of every sequence
    float* d_cart;                  // [S][n*n] f32 Cartesian image of the current frame (detector input)
    void* d_det; DetectWs det;
    int32_t* d_pair_idx;
    float* d_feats; int32_t* d_counts;
    double *d_kf_und, *d_kf_pose, *d_prev_pose, *d_vel; int32_t* d_nkf;
    int32_t* d_flags;
    float* d_next; uint8_t* d_status; float* d_err;
    float *d_good_old, *d_good_new; int32_t *d_good_src, *d_ngood;
    void* d_clique_ws;
    double *d_R, *d_h, *d_x; int32_t* d_iters; double* d_mds_scratch;
    rf_seq_result* d_results; int ring;
    SeqDesc* d_desc; SeqDesc* h_desc; cudaEvent_t ev_desc[SEQ_DESC_RING]; int desc_head;
    int steps;                      // steps queued since the reset (0 = only the reset)
    cudaStream_t stream;
    cudaEvent_t ev_uploaded, ev_join;
    // recent steps (arena ranges still being read) so that uploads can wait for exactly those
    struct { int lo, hi; cudaEvent_t ev; bool used; } hist[SEQ_HIST]; int hist_head;
    std::map<int, cudaGraphExec_t> graphs;   // key = parity | with_mds << 1
    int launches_per_step;
};

struct SeqScope {   // launchers take their stream from the handle: run a scope on the runner's stream
    rf_handle* h; cudaStream_t saved;
    SeqScope(rf_handle* h_, cudaStream_t s) : h(h_), saved(h_->stream) { h->stream = s; }
    ~SeqScope() { h->stream = saved; }
};

// ------------------------------------------------------------------------------------
// f32 Cartesian image of the current scan of every FLAGGED sequence (cv2.warpPolar, bit-exact: the same
// arithmetic as k_polar2cart in k_image.cu).  Only re-detecting sequences pay for it.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_seq_cart_f32(const uint8_t* __restrict__ arena, size_t frame_stride, const SeqDesc* __restrict__ desc, int row_pitch, int A,
               int W, const uint32_t* __restrict__ map, int n, float* __restrict__ cart, size_t cart_stride,
               const int32_t* __restrict__ flags) {
    const int s_ = blockIdx.y;     // grid = (tile workers, sequences): un-flagged sequences cost gridDim.x empty CTAs, not one per tile
    if (!flags[s_]) return;
    __shared__ float lut[256];
    const int tid = threadIdx.y * 16 + threadIdx.x;
    lut[tid] = __fdiv_rn((float)tid, 255.0f);  // parseData.py:43  u8 -> f32 / 255. (IEEE division)
    __syncthreads();
    const int n4 = n >> 2;
    const int tiles_x = (n4 + 15) / 16, ntiles = tiles_x * ((n + 15) / 16);
    const uint8_t* s = arena + (size_t)(desc->base + s_ * desc->stride) * frame_stride;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int x4 = (t % tiles_x) * 16 + threadIdx.x;
    const int y = (t / tiles_x) * 16 + threadIdx.y;
    if (x4 >= n4 || y >= n) continue;
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
        int r0 = iy - 1; r0 = r0 < 0 ? r0 + A : (r0 >= A ? r0 - A : r0);
        int r1 = iy;     r1 = r1 >= A ? r1 - A : r1;
        const bool y0ok = iy < A + 2, y1ok = iy + 1 < A + 2;
        const bool x0ok = ix < W, x1ok = ix + 1 < W;
        const uint8_t* p0 = s + (size_t)r0 * row_pitch + ix;
        const uint8_t* p1 = s + (size_t)r1 * row_pitch + ix;
        const float v00 = (y0ok && x0ok) ? lut[__ldg(p0)] : 0.0f;
        const float v01 = (y0ok && x1ok) ? lut[__ldg(p0 + 1)] : 0.0f;
        const float v10 = (y1ok && x0ok) ? lut[__ldg(p1)] : 0.0f;
        const float v11 = (y1ok && x1ok) ? lut[__ldg(p1 + 1)] : 0.0f;
        float acc = __fmul_rn(v00, w00);
        acc = __fadd_rn(acc, __fmul_rn(v01, w01));
        acc = __fadd_rn(acc, __fmul_rn(v10, w10));
        acc = __fadd_rn(acc, __fmul_rn(v11, w11));
        o[k] = acc;
    }
    reinterpret_cast<float4*>(cart + (size_t)s_ * cart_stride + (size_t)y * n)[x4] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------
__global__ void k_seq_init(int S, const double* __restrict__ init_pose, double* __restrict__ prev_pose,
                           double* __restrict__ kf_pose, double* __restrict__ vel, int32_t* __restrict__ counts,
                           int32_t* __restrict__ flags, int32_t* __restrict__ nkf, int32_t* __restrict__ pair_idx,
                           rf_seq_result* __restrict__ res0) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int k = 0; k < 3; ++k) {
        const double v = init_pose ? init_pose[3 * s + k] : 0.0;
        prev_pose[3 * s + k] = v; kf_pose[3 * s + k] = v; vel[3 * s + k] = 0.0;
    }
    counts[s] = 0; flags[s] = 1; nkf[s] = 1;      // the first keyframe (RawROAMSystem.py:155-157)
    pair_idx[2 * s] = s; pair_idx[2 * s + 1] = s;
    rf_seq_result r;
    memset(&r, 0, sizeof(r));
    for (int k = 0; k < 3; ++k) r.pose[k] = prev_pose[3 * s + k];
    r.R[0] = r.R[3] = 1.0; r.kab_R[0] = r.kab_R[3] = 1.0;
    r.retrack = 1; r.keyframe_added = 1; r.n_keyframes = 1;
    res0[s] = r;
}

// MotionDistortionSolver.undistort (motionDistortion.py:127-153) of one metric point
__device__ __forceinline__ void undistort_pt(double vx, double vy, double vth, double period, double x, double y, double& ox,
                                             double& oy) {
    const double t = period * atan2(-y, -x) / (2.0 * M_PI);
    const double th = vth * t, c = cos(th), s = sin(th);
    ox = c * x - s * y + vx * t;
    oy = s * x + c * y + vy * t;
}

struct SeqUpdArgs {
    int S, Kmax, with_mds, retrack_thr;
    double center, res, period, rot_thr, trans_thr_sq;
    const SeqDesc* desc;
    const int32_t* n_good; const float* good_new; const int32_t* good_src;
    const uint8_t* cmask; int mask_stride; const int32_t* n_inl; const int32_t* nodes; const int32_t* cstatus;
    const double* kab_R; const double* kab_h; const double* x; const int32_t* iters;
    float* feats; int32_t* counts;
    double *prev_pose, *kf_pose, *kf_und, *vel; int32_t* nkf; int32_t* flags;
    rf_seq_result* results;    // ring base
};

// One warp per sequence: pose, relative transform, keyframe decision, feature carry-over and keyframe-point pruning.
__global__ void __launch_bounds__(128) k_seq_update(const SeqUpdArgs a) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const size_t base = (size_t)s * a.Kmax;
    const int K = a.counts[s], ng = a.n_good[s], ninl = a.n_inl[s];
    const double x0 = a.prev_pose[3 * s], y0 = a.prev_pose[3 * s + 1], t0 = a.prev_pose[3 * s + 2];
    const double c0 = cos(t0), s0 = sin(t0);
    const double* R = a.kab_R + (size_t)s * 4; const double* hp = a.kab_h + (size_t)s * 2;
    const double hx = hp[0] * a.res, hy = hp[1] * a.res;                     // Tracker.py:125-126
    double pose[3], v[3], mx[6];
    if (a.with_mds) {
        for (int k = 0; k < 6; ++k) mx[k] = a.x[(size_t)s * 6 + k];
        pose[0] = mx[3]; pose[1] = mx[4]; pose[2] = mx[5];                  // RawROAMSystem.py:212
        v[0] = mx[0]; v[1] = mx[1]; v[2] = mx[2];                            // :232
    } else {
        // T_wj = prev_pose @ [[R, h], [0, 1]]   (RawROAMSystem.py:201): the pose without motion compensation
        pose[0] = c0 * hx - s0 * hy + x0; pose[1] = s0 * hx + c0 * hy + y0;
        pose[2] = atan2(s0 * R[0] + c0 * R[2], c0 * R[0] - s0 * R[2]);
        v[0] = v[1] = v[2] = 0.0;
        mx[0] = mx[1] = mx[2] = 0.0; mx[3] = pose[0]; mx[4] = pose[1]; mx[5] = pose[2];
    }
    // relative_transform = T_wj0^-1 @ T(pose)   (RawROAMSystem.py:214)
    const double c = cos(pose[2]), sn = sin(pose[2]);
    const double tx = pose[0] - x0, ty = pose[1] - y0;
    const double r00 = c0 * c + s0 * sn, r01 = -c0 * sn + s0 * c, r10 = -s0 * c + c0 * sn, r11 = s0 * sn + c0 * c;
    const double rhx = c0 * tx + s0 * ty, rhy = -s0 * tx + c0 * ty;
    // keyframe criteria (Mapping.py:149-174) against the latest keyframe, and the retrack rule
    const double dth = a.kf_pose[3 * s + 2] - pose[2];
    const double ddx = a.kf_pose[3 * s] - pose[0], ddy = a.kf_pose[3 * s + 1] - pose[1];
    const bool retrack = ninl <= a.retrack_thr;                               // RawROAMSystem.py:249-250
    const bool good_kf = fabs(dth) >= a.rot_thr || (ddx * ddx + ddy * ddy) >= a.trans_thr_sq;
    const bool new_kf = retrack || good_kf;
    // blobCoord = good_new (clique inliers, order preserved); keyframe points pruned with corrStatus, or re-derived
    // from the new features when this frame becomes the keyframe (possible_kf.updateInfo, Mapping.py:37-66)
    const float* gn = a.good_new + base * 2;
    const int32_t* gs = a.good_src + base;
    const uint8_t* cm = a.cmask + (size_t)s * a.mask_stride;
    float* feats = a.feats + base * 2;
    double* und = a.kf_und + base * 2;
    int cnt = 0;
    for (int j0 = 0; j0 < ng; j0 += 32) {
        const int j = j0 + lane;
        const bool ok = j < ng && cm[j];
        const unsigned bm = __ballot_sync(0xffffffffu, ok);
        float fx = 0.f, fy = 0.f; double ux = 0, uy = 0;
        if (ok) {
            fx = gn[2 * j]; fy = gn[2 * j + 1];
            if (new_kf) {
                const double px = ((double)fx - a.center) * a.res, py = ((double)fy - a.center) * a.res;   // RawROAMSystem.py:198-199
                undistort_pt(v[0], v[1], v[2], a.period, px, py, ux, uy);
            } else {
                ux = und[2 * gs[j]]; uy = und[2 * gs[j] + 1];
            }
        }
        __syncwarp();      // every read of this chunk precedes its writes (dst <= src row: no later read is clobbered)
        if (ok) {
            const int d = cnt + __popc(bm & ((1u << lane) - 1u));
            feats[2 * d] = fx; feats[2 * d + 1] = fy;
            und[2 * d] = ux; und[2 * d + 1] = uy;
        }
        cnt += __popc(bm);
        __syncwarp();
    }
    __syncwarp();          // every lane has read the state lane 0 is about to overwrite
    if (lane == 0) {
        int nk = a.nkf[s];
        if (new_kf) {
            ++nk;
            a.nkf[s] = nk;
            for (int k = 0; k < 3; ++k) a.kf_pose[3 * s + k] = pose[k];
        }
        for (int k = 0; k < 3; ++k) { a.prev_pose[3 * s + k] = pose[k]; a.vel[3 * s + k] = v[k]; }
        a.counts[s] = cnt;
        a.flags[s] = retrack ? 1 : 0;
        rf_seq_result r;
        for (int k = 0; k < 3; ++k) r.pose[k] = pose[k];
        r.R[0] = r00; r.R[1] = r01; r.R[2] = r10; r.R[3] = r11; r.h[0] = rhx; r.h[1] = rhy;
        for (int k = 0; k < 6; ++k) r.mds_x[k] = mx[k];
        for (int k = 0; k < 4; ++k) r.kab_R[k] = R[k];
        r.kab_h[0] = hx; r.kab_h[1] = hy;
        r.n_features_in = K; r.n_good = ng; r.n_tracked = ninl; r.retrack = retrack; r.keyframe_added = new_kf;
        r.n_keyframes = nk; r.n_features_out = cnt; r.n_candidates = 0;
        r.mds_iters = a.with_mds ? a.iters[s] : 0; r.clique_nodes = a.nodes[s]; r.status = a.cstatus[s]; r.reserved = 0;
        a.results[(size_t)a.desc->ring_slot * a.S + s] = r;
    }
}

// getFeatures.appendNewFeatures (getFeatures.py:98-118) for every flagged sequence: vstack(old, new), np.unique(axis=0)
// keeping first occurrences in order, f32; then old_kf.updateInfo(latestPose, centered_new, ..., velocity)
// (RawROAMSystem.py:264-268): the keyframe points are re-derived for the whole new feature set.
struct SeqAppendArgs {
    int S, Kmax; double center, res, period;
    const SeqDesc* desc; const int32_t* flags;
    const int32_t* sel_idx; const double2* rc; unsigned ssc_cap; const int32_t* m; const int32_t* dstatus; const unsigned* ncand;
    float* feats; int32_t* counts; double* kf_und; const double* vel;
    rf_seq_result* results;
};

__global__ void __launch_bounds__(256) k_seq_append(const SeqAppendArgs a) {
    extern __shared__ float2 s_pts[];            // [Kmax] combined rows, then uint8 dup[Kmax]
    uint8_t* s_dup = reinterpret_cast<uint8_t*>(s_pts + a.Kmax);
    __shared__ int s_total;
    const int s = blockIdx.x;
    if (!a.flags[s]) return;
    const int tid = threadIdx.x;
    const size_t base = (size_t)s * a.Kmax;
    const int n_old = a.counts[s];
    int m = a.m[s];
    int status = a.dstatus[s];
    if (n_old + m > a.Kmax) { m = a.Kmax - n_old; status = RF_E_CAPACITY; }
    const int tot = n_old + m;
    const int32_t* sel = a.sel_idx + (size_t)s * a.ssc_cap;
    const double2* rc = a.rc + (size_t)s * a.ssc_cap;
    for (int i = tid; i < tot; i += 256) {
        if (i < n_old) s_pts[i] = make_float2(a.feats[(base + i) * 2], a.feats[(base + i) * 2 + 1]);
        else { const double2 q = rc[sel[i - n_old]]; s_pts[i] = make_float2((float)q.y, (float)q.x); }   // np.fliplr: (row, col) -> (x, y)
    }
    __syncthreads();
    for (int i = tid; i < tot; i += 256) {
        const float2 p = s_pts[i];
        bool dup = false;
        for (int j = 0; j < i && !dup; ++j) dup = s_pts[j].x == p.x && s_pts[j].y == p.y;
        s_dup[i] = dup;
    }
    __syncthreads();
    if (tid < 32) {
        int cnt = 0;
        for (int i0 = 0; i0 < tot; i0 += 32) {
            const int i = i0 + tid;
            const bool ok = i < tot && !s_dup[i];
            const unsigned bm = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int d = cnt + __popc(bm & ((1u << tid) - 1u));
                const float2 p = s_pts[i];
                a.feats[(base + d) * 2] = p.x; a.feats[(base + d) * 2 + 1] = p.y;
                const double px = ((double)p.x - a.center) * a.res, py = ((double)p.y - a.center) * a.res;
                double ux, uy;
                undistort_pt(a.vel[3 * s], a.vel[3 * s + 1], a.vel[3 * s + 2], a.period, px, py, ux, uy);
                a.kf_und[(base + d) * 2] = ux; a.kf_und[(base + d) * 2 + 1] = uy;
            }
            cnt += __popc(bm);
        }
        if (tid == 0) s_total = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        a.counts[s] = s_total;
        rf_seq_result* r = a.results + (size_t)a.desc->ring_slot * a.S + s;
        r->n_features_out = s_total;
        r->n_candidates = (int32_t)a.ncand[s];
        if (status != RF_OK && r->status == RF_OK) r->status = status;
    }
}

// ------------------------------------------------------------------------------------
static void seq_free(rf_seq* q) {
    if (!q) return;
    auto F = [](void* p) { if (p) cudaFree(p); };
    for (auto& kv : q->graphs) cudaGraphExecDestroy(kv.second);
    F(q->d_arena); F(q->d_rawi); rf_frameset_free(&q->fs[0]); rf_frameset_free(&q->fs[1]); F(q->d_cart); F(q->d_det);
    F(q->d_pair_idx); F(q->d_feats); F(q->d_counts); F(q->d_kf_und); F(q->d_kf_pose); F(q->d_prev_pose); F(q->d_vel);
    F(q->d_nkf); F(q->d_flags); F(q->d_next); F(q->d_status); F(q->d_err); F(q->d_good_old); F(q->d_good_new);
    F(q->d_good_src); F(q->d_ngood); F(q->d_clique_ws); F(q->d_R); F(q->d_h); F(q->d_x); F(q->d_iters); F(q->d_mds_scratch);
    F(q->d_results); F(q->d_desc);
    if (q->h_desc) cudaFreeHost(q->h_desc);
    for (int i = 0; i < SEQ_DESC_RING; ++i) if (q->ev_desc[i]) cudaEventDestroy(q->ev_desc[i]);
    for (int i = 0; i < SEQ_HIST; ++i) if (q->hist[i].ev) cudaEventDestroy(q->hist[i].ev);
    if (q->ev_uploaded) cudaEventDestroy(q->ev_uploaded);
    if (q->ev_join) cudaEventDestroy(q->ev_join);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
}

#define RF_SALLOC(ptr, bytes)                                                                       \
    do {                                                                                            \
        if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) {                                   \
            cudaGetLastError(); seq_free(q);                                                        \
            return rf_fail(h, RF_E_NOMEM, "rf_seq_create: cudaMalloc(%zu) failed", (size_t)(bytes)); \
        }                                                                                           \
    } while (0)

// k_doh.cu: determinant-of-Hessian candidates (keys) for every flagged problem
int rf_launch_doh_keypoints(rf_handle* h, const DetectWs& ws, const float* d_cart, size_t cart_stride, int n, void* d_doh_ws,
                             const int32_t* d_flags);
size_t rf_doh_ws_bytes(const rf_handle* h, int S);
int rf_doh_prepare(rf_handle* h);

static std::map<rf_seq*, void*> g_doh_ws;   // DoH scratch (integral images) of runners created with detector_mode 1

extern "C" {

int rf_seq_create(rf_handle* h, int n_seq, int arena_frames, int detector_mode, rf_seq** out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !out || n_seq < 1 || arena_frames < n_seq) return rf_fail(h, RF_E_BADARG, "rf_seq_create: bad argument (arena_frames must be >= n_seq)");
    if (detector_mode != 0 && detector_mode != 1) return rf_fail(h, RF_E_BADARG, "rf_seq_create: detector_mode %d (0 = structure tensor, 1 = determinant of Hessian)", detector_mode);
    *out = nullptr;
    cudaSetDevice(h->device);
    const rf_config& c = h->cfg;
    if (c.klt_max_level < 1) return rf_fail(h, RF_E_BADARG, "rf_seq_create: the fused image path needs at least two pyramid levels");
    rf_seq* q = new rf_seq();
    q->S = n_seq; q->Kmax = c.max_features; q->arena_frames = arena_frames; q->detector_mode = detector_mode;
    q->raw_cols = c.range_bins; q->raw_pitch = (q->raw_cols + 15) & ~15;
    q->frame_stride = (size_t)c.azimuths * q->raw_pitch;
    q->d_arena = nullptr; q->d_rawi = nullptr; memset(&q->fs[0], 0, sizeof(FrameSet)); memset(&q->fs[1], 0, sizeof(FrameSet));
    q->cur = 0; q->d_cart = nullptr; q->d_det = nullptr; q->d_pair_idx = nullptr; q->d_feats = nullptr; q->d_counts = nullptr;
    q->d_kf_und = q->d_kf_pose = q->d_prev_pose = q->d_vel = nullptr; q->d_nkf = nullptr; q->d_flags = nullptr;
    q->d_next = nullptr; q->d_status = nullptr; q->d_err = nullptr; q->d_good_old = q->d_good_new = nullptr;
    q->d_good_src = q->d_ngood = nullptr; q->d_clique_ws = nullptr; q->d_R = q->d_h = q->d_x = nullptr; q->d_iters = nullptr;
    q->d_mds_scratch = nullptr; q->d_results = nullptr; q->d_desc = nullptr; q->h_desc = nullptr; q->desc_head = 0; q->steps = 0;
    q->stream = nullptr; q->ev_uploaded = nullptr; q->ev_join = nullptr; q->hist_head = 0; q->launches_per_step = 0;
    for (int i = 0; i < SEQ_DESC_RING; ++i) q->ev_desc[i] = nullptr;
    for (int i = 0; i < SEQ_HIST; ++i) { q->hist[i].ev = nullptr; q->hist[i].used = false; }
    q->ring = std::max(8, std::min(arena_frames, 4096));
    const size_t S = n_seq, K = q->Kmax, n2 = (size_t)h->n * h->n;
    RF_SALLOC(q->d_arena, (size_t)arena_frames * q->frame_stride + 16);
    RF_SALLOC(q->d_rawi, rf_interleave_words(h, n_seq) * sizeof(uint32_t));
    int rc;
    if ((rc = rf_frameset_alloc(h, &q->fs[0], n_seq, false)) || (rc = rf_frameset_alloc(h, &q->fs[1], n_seq, false))) { seq_free(q); return rc; }
    RF_SALLOC(q->d_cart, S * n2 * sizeof(float));
    const unsigned key_cap = 1u << 17, ssc_cap = RF_SSC_MAX_CANDIDATES, cells_cap = 1u << 18;
    const size_t det_bytes = rf_detect_ws_bytes(n_seq, h->n, h->n, key_cap, ssc_cap, cells_cap, detector_mode == 0);
    RF_SALLOC(q->d_det, det_bytes);
    q->det = rf_detect_ws_carve(q->d_det, n_seq, h->n, h->n, key_cap, ssc_cap, cells_cap, detector_mode == 0);
    RF_SALLOC(q->d_pair_idx, S * 2 * sizeof(int32_t));
    RF_SALLOC(q->d_feats, S * K * 2 * sizeof(float));
    RF_SALLOC(q->d_counts, S * sizeof(int32_t));
    RF_SALLOC(q->d_kf_und, S * K * 2 * sizeof(double));
    RF_SALLOC(q->d_kf_pose, S * 3 * sizeof(double));
    RF_SALLOC(q->d_prev_pose, S * 3 * sizeof(double));
    RF_SALLOC(q->d_vel, S * 3 * sizeof(double));
    RF_SALLOC(q->d_nkf, S * sizeof(int32_t));
    RF_SALLOC(q->d_flags, S * sizeof(int32_t));
    RF_SALLOC(q->d_next, S * K * 2 * sizeof(float));
    RF_SALLOC(q->d_status, S * K);
    RF_SALLOC(q->d_err, S * K * sizeof(float));
    RF_SALLOC(q->d_good_old, S * K * 2 * sizeof(float));
    RF_SALLOC(q->d_good_new, S * K * 2 * sizeof(float));
    RF_SALLOC(q->d_good_src, S * K * sizeof(int32_t));
    RF_SALLOC(q->d_ngood, S * sizeof(int32_t));
    RF_SALLOC(q->d_clique_ws, rf_clique_ws_total((int)K, n_seq));
    RF_SALLOC(q->d_R, S * 4 * sizeof(double));
    RF_SALLOC(q->d_h, S * 2 * sizeof(double));
    RF_SALLOC(q->d_x, S * 6 * sizeof(double));
    RF_SALLOC(q->d_iters, S * sizeof(int32_t));
    RF_SALLOC(q->d_mds_scratch, S * K * 5 * sizeof(double));
    RF_SALLOC(q->d_results, (size_t)q->ring * S * sizeof(rf_seq_result));
    RF_SALLOC(q->d_desc, sizeof(SeqDesc));
    if (cudaMallocHost((void**)&q->h_desc, SEQ_DESC_RING * sizeof(SeqDesc)) != cudaSuccess) { seq_free(q); return rf_fail(h, RF_E_NOMEM, "rf_seq_create: pinned allocation failed"); }
    bool ok = cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&q->ev_uploaded, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&q->ev_join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < SEQ_DESC_RING; ++i) ok = cudaEventCreateWithFlags(&q->ev_desc[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < SEQ_HIST; ++i) ok = cudaEventCreateWithFlags(&q->hist[i].ev, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { seq_free(q); return rf_fail(h, RF_E_CUDA, "rf_seq_create: stream/event creation failed"); }
    if (detector_mode == 1) {
        void* dw = nullptr;
        if (cudaMalloc(&dw, rf_doh_ws_bytes(h, n_seq)) != cudaSuccess) { cudaGetLastError(); seq_free(q); return rf_fail(h, RF_E_NOMEM, "rf_seq_create: DoH workspace allocation failed"); }
        g_doh_ws[q] = dw;
    }
    {
        SeqScope sc(h, q->stream);
        if ((rc = rf_detect_ws_init(h, q->det)) || (rc = rf_detect_prepare(h)) || (detector_mode == 1 && (rc = rf_doh_prepare(h)))) { seq_free(q); return rc; }
        cudaMemsetAsync(q->d_next, 0, S * K * 2 * sizeof(float), q->stream);
        cudaMemsetAsync(q->d_status, 0, S * K, q->stream);
        cudaMemsetAsync(q->d_err, 0, S * K * sizeof(float), q->stream);
        cudaMemsetAsync(q->d_feats, 0, S * K * 2 * sizeof(float), q->stream);
        cudaMemsetAsync(q->d_kf_und, 0, S * K * 2 * sizeof(double), q->stream);
        cudaMemsetAsync(q->d_results, 0, (size_t)q->ring * S * sizeof(rf_seq_result), q->stream);
        cudaMemsetAsync(q->d_counts, 0, S * sizeof(int32_t), q->stream);
        if (cudaStreamSynchronize(q->stream) != cudaSuccess) { seq_free(q); return rf_fail(h, RF_E_CUDA, "rf_seq_create: initialisation failed"); }
    }
    h->seqs.push_back(q);
    *out = q;
    return RF_OK;
}

void rf_seq_destroy(rf_handle* h, rf_seq* q) {
    RfDeviceGuard rf_guard_(h);
    if (!q) return;
    if (h) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(q->stream);
        if (h->stream_copy) cudaStreamSynchronize(h->stream_copy);
        h->seqs.erase(std::remove(h->seqs.begin(), h->seqs.end(), q), h->seqs.end());
    }
    auto it = g_doh_ws.find(q);
    if (it != g_doh_ws.end()) { cudaFree(it->second); g_doh_ws.erase(it); }
    seq_free(q);
}

int rf_seq_upload_async(rf_handle* h, rf_seq* q, int first_frame, int n_frames, const uint8_t* raw) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q || !raw || first_frame < 0 || n_frames < 0 || first_frame + n_frames > q->arena_frames)
        return rf_fail(h, RF_E_BADARG, "rf_seq_upload: frames [%d, %d) outside the arena of %d", first_frame, first_frame + n_frames, q ? q->arena_frames : 0);
    if (!n_frames) return RF_OK;
    cudaSetDevice(h->device);
    const rf_config& c = h->cfg;
    cudaStream_t cs = h->stream_copy;
    const int lo = first_frame, hi = first_frame + n_frames - 1;
    for (int i = 0; i < SEQ_HIST; ++i)      // queued steps that still read these frames
        if (q->hist[i].used && !(q->hist[i].hi < lo || q->hist[i].lo > hi)) RF_CUDA(h, cudaStreamWaitEvent(cs, q->hist[i].ev, 0));
    RF_CUDA(h, cudaMemcpy2DAsync(q->d_arena + (size_t)first_frame * q->frame_stride, q->raw_pitch, raw + c.meta_bytes, c.raw_width,
                                 q->raw_cols, (size_t)n_frames * c.azimuths, cudaMemcpyHostToDevice, cs));
    RF_CUDA(h, cudaEventRecord(q->ev_uploaded, cs));
    return RF_OK;
}

// descriptor of the next enqueue: pinned ring -> device (stream-ordered, outside the graph)
static int seq_push_desc(rf_handle* h, rf_seq* q, int base, int stride, int ring_slot) {
    const int k = q->desc_head;
    RF_CUDA(h, cudaEventSynchronize(q->ev_desc[k]));          // the copy that last used this pinned slot has run
    q->h_desc[k].base = base; q->h_desc[k].stride = stride; q->h_desc[k].ring_slot = ring_slot; q->h_desc[k].pad = 0;
    RF_CUDA(h, cudaMemcpyAsync(q->d_desc, &q->h_desc[k], sizeof(SeqDesc), cudaMemcpyHostToDevice, q->stream));
    RF_CUDA(h, cudaEventRecord(q->ev_desc[k], q->stream));
    q->desc_head = (k + 1) % SEQ_DESC_RING;
    return RF_OK;
}

static int seq_check_range(rf_handle* h, rf_seq* q, int base, int stride) {
    const long long last = (long long)base + (long long)(q->S - 1) * stride;
    if (base < 0 || base >= q->arena_frames || last < 0 || last >= q->arena_frames)
        return rf_fail(h, RF_E_BADARG, "rf_seq: scans base %d + s * %d (s < %d) leave the arena of %d frames", base, stride, q->S, q->arena_frames);
    return RF_OK;
}

static void seq_note_step(rf_handle* h, rf_seq* q, int base, int stride) {
    const int a = base, b = base + (q->S - 1) * stride;
    auto& e = q->hist[q->hist_head];
    e.lo = std::min(a, b); e.hi = std::max(a, b); e.used = true;
    cudaEventRecord(e.ev, q->stream);
    q->hist_head = (q->hist_head + 1) % SEQ_HIST;
}

// image path of the current scans into pyramid set `dst`
static int seq_enqueue_image(rf_handle* h, rf_seq* q, int dst) {
    int rc;
    const int32_t* sel = reinterpret_cast<const int32_t*>(q->d_desc);
    if ((rc = rf_launch_interleave(h, q->d_arena, q->frame_stride, q->raw_pitch, q->S, q->d_rawi, sel))) return rc;
    if ((rc = rf_launch_scan_to_l0l1(h, q->d_rawi, q->fs[dst], q->S))) return rc;
    return rf_launch_pyr_levels(h, q->fs[dst], 2, q->S);
}

// re-detection + append for the flagged sequences (getFeatures.appendNewFeatures on the CURRENT scan)
static int seq_enqueue_detect(rf_handle* h, rf_seq* q) {
    const rf_config& c = h->cfg;
    int rc;
    const size_t n2 = (size_t)h->n * h->n;
    dim3 blk(16, 16), grd(rf_tile_workers(h, (((h->n >> 2) + 15) / 16) * ((h->n + 15) / 16), q->S), q->S);
    k_seq_cart_f32<<<grd, blk, 0, h->stream>>>(q->d_arena, q->frame_stride, q->d_desc, q->raw_pitch, c.azimuths, c.range_bins, h->map,
                                               h->n, q->d_cart, n2, q->d_flags);
    RF_CHECK_LAUNCH(h);
    if (q->detector_mode == 0) {
        if ((rc = rf_launch_detect_clear(h, q->det))) return rc;
        if ((rc = rf_launch_min_eig(h, q->d_cart, n2, h->n, q->det.resp, q->det.resp_stride, q->S, q->d_flags, q->det.maxbits))) return rc;
        if ((rc = rf_launch_select_sorted(h, q->det, q->det.resp, q->det.resp_stride, (float)-c.detect_quality, q->d_flags, true))) return rc;
        if ((rc = rf_launch_ssc_from_keys(h, q->det, c.ssc_num_ret, c.ssc_tolerance, q->d_flags))) return rc;
    } else {
        // the reference's detector: blob_doh -> adaptiveNMS' argsort by sigma -> ssc (getFeatures.py:47-51,66-72)
        if ((rc = rf_launch_doh_keypoints(h, q->det, q->d_cart, n2, h->n, g_doh_ws[q], q->d_flags))) return rc;
        if ((rc = rf_launch_ssc(h, q->det, c.ssc_num_ret, c.ssc_tolerance, h->n, h->n, q->d_flags))) return rc;
    }
    SeqAppendArgs a;
    a.S = q->S; a.Kmax = q->Kmax; a.center = (double)h->R; a.res = c.cart_res_m; a.period = c.mds_period;
    a.desc = q->d_desc; a.flags = q->d_flags; a.sel_idx = q->det.sel_idx; a.rc = q->det.rc; a.ssc_cap = q->det.ssc_cap;
    a.m = q->det.m; a.dstatus = q->det.status; a.ncand = q->det.count; a.feats = q->d_feats; a.counts = q->d_counts;
    a.kf_und = q->d_kf_und; a.vel = q->d_vel; a.results = q->d_results;
    k_seq_append<<<q->S, 256, (size_t)q->Kmax * 9, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// everything of one step after the descriptor is in place (graph-capturable: static grids, device-side predicates)
static int seq_enqueue_step(rf_handle* h, rf_seq* q, int parity, int with_mds) {
    const rf_config& c = h->cfg;
    const int S = q->S, K = q->Kmax, prev = parity, nxt = parity ^ 1;
    int rc;
    if ((rc = seq_enqueue_image(h, q, nxt))) return rc;
    if ((rc = rf_launch_klt(h, q->fs[prev], q->fs[nxt], q->d_pair_idx, q->d_feats, q->d_counts, K, S, q->d_next, q->d_status, q->d_err, 1))) return rc;
    if ((rc = rf_launch_compact_good(h, q->d_feats, q->d_next, q->d_status, q->d_counts, K, S, q->d_good_old, q->d_good_new, q->d_good_src, q->d_ngood))) return rc;
    uint8_t* d_mask; int mask_stride; int32_t *d_ninl, *d_nodes, *d_cstatus;
    if ((rc = rf_launch_reject(h, q->d_clique_ws, q->d_good_old, q->d_good_new, q->d_ngood, K, S, &d_mask, &mask_stride, &d_ninl, &d_nodes, &d_cstatus))) return rc;
    // Tracker.getTransform(good_old, good_new): src = good_old, target = good_new
    if ((rc = rf_launch_kabsch(h, q->d_good_old, q->d_good_new, d_mask, mask_stride, q->d_ngood, K, S, q->d_R, q->d_h, nullptr))) return rc;
    if (with_mds &&
        (rc = rf_launch_mds_chain(h, q->d_good_old, q->d_good_new, d_mask, mask_stride, q->d_ngood, K, S, q->d_R, q->d_h, q->d_prev_pose,
                                  q->d_kf_und, q->d_kf_pose, q->d_good_src, q->d_mds_scratch, q->d_x, q->d_iters))) return rc;
    SeqUpdArgs u;
    u.S = S; u.Kmax = K; u.with_mds = with_mds; u.retrack_thr = c.retrack_threshold; u.center = (double)h->R; u.res = c.cart_res_m;
    u.period = c.mds_period; u.rot_thr = c.kf_rot_thr; u.trans_thr_sq = c.kf_trans_thr * c.kf_trans_thr; u.desc = q->d_desc;
    u.n_good = q->d_ngood; u.good_new = q->d_good_new; u.good_src = q->d_good_src; u.cmask = d_mask; u.mask_stride = mask_stride;
    u.n_inl = d_ninl; u.nodes = d_nodes; u.cstatus = d_cstatus; u.kab_R = q->d_R; u.kab_h = q->d_h; u.x = q->d_x; u.iters = q->d_iters;
    u.feats = q->d_feats; u.counts = q->d_counts; u.prev_pose = q->d_prev_pose; u.kf_pose = q->d_kf_pose; u.kf_und = q->d_kf_und;
    u.vel = q->d_vel; u.nkf = q->d_nkf; u.flags = q->d_flags; u.results = q->d_results;
    k_seq_update<<<(S + 3) / 4, 128, 0, h->stream>>>(u);
    RF_CHECK_LAUNCH(h);
    return seq_enqueue_detect(h, q);
}

int rf_seq_reset_async(rf_handle* h, rf_seq* q, int base, int stride, const double* init_pose) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q) return rf_fail(h, RF_E_BADARG, "rf_seq_reset: null argument");
    cudaSetDevice(h->device);
    int rc = seq_check_range(h, q, base, stride);
    if (rc) return rc;
    SeqScope sc(h, q->stream);
    RF_CUDA(h, cudaStreamWaitEvent(q->stream, q->ev_uploaded, 0));
    if (init_pose) {
        // staged through the MDS solution buffer (free until the first step); pageable source: the copy call returns
        // once the data has been staged, so the caller's array is not referenced afterwards
        RF_CUDA(h, cudaMemcpyAsync(q->d_mds_scratch, init_pose, (size_t)q->S * 3 * sizeof(double), cudaMemcpyHostToDevice, q->stream));
    }
    q->steps = 0; q->cur = 0;
    if ((rc = seq_push_desc(h, q, base, stride, 0))) return rc;
    k_seq_init<<<(q->S + 127) / 128, 128, 0, h->stream>>>(q->S, init_pose ? q->d_mds_scratch : nullptr, q->d_prev_pose, q->d_kf_pose,
                                                          q->d_vel, q->d_counts, q->d_flags, q->d_nkf, q->d_pair_idx, q->d_results);
    RF_CHECK_LAUNCH(h);
    if ((rc = seq_enqueue_image(h, q, q->cur))) return rc;
    if ((rc = seq_enqueue_detect(h, q))) return rc;
    seq_note_step(h, q, base, stride);
    return RF_OK;
}

int rf_seq_step_async(rf_handle* h, rf_seq* q, int base, int stride, int flags) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q) return rf_fail(h, RF_E_BADARG, "rf_seq_step: null argument");
    cudaSetDevice(h->device);
    int rc = seq_check_range(h, q, base, stride);
    if (rc) return rc;
    const int with_mds = (flags & RF_SEQ_MDS) ? 1 : 0;
    SeqScope sc(h, q->stream);
    RF_CUDA(h, cudaStreamWaitEvent(q->stream, q->ev_uploaded, 0));
    const int step = q->steps + 1;
    if ((rc = seq_push_desc(h, q, base, stride, step % q->ring))) return rc;
    if (flags & RF_SEQ_GRAPH) {
        const int key = q->cur | (with_mds << 1);
        auto it = q->graphs.find(key);
        if (it == q->graphs.end()) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = h->launches;
            RF_CUDA(h, cudaStreamBeginCapture(q->stream, cudaStreamCaptureModeRelaxed));
            rc = seq_enqueue_step(h, q, q->cur, with_mds);
            cudaError_t e = cudaStreamEndCapture(q->stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return rf_fail(h, RF_E_CUDA, "rf_seq_step: graph capture failed: %s", cudaGetErrorString(e));
            q->launches_per_step = (int)(h->launches - l0);
            h->launches = l0;                      // capturing launched nothing
            cudaGraphExec_t ge = nullptr;
            e = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return rf_fail(h, RF_E_CUDA, "rf_seq_step: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            it = q->graphs.emplace(key, ge).first;
        }
        RF_CUDA(h, cudaGraphLaunch(it->second, q->stream));
        h->launches += q->launches_per_step;
    } else {
        const int64_t l0 = h->launches;
        if ((rc = seq_enqueue_step(h, q, q->cur, with_mds))) return rc;
        q->launches_per_step = (int)(h->launches - l0);
    }
    q->cur ^= 1;
    q->steps = step;
    seq_note_step(h, q, base, stride);
    return RF_OK;
}

int rf_seq_results_async(rf_handle* h, rf_seq* q, int step, rf_seq_result* out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q || !out) return rf_fail(h, RF_E_BADARG, "rf_seq_results: null argument");
    if (step < 0 || step > q->steps || step <= q->steps - q->ring)
        return rf_fail(h, RF_E_BADARG, "rf_seq_results: step %d is not among the last %d of %d", step, q->ring, q->steps);
    cudaSetDevice(h->device);
    RF_CUDA(h, cudaMemcpyAsync(out, q->d_results + (size_t)(step % q->ring) * q->S, (size_t)q->S * sizeof(rf_seq_result),
                               cudaMemcpyDeviceToHost, q->stream));
    return RF_OK;
}

int rf_seq_results(rf_handle* h, rf_seq* q, int step, rf_seq_result* out) {
    RfDeviceGuard rf_guard_(h);
    int rc = rf_seq_results_async(h, q, step, out);
    if (rc) return rc;
    RF_CUDA(h, cudaStreamSynchronize(q->stream));
    return RF_OK;
}

int rf_seq_ring(const rf_seq* q) { return q ? q->ring : 0; }
int rf_seq_steps_done(const rf_seq* q) { return q ? q->steps : 0; }
int rf_seq_launches_per_step(const rf_seq* q) { return q ? q->launches_per_step : 0; }

int rf_seq_features(rf_handle* h, rf_seq* q, float* feats, int32_t* counts) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q) return rf_fail(h, RF_E_BADARG, "rf_seq_features: null argument");
    cudaSetDevice(h->device);
    if (feats) RF_CUDA(h, cudaMemcpyAsync(feats, q->d_feats, (size_t)q->S * q->Kmax * 2 * sizeof(float), cudaMemcpyDeviceToHost, q->stream));
    if (counts) RF_CUDA(h, cudaMemcpyAsync(counts, q->d_counts, (size_t)q->S * sizeof(int32_t), cudaMemcpyDeviceToHost, q->stream));
    RF_CUDA(h, cudaStreamSynchronize(q->stream));
    return RF_OK;
}

int rf_seq_sync(rf_handle* h, rf_seq* q) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !q) return rf_fail(h, RF_E_BADARG, "rf_seq_sync: null argument");
    cudaSetDevice(h->device);
    RF_CUDA(h, cudaStreamSynchronize(h->stream_copy));
    RF_CUDA(h, cudaStreamSynchronize(q->stream));
    return RF_OK;
}

}  // extern "C"

// joins for the handle-level timers / rf_sync (k_batch.cu calls these)
int rf_seq_join_all(rf_handle* h) {
    for (rf_seq* q : h->seqs) {
        RF_CUDA(h, cudaEventRecord(q->ev_join, q->stream));
        RF_CUDA(h, cudaStreamWaitEvent(h->stream, q->ev_join, 0));
    }
    return RF_OK;
}
int rf_seq_sync_all(rf_handle* h) {
    for (rf_seq* q : h->seqs) RF_CUDA(h, cudaStreamSynchronize(q->stream));
    return RF_OK;
}
