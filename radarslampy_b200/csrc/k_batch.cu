// k_batch.cu — the fused per-pair front end over a device-resident batch of independent
// frame pairs: what Tracker.track + Tracker.getTransform (+ the MotionDistortionSolver call
// in RawROAMSystem.run) do for one pair, for every pair of the batch, one launch per stage.
//
// Replaces (reference file:line):
//   RawROAMSystem.py:164-165   load + convertPolarImageToCartesian     -> k_polar2cart / k_pyr_down
//   Tracker.py:75-76           getTrackedPointsKLT                     -> k_klt
//   getTransformKLT.py:368-376 good/bad split (order-preserving)       -> k_compact_good
//   Tracker.py:95              rejectOutliers                          -> k_adjacency + k_clique
//   Tracker.py:103-104         corrStatus[good rows] &= clique mask    -> k_finish_pairs
//   Tracker.py:108-127         getTransform (Kabsch, h * 0.0864)       -> k_kabsch (+ k_finish_pairs)
//   RawROAMSystem.py:190-212   T_wj, p_w, MDS.update_problem/optimize  -> k_mds (fused mode)
//
// HBM layout of an rf_batch (all sized once at rf_batch_create from rf_config maxima):
//   raw      [max_frames][A][raw_pitch] u8   only the used columns (11 metadata + range_bins)
//                                            are uploaded (2-D copy), raw_pitch = 16-byte multiple
//   frames   FrameSet(count = max_frames): u8 pyramid levels (+ f32 plane if write_cart_f32)
//   pairs    pair_idx[P][2], feats[P][Kmax][2], counts[P], prev_pose[P][3]
//   tracks   next[P][Kmax][2], status[P][Kmax], err[P][Kmax]
//   good     old/new[P][Kmax][2], src_row[P][Kmax], n_good[P]
//   clique   workspace of k_clique.cu (adjacency bitsets, search stacks, masks)
//   poses    R[P][4], h[P][2], mds_x[P][6], results[P] (rf_pair_result)
//
// Streams.  A batch is pipelined over three streams so that, with two batches alternating on
// one handle, the PCIe upload of batch B, the image + KLT kernels of batch A and the
// latency-bound clique search of the batch before overlap:
//   handle.stream_copy  H2D staging                        -> ev_uploaded
//   handle.stream       interleave, scan->pyramid, KLT, compaction      -> ev_main_done
//   batch.tail          adjacency, clique, Kabsch, MDS, finish, D2H     -> ev_tail_done
// An upload waits for the previous run of the SAME batch (ev_main_done, ev_tail_done) before
// it overwrites the inputs; KLT waits for the previous tail of the same batch before it
// overwrites the track buffers.  rf_sync / the timers join all three.
#include <algorithm>

#include <stdlib.h>

#include "common.cuh"

#define RF_PROFILE_RING 64
#define RF_N_STAGES 9   // p2c|interleave, scan->L0+L1, pyrDown (rest), klt, compact, reject, kabsch, mds, finish

struct rf_batch {
    int max_frames, max_pairs, Kmax;
    int raw_pitch, raw_cols;
    uint8_t* d_raw;
    uint32_t* d_rawi;            // frame-interleaved scans (4 frames per word), input of the fused image kernel
    FrameSet fs;
    int32_t* d_pair_idx; float* d_feats; int32_t* d_counts; double* d_prev_pose;
    float* d_next; uint8_t* d_status; float* d_err;
    float *d_good_old, *d_good_new; int32_t* d_good_src; int32_t* d_ngood;
    void* d_clique_ws;
    double *d_R, *d_h, *d_x; int32_t* d_iters; double* d_mds_scratch;
    uint8_t* d_corr;             // [P][Kmax] corrStatus after the clique fix-up
    rf_pair_result* d_results;
    int n_frames, n_pairs; bool has_pose; bool ran_mds;
    // per-stage CUDA events of the most recent runs (ring)
    bool profiling; int prof_head; int prof_count;
    cudaEvent_t ev[RF_PROFILE_RING][RF_N_STAGES + 1];
    // pipeline
    cudaStream_t tail;
    cudaEvent_t ev_uploaded, ev_main_done, ev_tail_done;
};

// RADARFE_KLT_ON_TAIL=1 moves KLT from the handle's main stream to the batch's tail stream.  Measured on B200 (r01g):
// +2.5 % with >= 4 batches in flight, -4.5 % with 3 — the scan kernel's CTAs fill the shared memory of every SM, so the
// two kernels take turns rather than co-run; the default keeps KLT on the main stream.
static const bool g_klt_on_tail = []() { const char* e = getenv("RADARFE_KLT_ON_TAIL"); return e && e[0] == '1'; }();

// Launchers take their stream from the handle: run a scope on another stream.
struct StreamScope {
    rf_handle* h; cudaStream_t saved;
    StreamScope(rf_handle* h_, cudaStream_t s) : h(h_), saved(h_->stream) { h->stream = s; }
    ~StreamScope() { h->stream = saved; }
};

int rf_seq_join_all(rf_handle* h);   // k_seq.cu
int rf_seq_sync_all(rf_handle* h);

int rf_launch_compact_good(rf_handle* h, const float* d_feats, const float* d_next, const uint8_t* d_status, const int32_t* d_counts,
                           int Kmax, int P, float* d_good_old, float* d_good_new, int32_t* d_good_src, int32_t* d_ngood);

int rf_join_streams(rf_handle* h) {
    if (h->stream_copy) {
        RF_CUDA(h, cudaEventRecord(h->ev_copy, h->stream_copy));
        RF_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    }
    for (rf_batch* b : h->batches) {
        RF_CUDA(h, cudaEventRecord(b->ev_tail_done, b->tail));
        RF_CUDA(h, cudaStreamWaitEvent(h->stream, b->ev_tail_done, 0));
    }
    return rf_seq_join_all(h);
}

int rf_sync_all(rf_handle* h) {
    if (h->stream_copy) RF_CUDA(h, cudaStreamSynchronize(h->stream_copy));
    if (h->stream) RF_CUDA(h, cudaStreamSynchronize(h->stream));
    for (rf_batch* b : h->batches) RF_CUDA(h, cudaStreamSynchronize(b->tail));
    return rf_seq_sync_all(h);
}

// ------------------------------------------------------------------------------------
// getTransformKLT.py:368-376: good_new = nextPts[status == 1] (order preserved).  One warp
// per pair.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_compact_good(const float* __restrict__ feats, const float* __restrict__ next, const uint8_t* __restrict__ status,
               const int32_t* __restrict__ counts, int Kmax, int P, float* __restrict__ good_old,
               float* __restrict__ good_new, int32_t* __restrict__ good_src, int32_t* __restrict__ n_good) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int K = counts[p];
    const size_t base = (size_t)p * Kmax;
    int n = 0;
    for (int i0 = 0; i0 < K; i0 += 32) {
        const int i = i0 + lane;
        const bool ok = i < K && status[base + i];
        const unsigned bm = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const size_t j = base + n + __popc(bm & ((1u << lane) - 1));
            reinterpret_cast<float2*>(good_old)[j] = reinterpret_cast<const float2*>(feats)[base + i];
            reinterpret_cast<float2*>(good_new)[j] = reinterpret_cast<const float2*>(next)[base + i];
            good_src[j] = i;
        }
        n += __popc(bm);
    }
    if (lane == 0) n_good[p] = n;
}

int rf_launch_compact_good(rf_handle* h, const float* d_feats, const float* d_next, const uint8_t* d_status, const int32_t* d_counts,
                           int Kmax, int P, float* d_good_old, float* d_good_new, int32_t* d_good_src, int32_t* d_ngood) {
    k_compact_good<<<(P + 3) / 4, 128, 0, h->stream>>>(d_feats, d_next, d_status, d_counts, Kmax, P, d_good_old, d_good_new, d_good_src, d_ngood);
    RF_CHECK_LAUNCH(h);
    return RF_OK;
}

// Assemble rf_pair_result and the caller-facing corrStatus (Tracker.py:98-104).
__global__ void __launch_bounds__(128)
k_finish_pairs(int P, int Kmax, const int32_t* __restrict__ counts, const int32_t* __restrict__ n_good,
               const uint8_t* __restrict__ status, const int32_t* __restrict__ good_src,
               const uint8_t* __restrict__ cmask, int mask_stride, const int32_t* __restrict__ n_inl,
               const int32_t* __restrict__ nodes, const int32_t* __restrict__ cstatus, const double* __restrict__ R,
               const double* __restrict__ hpx, double res, const double* __restrict__ x, const int32_t* __restrict__ iters,
               int with_mds, uint8_t* __restrict__ corr, rf_pair_result* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const size_t base = (size_t)p * Kmax;
    const int K = counts[p], ng = n_good[p];
    for (int i = lane; i < Kmax; i += 32) corr[base + i] = 0;
    __syncwarp();
    for (int j = lane; j < ng; j += 32)
        if (cmask[(size_t)p * mask_stride + j]) corr[base + good_src[base + j]] = 1;
    if (lane == 0) {
        rf_pair_result r;
        for (int k = 0; k < 4; ++k) r.R[k] = R[(size_t)p * 4 + k];
        r.h[0] = hpx[(size_t)p * 2] * res; r.h[1] = hpx[(size_t)p * 2 + 1] * res;   // Tracker.py:125-126
        for (int k = 0; k < 6; ++k) r.mds_x[k] = with_mds ? x[(size_t)p * 6 + k] : 0.0;
        r.n_features = K; r.n_good = ng; r.n_inliers = n_inl[p];
        r.mds_iters = with_mds ? iters[p] : 0;
        r.status = cstatus[p]; r.clique_nodes = nodes[p];
        out[p] = r;
    }
}

// ------------------------------------------------------------------------------------
static void batch_free(rf_batch* b) {
    if (!b) return;
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(b->d_raw); F(b->d_rawi); rf_frameset_free(&b->fs);
    F(b->d_pair_idx); F(b->d_feats); F(b->d_counts); F(b->d_prev_pose);
    F(b->d_next); F(b->d_status); F(b->d_err);
    F(b->d_good_old); F(b->d_good_new); F(b->d_good_src); F(b->d_ngood);
    F(b->d_clique_ws); F(b->d_R); F(b->d_h); F(b->d_x); F(b->d_iters); F(b->d_mds_scratch);
    F(b->d_corr); F(b->d_results);
    for (int i = 0; i < RF_PROFILE_RING; ++i)
        for (int s = 0; s <= RF_N_STAGES; ++s)
            if (b->ev[i][s]) cudaEventDestroy(b->ev[i][s]);
    if (b->ev_uploaded) cudaEventDestroy(b->ev_uploaded);
    if (b->ev_main_done) cudaEventDestroy(b->ev_main_done);
    if (b->ev_tail_done) cudaEventDestroy(b->ev_tail_done);
    if (b->tail) cudaStreamDestroy(b->tail);
    delete b;
}

#define RF_BALLOC(ptr, bytes)                                                                       \
    do {                                                                                            \
        if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) {                                   \
            cudaGetLastError(); batch_free(b);                                                      \
            return rf_fail(h, RF_E_NOMEM, "rf_batch_create: cudaMalloc(%zu) failed", (size_t)(bytes)); \
        }                                                                                           \
    } while (0)

extern "C" {

static int batch_create_sized(rf_handle* h, int max_frames, int max_pairs, rf_batch** out);

int rf_batch_create(rf_handle* h, rf_batch** out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !out) return rf_fail(h, RF_E_BADARG, "rf_batch_create: null argument");
    return batch_create_sized(h, h->cfg.max_frames, h->cfg.max_pairs, out);
}

static int batch_create_sized(rf_handle* h, int max_frames, int max_pairs, rf_batch** out) {
    *out = nullptr;
    const rf_config& c = h->cfg;
    rf_batch* b = new rf_batch();
    memset(b, 0, sizeof(*b));
    b->max_frames = max_frames; b->max_pairs = max_pairs; b->Kmax = c.max_features;
    b->raw_cols = c.range_bins;                 // power bins only: the 11 metadata bytes are decoded on the host
    b->raw_pitch = (b->raw_cols + 15) & ~15;
    const size_t P = b->max_pairs, K = b->Kmax;
    RF_BALLOC(b->d_raw, (size_t)b->max_frames * c.azimuths * b->raw_pitch + 16);
    RF_BALLOC(b->d_rawi, rf_interleave_words(h, b->max_frames) * sizeof(uint32_t));
    int rc = rf_frameset_alloc(h, &b->fs, b->max_frames, c.write_cart_f32 != 0);
    if (rc) { batch_free(b); return rc; }
    RF_BALLOC(b->d_pair_idx, P * 2 * sizeof(int32_t));
    RF_BALLOC(b->d_feats, P * K * 2 * sizeof(float));
    RF_BALLOC(b->d_counts, P * sizeof(int32_t));
    RF_BALLOC(b->d_prev_pose, P * 3 * sizeof(double));
    RF_BALLOC(b->d_next, P * K * 2 * sizeof(float));
    RF_BALLOC(b->d_status, P * K);
    RF_BALLOC(b->d_err, P * K * sizeof(float));
    RF_BALLOC(b->d_good_old, P * K * 2 * sizeof(float));
    RF_BALLOC(b->d_good_new, P * K * 2 * sizeof(float));
    RF_BALLOC(b->d_good_src, P * K * sizeof(int32_t));
    RF_BALLOC(b->d_ngood, P * sizeof(int32_t));
    RF_BALLOC(b->d_clique_ws, rf_clique_ws_total((int)K, (int)P));
    RF_BALLOC(b->d_R, P * 4 * sizeof(double));
    RF_BALLOC(b->d_h, P * 2 * sizeof(double));
    RF_BALLOC(b->d_x, P * 6 * sizeof(double));
    RF_BALLOC(b->d_iters, P * sizeof(int32_t));
    RF_BALLOC(b->d_mds_scratch, P * K * 5 * sizeof(double));
    RF_BALLOC(b->d_corr, P * K);
    RF_BALLOC(b->d_results, P * sizeof(rf_pair_result));
    // slots k >= feat_counts[p] are never written by the kernels: define them (0) instead of leaving whatever the
    // allocator recycled, so that the [P][Kmax] outputs are reproducible bit for bit (found under compute-sanitizer)
    if (cudaMemset(b->d_next, 0, P * K * 2 * sizeof(float)) != cudaSuccess || cudaMemset(b->d_status, 0, P * K) != cudaSuccess ||
        cudaMemset(b->d_err, 0, P * K * sizeof(float)) != cudaSuccess || cudaMemset(b->d_corr, 0, P * K) != cudaSuccess ||
        cudaMemset(b->d_feats, 0, P * K * 2 * sizeof(float)) != cudaSuccess) {
        batch_free(b);
        return rf_fail(h, RF_E_CUDA, "rf_batch_create: cudaMemset failed");
    }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // tail streams run above the image kernels of later batches
    { const char* e = getenv("RADARFE_TAIL_PRIO"); if (e && e[0] == 'l') prio_hi = prio_lo; }   // diagnostic: tail at the image kernels' priority
    if (cudaStreamCreateWithPriority(&b->tail, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_uploaded, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_main_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_tail_done, cudaEventDisableTiming) != cudaSuccess) {
        batch_free(b);
        return rf_fail(h, RF_E_CUDA, "rf_batch_create: stream/event creation failed");
    }
    h->batches.push_back(b);
    *out = b;
    return RF_OK;
}

void rf_batch_destroy(rf_handle* h, rf_batch* b) {
    RfDeviceGuard rf_guard_(h);
    if (!b) return;
    if (h) {
        cudaSetDevice(h->device);
        rf_sync_all(h);
        h->batches.erase(std::remove(h->batches.begin(), h->batches.end(), b), h->batches.end());
    }
    batch_free(b);
}

int rf_batch_upload_async(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx,
                          int n_pairs, const float* feats, const int32_t* feat_counts, const double* prev_pose) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b || n_frames < 0 || n_pairs < 0 || (n_frames > 0 && !raw) ||
        (n_pairs > 0 && (!pair_idx || !feats || !feat_counts)))
        return rf_fail(h, RF_E_BADARG, "rf_batch_upload: null argument");
    if (n_frames > b->max_frames || n_pairs > b->max_pairs)
        return rf_fail(h, RF_E_CAPACITY, "rf_batch_upload: %d frames / %d pairs exceed the batch capacity (%d / %d)",
                       n_frames, n_pairs, b->max_frames, b->max_pairs);
    const rf_config& c = h->cfg;
    for (int p = 0; p < n_pairs; ++p) {
        if (feat_counts[p] < 0 || feat_counts[p] > b->Kmax)
            return rf_fail(h, RF_E_CAPACITY, "rf_batch_upload: pair %d has %d features (max_features = %d)", p,
                           feat_counts[p], b->Kmax);
        if (pair_idx[2 * p] < 0 || pair_idx[2 * p] >= n_frames || pair_idx[2 * p + 1] < 0 || pair_idx[2 * p + 1] >= n_frames)
            return rf_fail(h, RF_E_BADARG, "rf_batch_upload: pair %d references a frame outside [0, %d)", p, n_frames);
    }
    cudaStream_t cs = h->stream_copy;
    // the previous run of this batch must have consumed its inputs before they are overwritten
    RF_CUDA(h, cudaStreamWaitEvent(cs, b->ev_main_done, 0));
    RF_CUDA(h, cudaStreamWaitEvent(cs, b->ev_tail_done, 0));
    if (n_frames)
        RF_CUDA(h, cudaMemcpy2DAsync(b->d_raw, b->raw_pitch, raw + c.meta_bytes, c.raw_width, b->raw_cols, (size_t)n_frames * c.azimuths,
                                     cudaMemcpyHostToDevice, cs));
    if (n_pairs) {
        RF_CUDA(h, cudaMemcpyAsync(b->d_pair_idx, pair_idx, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
        RF_CUDA(h, cudaMemcpyAsync(b->d_feats, feats, (size_t)n_pairs * b->Kmax * 2 * sizeof(float), cudaMemcpyHostToDevice, cs));
        RF_CUDA(h, cudaMemcpyAsync(b->d_counts, feat_counts, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
        if (prev_pose)
            RF_CUDA(h, cudaMemcpyAsync(b->d_prev_pose, prev_pose, (size_t)n_pairs * 3 * sizeof(double), cudaMemcpyHostToDevice, cs));
    }
    RF_CUDA(h, cudaEventRecord(b->ev_uploaded, cs));
    b->n_frames = n_frames; b->n_pairs = n_pairs; b->has_pose = prev_pose != nullptr;
    return RF_OK;
}

int rf_batch_upload(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx, int n_pairs,
                    const float* feats, const int32_t* feat_counts, const double* prev_pose) {
    RfDeviceGuard rf_guard_(h);
    int rc = rf_batch_upload_async(h, b, raw, n_frames, pair_idx, n_pairs, feats, feat_counts, prev_pose);
    if (rc) return rc;
    RF_CUDA(h, cudaStreamSynchronize(h->stream_copy));
    return RF_OK;
}

int rf_batch_set_profiling(rf_handle* h, rf_batch* b, int on) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b) return rf_fail(h, RF_E_BADARG, "rf_batch_set_profiling: null argument");
    if (on && !b->ev[0][0]) {
        for (int i = 0; i < RF_PROFILE_RING; ++i)
            for (int s = 0; s <= RF_N_STAGES; ++s) RF_CUDA(h, cudaEventCreate(&b->ev[i][s]));
    }
    b->profiling = on != 0; b->prof_head = 0; b->prof_count = 0;
    return RF_OK;
}

int rf_batch_run_async(rf_handle* h, rf_batch* b, int with_mds) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b) return rf_fail(h, RF_E_BADARG, "rf_batch_run_async: null argument");
    const rf_config& c = h->cfg;
    const int P = b->n_pairs, F = b->n_frames, K = b->Kmax;
    cudaEvent_t* ev = b->profiling ? b->ev[b->prof_head] : nullptr;
    int stage = 0;
    auto mark = [&]() { if (ev) cudaEventRecord(ev[stage], h->stream); ++stage; };
    int rc;
    RF_CUDA(h, cudaStreamWaitEvent(h->stream, b->ev_uploaded, 0));
    // with KLT on the tail stream the previous run's tail still reads this batch's pyramid: wait before rebuilding it
    if (g_klt_on_tail && P) RF_CUDA(h, cudaStreamWaitEvent(h->stream, b->ev_tail_done, 0));
    mark();
    const bool fused = b->fs.cart == nullptr && b->fs.n_levels >= 2;
    if (F && fused) {
        // stage 0: frame interleave; stage 1: level 0 + level 1 in one kernel; stage 2: the higher levels
        if ((rc = rf_launch_interleave(h, b->d_raw, (size_t)c.azimuths * b->raw_pitch, b->raw_pitch, F, b->d_rawi))) return rc;
        mark();
        if ((rc = rf_launch_scan_to_l0l1(h, b->d_rawi, b->fs, F))) return rc;
        mark();
        if ((rc = rf_launch_pyr_levels(h, b->fs, 2, F))) return rc;
        mark();
    } else {
        // stage 0: polar->Cartesian (f32 + u8 level 0); stage 2: the whole pyramid
        if (F && (rc = rf_launch_polar2cart_u8(h, b->d_raw, (size_t)c.azimuths * b->raw_pitch, b->raw_pitch, 0, b->fs, 0, F,
                                                b->fs.cart != nullptr))) return rc;
        mark();
        mark();
        if (F && (rc = rf_launch_pyramid(h, b->fs, 0, F))) return rc;
        mark();
    }
    if (P) {
        // KLT either closes the main-stream part (klt_on_tail = 0) or opens the tail-stream part: on the (high-priority)
        // tail stream it co-runs with the image kernels of the NEXT batch instead of queueing behind/before them
        const bool klt_on_tail = g_klt_on_tail;
        if (klt_on_tail) {
            RF_CUDA(h, cudaEventRecord(b->ev_main_done, h->stream));
            RF_CUDA(h, cudaStreamWaitEvent(b->tail, b->ev_main_done, 0));
        } else {
            // the tail of this batch's previous run still reads the track buffers KLT is about to overwrite
            RF_CUDA(h, cudaStreamWaitEvent(h->stream, b->ev_tail_done, 0));
        }
        StreamScope on_tail(h, klt_on_tail ? b->tail : h->stream);
        if ((rc = rf_launch_klt(h, b->fs, b->fs, b->d_pair_idx, b->d_feats, b->d_counts, K, P, b->d_next, b->d_status,
                                b->d_err, 1))) return rc;
        mark();
        k_compact_good<<<(P + 3) / 4, 128, 0, h->stream>>>(b->d_feats, b->d_next, b->d_status, b->d_counts, K, P,
                                                           b->d_good_old, b->d_good_new, b->d_good_src, b->d_ngood);
        RF_CHECK_LAUNCH(h);
        if (!klt_on_tail) {
            RF_CUDA(h, cudaEventRecord(b->ev_main_done, h->stream));
            RF_CUDA(h, cudaStreamWaitEvent(b->tail, b->ev_main_done, 0));
            h->stream = b->tail;             // everything below runs on the batch's tail stream (restored by on_tail)
        }
        mark();
        uint8_t* d_mask; int mask_stride; int32_t *d_ninl, *d_nodes, *d_cstatus;
        if ((rc = rf_launch_reject(h, b->d_clique_ws, b->d_good_old, b->d_good_new, b->d_ngood, K, P, &d_mask, &mask_stride,
                                   &d_ninl, &d_nodes, &d_cstatus))) return rc;
        mark();
        // Tracker.getTransform(good_old, good_new): src = good_old, target = good_new
        if ((rc = rf_launch_kabsch(h, b->d_good_old, b->d_good_new, d_mask, mask_stride, b->d_ngood, K, P, b->d_R, b->d_h,
                                   nullptr))) return rc;
        mark();
        if (with_mds) {
            if ((rc = rf_launch_mds_fused(h, b->d_good_old, b->d_good_new, d_mask, mask_stride, b->d_ngood, K, P, b->d_R,
                                          b->d_h, b->has_pose ? b->d_prev_pose : nullptr, b->d_mds_scratch, b->d_x,
                                          b->d_iters))) return rc;
        }
        mark();
        k_finish_pairs<<<(P + 3) / 4, 128, 0, h->stream>>>(P, K, b->d_counts, b->d_ngood, b->d_status, b->d_good_src, d_mask,
                                                           mask_stride, d_ninl, d_nodes, d_cstatus, b->d_R, b->d_h,
                                                           c.cart_res_m, b->d_x, b->d_iters, with_mds, b->d_corr,
                                                           b->d_results);
        RF_CHECK_LAUNCH(h);
        mark();
        RF_CUDA(h, cudaEventRecord(b->ev_tail_done, b->tail));
    } else {
        while (stage <= RF_N_STAGES) mark();
        RF_CUDA(h, cudaEventRecord(b->ev_main_done, h->stream));
    }
    b->ran_mds = with_mds != 0;
    if (ev) { b->prof_head = (b->prof_head + 1) % RF_PROFILE_RING; if (b->prof_count < RF_PROFILE_RING) b->prof_count++; }
    return RF_OK;
}

int rf_batch_stage_times(rf_handle* h, rf_batch* b, float* ms_sum, int cap, int* n_runs) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b || !ms_sum || cap < RF_N_STAGES) return rf_fail(h, RF_E_BADARG, "rf_batch_stage_times: bad argument");
    for (int s = 0; s < cap; ++s) ms_sum[s] = 0.f;
    if (n_runs) *n_runs = b->prof_count;
    if (!b->profiling) return RF_OK;
    int rcs = rf_sync_all(h);
    if (rcs) return rcs;
    for (int r = 0; r < b->prof_count; ++r) {
        const int i = (b->prof_head - 1 - r + 2 * RF_PROFILE_RING) % RF_PROFILE_RING;
        for (int s = 0; s < RF_N_STAGES; ++s) {
            float ms = 0.f;
            RF_CUDA(h, cudaEventElapsedTime(&ms, b->ev[i][s], b->ev[i][s + 1]));
            ms_sum[s] += ms;
        }
    }
    b->prof_count = 0;
    return RF_OK;
}

int rf_batch_download_async(rf_handle* h, rf_batch* b, rf_pair_result* results, float* next_xy, uint8_t* status) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b || (b->n_pairs > 0 && !results)) return rf_fail(h, RF_E_BADARG, "rf_batch_download: null argument");
    const size_t P = b->n_pairs, K = b->Kmax;
    if (!P) return RF_OK;
    RF_CUDA(h, cudaMemcpyAsync(results, b->d_results, P * sizeof(rf_pair_result), cudaMemcpyDeviceToHost, b->tail));
    if (next_xy) RF_CUDA(h, cudaMemcpyAsync(next_xy, b->d_next, P * K * 2 * sizeof(float), cudaMemcpyDeviceToHost, b->tail));
    if (status) RF_CUDA(h, cudaMemcpyAsync(status, b->d_corr, P * K, cudaMemcpyDeviceToHost, b->tail));
    RF_CUDA(h, cudaEventRecord(b->ev_tail_done, b->tail));
    return RF_OK;
}

int rf_batch_download(rf_handle* h, rf_batch* b, rf_pair_result* results, float* next_xy, uint8_t* status) {
    RfDeviceGuard rf_guard_(h);
    int rc = rf_batch_download_async(h, b, results, next_xy, status);
    if (rc) return rc;
    RF_CUDA(h, cudaStreamSynchronize(b->tail));
    return RF_OK;
}

int rf_batch_wait(rf_handle* h, rf_batch* b) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b) return rf_fail(h, RF_E_BADARG, "rf_batch_wait: null argument");
    RF_CUDA(h, cudaStreamSynchronize(b->tail));
    return RF_OK;
}

int rf_batch_klt_status(rf_handle* h, rf_batch* b, uint8_t* klt_status, float* err) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b) return rf_fail(h, RF_E_BADARG, "rf_batch_klt_status: null argument");
    const size_t P = b->n_pairs, K = b->Kmax;
    if (!P) return RF_OK;
    { int rcs = rf_sync_all(h); if (rcs) return rcs; }
    if (klt_status) RF_CUDA(h, cudaMemcpyAsync(klt_status, b->d_status, P * K, cudaMemcpyDeviceToHost, h->stream));
    if (err) RF_CUDA(h, cudaMemcpyAsync(err, b->d_err, P * K * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

// Tracker.py:62-63 for every pair of the batch: FMT rotation prior from the RESIDENT raw scans (no extra upload).
// Runs on the handle's main stream after the batch's upload; synchronous on return.
int rf_batch_fmt(rf_handle* h, rf_batch* b, int downsample, int clip_px, double* angle_rad, double* scale, double* response,
                 double* shift_xy) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b || !angle_rad) return rf_fail(h, RF_E_BADARG, "rf_batch_fmt: null argument");
    const int P = b->n_pairs, F = b->n_frames;
    if (!P) return RF_OK;
    RF_CUDA(h, cudaStreamWaitEvent(h->stream, b->ev_uploaded, 0));
    double* d_out = nullptr;
    int sz = 0;
    double log_base = 1.0;
    int rc = rf_fmt_resident_u8(h, b->d_raw, F, h->cfg.azimuths, b->raw_cols, (size_t)b->raw_pitch, b->d_pair_idx, P, downsample,
                                clip_px, &d_out, &sz, &log_base);
    if (rc) return rc;
    std::vector<double> out((size_t)P * 3);
    RF_CUDA(h, cudaMemcpyAsync(out.data(), d_out, (size_t)P * 24, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    rf_fmt_finish(out.data(), P, sz, log_base, angle_rad, scale, response, shift_xy);
    return RF_OK;
}

int rf_track_batch(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx, int n_pairs,
                   const float* feats, const int32_t* feat_counts, const double* prev_pose, int with_mds,
                   rf_pair_result* results, float* next_xy, uint8_t* status) {
    RfDeviceGuard rf_guard_(h);
    int rc = rf_batch_upload_async(h, b, raw, n_frames, pair_idx, n_pairs, feats, feat_counts, prev_pose);
    if (rc) return rc;
    if ((rc = rf_batch_run_async(h, b, with_mds))) return rc;
    return rf_batch_download(h, b, results, next_xy, status);
}

// Tracker.track + getTransform (+ MDS) for ONE pair given as two separate raw scans (SURVEY.md 8b rf_track_pair): a
// two-frame, one-pair batch owned by the handle, created on first use.  feats_xy [K, 2] f32, K <= max_features;
// next_xy [K, 2] / corr_status [K] may be NULL.  Synchronous on return.
int rf_track_pair(rf_handle* h, const uint8_t* raw_prev, const uint8_t* raw_next, const float* feats_xy, int K,
                  const double* prev_pose, int with_mds, rf_pair_result* result, float* next_xy, uint8_t* corr_status) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !raw_prev || !raw_next || !result || K < 0 || (K > 0 && !feats_xy))
        return rf_fail(h, RF_E_BADARG, "rf_track_pair: bad argument");
    const rf_config& c = h->cfg;
    if (K > c.max_features) return rf_fail(h, RF_E_CAPACITY, "rf_track_pair: %d features exceed max_features = %d", K, c.max_features);
    int rc;
    if (!h->pair_batch && (rc = batch_create_sized(h, 2, 1, &h->pair_batch))) return rc;
    rf_batch* b = h->pair_batch;
    const size_t scan = (size_t)c.azimuths * c.raw_width, kmax = (size_t)c.max_features;
    const size_t off_feats = (2 * scan + 255) & ~(size_t)255, off_next = off_feats + ((kmax * 8 + 255) & ~(size_t)255);
    const size_t off_st = off_next + ((kmax * 8 + 255) & ~(size_t)255);
    if ((rc = rf_ensure_pinned(h, off_st + kmax + 256))) return rc;
    uint8_t* hp = (uint8_t*)h->h_pinned;
    memcpy(hp, raw_prev, scan);
    memcpy(hp + scan, raw_next, scan);
    float* hf = (float*)(hp + off_feats);
    memset(hf, 0, kmax * 8);
    if (K) memcpy(hf, feats_xy, (size_t)K * 8);
    const int32_t pair[2] = {0, 1}, count = K;
    if ((rc = rf_batch_upload_async(h, b, hp, 2, pair, 1, hf, &count, prev_pose))) return rc;
    if ((rc = rf_batch_run_async(h, b, with_mds))) return rc;
    float* hn = (float*)(hp + off_next);
    uint8_t* hs = hp + off_st;
    if ((rc = rf_batch_download(h, b, result, next_xy ? hn : nullptr, corr_status ? hs : nullptr))) return rc;
    if (next_xy && K) memcpy(next_xy, hn, (size_t)K * 8);
    if (corr_status && K) memcpy(corr_status, hs, (size_t)K);
    return RF_OK;
}

int rf_batch_frame_download(rf_handle* h, const rf_batch* b, int frame, int what, void* out, int* rows, int* cols) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !b || frame < 0 || frame >= b->n_frames) return rf_fail(h, RF_E_BADARG, "rf_batch_frame_download: bad frame");
    { int rcs = rf_sync_all(h); if (rcs) return rcs; }
    if (what == 0) {
        if (!b->fs.cart) return rf_fail(h, RF_E_BADARG, "rf_batch_frame_download: batch was created with write_cart_f32 = 0");
        if (rows) *rows = h->n;
        if (cols) *cols = h->n;
        if (out) RF_CUDA(h, cudaMemcpyAsync(out, b->fs.cart + (size_t)frame * b->fs.cart_stride, (size_t)h->n * h->n * 4,
                                            cudaMemcpyDeviceToHost, h->stream));
    } else {
        const int l = what - 1;
        if (l < 0 || l >= b->fs.n_levels) return rf_fail(h, RF_E_BADARG, "rf_batch_frame_download: no pyramid level %d", l);
        if (rows) *rows = b->fs.h[l];
        if (cols) *cols = b->fs.w[l];
        if (out) RF_CUDA(h, cudaMemcpy2DAsync(out, b->fs.w[l], b->fs.lvl[l] + (size_t)frame * b->fs.lvl_stride[l], b->fs.pitch[l],
                                              b->fs.w[l], b->fs.h[l], cudaMemcpyDeviceToHost, h->stream));
    }
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

// pinned host staging for callers that want asynchronous uploads (NumPy arrays can wrap it)
int rf_host_alloc(size_t bytes, void** out) {
    if (!out) return RF_E_BADARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return rf_fail(nullptr, RF_E_NOMEM, "rf_host_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    return RF_OK;
}
void rf_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
