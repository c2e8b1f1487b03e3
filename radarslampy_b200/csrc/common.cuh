// common.cuh — handle layout, error plumbing and small device helpers shared by every
// translation unit of libradarfe.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/radarfe.h"

#define RF_MAX_LEVELS 4  // pyramid levels 0..3 (klt_max_level <= 3)

// A set of `count` device-resident frames with uniform strides (count == 1 for rf_frame).
struct FrameSet {
    float* cart;                   // [count][n][n] f32 (null when the batch path elides it)
    size_t cart_stride;            // elements between frames
    uint8_t* lvl[RF_MAX_LEVELS];   // u8 pyramid levels, rows `pitch` bytes apart
    size_t lvl_stride[RF_MAX_LEVELS];  // bytes between frames (multiple of 256)
    int w[RF_MAX_LEVELS], h[RF_MAX_LEVELS];
    int pitch[RF_MAX_LEVELS];      // row pitch in bytes: w rounded up to 16 (TMA tensor maps need 16-byte strides)
    int n_levels;
    int count;
};
struct rf_frame {
    FrameSet fs;
};

struct rf_batch;
struct rf_seq;
struct rf_handle {
    rf_config cfg;
    int device;
    cudaStream_t stream;       // main stream: every single-stage entry point, and the image + KLT stages of a batch
    bool owns_stream;
    cudaStream_t stream_copy;  // H2D staging of the batch path (overlaps the kernels of the previous batch)
    cudaEvent_t ev_copy;       // join point of stream_copy for the timers / rf_sync
    std::vector<rf_batch*> batches;   // live batches (each owns a tail stream for rejection + solves)
    rf_batch* pair_batch;             // two-frame / one-pair batch behind rf_track_pair (created on first use)
    std::vector<rf_seq*> seqs;        // live lock-step sequence runners (each owns a stream)
    int n;          // cartesian size 2R
    int R;
    int sm_count;
    // geometry table: packed fixed-point sample coordinates of cv2.warpPolar's inverse map
    uint32_t* map;  // [n][n]  (sx | sy << 17), sx = round(32*rho), sy = round(32*(phi+1))
    uint4* map2;    // [n][n]  geometry records of the fused batch path (k_fused.cu)
    // staging
    uint8_t* d_raw;      // one raw scan
    float* d_polar;      // one f32 polar image
    uint8_t* d_polar_u8; // [A][range_bins] u8 recovered from an f32 polar
    void* d_scratch; size_t scratch_bytes;    // generic device scratch (grown on demand)
    uint64_t scratch_gen;                     // bumped whenever the arena is reallocated (contents lost)
    void* h_pinned; size_t pinned_bytes;      // generic pinned host staging
    cudaEvent_t ev0, ev1;
    int64_t launches;
    std::string err;
};

// Every extern "C" entry point that takes a handle runs on the handle's device whatever the calling thread's current
// device is (a second handle on another GPU, a Python thread, a host application that called cudaSetDevice), and
// leaves the caller's current device as it found it.
struct RfDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit RfDeviceGuard(const rf_handle* h) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = cudaSetDevice(h->device) == cudaSuccess;
    }
    ~RfDeviceGuard() { if (switched) cudaSetDevice(prev); }
    RfDeviceGuard(const RfDeviceGuard&) = delete;
    RfDeviceGuard& operator=(const RfDeviceGuard&) = delete;
};

extern thread_local std::string g_rf_err;  // for failures without a handle

int rf_fail(rf_handle* h, int code, const char* fmt, ...);
int rf_ensure_scratch(rf_handle* h, size_t bytes);
int rf_ensure_pinned(rf_handle* h, size_t bytes);
// k_batch.cu: make the main stream wait for everything outstanding on the copy / tail streams; full sync
int rf_join_streams(rf_handle* h);
int rf_sync_all(rf_handle* h);

#define RF_CUDA(h, expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return rf_fail((h), RF_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                  \
    } while (0)

#define RF_CHECK_LAUNCH(h)                                                                       \
    do {                                                                                         \
        (h)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess)                                                                   \
            return rf_fail((h), RF_E_CUDA, "kernel launch failed: %s (%s:%d)",                   \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                          \
    } while (0)

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
    // cv::BORDER_REFLECT_101 for |overshoot| < len (always true here: borders <= 16 px, len > 15)
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return p;
}
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }  // ties-to-even, like cvRound
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

// ---- per-stage launchers (defined in the k_*.cu files) ------------------------------
int rf_frameset_alloc(rf_handle* h, FrameSet* fs, int count, bool with_f32);
void rf_frameset_free(FrameSet* fs);
int rf_launch_build_map(rf_handle* h);
// src: `n_frames` scans, frame i at d_src + i*src_frame_stride (elements), row pitch and first
// power column given; tap type u8 (raw scan) or f32 (caller polar image)
int rf_launch_polar2cart_u8(rf_handle* h, const uint8_t* d_src, size_t src_frame_stride, int row_pitch,
                            int col0, const FrameSet& dst, int first, int n_frames, bool write_f32);
int rf_launch_polar2cart_f32(rf_handle* h, const float* d_src, int row_pitch, const FrameSet& dst);
int rf_launch_cart_to_u8(rf_handle* h, const FrameSet& fs);
int rf_launch_pyramid(rf_handle* h, const FrameSet& fs, int first, int n_frames);
int rf_launch_extract(rf_handle* h, const uint8_t* d_raw, float* d_polar);

// k_fused.cu — batch image path (interleaved scans -> level 0 + pyramid)
int rf_fused_wp(const rf_handle* h);
size_t rf_interleave_words(const rf_handle* h, int max_frames);
int rf_launch_build_map2(rf_handle* h);
// frame_sel: optional device pointer to {base, stride}: frame f of the group is read from source frame base + f * stride
int rf_launch_interleave(rf_handle* h, const uint8_t* d_raw, size_t frame_stride, int pitch, int n_frames, uint32_t* d_out,
                         const int32_t* d_frame_sel = nullptr);
int rf_launch_scan_to_l0l1(rf_handle* h, const uint32_t* d_rawi, const FrameSet& fs, int n_frames);
int rf_launch_pyr_levels(rf_handle* h, const FrameSet& fs, int first_level, int n_frames);
// k_klt.cu
int rf_launch_klt(rf_handle* h, const FrameSet& prev, const FrameSet& next, const int32_t* d_pair_idx, const float* d_pts,
                  const int32_t* d_counts, int Kmax, int P, float* d_next, uint8_t* d_status, float* d_err, int gate);
// k_batch.cu
int rf_launch_compact_good(rf_handle* h, const float* d_feats, const float* d_next, const uint8_t* d_status, const int32_t* d_counts,
                           int Kmax, int P, float* d_good_old, float* d_good_new, int32_t* d_good_src, int32_t* d_ngood);
// k_clique.cu
size_t rf_clique_ws_total(int Kmax, int P);
int rf_launch_reject(rf_handle* h, void* ws_base, const float* d_prev, const float* d_new, const int32_t* d_counts,
                     int Kstride, int P, uint8_t** d_mask_out, int* mask_stride, int32_t** d_ninl, int32_t** d_nodes,
                     int32_t** d_status);
// k_solve.cu
int rf_launch_kabsch(rf_handle* h, const float* d_src, const float* d_tgt, const uint8_t* d_mask, int mask_stride,
                     const int32_t* d_counts, int Kstride, int P, double* d_R, double* d_h, int32_t* d_nused);
int rf_launch_mds_fused(rf_handle* h, const float* d_old, const float* d_new, const uint8_t* d_mask, int mask_stride,
                        const int32_t* d_counts, int Kstride, int P, const double* d_R, const double* d_h,
                        const double* d_prev_pose, double* d_scratch, double* d_x, int32_t* d_iters);
int rf_launch_mds_chain(rf_handle* h, const float* d_old, const float* d_new, const uint8_t* d_mask, int mask_stride,
                        const int32_t* d_counts, int Kstride, int P, const double* d_R, const double* d_h,
                        const double* d_prev_pose, const double* d_kf_und, const double* d_kf_pose, const int32_t* d_good_src,
                        double* d_scratch, double* d_x, int32_t* d_iters);
// k_fmt.cu — Fourier-Mellin rotation prior over device-resident scans (used by rf_batch_fmt)
int rf_fmt_resident_u8(rf_handle* h, const uint8_t* d_raw, int F, int A, int W, size_t pitch, const int32_t* d_pairs, int P,
                       int downsample, int clip_px, double** d_out_p, int* sz_out, double* log_base_out);
void rf_fmt_finish(const double* out3, int P, int sz, double log_base, double* angle_rad, double* scale, double* response,
                   double* shift_xy);
