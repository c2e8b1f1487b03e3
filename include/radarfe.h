/*
 * radarfe.h — C ABI of libradarfe.so: the B200 (sm_100a) radar-odometry front end.
 *
 * This is the drop-in boundary for the per-frame hot path of Samleo8/RadarSLAMPy
 * ("RAW-ROAM").  The reference has no FFI of its own (it is pure Python over
 * OpenCV / SciPy / networkx); each entry point below therefore replaces one Python
 * call site of the reference, cited as file:line relative to the reference root.
 * The Python modules in radarslampy_b200/ keep the reference's names and signatures
 * and call these symbols through ctypes.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every function returns RF_OK (0) or a
 *     negative rf_status; rf_last_error() gives the message for the last failure.
 *   - the caller owns every host buffer (C-contiguous, sizes passed explicitly); the
 *     library never keeps a host pointer after the call returns.
 *   - device memory, the CUDA stream and all workspaces are owned by the rf_handle.
 *   - a handle is single-threaded; different handles may be used from different threads.
 *   - entry points are synchronous on return unless their name ends in _async.
 *   - there is NO CPU fallback: without a CUDA device rf_create fails with RF_E_CUDA.
 */
#ifndef RADARFE_H
#define RADARFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADARFE_VERSION 100 /* 0.1.0 */
#if defined(__GNUC__)
#define RF_API __attribute__((visibility("default")))
#else
#define RF_API
#endif

typedef enum rf_status {
    RF_OK = 0,
    RF_E_BADARG = -1,     /* null pointer, negative size, shape mismatch                  */
    RF_E_CAPACITY = -2,   /* input exceeds a maximum fixed at rf_create (see rf_config)   */
    RF_E_CUDA = -3,       /* CUDA runtime / driver error, or no sm_100 device             */
    RF_E_WORKLIMIT = -4,  /* clique search exceeded rf_config.clique_node_limit           */
    RF_E_NOMEM = -5
} rf_status;

/* Constants of the reference, as one POD.  rf_default_config() fills the reference's
 * values; parity tests run with the defaults. */
typedef struct rf_config {
    int32_t azimuths;          /* 400   rows of an Oxford scan            parseData.py:39-43  */
    int32_t raw_width;         /* 3779  bytes per row: 11 metadata + 3768 bins                */
    int32_t meta_bytes;        /* 11                                                          */
    int32_t range_bins;        /* 2025  int(87.5 / 0.0432)                parseData.py:49-51  */
    int32_t downsample;        /* 2     => R = range_bins / 2, image 2R x 2R   parseData.py:119-125 */
    int32_t max_features;      /* capacity per frame pair (K)                                  */
    int32_t max_pairs;         /* capacity of one rf_batch                                     */
    int32_t max_frames;        /* frames resident in one rf_batch                              */
    int32_t klt_win;           /* 15    getTransformKLT.py:343                                 */
    int32_t klt_max_level;     /* 3     getTransformKLT.py:77-84                               */
    int32_t klt_max_iters;     /* 10                                                           */
    float   klt_eps;           /* 0.03                                                         */
    float   klt_min_eig;       /* 1e-4  cv2 default minEigThreshold                            */
    float   klt_err_thr;       /* 10    ERR_THRESHOLD getTransformKLT.py:84,365                */
    double  dist_thr_px;       /* 0.5 / 0.0864 = 5.787037037037036  outlierRejection.py:10-11  */
    double  cart_res_m;        /* 0.0864 m per pixel                parseData.py:10-13         */
    double  mds_period;        /* 0.25 s  (1 / RADAR_SCAN_FREQUENCY) motionDistortion.py:36    */
    double  mds_sigma_p[2];    /* (4, 4)                             RawROAMSystem.py:135-139  */
    double  mds_sigma_v[3];    /* (1, 1, (5 deg)^2)                                            */
    int64_t clique_node_limit; /* per-pair bound on search-tree descents (RF_E_WORKLIMIT)      */
    int32_t write_cart_f32;    /* batch path: also materialise the f32 Cartesian image         */
    int32_t retrack_threshold; /* 60    N_FEATURES_BEFORE_RETRACK getFeatures.py:57, RawROAMSystem.py:250 */
    int32_t ssc_num_ret;       /* 200   adaptiveNMS ret_points        getFeatures.py:66             */
    int32_t doh_num_sigma;     /* 3     DEFAULT_FEATURE_PARAMS        getFeatures.py:13-18          */
    double  ssc_tolerance;     /* 0.1                                                               */
    double  kf_rot_thr;        /* 0.2 rad  ROT_THRESHOLD              Mapping.py:13-15              */
    double  kf_trans_thr;      /* 2.0 m    TRANS_THRESHOLD                                          */
    double  detect_quality;    /* mode 0: candidates above this fraction of the strongest response (0.01) */
    double  doh_min_sigma;     /* 0.01                                getFeatures.py:13-18          */
    double  doh_max_sigma;     /* 10                                                                */
    double  doh_threshold;     /* 0.0005                                                            */
} rf_config;

typedef struct rf_handle rf_handle; /* stream + workspaces + geometry tables            */
typedef struct rf_frame rf_frame;   /* one device-resident scan: f32 cart + u8 pyramid   */
typedef struct rf_batch rf_batch;   /* device-resident batch of independent frame pairs  */

/* Result of one frame pair (what Tracker.track + Tracker.getTransform + the MDS solve
 * return for it).  `src = R * target + h`, h in metres (getTransformKLT.py:129-162,
 * Tracker.py:108-127). */
typedef struct rf_pair_result {
    double R[4];      /* row-major 2x2                                                       */
    double h[2];      /* metres                                                              */
    double mds_x[6];  /* [vx, vy, vtheta, Tx, Ty, Ttheta]   motionDistortion.py:295-325      */
    int32_t n_features;  /* K handed in                                                      */
    int32_t n_good;      /* after KLT status & (err < thr)      getTransformKLT.py:364-376   */
    int32_t n_inliers;   /* clique size                          outlierRejection.py:71-78   */
    int32_t mds_iters;
    int32_t status;      /* RF_OK or RF_E_WORKLIMIT for this pair                            */
    int32_t clique_nodes; /* search-tree descents spent                                      */
} rf_pair_result;

/* ---- lifetime ------------------------------------------------------------------- */
RF_API void rf_default_config(rf_config* cfg);
RF_API int rf_create(const rf_config* cfg, int device, void* cuda_stream_or_null, rf_handle** out);
RF_API void rf_destroy(rf_handle* h);
RF_API const char* rf_last_error(const rf_handle* h_or_null);
RF_API int rf_version(void);
RF_API int rf_cart_size(const rf_handle* h);          /* 2R */
RF_API void* rf_stream(const rf_handle* h);           /* cudaStream_t the kernels run on */
/* CUDA-event timing on the handle's stream (bench.py uses it: torch events do not see
 * this stream).  rf_timer_stop_ms synchronises and returns elapsed milliseconds. */
RF_API int rf_timer_start(rf_handle* h);
RF_API int rf_timer_stop_ms(rf_handle* h, float* ms);
/* number of kernels launched by this handle since creation (bench.py: gpu_launches) */
RF_API int64_t rf_launch_count(const rf_handle* h);

/* ---- a1  parseData.extractDataFromRadarImage            parseData.py:17-53 ------- */
/* raw u8 [A, raw_width] -> polar f32 [A, range_bins] (= u8 / 255.f, IEEE division),
 * timestamps i64 [A], azimuths f32 [A] (radians), valid u8 [A]. Any output may be NULL. */
RF_API int rf_extract_polar(rf_handle* h, const uint8_t* raw, float* polar, int64_t* timestamps,
                     float* azimuths, uint8_t* valid);

/* ---- frames ---------------------------------------------------------------------- */
RF_API int rf_frame_create(rf_handle* h, rf_frame** out);
RF_API void rf_frame_destroy(rf_handle* h, rf_frame* f);

/* ---- a2+a3  parseData.convertPolarImageToCartesian       parseData.py:100-135 -----
 *             (img*255).astype(uint8) + cv2 pyramid         getTransformKLT.py:356-360
 * Exactly one of raw (u8 [A, raw_width]) / polar (f32 [A, range_bins], values k/255)
 * is non-NULL.  Fills `frame` (f32 cart, u8 image, pyramid levels 1..klt_max_level) and,
 * if cart_out != NULL, copies the f32 image [2R, 2R] to the host. */
RF_API int rf_polar_to_cart(rf_handle* h, const uint8_t* raw, const float* polar, rf_frame* frame,
                     float* cart_out);
/* convertPolarImageToCartesian(imgPolar, logPolarMode=True): cv2.warpPolar with WARP_POLAR_LOG | WARP_INVERSE_MAP
 *                                                              parseData.py:100-135 (:131-133)
 * polar f32 [A, range_bins] (host) -> cart_out f32 [2R, 2R] (host).  Never taken on the reference's live path; matches
 * cv2 to the tolerance of its vendor logf (1 ulp on the radius -> a different 1/32-pixel sample position for a few pixels
 * in a thousand). */
RF_API int rf_polar_to_cart_log(rf_handle* h, const float* polar, float* cart_out);
/* Build a frame from a caller-supplied f32 Cartesian image [n, n], n == 2R (the path
 * getTrackedPointsKLT takes when handed plain NumPy images). */
RF_API int rf_frame_from_cart(rf_handle* h, const float* cart, int n, rf_frame* frame);
/* what: 0 = f32 cart [2R,2R]; 1+l = u8 pyramid level l.  *rows/*cols receive the shape. */
RF_API int rf_frame_download(rf_handle* h, const rf_frame* f, int what, void* out, int* rows, int* cols);

/* ---- a4+a5  cv2.calcOpticalFlowPyrLK + err gating        getTransformKLT.py:359-376 */
/* pts [K,2] (x,y) -> next [K,2], status u8 [K] (cv2 status & (err < klt_err_thr) when
 * apply_err_gate != 0), err f32 [K]. */
RF_API int rf_klt(rf_handle* h, const rf_frame* prev, const rf_frame* next, const float* pts, int K,
           int apply_err_gate, float* next_xy, uint8_t* status, float* err);

/* ---- a6  outlierRejection.rejectOutliers                 outlierRejection.py:16-95 */
/* mask u8 [K] = membership in the first largest maximal clique in networkx order. */
RF_API int rf_reject_outliers(rf_handle* h, const float* prev_xy, const float* new_xy, int K,
                       uint8_t* mask, int* n_inliers, int* nodes_or_null);
/* adjacency only (K x K bytes), for tests                   outlierRejection.py:49-58 */
RF_API int rf_consistency_adjacency(rf_handle* h, const float* prev_xy, const float* new_xy, int K,
                             uint8_t* adj);
/* The same two calls for float64 coordinates: rejectOutliers / cdist run in float64 on whatever dtype the caller holds
 * (outlierRejection.py:49-58; metric or undistorted points are float64), so nothing is rounded to float32 first. */
RF_API int rf_reject_outliers_f64(rf_handle* h, const double* prev_xy, const double* new_xy, int K,
                           uint8_t* mask, int* n_inliers, int* nodes_or_null);
RF_API int rf_consistency_adjacency_f64(rf_handle* h, const double* prev_xy, const double* new_xy, int K,
                                 uint8_t* adj);

/* Test hook: the clique search alone on a caller-supplied adjacency matrix (K x K bytes,
 * diagonal ignored).  prune == 0 enumerates EVERY maximal clique in networkx order and
 * returns their count plus an order-sensitive FNV-1a hash over (size, members...) of each
 * yield, so tests can pin the enumeration ORDER against the oracle.  prune | 2 runs the
 * production configuration (no hash, clique shortcut enabled) and reports only mask/size/nodes.
 * mask_out: K int32. */
RF_API int rf_clique_search(rf_handle* h, const uint8_t* adj, int K, int prune, int32_t* mask_out, int* size,
                     int64_t* n_yields, uint64_t* order_hash, int64_t* nodes);

/* ---- a7  getTransformKLT.calculateTransformSVD           getTransformKLT.py:129-162 */
/* src = R * tgt + h (pixel units; Tracker.getTransform scales h by cart_res_m). */
RF_API int rf_kabsch(rf_handle* h, const float* src_xy, const float* tgt_xy, int N, double R[4], double hvec[2]);
/* float64 coordinates: means, centring and the cross-covariance in float64, as NumPy computes them for a float64 input
 * (getTransformKLT.py:141-150); rf_kabsch reproduces NumPy's float32 means for the float32 points cv2 returns. */
RF_API int rf_kabsch_f64(rf_handle* h, const double* src_xy, const double* tgt_xy, int N, double R[4], double hvec[2]);

/* ---- a8  MotionDistortionSolver.optimize_library         motionDistortion.py:80-325 */
/* sigma_p[2] / sigma_v[3] are the DIAGONALS the solver object was built with (the reference
 * weights residuals by 1/sigma, motionDistortion.py:96-99); NULL -> rf_config values;
 * period <= 0 -> rf_config.mds_period. */
RF_API int rf_mds_solve(rf_handle* h, const double T_wj0[9], const double* p_w, const double* p_jt, int N,
                 const double T_wj[9], const double* sigma_p, const double* sigma_v, double period,
                 double x_out[6], int* iters, double* cost);
/* static MotionDistortionSolver.undistort                  motionDistortion.py:127-153 */
RF_API int rf_mds_undistort(rf_handle* h, const double v[3], const double* pts_xy, int N, double period,
                     double* out_xy);
/* the same with explicit per-point time offsets: undistort(v_j, points, times=...)   motionDistortion.py:135-138 */
RF_API int rf_mds_undistort_times(rf_handle* h, const double v[3], const double* pts_xy, const double* times, int N,
                           double* out_xy);

/* ---- a9  ANMS.ssc                                        ANMS.py:5-102 ------------- */
/* kp [n,3] f64 (row, col, sigma) in caller order -> sel_idx [<= n] in selection order. */
RF_API int rf_ssc(rf_handle* h, const double* kp, int n, int num_ret, double tol, int cols, int rows,
           int32_t* sel_idx, int* m);

/* ---- a10 detector (response + 3x3 NMS + threshold)       getFeatures.py:22-53 ------ */
/* mode 0: min-eigenvalue structure tensor (Sobel 3x3, 3x3 box), cv2.cornerMinEigenVal
 * semantics on the f32 Cartesian image.  Candidates are sorted by (response desc, index desc):
 * out [cap,3] f64 (row, col, response).  *n receives the total candidate count.
 * threshold >= 0 is absolute; threshold < 0 means the fraction -threshold of the maximum response
 * (cv2.goodFeaturesToTrack's qualityLevel). */
/* mode 1: the reference's own detector, determinant-of-Hessian blobs = skimage.feature.blob_doh with the handle's
 * doh_* parameters (getFeatures.DEFAULT_FEATURE_PARAMS); threshold < 0 selects cfg.doh_threshold.
 * out [cap,3] f64 (row, col, sigma) — blob_doh's return value.  See rf_detect_doh. */
RF_API int rf_detect(rf_handle* h, const rf_frame* f, int mode, float threshold, double* out, int cap, int* n);
/* getFeatures.getBlobsFromCart(cartImage, min_sigma, max_sigma, num_sigma, threshold, method="doh")
 *                                                           getFeatures.py:22-53
 * = skimage.feature.blob_doh(img.astype(double), ...) on the frame's f32 Cartesian image: float64 integral image,
 * box-filter Hessian determinant per sigma of np.linspace(min_sigma, max_sigma, num_sigma) (num_sigma <= 16),
 * 3x3x3 peak_local_max above `threshold`, blobs ordered by descending response, _prune_blobs(overlap).
 * out [cap,3] f64 (row, col, sigma), *n = number of blobs.  scikit-image 0.19.2 is not part of the reference tree:
 * parity with it is UNPINNED; the algorithm and its three implementation-defined points (all-NaN plane of a
 * zero-size box, order of equal responses, pruning order) are written down in oracle/doh_restate.py. */
RF_API int rf_detect_doh(rf_handle* h, const rf_frame* f, double min_sigma, double max_sigma, int num_sigma,
                  double threshold, double overlap, double* out, int cap, int* n);
/* parity-test hook: the float64 integral image (sigma_index < 0) or the Hessian-determinant plane of one scale
 * (f64 [2R,2R]) */
RF_API int rf_doh_response(rf_handle* h, const rf_frame* f, double min_sigma, double max_sigma, int num_sigma,
                    int sigma_index, double* out);
/* response map only (f32 [2R,2R]) — used by the parity tests */
RF_API int rf_corner_response(rf_handle* h, const rf_frame* f, int mode, float* resp);
/* selection half of rf_detect on a caller-supplied response map (f32 [rows, cols]): interior
 * pixels with resp > threshold that equal their 3x3 maximum, ordered like
 * cv2.goodFeaturesToTrack orders its candidates (response descending, then pixel index
 * descending).  out [cap,3] f64 (row, col, response); *n = total candidate count.  Bit-exact. */
RF_API int rf_nms_select(rf_handle* h, const float* resp, int rows, int cols, float threshold, double* out, int cap,
                  int* n);

/* ---- a12 getPointCloud.getPointCloudPolarInd             getPointCloud.py:11-54 ---- */
/* polar f32 [A, W] -> out [cap,2] i64 (azimuth, range), azimuth-major order. */
RF_API int rf_polar_peaks(rf_handle* h, const float* polar, int A, int W, int64_t* out, int64_t cap, int64_t* n);

/* ---- N1  Fourier-Mellin rotation prior                   FMT.py:13-90 --------------- */
/* FMT.getRotationUsingFMT for n_pairs pairs drawn from n_frames f32 polar images [n_frames, A, W] (host):
 * clip to clip_px range bins (<= 0: none; FMT.py:55-58), cv2.resize to clip_px / downsample bins, log-polar
 * resampling (parseData.convertPolarImgToLogPolar, parseData.py:138-157), Hann-windowed phase correlation.
 * Outputs per pair: angle_rad (normalised, R(angle) @ src = target), scale, response (NULL to skip) and the raw
 * phaseCorrelate shift (dx, dy) in shift_xy [n_pairs, 2] (NULL to skip). */
RF_API int rf_fmt_rotation(rf_handle* h, const float* polar, int n_frames, int A, int W, const int32_t* pair_idx,
                    int n_pairs, int downsample, int clip_px, double* angle_rad, double* scale, double* response,
                    double* shift_xy);
/* The same with one pointer per image ([A, W] f32 each): the reference passes two separately allocated arrays. */
RF_API int rf_fmt_rotation_frames(rf_handle* h, const float* const* frames, int n_frames, int A, int W,
                           const int32_t* pair_idx, int n_pairs, int downsample, int clip_px, double* angle_rad,
                           double* scale, double* response, double* shift_xy);
/* The log-polar image itself (stage-level parity): out [h_lp, w_lp] f32; out == NULL only queries the size. */
RF_API int rf_fmt_log_polar(rf_handle* h, const float* polar, int A, int W, int downsample, int clip_px, float* out,
                     int64_t out_cap, int* h_lp, int* w_lp);
/* parseData.convertCartesianImageToPolar(imgCart, logPolarMode, shapeHW)                parseData.py:69-97
 * cv2.warpPolar forward map of a square f32 image [n, n] (host), linear or semi-log radius, INTER_LINEAR with
 * outliers filled with 0.  rows_out / cols_out <= 0 select cv2's default size (round(n / 2 * pi), round(n / 2));
 * out [rows, cols] f32 (host); out == NULL only queries the size. */
RF_API int rf_cart_to_polar(rf_handle* h, const float* cart, int n, int log_mode, int rows_out, int cols_out, float* out,
                     int64_t out_cap, int* rows, int* cols);
/* cv2.phaseCorrelate(a, b, cv2.createHanningWindow((cols, rows), CV_32F))   FMT.py:13-33
 * a, b [rows, cols] f32 (host) -> (dx, dy), response. */
RF_API int rf_phase_correlate(rf_handle* h, const float* a, const float* b, int rows, int cols, double* dx, double* dy,
                       double* response);

/* The same for every pair of an uploaded batch, from the scans already resident on the device (u8 power bins,
 * decoded as parseData.py:43 does): Tracker.py:62-63 without a second upload.  Synchronous on return. */
RF_API int rf_batch_fmt(rf_handle* h, rf_batch* b, int downsample, int clip_px, double* angle_rad, double* scale,
                 double* response, double* shift_xy);

/* ---- N3  trajectory chaining                              trajectoryPlotting.py:27-60 */
/* Poses [P+1, 3] (x, y, theta) from per-pair relative transforms R [P, 4] (row-major 2x2), h [P, 2], as a parallel
 * prefix product over SE(2), starting at start_pose (x, y, theta; NULL = origin).
 * left_multiply = 1: T_k = A_k T_{k-1} (Trajectory.appendRelativeTransform, trajectoryPlotting.py:54);
 * left_multiply = 0: T_k = T_{k-1} A_k (Trajectory.appendRelativeDeltas, trajectoryPlotting.py:27-35). */
RF_API int rf_chain_poses(rf_handle* h, const double* R, const double* hv, int P, const double* start_pose,
                   int left_multiply, double* poses_out);

/* ---- N4  raw-scan ingest                                  parseData.py:160-179 ------ */
/* Parallel host-side decode of the radar scan container (8-bit grayscale non-interlaced PNG) — what
 * cv2.imread(path, cv2.IMREAD_GRAYSCALE) does one file at a time.  n files -> out [n, rows, cols] u8 (typically a
 * pinned rf_host_alloc buffer handed to rf_batch_upload_async); threads <= 0 uses every host core.  No handle:
 * errors are reported through rf_last_error(NULL). */
RF_API int rf_png_info(const char* path, int* rows, int* cols);
RF_API int rf_ingest_png(const char* const* paths, int n, int rows, int cols, int threads, uint8_t* out);

/* ---- a11 fused pair / batch: Tracker.track + getTransform (+ MDS)  Tracker.py:35-127 */
/* One pair from two separately held raw scans [A, raw_width] u8: the fused path above on a two-frame batch the
 * handle owns.  feats_xy [K, 2] f32 (K <= max_features); prev_pose (x, y, theta) or NULL; next_xy [K, 2] and
 * corr_status [K] (Tracker.track's corrStatus) may be NULL.  Synchronous on return. */
RF_API int rf_track_pair(rf_handle* h, const uint8_t* raw_prev, const uint8_t* raw_next, const float* feats_xy, int K,
                  const double* prev_pose, int with_mds, rf_pair_result* result, float* next_xy, uint8_t* corr_status);
RF_API int rf_batch_create(rf_handle* h, rf_batch** out);
RF_API void rf_batch_destroy(rf_handle* h, rf_batch* b);
/* Host -> device staging.  raw: [n_frames, A, raw_width] u8; pair_idx: [n_pairs,2]
 * (prev frame, next frame) indices into raw; feats: [n_pairs, max_features, 2] f32 with
 * feat_counts[n_pairs]; prev_pose: [n_pairs,3] (x,y,theta) or NULL for identity. */
RF_API int rf_batch_upload(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx,
                    int n_pairs, const float* feats, const int32_t* feat_counts, const double* prev_pose);
/* Same without the trailing synchronisation: the host buffers must stay valid (and should
 * be pinned, see rf_host_alloc) until rf_sync returns. */
RF_API int rf_batch_upload_async(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx,
                          int n_pairs, const float* feats, const int32_t* feat_counts, const double* prev_pose);
/* Device-only: every stage for every pair, no host synchronisation before return.
 * Pipelining: uploads run on a copy stream, the image + KLT stages on the handle's stream and the
 * rejection / solve stages (and the download) on a per-batch tail stream, chained by events; with two
 * batches used alternately on one handle the upload of one overlaps the kernels of the other.  An
 * upload into a batch waits (on the device) until the previous run of that batch has consumed its
 * inputs. */
RF_API int rf_batch_run_async(rf_handle* h, rf_batch* b, int with_mds);
/* Block until everything queued on the handle (all its streams, all its batches) has finished. */
RF_API int rf_sync(rf_handle* h);
/* Block until the last run + download queued for this batch has finished (other batches keep going). */
RF_API int rf_batch_wait(rf_handle* h, rf_batch* b);
/* Device -> host: results [n_pairs]; next_xy [n_pairs,max_features,2] and status
 * [n_pairs,max_features] may be NULL. */
RF_API int rf_batch_download(rf_handle* h, rf_batch* b, rf_pair_result* results, float* next_xy, uint8_t* status);
RF_API int rf_batch_download_async(rf_handle* h, rf_batch* b, rf_pair_result* results, float* next_xy, uint8_t* status);
/* KLT-stage outputs of the last run: status [n_pairs,max_features] = cv2 status & (err < thr)
 * (getTransformKLT.py:365) before the clique fix-up, and err. Either may be NULL. */
RF_API int rf_batch_klt_status(rf_handle* h, rf_batch* b, uint8_t* klt_status, float* err);
/* One frame of the batch back to the host (what: 0 = f32 cart, needs write_cart_f32;
 * 1+l = u8 pyramid level l). */
RF_API int rf_batch_frame_download(rf_handle* h, const rf_batch* b, int frame, int what, void* out, int* rows, int* cols);
/* Per-stage CUDA-event timing on the handle's stream.  While enabled, every
 * rf_batch_run_async records events at its stage boundaries (ring of 64 runs);
 * rf_batch_stage_times synchronises, sums the elapsed ms of the runs recorded since the last
 * call into ms_sum[9] = {polar->cart | frame interleave, scan -> level 0 + 1, pyrDown of the remaining
 * levels, klt, compact, reject, kabsch, mds, finish} and resets the ring.  Stages of different batches
 * overlap when batches are pipelined, so the sums describe kernels, not the critical path. */
RF_API int rf_batch_set_profiling(rf_handle* h, rf_batch* b, int on);
RF_API int rf_batch_stage_times(rf_handle* h, rf_batch* b, float* ms_sum, int cap, int* n_runs);
/* Pinned (page-locked) host memory for the *_async entry points. */
RF_API int rf_host_alloc(size_t bytes, void** out);
RF_API void rf_host_free(void* p);
/* upload + run + download in one call (what Tracker.track/getTransform amount to). */
RF_API int rf_track_batch(rf_handle* h, rf_batch* b, const uint8_t* raw, int n_frames, const int32_t* pair_idx,
                   int n_pairs, const float* feats, const int32_t* feat_counts, const double* prev_pose,
                   int with_mds, rf_pair_result* results, float* next_xy, uint8_t* status);

/* ---- N2  chained odometry on the device: RawROAMSystem.run          RawROAMSystem.py:141-300
 *          + Keyframe / Map bookkeeping                               Mapping.py:37-66,97-125,149-174
 *          + appendNewFeatures / adaptiveNMS                          getFeatures.py:66-118
 * n_seq independent sequences advance one frame per step in lock step.  Everything the reference's loop carries
 * from frame k-1 to frame k stays on the device: the tracked features (blobCoord), the previous image pyramid,
 * prev_pose, the keyframe the features belong to (pose + undistorted local points, pruned with corrStatus), the
 * keyframe count.  A step is: scan -> Cartesian u8 pyramid, KLT from the previous frame, err gating, clique
 * rejection, Kabsch, motion-distortion solve (world points from the keyframe), pose / keyframe update, and — for
 * exactly the sequences whose feature count fell to <= 60 (RawROAMSystem.py:250-271) — re-detection on the device:
 * detector response, 3x3 NMS, candidate sort, the whole SSC bisection, append + order-preserving de-duplication.
 * No host synchronisation inside a step; with RF_SEQ_GRAPH the step is one CUDA graph launch.
 *
 * Scans live in an arena of `arena_frames` device-resident scans; a step reads scan (base + s * stride) for
 * sequence s, so both "one long drive, sequence s starts at frame s" (stride 1) and "one slot of n_seq fresh scans
 * per step" layouts work without copies. */
typedef struct rf_seq rf_seq;
typedef struct rf_seq_result {
    double pose[3];        /* absolute pose (x, y, theta) after this frame      RawROAMSystem.py:212,237           */
    double R[4];           /* relative transform T_wj0^-1 T(pose), row-major     RawROAMSystem.py:214               */
    double h[2];           /* metres                                                                                */
    double mds_x[6];       /* [vx, vy, vtheta, Tx, Ty, Ttheta]                   motionDistortion.py:295-325        */
    double kab_R[4];       /* Tracker.getTransform(good_old, good_new)           Tracker.py:108-127 (h in metres)   */
    double kab_h[2];
    int32_t n_features_in; /* len(blobCoord) handed to Tracker.track                                                */
    int32_t n_good;        /* after KLT status & (err < thr)                                                        */
    int32_t n_tracked;     /* clique size = good_new.shape[0]                    RawROAMSystem.py:249               */
    int32_t retrack;       /* n_tracked <= 60 -> appendNewFeatures               RawROAMSystem.py:250,264           */
    int32_t keyframe_added;/* retrack or Map.isGoodKeyframe                      RawROAMSystem.py:251-271           */
    int32_t n_keyframes;   /* len(map.keyframes) after this frame                                                   */
    int32_t n_features_out;/* len(blobCoord) carried to the next frame                                              */
    int32_t n_candidates;  /* detector candidates of this frame's re-detection (0 if none)                          */
    int32_t mds_iters;
    int32_t clique_nodes;
    int32_t status;        /* RF_OK, RF_E_WORKLIMIT (clique), RF_E_CAPACITY (features / candidates truncated)       */
    int32_t reserved;
} rf_seq_result;

#define RF_SEQ_MDS 1      /* pose from the motion-distortion solve (configs[2], what the reference always does);
                             without it the pose is T_wj = prev_pose @ [R, h] (configs[1]) and velocity = 0        */
#define RF_SEQ_GRAPH 2    /* replay the step as a CUDA graph (captured on first use per step shape)               */

/* detector_mode: 0 = structure-tensor min eigenvalue, 1 = determinant of Hessian (rf_detect).  The runner owns a
 * CUDA stream: several runners on one handle overlap (the latency-bound clique / solve tail of one under the image
 * kernels of another). */
RF_API int rf_seq_create(rf_handle* h, int n_seq, int arena_frames, int detector_mode, rf_seq** out);
RF_API void rf_seq_destroy(rf_handle* h, rf_seq* s);
/* raw [n_frames][A][raw_width] u8 (host; pinned for a truly asynchronous copy) -> arena frames [first, first + n_frames).
 * Waits (on the device) for queued steps that still read those arena frames. */
RF_API int rf_seq_upload_async(rf_handle* h, rf_seq* s, int first_frame, int n_frames, const uint8_t* raw);
/* Frame 0 of every sequence: pyramid, feature detection, first keyframe.  init_pose [n_seq][3] or NULL (origin). */
RF_API int rf_seq_reset_async(rf_handle* h, rf_seq* s, int base, int stride, const double* init_pose);
/* One frame for every sequence.  Asynchronous; results land in a device ring (rf_seq_results). */
RF_API int rf_seq_step_async(rf_handle* h, rf_seq* s, int base, int stride, int flags);
/* Results of step `step` (1 = first step after the reset; 0 = the reset itself: feature counts only), which must be
 * one of the last rf_seq_ring() steps.  Synchronises the runner's stream.  out [n_seq]. */
RF_API int rf_seq_results(rf_handle* h, rf_seq* s, int step, rf_seq_result* out);
RF_API int rf_seq_results_async(rf_handle* h, rf_seq* s, int step, rf_seq_result* out_pinned);
RF_API int rf_seq_ring(const rf_seq* s);
RF_API int rf_seq_steps_done(const rf_seq* s);
/* Current blobCoord of every sequence: feats [n_seq][max_features][2] f32, counts [n_seq].  Synchronises. */
RF_API int rf_seq_features(rf_handle* h, rf_seq* s, float* feats, int32_t* counts);
RF_API int rf_seq_sync(rf_handle* h, rf_seq* s);
/* Number of kernels one step launches (eager or inside the graph). */
RF_API int rf_seq_launches_per_step(const rf_seq* s);

#ifdef __cplusplus
}
#endif
#endif /* RADARFE_H */
