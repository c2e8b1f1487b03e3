// detect.cuh — device-side feature detection for a batch of frames (k_detect.cu): response map,
// 3x3 NMS candidates, in-kernel key sort, and the whole SSC bisection (ANMS.py:5-102) in one
// launch per batch with no host synchronisation.  Shared by the single-frame entry points
// (rf_detect, rf_ssc, rf_nms_select) and the lock-step sequence runner (k_seq.cu).
#pragma once
#include "common.cuh"

#define RF_SSC_MAX_CANDIDATES 65536   // getFeatures.MAX_CANDIDATES: the strongest candidates handed to adaptiveNMS

// Workspace of `S` independent detection problems (all device memory, sized once).
struct DetectWs {
    int S;                 // problems (frames)
    int rows, cols;        // response-map shape
    unsigned key_cap;      // candidate keys per problem (power of two)
    unsigned ssc_cap;      // keypoints per problem the SSC arrays hold
    unsigned cells_cap;    // cells of the global cover grid per problem
    float* resp; size_t resp_stride;       // [S][rows*cols]  (null when the caller supplies responses)
    unsigned* maxbits;                      // [S] bit pattern of the maximum response
    unsigned* count;                        // [S] candidates found (may exceed key_cap: overflow)
    unsigned long long* keys;               // [S][key_cap]
    double2* rc;                            // [S][ssc_cap]  (row, col) of the keypoints in priority order
    int32_t* n_kp;                          // [S] keypoints handed to SSC
    uint32_t* cell;                         // [S][ssc_cap]  linear cell index of every keypoint for the current width
    uint32_t* alive;                        // [S][2][ssc_cap]
    uint32_t* selmask;                      // [S][2][ssc_cap / 32]
    uint32_t* grid;                         // [S][cells_cap]  cover grid (kept all-EMPTY between passes)
    int32_t* sel_idx;                       // [S][ssc_cap]  selected keypoints in selection order
    int32_t* m;                             // [S] number selected
    int32_t* status;                        // [S] RF_OK / RF_E_BADARG (the reference raises) / RF_E_CAPACITY
    void* base; size_t bytes;
};

// CTAs per problem for the tile-walking kernels of a batch of S problems (some of which may be switched off by a device
// flag): enough to fill the machine 16 deep when every problem is live, never fewer than 64 nor more than the tiles
static inline int rf_tile_workers(const rf_handle* h, int ntiles, int S) {
    int w = (h->sm_count * 16 + S - 1) / (S > 0 ? S : 1);
    // In a lock-step step only a quarter of the problems is live, and a live problem's strips are walked by its own CTAs
    // only: with 64 of them each warp walked 8 strips row by row, one exposed load latency per row (k_min_eig / k_nms_select at
    // 1.2 TB/s).  256 per problem; the CTAs of a switched-off problem exit on their first instruction.
    w = w < 256 ? 256 : w;
    return w > ntiles ? ntiles : w;
}

size_t rf_detect_ws_bytes(int S, int rows, int cols, unsigned key_cap, unsigned ssc_cap, unsigned cells_cap, bool with_resp);
// carve a workspace out of device memory `base` (rf_detect_ws_bytes); grid cells are initialised by rf_detect_ws_init
DetectWs rf_detect_ws_carve(void* base, int S, int rows, int cols, unsigned key_cap, unsigned ssc_cap, unsigned cells_cap, bool with_resp);
int rf_detect_ws_init(rf_handle* h, const DetectWs& ws);
int rf_detect_prepare(rf_handle* h);   // kernel attributes (call once per device before any stream capture)

// response of mode 0 (structure-tensor minimum eigenvalue) for every flagged problem: img [S][n][n] f32
// d_maxbits (optional, [S] u32 cleared by the caller): the maximum response of every problem, accumulated while the
// response is written (saves the separate pass over the response plane a relative threshold would need)
int rf_launch_min_eig(rf_handle* h, const float* d_img, size_t img_stride, int n, float* d_resp, size_t resp_stride, int S,
                      const int32_t* d_flags, unsigned* d_maxbits = nullptr);
// clear the candidate counters and the response maxima of a workspace (first launch of a detection)
int rf_launch_detect_clear(rf_handle* h, const DetectWs& ws);
// threshold (absolute if >= 0, else -threshold x max response) + 3x3 NMS -> sortable keys, then sort them.
// have_max: ws.count / ws.maxbits were cleared by rf_launch_detect_clear and ws.maxbits filled by rf_launch_min_eig
int rf_launch_select_sorted(rf_handle* h, const DetectWs& ws, const float* d_resp, size_t resp_stride, float threshold,
                            const int32_t* d_flags, bool have_max = false);
// keys -> (row, col) keypoints (at most RF_SSC_MAX_CANDIDATES strongest), then the SSC bisection
int rf_launch_ssc_from_keys(rf_handle* h, const DetectWs& ws, int num_ret, double tol, const int32_t* d_flags);
// SSC bisection on keypoints already in ws.rc / ws.n_kp
int rf_launch_ssc(rf_handle* h, const DetectWs& ws, int num_ret, double tol, int cols, int rows, const int32_t* d_flags);
