// api.cu — the C ABI of libradarfe.so (include/radarfe.h): handle lifetime, host<->device
// staging and the per-stage entry points.  Kernels live in the k_*.cu files.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

thread_local std::string g_rf_err;

int rf_fail(rf_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_rf_err = buf;
    return code;
}

int rf_ensure_scratch(rf_handle* h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return RF_OK;
    if (h->d_scratch) { cudaStreamSynchronize(h->stream); cudaFree(h->d_scratch); h->d_scratch = nullptr; h->scratch_bytes = 0; }
    bytes = (bytes + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    RF_CUDA(h, cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    h->scratch_gen++;
    return RF_OK;
}

int rf_ensure_pinned(rf_handle* h, size_t bytes) {
    if (bytes <= h->pinned_bytes) return RF_OK;
    if (h->h_pinned) { cudaStreamSynchronize(h->stream); cudaFreeHost(h->h_pinned); h->h_pinned = nullptr; h->pinned_bytes = 0; }
    bytes = (bytes + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    RF_CUDA(h, cudaMallocHost(&h->h_pinned, bytes));
    h->pinned_bytes = bytes;
    return RF_OK;
}

extern "C" {

void rf_default_config(rf_config* c) {
    memset(c, 0, sizeof(*c));
    c->azimuths = 400;
    c->raw_width = 3779;
    c->meta_bytes = 11;
    c->range_bins = (int)(87.5 / 0.0432);  // 2025, parseData.py:49-51
    c->downsample = 2;
    c->max_features = 512;
    c->max_pairs = 256;
    c->max_frames = 512;
    c->klt_win = 15;
    c->klt_max_level = 3;
    c->klt_max_iters = 10;
    c->klt_eps = 0.03f;
    c->klt_min_eig = 1e-4f;
    c->klt_err_thr = 10.0f;
    c->cart_res_m = 0.0432 * 2;                 // parseData.py:9-13
    c->dist_thr_px = 0.5 / (0.0432 * 2);        // outlierRejection.py:10-11
    c->mds_period = 1.0 / 4;                    // motionDistortion.py:36
    c->mds_sigma_p[0] = 4; c->mds_sigma_p[1] = 4;  // RawROAMSystem.py:135-139
    c->mds_sigma_v[0] = 1; c->mds_sigma_v[1] = 1;
    c->mds_sigma_v[2] = (5 * M_PI / 180) * (5 * M_PI / 180);
    c->clique_node_limit = 4000000;
    c->write_cart_f32 = 1;
    c->retrack_threshold = 60;                  // getFeatures.py:57
    c->ssc_num_ret = 200; c->ssc_tolerance = 0.1;   // getFeatures.py:66
    c->kf_rot_thr = 0.2; c->kf_trans_thr = 2.0;     // Mapping.py:13-15
    c->detect_quality = 0.01;
    c->doh_min_sigma = 0.01; c->doh_max_sigma = 10; c->doh_num_sigma = 3; c->doh_threshold = 0.0005;   // getFeatures.py:13-18
}

int rf_version(void) { return RADARFE_VERSION; }

const char* rf_last_error(const rf_handle* h) { return h ? h->err.c_str() : g_rf_err.c_str(); }

int rf_create(const rf_config* cfg, int device, void* stream, rf_handle** out) {
    if (!cfg || !out) return rf_fail(nullptr, RF_E_BADARG, "rf_create: null argument");
    *out = nullptr;
    if (cfg->azimuths < 2 || cfg->range_bins < 4 || cfg->raw_width < cfg->meta_bytes + cfg->range_bins ||
        cfg->downsample < 1 || cfg->klt_win != 15 || cfg->klt_max_level < 0 || cfg->klt_max_level >= RF_MAX_LEVELS ||
        cfg->max_features < 1 || cfg->max_features > 2048 || cfg->max_pairs < 1 || cfg->max_frames < 2)
        return rf_fail(nullptr, RF_E_BADARG, "rf_create: unsupported configuration");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return rf_fail(nullptr, RF_E_CUDA, "rf_create: no CUDA device (%s) - libradarfe has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return rf_fail(nullptr, RF_E_BADARG, "rf_create: device %d out of range", device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return rf_fail(nullptr, RF_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return rf_fail(nullptr, RF_E_CUDA, "rf_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    rf_handle* h = new rf_handle();
    h->cfg = *cfg;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->R = cfg->downsample > 1 ? cfg->range_bins / cfg->downsample : cfg->range_bins;  // parseData.py:119-122
    h->n = 2 * h->R;
    h->map = nullptr; h->map2 = nullptr; h->d_raw = nullptr; h->d_polar = nullptr; h->d_polar_u8 = nullptr;
    h->d_scratch = nullptr; h->scratch_bytes = 0; h->scratch_gen = 0; h->h_pinned = nullptr; h->pinned_bytes = 0;
    h->launches = 0;
    h->pair_batch = nullptr;
    h->ev0 = h->ev1 = nullptr; h->stream_copy = nullptr; h->ev_copy = nullptr; h->stream = nullptr; h->owns_stream = false;
    int rc = RF_OK;
    auto bail = [&](int code) { rf_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(rf_fail(nullptr, RF_E_CUDA, "cudaSetDevice failed"));
    if (stream) { h->stream = (cudaStream_t)stream; h->owns_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
            return bail(rf_fail(nullptr, RF_E_CUDA, "cudaStreamCreate failed"));
        h->owns_stream = true;
    }
    if (cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking) != cudaSuccess)
        return bail(rf_fail(nullptr, RF_E_CUDA, "cudaStreamCreate failed"));
    if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess)
        return bail(rf_fail(nullptr, RF_E_CUDA, "cudaEventCreate failed"));
    if (cudaMalloc(&h->map, (size_t)h->n * h->n * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&h->map2, (size_t)h->n * h->n * sizeof(uint4)) != cudaSuccess ||
        cudaMalloc(&h->d_raw, (size_t)cfg->azimuths * cfg->raw_width) != cudaSuccess ||
        cudaMalloc(&h->d_polar, (size_t)cfg->azimuths * cfg->range_bins * sizeof(float)) != cudaSuccess)
        return bail(rf_fail(nullptr, RF_E_NOMEM, "rf_create: device allocation failed"));
    if ((rc = rf_launch_build_map(h)) != RF_OK || (rc = rf_launch_build_map2(h)) != RF_OK) { g_rf_err = h->err; return bail(rc); }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(rf_fail(nullptr, RF_E_CUDA, "map build failed"));
    *out = h;
    return RF_OK;
}

void rf_destroy(rf_handle* h) {
    RfDeviceGuard rf_guard_(h);
    if (!h) return;
    cudaSetDevice(h->device);
    rf_sync_all(h);
    if (h->pair_batch) { rf_batch_destroy(h, h->pair_batch); h->pair_batch = nullptr; }
    if (h->map) cudaFree(h->map);
    if (h->map2) cudaFree(h->map2);
    if (h->d_raw) cudaFree(h->d_raw);
    if (h->d_polar) cudaFree(h->d_polar);
    if (h->d_polar_u8) cudaFree(h->d_polar_u8);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->stream_copy) cudaStreamDestroy(h->stream_copy);
    if (h->owns_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int rf_cart_size(const rf_handle* h) { return h ? h->n : 0; }
void* rf_stream(const rf_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t rf_launch_count(const rf_handle* h) { return h ? h->launches : 0; }

int rf_timer_start(rf_handle* h) {
    RfDeviceGuard rf_guard_(h);
    if (!h) return RF_E_BADARG;
    int rc = rf_join_streams(h);
    if (rc) return rc;
    RF_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    return RF_OK;
}
int rf_timer_stop_ms(rf_handle* h, float* ms) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !ms) return RF_E_BADARG;
    int rc = rf_join_streams(h);   // the interval ends when the copy and tail streams have drained too
    if (rc) return rc;
    RF_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    RF_CUDA(h, cudaEventSynchronize(h->ev1));
    RF_CUDA(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return RF_OK;
}
int rf_sync(rf_handle* h) {
    RfDeviceGuard rf_guard_(h);
    if (!h) return RF_E_BADARG;
    return rf_sync_all(h);
}

// ---- a1 -------------------------------------------------------------------------------
int rf_extract_polar(rf_handle* h, const uint8_t* raw, float* polar, int64_t* timestamps, float* azimuths,
                     uint8_t* valid) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !raw) return rf_fail(h, RF_E_BADARG, "rf_extract_polar: null argument");
    const rf_config& c = h->cfg;
    if (polar) {
        size_t rb = (size_t)c.azimuths * c.raw_width;
        RF_CUDA(h, cudaMemcpyAsync(h->d_raw, raw, rb, cudaMemcpyHostToDevice, h->stream));
        int rc = rf_launch_extract(h, h->d_raw, h->d_polar);
        if (rc) return rc;
        RF_CUDA(h, cudaMemcpyAsync(polar, h->d_polar, (size_t)c.azimuths * c.range_bins * sizeof(float),
                                   cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    // 11 metadata bytes per azimuth (parseData.py:39-42): byte shuffling, done on the host side of the ABI
    if (c.meta_bytes >= 11) {
        for (int a = 0; a < c.azimuths; ++a) {
            const uint8_t* row = raw + (size_t)a * c.raw_width;
            if (timestamps) memcpy(&timestamps[a], row, 8);
            if (azimuths) {
                uint16_t enc; memcpy(&enc, row + 8, 2);
                azimuths[a] = (float)((double)enc / 5600.0 * 2 * M_PI);
            }
            if (valid) valid[a] = row[10] == 255;
        }
    }
    return RF_OK;
}

// ---- frames ---------------------------------------------------------------------------
int rf_frame_create(rf_handle* h, rf_frame** out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !out) return rf_fail(h, RF_E_BADARG, "rf_frame_create: null argument");
    rf_frame* f = new rf_frame();
    int rc = rf_frameset_alloc(h, &f->fs, 1, true);
    if (rc) { rf_frameset_free(&f->fs); delete f; return rc; }
    *out = f;
    return RF_OK;
}

void rf_frame_destroy(rf_handle* h, rf_frame* f) {
    RfDeviceGuard rf_guard_(h);
    if (!f) return;
    if (h) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
    rf_frameset_free(&f->fs);
    delete f;
}

int rf_polar_to_cart(rf_handle* h, const uint8_t* raw, const float* polar, rf_frame* frame, float* cart_out) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !frame || (!raw == !polar)) return rf_fail(h, RF_E_BADARG, "rf_polar_to_cart: need exactly one of raw/polar");
    const rf_config& c = h->cfg;
    int rc;
    if (raw) {
        RF_CUDA(h, cudaMemcpyAsync(h->d_raw, raw, (size_t)c.azimuths * c.raw_width, cudaMemcpyHostToDevice, h->stream));
        rc = rf_launch_polar2cart_u8(h, h->d_raw, 0, c.raw_width, c.meta_bytes, frame->fs, 0, 1, true);
    } else {
        RF_CUDA(h, cudaMemcpyAsync(h->d_polar, polar, (size_t)c.azimuths * c.range_bins * sizeof(float),
                                   cudaMemcpyHostToDevice, h->stream));
        rc = rf_launch_polar2cart_f32(h, h->d_polar, c.range_bins, frame->fs);
    }
    if (rc) return rc;
    if ((rc = rf_launch_pyramid(h, frame->fs, 0, 1))) return rc;
    if (cart_out)
        RF_CUDA(h, cudaMemcpyAsync(cart_out, frame->fs.cart, (size_t)h->n * h->n * sizeof(float), cudaMemcpyDeviceToHost,
                                   h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

int rf_frame_from_cart(rf_handle* h, const float* cart, int n, rf_frame* frame) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !cart || !frame) return rf_fail(h, RF_E_BADARG, "rf_frame_from_cart: null argument");
    if (n != h->n) return rf_fail(h, RF_E_BADARG, "rf_frame_from_cart: image is %d x %d, handle expects %d", n, n, h->n);
    RF_CUDA(h, cudaMemcpyAsync(frame->fs.cart, cart, (size_t)n * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    int rc = rf_launch_cart_to_u8(h, frame->fs);
    if (rc) return rc;
    if ((rc = rf_launch_pyramid(h, frame->fs, 0, 1))) return rc;
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

int rf_frame_download(rf_handle* h, const rf_frame* f, int what, void* out, int* rows, int* cols) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !f) return rf_fail(h, RF_E_BADARG, "rf_frame_download: null argument");
    if (what == 0) {
        if (rows) *rows = h->n;
        if (cols) *cols = h->n;
        if (out) RF_CUDA(h, cudaMemcpyAsync(out, f->fs.cart, (size_t)h->n * h->n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    } else {
        int l = what - 1;
        if (l < 0 || l >= f->fs.n_levels) return rf_fail(h, RF_E_BADARG, "rf_frame_download: no pyramid level %d", l);
        if (rows) *rows = f->fs.h[l];
        if (cols) *cols = f->fs.w[l];
        if (out) RF_CUDA(h, cudaMemcpy2DAsync(out, f->fs.w[l], f->fs.lvl[l], f->fs.pitch[l], f->fs.w[l], f->fs.h[l], cudaMemcpyDeviceToHost, h->stream));
    }
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

}  // extern "C"
