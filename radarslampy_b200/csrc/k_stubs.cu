// TEMPORARY: entry points not implemented yet (removed as the stages land).
#include "common.cuh"
#define NI(h) rf_fail((h), RF_E_BADARG, "%s: not implemented yet", __func__)
extern "C" {
int rf_ssc(rf_handle* h, const double*, int, int, double, int, int, int32_t*, int*) { return NI(h); }
int rf_detect(rf_handle* h, const rf_frame*, int, float, double*, int, int*) { return NI(h); }
int rf_corner_response(rf_handle* h, const rf_frame*, int, float*) { return NI(h); }
int rf_polar_peaks(rf_handle* h, const float*, int, int, int64_t*, int64_t, int64_t*) { return NI(h); }
int rf_batch_create(rf_handle* h, rf_batch**) { return NI(h); }
void rf_batch_destroy(rf_handle*, rf_batch*) {}
int rf_batch_upload(rf_handle* h, rf_batch*, const uint8_t*, int, const int32_t*, int, const float*, const int32_t*, const double*) { return NI(h); }
int rf_batch_run_async(rf_handle* h, rf_batch*, int) { return NI(h); }
int rf_batch_download(rf_handle* h, rf_batch*, rf_pair_result*, float*, uint8_t*) { return NI(h); }
int rf_track_batch(rf_handle* h, rf_batch*, const uint8_t*, int, const int32_t*, int, const float*, const int32_t*, const double*, int, rf_pair_result*, float*, uint8_t*) { return NI(h); }
}
