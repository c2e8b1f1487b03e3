// TEMPORARY: entry points not implemented yet (removed as the stages land).
#include "common.cuh"
#define NI(h) rf_fail((h), RF_E_BADARG, "%s: not implemented yet", __func__)
extern "C" {
int rf_ssc(rf_handle* h, const double*, int, int, double, int, int, int32_t*, int*) { return NI(h); }
int rf_detect(rf_handle* h, const rf_frame*, int, float, double*, int, int*) { return NI(h); }
int rf_corner_response(rf_handle* h, const rf_frame*, int, float*) { return NI(h); }
int rf_polar_peaks(rf_handle* h, const float*, int, int, int64_t*, int64_t, int64_t*) { return NI(h); }
}
