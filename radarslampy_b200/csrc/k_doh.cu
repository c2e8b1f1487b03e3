// k_doh.cu — determinant-of-Hessian blob response (placeholder until the kernels land)
#include "detect.cuh"
int rf_doh_detect_host(rf_handle* h, const float* d_cart, int n, float threshold, double* out, int cap, int* n_out) {
    return rf_fail(h, RF_E_BADARG, "rf_detect: mode 1 not built yet");
}
int rf_launch_doh_candidates(rf_handle* h, const DetectWs& ws, const float* d_cart, size_t cart_stride, int n, void* d_doh_ws,
                             const int32_t* d_flags) {
    return rf_fail(h, RF_E_BADARG, "DoH mode not built yet");
}
size_t rf_doh_ws_bytes(const rf_handle* h, int S) { return 256; }
