// h2d_probe_multi.cu — aggregate pinned-H2D ceiling of ONE box with N GPUs copying at the same time (tools only).
// One host thread per device, each with its own pinned buffer of 256 scans (400 x 3779 bytes), all released by one
// barrier; per device and aggregate GB/s of (1) flat copies and (2) the 2-D used-column copies rf_batch_upload
// issues (2008 of 3779 bytes per row).  Answers "is the e2e arm of bench.py at the box's ceiling at N GPUs?".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread tools/h2d_probe_multi.cu -o gpurun_out/h2d_probe_multi
//   gpurun_out/h2d_probe_multi [n_devices]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

static const int A = 400, RAW = 3779, USED = 2008, F = 256, REPS = 12;
static pthread_barrier_t g_bar;
struct Job { int dev; double flat_gbs, d2_gbs; int ok; };

static void* run(void* arg) {
    Job* j = (Job*)arg;
    j->ok = 0;
    if (cudaSetDevice(j->dev) != cudaSuccess) { pthread_barrier_wait(&g_bar); pthread_barrier_wait(&g_bar); pthread_barrier_wait(&g_bar); pthread_barrier_wait(&g_bar); return nullptr; }
    const size_t rows = (size_t)F * A, host_bytes = rows * RAW, dpitch = 2016, dev_bytes = rows * dpitch;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t st;
    cudaMallocHost(&h, host_bytes); cudaMalloc(&d, host_bytes > dev_bytes ? host_bytes : dev_bytes); cudaStreamCreate(&st);
    memset(h, 1, host_bytes);
    cudaMemcpyAsync(d, h, host_bytes, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
    // flat
    pthread_barrier_wait(&g_bar);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < REPS; ++r) cudaMemcpyAsync(d, h, host_bytes, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    j->flat_gbs = (double)host_bytes * REPS / s / 1e9;
    pthread_barrier_wait(&g_bar);
    // 2-D used columns
    cudaMemcpy2DAsync(d, dpitch, h, RAW, USED, rows, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
    pthread_barrier_wait(&g_bar);
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < REPS; ++r) cudaMemcpy2DAsync(d, dpitch, h, RAW, USED, rows, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    j->d2_gbs = (double)USED * rows * REPS / s / 1e9;
    pthread_barrier_wait(&g_bar);
    j->ok = cudaGetLastError() == cudaSuccess;
    cudaFreeHost(h); cudaFree(d); cudaStreamDestroy(st);
    return nullptr;
}

int main(int argc, char** argv) {
    int n = 0;
    cudaGetDeviceCount(&n);
    if (argc > 1 && atoi(argv[1]) > 0 && atoi(argv[1]) < n) n = atoi(argv[1]);
    if (n < 1) { printf("{\"error\": \"no device\"}\n"); return 1; }
    pthread_barrier_init(&g_bar, nullptr, n);
    pthread_t th[64]; Job jobs[64];
    for (int i = 0; i < n; ++i) { jobs[i].dev = i; pthread_create(&th[i], nullptr, run, &jobs[i]); }
    double flat = 0, d2 = 0;
    for (int i = 0; i < n; ++i) { pthread_join(th[i], nullptr); flat += jobs[i].flat_gbs; d2 += jobs[i].d2_gbs; }
    printf("{\"devices\": %d, \"flat_gbs_total\": %.1f, \"used_columns_2d_gbs_total\": %.1f, \"per_device_flat\": [", n, flat, d2);
    for (int i = 0; i < n; ++i) printf("%s%.1f", i ? ", " : "", jobs[i].flat_gbs);
    printf("], \"per_device_2d\": [");
    for (int i = 0; i < n; ++i) printf("%s%.1f", i ? ", " : "", jobs[i].d2_gbs);
    printf("], \"frames_per_s_ceiling_2d\": %.0f, \"bytes_per_frame\": %d}\n", d2 * 1e9 / ((double)USED * A), USED * A);
    return 0;
}
