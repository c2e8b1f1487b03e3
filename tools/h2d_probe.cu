// h2d_probe.cu — PCIe staging probe for the raw-scan upload (tools only; not part of libradarfe.so).
// Measures H2D GB/s of (1) a flat copy, (2) the 2-D copy rf_batch_upload issues (1997-byte rows out of
// 3779-byte scan rows), (3) a zero-copy kernel reading the same rows from mapped pinned memory.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/h2d_probe.cu -o gpurun_out/h2d_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// one warp per scan row: aligned 16-byte loads over [row + 11, row + 11 + W), written to a pitched device row
__global__ void __launch_bounds__(256) k_zero_copy(const uint8_t* __restrict__ src, int pitch, int col0, int W, int rows,
                                                   uint8_t* __restrict__ dst, int dpitch) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        const uint8_t* p = src + (size_t)r * pitch + col0;
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15;
        const int shift = (int)(reinterpret_cast<uintptr_t>(p) - a0);
        const uint4* ap = reinterpret_cast<const uint4*>(a0);
        const int nvec = (shift + W + 15) >> 4;
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)r * dpitch);
        for (int i = lane; i < nvec; i += 32) {
            uint4 v = ap[i];
            d[i] = v;   // (realignment by `shift` would happen in the consumer; the probe measures the link)
        }
    }
}

int main() {
    const int A = 400, PITCH = 3779, W = 1997, F = 256, DP = 2016 + 16;
    const size_t rows = (size_t)A * F, src_bytes = rows * PITCH, used = rows * W;
    uint8_t *h, *d, *d2;
    CK(cudaHostAlloc((void**)&h, src_bytes + 64, cudaHostAllocMapped));
    memset(h, 7, src_bytes + 64);
    CK(cudaMalloc((void**)&d, src_bytes));
    CK(cudaMalloc((void**)&d2, rows * DP + 64));
    uint8_t* hd; CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, s); CK(cudaMemcpyAsync(d, h, used, cudaMemcpyHostToDevice, s)); cudaEventRecord(e1, s);
        CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1);
        printf("flat   %8.1f MB  %.3f ms  %.1f GB/s\n", used / 1e6, ms, used / ms / 1e6);
        cudaEventRecord(e0, s); CK(cudaMemcpyAsync(d, h, src_bytes, cudaMemcpyHostToDevice, s)); cudaEventRecord(e1, s);
        CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1);
        printf("flatall%8.1f MB  %.3f ms  %.1f GB/s\n", src_bytes / 1e6, ms, src_bytes / ms / 1e6);
        cudaEventRecord(e0, s);
        CK(cudaMemcpy2DAsync(d2, 2000, h + 11, PITCH, W, rows, cudaMemcpyHostToDevice, s));
        cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1);
        printf("2d     %8.1f MB  %.3f ms  %.1f GB/s (useful)\n", used / 1e6, ms, used / ms / 1e6);
        // per-frame 2-D copies (what a streaming caller would issue)
        cudaEventRecord(e0, s);
        for (int f = 0; f < F; ++f)
            CK(cudaMemcpy2DAsync(d2 + (size_t)f * A * 2000, 2000, h + (size_t)f * A * PITCH + 11, PITCH, W, A, cudaMemcpyHostToDevice, s));
        cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1);
        printf("2d/frm %8.1f MB  %.3f ms  %.1f GB/s (useful)\n", used / 1e6, ms, used / ms / 1e6);
        for (int blocks = 148; blocks <= 148 * 8; blocks *= 2) {
            cudaEventRecord(e0, s);
            k_zero_copy<<<blocks, 256, 0, s>>>(hd, PITCH, 11, W, (int)rows, d2, DP);
            cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1);
            printf("zcopy b=%4d %6.1f MB  %.3f ms  %.1f GB/s (useful)\n", blocks, used / 1e6, ms, used / ms / 1e6);
        }
    }
    return 0;
}
