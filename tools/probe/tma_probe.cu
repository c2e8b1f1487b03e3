// Stand-alone probe of the TMA box load used by k_pyr_down_tma (diagnostic tool, not part of the library).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int bytes, uint8_t* out, int* status) {
    __shared__ __align__(128) uint8_t buf[8192];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(d), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(d), "l"(&tm), "r"(b), "r"(c0), "r"(c1) : "memory");
    }
    __syncthreads();
    int spins = 0; uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        if (++spins > (1 << 20)) { if (threadIdx.x == 0) *status = -1; return; }
    }
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
    if (threadIdx.x == 0) *status = spins;
}

int main(int argc, char** argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 998, h = argc > 2 ? atoi(argv[2]) : 998, count = argc > 3 ? atoi(argv[3]) : 16;
    const int boxw = argc > 4 ? atoi(argv[4]) : 144, boxh = argc > 5 ? atoi(argv[5]) : 19, rank = argc > 6 ? atoi(argv[6]) : 3;
    const int c0 = argc > 7 ? atoi(argv[7]) : 382, c1 = argc > 8 ? atoi(argv[8]) : 750, c2 = argc > 9 ? atoi(argv[9]) : 4;
    const int pitch = (w + 15) & ~15;
    const size_t stride = ((size_t)pitch * h + 255) & ~(size_t)255;
    std::vector<uint8_t> img(stride * count);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d_img, *d_out; int* d_st;
    cudaMalloc(&d_img, img.size() + 256); cudaMalloc(&d_out, 8192); cudaMalloc(&d_st, 4);
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_st, 0, 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)count};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
    const cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  w=%d h=%d count=%d pitch=%d box=%dx%d rank=%d coord=(%d,%d,%d)\n", (int)r, w, h, count, pitch, boxw, boxh, rank, c0, c1, c2);
    if (r != CUDA_SUCCESS) return 1;
    const int bytes = boxw * boxh;
    if (rank == 3) k_probe<3><<<1, 128>>>(tm, c0, c1, c2, bytes, d_out, d_st);
    else k_probe<2><<<1, 128>>>(tm, c0, c1, c2, bytes, d_out, d_st);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; std::vector<uint8_t> out(8192);
    cudaMemcpy(&st, d_st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(out.data(), d_out, 8192, cudaMemcpyDeviceToHost);
    printf("sync: %s, status(spins)=%d\n", cudaGetErrorString(e), st);
    if (e != cudaSuccess) return 1;
    int bad = 0;
    for (int y = 0; y < boxh; ++y) for (int x = 0; x < boxw; ++x) {
        const int gx = c0 + x, gy = c1 + y;
        const uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[(rank == 3 ? c2 * stride : 0) + (size_t)gy * pitch + gx] : 0;
        if (out[y * boxw + x] != want) ++bad;
    }
    printf("mismatches: %d of %d\n", bad, bytes);
    return 0;
}
