// TMEM read-bandwidth micro-benchmark (sm_100a): 8 warps (2 per lane quarter) read a 128 x 256 fp32
// accumulator region repeatedly with different tcgen05.ld widths.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_cols<8>(uint32_t taddr) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= r[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<16>(uint32_t taddr) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
    for (int i = 0; i < 16; ++i) s ^= r[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
    for (int i = 0; i < 32; ++i) s ^= r[i];
    return s;
}

template <int W>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, long long* cyc, int iters, int nwarps) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        const int half = warp >> 2;              // column half (128 columns each when 8 warps)
        const int ncol = nwarps > 4 ? 128 : 256;
        for (int it = 0; it < iters; ++it)
            for (int c = 0; c < ncol; c += W) acc ^= ld_cols<W>(base + half * 128 + c);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int W>
void run(int nwarps) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 200;
    k<W><<<148, 256>>>(out, cyc, iters, nwarps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double bytes = 128.0 * 256 * 4 * iters;
    printf("x%-2d warps=%d: %s  %lld cycles for %d reads of 128x256 fp32 -> %.1f cycles per accumulator, %.1f B/clk/SM\n", W, nwarps,
           cudaGetErrorString(e), h, iters, (double)h / iters, bytes / h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<8>(4); run<16>(4); run<32>(4);
    run<8>(8); run<16>(8); run<32>(8);
    return 0;
}
