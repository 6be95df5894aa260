// Does tcgen05.ld (epilogue TMEM reads) overlap tcgen05.mma (accumulator writes) on sm_100a?
//
// One CTA per SM, 320 threads.  Thread 32 issues `iters` "tiles" of NK x UMMA 128x256x16 (cta_group::1, bf16, smem
// operands = zeros) into alternating 256-column accumulator stages; warps 2..9 read a 128 x 256 fp32 accumulator
// per tile with a chosen tcgen05.ld width.  Modes: 0 = MMA only, 1 = LD only, 2 = both, unsynchronised (pure
// port-contention test).  Prints cycles per tile for each mode and ld shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_mma_ld tmem_mma_ld.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>
struct Ld;
template <>
struct Ld<16> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s ^= r[i];
        return s;
    }
};
template <>
struct Ld<32> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(t));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) s ^= r[i];
        return s;
    }
};
template <>
struct Ld<64> {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[64];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,"
            "%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
              "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
              "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
              "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
            : "r"(t));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) s ^= r[i];
        return s;
    }
};
// two x16 loads in flight per wait (what a software-pipelined epilogue would do)
struct Ld16x2 {
    static __device__ __forceinline__ uint32_t go(uint32_t t) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(t + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) s ^= r[i];
        return s;
    }
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {   // K-major, 128-byte swizzle, SBO = 1024
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1) << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

template <class L, int W>
__global__ void __launch_bounds__(320, 1) k(uint32_t* out, long long* cyc, int iters, int mode, int nk, int ld_warps) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot;
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp == 1) {
        if (lane == 0 && mode != 1) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a = smem_u32(smem), b = a + 16384;
            uint32_t ph[2] = {0, 0};
            for (int it = 0; it < iters; ++it) {
                const int s = it & 1;
                if (it >= 2) {   // the commit of tile it-2 (same stage) has fired
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(bar + s)), "r"(ph[s]) : "memory");
                    ph[s] ^= 1;
                }
                for (int kk = 0; kk < nk; ++kk) {
                    const uint64_t ad = smem_desc(a + (kk & 3) * 32), bd = smem_desc(b + (kk & 3) * 32);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                                     base + s * 256), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(kk != 0)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + s)) : "memory");
            }
            for (int s = 0; s < 2; ++s) {   // drain
                int n = (iters + 1 - s) / 2;  // commits on stage s
                if (n > 0) {
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(bar + s)), "r"(ph[s]) : "memory");
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 2 + ld_warps && mode != 0) {
        // ld_warps = 8: two warps per lane quarter, each half of the 256 columns; 4: one warp per quarter, all columns
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int ncol = ld_warps > 4 ? 128 : 256;
        const uint32_t tb = base + ((uint32_t)(q * 32) << 16) + half * ncol;
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = tb + (it & 1) * 256;
            for (int c = 0; c < ncol; c += W) acc ^= L::go(st + c);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) cyc[1 + warp] = clock64() - t0;
    out[blockIdx.x * 320 + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
    }
}

template <class L, int W>
void run(const char* name, int nk, int ld_warps) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 320 * 4); cudaMalloc(&cyc, 8 * 16);
    const int iters = 400, smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(k<L, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    double c[3];
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(cyc, 0, 8 * 16);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0);
        k<L, W><<<148, 320, smem>>>(out, cyc, iters, mode, nk, ld_warps);
        cudaError_t e = cudaGetLastError();
        cudaEventRecord(e1);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long hh[16]; cudaMemcpy(hh, cyc, 8 * 16, cudaMemcpyDeviceToHost);
        if (getenv("UB_DEBUG")) printf("   mode %d: %.3f ms; cyc thread0 %lld; per warp: %lld %lld %lld %lld %lld\n", mode, ms, hh[0], hh[1], hh[2], hh[3], hh[4], hh[10]);
        if (e != cudaSuccess) { printf("%s mode %d: %s\n", name, mode, cudaGetErrorString(e)); return; }
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        c[mode] = (double)h / iters;
    }
    printf("%-8s nk=%2d ld_warps=%d: cycles/tile  mma-only %7.1f   ld-only %7.1f (%.1f B/clk/SM)   both %7.1f   (sum %7.1f, max %7.1f)\n", name, nk,
           ld_warps, c[0], c[1], 128.0 * 256 * 4 / c[1], c[2], c[0] + c[1], c[0] > c[1] ? c[0] : c[1]);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int nk : {24, 96}) {
        run<Ld<16>, 16>("x16", nk, 8);
        run<Ld<32>, 32>("x32", nk, 8);
        run<Ld<64>, 64>("x64", nk, 8);
        run<Ld16x2, 32>("x16x2", nk, 8);
        run<Ld<32>, 32>("x32", nk, 4);
        run<Ld<64>, 64>("x64", nk, 4);
    }
    return 0;
}
