// Legacy tensor-core path on sm_100a: issue rate of mma.sync m16n8k16 (bf16) and m16n8k8 (tf32) per SM.
// The gate kernels (skinny T x d x E contractions, E = 8..64) use this path; this prints what it can sustain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void __launch_bounds__(256) bf16_kernel(float* out, int iters) {
    float c[ILP][4];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    a[0] += threadIdx.x & 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) tf32_kernel(float* out, int iters) {
    float c[ILP][4];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run(const char* name, K kern, double flop_per_mma, int ilp, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * 256);
    const int iters = 20000;
    kern<<<sms * ctas_per_sm, 256>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<sms * ctas_per_sm, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = double(sms) * ctas_per_sm * 8 * iters * ilp;
    printf("%-28s ilp %d, %d CTA/SM: %.3f ms  %.1f TFLOP/s  (%.2f ns per MMA per SM)\n", name, ilp, ctas_per_sm, ms,
           mmas * flop_per_mma / (ms * 1e-3) / 1e12, ms * 1e6 / (mmas / sms));
    cudaFree(out);
}

int main() {
    run("mma.sync m16n8k16 bf16", bf16_kernel<4>, 2.0 * 16 * 8 * 16, 4, 1);
    run("mma.sync m16n8k16 bf16", bf16_kernel<8>, 2.0 * 16 * 8 * 16, 8, 1);
    run("mma.sync m16n8k16 bf16", bf16_kernel<8>, 2.0 * 16 * 8 * 16, 8, 2);
    run("mma.sync m16n8k8 tf32", tf32_kernel<8>, 2.0 * 16 * 8 * 8, 8, 1);
    run("mma.sync m16n8k8 tf32", tf32_kernel<8>, 2.0 * 16 * 8 * 8, 8, 2);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
