// gemm.cuh — grouped expert GEMM on tcgen05 / TMEM, fed by TMA (sm_100a only).
//
// One persistent, warp-specialised kernel template covers every dense contraction of the
// expert FFN (replaces FastMoE's per-expert cuBLAS loop
// `fmoe_cuda.linear_forward/backward`, reached from /root/reference/models/resMoE.py:27-29):
//
//   ROWS mode  (M = packed token rows, one weight matrix per 128-row tile):
//     fc1   : U = X  W1^T + b1, H = gelu_erf(U)     A K-major, B K-major, EPI_BIAS_GELU_DUAL
//     fc2   : Y = H  W2^T + b2                      A K-major, B K-major, EPI_BIAS
//     dgelu : dU = (dY W2) * gelu'(U)               A K-major, B MN-major, EPI_DGELU
//     dgrad : dX = dU W1                            A K-major, B MN-major, EPI_PLAIN
//   WGRAD mode (K = the rows of one expert's segment, fp32 output per expert):
//     dW2[e] = dY_e^T H_e , dW1[e] = dU_e^T X_e     A MN-major, B MN-major, EPI_F32
//
// Layout contract: the packed row buffers are [rows_cap, cols] bf16, every expert's segment
// starts at a multiple of 128 rows (so a 128-row tile never straddles two experts) and pad
// rows are zero in X and dY (so they contribute nothing to WGRAD).
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  Pipelines: smem ring full/empty
// (TMA <-> MMA), two TMEM accumulator stages full/empty (MMA <-> epilogue), epilogue staging
// ring drained by TMA stores.  Tile = 128 x BN x 64, UMMA 128 x BN x 16, cta_group::1.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace moe {

enum : int { EPI_BIAS_GELU_DUAL = 0, EPI_BIAS = 1, EPI_DGELU = 2, EPI_PLAIN = 3, EPI_F32 = 4 };

struct GemmParams {
    const int* tile_expert;  // ROWS: expert of each 128-row tile            [max_mtiles]
    const int* num_mtiles;   // ROWS: number of live 128-row tiles (device scalar)
    const int* seg_start;    // WGRAD: first row of each expert segment      [E+1]
    const float* bias;       // [E, N] fp32 or nullptr
    const __nv_bfloat16* aux;  // EPI_DGELU: pre-activation U [rows_cap, N]
    int E;
    int M;  // WGRAD: output rows per expert
    int N;  // output columns (per expert)
    int K;  // ROWS: reduction length
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kStagingBytes = 16384;  // 128 rows x 128 B
constexpr int kSmemLimit = 232448;    // 227 KB

template <int BN, int EPI>
struct GemmCfg {
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = BN * kBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NOUT = (EPI == EPI_BIAS_GELU_DUAL) ? 2 : 1;
    static constexpr int NBUF = 2 * NOUT;
    static constexpr int CHUNK_COLS = (EPI == EPI_F32) ? 32 : 64;
    static constexpr int NCHUNK = BN / CHUNK_COLS;
    static constexpr int BAR_BYTES = 256;
    static constexpr int STAGES_RAW = (kSmemLimit - 1024 - BAR_BYTES - NBUF * kStagingBytes) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + NBUF * kStagingBytes + BAR_BYTES;
    static_assert(STAGES >= 2, "not enough shared memory for a pipelined tile");
    static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64, at most 256");
};

__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad_f(float u) {
    float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
    float pdf = __expf(-0.5f * u * u) * 0.39894228040143268f;
    return cdf + u * pdf;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

struct TileCoord {
    int e;      // expert (weight index)
    int m0;     // ROWS: first packed row of the tile.  WGRAD: first output row inside the expert
    int n0;     // first output column
    int row0;   // WGRAD: first packed row of the expert segment
    int kb;     // number of 64-deep k-blocks
};

template <int BN, bool WGRAD>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int n_ntiles) {
    TileCoord c;
    if constexpr (!WGRAD) {
        int m = tile / n_ntiles;
        c.e = __ldg(p.tile_expert + m);
        c.m0 = m * kBM;
        c.n0 = (tile - m * n_ntiles) * BN;
        c.row0 = 0;
        c.kb = p.K / kBK;
    } else {
        int m_tiles = (p.M + kBM - 1) / kBM;
        int per_e = m_tiles * n_ntiles;
        c.e = tile / per_e;
        int rem = tile - c.e * per_e;
        int mt = rem / n_ntiles;
        c.m0 = mt * kBM;
        c.n0 = (rem - mt * n_ntiles) * BN;
        c.row0 = __ldg(p.seg_start + c.e);
        c.kb = (__ldg(p.seg_start + c.e + 1) - c.row0) / kBK;
    }
    return c;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool WGRAD>
__global__ void __launch_bounds__(kGemmThreads, 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                    const GemmParams p) {
    using Cfg = GemmCfg<BN, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(!WGRAD || (A_MN && B_MN && EPI == EPI_F32), "WGRAD = MN-major operands, fp32 output");
    static_assert(WGRAD || !A_MN, "ROWS mode reads the packed rows K-major");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::NBUF * kStagingBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO0);
        if constexpr (Cfg::NOUT == 2) tma_prefetch_desc(&tmO1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + s, 1);
            mbar_init(tempty_bar + s, 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_ntiles = p.N / BN;
    int total_tiles;
    if constexpr (WGRAD) total_tiles = p.E * ((p.M + kBM - 1) / kBM) * n_ntiles;
    else total_tiles = __ldg(p.num_mtiles) * n_ntiles;

    if (warp == 0) {
        // ================================ TMA producer (one thread) ================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles);
                for (int kb = 0; kb < c.kb; ++kb) {
                    mbar_wait(empty_bar + s, ph ^ 1);
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar + s, Cfg::STAGE_BYTES);
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tmA, full_bar + s, kb * kBK, c.m0);
                    } else {
                        const int krow = c.row0 + kb * kBK;
                        tma_load_2d(sa, &tmA, full_bar + s, c.m0, krow);
                        tma_load_2d(sa + 8192, &tmA, full_bar + s, c.m0 + 64, krow);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(sb, &tmB, full_bar + s, kb * kBK, c.e * p.N + c.n0);
                    } else {
                        const int krow = WGRAD ? (c.row0 + kb * kBK) : (c.e * p.K + kb * kBK);
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_2d(sb + i * 8192, &tmB, full_bar + s, c.n0 + i * 64, krow);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (one thread) ==================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN, B_MN);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles);
                if (c.kb == 0) continue;
                mbar_wait(tempty_bar + as, aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * 256;
                for (int kb = 0; kb < c.kb; ++kb) {
                    mbar_wait(full_bar + s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < kBK / 16; ++k4) {
                        const uint64_t ad = A_MN ? umma_smem_desc(a_addr + k4 * 2048, 8192, 1024)
                                                 : umma_smem_desc(a_addr + k4 * 32, 16, 1024);
                        const uint64_t bd = B_MN ? umma_smem_desc(b_addr + k4 * 2048, 8192, 1024)
                                                 : umma_smem_desc(b_addr + k4 * 32, 16, 1024);
                        umma_bf16(tmem_d, ad, bd, idesc, (kb | k4) != 0);
                    }
                    umma_commit(empty_bar + s);  // frees the smem slot once these MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar + as);  // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue (4 warps, 128 threads) ==========================
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;       // row inside the 128-row tile
        const int ep_tid = threadIdx.x - 64;
        uint8_t* my_row = staging + r * 128;
        const int sw = r & 7;
        int as = 0;
        uint32_t aph = 0;
        uint32_t step = 0;  // staging ring position (monotonic over the CTA's lifetime)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles);
            const bool live = c.kb != 0;
            if (live) {
                mbar_wait(tfull_bar + as, aph);
                tc_fence_after();
            }
            const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
            const float* bias = (EPI == EPI_BIAS_GELU_DUAL || EPI == EPI_BIAS) && p.bias != nullptr
                                    ? p.bias + static_cast<size_t>(c.e) * p.N + c.n0
                                    : nullptr;
#pragma unroll 1
            for (int ch = 0; ch < Cfg::NCHUNK; ++ch, ++step) {
                uint8_t* buf0 = my_row + ((step & 1) * Cfg::NOUT) * kStagingBytes;
                // staging buffers of ring slot (step & 1) were last used two steps ago
                if (ep_tid == 0) tma_store_wait_read<1>();
                named_bar_sync(1, 128);

                uint4 auxv[8];
                if constexpr (EPI == EPI_DGELU) {
                    const uint4* ap = reinterpret_cast<const uint4*>(
                        p.aux + static_cast<size_t>(c.m0 + r) * p.N + c.n0 + ch * 64);
#pragma unroll
                    for (int i = 0; i < 8; ++i) auxv[i] = __ldg(ap + i);
                }
#pragma unroll
                for (int half = 0; half < Cfg::CHUNK_COLS / 32; ++half) {
                    uint32_t acc[32];
                    if (live) {
                        tmem_ld32(tmem_row + ch * Cfg::CHUNK_COLS + half * 32, acc);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0u;
                    }
                    if constexpr (EPI == EPI_F32) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint4 v = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                            *reinterpret_cast<uint4*>(buf0 + ((j ^ sw) << 4)) = v;
                        }
                    } else {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
                        if (bias != nullptr) {
                            const float4* bp = reinterpret_cast<const float4*>(bias + ch * 64 + half * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float4 b = __ldg(bp + i);
                                v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                            }
                        }
                        if constexpr (EPI == EPI_DGELU) {
                            const uint32_t* aw = reinterpret_cast<const uint32_t*>(auxv) + half * 16;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                __nv_bfloat162 u2 = *reinterpret_cast<const __nv_bfloat162*>(aw + i);
                                v[2 * i] *= gelu_erf_grad_f(__low2float(u2));
                                v[2 * i + 1] *= gelu_erf_grad_f(__high2float(u2));
                            }
                        }
                        // first (or only) output: the linear result itself
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 o = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                            *reinterpret_cast<uint4*>(buf0 + (((half * 4 + j) ^ sw) << 4)) = o;
                        }
                        if constexpr (EPI == EPI_BIAS_GELU_DUAL) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 o = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                                *reinterpret_cast<uint4*>(buf0 + kStagingBytes + (((half * 4 + j) ^ sw) << 4)) = o;
                            }
                        }
                    }
                }
                if (live && ch == Cfg::NCHUNK - 1) {
                    // every TMEM read of this accumulator stage has completed -> hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar + as);
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (ep_tid == 0) {
                    const uint8_t* sbuf = staging + ((step & 1) * Cfg::NOUT) * kStagingBytes;
                    if constexpr (WGRAD) {
                        tma_store_3d(&tmO0, sbuf, c.n0 + ch * Cfg::CHUNK_COLS, c.m0, c.e);
                    } else {
                        tma_store_2d(&tmO0, sbuf, c.n0 + ch * Cfg::CHUNK_COLS, c.m0);
                        if constexpr (Cfg::NOUT == 2)
                            tma_store_2d(&tmO1, sbuf + kStagingBytes, c.n0 + ch * Cfg::CHUNK_COLS, c.m0);
                    }
                    tma_store_commit();
                }
            }
            if (live) {
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
        if (ep_tid == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace moe
