// gemm.cuh — grouped expert GEMM on tcgen05 / TMEM, fed by TMA, on CTA PAIRS (sm_100a only).
//
// One persistent, warp-specialised kernel template covers every dense contraction of the
// expert FFN (replaces FastMoE's per-expert cuBLAS loop
// `fmoe_cuda.linear_forward/backward`, reached from /root/reference/models/resMoE.py:27-29):
//
//   ROWS mode  (M = packed token rows, one weight matrix per 256-row pair tile; A and B K-major):
//     fc1   : U = X  W1^T + b1, H = gelu_erf(U)     B = W1   [E, h, d]      EPI_BIAS_GELU_DUAL
//     fc2   : Y = H  W2^T + b2                      B = W2   [E, d, h]      EPI_BIAS
//     dgelu : dU = (dY W2) * gelu'(U)               B = W2^T [E, h, d]      EPI_DGELU
//     dgrad : dX = dU W1                            B = W1^T [E, d, h]      EPI_PLAIN
//   WGRAD mode (K = the rows of one expert's segment, fp32 output per expert; A and B MN-major):
//     dW2[e] = dY_e^T H_e , dW1[e] = dU_e^T X_e                             EPI_F32
//
// Layout contract: the packed row buffers are [rows_cap, cols] bf16, every expert's segment
// starts at a multiple of 256 rows (so a 256-row pair tile never straddles two experts) and pad
// rows are zero in X and dY (so they contribute nothing to WGRAD).
//
// Why pairs: with one CTA per 128 x BN tile every SM pulls (128 + BN) x 64 x 2 bytes from L2 per
// 2 x 128 x BN x 64 flop — 0.012-0.013 B/flop, which at the ~6.3 KB/clk L2->SM ceiling caps the
// chip at ~0.9 PFLOP/s (measured round 1a: tensor pipe 53 % active, fc2 at 800 TFLOP/s).  A CTA
// pair (`tcgen05.mma.cta_group::2`, UMMA 256 x BN x 16) shares B: each CTA loads its own 128 A
// rows and HALF of the B tile, 0.0078 B/flop at BN = 256.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM owner (+ single-thread MMA issuer in
// the leader CTA), warps 2..9 = epilogue: two groups of four warps (one TMEM lane quarter each),
// the groups take alternate column chunks of the accumulator; every warp stages and TMA-stores its own
// 32-row slab, so the epilogue has no CTA-level barrier.  Pipelines: smem ring full/empty
// (TMA <-> MMA; `full` lives in the leader and is credited by both CTAs' TMA loads, `empty` is
// multicast to both CTAs by tcgen05.commit), two TMEM accumulator stages full/empty
// (MMA <-> epilogue of both CTAs), one store-staging buffer per group and output drained by TMA.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace moe {

enum : int { EPI_BIAS_GELU_DUAL = 0, EPI_BIAS = 1, EPI_DGELU = 2, EPI_PLAIN = 3, EPI_F32 = 4 };

struct GemmParams {
    const int* tile_expert;  // ROWS: expert of each 256-row pair tile       [max_mtiles]
    const int* num_mtiles;   // ROWS: number of live 256-row tiles (device scalar)
    const int* seg_start;    // WGRAD: first row of each expert segment      [E+1]
    const float* bias;       // [E, N] fp32 or nullptr
    const __nv_bfloat16* aux;  // EPI_DGELU: pre-activation U [rows_cap, N]
    int E;
    int M;  // WGRAD: output rows per expert
    int N;  // output columns (per expert)
    int K;  // ROWS: reduction length
};

constexpr int kBM = 128;     // accumulator rows per CTA
constexpr int kPairM = 256;  // rows per pair tile (UMMA M with cta_group::2)
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemLimit = 232448;    // 227 KB
#ifndef MOE_PREFETCH_DIST
#define MOE_PREFETCH_DIST 8
#endif
constexpr int kPrefetchDist = MOE_PREFETCH_DIST;  // k-blocks the L2 prefetch cursor runs ahead of the smem ring

template <int BN, int EPI>
struct GemmCfg {
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = (BN / 2) * kBK * 2;  // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NOUT = (EPI == EPI_BIAS_GELU_DUAL) ? 2 : 1;
    // The epilogue works in chunks of 32 accumulator columns.  Every epilogue warp owns one 32-row slab per output
    // (32 x 64 B of bf16, 64-byte swizzle; 32 x 128 B of fp32 for WGRAD, 128-byte swizzle) that it fills and TMA-stores
    // on its own, and for DGELU two more 2 KB slabs into which it TMA-loads its rows of the pre-activation U one
    // chunk ahead.  Small slabs leave the shared memory to the operand ring.
    static constexpr int NCHUNK = BN / 32;
    static constexpr int SLAB_BYTES = (EPI == EPI_F32) ? 4096 : 2048;
    static constexpr int OUT_BYTES = kEpiWarps * NOUT * SLAB_BYTES;
    static constexpr int AUX_BYTES = (EPI == EPI_DGELU) ? kEpiWarps * 2 * 2048 : 0;
    static constexpr int STAGING_BYTES = OUT_BYTES + AUX_BYTES;
    static constexpr int BAR_BYTES = 512 + kEpiWarps * 128 * 4;  // mbarriers + TMEM slot, then the per-warp bias copies
    static constexpr int STAGES_RAW = (kSmemLimit - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
    static_assert(STAGES >= 3, "not enough shared memory for a pipelined tile");
    static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64, at most 256");
};

// ------------------------------------------------------------------------------------------------
// exact-erf GELU and its derivative for the epilogues (coefficients: tools/fit_gelu.py)
//   gelu(u)  = relu(u) - |u| Q(|u|),              Q(a) = Phi(-a) = exp2(P6(a))
//   gelu'(u) = u < 0 ? m(|u|) : 1 - m(|u|),       m(a) = Phi(-a) - a phi(a) = exp(-a^2/2) w8(a)
// P6 / w8: polynomials on [0, 6.5] (|u| is clamped; beyond it both corrections are < 1e-9).
// Max abs error vs float64 erfc: 1.0e-7 (gelu — the level of CUDA's erff), 1.2e-6 (gelu', a factor that is
// rounded to bf16 right after).  One MUFU (ex2) per element, Horner chains on packed fp32 FMAs (FFMA2).
// Why not erff/expf: the fp32 pipe retires 128 lane-FMAs per clock per SM, a 128 x 256 tile at K = 384
// leaves ~20 of them per element; erff + expf + the rest needs about twice that (measured round 1b).
// ------------------------------------------------------------------------------------------------
constexpr float kGeluAMax = 6.5f;
constexpr float kNegHalfLog2e = -0.72134752044448170f;
__device__ constexpr float kGeluLogQ[7] = {  // P6(a) = log2 Phi(-a)
    -9.999921094e-01f, -1.151212416e+00f, -4.587352391e-01f, -5.346286518e-02f, 8.115032863e-03f, -7.800564408e-04f,
    3.437071280e-05f};
__device__ constexpr float kGeluGradW[9] = {  // w8(a)
    4.999988248e-01f, -7.978099009e-01f, 2.492143745e-01f, -1.297814929e-01f, 5.588059320e-02f, -1.862516973e-02f,
    4.319766638e-03f, -5.991640005e-04f, 3.660521692e-05f};

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ uint32_t pack_bf16x2(float2 v) {
    __nv_bfloat162 b = __float22bfloat162_rn(v);
    return *reinterpret_cast<uint32_t*>(&b);
}

// Q(a) = Phi(-a) for 8 element pairs.  The eight Horner chains are written step-major so that
// consecutive FFMA2s are independent.
__device__ __forceinline__ void gelu_q8(const float2 (&a)[8], float2 (&q)[8]) {
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(splat2(kGeluLogQ[6]), a[i], splat2(kGeluLogQ[5]));
#pragma unroll
    for (int k = 4; k >= 0; --k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], a[i], splat2(kGeluLogQ[k]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = make_float2(ex2_approx(p[i].x), ex2_approx(p[i].y));
}
// m(a) = Phi(-a) - a phi(a) for 8 element pairs
__device__ __forceinline__ void gelu_m8(const float2 (&a)[8], float2 (&m)[8]) {
    float2 g[8], p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = __fmul2_rn(__fmul2_rn(a[i], splat2(kNegHalfLog2e)), a[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = make_float2(ex2_approx(g[i].x), ex2_approx(g[i].y));
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(splat2(kGeluGradW[8]), a[i], splat2(kGeluGradW[7]));
#pragma unroll
    for (int k = 6; k >= 0; --k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], a[i], splat2(kGeluGradW[k]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = __fmul2_rn(g[i], p[i]);
}

// One block of 16 accumulator columns (8 pairs) of one row through the epilogue.
//   acc: 16 fp32 accumulators; bias_saddr: shared-memory address of 16 fp32 (BIAS epilogues); aux: 8 packed bf16x2 of U (DGELU)
//   o0 / o1: 8 packed bf16x2 outputs each (o1 = gelu, fc1 only)
template <int EPI>
__device__ __forceinline__ void epilogue_block16(const uint32_t* acc, uint32_t bias_saddr, const uint32_t* aux,
                                                 uint32_t* o0, uint32_t* o1) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    if constexpr (EPI == EPI_BIAS_GELU_DUAL || EPI == EPI_BIAS) {
        float2 b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 b4;   // shared memory, warp-uniform address (broadcast)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(bias_saddr + i * 16));
            b[2 * i] = make_float2(b4.x, b4.y);
            b[2 * i + 1] = make_float2(b4.z, b4.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __fadd2_rn(v[i], b[i]);
    }
    if constexpr (EPI == EPI_DGELU) {
        float2 u[8], a[8], m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u[i] = make_float2(__uint_as_float(aux[i] << 16), __uint_as_float(aux[i] & 0xffff0000u));
            a[i] = make_float2(fminf(fabsf(u[i].x), kGeluAMax), fminf(fabsf(u[i].y), kGeluAMax));
        }
        gelu_m8(a, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 om = __ffma2_rn(m[i], splat2(-1.0f), splat2(1.0f));  // 1 - m
            const float2 gp = make_float2(u[i].x < 0.0f ? m[i].x : om.x, u[i].y < 0.0f ? m[i].y : om.y);
            o0[i] = pack_bf16x2(__fmul2_rn(v[i], gp));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o0[i] = pack_bf16x2(v[i]);
    }
#ifdef MOE_DBG_NO_GELU
    if constexpr (EPI == EPI_BIAS_GELU_DUAL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o1[i] = pack_bf16x2(v[i]);
    }
    return;
#endif
    if constexpr (EPI == EPI_BIAS_GELU_DUAL) {
        float2 a[8], qq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = make_float2(fminf(fabsf(v[i].x), kGeluAMax), fminf(fabsf(v[i].y), kGeluAMax));
        gelu_q8(a, qq);
#pragma unroll
        for (int i = 0; i < 8; ++i)  // relu(v) - a Q(a)
            o1[i] = pack_bf16x2(__ffma2_rn(make_float2(-a[i].x, -a[i].y), qq[i], make_float2(fmaxf(v[i].x, 0.0f), fmaxf(v[i].y, 0.0f))));
    }
}

// -DMOE_DBG_TIMELINE (kernel experiments only): per-tile clock64 stamps of the three roles of CTA 0 / CTA 1
//   g_tl[cta][role][tile][event]; role 0 = TMA producer, 1 = MMA issuer, 2 = epilogue warp 2, 3 = epilogue warp 6
#ifdef MOE_DBG_TIMELINE
__device__ long long g_tl[2][4][64][4];
#define MOE_TL(role, ti, ev)                                                                        \
    do {                                                                                            \
        if (blockIdx.x < 2 && (ti) < 64) g_tl[blockIdx.x][role][ti][ev] = clock64();                \
    } while (0)
#else
#define MOE_TL(role, ti, ev) do { } while (0)
#endif

struct TileCoord {
    int e;      // expert (weight index)
    int m0;     // ROWS: first packed row of THIS CTA's half.  WGRAD: first output row of this CTA's half
    int n0;     // first output column of the pair tile
    int row0;   // WGRAD: first packed row of the expert segment
    int kb;     // number of 64-deep k-blocks
};

template <int BN, bool WGRAD>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int n_ntiles, int rank) {
    TileCoord c;
    if constexpr (!WGRAD) {
        int m = tile / n_ntiles;
        c.e = __ldg(p.tile_expert + m);
        c.m0 = m * kPairM + rank * kBM;
        c.n0 = (tile - m * n_ntiles) * BN;
        c.row0 = 0;
        c.kb = p.K / kBK;
    } else {
        int m_tiles = (p.M + kPairM - 1) / kPairM;
        int per_e = m_tiles * n_ntiles;
        c.e = tile / per_e;
        int rem = tile - c.e * per_e;
        int mt = rem / n_ntiles;
        c.m0 = mt * kPairM + rank * kBM;
        c.n0 = (rem - mt * n_ntiles) * BN;
        c.row0 = __ldg(p.seg_start + c.e);
        c.kb = (__ldg(p.seg_start + c.e + 1) - c.row0) / kBK;
    }
    return c;
}

template <int BN, int EPI, bool WGRAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                    const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
    using Cfg = GemmCfg<BN, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(WGRAD == (EPI == EPI_F32), "WGRAD <=> fp32 output");
    static_assert(!WGRAD || BN % 128 == 0, "MN-major B: each CTA's half must be whole 64-column swizzle atoms");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    [[maybe_unused]] uint64_t* aux_bar = tempty_bar + 2;                                  // [8 warps][2]  (DGELU)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2 * kEpiWarps);
    [[maybe_unused]] float* bias_s = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES + 512);  // [8 warps][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());  // 0 = leader (issues the MMAs)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO0);
        if constexpr (Cfg::NOUT == 2) tma_prefetch_desc(&tmO1);
        if constexpr (EPI == EPI_DGELU) {
            tma_prefetch_desc(&tmAux);
            for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(aux_bar + i, 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);   // leader's producer arrive.expect_tx; bytes from both CTAs
            mbar_init(empty_bar + s, 1);  // one multicast tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + s, 1);                // one multicast tcgen05.commit
            mbar_init(tempty_bar + s, 2 * kEpiWarps);   // every epilogue warp of both CTAs (leader's copy is used)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits + TMEM allocation visible in both CTAs before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_ntiles = (p.N + BN - 1) / BN;
    int total_tiles;
    if constexpr (WGRAD) total_tiles = p.E * ((p.M + kPairM - 1) / kPairM) * n_ntiles;
    else total_tiles = __ldg(p.num_mtiles) * n_ntiles;
    const int first_tile = blockIdx.x >> 1;
    const int tile_stride = gridDim.x >> 1;

    if (warp == 0) {
        // ================================ TMA producer (one thread per CTA) =========================
        // -DMOE_L2_PREFETCH (experiment, off): a second cursor runs kPrefetchDist k-blocks ahead of the loads and pulls
        // those boxes into L2 (cp.async.bulk.prefetch.tensor).  Measured round 1e: 20-35 % SLOWER on every op (fc2
        // 68 -> 84 us) — the mainloop is bound by L2 -> SM request throughput, not by HBM latency, and the
        // prefetches double the requests.
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            [[maybe_unused]] int ti = 0;
            auto issue = [&](const TileCoord& c, int kb, uint8_t* sa, uint64_t* bar) {   // sa == nullptr: L2 prefetch only
                uint8_t* const sb = sa != nullptr ? sa + Cfg::A_BYTES : nullptr;
                if constexpr (!WGRAD) {
                    const int ca = kb * kBK, ra = c.m0, rb = c.e * p.N + c.n0 + rank * (BN / 2);
                    if (sa == nullptr) { tma_prefetch_2d(&tmA, ca, ra); tma_prefetch_2d(&tmB, ca, rb); return; }
                    tma_load_2d_pair(sa, &tmA, bar, ca, ra);
                    tma_load_2d_pair(sb, &tmB, bar, ca, rb);
                } else {
                    const int krow = c.row0 + kb * kBK;
                    if (sa == nullptr) {
                        tma_prefetch_2d(&tmA, c.m0, krow);
                        tma_prefetch_2d(&tmA, c.m0 + 64, krow);
#pragma unroll
                        for (int i = 0; i < BN / 128; ++i) tma_prefetch_2d(&tmB, c.n0 + rank * (BN / 2) + i * 64, krow);
                        return;
                    }
                    tma_load_2d_pair(sa, &tmA, bar, c.m0, krow);
                    tma_load_2d_pair(sa + 8192, &tmA, bar, c.m0 + 64, krow);
#pragma unroll
                    for (int i = 0; i < BN / 128; ++i)
                        tma_load_2d_pair(sb + i * 8192, &tmB, bar, c.n0 + rank * (BN / 2) + i * 64, krow);
                }
            };
            // prefetch cursor
            int ptile = first_tile, pkb = 0;
            TileCoord pc{};
            if (ptile < total_tiles) pc = decode_tile<BN, WGRAD>(p, ptile, n_ntiles, rank);
            [[maybe_unused]] auto prefetch_next = [&]() {
                while (ptile < total_tiles && pkb >= pc.kb) {   // next non-empty tile
                    ptile += tile_stride;
                    pkb = 0;
                    if (ptile < total_tiles) pc = decode_tile<BN, WGRAD>(p, ptile, n_ntiles, rank);
                }
                if (ptile >= total_tiles) return;
                issue(pc, pkb, nullptr, nullptr);
                ++pkb;
            };
#ifdef MOE_L2_PREFETCH
            for (int i = 0; i < STAGES + kPrefetchDist; ++i) prefetch_next();
#endif
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++ti) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
                MOE_TL(0, ti, 0);
                for (int kb = 0; kb < c.kb; ++kb) {
                    mbar_wait(empty_bar + s, ph ^ 1);
                    if (kb == 0) MOE_TL(0, ti, 1);
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(full_bar + s, 2 * Cfg::STAGE_BYTES);
                    issue(c, kb, sa, full_bar + s);
#ifdef MOE_L2_PREFETCH
                    prefetch_next();
#endif
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                MOE_TL(0, ti, 2);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (one thread of the leader CTA) =================
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kPairM, BN, WGRAD, WGRAD);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            [[maybe_unused]] int ti = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++ti) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
                if (c.kb == 0) continue;
                MOE_TL(1, ti, 0);
                mbar_wait(tempty_bar + as, aph ^ 1);
                tc_fence_after();
                MOE_TL(1, ti, 1);
                const uint32_t tmem_d = tmem_base + as * 256;
                for (int kb = 0; kb < c.kb; ++kb) {
                    mbar_wait(full_bar + s, ph);
                    tc_fence_after();
                    if (kb == 0) MOE_TL(1, ti, 2);
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < kBK / 16; ++k4) {
                        const uint64_t ad = WGRAD ? umma_smem_desc(a_addr + k4 * 2048, 8192, 1024)
                                                  : umma_smem_desc(a_addr + k4 * 32, 16, 1024);
                        const uint64_t bd = WGRAD ? umma_smem_desc(b_addr + k4 * 2048, 8192, 1024)
                                                  : umma_smem_desc(b_addr + k4 * 32, 16, 1024);
                        umma_bf16(tmem_d, ad, bd, idesc, (kb | k4) != 0);
                    }
                    umma_commit_pair(empty_bar + s);  // frees the smem slot in both CTAs once these MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_pair(tfull_bar + as);  // accumulator complete -> epilogues of both CTAs
                MOE_TL(1, ti, 3);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue (2 groups x 4 warps) ============================
        // Every warp is independent: it owns 32 accumulator rows (its TMEM lane quarter; thread = row), the two groups
        // take alternate 32-column chunks, and each warp stages and TMA-stores its own slabs, so there is no CTA-level
        // barrier anywhere in the epilogue.
        const int q = warp & 3;                  // TMEM lane quarter this warp may touch
        const int grp = (warp - 2) >> 2;         // column-chunk parity this group owns
        const int ew = warp - 2;                 // 0..7
        float* const wbias = bias_s + ew * 128;  // this warp's copy of the bias values of its chunks
        int as = 0;
        uint32_t aph = 0;
        [[maybe_unused]] int ti = 0;
        [[maybe_unused]] const int tl_role = warp == 2 ? 2 : 3;
        [[maybe_unused]] const bool tl_on = (warp == 2 || warp == 6) && lane == 0;
        if constexpr (EPI == EPI_F32) {
            uint8_t* const slab = staging + ew * Cfg::SLAB_BYTES;     // 32 rows x 128 B, 128-byte swizzle
            uint8_t* const my_row = slab + lane * 128;
            const int sw = lane & 7;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++ti) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
                const bool live = c.kb != 0;
                if (tl_on) MOE_TL(tl_role, ti, 0);
                if (live) {
                    mbar_wait(tfull_bar + as, aph);
                    tc_fence_after();
                }
                if (tl_on) MOE_TL(tl_role, ti, 1);
                const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
#pragma unroll 1
                for (int ch = grp; ch < Cfg::NCHUNK; ch += 2) {
                    const bool last_chunk = (ch + 2 >= Cfg::NCHUNK);
                    uint32_t acc[32];
                    if (live) {
                        tmem_ld32(tmem_row + ch * 32, acc);
                        tmem_ld_wait();
                        if (last_chunk) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                            if (tl_on) MOE_TL(tl_role, ti, 2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0u;
                    }
                    if (lane == 0) tma_store_wait_read<0>();   // this warp's previous store has left its slab
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) =
                            make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && c.m0 + q * 32 < p.M && c.n0 + ch * 32 < p.N) {
                        tma_store_3d(&tmO0, slab, c.n0 + ch * 32, c.m0 + q * 32, c.e);
                        tma_store_commit();
                    }
                }
                if (tl_on) MOE_TL(tl_role, ti, 3);
                if (live && ++as == 2) { as = 0; aph ^= 1; }
            }
        } else {
            // bf16 outputs: slab rows are 64 B (four 16-byte units), unit u of row r lives at unit u ^ ((r >> 1) & 3)
            // (CU_TENSOR_MAP_SWIZZLE_64B); a quarter-warp then touches all 32 banks exactly once.
            const uint32_t out_s = smem_u32(staging) + ew * Cfg::NOUT * 2048;                 // this warp's output slab(s)
            const uint32_t my_out = out_s + lane * 64;
            const int sw = (lane >> 1) & 3;
            [[maybe_unused]] const uint32_t aux_s = smem_u32(staging) + Cfg::OUT_BYTES + ew * 4096;   // two 2 KB slabs
            [[maybe_unused]] uint64_t* const my_aux_bar = aux_bar + ew * 2;
            [[maybe_unused]] uint32_t aux_issued = 0, aux_used = 0;   // running chunk counters of this warp (slab = n & 1)
            constexpr int MYCH = Cfg::NCHUNK / 2;                     // chunks of one group per tile
            // DGELU: TMA-load this warp's 32 rows x 32 columns of U for chunk `ch` of the tile at `c`
            [[maybe_unused]] auto issue_aux = [&](const TileCoord& c, int ch) {
                if (lane == 0) {
                    uint64_t* bar = my_aux_bar + (aux_issued & 1);
                    mbar_arrive_expect_tx(bar, 2048);
                    tma_load_2d_s(aux_s + (aux_issued & 1) * 2048, &tmAux, bar, c.n0 + ch * 32, c.m0 + q * 32);
                }
                ++aux_issued;
            };
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++ti) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
                const bool live = c.kb != 0;
                if (tl_on) MOE_TL(tl_role, ti, 0);
                if constexpr (EPI == EPI_BIAS_GELU_DUAL || EPI == EPI_BIAS) {
                    // lane l fetches 4 consecutive bias values of the group's (at most four) chunks
                    const int lc = grp + 2 * (lane >> 3);       // chunk the lane's values belong to
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lc < Cfg::NCHUNK)
                        bv = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(c.e) * p.N + c.n0 + lc * 32) + (lane & 7));
                    __syncwarp();                                // previous tile's reads of wbias are done
                    *reinterpret_cast<float4*>(wbias + lane * 4) = bv;
                    __syncwarp();
                }
                if constexpr (EPI == EPI_DGELU) issue_aux(c, grp);   // lands while this warp waits for the accumulator
                if (live) {
                    mbar_wait(tfull_bar + as, aph);
                    tc_fence_after();
                }
                if (tl_on) MOE_TL(tl_role, ti, 1);
                const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
#pragma unroll 1
                for (int i = 0; i < MYCH; ++i) {
                    const int ch = grp + 2 * i;
                    const bool last_chunk = (i + 1 == MYCH);
#ifdef MOE_DBG_NO_EPI
                    if (last_chunk) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                    }
                    continue;
#endif
                    [[maybe_unused]] uint32_t aux_rd = 0;
                    if constexpr (EPI == EPI_DGELU) {
                        // the other aux slab is free (its chunk was consumed in the previous iteration): prefetch the next chunk
                        if (!last_chunk) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            issue_aux(c, ch + 2);
                        }
                        mbar_wait(my_aux_bar + (aux_used & 1), (aux_used >> 1) & 1);
                        aux_rd = aux_s + (aux_used & 1) * 2048 + lane * 64;
                        ++aux_used;
                    }
                    const uint32_t cbias = smem_u32(wbias) + i * 128;
                    uint32_t acc[2][16];
                    tmem_ld16(tmem_row + ch * 32, acc[0]);
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {   // 16 accumulator columns at a time
                        tmem_ld_wait();
                        if (blk == 0) {
                            tmem_ld16(tmem_row + ch * 32 + 16, acc[1]);
                        } else if (last_chunk) {
                            // every TMEM read of this accumulator stage by this warp is done -> hand it back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                            if (tl_on) MOE_TL(tl_role, ti, 2);
                        }
#pragma unroll
                        for (int r = 0; r < 16; ++r) asm volatile("" : "+r"(acc[blk][r]));
                        [[maybe_unused]] uint32_t aux[8];
                        if constexpr (EPI == EPI_DGELU) {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(aux[4 * j]), "=r"(aux[4 * j + 1]), "=r"(aux[4 * j + 2]), "=r"(aux[4 * j + 3])
                                             : "r"(aux_rd + (((blk * 2 + j) ^ sw) << 4)));
                        }
                        uint32_t o0[8];                       // 16 columns of output 0, packed bf16x2
                        [[maybe_unused]] uint32_t o1[8];      // 16 columns of output 1 (fc1: gelu)
                        epilogue_block16<EPI>(acc[blk], cbias + blk * 64, aux, o0, o1);
                        if (blk == 0) {
                            if (lane == 0) tma_store_wait_read<0>();   // this warp's previous store has left its slab(s)
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t slot = my_out + (((blk * 2 + j) ^ sw) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(o0[4 * j]), "r"(o0[4 * j + 1]),
                                         "r"(o0[4 * j + 2]), "r"(o0[4 * j + 3]) : "memory");
                            if constexpr (EPI == EPI_BIAS_GELU_DUAL)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 2048), "r"(o1[4 * j]),
                                             "r"(o1[4 * j + 1]), "r"(o1[4 * j + 2]), "r"(o1[4 * j + 3]) : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
#ifndef MOE_DBG_NO_STORE
                    if (lane == 0) {
                        tma_store_2d_s(&tmO0, out_s, c.n0 + ch * 32, c.m0 + q * 32);
                        if constexpr (Cfg::NOUT == 2) tma_store_2d_s(&tmO1, out_s + 2048, c.n0 + ch * 32, c.m0 + q * 32);
                        tma_store_commit();
                    }
#endif
                }
                if (tl_on) MOE_TL(tl_role, ti, 3);
                if (live && ++as == 2) { as = 0; aph ^= 1; }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    // teardown: neither CTA may exit (or free TMEM) while its peer can still signal or read it
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace moe
