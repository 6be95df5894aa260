// gemm.cuh — grouped expert GEMM on tcgen05 / TMEM, fed by TMA, on CTA PAIRS (sm_100a only).
//
// One persistent, warp-specialised kernel template covers every dense contraction of the
// expert FFN (replaces FastMoE's per-expert cuBLAS loop
// `fmoe_cuda.linear_forward/backward`, reached from /root/reference/models/resMoE.py:27-29):
//
//   ROWS mode  (M = packed token rows, one weight matrix per 256-row pair tile; A K-major, B K-major (forward) or MN-major (backward)):
//     fc1   : U = X  W1^T + b1 -> G = gelu_erf'(U), H = gelu_erf(U)   B = W1   [E, h, d]   EPI_BIAS_GELU_DUAL
//     fc2   : Y = H  W2^T + b2                      B = W2   [E, d, h]      EPI_BIAS
//     dgelu : dU = (dY W2) * G                      B = W2   [E, d, h] MN-major (N = h contiguous)   EPI_DGELU
//     dgrad : dX = dU W1                            B = W1   [E, h, d] MN-major (N = d contiguous)   EPI_PLAIN
//   (the two backward contractions read the SAME bf16 weight copies as the forward ones, through MN-major UMMA
//    descriptors — no transposed copies: the per-step weight cast writes half as much)
//   WGRAD mode (K = the rows of one expert's segment, fp32 output per expert; A and B MN-major):
//     dW1[e] = dU_e^T X_e   [E, M = h, N = d]                               EPI_F32
//     dW2[e] = (H_e^T dY_e)^T: computed as [M = h, N = d], stored transposed EPI_F32_T
//   (both weight gradients then have M = h, N = d: at d = 384 that is six 256-row M tiles and two 192-column N
//    tiles per expert with no padding, where dW2 as [M = d, N = h] left half of every second M tile empty)
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
// CTA = 64 + 32 W threads: warp 0 = TMA producer, warp 1 = TMEM owner (+ single-thread MMA issuer in the leader
// CTA), W epilogue warps (8; 12 for fc1) in groups of four (one TMEM lane quarter each); the groups take alternate
// 32-column chunks of the accumulator, and every warp stages and TMA-stores its own 32-row slab, so the epilogue
// has no CTA-level barrier.  Pipelines: smem ring full/empty (TMA <-> MMA; `full` lives in the leader and is
// credited by both CTAs' TMA loads, `empty` is multicast to both CTAs by tcgen05.commit), two TMEM accumulator
// stages full/empty (MMA <-> epilogue of both CTAs), per-warp store slabs drained by TMA.
//
// What bounds it (round 1e, -DMOE_DBG_TIMELINE stamps + tools/ubench/tmem_mma_ld.cu):
//   * tcgen05.ld does NOT contend with tcgen05.mma: 128 KB of accumulator read in ~410 clocks (320 B/clk/SM) with
//     or without MMAs in flight.
//   * the mainloop runs at ~1 000 clocks per 64-deep k-block against 512 clocks of MMA issue: operand feed.  An L2
//     prefetch cursor ahead of the ring made every op 20-35 % slower, and keeping the weight tile resident in shared
//     memory for the K = 384 ops (half the L2 -> SM traffic) made them 7-10 % slower: it is neither HBM latency nor
//     L2 bandwidth.  Per tile the shared memory moves the TMA fill, the UMMA operand reads (A 4 KB + B 8 KB per
//     MMA and SM, the B half served twice) and the epilogue's staging + TMA-store reads — ~740 KB for an fc1 tile,
//     5 800 clocks at 128 B/clk, which is what a tile takes.
//   * epilogues: gelu' moved from dgelu into fc1 (one shared exp + polynomial), dgelu reads G through TMA
//     most of a tile ahead (per-lane row loads cost 4 000 LSU cycles per tile, a one-chunk-ahead TMA exposed its
//     ~1 500-clock latency four times per tile), fc1 runs 12 epilogue warps.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace moe {

enum : int { EPI_BIAS_GELU_DUAL = 0, EPI_BIAS = 1, EPI_DGELU = 2, EPI_PLAIN = 3, EPI_F32 = 4, EPI_F32_T = 5,
              EPI_BIAS_GELU = 6 /* fc1 without gelu': forward-only (evaluation) passes */ };

struct GemmParams {
    const int* tile_expert;  // ROWS: expert of each 256-row pair tile       [max_mtiles]
    const int* num_mtiles;   // ROWS: number of live 256-row tiles (device scalar)
    const int* seg_start;    // WGRAD: first row of each expert segment      [E+1]
    const float* bias;       // [E, N] fp32 or nullptr
    const __nv_bfloat16* aux;  // EPI_DGELU: G = gelu'(U) [rows_cap, N] (read through its tensor map)
    float* colsum;           // EPI_DGELU (optional): column sums of every 32-row output slab, [rows_cap / 32, N] fp32
    int* flags;              // WGRAD stream-K: one int per (tile, CTA rank, epilogue warp), zero between launches
    int streamk;             // WGRAD: 1 = every CTA pair takes an equal share of the linearised (tile, k-block) space (StreamK below)
    int ksplit;              // WGRAD, !streamk: S >= 2 = every tile's K range is cut in S equal parts = S work units (round-robin)
    int E;
    int M;  // WGRAD: output rows per expert
    int N;  // output columns (per expert)
    int K;  // ROWS: reduction length
};

constexpr int kBM = 128;     // accumulator rows per CTA
constexpr int kPairM = 256;  // rows per pair tile (UMMA M with cta_group::2)
constexpr int kBK = 64;
// epilogue warps per CTA (a multiple of 4: one warp per TMEM lane quarter and column group).  The fc1 epilogue
// evaluates gelu and gelu' (~13 instructions per element) and is latency-bound at two warps per scheduler
// (measured: 6 300 clocks per tile against a 6 600-clock mainloop); a third group of four warps hides that.
#ifndef MOE_FC1_EPI_WARPS
#define MOE_FC1_EPI_WARPS 12
#endif
// The fp32 (weight-gradient) epilogues also run 12 warps: their tile is up to 384 columns wide and not overlapped.
#ifndef MOE_WIDE_EPI_WARPS
#define MOE_WIDE_EPI_WARPS 8
#endif
#ifndef MOE_DGELU_EPI_WARPS
#define MOE_DGELU_EPI_WARPS 8
#endif
__host__ __device__ constexpr int epi_warps(int epi, int bn) {
    return epi == 0 /* EPI_BIAS_GELU_DUAL */ ? MOE_FC1_EPI_WARPS : (epi == 4 || epi == 5) /* EPI_F32, EPI_F32_T */ ? 12
           : epi == 2 /* EPI_DGELU */ ? MOE_DGELU_EPI_WARPS : bn > 256 ? MOE_WIDE_EPI_WARPS : 8;
}
constexpr int kSmemLimit = 232448;    // 227 KB

template <int BN, int EPI>
struct GemmCfg {
    // BN <= 256: one UMMA per k-step and two TMEM accumulator stages (the epilogue of a tile overlaps the next
    // tile's mainloop).  BN = 384: ONE 384-column accumulator fed by two UMMAs per k-step (N = 256 + N = 128) that
    // share the A tile: 40 KB of operands per 256 x 384 x 64 block instead of 2 x 28 KB (BN = 192) — the mainloop
    // is bound by the L2 -> SM operand feed (~35 B/clk/SM measured on every variant), so bytes per flop is what
    // counts; the epilogue no longer overlaps, which long-K launches (K >= 768, WGRAD) amortise.
    static constexpr int NSUB = (BN + 255) / 256;
    static constexpr int ACC_STAGES = BN <= 256 ? 2 : 1;
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = (BN / 2) * kBK * 2;  // this CTA's half of the B tile (one k-block)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NOUT = (EPI == EPI_BIAS_GELU_DUAL) ? 2 : 1;
    static constexpr int EPI_WARPS = epi_warps(EPI, BN);
    static constexpr int NGRP = EPI_WARPS / 4;          // column groups: group g takes chunks g, g + NGRP, ...
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;  // warp 0 = TMA producer, warp 1 = TMEM owner / MMA issuer
    // The epilogue works in chunks of 32 accumulator columns.  Every epilogue warp owns one 32-row slab per output
    // (32 x 64 B of bf16, 64-byte swizzle; 32 x 128 B of fp32 for WGRAD, 128-byte swizzle) that it fills and TMA-stores
    // on its own.  DGELU works in place: a warp has one slab per chunk of its tile share, TMA-loads its rows of
    // G = gelu'(U) into it most of a tile ahead (a TMA round trip is ~1 500 clocks, four times a chunk's arithmetic),
    // multiplies in place and stores the same slab.  Small slabs leave the shared memory to the operand ring.
    static constexpr int NCHUNK = BN / 32;
    static constexpr int MAXCH = (NCHUNK + NGRP - 1) / NGRP;   // chunks of one warp per tile
    static constexpr bool F32 = (EPI == EPI_F32 || EPI == EPI_F32_T);
    // B is MN-major (N contiguous in global memory, read in [64 k] x [B_ATOM n] boxes) for the weight gradients and for
    // the two backward row-mode contractions, which read the forward weights W2 [E, d, h] / W1 [E, h, d] as [K, N]
    static constexpr bool B_MN = F32 || EPI == EPI_DGELU || EPI == EPI_PLAIN;
    // WGRAD reads B MN-major: whole 128-byte swizzle atoms (64 columns) when this CTA's half allows it, else
    // 64-byte swizzle atoms (32 columns): BN = 192 -> 96 columns per CTA = three 32-column atoms
    static constexpr int B_ATOM = ((BN / 2) % 64 == 0) ? 64 : 32;
    static constexpr int SLAB_BYTES = F32 ? 4096 : 2048;
    static constexpr int OUT_BYTES = EPI_WARPS * SLAB_BYTES * (EPI == EPI_DGELU ? MAXCH : NOUT);
    static constexpr int STAGING_BYTES = OUT_BYTES;
    static constexpr int BIAS_FLOATS = (MAXCH > 4 ? MAXCH : 4) * 32;   // per-warp copy of the bias of its chunks
    static constexpr int BAR_BYTES = 512 + EPI_WARPS * BIAS_FLOATS * 4;  // mbarriers + TMEM slot, then the bias copies
    static constexpr int STAGES_RAW = (kSmemLimit - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
    static_assert(STAGES >= 3, "not enough shared memory for a pipelined tile");
    static_assert(BN % 64 == 0 && (BN <= 256 || BN == 384), "BN must be a multiple of 64, at most 256, or 384");
    static_assert(BN <= 256 || (EPI != EPI_DGELU && EPI != EPI_BIAS_GELU_DUAL), "wide tiles: fc2 / dgrad / wgrad only");
};

// ------------------------------------------------------------------------------------------------
// exact-erf GELU and its derivative, both in the fc1 epilogue (coefficients: tools/fit_gelu.py)
//   a = min(|u|, 6.5),  g = exp(-a^2 / 2) (one ex2.approx),  W(a) = Phi(-a) / g = 0.5 erfcx(a / sqrt 2): degree-7 polynomial
//   gelu(u)  = relu(u) - a g W(a)
//   gelu'(u) = u < 0 ? m : 1 - m,   m = Phi(-a) - a phi(a) = g (W(a) - a / sqrt(2 pi))
// fc1 writes H = gelu(U) and G = gelu'(U) (bf16) instead of the pre-activation itself: backward needs U only
// through gelu'(U), so the dgelu epilogue shrinks to one multiply per element, and gelu' is evaluated from the fp32
// pre-activation, not from its bf16 rounding.  Max abs error vs float64 erfc: 5.3e-6 (gelu), 1.2e-5 (gelu') — both
// far below the bf16 rounding of the stored values (2^-9 relative).  One MUFU per element, the Horner chain on packed
// fp32 FMAs (FFMA2).
// ------------------------------------------------------------------------------------------------
constexpr float kGeluAMax = 6.5f;
constexpr float kNegHalfLog2e = -0.72134752044448170f;
constexpr float kNegInvSqrt2Pi = -0.39894228040143268f;
#ifndef MOE_GELU_DEG
#define MOE_GELU_DEG 7
#endif
constexpr int kGeluDeg = MOE_GELU_DEG;
#if MOE_GELU_DEG == 7
__device__ constexpr float kGeluW[8] = {  // W(a) = 0.5 erfcx(a / sqrt 2) on [0, 6.5]
    4.999879883e-01f, -3.984107670e-01f, 2.460856669e-01f, -1.217100563e-01f, 4.573317066e-02f, -1.176512435e-02f,
    1.782060358e-03f, -1.172508184e-04f};
#elif MOE_GELU_DEG == 5   // kernel experiments: 7.9e-5 (gelu) / 1.7e-4 (gelu') max abs error
__device__ constexpr float kGeluW[6] = {4.998291619e-01f, -3.940517119e-01f, 2.266079638e-01f, -8.922066891e-02f, 2.040228419e-02f,
                                        -1.967727139e-03f};
#endif

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ uint32_t pack_bf16x2(float2 v) {
    __nv_bfloat162 b = __float22bfloat162_rn(v);
    return *reinterpret_cast<uint32_t*>(&b);
}

// One block of 16 accumulator columns (8 pairs) of one row through the epilogue.
//   acc: 16 fp32 accumulators; bias_saddr: shared-memory address of 16 fp32 (BIAS epilogues); aux: 8 packed bf16x2 of G (DGELU)
//   o0 / o1: 8 packed bf16x2 outputs each (fc1: o0 = gelu'(u), o1 = gelu(u))
template <int EPI>
__device__ __forceinline__ void epilogue_block16(const uint32_t* acc, uint32_t bias_saddr, const uint32_t* aux,
                                                 uint32_t* o0, uint32_t* o1) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    if constexpr (EPI == EPI_BIAS_GELU_DUAL || EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
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
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // dU = (dY W2) * gelu'(U)
            const float2 gp = make_float2(__uint_as_float(aux[i] << 16), __uint_as_float(aux[i] & 0xffff0000u));
            o0[i] = pack_bf16x2(__fmul2_rn(v[i], gp));
        }
    } else if constexpr (EPI == EPI_BIAS_GELU_DUAL) {
#ifdef MOE_DBG_NO_GELU
#pragma unroll
        for (int i = 0; i < 8; ++i) o0[i] = o1[i] = pack_bf16x2(v[i]);
        return;
#endif
        float2 a[8], g[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = make_float2(fminf(fabsf(v[i].x), kGeluAMax), fminf(fabsf(v[i].y), kGeluAMax));
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = __fmul2_rn(__fmul2_rn(a[i], splat2(kNegHalfLog2e)), a[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = make_float2(ex2_approx(g[i].x), ex2_approx(g[i].y));
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ffma2_rn(splat2(kGeluW[kGeluDeg]), a[i], splat2(kGeluW[kGeluDeg - 1]));
#pragma unroll
        for (int k = kGeluDeg - 2; k >= 0; --k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ffma2_rn(w[i], a[i], splat2(kGeluW[k]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 q = __fmul2_rn(g[i], w[i]);                                                   // Phi(-a)
            o1[i] = pack_bf16x2(__ffma2_rn(make_float2(-a[i].x, -a[i].y), q, make_float2(fmaxf(v[i].x, 0.0f), fmaxf(v[i].y, 0.0f))));
            const float2 m = __fmul2_rn(g[i], __ffma2_rn(a[i], splat2(kNegInvSqrt2Pi), w[i]));          // Phi(-a) - a phi(a)
            const float2 om = __ffma2_rn(m, splat2(-1.0f), splat2(1.0f));
            o0[i] = pack_bf16x2(make_float2(v[i].x < 0.0f ? m.x : om.x, v[i].y < 0.0f ? m.y : om.y));
        }
    } else if constexpr (EPI == EPI_BIAS_GELU) {   // same arithmetic as the gelu half of the dual epilogue (bit-identical H)
        float2 a[8], g[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = make_float2(fminf(fabsf(v[i].x), kGeluAMax), fminf(fabsf(v[i].y), kGeluAMax));
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = __fmul2_rn(__fmul2_rn(a[i], splat2(kNegHalfLog2e)), a[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = make_float2(ex2_approx(g[i].x), ex2_approx(g[i].y));
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ffma2_rn(splat2(kGeluW[kGeluDeg]), a[i], splat2(kGeluW[kGeluDeg - 1]));
#pragma unroll
        for (int k = kGeluDeg - 2; k >= 0; --k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ffma2_rn(w[i], a[i], splat2(kGeluW[k]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 q = __fmul2_rn(g[i], w[i]);                                                   // Phi(-a)
            o0[i] = pack_bf16x2(__ffma2_rn(make_float2(-a[i].x, -a[i].y), q, make_float2(fmaxf(v[i].x, 0.0f), fmaxf(v[i].y, 0.0f))));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o0[i] = pack_bf16x2(v[i]);
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
    int kb0;    // WGRAD: first k-block of this work unit inside the expert segment
    int part;   // WGRAD stream-K: this fragment is the part-th piece of its tile's K range: 0 = plain store, s > 0 waits for
                //   part s - 1 and reduce-adds (fixed order); -1 = the whole tile (no flags)
    bool last;  // WGRAD stream-K: the fragment that ends the tile's K range (leaves the flag at zero)
    int tile;   // WGRAD: output tile index (flags)
};
// first B column (inside the pair tile) of this CTA's i-th 64-column block: BN <= 256: this CTA's half;
// BN = 384: blocks 0, 1 belong to the N = 256 UMMA (columns 0..255), block 2 to the N = 128 UMMA (columns 256..383)
template <int BN>
__device__ __forceinline__ int b_block_col(int rank, int i) {
    if constexpr (BN <= 256) return rank * (BN / 2) + i * 64;
    else return i < 2 ? rank * 128 + i * 64 : 256 + rank * 64;
}

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
        c.kb0 = 0; c.part = -1; c.tile = tile;
    } else {
        int m_tiles = (p.M + kPairM - 1) / kPairM;
        int per_e = m_tiles * n_ntiles;
        c.part = -1;
        if (p.ksplit >= 2) {   // units [s tiles, (s + 1) tiles) are part s of every tile: a part only ever waits on a lower unit index
            const int ntile = p.E * per_e;
            c.part = tile / ntile;
            tile -= c.part * ntile;
        }
        c.tile = tile;
        c.e = tile / per_e;
        int rem = tile - c.e * per_e;
        int mt = rem / n_ntiles;
        c.m0 = mt * kPairM + rank * kBM;
        c.n0 = (rem - mt * n_ntiles) * BN;
        c.row0 = __ldg(p.seg_start + c.e);
        c.kb = (__ldg(p.seg_start + c.e + 1) - c.row0) / kBK;
        c.kb0 = 0;
        if (c.part >= 0) {     // part s covers k-blocks [kb s / S, kb (s + 1) / S)
            const int b0 = static_cast<int>(static_cast<long long>(c.kb) * c.part / p.ksplit);
            const int b1 = static_cast<int>(static_cast<long long>(c.kb) * (c.part + 1) / p.ksplit);
            c.kb0 = b0;
            c.kb = b1 - b0;
            c.last = c.part + 1 == p.ksplit;
            return c;
        }
    }
    c.last = true;
    return c;
}

// ------------------------------------------------------------------------------------------------
// Stream-K schedule of the weight gradients (WGRAD, p.streamk).
//
// 96 tiles of 256 x 384 (config 2) on 74 CTA pairs are 1.3 rounds; cutting every tile's K range in two equal work units
// (round 2a) made that 2.6 -> 3 rounds of half length.  Here the (tile, k-block) space is laid out on one line — the tiles
// in (expert, m, n) order, each as long as its expert's k-blocks plus kEpiKb units that stand for its epilogue — and pair p
// takes the p-th of `pairs` equal spans of that line.  A span boundary that would leave fewer than kSkMinKb k-blocks of
// a tile on one side (or that falls into the epilogue units) moves to the tile boundary.  A pair walks its span BACKWARDS:
// the piece of a tile that starts the tile's K range (part 0, a plain store) is then the first thing its pair runs, and the
// piece that continues a tile begun by the previous pair (part s > 0: waits for part s - 1 through the per-warp flag, then
// reduce-adds) is the last thing its pair runs — by then the store it waits for is long done, and every wait points at a
// fragment that is first on a lower-numbered pair: no circular wait as long as all pairs are resident (one CTA per SM).
// One store then the adds in one fixed order: bit-reproducible.  The tables are rebuilt by every CTA from seg_start
// (device data: ragged and empty experts need no host involvement).
// ------------------------------------------------------------------------------------------------
constexpr int kSkMaxE = 128;       // experts per launch the tables hold (more experts => more tiles than 3 rounds: no stream-K)
constexpr int kSkMaxPairs = 128;
constexpr int kSkMinKb = 4;        // no fragment shorter than this many k-blocks
constexpr int kSkMinSpan = 24;     // no span shorter than this many units of the line
struct SkTab {
    int off[kSkMaxE + 1];    // position of each expert's first tile on the line
    int kb[kSkMaxE];         // k-blocks of each expert segment
    int cut[kSkMaxPairs + 1];   // span boundaries: pair p owns [cut[p], cut[p + 1])
};

// one full warp, before the prologue's cluster barrier
template <int EPI_KB>
__device__ __forceinline__ void sk_build(SkTab* t, const GemmParams& p, int per_e, int pairs, int lane) {
    for (int e = lane; e < p.E; e += 32) t->kb[e] = (__ldg(p.seg_start + e + 1) - __ldg(p.seg_start + e)) / kBK;
    __syncwarp();
    const int chunk = (p.E + 31) / 32;   // experts per lane (contiguous)
    int s = 0;
    for (int i = 0; i < chunk; ++i) {
        const int e = lane * chunk + i;
        if (e < p.E) s += per_e * (t->kb[e] + EPI_KB);
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int run = incl - s;
    for (int i = 0; i < chunk; ++i) {
        const int e = lane * chunk + i;
        if (e < p.E) { t->off[e] = run; run += per_e * (t->kb[e] + EPI_KB); }
    }
    const int L = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) t->off[p.E] = L;
    __syncwarp();
    // small problems: a span shorter than kSkMinSpan units is mostly epilogue — use fewer pairs (the rest get empty spans)
    const int act = max(1, min(pairs, L / kSkMinSpan));
    for (int q = lane; q <= pairs; q += 32) {
        int c = q >= act ? L : static_cast<int>(static_cast<long long>(L) * q / act);
        if (q > 0 && q < act) {
            int lo = 0, hi = p.E - 1;   // expert holding c: the largest e with off[e] <= c
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (t->off[mid] <= c) lo = mid; else hi = mid - 1;
            }
            const int kb = t->kb[lo], len = kb + EPI_KB;
            const int within = c - t->off[lo];
            const int ti = within / len, o = within - ti * len;
            const int o_t = t->off[lo] + ti * len;
            if (o == 0) c = o_t;
            else if (o >= kb - kSkMinKb) c = o_t + len;   // too close to the end of the K range, or inside the epilogue units
            else if (o < kSkMinKb) c = o_t;
        }
        t->cut[q] = c;
    }
    __syncwarp();
}

// The work units of one CTA pair, in the order the pair runs them: whole tiles p, p + pairs, ... — or the stream-K walk.
template <int BN, bool WGRAD, int EPI_KB>
struct WorkIter {
    const GemmParams& p;
    const SkTab* sk;
    int n_ntiles, rank, tile, total, stride;
    int cur, c0, e, per_e;
    __device__ __forceinline__ WorkIter(const GemmParams& p_, const SkTab* sk_, int n_ntiles_, int rank_, int total_)
        : p(p_), sk(sk_), n_ntiles(n_ntiles_), rank(rank_), tile(blockIdx.x >> 1), total(total_), stride(gridDim.x >> 1) {
        if constexpr (WGRAD) {
            if (p.streamk) {
                c0 = sk->cut[tile];
                cur = sk->cut[tile + 1];
                e = p.E - 1;
                per_e = ((p.M + kPairM - 1) / kPairM) * n_ntiles;
            }
        }
    }
    __device__ __forceinline__ bool next(TileCoord& c) {
        if constexpr (WGRAD) {
            if (p.streamk) {
                if (cur <= c0) return false;
                const int pos = cur - 1;
                while (sk->off[e] > pos) --e;
                const int kbe = sk->kb[e], len = kbe + EPI_KB;
                const int ti = (pos - sk->off[e]) / len;
                const int o_t = sk->off[e] + ti * len;
                const int kend = min(cur - o_t, kbe);
                const int kstart = max(c0 - o_t, 0);
                const int mt = ti / n_ntiles;
                c.e = e;
                c.m0 = mt * kPairM + rank * kBM;
                c.n0 = (ti - mt * n_ntiles) * BN;
                c.row0 = __ldg(p.seg_start + e);
                c.kb0 = kstart;
                c.kb = kend - kstart;
                c.tile = e * per_e + ti;
                c.last = kend == kbe;
                if (kstart == 0) {
                    c.part = c.last ? -1 : 0;
                } else {   // parts before this one: the non-empty spans of lower pairs that reach into this tile
                    int part = 0;
                    for (int q = tile - 1; q >= 0 && sk->cut[q + 1] > o_t; --q) part += sk->cut[q] < sk->cut[q + 1] ? 1 : 0;
                    c.part = part;
                }
                cur = o_t;
                return true;
            }
        }
        if (tile >= total) return false;
        c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
        tile += stride;
        return true;
    }
};

template <int BN, int EPI, bool WGRAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((GemmCfg<BN, EPI>::THREADS), 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                    const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
    using Cfg = GemmCfg<BN, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(WGRAD == Cfg::F32, "WGRAD <=> fp32 output");
    static_assert(!WGRAD || BN % 64 == 0, "MN-major B: each CTA's half must be whole 32-column swizzle atoms");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    [[maybe_unused]] uint64_t* aux_bar = tempty_bar + 2;                                  // [warps][MAXCH]  (DGELU)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + (EPI == EPI_DGELU ? Cfg::EPI_WARPS * Cfg::MAXCH : 0));
    static_assert((2 * STAGES + 4 + (EPI == EPI_DGELU ? Cfg::EPI_WARPS * Cfg::MAXCH : 0)) * 8 + 4 <= 512, "barrier block overflows its 512 bytes");
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
            for (int i = 0; i < Cfg::EPI_WARPS * Cfg::MAXCH; ++i) mbar_init(aux_bar + i, 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);   // leader's producer arrive.expect_tx; bytes from both CTAs
            mbar_init(empty_bar + s, 1);  // one multicast tcgen05.commit
        }
        for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
            mbar_init(tfull_bar + s, 1);                // one multicast tcgen05.commit
            mbar_init(tempty_bar + s, 2 * Cfg::EPI_WARPS);   // every epilogue warp of both CTAs (leader's copy is used)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    const int n_ntiles = (p.N + BN - 1) / BN;
    // WGRAD has no bias: the stream-K tables live in the bias area
    constexpr int kEpiKb = Cfg::ACC_STAGES == 1 ? 6 : 2;   // epilogue of one work unit, in k-blocks of mainloop time
    static_assert(sizeof(SkTab) <= Cfg::EPI_WARPS * Cfg::BIAS_FLOATS * 4, "stream-K tables do not fit the bias area");
    const SkTab* const sk = reinterpret_cast<const SkTab*>(bias_s);
    if constexpr (WGRAD) {
        if (p.streamk && warp == 2)
            sk_build<kEpiKb>(reinterpret_cast<SkTab*>(bias_s), p, ((p.M + kPairM - 1) / kPairM) * n_ntiles, gridDim.x >> 1, lane);
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits, TMEM allocation (and the stream-K tables) visible before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int total_tiles;
    if constexpr (WGRAD) total_tiles = p.E * ((p.M + kPairM - 1) / kPairM) * n_ntiles * (p.ksplit >= 2 ? p.ksplit : 1);
    else total_tiles = __ldg(p.num_mtiles) * n_ntiles;
    [[maybe_unused]] const int first_tile = blockIdx.x >> 1;   // pair p takes tiles p, p + npairs, ... (or its stream-K span)
    [[maybe_unused]] const int tile_stride = gridDim.x >> 1;
    using Work = WorkIter<BN, WGRAD, kEpiKb>;

    if (warp == 0) {
        // ================================ TMA producer (one thread per CTA) =========================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            [[maybe_unused]] int ti = 0;
            [[maybe_unused]] const uint64_t pol_a = WGRAD ? l2_policy_evict_first() : 0, pol_b = WGRAD ? l2_policy_evict_last() : 0;
            Work work(p, sk, n_ntiles, rank, total_tiles);
            for (TileCoord c; work.next(c); ++ti) {
                MOE_TL(0, ti, 0);
                for (int kb = 0; kb < c.kb; ++kb) {
                    mbar_wait(empty_bar + s, ph ^ 1);
                    if (kb == 0) MOE_TL(0, ti, 1);
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(full_bar + s, 2 * Cfg::STAGE_BYTES);
                    if constexpr (!WGRAD) {
                        tma_load_2d_pair(sa, &tmA, full_bar + s, kb * kBK, c.m0);
                        if constexpr (Cfg::B_MN) {
                            const int krow = c.e * p.K + kb * kBK;   // B = [E * K, N]
#pragma unroll
                            for (int i = 0; i < (BN / 2) / Cfg::B_ATOM; ++i)
                                tma_load_2d_pair(sb + i * (Cfg::B_ATOM * kBK * 2), &tmB, full_bar + s,
                                                 c.n0 + (Cfg::B_ATOM == 64 ? b_block_col<BN>(rank, i) : rank * (BN / 2) + i * Cfg::B_ATOM),
                                                 krow);
                        } else if constexpr (BN <= 256) {
                            tma_load_2d_pair(sb, &tmB, full_bar + s, kb * kBK, c.e * p.N + c.n0 + rank * (BN / 2));
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)   // 64-row boxes
                                tma_load_2d_pair(sb + i * 8192, &tmB, full_bar + s, kb * kBK,
                                                 c.e * p.N + c.n0 + b_block_col<BN>(rank, i));
                        }
                    } else {
                        // Stream-K (only chosen when N is one tile wide and B fits L2, gemm_launch.cu): A (the dU / H columns of
                        // this tile) is read once per launch: evict_first.  B (the expert's X / dY rows) is re-read by every M
                        // tile of the expert, under stream-K at a different time by each of them: evict_last keeps it in L2
                        // (38.7 MB at config 2; measured 240 -> 209 MB of DRAM reads, 71 -> 67 us).  Lock-step schedules
                        // (whole tiles, equal split-K parts) share B in time and measured slower with the hints.
                        const int krow = c.row0 + (c.kb0 + kb) * kBK;
                        if (p.streamk) {
                            tma_load_2d_pair_hint(sa, &tmA, full_bar + s, c.m0, krow, pol_a);
                            tma_load_2d_pair_hint(sa + 8192, &tmA, full_bar + s, c.m0 + 64, krow, pol_a);
#pragma unroll
                            for (int i = 0; i < (BN / 2) / Cfg::B_ATOM; ++i)
                                tma_load_2d_pair_hint(sb + i * (Cfg::B_ATOM * kBK * 2), &tmB, full_bar + s,
                                                      c.n0 + (Cfg::B_ATOM == 64 ? b_block_col<BN>(rank, i) : rank * (BN / 2) + i * Cfg::B_ATOM),
                                                      krow, pol_b);
                        } else {
                            tma_load_2d_pair(sa, &tmA, full_bar + s, c.m0, krow);
                            tma_load_2d_pair(sa + 8192, &tmA, full_bar + s, c.m0 + 64, krow);
#pragma unroll
                            for (int i = 0; i < (BN / 2) / Cfg::B_ATOM; ++i)
                                tma_load_2d_pair(sb + i * (Cfg::B_ATOM * kBK * 2), &tmB, full_bar + s,
                                                 c.n0 + (Cfg::B_ATOM == 64 ? b_block_col<BN>(rank, i) : rank * (BN / 2) + i * Cfg::B_ATOM),
                                                 krow);
                        }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                MOE_TL(0, ti, 2);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (one thread of the leader CTA) =================
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kPairM, BN <= 256 ? BN : 256, WGRAD, Cfg::B_MN);
            [[maybe_unused]] constexpr uint32_t idesc1 = umma_idesc_bf16(kPairM, BN <= 256 ? 16 : BN - 256, WGRAD, Cfg::B_MN);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            [[maybe_unused]] int ti = 0;
            Work work(p, sk, n_ntiles, rank, total_tiles);
            for (TileCoord c; work.next(c); ++ti) {
                if (c.kb == 0) continue;
                MOE_TL(1, ti, 0);
                mbar_wait(tempty_bar + as, aph ^ 1);
                tc_fence_after();
                MOE_TL(1, ti, 1);
                const uint32_t tmem_d = tmem_base + as * 256;   // wide tiles: one stage, columns [0, BN)
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
                        const uint64_t bd = !Cfg::B_MN ? umma_smem_desc(b_addr + k4 * 32, 16, 1024)
                                            : Cfg::B_ATOM == 64 ? umma_smem_desc(b_addr + k4 * 2048, 8192, 1024)
                                                                : umma_smem_desc(b_addr + k4 * 1024, 4096, 512, 4);
                        umma_bf16(tmem_d, ad, bd, idesc, (kb | k4) != 0);
                        if constexpr (Cfg::NSUB == 2)   // second UMMA of the k-step: B blocks past the first 128 rows / columns
                            umma_bf16(tmem_d + 256, ad, bd + (16384 >> 4), idesc1, (kb | k4) != 0);
                    }
                    umma_commit_pair(empty_bar + s);  // frees the smem slot in both CTAs once these MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_pair(tfull_bar + as);  // accumulator complete -> epilogues of both CTAs
                MOE_TL(1, ti, 3);
                if (++as == Cfg::ACC_STAGES) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue (2 groups x 4 warps) ============================
        // Every warp is independent: it owns 32 accumulator rows (its TMEM lane quarter; thread = row), the two groups
        // take alternate 32-column chunks, and each warp stages and TMA-stores its own slabs, so there is no CTA-level
        // barrier anywhere in the epilogue.
        const int q = warp & 3;                  // TMEM lane quarter this warp may touch
        const int grp = (warp - 2) >> 2;         // column group this warp belongs to
        const int ew = warp - 2;                 // 0..7
        float* const wbias = bias_s + ew * Cfg::BIAS_FLOATS;  // this warp's copy of the bias values of its chunks
        int as = 0;
        uint32_t aph = 0;
        [[maybe_unused]] int ti = 0;
        [[maybe_unused]] const int tl_role = warp == 2 ? 2 : 3;
        [[maybe_unused]] const bool tl_on = (warp == 2 || warp == 6) && lane == 0;
        if constexpr (Cfg::F32) {
            // Two half-slabs per warp (16 accumulator columns each, 2 KB), used alternately: the TMA store of one
            // half reads its slab while the next half is loaded from TMEM and staged (one full slab per warp exposed the
            // store's ~600-clock shared-memory read latency once per chunk — with a single 384-column accumulator the
            // epilogue does not overlap the mainloop, so its length is paid in full).
            //   EPI_F32  : slab = 32 rows (m) x 64 B, 64-byte swizzle (unit u of row r at u ^ ((r >> 1) & 3))
            //   EPI_F32_T: slab = 16 rows (n) x 128 B, 128-byte swizzle (row j = accumulator column j, this lane = column)
            uint8_t* const slab0 = staging + ew * Cfg::SLAB_BYTES;
            uint32_t nhalf = 0;   // half-slabs filled so far (parity selects the slab)
            Work work(p, sk, n_ntiles, rank, total_tiles);
            for (TileCoord c; work.next(c); ++ti) {
                const bool live = c.kb != 0;
                if (tl_on) MOE_TL(tl_role, ti, 0);
                // stream-K: this warp's slabs of the tile are also written by the same warp of the pairs that run the other
                // parts.  Part 0 stores and sets the flag to 1; part s waits for the flag to read s, reduce-adds and passes
                // s + 1 on (the last part leaves 0): one store, then the adds in one fixed order — bit-reproducible.
                int* const my_flag = c.part >= 0 ? p.flags + (static_cast<size_t>(c.tile) * 2 + rank) * Cfg::EPI_WARPS + ew : nullptr;
                const int pass_on = c.last ? 0 : c.part + 1;
                if (c.part > 0) {
                    if (lane == 0) {
                        uint32_t spins = 0;
                        while (ld_acquire_gpu(my_flag) != c.part) {
                            __nanosleep(64);
                            if (++spins > (1u << 27)) __trap();
                        }
                        fence_proxy_async_all();
                        if (!live) st_release_gpu(my_flag, pass_on);   // nothing to add: hand the tile on
                    }
                    __syncwarp();
                    if (!live) continue;
                }
                if (live) {
                    mbar_wait(tfull_bar + as, aph);
                    tc_fence_after();
                }
                if (tl_on) MOE_TL(tl_role, ti, 1);
                const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
#pragma unroll 1
                for (int ch = grp; ch < Cfg::NCHUNK; ch += Cfg::NGRP) {
                    const bool last_chunk = (ch + Cfg::NGRP >= Cfg::NCHUNK);
                    uint32_t acc[2][16];
                    if (live) tmem_ld16(tmem_row + ch * 32, acc[0]);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {   // 16 accumulator columns at a time
                        if (live) {
                            tmem_ld_wait();
                            if (hb == 0) {
                                tmem_ld16(tmem_row + ch * 32 + 16, acc[1]);
                            } else if (last_chunk) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                                if (tl_on) MOE_TL(tl_role, ti, 2);
                            }
#pragma unroll
                            for (int r = 0; r < 16; ++r) asm volatile("" : "+r"(acc[hb][r]));
                        } else {
#pragma unroll
                            for (int r = 0; r < 16; ++r) acc[hb][r] = 0u;
                        }
                        uint8_t* const slab = slab0 + (nhalf & 1) * 2048;
                        ++nhalf;
                        if (lane == 0) tma_store_wait_read<1>();   // the store issued two halves ago has left this slab
                        __syncwarp();
                        if constexpr (EPI == EPI_F32) {
                            uint8_t* const my_row = slab + lane * 64;
                            const int sw = (lane >> 1) & 3;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) =
                                    make_uint4(acc[hb][4 * j], acc[hb][4 * j + 1], acc[hb][4 * j + 2], acc[hb][4 * j + 3]);
                        } else {
                            uint8_t* const my_col = slab + (lane & 3) * 4;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                *reinterpret_cast<uint32_t*>(my_col + j * 128 + (((lane >> 2) ^ (j & 7)) << 4)) = acc[hb][j];
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (c.m0 + q * 32 < p.M && c.n0 + ch * 32 + hb * 16 < p.N) {
                                const int o_col = EPI == EPI_F32 ? c.n0 + ch * 32 + hb * 16 : c.m0 + q * 32;
                                const int o_row = EPI == EPI_F32 ? c.m0 + q * 32 : c.n0 + ch * 32 + hb * 16;
                                if (c.part > 0) tma_reduce_add_3d(&tmO0, slab, o_col, o_row, c.e);
                                else tma_store_3d(&tmO0, slab, o_col, o_row, c.e);
                            }
                            tma_store_commit();   // one group per half even when it is empty: wait_group.read 1 counts groups
                        }
                    }
                }
                if (c.part >= 0) {   // this part's stores / adds are complete and visible before the next part is let in
                    if (lane == 0) {
                        tma_store_wait_all<0>();
                        __threadfence();
                        st_release_gpu(my_flag, pass_on);
                    }
                    __syncwarp();
                }
                if (tl_on) MOE_TL(tl_role, ti, 3);
                if (live && ++as == Cfg::ACC_STAGES) { as = 0; aph ^= 1; }
            }
        } else {
            // bf16 outputs: slab rows are 64 B (four 16-byte units), unit u of row r lives at unit u ^ ((r >> 1) & 3)
            // (CU_TENSOR_MAP_SWIZZLE_64B); a quarter-warp then touches all 32 banks exactly once.
            // this warp's slab(s): [output 0][output 1] — or, DGELU, one in-place slab per chunk
            const uint32_t out_s = smem_u32(staging) + ew * (EPI == EPI_DGELU ? Cfg::MAXCH : Cfg::NOUT) * 2048;
            const int sw = (lane >> 1) & 3;
            [[maybe_unused]] uint64_t* const my_aux_bar = aux_bar + ew * Cfg::MAXCH;
            // DGELU: TMA-load this warp's 32 rows x 32 columns of G for its i-th chunk of the tile at `c` into slab i
            [[maybe_unused]] auto issue_aux = [&](const TileCoord& c, int i) {
                mbar_arrive_expect_tx(my_aux_bar + i, 2048);
                tma_load_2d_s(out_s + i * 2048, &tmAux, my_aux_bar + i, c.n0 + (grp + Cfg::NGRP * i) * 32, c.m0 + q * 32);
            };
            [[maybe_unused]] uint32_t it = 0;   // tiles done by this warp
            if constexpr (EPI == EPI_DGELU) {
                if (lane == 0 && first_tile < total_tiles) {
                    const TileCoord c0 = decode_tile<BN, WGRAD>(p, first_tile, n_ntiles, rank);
                    for (int i = 0; grp + Cfg::NGRP * i < Cfg::NCHUNK; ++i) issue_aux(c0, i);
                }
            }
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++ti) {
                const TileCoord c = decode_tile<BN, WGRAD>(p, tile, n_ntiles, rank);
                const bool live = c.kb != 0;
                if (tl_on) MOE_TL(tl_role, ti, 0);
                if constexpr (EPI == EPI_BIAS_GELU_DUAL || EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
                    // lane l fetches 4 consecutive bias values of the group's chunks, four chunks per pass
                    __syncwarp();                                // previous tile's reads of wbias are done
#pragma unroll
                    for (int i0 = 0; i0 < Cfg::MAXCH; i0 += 4) {
                        const int lc = grp + Cfg::NGRP * (i0 + (lane >> 3));   // chunk the lane's values belong to
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (lc < Cfg::NCHUNK)
                            bv = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(c.e) * p.N + c.n0 + lc * 32) + (lane & 7));
                        if (i0 + (lane >> 3) < Cfg::BIAS_FLOATS / 32) *reinterpret_cast<float4*>(wbias + i0 * 32 + lane * 4) = bv;
                    }
                    __syncwarp();
                }
                [[maybe_unused]] TileCoord cn{};                  // DGELU: next tile of this pair, whose G slabs are prefetched
                [[maybe_unused]] bool have_next = false;
                if constexpr (EPI == EPI_DGELU) {
                    have_next = tile + tile_stride < total_tiles;
                    if (have_next) cn = decode_tile<BN, WGRAD>(p, tile + tile_stride, n_ntiles, rank);
                    if (it > 0 && lane == 0) {   // the last slab of the previous tile: its store was the last group committed
                        tma_store_wait_read<0>();
                        issue_aux(c, (Cfg::NCHUNK - 1 - grp) / Cfg::NGRP);
                    }
                }
                if (live) {
                    mbar_wait(tfull_bar + as, aph);
                    tc_fence_after();
                }
                if (tl_on) MOE_TL(tl_role, ti, 1);
                const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
                if (Cfg::NGRP > Cfg::NCHUNK && grp >= Cfg::NCHUNK) {   // a group without chunks (narrow tile) still releases the stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                }
#pragma unroll 1
                for (int i = 0; grp + Cfg::NGRP * i < Cfg::NCHUNK; ++i) {
                    const int ch = grp + Cfg::NGRP * i;
                    const bool last_chunk = (ch + Cfg::NGRP >= Cfg::NCHUNK);
#ifdef MOE_DBG_NO_EPI
                    if (last_chunk) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_bar + as, 0);
                    }
                    continue;
#endif
                    const uint32_t my_out = out_s + (EPI == EPI_DGELU ? i * 2048 : 0) + lane * 64;   // this lane's row of the slab
                    if constexpr (EPI == EPI_DGELU) mbar_wait(my_aux_bar + i, it & 1);                  // G rows of this chunk have landed
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
                                             : "r"(my_out + (((blk * 2 + j) ^ sw) << 4)));
                        }
                        uint32_t o0[8];                       // 16 columns of output 0, packed bf16x2
                        [[maybe_unused]] uint32_t o1[8];      // 16 columns of output 1 (fc1: gelu)
                        epilogue_block16<EPI>(acc[blk], cbias + blk * 64, aux, o0, o1);
                        if constexpr (EPI == EPI_DGELU) {
                            // db1 on the way out: column sums of this warp's 32 rows x 16 columns of dU (the bf16 values
                            // that are stored), by a transposing butterfly — lane l ends with column l >> 1 — written
                            // per 32-row slab; moe_slab_colsum_final adds the slabs of an expert in row order.
                            if (p.colsum != nullptr) {
                                float v[16];
#pragma unroll
                                for (int r = 0; r < 8; ++r) {
                                    v[2 * r] = __uint_as_float(o0[r] << 16);
                                    v[2 * r + 1] = __uint_as_float(o0[r] & 0xffff0000u);
                                }
#pragma unroll
                                for (int lvl = 0; lvl < 4; ++lvl) {
                                    const int off = 16 >> lvl, n = 8 >> lvl;
                                    const bool upper = (lane & off) != 0;
#pragma unroll
                                    for (int j = 0; j < n; ++j) {
                                        const float send = upper ? v[j] : v[j + n];
                                        const float keep = upper ? v[j + n] : v[j];
                                        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                                    }
                                }
                                v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
                                if ((lane & 1) == 0)
                                    p.colsum[static_cast<size_t>((c.m0 + q * 32) >> 5) * p.N + c.n0 + ch * 32 + blk * 16 + (lane >> 1)] = v[0];
                            }
                        }
                        if (EPI != EPI_DGELU && blk == 0) {
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
                        tma_store_2d_s(&tmO0, my_out - lane * 64, c.n0 + ch * 32, c.m0 + q * 32);
                        if constexpr (Cfg::NOUT == 2) tma_store_2d_s(&tmO1, out_s + 2048, c.n0 + ch * 32, c.m0 + q * 32);
                        tma_store_commit();
                        if constexpr (EPI == EPI_DGELU) {
                            // slab i-1 was stored one chunk ago: once that store has read it, refill it for the next tile
                            if (i >= 1 && have_next) {
                                tma_store_wait_read<1>();
                                issue_aux(cn, i - 1);
                            }
                        }
                    }
#endif
                }
                if (tl_on) MOE_TL(tl_role, ti, 3);
                ++it;
                if (live && ++as == Cfg::ACC_STAGES) { as = 0; aph ^= 1; }
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
