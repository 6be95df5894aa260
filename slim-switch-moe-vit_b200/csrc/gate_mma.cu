// gate_mma.cu — the gate projection on the 5th-generation tensor cores (tcgen05 / TMEM), with CERTIFIED routing (sm_100a).
//
// The gate of the MoE layer (FastMoE NaiveGate / SwitchGate / GShardGate: nn.Linear(d, E) + top-k + softmax, reached
// from /root/reference/models/resMoE.py:27-29; attribute path pinned by models/resmoe_flop_hook.py:7) is a skinny
// contraction logits[T, E] = x[T, d] Wg[E, d]^T with E = 8..64.  On the CUDA cores it costs T d E fp32 FMAs (round 1:
// 48 us at the config-2 shape against a 6 us HBM floor, and 4x that at E = 64); a first mma.sync version was bound by
// the per-CTA latency chain (31 us) and, at E = 64, by mma.sync's own rate.  This version is a persistent,
// warp-specialised pipeline — one CTA per SM, token tiles of 128:
//
//   warp 0      TMA producer: 64-feature chunks of x (128 tokens) and of the two weight planes stream through a ring
//               of shared-memory stages (128-byte swizzle), running ahead across tile boundaries;
//   warp 1      one thread issues tcgen05.mma (cta_group::1, M = 128 tokens, N = 2 E_PAD, K = 16) straight from shared
//               memory into a double-buffered TMEM accumulator: columns [0, E_PAD) = x w0^T, [E_PAD, 2 E_PAD) = x w1^T;
//   warps 2-3   two token rows per thread: ||x_t||^2 from the staged chunks (needed by the certification bound);
//   warps 4-11  two epilogue groups of four warps, alternating tiles (each owns one accumulator stage), one thread per
//               token (its TMEM lane): logits, certification, top-k, scores, per-tile histogram and probability sums —
//               while the other warps are already two tiles ahead.  (ncu, round 2: with one group the kernel was bound
//               by the epilogue's instruction issue — four warps, one per scheduler, with dependent chains.)
//
//   * x is bf16 (exact operand).  Wg (fp32) is split once per forward into two bf16 planes w0 = bf16(w),
//     w1 = bf16(w - w0) (|w - w0 - w1| <= 2^-18 |w|); products are exact in fp32, the planes accumulate separately
//     and are added once at the end.
//   * The routing INTEGERS must stay bit-identical to the CPU oracle, whose logits are fp32 FMA chains in a fixed
//     order (LOGIT ORDER v1, oracle/gate_ref.c).  The tensor-core logits differ from those by rounding only, and the
//     difference is bounded per token:  |v_e - L_e| <= B_t = kappa * ||x_t||_2 * max_e ||w_e||_2  (Cauchy-Schwarz on
//     sum_i |x_i w_ei|; kappa covers the accumulation error of both sides and the split residual, see gate_kappa()).
//     A token whose k + 1 largest logits are pairwise further apart than 2 B_t + 16 u max|v| cannot be routed
//     differently by the oracle: its top-k selection AND order are certified.  For every other token (a fraction of a
//     percent, and every exact tie) the experts that can still reach the oracle's top-k — those within the margin of
//     the k-th largest logit — are recomputed by the token's warp in LOGIT ORDER v1 (bit-exact) before the selection
//     is redone.  So idx / counts / positions are bit-exact against the oracle for every token, and every emitted
//     logit is either bit-exact (recomputed candidates) or within B_t of the oracle's.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "ptx.cuh"

namespace moe {

constexpr int kGmTok = 128;       // tokens per CTA = two 64-token routing tiles
constexpr int kGmKC = 64;         // features per pipeline stage (one 128-byte swizzle row)
constexpr int kGmThreads = 384;   // producer, MMA issuer, 2 row-norm warps, 2 x 4 epilogue warps
constexpr int kGmMaxStages = 8;
constexpr int kGmMaxK = 8;

// kappa of the certification bound (see the header): accumulation error of the tensor-core path (d / 16 dependent
// MMAs per accumulator, each assumed accurate to 2^-21 of the magnitudes it adds — several times what fp32-accumulate
// tensor cores are known to lose to truncation), the split residual 2^-18, and the oracle's own fp32 rounding
// ((d / 32 + 5) operations of 2^-24 each, + 3 for the bias / noise adds on either side).
static float gate_kappa(int d) {
    return static_cast<float>((d / 16.0) * ldexp(1.0, -21) + ldexp(1.0, -18) + (d / 32.0 + 8.0) * ldexp(1.0, -24));
}

// ------------------------------------------------------------------------------------------------
// prologue: Wg [E, d] fp32 -> planes [2][E_pad][d] bf16 (rows >= E zero) + wnorm[E_pad] = ||w_e||_2 (rounded up)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_split_kernel(const float* __restrict__ Wg, int E, int E_pad, int d, __nv_bfloat16* __restrict__ planes,
                  float* __restrict__ wnorm) {
    __shared__ float red[8];
    const int e = blockIdx.x;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < d; i += 256) {
        const float w = e < E ? Wg[static_cast<size_t>(e) * d + i] : 0.0f;
        const __nv_bfloat16 w0 = __float2bfloat16_rn(w);
        const __nv_bfloat16 w1 = __float2bfloat16_rn(w - __bfloat162float(w0));
        planes[static_cast<size_t>(e) * d + i] = w0;
        planes[(static_cast<size_t>(E_pad) + e) * d + i] = w1;
        ss = fmaf(w, w, ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[w];
        wnorm[e] = sqrtf(t) * 1.0001f;   // an upper bound whatever the rounding of the sum
    }
}

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Up to four logits of one token in LOGIT ORDER v1 (bit-exact with oracle/gate_ref.c), computed by a whole warp: feature i
// belongs to lane (i / 4) % 32, ascending-i FMA chains from +0, xor butterfly 16..1 (the caller adds bias and noise last).
// Out of line on purpose: it runs for a fraction of a percent of the tokens and must not cost the main path registers or
// code size.  Experts e[j] < 0 are skipped.  The loads of four iterations are independent and issued together.
__device__ __noinline__ float4 gate_exact_dots(const __nv_bfloat16* __restrict__ xr, const float* __restrict__ Wg, int d,
                                               int e0, int e1, int e2, int e3, int lane) {
    const float* w0 = Wg + static_cast<size_t>(max(e0, 0)) * d;
    const float* w1 = Wg + static_cast<size_t>(max(e1, 0)) * d;
    const float* w2 = Wg + static_cast<size_t>(max(e2, 0)) * d;
    const float* w3 = Wg + static_cast<size_t>(max(e3, 0)) * d;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
    for (int c4 = lane; c4 < d / 4; c4 += 32) {
        const uint2 xb = __ldg(reinterpret_cast<const uint2*>(xr + c4 * 4));
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(w0 + c4 * 4));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(w1 + c4 * 4));
        const float4 q2 = __ldg(reinterpret_cast<const float4*>(w2 + c4 * 4));
        const float4 q3 = __ldg(reinterpret_cast<const float4*>(w3 + c4 * 4));
        const float x0 = bf_lo(xb.x), x1 = bf_hi(xb.x), x2 = bf_lo(xb.y), x3 = bf_hi(xb.y);
        a0 = fmaf(x0, q0.x, a0); a0 = fmaf(x1, q0.y, a0); a0 = fmaf(x2, q0.z, a0); a0 = fmaf(x3, q0.w, a0);
        a1 = fmaf(x0, q1.x, a1); a1 = fmaf(x1, q1.y, a1); a1 = fmaf(x2, q1.z, a1); a1 = fmaf(x3, q1.w, a1);
        a2 = fmaf(x0, q2.x, a2); a2 = fmaf(x1, q2.y, a2); a2 = fmaf(x2, q2.z, a2); a2 = fmaf(x3, q2.w, a2);
        a3 = fmaf(x0, q3.x, a3); a3 = fmaf(x1, q3.y, a3); a3 = fmaf(x2, q3.z, a3); a3 = fmaf(x3, q3.w, a3);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        a0 = a0 + __shfl_xor_sync(0xffffffffu, a0, off);
        a1 = a1 + __shfl_xor_sync(0xffffffffu, a1, off);
        a2 = a2 + __shfl_xor_sync(0xffffffffu, a2, off);
        a3 = a3 + __shfl_xor_sync(0xffffffffu, a3, off);
    }
    return make_float4(a0, a1, a2, a3);
}

// top-NP of one token's logits held by ONE thread: descending value, ties -> lowest expert index.  Everything is
// compile-time indexed so that v / pv / pi stay in registers.  NP = 2 and 3 (top-1 and top-2 gates: k + 1 picks) are a
// single insertion pass over the experts; larger NP selects pick by pick.
template <int EP, int NP>
__device__ __forceinline__ void thread_topk(const float (&v)[EP], int E, int npick, float (&pv)[NP], int (&pi)[NP]) {
    if constexpr (NP <= 3) {
#pragma unroll
        for (int p = 0; p < NP; ++p) { pv[p] = -INFINITY; pi[p] = 0x7fffffff; }
#pragma unroll
        for (int e = 0; e < EP; ++e) {
            if (e < E) {   // ascending e + strict comparisons: equal values keep the lower index in front
                const float x = v[e];
                if (pi[0] == 0x7fffffff || x > pv[0]) {
                    if constexpr (NP == 3) { pv[2] = pv[1]; pi[2] = pi[1]; }
                    pv[1] = pv[0]; pi[1] = pi[0];
                    pv[0] = x; pi[0] = e;
                } else if (pi[1] == 0x7fffffff || x > pv[1]) {
                    if constexpr (NP == 3) { pv[2] = pv[1]; pi[2] = pi[1]; }
                    pv[1] = x; pi[1] = e;
                } else if constexpr (NP == 3) {
                    if (pi[2] == 0x7fffffff || x > pv[2]) { pv[2] = x; pi[2] = e; }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (p >= npick) { pv[p] = -INFINITY; pi[p] = 0x7fffffff; }
        return;
    }
    uint64_t used = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float bv = -INFINITY;
        int be = 0x7fffffff;
        if (p < npick) {
#pragma unroll
            for (int e = 0; e < EP; ++e) {
                const bool ok = e < E && !((used >> e) & 1ull);
                if (ok && (be == 0x7fffffff || v[e] > bv)) { bv = v[e]; be = e; }
            }
            if (be != 0x7fffffff) used |= 1ull << be;
        }
        pv[p] = bv;
        pi[p] = be;
    }
}

// lane l returns sum over the warp's lanes of their v[l] (fixed xor-butterfly order); v is destroyed
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int off = 16 >> lvl;
        const int n = 16 >> lvl;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = upper ? v[j] : v[j + n];
            const float keep = upper ? v[j + n] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int EP, int NP>   // EP: padded expert count (16 / 32 / 64); NP: picks held in registers (>= min(k + 1, E))
__global__ void __launch_bounds__(kGmThreads, 1)
gate_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __nv_bfloat16* __restrict__ x, const float* __restrict__ Wg, const float* __restrict__ bg,
                     const float* __restrict__ noise, const uint8_t* __restrict__ token_mask, const float* __restrict__ wnorm,
                     float kappa, int64_t T, int d, int E, int k, int score_mode, int want_psum, int ntiles, int nstages,
                     float* __restrict__ logits, int* __restrict__ idx, float* __restrict__ score, int* __restrict__ tile_hist,
                     float* __restrict__ tile_psum) {
    constexpr int NCOL = 2 * EP;                      // accumulator columns: both planes
    constexpr int X_BYTES = kGmTok * 128;             // 128 token rows x 64 bf16
    constexpr int W_BYTES = NCOL * 128;               // both planes, EP rows each
    constexpr int STAGE_BYTES = X_BYTES + W_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * NCOL < 32 ? 32 : 2 * NCOL;   // two accumulator stages (a power of two: 64 / 128 / 256)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + nstages * STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);              // [8]
    uint64_t* empty_bar = full_bar + kGmMaxStages;                       // [8]
    uint64_t* tfull_bar = empty_bar + kGmMaxStages;                      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                                // [2]
    uint64_t* ssfull_bar = tempty_bar + 2;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ssfull_bar + 2);   // [1] (+ pad)
    float* ss_s = reinterpret_cast<float*>(tail + 256);                  // [2][128] squared row norms per accumulator stage
    float* bias_s = ss_s + 2 * kGmTok;                                   // [EP]
    int* hist_s = reinterpret_cast<int*>(bias_s + EP);                   // [2 groups][2 parities][2 halves][EP]
    float* psum_s = reinterpret_cast<float*>(hist_s + 8 * EP);           // [2 groups][2 parities][4 quarters][EP]
    float* exact_s = psum_s + 16 * EP;                                   // [8 warps][EP]
    float* wmax_s = exact_s + 8 * EP;                                    // [1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = d / kGmKC;
    const int ntiles128 = (ntiles + 1) >> 1;

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full_bar + s, 1);       // producer's arrive.expect_tx
            mbar_init(empty_bar + s, 3);      // tcgen05.commit + the two row-norm warps
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + a, 1);      // tcgen05.commit
            mbar_init(tempty_bar + a, 4);     // the four warps of the epilogue group that owns the stage
            mbar_init(ssfull_bar + a, 2);     // the two row-norm warps
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 8 * EP; i += kGmThreads) hist_s[i] = 0;
    for (int i = tid; i < EP; i += kGmThreads) bias_s[i] = (bg != nullptr && i < E) ? __ldg(bg + i) : 0.0f;
    if (warp == 2) {
        float m = 0.0f;
        for (int e = lane; e < E; e += 32) m = fmaxf(m, __ldg(wnorm + e));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) *wmax_s = m;
    }
    if (warp == 1) {
        tmem_alloc1(tmem_slot, TMEM_COLS);
        tmem_relinquish1();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ===========================================
        if (lane == 0) {
            tma_prefetch_desc(&tmX);
            tma_prefetch_desc(&tmW);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles128; tile += gridDim.x) {
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(empty_bar + s, ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(full_bar + s, STAGE_BYTES);
                    tma_load_2d(st, &tmX, full_bar + s, kc * kGmKC, tile * kGmTok);   // rows past T: zero-filled
                    tma_load_2d(st + X_BYTES, &tmW, full_bar + s, kc * kGmKC, 0);
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (one thread) ================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kGmTok, NCOL, false, false);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            for (int tile = blockIdx.x; tile < ntiles128; tile += gridDim.x) {
                mbar_wait(tempty_bar + as, aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * NCOL;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(full_bar + s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + X_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < kGmKC / 16; ++k4)
                        umma_bf16_1(tmem_d, umma_smem_desc(a_addr + k4 * 32, 16, 1024), umma_smem_desc(b_addr + k4 * 32, 16, 1024),
                                    idesc, (kc | k4) != 0);
                    umma_commit1(empty_bar + s);
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
                umma_commit1(tfull_bar + as);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else if (warp < 4) {
        // ================================ row norms: two token rows per thread =====================
        const int r0 = (warp - 2) * 64 + lane, r1 = r0 + 32;
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < ntiles128; tile += gridDim.x) {
            float sa = 0.0f, sb = 0.0f;
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait(full_bar + s, ph);
                const uint8_t* xa = smem + s * STAGE_BYTES + r0 * 128;
                const uint8_t* xb = smem + s * STAGE_BYTES + r1 * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {   // 128-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ (r & 7); r1 & 7 == r0 & 7
                    const uint4 q = *reinterpret_cast<const uint4*>(xa + ((c ^ (r0 & 7)) << 4));
                    const uint4 u = *reinterpret_cast<const uint4*>(xb + ((c ^ (r0 & 7)) << 4));
                    sa = fmaf(bf_lo(q.x), bf_lo(q.x), sa); sa = fmaf(bf_hi(q.x), bf_hi(q.x), sa);
                    sb = fmaf(bf_lo(u.x), bf_lo(u.x), sb); sb = fmaf(bf_hi(u.x), bf_hi(u.x), sb);
                    sa = fmaf(bf_lo(q.y), bf_lo(q.y), sa); sa = fmaf(bf_hi(q.y), bf_hi(q.y), sa);
                    sb = fmaf(bf_lo(u.y), bf_lo(u.y), sb); sb = fmaf(bf_hi(u.y), bf_hi(u.y), sb);
                    sa = fmaf(bf_lo(q.z), bf_lo(q.z), sa); sa = fmaf(bf_hi(q.z), bf_hi(q.z), sa);
                    sb = fmaf(bf_lo(u.z), bf_lo(u.z), sb); sb = fmaf(bf_hi(u.z), bf_hi(u.z), sb);
                    sa = fmaf(bf_lo(q.w), bf_lo(q.w), sa); sa = fmaf(bf_hi(q.w), bf_hi(q.w), sa);
                    sb = fmaf(bf_lo(u.w), bf_lo(u.w), sb); sb = fmaf(bf_hi(u.w), bf_hi(u.w), sb);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar + s);
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
            mbar_wait(tempty_bar + as, aph ^ 1);     // the epilogue has read this stage's norms of two tiles ago
            ss_s[as * kGmTok + r0] = sa;
            ss_s[as * kGmTok + r1] = sb;
            __syncwarp();
            if (lane == 0) mbar_arrive(ssfull_bar + as);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    } else {
        // ================================ epilogue: one thread per token (TMEM lane) ==============
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int grp = (warp - 4) >> 2;              // epilogue group: tiles it = grp, grp + 2, ... of this CTA; accumulator stage grp
        const int r = q * 32 + lane;                  // token row inside the tile
        const int et = tid - (4 + 4 * grp) * 32;      // 0..127 inside the group
        const float wmax = *wmax_s;
        const int npick = min(k + 1, E);
        const bool need_p = (score_mode == 1) || want_psum;
        const int as = grp;
        int it = 0;
        uint32_t aph = 0;
        int* const hist_g = hist_s + grp * 4 * EP;
        float* const psum_g = psum_s + grp * 8 * EP;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles128; tile += 2 * gridDim.x, ++it) {
            const int par = it & 1;
            const int64_t tok = static_cast<int64_t>(tile) * kGmTok + r;
            const bool in_range = tok < T;
            const bool masked = in_range && token_mask != nullptr && token_mask[tok] == 0;
            mbar_wait(tfull_bar + as, aph);
            mbar_wait(ssfull_bar + as, aph);
            tc_fence_after();
            float v[EP];
            const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * NCOL;
#pragma unroll
            for (int c = 0; c < EP / 16; ++c) {
                uint32_t hi[16], lo[16];
                tmem_ld16(trow + c * 16, hi);
                tmem_ld16(trow + EP + c * 16, lo);
                tmem_ld_wait();
                // tcgen05.ld is asynchronous: its destination registers are valid only after the wait.  Without this pin the
                // compiler may read them — or hand them to something else — before the data lands (seen as an illegal address
                // at E = 64 once the surrounding code changed the register allocation)
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(hi[i]), "+r"(lo[i]));
#pragma unroll
                for (int i = 0; i < 16; ++i) v[c * 16 + i] = __uint_as_float(hi[i]) + __uint_as_float(lo[i]);
            }
            const float ss = ss_s[as * kGmTok + r];
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + as);   // accumulator stage and its norms are free again
            aph ^= 1;

            float vmax = 0.0f;
#pragma unroll
            for (int e = 0; e < EP; ++e) {
                if (e < E) {
                    float val = v[e] + bias_s[e];
                    if (noise != nullptr && in_range) val += __ldg(noise + tok * E + e);
                    v[e] = val;
                    vmax = fmaxf(vmax, fabsf(val));
                }
            }
            float pv[NP];
            int pi[NP];
            thread_topk<EP, NP>(v, E, npick, pv, pi);
            const float margin = 2.0f * kappa * sqrtf(ss) * wmax + 9.5367431640625e-07f * vmax;   // 16 u max|v|
            bool amb = false;
#pragma unroll
            for (int p = 0; p + 1 < NP; ++p)
                if (p + 1 < npick) amb = amb || !(pv[p] - pv[p + 1] > margin);
            amb = amb && in_range && !masked;

            // ---- uncertified tokens.  The oracle's top-k can only contain experts whose tensor-core logit lies within
            // `margin` (> 2 B_t) of the k-th largest one: every other expert is provably below k exact values.  Only those
            // candidates (typically k + 1 of them) are recomputed — by the whole warp, in LOGIT ORDER v1, bit-exact with the
            // oracle — and the selection is redone on the corrected values.
            uint64_t cand = 0;
            if (amb) {
                float kth = pv[0];
#pragma unroll
                for (int p = 1; p < NP; ++p)
                    if (p < k) kth = pv[p];
                const float lim = kth - margin;
#pragma unroll
                for (int e = 0; e < EP; ++e)
                    if (e < E && v[e] >= lim) cand |= 1ull << e;
            }
            unsigned flagged = __ballot_sync(0xffffffffu, amb);
            if (flagged) {
                float* ex = exact_s + (warp - 4) * EP;
                while (flagged) {
                    const int src = __ffs(flagged) - 1;
                    flagged &= flagged - 1;
                    const int64_t tk = static_cast<int64_t>(tile) * kGmTok + q * 32 + src;
                    uint32_t c_lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(cand), src);
                    uint32_t c_hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(cand >> 32), src);
                    uint64_t cm = (static_cast<uint64_t>(c_hi) << 32) | c_lo;
                    while (cm) {
                        int e4[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            e4[j] = cm ? __ffsll(static_cast<long long>(cm)) - 1 : -1;
                            if (cm) cm &= cm - 1;
                        }
                        const float4 dots = gate_exact_dots(x + tk * d, Wg, d, e4[0], e4[1], e4[2], e4[3], lane);
                        if (lane == 0) {
                            const float dv[4] = {dots.x, dots.y, dots.z, dots.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (e4[j] >= 0) {
                                    float val = dv[j] + (bg != nullptr ? __ldg(bg + e4[j]) : 0.0f);
                                    if (noise != nullptr) val += __ldg(noise + tk * E + e4[j]);
                                    ex[e4[j]] = val;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == src) {
#pragma unroll
                        for (int e = 0; e < EP; ++e)
                            if ((cand >> e) & 1ull) v[e] = ex[e];
                    }
                    __syncwarp();
                }
                if (amb) thread_topk<EP, NP>(v, E, npick, pv, pi);
            }

            // ---- outputs: logits, idx, score, per-tile histogram, per-tile probability sums
            if (in_range) {
                float* lrow = logits + tok * E;
                if ((E & 3) == 0) {
#pragma unroll
                    for (int e = 0; e < EP; e += 4)
                        if (e < E) *reinterpret_cast<float4*>(lrow + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < EP; ++e)
                        if (e < E) lrow[e] = v[e];
                }
            }
            const bool live = in_range && !masked;
            const float m = pv[0];
            float rz = 0.0f;
            if (need_p) {
                float z = 0.0f;
#pragma unroll
                for (int e = 0; e < EP; ++e) {
                    // v now holds the un-normalised probabilities.  ex2.approx (2^-22 relative) is enough for the column sums and
                    // the normaliser; the k selected scores below use expf
                    v[e] = e < E ? ex2_approx((v[e] - m) * 1.4426950408889634f) : 0.0f;
                    z += v[e];
                }
                rz = 1.0f / z;
                if (want_psum) {
#pragma unroll
                    for (int g0 = 0; g0 < EP; g0 += 32) {
                        float cols[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) cols[i] = (g0 + i < EP && live) ? v[g0 + i < EP ? g0 + i : 0] * rz : 0.0f;
                        const float colsum = warp_transpose_sum32(cols, lane);
                        if (g0 + lane < EP) psum_g[(par * 4 + q) * EP + g0 + lane] = colsum;
                    }
                }
            }
            if (in_range) {
                int* irow = idx + tok * k;
                float* srow = score + tok * k;
                if (masked) {
                    for (int p = 0; p < k; ++p) { irow[p] = -1; srow[p] = 0.0f; }
                } else {
                    float ssum = 0.0f;
                    if (score_mode == 0) {
#pragma unroll
                        for (int p = 0; p < NP; ++p)
                            if (p < k) ssum += expf(pv[p] - m);
                    }
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        if (p < k) {
                            const float w = expf(pv[p] - m);
                            irow[p] = pi[p];
                            srow[p] = score_mode == 0 ? w / ssum : w * rz;
                            atomicAdd(hist_g + (par * 2 + (q >> 1)) * EP + pi[p], 1);
                        }
                    }
                }
            }
            named_bar_sync(2 + grp, 128);   // the group's four warps: this tile's histogram and column sums are complete
            for (int i = et; i < 2 * EP; i += 128) {
                const int hf = i / EP, e = i - hf * EP;
                const int rt = tile * 2 + hf;             // 64-token routing tile
                if (e < E && rt < ntiles) {
                    tile_hist[static_cast<size_t>(e) * ntiles + rt] = hist_g[par * 2 * EP + i];
                    if (want_psum)
                        tile_psum[static_cast<size_t>(e) * ntiles + rt] =
                            psum_g[(par * 4 + 2 * hf) * EP + e] + psum_g[(par * 4 + 2 * hf + 1) * EP + e];
                }
                hist_g[par * 2 * EP + i] = 0;             // reused two of the group's tiles later, behind its next barrier
            }
        }
    }

    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc1(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int gate_e_pad(int E) { return E <= 16 ? 16 : E <= 32 ? 32 : 64; }

bool gate_mma_supported(int x_dtype, int d, int E) { return x_dtype == MOE_DTYPE_BF16 && E <= 64 && d % 64 == 0 && d >= 64; }

size_t gate_fwd_workspace_bytes(int d, int E) {
    if (E > 64) return 256;
    const size_t ep = static_cast<size_t>(gate_e_pad(E));
    return (2 * ep * d * 2 + ep * 4 + 255) / 256 * 256;
}

template <int EP, int NP>
static int launch_gate_mma_t(const CUtensorMap& tX, const CUtensorMap& tW, const __nv_bfloat16* x, const float* Wg, const float* bg,
                             const float* noise, const uint8_t* token_mask, const float* wnorm, int64_t T, int d, int E, int k,
                             int score_mode, int want_psum, float* logits, int* idx, float* score, int* tile_hist,
                             float* tile_psum, cudaStream_t st) {
    constexpr int STAGE = kGmTok * 128 + 2 * EP * 128;
    const int ntiles = static_cast<int>((T + MOE_TOKEN_TILE - 1) / MOE_TOKEN_TILE);
    const int ntiles128 = (ntiles + 1) / 2;
    int nstages = (190 * 1024) / STAGE;   // one CTA per SM: the ring runs ahead across tile boundaries
    if (nstages > kGmMaxStages) nstages = kGmMaxStages;
    const size_t smem = 1024 + static_cast<size_t>(nstages) * STAGE + 256 + (2 * kGmTok + EP + 8 * EP + 16 * EP + 8 * EP + 4) * 4;
    auto kfn = gate_fwd_umma_kernel<EP, NP>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) { set_error("gate_fwd_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int grid = ntiles128 < sm_count() ? ntiles128 : sm_count();
    kfn<<<grid, kGmThreads, smem, st>>>(tX, tW, x, Wg, bg, noise, token_mask, wnorm, gate_kappa(d), T, d, E, k, score_mode,
                                        want_psum, ntiles, nstages, logits, idx, score, tile_hist, tile_psum);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("gate_fwd_mma launch: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

int launch_gate_fwd_mma(const void* x, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask,
                        int64_t T, int d, int E, int k, int score_mode, int want_psum, float* logits, int* idx, float* score,
                        int* tile_hist, float* tile_psum, void* workspace, cudaStream_t st) {
    const int ep = gate_e_pad(E);
    __nv_bfloat16* planes = static_cast<__nv_bfloat16*>(workspace);
    float* wnorm = reinterpret_cast<float*>(planes + static_cast<size_t>(2) * ep * d);
    gate_split_kernel<<<ep, 256, 0, st>>>(Wg, E, ep, d, planes, wnorm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("gate_split launch: %s", cudaGetErrorString(e)); return 1; }
    CUtensorMap tX, tW;
    if (!encode_tmap_2d_bf16(&tX, x, static_cast<uint64_t>(d), static_cast<uint64_t>(T), 64, kGmTok)) return 1;
    if (!encode_tmap_2d_bf16(&tW, planes, static_cast<uint64_t>(d), static_cast<uint64_t>(2 * ep), 64, static_cast<uint32_t>(2 * ep))) return 1;
    auto xb = static_cast<const __nv_bfloat16*>(x);
    // picks kept in registers: top-(k + 1); the common gates (k = 1, 2) get their own instantiation
#define MOE_GATE_LAUNCH(EP_, NP_) \
    return launch_gate_mma_t<EP_, NP_>(tX, tW, xb, Wg, bg, noise, token_mask, wnorm, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, tile_psum, st)
    const int np = k + 1 <= 2 ? 2 : k + 1 <= 3 ? 3 : kGmMaxK + 1;
    if (ep == 16) { if (np == 2) MOE_GATE_LAUNCH(16, 2); if (np == 3) MOE_GATE_LAUNCH(16, 3); MOE_GATE_LAUNCH(16, kGmMaxK + 1); }
    if (ep == 32) { if (np == 2) MOE_GATE_LAUNCH(32, 2); if (np == 3) MOE_GATE_LAUNCH(32, 3); MOE_GATE_LAUNCH(32, kGmMaxK + 1); }
    if (np == 2) MOE_GATE_LAUNCH(64, 2);
    if (np == 3) MOE_GATE_LAUNCH(64, 3);
    MOE_GATE_LAUNCH(64, kGmMaxK + 1);
#undef MOE_GATE_LAUNCH
}

}  // namespace moe
