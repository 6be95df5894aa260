// gate_mma.cu — the gate projection on the tensor cores, with CERTIFIED routing (sm_100a).
//
// The gate of the MoE layer (FastMoE NaiveGate / SwitchGate / GShardGate: nn.Linear(d, E) + top-k + softmax, reached
// from /root/reference/models/resMoE.py:27-29; attribute path pinned by models/resmoe_flop_hook.py:7) is a skinny
// contraction logits[T, E] = x[T, d] Wg[E, d]^T with E = 8..64.  On the CUDA cores it costs T d E fp32 FMAs — 4x the
// time at E = 64 of what it costs at E = 16, and far above the HBM time of reading x once (round 1: 48 us at the
// config-2 shape against a 6 us HBM floor).  Here it runs on mma.sync m16n8k16 (bf16 x bf16 -> fp32):
//
//   * x is bf16 (exact operand).  Wg (fp32) is split once per forward into two bf16 planes w0 = bf16(w),
//     w1 = bf16(w - w0) (|w - w0 - w1| <= 2^-18 |w|); products are exact in fp32, the planes go to two accumulators
//     (hi, lo) that are added once at the end.
//   * The routing INTEGERS must stay bit-identical to the CPU oracle, whose logits are fp32 FMA chains in a fixed
//     order (LOGIT ORDER v1, oracle/gate_ref.c).  The tensor-core logits differ from those by rounding only, and the
//     difference is bounded per token:  |v_e - L_e| <= B_t = kappa * ||x_t||_2 * max_e ||w_e||_2  (Cauchy-Schwarz on
//     sum_i |x_i w_ei|; kappa covers the accumulation error of both sides and the split residual, see gate_kappa()).
//     A token whose k + 1 largest logits are pairwise further apart than 2 B_t + 16 u max|v| cannot be routed
//     differently by the oracle: its top-k selection AND order are certified.  Every other token (a fraction of a
//     percent, and every exact tie) is recomputed by its warp in LOGIT ORDER v1 — bit-exact logits — before top-k.
//     So idx / counts / positions are bit-exact against the oracle for every token, and the emitted logits are
//     either bit-exact (recomputed tokens) or within B_t of the oracle's (certified tokens).
//   * top-k runs on 4 lanes per token (the quad that holds the token's accumulator row), scores / softmax sums in
//     registers; per-64-token-tile histograms and probability sums leave through shared memory in a fixed order.
//
// Pipeline: one producer warp streams 64-feature chunks of x (128 tokens) and of the two weight planes through a
// ring of shared-memory stages with TMA (128-byte swizzle, read back with ldmatrix), 8 MMA warps of 16 tokens each.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "ptx.cuh"

namespace moe {

constexpr int kGmTok = 128;       // tokens per CTA = two 64-token routing tiles
constexpr int kGmKC = 64;         // features per pipeline stage (one 128-byte swizzle row)
constexpr int kGmThreads = 288;   // 8 MMA warps + 1 producer warp
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

__device__ __forceinline__ void ldsm_x4(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// top-`npick` of one token's logits held by a quad (lane t of the quad owns experts 8j + 2t + c): descending value,
// ties -> lowest expert index.  v is not modified; `pv` / `pi` are returned in every lane of the quad.
template <int NT>
__device__ __forceinline__ void quad_topk(const float (&v)[NT][2], int t, int E, int npick, float (&pv)[kGmMaxK + 1],
                                          int (&pi)[kGmMaxK + 1]) {
    uint32_t used = 0;
    for (int p = 0; p < npick; ++p) {
        float bv = -INFINITY;
        int be = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int e = 8 * j + 2 * t + c;
                const bool ok = e < E && !((used >> (2 * j + c)) & 1u);
                if (ok && (be == 0x7fffffff || v[j][c] > bv)) { bv = v[j][c]; be = e; }
            }
        }
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oe = __shfl_xor_sync(0xffffffffu, be, off);
            if (oe != 0x7fffffff && (be == 0x7fffffff || ov > bv || (ov == bv && oe < be))) { bv = ov; be = oe; }
        }
        pv[p] = bv;
        pi[p] = be;
        if (be != 0x7fffffff && ((be & 7) >> 1) == t) used |= 1u << (2 * (be >> 3) + (be & 1));
    }
}

template <int NT>
__global__ void __launch_bounds__(kGmThreads, 2)
gate_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __nv_bfloat16* __restrict__ x, const float* __restrict__ Wg, const float* __restrict__ bg,
                    const float* __restrict__ noise, const uint8_t* __restrict__ token_mask, const float* __restrict__ wnorm,
                    float kappa, int64_t T, int d, int E, int k, int score_mode, int want_psum, int ntiles, int nstages,
                    float* __restrict__ logits, int* __restrict__ idx, float* __restrict__ score, int* __restrict__ tile_hist,
                    float* __restrict__ tile_psum) {
    constexpr int E_PAD = NT * 8;
    constexpr int X_BYTES = kGmTok * 128;            // 128 token rows x 64 bf16
    constexpr int W_BYTES = 2 * E_PAD * 128;         // both planes, E_PAD rows each
    constexpr int STAGE_BYTES = X_BYTES + W_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + nstages * STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);        // [nstages]
    uint64_t* empty_bar = full_bar + 8;                            // [nstages]
    int* hist_s = reinterpret_cast<int*>(tail + 128);              // [2][E_PAD]
    float* psum_s = reinterpret_cast<float*>(hist_s + 2 * E_PAD);  // [8 warps][E_PAD]
    float* exact_s = psum_s + 8 * E_PAD;                           // [8 warps][E_PAD]
    float* wmax_s = exact_s + 8 * E_PAD;                           // [1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kGmTok;
    const int nk = d / kGmKC;

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 8);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * E_PAD; i += kGmThreads) hist_s[i] = 0;
    if (warp == 0) {
        float m = 0.0f;
        for (int e = lane; e < E; e += 32) m = fmaxf(m, __ldg(wnorm + e));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) *wmax_s = m;
    }
    __syncthreads();

    if (warp == 8) {
        // ================================ producer: TMA, one elected lane ==========================
        if (lane == 0) {
            tma_prefetch_desc(&tmX);
            tma_prefetch_desc(&tmW);
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % nstages;
                if (kc >= nstages) mbar_wait(empty_bar + s, ((kc / nstages) - 1) & 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_arrive_expect_tx(full_bar + s, STAGE_BYTES);
                tma_load_2d(st, &tmX, full_bar + s, kc * kGmKC, static_cast<int>(t_base));   // rows past T: zero-filled
                tma_load_2d(st + X_BYTES, &tmW, full_bar + s, kc * kGmKC, 0);
            }
        }
        return;
    }

    // ==================================== 8 MMA warps, 16 tokens each ==============================
    const int g = lane >> 2, t = lane & 3;
    float hi[NT][4], lo[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { hi[j][c] = 0.0f; lo[j][c] = 0.0f; }
    }
    float ss0 = 0.0f, ss1 = 0.0f;   // partial sum of squares of token rows g and g + 8 (this lane's k columns)
    // ldmatrix lane addresses (128-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ (r & 7))
    const int a_row = warp * 16 + (lane & 15);            // A: lanes 0-15 rows 0-15 at k-chunk 2 ks, lanes 16-31 at 2 ks + 1
    const int a_kc = lane >> 4;
    const int b_row = ((lane >> 4) << 3) + (lane & 7);    // B: matrices (n 0-7, k lo), (n 0-7, k hi), (n 8-15, k lo), (n 8-15, k hi)
    const int b_kc = (lane >> 3) & 1;
    for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % nstages;
        mbar_wait(full_bar + s, (kc / nstages) & 1);
        const uint32_t xs = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t ws = xs + X_BYTES;
#pragma unroll
        for (int ks = 0; ks < kGmKC / 16; ++ks) {
            uint32_t a[4];
            ldsm_x4(xs + a_row * 128 + (((2 * ks + a_kc) ^ (a_row & 7)) << 4), a);
            ss0 = fmaf(bf_lo(a[0]), bf_lo(a[0]), ss0); ss0 = fmaf(bf_hi(a[0]), bf_hi(a[0]), ss0);
            ss0 = fmaf(bf_lo(a[2]), bf_lo(a[2]), ss0); ss0 = fmaf(bf_hi(a[2]), bf_hi(a[2]), ss0);
            ss1 = fmaf(bf_lo(a[1]), bf_lo(a[1]), ss1); ss1 = fmaf(bf_hi(a[1]), bf_hi(a[1]), ss1);
            ss1 = fmaf(bf_lo(a[3]), bf_lo(a[3]), ss1); ss1 = fmaf(bf_hi(a[3]), bf_hi(a[3]), ss1);
#pragma unroll
            for (int jp = 0; jp < NT / 2; ++jp) {
                const int r0 = jp * 16 + b_row;             // plane 0 row
                const int r1 = E_PAD + r0;                  // plane 1 row
                uint32_t b[4];
                ldsm_x4(ws + r0 * 128 + (((2 * ks + b_kc) ^ (r0 & 7)) << 4), b);
                mma_bf16_16816(hi[2 * jp], a, b[0], b[1]);
                mma_bf16_16816(hi[2 * jp + 1], a, b[2], b[3]);
                ldsm_x4(ws + r1 * 128 + (((2 * ks + b_kc) ^ (r1 & 7)) << 4), b);
                mma_bf16_16816(lo[2 * jp], a, b[0], b[1]);
                mma_bf16_16816(lo[2 * jp + 1], a, b[2], b[3]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + s);
    }

    // ---- epilogue: everything stays in the quad that owns the token row
    ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
    ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
    const float wmax = *wmax_s;
    const int npick = min(k + 1, E);
    float v[2][NT][2];
    float pv[2][kGmMaxK + 1];
    int pi[2][kGmMaxK + 1];
    bool amb[2];
    int64_t tok[2];
    bool in_range[2], masked[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        tok[rr] = t_base + warp * 16 + g + rr * 8;
        in_range[rr] = tok[rr] < T;
        masked[rr] = in_range[rr] && token_mask != nullptr && token_mask[tok[rr]] == 0;
        float vmax = 0.0f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int e = 8 * j + 2 * t + c;
                float val = hi[j][rr * 2 + c] + lo[j][rr * 2 + c];
                if (e < E) {
                    if (bg != nullptr) val += __ldg(bg + e);
                    if (noise != nullptr && in_range[rr]) val += __ldg(noise + tok[rr] * E + e);
                    vmax = fmaxf(vmax, fabsf(val));
                }
                v[rr][j][c] = val;
            }
        }
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 1));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 2));
        quad_topk<NT>(v[rr], t, E, npick, pv[rr], pi[rr]);
        const float margin = 2.0f * kappa * sqrtf(rr == 0 ? ss0 : ss1) * wmax + 9.5367431640625e-07f * vmax;   // 16 u max|v|
        bool a = false;
        for (int p = 0; p + 1 < npick; ++p) a = a || !(pv[rr][p] - pv[rr][p + 1] > margin);
        amb[rr] = a && in_range[rr] && !masked[rr];
    }

    // ---- uncertified tokens: the warp recomputes the token's logits in LOGIT ORDER v1 (bit-exact with the oracle)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        unsigned flagged = __ballot_sync(0xffffffffu, amb[rr] && t == 0);   // one bit per ambiguous row (lane 4 g)
        while (flagged) {
            const int src = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            const int64_t tk = t_base + warp * 16 + (src >> 2) + rr * 8;
            const __nv_bfloat16* xr = x + tk * d;
            float* ex = exact_s + warp * E_PAD;
            for (int e0 = 0; e0 < E; e0 += 16) {
                float acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
                for (int c4 = lane; c4 < d / 4; c4 += 32) {   // feature i belongs to lane (i / 4) % 32, ascending i
                    const uint2 xb = __ldg(reinterpret_cast<const uint2*>(xr + c4 * 4));
                    const float x0 = bf_lo(xb.x), x1 = bf_hi(xb.x), x2 = bf_lo(xb.y), x3 = bf_hi(xb.y);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (e0 + i < E) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e0 + i) * d + c4 * 4));
                            float a_ = acc[i];
                            a_ = fmaf(x0, w.x, a_); a_ = fmaf(x1, w.y, a_); a_ = fmaf(x2, w.z, a_); a_ = fmaf(x3, w.w, a_);
                            acc[i] = a_;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p = acc[i];
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, off);
                    if (lane == 0 && e0 + i < E) {
                        float val = p + (bg != nullptr ? __ldg(bg + e0 + i) : 0.0f);
                        if (noise != nullptr) val += __ldg(noise + tk * E + e0 + i);
                        ex[e0 + i] = val;
                    }
                }
            }
            __syncwarp();
            if ((lane >> 2) == (src >> 2)) {   // the quad that owns the row takes the exact values
#pragma unroll
                for (int j = 0; j < NT; ++j) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int e = 8 * j + 2 * t + c;
                        if (e < E) v[rr][j][c] = ex[e];
                    }
                }
            }
            __syncwarp();
        }
        if (__any_sync(0xffffffffu, amb[rr])) quad_topk<NT>(v[rr], t, E, npick, pv[rr], pi[rr]);   // warp-uniform branch
    }

    // ---- outputs: logits, idx, score, per-tile histogram, per-tile probability sums
    const int half = warp >> 2;                       // routing tile of this warp inside the CTA
    const bool need_p = (score_mode == 1) || want_psum;
    float pcol[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { pcol[j][0] = 0.0f; pcol[j][1] = 0.0f; }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        if (in_range[rr]) {
            float* lrow = logits + tok[rr] * E;
            if ((E & 1) == 0) {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    if (8 * j + 2 * t < E) *reinterpret_cast<float2*>(lrow + 8 * j + 2 * t) = make_float2(v[rr][j][0], v[rr][j][1]);
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (8 * j + 2 * t < E) lrow[8 * j + 2 * t] = v[rr][j][0];
                    if (8 * j + 2 * t + 1 < E) lrow[8 * j + 2 * t + 1] = v[rr][j][1];
                }
            }
        }
        const bool live = in_range[rr] && !masked[rr];
        const float m = pv[rr][0];
        float z = 0.0f, rz = 0.0f;
        if (need_p) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (8 * j + 2 * t + c < E) z += expf(v[rr][j][c] - m);
            }
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            rz = 1.0f / z;
            if (want_psum && live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (8 * j + 2 * t + c < E) pcol[j][c] += expf(v[rr][j][c] - m) / z;
                }
            }
        }
        if (t == 0 && in_range[rr]) {
            if (masked[rr]) {
                for (int p = 0; p < k; ++p) { idx[tok[rr] * k + p] = -1; score[tok[rr] * k + p] = 0.0f; }
            } else {
                float w[kGmMaxK], ssum = 0.0f;
                if (score_mode == 0)
                    for (int p = 0; p < k; ++p) { w[p] = expf(pv[rr][p] - m); ssum += w[p]; }
                for (int p = 0; p < k; ++p) {
                    idx[tok[rr] * k + p] = pi[rr][p];
                    score[tok[rr] * k + p] = score_mode == 0 ? w[p] / ssum : expf(pv[rr][p] - m) * rz;
                    atomicAdd(hist_s + half * E_PAD + pi[rr][p], 1);
                }
            }
        }
    }
    if (want_psum) {   // column sums over the warp's 16 tokens: lanes with the same t, fixed xor order
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float p = pcol[j][c];
                p += __shfl_xor_sync(0xffffffffu, p, 4);
                p += __shfl_xor_sync(0xffffffffu, p, 8);
                p += __shfl_xor_sync(0xffffffffu, p, 16);
                if (g == 0) psum_s[warp * E_PAD + 8 * j + 2 * t + c] = p;
            }
        }
    }
    named_bar_sync(1, 256);   // the 8 MMA warps (the producer warp has left)
    for (int i = tid; i < 2 * E_PAD; i += 256) {
        const int hf = i / E_PAD, e = i - hf * E_PAD;
        const int tile = blockIdx.x * 2 + hf;
        if (e < E && tile < ntiles) {
            tile_hist[static_cast<size_t>(e) * ntiles + tile] = hist_s[i];
            if (want_psum) {
                const float* ps = psum_s + (hf * 4) * E_PAD + e;
                tile_psum[static_cast<size_t>(e) * ntiles + tile] = ((ps[0] + ps[E_PAD]) + ps[2 * E_PAD]) + ps[3 * E_PAD];
            }
        }
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

template <int NT>
static int launch_gate_mma_t(const CUtensorMap& tX, const CUtensorMap& tW, const __nv_bfloat16* x, const float* Wg, const float* bg,
                             const float* noise, const uint8_t* token_mask, const float* wnorm, int64_t T, int d, int E, int k,
                             int score_mode, int want_psum, float* logits, int* idx, float* score, int* tile_hist,
                             float* tile_psum, cudaStream_t st) {
    constexpr int E_PAD = NT * 8;
    constexpr int STAGE = kGmTok * 128 + 2 * E_PAD * 128;
    const int ntiles = static_cast<int>((T + MOE_TOKEN_TILE - 1) / MOE_TOKEN_TILE);
    const int nk = d / kGmKC;
    int nstages = (110 * 1024) / STAGE;   // two CTAs per SM: one's epilogue overlaps the other's loads
    if (nstages > 6) nstages = 6;
    if (nstages > nk) nstages = nk;
    if (nstages < 2) nstages = nk < 2 ? 1 : 2;
    const size_t smem = 1024 + static_cast<size_t>(nstages) * STAGE + 128 + (2 * E_PAD + 16 * E_PAD + 4) * 4;
    auto kfn = gate_fwd_mma_kernel<NT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error("gate_fwd_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int grid = (ntiles + 1) / 2;
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
    switch (ep) {
        case 16: return launch_gate_mma_t<2>(tX, tW, xb, Wg, bg, noise, token_mask, wnorm, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, tile_psum, st);
        case 32: return launch_gate_mma_t<4>(tX, tW, xb, Wg, bg, noise, token_mask, wnorm, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, tile_psum, st);
        default: return launch_gate_mma_t<8>(tX, tW, xb, Wg, bg, noise, token_mask, wnorm, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, tile_psum, st);
    }
}

}  // namespace moe
