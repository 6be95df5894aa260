// routing.cu — the HBM-bound half of the MoE layer, hand-written for sm_100a:
//   gate (projection + top-k + scores + per-tile histogram)      replaces NaiveGate's Linear/topk/softmax
//   route_scan (deterministic capacity prefix sums)              replaces fmoe_cuda.expert_count + cumsum + .item()
//   dispatch (token -> packed per-expert rows, bf16)             replaces fmoe_cuda.assign_pos + MOEScatter
//   combine (gate-weighted gather)                               replaces MOEGather + torch.bmm
//   and the matching backward kernels.
// Upstream semantics: SURVEY.md §3.3 / §8a (FastMoE is not vendored in /root/reference; call site
// /root/reference/models/resMoE.py:15-29, caller models/vision_transformer.py:319-322).
//
// Determinism contract: logits use LOGIT ORDER v1 (fixed fp32 FMA chains + xor
// butterfly) so they are bit-identical to oracle/gate_ref.c; top-k ties go to the lowest expert
// index; pairs are ranked inside an expert by ascending flattened index t*k+j.  No atomics on
// floating point anywhere; integer shared-memory atomics only where the result is order-free.
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace moe {

constexpr int kTokTile = MOE_TOKEN_TILE;  // tokens per routing tile (one CTA): small tiles = many CTAs in flight
constexpr int kTokPerWarp = kTokTile / 8;
static_assert(kTokTile == 64, "gate / dispatch kernels assume 8 warps x 8 tokens");
constexpr int kMaxK = 8;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_x4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load_x4(const __nv_bfloat16* p) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
// 8 consecutive elements as fp32
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ float warp_sum_xor(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Transposing butterfly: every lane holds 32 partial values v[0..31]; on return lane l holds in
// v[0] the sum over all 32 lanes of value l.  Each level adds exactly the pair (l, l^off) of the
// plain xor butterfly, so the result is bit-identical to LOGIT ORDER v1's reduction tree.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
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

// ------------------------------------------------------------------------------------------------
// K1: gate forward
// ------------------------------------------------------------------------------------------------
template <typename XT, int EG>
__global__ void __launch_bounds__(256)
gate_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ Wg, const float* __restrict__ bg,
                const float* __restrict__ noise, const uint8_t* __restrict__ token_mask, int64_t T, int d, int E, int k, int score_mode,
                int want_psum, float* __restrict__ logits, int* __restrict__ idx,
                float* __restrict__ score, int* __restrict__ tile_hist, float* __restrict__ tile_psum) {
    constexpr int NT = 64 / EG;  // tokens per warp pass: NT*EG = 64 accumulators per lane (NT <= kTokPerWarp)
    extern __shared__ float smem_f[];
    float* wg_s = smem_f;                          // [EG][d]
    float* lg_s = wg_s + EG * d;                   // [kTokTile][E+1]
    float* m_s = lg_s + kTokTile * (E + 1);        // [kTokTile] row max
    float* rz_s = m_s + kTokTile;                  // [kTokTile] Z
    int* hist_s = reinterpret_cast<int*>(rz_s + kTokTile);  // [E]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kTokTile;
    const int n_groups = (E + EG - 1) / EG;
    const int n_chunks = (d + 127) / 128;
    const int ldl = E + 1;

    for (int e = tid; e < E; e += 256) hist_s[e] = 0;

    for (int g = 0; g < n_groups; ++g) {
        __syncthreads();  // previous group's readers are done with wg_s
        for (int i = tid * 4; i < EG * d; i += 256 * 4) {
            const int e = g * EG + i / d;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < E) w = __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(g) * EG * d + i));
            *reinterpret_cast<float4*>(wg_s + i) = w;
        }
        __syncthreads();

        for (int it = 0; it < kTokPerWarp / NT; ++it) {
            const int64_t tok0 = t_base + warp * kTokPerWarp + it * NT;
            if (tok0 >= T) break;  // warp-uniform
            float acc[NT * EG];
#pragma unroll
            for (int i = 0; i < NT * EG; ++i) acc[i] = 0.0f;
            for (int c = 0; c < n_chunks; ++c) {
                const int i0 = c * 128 + lane * 4;
                if (i0 < d) {
                    float4 xv[NT];   // rows past T are clamped to the last row: their logits are never stored
#pragma unroll
                    for (int n = 0; n < NT; ++n) xv[n] = load_x4(x + min(tok0 + n, T - 1) * d + i0);
#pragma unroll
                    for (int e = 0; e < EG; ++e) {
                        const float4 w = *reinterpret_cast<const float4*>(wg_s + e * d + i0);
#pragma unroll
                        for (int n = 0; n < NT; ++n) {
                            float a = acc[n * EG + e];
                            a = fmaf(xv[n].x, w.x, a);
                            a = fmaf(xv[n].y, w.y, a);
                            a = fmaf(xv[n].z, w.z, a);
                            a = fmaf(xv[n].w, w.w, a);
                            acc[n * EG + e] = a;
                        }
                    }
                }
            }
#pragma unroll
            for (int rd = 0; rd < 2; ++rd) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = acc[rd * 32 + i];
                const float tot = transpose_reduce32(v, lane);
                const int j = rd * 32 + lane;
                const int n = j / EG, e = g * EG + (j % EG);
                const int64_t tok = tok0 + n;
                if (e < E && tok < T) {
                    float val = tot + (bg != nullptr ? __ldg(bg + e) : 0.0f);
                    if (noise != nullptr) val += __ldg(noise + tok * E + e);
                    lg_s[static_cast<int>(tok - t_base) * ldl + e] = val;
                    logits[tok * E + e] = val;
                }
            }
        }
    }
    __syncthreads();

    // ---- per-token top-k (ties -> lowest index), scores, histogram
    {
        const int64_t tok = t_base + tid;
        if (tid < kTokTile && tok < T && token_mask != nullptr && token_mask[tok] == 0) {
            // token-skip mask (reference models/resMoE.py:126-145): the token is not routed at all —
            // idx = -1, score = 0, no histogram entry, no share in psum (exp(l - inf) = 0)
            for (int j = 0; j < k; ++j) { idx[tok * k + j] = -1; score[tok * k + j] = 0.0f; }
            m_s[tid] = INFINITY;
            rz_s[tid] = 1.0f;
        } else if (tid < kTokTile && tok < T) {
            const float* lr = lg_s + tid * ldl;
            int picked[kMaxK];
            float pv[kMaxK];
            for (int j = 0; j < k; ++j) {
                int besti = -1;
                float best = 0.0f;
                for (int e = 0; e < E; ++e) {
                    bool used = false;
                    for (int qq = 0; qq < j; ++qq) used |= (picked[qq] == e);
                    if (used) continue;
                    const float v = lr[e];
                    if (besti < 0 || v > best) { besti = e; best = v; }
                }
                picked[j] = besti;
                pv[j] = best;
                idx[tok * k + j] = besti;
                atomicAdd(hist_s + besti, 1);
            }
            const float m = pv[0];
            float z = 0.0f;
            if (score_mode == 1 || want_psum) {
                for (int e = 0; e < E; ++e) z += expf(lr[e] - m);
                m_s[tid] = m;
                rz_s[tid] = z;
            }
            if (score_mode == 0) {
                float w[kMaxK], s = 0.0f;
                for (int j = 0; j < k; ++j) { w[j] = expf(pv[j] - m); s += w[j]; }
                for (int j = 0; j < k; ++j) score[tok * k + j] = w[j] / s;
            } else {
                for (int j = 0; j < k; ++j) score[tok * k + j] = expf(pv[j] - m) / z;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < E; e += 256) tile_hist[static_cast<size_t>(e) * gridDim.x + blockIdx.x] = hist_s[e];

    if (want_psum) {
        const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTokTile), T - t_base));
        for (int e = warp; e < E; e += 8) {
            float part = 0.0f;
            for (int t = lane; t < n_tok; t += 32) part += expf(lg_s[t * ldl + e] - m_s[t]) / rz_s[t];
            part = warp_sum_xor(part);
            if (lane == 0) tile_psum[static_cast<size_t>(e) * gridDim.x + blockIdx.x] = part;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: scan of the per-tile histograms -> tile bases, counts, capacity-clamped segments, GEMM tile table
// ------------------------------------------------------------------------------------------------
// K2a: one CTA per expert — exclusive prefix sum of its row of per-tile histograms (tile_base), its total (count) and the
// sum of its row of per-tile probability sums (psum), all in a fixed order.  (Round 1 ran the whole scan in one CTA: 72 us
// at E = 64, T = 262144 — one SM pulling every row through its L2 port.)
__global__ void __launch_bounds__(256)
route_scan_rows_kernel(const int* __restrict__ tile_hist, const float* __restrict__ tile_psum, int ntiles,
                       int* __restrict__ tile_base, int* __restrict__ count, float* __restrict__ psum) {
    __shared__ int wtot[8];
    __shared__ float wps[8];
    __shared__ int carry_s;
    const int e = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int* hrow = tile_hist + static_cast<size_t>(e) * ntiles;
    int* brow = tile_base + static_cast<size_t>(e) * ntiles;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int g0 = 0; g0 < ntiles; g0 += 256 * 8) {      // 8 consecutive tiles per thread
        int v[8], tot = 0;
        const int b0 = g0 + tid * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = b0 + i < ntiles ? hrow[b0 + i] : 0;
            tot += v[i];
        }
        int incl = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        int base = carry_s;
        for (int w = 0; w < warp; ++w) base += wtot[w];
        int run = base + incl - tot;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (b0 + i < ntiles) brow[b0 + i] = run;
            run += v[i];
        }
        __syncthreads();
        if (tid == 255) carry_s = run;                   // thread 255 holds the running total of the whole chunk
        __syncthreads();
    }
    if (tid == 0) count[e] = carry_s;
    if (tile_psum != nullptr && psum != nullptr) {       // fixed order: thread-strided partial sums, xor butterfly, warps in order
        const float* prow = tile_psum + static_cast<size_t>(e) * ntiles;
        float part = 0.0f;
#pragma unroll 4
        for (int b = tid; b < ntiles; b += 256) part += prow[b];
        part = warp_sum_xor(part);
        if (lane == 0) wps[warp] = part;
        __syncthreads();
        if (tid == 0) {
            float t = 0.0f;
            for (int w = 0; w < 8; ++w) t += wps[w];
            psum[e] = t;
        }
    }
}

// K2b: one CTA — capacity-clamped segments, GEMM tile table, load-balancing loss from the per-expert totals
__global__ void __launch_bounds__(1024)
route_scan_kernel(int E, long long capacity, int* __restrict__ count, int* __restrict__ kept,
                  int* __restrict__ seg_start, int* __restrict__ tile_expert, int* __restrict__ num_mtiles,
                  int max_mtiles, const float* __restrict__ psum, int aux_mode, long long tokens, int k, float* __restrict__ aux_loss,
                  float* __restrict__ aux_coef, long long slab_rows) {
    extern __shared__ int smem_i[];
    int* cnt_s = smem_i;          // [E]
    int* seg_s = smem_i + E;      // [E+1]
    float* ps_s = reinterpret_cast<float*>(smem_i + 2 * E + 1);  // [E]
    const int tid = threadIdx.x;
    for (int e = tid; e < E; e += 1024) {
        cnt_s[e] = count[e];
        ps_s[e] = (aux_mode != 0 && psum != nullptr) ? psum[e] : 0.0f;
    }
    __syncthreads();
    if (tid == 0) {
        int start = 0;
        for (int e = 0; e < E; ++e) {
            const int c = cnt_s[e];
            const int kp = static_cast<long long>(c) < capacity ? c : static_cast<int>(capacity);
            kept[e] = kp;
            seg_s[e] = start;
            seg_start[e] = start;
            // packed layout: the segment is as long as the expert needs; expert-parallel layout: fixed slabs, so that
            // the all-to-all that follows has static message sizes (no host round trip for the counts)
            start += slab_rows > 0 ? static_cast<int>(slab_rows) : (kp + MOE_ROW_ALIGN - 1) / MOE_ROW_ALIGN * MOE_ROW_ALIGN;
        }
        seg_s[E] = start;
        seg_start[E] = start;
        *num_mtiles = start / MOE_ROW_ALIGN;
        // load-balancing loss = sum_e coef_e * psum_e with coef_e = E/T * (share of expert e); the coefficients are
        // its gradient w.r.t. psum (the shares are integers, not differentiable).
        //   MOE_AUX_SWITCH: share = kept_e / sum(kept)        (Switch: E * sum_e f_e P_e)
        //   MOE_AUX_GSHARD: share = count_e / (tokens * k)    (GShard: mean(c_e m_e) * E^2)
        if (aux_mode != 0 && aux_loss != nullptr) {
            long long tot_kept = 0;
            for (int e = 0; e < E; ++e) tot_kept += cnt_s[e] < capacity ? cnt_s[e] : capacity;
            float loss = 0.0f;
            for (int e = 0; e < E; ++e) {
                const long long kp = cnt_s[e] < capacity ? cnt_s[e] : capacity;
                const float share = aux_mode == MOE_AUX_SWITCH
                                        ? static_cast<float>(kp) / static_cast<float>(tot_kept > 0 ? tot_kept : 1)
                                        : static_cast<float>(cnt_s[e]) / static_cast<float>(tokens * k);
                const float coef = share * static_cast<float>(E) / static_cast<float>(tokens);
                aux_coef[e] = coef;
                loss = fmaf(coef, ps_s[e], loss);
            }
            *aux_loss = loss;
        }
    }
    __syncthreads();
    const int nm = seg_s[E] / MOE_ROW_ALIGN;
    for (int m = tid; m < max_mtiles; m += 1024) {
        int e = -1;
        if (m < nm) {
            const int row = m * MOE_ROW_ALIGN;
            e = 0;
            while (e + 1 < E && seg_s[e + 1] <= row) ++e;
        }
        tile_expert[m] = e;
    }
}

// zero the pad rows [seg_start[e]+kept[e], seg_start[e+1]) of a packed bf16 buffer
__device__ __forceinline__ void zero_pad_rows(__nv_bfloat16* buf, int d, const int* seg_start, const int* kept, int e,
                                              int* row_src) {
    const int r0 = seg_start[e] + kept[e], r1 = seg_start[e + 1];
    const int per_row = d / 8;
    const int items = (r1 - r0) * per_row;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int r = r0 + it / per_row, c = (it % per_row) * 8;
        *reinterpret_cast<uint4*>(buf + static_cast<size_t>(r) * d + c) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (row_src != nullptr)
        for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) row_src[r] = -1;
}

// ------------------------------------------------------------------------------------------------
// K3: dispatch forward — positions + packed bf16 copy
// ------------------------------------------------------------------------------------------------
// `seg_start[e]` is the (global) row of expert e's first pair of THIS token set: the packed segment start on one GPU, or —
// under expert parallelism over peer memory — owner_rank * rows_per_rank + the row inside the owner's packed buffer
// where this source rank's share of expert e begins (moe_ep_exchange_counts).  Rows are written through `xrows`.
// Blocks >= ntiles zero the pad rows of the LOCAL buffer `xpad` (segments pad_seg / pad_kept).
template <typename XT>
__global__ void __launch_bounds__(256)
dispatch_fwd_kernel(const XT* __restrict__ x, const int* __restrict__ idx, const int* __restrict__ tile_base,
                    const int* __restrict__ seg_start, int64_t T, int d, int E, int k,
                    long long capacity, int ntiles, int* __restrict__ pos, int* __restrict__ row_src,
                    PeerRows xrows, __nv_bfloat16* __restrict__ xpad, const int* __restrict__ pad_seg,
                    const int* __restrict__ pad_kept) {
    if (static_cast<int>(blockIdx.x) >= ntiles) {
        zero_pad_rows(xpad, d, pad_seg, pad_kept, blockIdx.x - ntiles, row_src);
        return;
    }
    extern __shared__ int smem_i[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kTokTile;
    const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTokTile), T - t_base));
    const int n_ent = n_tok * k;
    const int n_chunks = (n_ent + 31) / 32;
    int* chunk_cnt = smem_i;                       // [n_chunks_max][E]
    int* ent_row = smem_i + (kTokTile * k / 32) * E;  // [256*k]
    const int64_t i_base = t_base * k;

    for (int i = tid; i < n_chunks * E; i += 256) chunk_cnt[i] = 0;
    __syncthreads();
    // A: per-32-entry chunk: rank inside the chunk + per-chunk expert counts
    for (int ch = warp; ch < n_chunks; ch += 8) {
        const int il = ch * 32 + lane;
        const int e = il < n_ent ? idx[i_base + il] : -1;   // -1: past the end, or a masked (skipped) token
        const bool valid = e >= 0;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(act, e);
            const int rk = __popc(peers & ((1u << lane) - 1u));
            if (rk == 0) chunk_cnt[ch * E + e] = __popc(peers);
            ent_row[il] = rk;  // temporarily: rank inside chunk
        }
    }
    __syncthreads();
    // B: exclusive scan over chunks per expert, seeded with this tile's global base
    for (int e = tid; e < E; e += 256) {
        int running = tile_base[static_cast<size_t>(e) * ntiles + blockIdx.x];
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int c = chunk_cnt[ch * E + e];
            chunk_cnt[ch * E + e] = running;
            running += c;
        }
    }
    __syncthreads();
    // C: final positions
    for (int il = tid; il < n_ent; il += 256) {
        const int e = idx[i_base + il];
        const long long rank = e >= 0 ? static_cast<long long>(chunk_cnt[(il >> 5) * E + e]) + ent_row[il] : capacity;
        const int row = rank < capacity ? seg_start[e] + static_cast<int>(rank) : -1;
        pos[i_base + il] = row;
        if (row >= 0 && row_src != nullptr) row_src[row] = static_cast<int>(i_base + il);
        ent_row[il] = row;
    }
    __syncthreads();
    // D: packed copy, 16 bytes of bf16 per work item, consecutive threads -> consecutive pieces of a row
    const int per_row = d / 8;
    const int items = n_ent * per_row;
    for (int it = tid; it < items; it += 256) {
        const int il = it / per_row, c = (it - il * per_row) * 8;
        const int row = ent_row[il];
        if (row < 0) continue;
        float v[8];
        load8(x + (t_base + il / k) * d + c, v);
        store8(peer_row<__nv_bfloat16>(xrows, row, d) + c, v);
    }
}

// ------------------------------------------------------------------------------------------------
// K4: combine forward   out[t] = sum_j score[t,j] * Y[pos[t,j]]
// ------------------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(PeerRows yrows, const int* __restrict__ pos, const float* __restrict__ score,
                   int64_t T, int d, int k, OT* __restrict__ out) {
    const int per_row = d / 8;
    const int64_t items = T * per_row;
    for (int64_t it = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; it < items;
         it += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t t = it / per_row;
        const int c = static_cast<int>(it - t * per_row) * 8;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        for (int j = 0; j < k; ++j) {
            const int row = __ldg(pos + t * k + j);
            if (row < 0) continue;
            const float s = __ldg(score + t * k + j);
            float y[8];
            load8(peer_row<const __nv_bfloat16>(yrows, row, d) + c, y);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(s, y[i], acc[i]);
        }
        store8(out + t * d + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// K5: combine backward   dYbuf[pos[t,j]] = bf16(score * dy[t]),  dscore[t,j] = <dy[t], Y[pos[t,j]]>
// ------------------------------------------------------------------------------------------------
template <typename GT>
__global__ void __launch_bounds__(256)
combine_bwd_kernel(const GT* __restrict__ dy, PeerRows yrows, const int* __restrict__ pos,
                   const float* __restrict__ score, const int* __restrict__ pad_seg, const int* __restrict__ pad_kept,
                   int64_t T, int d, int k, int n_tok_blocks, PeerRows dyrows, __nv_bfloat16* __restrict__ dypad,
                   float* __restrict__ dscore) {
    if (static_cast<int>(blockIdx.x) >= n_tok_blocks) {   // pad rows of the LOCAL buffer
        zero_pad_rows(dypad, d, pad_seg, pad_kept, blockIdx.x - n_tok_blocks, nullptr);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = static_cast<int64_t>(blockIdx.x) * 8 + warp;  // one warp per token
    if (t >= T) return;
    for (int j = 0; j < k; ++j) {
        const int row = __ldg(pos + t * k + j);
        float dot = 0.0f;
        if (row >= 0) {
            const float s = __ldg(score + t * k + j);
            const __nv_bfloat16* yr = peer_row<const __nv_bfloat16>(yrows, row, d);
            __nv_bfloat16* dyr = peer_row<__nv_bfloat16>(dyrows, row, d);
            for (int c = lane * 8; c < d; c += 256) {
                float g[8], y[8], o[8];
                load8(dy + t * d + c, g);
                load8(yr + c, y);
#pragma unroll
                for (int i = 0; i < 8; ++i) { dot = fmaf(g[i], y[i], dot); o[i] = s * g[i]; }
                store8(dyr + c, o);
            }
            dot = warp_sum_xor(dot);
        }
        if (lane == 0) dscore[t * k + j] = dot;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: gate backward   dlogits[T,E] from dscore (+ dpsum for the load-balancing loss)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ logits, const int* __restrict__ idx, const float* __restrict__ score,
                const float* __restrict__ dscore, const float* __restrict__ dpsum, int64_t T, int E, int k,
                int score_mode, float* __restrict__ dlogits) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (t >= T) return;
    const float* lr = logits + t * E;
    float* dl = dlogits + t * E;
    int pk[kMaxK];
    float s[kMaxK], g[kMaxK];
    for (int j = 0; j < k; ++j) { pk[j] = idx[t * k + j]; s[j] = score[t * k + j]; g[j] = dscore[t * k + j]; }
    const bool need_p = (score_mode == 1) || (dpsum != nullptr);
    float m = 0.0f, rz = 0.0f, pdot = 0.0f;
    const bool masked = pk[0] < 0;   // token-skip mask: no gate gradient at all
    if (need_p) {
        m = lr[max(pk[0], 0)];
        float z = 0.0f;
        for (int e = 0; e < E; ++e) z += expf(lr[e] - m);
        rz = 1.0f / z;
        if (dpsum != nullptr)
            for (int e = 0; e < E; ++e) pdot += expf(lr[e] - m) * rz * dpsum[e];
    }
    float inner = 0.0f;  // sum_j s_j g_j
    for (int j = 0; j < k; ++j) inner += s[j] * g[j];
    for (int e = 0; e < E; ++e) {
        float v = 0.0f;
        const float p = need_p ? expf(lr[e] - m) * rz : 0.0f;
        if (score_mode == 0) {
            for (int j = 0; j < k; ++j)
                if (pk[j] == e) v += s[j] * (g[j] - inner);
        } else {
            // score_j = p[idx_j]:  d/dl_e = sum_j g_j p_j (delta - p_e)
            for (int j = 0; j < k; ++j)
                if (pk[j] == e) v += g[j] * s[j];
            v -= inner * p;
        }
        if (dpsum != nullptr) v += p * (dpsum[e] - pdot);
        dl[e] = masked ? 0.0f : v;
    }
}

// ------------------------------------------------------------------------------------------------
// K7: dispatch backward   dx[t] = sum_j dXbuf[pos[t,j]] + sum_e dlogits[t,e] * Wg[e]
// ------------------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256)
dispatch_bwd_kernel(const __nv_bfloat16* __restrict__ dxbuf, const int* __restrict__ pos,
                    const float* __restrict__ dlogits, const int* __restrict__ idx, const float* __restrict__ Wg,
                    int64_t T, int d, int E, int k, int dense_dlogits, OT* __restrict__ dx) {
    const int per_row = d / 8;
    const int64_t items = T * per_row;
    for (int64_t it = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; it < items;
         it += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t t = it / per_row;
        const int c = static_cast<int>(it - t * per_row) * 8;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        if (dxbuf != nullptr) {
            for (int j = 0; j < k; ++j) {
                const int row = __ldg(pos + t * k + j);
                if (row < 0) continue;
                float v[8];
                load8(dxbuf + static_cast<size_t>(row) * d + c, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += v[i];
            }
        }
        if (dlogits != nullptr) {
            if (dense_dlogits) {
                for (int e = 0; e < E; ++e) {
                    const float g = __ldg(dlogits + t * E + e);
                    float w[8];
                    load8(Wg + static_cast<size_t>(e) * d + c, w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(g, w[i], acc[i]);
                }
            } else {
                for (int j = 0; j < k; ++j) {
                    const int e = __ldg(idx + t * k + j);
                    if (e < 0) continue;   // masked (skipped) token
                    const float g = __ldg(dlogits + t * E + e);
                    float w[8];
                    load8(Wg + static_cast<size_t>(e) * d + c, w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(g, w[i], acc[i]);
                }
            }
        }
        store8(dx + t * d + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// K7b: gate backward + dispatch backward in one pass (the layer's fused path)
//   dlogits[t,:] from dscore (+ dpsum)  -> written out for the gate weight gradient
//   dx[t] = sum_j dXbuf[pos[t,j]] + sum_e dlogits[t,e] * Wg[e]
// One CTA per 64-token tile: 64 threads compute the tile's dlogits rows into shared memory, then
// thread (cg, tg) owns 4 consecutive features for every TG-th token.  With dense dlogits and
// E <= 16 the thread keeps its Wg[:, cg*4..+3] slice in registers for the whole tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}

// Feature phase of the staged path for one (column group, token group) thread: n_it tokens, TG apart.
//   rp: this thread's 4 features of the first token's staged rows; dr: its dlogits row; op: its output
//   KT: compile-time k (0 = runtime k);  REGS: dense dlogits and E <= 16 -> the Wg slice lives in registers
//   idx_t: sparse dlogits only (NaiveGate without aux loss): the token's selected experts
template <typename OT, int KT, bool REGS>
__device__ __forceinline__ void gdb_tokens(const __nv_bfloat16* rp, const float* dr, const float* __restrict__ Wg,
                                           const int* idx_t, OT* op, int n_it, int TG, int d, int E, int k, int cg) {
    float4 wg[16];
    if constexpr (REGS) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
            wg[e] = e < E ? __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e) * d + cg * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int kk = KT > 0 ? KT : k;
    const size_t r_step = static_cast<size_t>(TG) * kk * d, o_step = static_cast<size_t>(TG) * d;
    const int ne4 = (E + 3) >> 2;
    if constexpr (!REGS) {
        if (idx_t == nullptr) {
            // dense dlogits, E > 16: four tokens per pass over the experts, 16 experts of Wg in registers at a time
            // (256 FMAs per 16 L1-resident loads instead of one load per expert and token)
            constexpr int TB = 4;
            for (int it0 = 0; it0 < n_it; it0 += TB) {
                float4 acc[TB];
#pragma unroll
                for (int u = 0; u < TB; ++u) {
                    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const __nv_bfloat16* ru = rp + static_cast<size_t>(min(it0 + u, n_it - 1)) * r_step;
                    for (int j = 0; j < kk; ++j) {
                        const uint2 r = *reinterpret_cast<const uint2*>(ru + j * d);
                        acc[u].x += __uint_as_float(r.x << 16); acc[u].y += __uint_as_float(r.x & 0xffff0000u);
                        acc[u].z += __uint_as_float(r.y << 16); acc[u].w += __uint_as_float(r.y & 0xffff0000u);
                    }
                }
                for (int e0 = 0; e0 < E; e0 += 16) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        wg[e] = e0 + e < E ? __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e0 + e) * d + cg * 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < TB; ++u) {
                        const float* du = dr + static_cast<size_t>(min(it0 + u, n_it - 1)) * TG * E + e0;
#pragma unroll
                        for (int e4 = 0; e4 < 4; ++e4) {
                            const float4 g4 = *reinterpret_cast<const float4*>(du + (e0 + e4 * 4 < E ? e4 * 4 : 0));   // past E: zero weights
                            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 w = wg[e4 * 4 + i];
                                acc[u].x = fmaf(gg[i], w.x, acc[u].x); acc[u].y = fmaf(gg[i], w.y, acc[u].y);
                                acc[u].z = fmaf(gg[i], w.z, acc[u].z); acc[u].w = fmaf(gg[i], w.w, acc[u].w);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < TB; ++u)
                    if (it0 + u < n_it) st4(op + static_cast<size_t>(it0 + u) * o_step, acc[u]);
            }
            return;
        }
    }
#pragma unroll 2
    for (int it = 0; it < n_it; ++it, rp += r_step, dr += TG * E, op += o_step) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < (KT > 0 ? KT : 1); ++j) {
            for (int jj = j; jj < kk; jj += (KT > 0 ? kk : 1)) {   // KT > 0: exactly one pass per j
                const uint2 r = *reinterpret_cast<const uint2*>(rp + jj * d);
                acc.x += __uint_as_float(r.x << 16); acc.y += __uint_as_float(r.x & 0xffff0000u);
                acc.z += __uint_as_float(r.y << 16); acc.w += __uint_as_float(r.y & 0xffff0000u);
            }
        }
        if constexpr (REGS) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                // experts past E carry zero weights: re-reading an in-range (finite) quad keeps the products at zero
                const float4 g4 = *reinterpret_cast<const float4*>(dr + (e4 < ne4 ? e4 * 4 : 0));
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = wg[e4 * 4 + i];
                    acc.x = fmaf(gg[i], w.x, acc.x); acc.y = fmaf(gg[i], w.y, acc.y);
                    acc.z = fmaf(gg[i], w.z, acc.z); acc.w = fmaf(gg[i], w.w, acc.w);
                }
            }
        } else {   // sparse dlogits (NaiveGate without an aux loss): only the selected experts
            const int* ip = idx_t + static_cast<size_t>(it) * TG * k;
            for (int j = 0; j < k; ++j) {
                const int e = __ldg(ip + j);
                if (e < 0) continue;
                const float g = dr[e];
                const float4 w = __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e) * d + cg * 4));
                acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y);
                acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
            }
        }
        st4(op, acc);
    }
}

template <typename OT>
__global__ void __launch_bounds__(256, 2)
gate_dispatch_bwd_kernel(const __nv_bfloat16* __restrict__ dxbuf, const int* __restrict__ pos,
                         const float* __restrict__ logits, const int* __restrict__ idx, const float* __restrict__ score,
                         const float* __restrict__ dscore, const float* __restrict__ dpsum, const float* __restrict__ Wg,
                         int64_t T, int d, int E, int k, int score_mode, float* __restrict__ dlogits, OT* __restrict__ dx,
                         int ts /* tokens per staged sub-tile; 0 = gather through registers */) {
    extern __shared__ float smem_f[];
    float* dl_s = smem_f;                                        // [kTokTile][E]
    int* pos_s = reinterpret_cast<int*>(smem_f + kTokTile * E);  // [kTokTile][k]
    __nv_bfloat16* rows_s = reinterpret_cast<__nv_bfloat16*>(smem_f + kTokTile * (E + k));   // [ts][k][d] staged dXbuf rows
    const int tid = threadIdx.x;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kTokTile;
    const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTokTile), T - t_base));
    const bool dense = (score_mode == 1) || (dpsum != nullptr);
    // rows of this tile's pairs, so that the gathers below do not wait on a dependent index load
    for (int i = tid; i < n_tok * k; i += 256) pos_s[i] = pos[t_base * k + i];
    // Staged gather: every packed row the sub-tile needs is requested with 16-byte cp.async up front (48 KB in flight
    // per CTA at d = 384, k = 1), so the gather runs at memory-level parallelism instead of one dependent 8-byte load
    // per thread and token (measured round 1e: 84 us = 0.98 TB/s with the register-pipelined gather).
    const bool staged = ts > 0 && dxbuf != nullptr;
    auto stage = [&](int tl_begin) {
        const int c16 = d >> 3;   // 16-byte chunks per row
        const int n16 = min(ts, n_tok - tl_begin) * k * c16;
        const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(rows_s));
        for (int i = tid; i < n16; i += 256) {
            const int pr = i / c16, c = i - pr * c16;
            const int row = pos_s[tl_begin * k + pr];
            if (row >= 0) {
                const __nv_bfloat16* src = dxbuf + static_cast<size_t>(row) * d + c * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + i * 16), "l"(src) : "memory");
            } else {
                reinterpret_cast<uint4*>(rows_s)[i] = make_uint4(0u, 0u, 0u, 0u);   // dropped / skipped pair
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (staged) {
        __syncthreads();
        stage(0);
    }

    {   // dlogits rows of the tile: 4 threads per token, experts strided over the 4 (same formulas as gate_bwd_kernel)
        const int tl = tid >> 2, part = tid & 3;
        const bool live = tl < n_tok;
        const int64_t t = t_base + (live ? tl : 0);
        const float* lr = logits + t * E;
        int pk[kMaxK];
        float sc[kMaxK], g[kMaxK];
        for (int j = 0; j < k; ++j) { pk[j] = idx[t * k + j]; sc[j] = score[t * k + j]; g[j] = dscore[t * k + j]; }
        float m = 0.0f, rz = 0.0f, pdot = 0.0f;
        const bool masked = pk[0] < 0;   // token-skip mask: no gate gradient at all
        if (dense) {
            m = lr[max(pk[0], 0)];
            float z = 0.0f;
#pragma unroll 1
            for (int e = part; e < E; e += 4) z += expf(lr[e] - m);
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            rz = 1.0f / z;
            if (dpsum != nullptr) {
#pragma unroll 1
                for (int e = part; e < E; e += 4) pdot += expf(lr[e] - m) * rz * dpsum[e];
                pdot += __shfl_xor_sync(0xffffffffu, pdot, 1);
                pdot += __shfl_xor_sync(0xffffffffu, pdot, 2);
            }
        }
        float inner = 0.0f;
        for (int j = 0; j < k; ++j) inner += sc[j] * g[j];
        if (live) {
            float* dl = dl_s + tl * E;
#pragma unroll 1
            for (int e = part; e < E; e += 4) {
                float v = 0.0f;
                const float p = dense ? expf(lr[e] - m) * rz : 0.0f;
                if (score_mode == 0) {
                    for (int j = 0; j < k; ++j)
                        if (pk[j] == e) v += sc[j] * (g[j] - inner);
                } else {
                    for (int j = 0; j < k; ++j)
                        if (pk[j] == e) v += g[j] * sc[j];
                    v -= inner * p;
                }
                if (dpsum != nullptr) v += p * (dpsum[e] - pdot);
                dl[e] = masked ? 0.0f : v;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n_tok * E; i += 256) dlogits[t_base * E + i] = dl_s[i];

    const int CG = d / 4;
    const int TG = CG >= 256 ? 1 : 256 / CG;
    if (staged) {
        const bool regs = dense && E <= 16;
        for (int sub = 0; sub < n_tok; sub += ts) {
            if (sub > 0) {
                __syncthreads();   // the previous sub-tile's rows have been consumed
                stage(sub);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const int sub_end = min(sub + ts, n_tok);
            for (int cgb = 0; cgb < CG; cgb += 256) {
                const int cg = CG >= 256 ? cgb + tid : tid % CG;
                const int tg = CG >= 256 ? 0 : tid / CG;
                if (cg >= CG || tg >= TG) continue;
                const __nv_bfloat16* rp = rows_s + static_cast<size_t>(tg) * k * d + cg * 4;
                const float* dr = dl_s + (sub + tg) * E;
                OT* op = dx + (t_base + sub + tg) * d + cg * 4;
                const int n_it = (sub_end - sub - tg + TG - 1) / TG;
                if (regs && k == 1) gdb_tokens<OT, 1, true>(rp, dr, Wg, nullptr, op, n_it, TG, d, E, k, cg);
                else if (regs && k == 2) gdb_tokens<OT, 2, true>(rp, dr, Wg, nullptr, op, n_it, TG, d, E, k, cg);
                else gdb_tokens<OT, 0, false>(rp, dr, Wg, dense ? nullptr : idx + (t_base + sub + tg) * k, op, n_it, TG, d, E, k, cg);
            }
        }
        return;
    }
    for (int cgb = 0; cgb < CG; cgb += 256) {
        const int cg = CG >= 256 ? cgb + tid : tid % CG;
        const int tg = CG >= 256 ? 0 : tid / CG;
        if (cg >= CG || tg >= TG) continue;
        const bool regs = dense && E <= 16;
        float4 wg[16];
        if (regs) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                wg[e] = e < E ? __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e) * d + cg * 4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        constexpr int U = 2;    // tokens per batch; the next batch's gathers are issued before this batch's FMAs
        constexpr int KP = 2;   // slots whose gathers are pipelined (k = 1, 2 cover every shipped gate)
        float4 nxt[U][KP];
        float nkeep[U][KP];
        auto gather = [&](int tl0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tl = min(tl0 + u * TG, n_tok - 1);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    const int row = (j < k && dxbuf != nullptr) ? pos_s[tl * k + j] : -1;
                    nkeep[u][j] = (row >= 0 && tl0 + u * TG < n_tok) ? 1.0f : 0.0f;
                    nxt[u][j] = (j < k && dxbuf != nullptr) ? load_x4(dxbuf + static_cast<size_t>(max(row, 0)) * d + cg * 4)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        gather(tg);
        for (int tl0 = tg; tl0 < n_tok; tl0 += U * TG) {
            float4 acc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    acc[u].x = fmaf(nkeep[u][j], nxt[u][j].x, acc[u].x); acc[u].y = fmaf(nkeep[u][j], nxt[u][j].y, acc[u].y);
                    acc[u].z = fmaf(nkeep[u][j], nxt[u][j].z, acc[u].z); acc[u].w = fmaf(nkeep[u][j], nxt[u][j].w, acc[u].w);
                }
            }
            if (tl0 + U * TG < n_tok) gather(tl0 + U * TG);
            if (dxbuf != nullptr) {
                for (int j = KP; j < k; ++j) {   // un-pipelined tail for k > 2
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int tl = tl0 + u * TG;
                        const int row = tl < n_tok ? pos_s[tl * k + j] : -1;
                        if (row >= 0) {
                            const float4 v = load_x4(dxbuf + static_cast<size_t>(row) * d + cg * 4);
                            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
                        }
                    }
                }
            }
            if (dense && !regs) {
                for (int e0 = 0; e0 < E; e0 += 16) {   // Wg slice of 16 experts: L1/L2-resident, amortised over the batch
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        wg[e] = e0 + e < E ? __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e0 + e) * d + cg * 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float* dr = dl_s + min(tl0 + u * TG, n_tok - 1) * E + e0;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float g = e0 + e < E ? dr[e] : 0.0f;
                            acc[u].x = fmaf(g, wg[e].x, acc[u].x); acc[u].y = fmaf(g, wg[e].y, acc[u].y);
                            acc[u].z = fmaf(g, wg[e].z, acc[u].z); acc[u].w = fmaf(g, wg[e].w, acc[u].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tl = tl0 + u * TG;
                if (tl >= n_tok) break;
                const int64_t t = t_base + tl;
                const float* dr = dl_s + tl * E;
                if (regs) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float g = e < E ? dr[e] : 0.0f;
                        acc[u].x = fmaf(g, wg[e].x, acc[u].x); acc[u].y = fmaf(g, wg[e].y, acc[u].y);
                        acc[u].z = fmaf(g, wg[e].z, acc[u].z); acc[u].w = fmaf(g, wg[e].w, acc[u].w);
                    }
                } else if (!dense) {
                    for (int j = 0; j < k; ++j) {
                        const int e = __ldg(idx + t * k + j);
                        if (e < 0) continue;
                        const float g = dr[e];
                        const float4 w = __ldg(reinterpret_cast<const float4*>(Wg + static_cast<size_t>(e) * d + cg * 4));
                        acc[u].x = fmaf(g, w.x, acc[u].x); acc[u].y = fmaf(g, w.y, acc[u].y);
                        acc[u].z = fmaf(g, w.z, acc[u].z); acc[u].w = fmaf(g, w.w, acc[u].w);
                    }
                }
                st4(dx + t * d + cg * 4, acc[u]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K8: gate weight gradient   dWg = dlogits^T x, dbg = colsum(dlogits)   (two deterministic stages)
// Stage 1 is persistent: block b owns the 64-token tiles b, b + grid, ... (a fixed assignment, so the
// summation order never changes) and keeps its [16 experts x 4 columns] accumulators in registers
// across all of them.  Thread (cg, tg): column group cg = 4 consecutive features, token group tg takes
// every TG-th token of a tile; the TG partial sums are combined in shared memory in tg order.
// Stage 2 sums the per-block partials in block order.
// ------------------------------------------------------------------------------------------------
constexpr int kWgTile = 64;   // tokens per step of the persistent loop
constexpr int kWgEG = 16;     // experts per register pass

__device__ __forceinline__ float4 ldx4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldx4(const __nv_bfloat16* p) { return load_x4(p); }

template <typename XT>
__global__ void __launch_bounds__(256)
gate_wgrad_partial_kernel(const float* __restrict__ dlogits, const XT* __restrict__ x, int64_t T, int d, int E,
                          int ntiles, float* __restrict__ part_w, float* __restrict__ part_b) {
    extern __shared__ float smem_f[];
    const int tid = threadIdx.x;
    const int CG = d / 4;                           // column groups of x
    const int TG = CG + 1 > 256 ? 1 : 256 / (CG + 1);  // token groups sharing a tile
    float* dl_s = smem_f;                           // [kWgTile][E]
    float* red_s = smem_f + kWgTile * E;            // [TG][kWgEG][d + 4] (only used when TG > 1)
    float* pw = part_w + static_cast<size_t>(blockIdx.x) * E * d;

    // dbg = colsum(dlogits) rides along as one virtual column group (cg == CG) whose x is (1, 0, 0, 0)
    const int CGX = CG + 1;
    for (int g0 = 0; g0 < E; g0 += kWgEG) {
        for (int cgb = 0; cgb < CGX; cgb += 256) {   // more than one sweep only when d >= 1024
            const int cg = CGX > 256 ? cgb + tid : tid % CGX;
            const int tg = CGX > 256 ? 0 : tid / CGX;
            const bool active = cg < CGX && tg < TG;
            const bool is_bias = cg == CG;
            float4 acc[kWgEG];
#pragma unroll
            for (int i = 0; i < kWgEG; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int64_t t_base = static_cast<int64_t>(tile) * kWgTile;
                const int n_tok = static_cast<int>(min(static_cast<int64_t>(kWgTile), T - t_base));
                __syncthreads();   // previous tile's readers are done with dl_s
                for (int i = tid; i < n_tok * E; i += 256) dl_s[i] = dlogits[t_base * E + i];
                __syncthreads();
                if (active) {
                    constexpr int U = 4;   // tokens per batch; the next batch is loaded before this one is consumed
                    float4 xn[U];
                    auto fetch = [&](int t0) {
#pragma unroll
                        for (int u = 0; u < U; ++u)   // always load (clamped); slots past the tile get weight 0 below
                            xn[u] = is_bias ? make_float4(1.f, 0.f, 0.f, 0.f)
                                            : ldx4(x + (t_base + min(t0 + u * TG, n_tok - 1)) * d + cg * 4);
                    };
                    fetch(tg);
                    for (int t0 = tg; t0 < n_tok; t0 += U * TG) {
                        float4 xv[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) xv[u] = xn[u];
                        if (t0 + U * TG < n_tok) fetch(t0 + U * TG);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const bool live = t0 + u * TG < n_tok;
                            const float* dr = dl_s + min(t0 + u * TG, n_tok - 1) * E + g0;
#pragma unroll
                            for (int i = 0; i < kWgEG; ++i) {
                                const float g = (live && g0 + i < E) ? dr[i] : 0.0f;
                                acc[i].x = fmaf(g, xv[u].x, acc[i].x);
                                acc[i].y = fmaf(g, xv[u].y, acc[i].y);
                                acc[i].z = fmaf(g, xv[u].z, acc[i].z);
                                acc[i].w = fmaf(g, xv[u].w, acc[i].w);
                            }
                        }
                    }
                }
            }
            float* pb = part_b + static_cast<size_t>(blockIdx.x) * E;
            if (TG == 1) {
                if (active) {
#pragma unroll
                    for (int i = 0; i < kWgEG; ++i) {
                        if (g0 + i >= E) break;
                        if (is_bias) pb[g0 + i] = acc[i].x;
                        else *reinterpret_cast<float4*>(pw + static_cast<size_t>(g0 + i) * d + cg * 4) = acc[i];
                    }
                }
            } else {
                const int dX = d + 4;   // row of the reduction buffer: d features + the bias group
                __syncthreads();
                if (active) {
#pragma unroll
                    for (int i = 0; i < kWgEG; ++i)
                        *reinterpret_cast<float4*>(red_s + (static_cast<size_t>(tg) * kWgEG + i) * dX + cg * 4) = acc[i];
                }
                __syncthreads();
                for (int o = tid; o < kWgEG * dX; o += 256) {
                    const int i = o / dX, c = o - i * dX;
                    if (g0 + i >= E) break;
                    if (c > d) continue;
                    float v = 0.0f;
                    for (int q = 0; q < TG; ++q) v += red_s[static_cast<size_t>(q) * kWgEG * dX + o];   // tg order
                    if (c < d) pw[static_cast<size_t>(g0 + i) * d + c] = v;
                    else pb[g0 + i] = v;
                }
            }
        }
    }
}

// Stage 1 for bf16 activations, on the tensor cores: per 16-expert group dWg^T-partial[16, d] += dlogits^T[16, 8] x[8, d]
// as m16n8k8 tf32 MMAs (x is exact in tf32, dlogits is rounded to tf32: 2^-11 relative).  T d E fp32 FMAs on the CUDA
// cores cost ~50 us at the config-2 shape; here the arithmetic disappears and the kernel streams x once per expert
// group (cp.async, double-buffered 32-token sub-tiles).  Same block -> tile assignment and partial layout as the SIMT
// kernel, so the result stays reproducible bit for bit.
constexpr int kWgSub = 32;     // tokens per staged sub-tile

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT: 16-expert m-tiles held in registers (ALL experts in one pass over x: E <= 16 MT).  gridDim.y feature slices: slice s
// owns features [s d / S, (s + 1) d / S) — it stages only that part of every x row — so that MT x (n-tiles per warp) stays
// within 16 accumulator tiles whatever E and d are (round 1 re-read x once per 16 experts: 4 passes at E = 64).
template <int MT>
__global__ void __launch_bounds__(256)
gate_wgrad_partial_mma_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ x, int64_t T, int d, int E,
                              int ntiles, float* __restrict__ part_w, float* __restrict__ part_b) {
    constexpr int EG = 16 * MT;            // experts per pass
    constexpr int LDL = EG + 8;            // floats per staged dlogits row (+8: conflict-free A-fragment reads)
    constexpr int MAXT = 16;               // accumulator tiles per thread (MT x n-tiles per warp)
    extern __shared__ __align__(16) uint8_t smem_wg[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ds = d / gridDim.y;                            // features of this slice
    const int f0 = blockIdx.y * ds;
    const int ldx = ds + 8;                                  // staged x row, in bf16 (16-byte multiple, bank-skewed)
    const size_t xbytes = static_cast<size_t>(kWgSub) * ldx * 2;
    auto xs = [&](int buf) { return reinterpret_cast<uint16_t*>(smem_wg + buf * xbytes); };
    auto dls = [&](int buf) { return reinterpret_cast<float*>(smem_wg + 2 * xbytes) + buf * (kWgSub * LDL); };
    const int fw = ds / 8;                                   // features per warp
    const int nt = fw / 8;                                   // n-tiles per warp (MT * nt <= MAXT, checked at launch)
    const int my_tiles = static_cast<int>(blockIdx.x) < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int steps = my_tiles * (kWgTile / kWgSub);
    float* pw = part_w + static_cast<size_t>(blockIdx.x) * E * d;
    const int chunks_per_row = ds / 8;                       // 16-byte pieces of one staged x row

    for (int g0 = 0; g0 < E; g0 += EG) {                     // one pass unless E > 16 MT
        float acc[MAXT][4];
#pragma unroll
        for (int j = 0; j < MAXT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
        float bsum = 0.0f;
        auto tok_base = [&](int st) {   // first token of step st
            const int tile = blockIdx.x + (st / (kWgTile / kWgSub)) * gridDim.x;
            return static_cast<int64_t>(tile) * kWgTile + (st % (kWgTile / kWgSub)) * kWgSub;
        };
        auto stage_x = [&](int st, int buf) {   // cp.async: rows past T are clamped (their dlogits are staged as 0)
            const int64_t t0 = tok_base(st);
            for (int c = tid; c < kWgSub * chunks_per_row; c += 256) {
                const int r = c / chunks_per_row, cc = c - r * chunks_per_row;
                const __nv_bfloat16* src = x + min(t0 + r, T - 1) * d + f0 + cc * 8;
                const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(xs(buf) + static_cast<size_t>(r) * ldx + cc * 8));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float dreg[2 * MT];
        auto fetch_dl = [&](int st) {   // 32 tokens x EG experts, 2 MT values per thread
            const int64_t t0 = tok_base(st);
#pragma unroll
            for (int i = 0; i < 2 * MT; ++i) {
                const int ix = tid + i * 256, r = ix / EG, e = g0 + (ix % EG);
                dreg[i] = (t0 + r < T && e < E) ? __ldg(dlogits + (t0 + r) * E + e) : 0.0f;
            }
        };
        auto store_dl = [&](int buf) {
#pragma unroll
            for (int i = 0; i < 2 * MT; ++i) {
                const int ix = tid + i * 256;
                dls(buf)[(ix / EG) * LDL + (ix % EG)] = dreg[i];
            }
        };
        if (steps > 0) {
            stage_x(0, 0);
            fetch_dl(0);
            store_dl(0);
        }
        for (int st = 0; st < steps; ++st) {
            const int buf = st & 1;
            if (st + 1 < steps) {
                stage_x(st + 1, buf ^ 1);   // that buffer's readers finished at the barrier that closed step st - 1
                fetch_dl(st + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                // x and dlogits of step st are visible to every warp
            const uint16_t* xb = xs(buf);
            const float* db = dls(buf);
#pragma unroll
            for (int ks = 0; ks < kWgSub / 8; ++ks) {
                const int k0 = ks * 8;
                uint32_t a[MT][4];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    a[m][0] = to_tf32(db[(k0 + t) * LDL + 16 * m + g]);
                    a[m][1] = to_tf32(db[(k0 + t) * LDL + 16 * m + g + 8]);
                    a[m][2] = to_tf32(db[(k0 + t + 4) * LDL + 16 * m + g]);
                    a[m][3] = to_tf32(db[(k0 + t + 4) * LDL + 16 * m + g + 8]);
                }
                const uint16_t* r0 = xb + static_cast<size_t>(k0 + t) * ldx + warp * fw + g;
                const uint16_t* r1 = r0 + 4 * static_cast<size_t>(ldx);
#pragma unroll
                for (int j = 0; j < MAXT / MT; ++j) {
                    if (j < nt) {
                        const uint32_t b0 = static_cast<uint32_t>(r0[j * 8]) << 16, b1 = static_cast<uint32_t>(r1[j * 8]) << 16;
#pragma unroll
                        for (int m = 0; m < MT; ++m) mma_tf32_16x8x8(acc[j * MT + m], a[m], b0, b1);
                    }
                }
            }
            if (blockIdx.y == 0 && tid < EG) {
#pragma unroll 8
                for (int r = 0; r < kWgSub; ++r) bsum += db[r * LDL + tid];
            }
            if (st + 1 < steps) store_dl(buf ^ 1);
            __syncthreads();                // readers of buffer `buf` are done before step st + 2 overwrites it
        }
        // C fragment: rows g, g + 8 = experts, columns 2t, 2t + 1 of each n-tile = features
#pragma unroll
        for (int j = 0; j < MAXT / MT; ++j) {
            if (j < nt) {
                const int col = f0 + warp * fw + j * 8 + 2 * t;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int e = g0 + 16 * m + g;
                    if (e < E) *reinterpret_cast<float2*>(pw + static_cast<size_t>(e) * d + col) = make_float2(acc[j * MT + m][0], acc[j * MT + m][1]);
                    if (e + 8 < E) *reinterpret_cast<float2*>(pw + static_cast<size_t>(e + 8) * d + col) = make_float2(acc[j * MT + m][2], acc[j * MT + m][3]);
                }
            }
        }
        if (blockIdx.y == 0 && tid < EG && g0 + tid < E) part_b[static_cast<size_t>(blockIdx.x) * E + g0 + tid] = bsum;
        __syncthreads();
    }
}

// out[o] = sum_b part[b][o].  CTA = 64 outputs x 16 slices of the partials (slice s takes b = s, s + 16, ...; four
// independent chains per thread), consecutive threads read consecutive outputs of one partial (coalesced: a warp per
// output with lanes striding the partials read 4 bytes per 32-byte sector and took 12.4 us for the 296 partials of config
// 2), slices combined in a fixed order: reproducible bit for bit.
constexpr int kWgRedSlices = 16;
__global__ void __launch_bounds__(64 * kWgRedSlices)
gate_wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nparts, int d, int E,
                         float* __restrict__ dWg, float* __restrict__ dbg) {
    __shared__ float red[kWgRedSlices][64];
    const int n = E * d, tot = n + (dbg != nullptr ? E : 0);
    const int cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
    const int o = blockIdx.x * 64 + cl;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (o < tot) {
        const float* src = o < n ? part_w + o : part_b + (o - n);
        const size_t stride = o < n ? static_cast<size_t>(n) : static_cast<size_t>(E);
        int b = slice;
        for (; b + 3 * kWgRedSlices < nparts; b += 4 * kWgRedSlices) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += src[static_cast<size_t>(b + i * kWgRedSlices) * stride];
        }
        for (int i = 0; b < nparts; b += kWgRedSlices, ++i) acc[i] += src[static_cast<size_t>(b) * stride];
    }
    red[slice][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (slice == 0 && o < tot) {
        float v = 0.0f;
#pragma unroll
        for (int sl = 0; sl < kWgRedSlices; sl += 4) v += (red[sl][cl] + red[sl + 1][cl]) + (red[sl + 2][cl] + red[sl + 3][cl]);
        if (o < n) dWg[o] = v;
        else dbg[o - n] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// misc: fp32 -> bf16 cast (weights), per-segment column sums (bias gradients)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n8) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n8;
         i += static_cast<int64_t>(gridDim.x) * 256) {
        float v[8];
        load8(src + i * 8, v);
        store8(dst + i * 8, v);
    }
}

// Both expert weight matrices in ONE launch (items [0, n8_0) belong to the first, the rest to the second), four independent
// 32-byte loads in flight per thread: the per-step weight cast of a layer.
__global__ void __launch_bounds__(256)
cast_bf16_pair_kernel(const float* __restrict__ src0, __nv_bfloat16* __restrict__ dst0, int64_t n8_0,
                      const float* __restrict__ src1, __nv_bfloat16* __restrict__ dst1, int64_t n8_1) {
    const int64_t total = n8_0 + n8_1, stride = static_cast<int64_t>(gridDim.x) * 256;
    for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i0 < total; i0 += 4 * stride) {
        float v[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j * stride;
            if (i < total) load8(i < n8_0 ? src0 + i * 8 : src1 + (i - n8_0) * 8, v[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j * stride;
            if (i < total) store8(i < n8_0 ? dst0 + i * 8 : dst1 + (i - n8_0) * 8, v[j]);
        }
    }
}

// Per-segment column sums (bias gradients) in two deterministic stages.
// Stage 1: CTA = 128 rows x 256 columns; warp w reads rows w*16..w*16+15 with all 16 16-byte loads in
// flight, the 8 warps are combined in shared memory in warp order -> part[rb][c].  128-row blocks never
// straddle experts (segments are 256-aligned); blocks beyond the last live row exit at once.
// Stage 2: out[e][c] = sum of the expert's row blocks in order.
__global__ void __launch_bounds__(256)
segment_colsum_partial_kernel(const __nv_bfloat16* __restrict__ buf, const int* __restrict__ seg_start, int E, int cols,
                              float* __restrict__ part) {
    __shared__ float red[8][256];
    const int rb = blockIdx.y, row0 = rb * 128;
    if (row0 >= __ldg(seg_start + E)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 256 + lane * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (c < cols) {
        uint4 v[16];
        const __nv_bfloat16* p = buf + static_cast<size_t>(row0 + warp * 16) * cols + c;
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = __ldg(reinterpret_cast<const uint4*>(p + static_cast<size_t>(r) * cols));
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[r]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc < cols) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
        part[static_cast<size_t>(rb) * cols + cc] = sum;
    }
}

// part holds one row of column sums per `rows_per_part` packed rows (128: segment_colsum_partial_kernel; 32: the
// slab sums the dgelu epilogue of the grouped GEMM leaves behind).  CTA = 64 columns x 16 slices of the expert's parts
// (slice s takes parts b0 + s, b0 + s + 16, ...), four independent chains per thread, slices combined in a fixed order:
// one dependent chain per column took 25 us for the 104 slabs per expert of config 2, four slices 9.8 us.
constexpr int kSegFinSlices = 16;
__global__ void __launch_bounds__(64 * kSegFinSlices)
segment_colsum_final_kernel(const float* __restrict__ part, const int* __restrict__ seg_start, int cols,
                            float* __restrict__ out, int rows_per_part) {
    __shared__ float red[kSegFinSlices][64];
    const int e = blockIdx.y;
    const int cl = threadIdx.x & 63, slice = threadIdx.x >> 6;
    const int c = blockIdx.x * 64 + cl;
    const int b0 = __ldg(seg_start + e) / rows_per_part, b1 = __ldg(seg_start + e + 1) / rows_per_part;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (c < cols) {
        int b = b0 + slice;
        for (; b + 3 * kSegFinSlices < b1; b += 4 * kSegFinSlices) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += part[static_cast<size_t>(b + kSegFinSlices * i) * cols + c];
        }
        for (int i = 0; b < b1; b += kSegFinSlices, ++i) acc[i] += part[static_cast<size_t>(b) * cols + c];
    }
    red[slice][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (slice == 0 && c < cols) {
        float v = 0.0f;
#pragma unroll
        for (int sl = 0; sl < kSegFinSlices; sl += 4) v += (red[sl][cl] + red[sl + 1][cl]) + (red[sl + 2][cl] + red[sl + 3][cl]);
        out[static_cast<size_t>(e) * cols + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Expert parallelism (receive side).  After the dispatch all-to-all a rank holds, for each source rank s
// and each LOCAL expert e, a fixed slab recv[s][e][slab_rows][d] of which the first kept_recv[s][e] rows
// are live.  ep_tables turns the W x E_local counts into the standard packed layout (expert-contiguous,
// 256-aligned segments, sources in rank order inside a segment) and ep_repack moves rows between the two
// layouts, so the expert FFN kernels run unchanged on the packed buffer.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ep_tables_kernel(const int* __restrict__ kept_recv, int W, int El, int* __restrict__ slab_dst, int* __restrict__ kept_loc,
                 int* __restrict__ seg_start, int* __restrict__ tile_expert, int* __restrict__ num_mtiles, int max_mtiles) {
    extern __shared__ int smem_i[];
    int* seg_s = smem_i;  // [El + 1]
    if (threadIdx.x == 0) {
        int start = 0;
        for (int e = 0; e < El; ++e) {
            seg_s[e] = start;
            seg_start[e] = start;
            int tot = 0;
            for (int s = 0; s < W; ++s) {
                slab_dst[s * El + e] = start + tot;   // first packed row of source s inside expert e's segment
                tot += kept_recv[s * El + e];
            }
            kept_loc[e] = tot;
            start += (tot + MOE_ROW_ALIGN - 1) / MOE_ROW_ALIGN * MOE_ROW_ALIGN;
        }
        seg_s[El] = start;
        seg_start[El] = start;
        *num_mtiles = start / MOE_ROW_ALIGN;
    }
    __syncthreads();
    const int nm = seg_s[El] / MOE_ROW_ALIGN;
    for (int m = threadIdx.x; m < max_mtiles; m += 256) {
        int e = -1;
        if (m < nm) {
            const int row = m * MOE_ROW_ALIGN;
            e = 0;
            while (e + 1 < El && seg_s[e + 1] <= row) ++e;
        }
        tile_expert[m] = e;
    }
}

// to_packed = 1: packed[slab_dst[s,e] + i] = slabs[(s*El + e)*slab_rows + i] for i < kept_recv[s,e]; pad rows of
//                every packed segment are zeroed (blocks >= W*El do that).
// to_packed = 0: the inverse copy; rows of a slab beyond kept are left untouched (the receiver never reads them).
__global__ void __launch_bounds__(256)
ep_repack_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, const int* __restrict__ kept_recv,
                 const int* __restrict__ slab_dst, const int* __restrict__ seg_start, const int* __restrict__ kept_loc,
                 int W, int El, long long slab_rows, int d, int to_packed, int row_blocks) {
    const int pair = blockIdx.x / row_blocks, rb = blockIdx.x - pair * row_blocks;
    if (pair >= W * El) {
        if (to_packed && rb == 0) zero_pad_rows(dst, d, seg_start, kept_loc, pair - W * El, nullptr);
        return;
    }
    const int n = kept_recv[pair];
    const size_t slab0 = static_cast<size_t>(pair) * slab_rows, pk0 = static_cast<size_t>(slab_dst[pair]);
    const int per_row = d / 8;
    const int r0 = rb * 64, r1 = min(n, r0 + 64);
    for (int it = threadIdx.x; it < (r1 - r0) * per_row; it += 256) {
        const int r = r0 + it / per_row, c = (it % per_row) * 8;
        const size_t a = (slab0 + r) * d + c, b = (pk0 + r) * d + c;
        if (to_packed) *reinterpret_cast<uint4*>(dst + b) = __ldg(reinterpret_cast<const uint4*>(src + a));
        else *reinterpret_cast<uint4*>(dst + a) = __ldg(reinterpret_cast<const uint4*>(src + b));
    }
}

// ================================================================================================
// host launchers (C++ linkage; the extern "C" surface lives in api.cu)
// ================================================================================================
static int pick_eg(int E, int d) {
    int eg = 64;
    while (eg > 8 && (eg / 2 >= E || static_cast<size_t>(eg) * d * 4 > 65536)) eg /= 2;
    return eg;
}

template <typename XT>
static cudaError_t launch_gate_fwd_t(const XT* x, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask, int64_t T, int d, int E, int k,
                                     int score_mode, int want_psum, float* logits, int* idx, float* score,
                                     int* tile_hist, float* tile_psum, cudaStream_t st) {
    const int ntiles = static_cast<int>((T + kTokTile - 1) / kTokTile);
    const int eg = pick_eg(E, d);
    const size_t smem = (static_cast<size_t>(eg) * d + static_cast<size_t>(kTokTile) * (E + 1) + 2 * kTokTile) * 4 +
                        static_cast<size_t>(E) * 4;
#define MOE_GATE_CASE(EGV)                                                                                         \
    case EGV: {                                                                                                    \
        auto kfn = gate_fwd_kernel<XT, EGV>;                                                                       \
        cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (err != cudaSuccess) return err;                                                                        \
        kfn<<<ntiles, 256, smem, st>>>(x, Wg, bg, noise, token_mask, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, \
                                       tile_psum);                                                                 \
        break;                                                                                                     \
    }
    switch (eg) {
        MOE_GATE_CASE(8)
        MOE_GATE_CASE(16)
        MOE_GATE_CASE(32)
        MOE_GATE_CASE(64)
    }
#undef MOE_GATE_CASE
    return cudaGetLastError();
}

cudaError_t launch_gate_fwd(const void* x, int x_dtype, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask, int64_t T, int d, int E, int k,
                            int score_mode, int want_psum, float* logits, int* idx, float* score, int* tile_hist,
                            float* tile_psum, cudaStream_t st) {
    if (x_dtype == MOE_DTYPE_F32)
        return launch_gate_fwd_t(static_cast<const float*>(x), Wg, bg, noise, token_mask, T, d, E, k, score_mode, want_psum, logits, idx,
                                 score, tile_hist, tile_psum, st);
    return launch_gate_fwd_t(static_cast<const __nv_bfloat16*>(x), Wg, bg, noise, token_mask, T, d, E, k, score_mode, want_psum, logits,
                             idx, score, tile_hist, tile_psum, st);
}

cudaError_t launch_route_scan(const int* tile_hist, const float* tile_psum, int ntiles, int E, long long capacity,
                              int* tile_base, int* count, int* kept, int* seg_start, int* tile_expert, int* num_mtiles,
                              int max_mtiles, float* psum, int aux_mode, long long tokens, int k, float* aux_loss,
                              float* aux_coef, long long slab_rows, cudaStream_t st) {
    route_scan_rows_kernel<<<E, 256, 0, st>>>(tile_hist, tile_psum, ntiles, tile_base, count, psum);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const size_t smem = (3 * static_cast<size_t>(E) + 1) * 4;
    route_scan_kernel<<<1, 1024, smem, st>>>(E, capacity, count, kept, seg_start, tile_expert, num_mtiles, max_mtiles, psum,
                                             aux_mode, tokens, k, aux_loss, aux_coef, slab_rows);
    return cudaGetLastError();
}

cudaError_t launch_dispatch_fwd_rows(const void* x, int x_dtype, const int* idx, const int* tile_base, const int* seg_start,
                                     int64_t T, int d, int E, int k, long long capacity, int* pos, int* row_src,
                                     const PeerRows& xrows, void* xpad, const int* pad_seg, const int* pad_kept, int n_pad,
                                     cudaStream_t st) {
    const int ntiles = static_cast<int>((T + kTokTile - 1) / kTokTile);
    const size_t smem = (static_cast<size_t>(kTokTile * k / 32) * E + static_cast<size_t>(kTokTile) * k) * 4;
    auto xp = static_cast<__nv_bfloat16*>(xpad);
    cudaError_t err;
    if (x_dtype == MOE_DTYPE_F32) {
        auto kfn = dispatch_fwd_kernel<float>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles + n_pad, 256, smem, st>>>(static_cast<const float*>(x), idx, tile_base, seg_start, T, d, E, k,
                                               capacity, ntiles, pos, row_src, xrows, xp, pad_seg, pad_kept);
    } else {
        auto kfn = dispatch_fwd_kernel<__nv_bfloat16>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles + n_pad, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(x), idx, tile_base, seg_start, T, d,
                                               E, k, capacity, ntiles, pos, row_src, xrows, xp, pad_seg, pad_kept);
    }
    return cudaGetLastError();
}

cudaError_t launch_dispatch_fwd(const void* x, int x_dtype, const int* idx, const int* tile_base, const int* seg_start,
                                const int* kept, int64_t T, int d, int E, int k, long long capacity, int* pos,
                                int* row_src, void* xbuf, cudaStream_t st) {
    return launch_dispatch_fwd_rows(x, x_dtype, idx, tile_base, seg_start, T, d, E, k, capacity, pos, row_src, local_rows(xbuf),
                                    xbuf, seg_start, kept, E, st);
}

static int grid_for(int64_t items, int sm_count) {
    int64_t blocks = (items + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count) * 8;  // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

cudaError_t launch_combine_fwd_rows(const PeerRows& yrows, const int* pos, const float* score, int64_t T, int d, int k, void* out,
                                    int out_dtype, int sm_count, cudaStream_t st) {
    const int grid = grid_for(T * (d / 8), sm_count);
    if (out_dtype == MOE_DTYPE_F32)
        combine_fwd_kernel<float><<<grid, 256, 0, st>>>(yrows, pos, score, T, d, k, static_cast<float*>(out));
    else
        combine_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(yrows, pos, score, T, d, k, static_cast<__nv_bfloat16*>(out));
    return cudaGetLastError();
}

cudaError_t launch_combine_fwd(const void* ybuf, const int* pos, const float* score, int64_t T, int d, int k, void* out,
                               int out_dtype, int sm_count, cudaStream_t st) {
    return launch_combine_fwd_rows(local_rows(ybuf), pos, score, T, d, k, out, out_dtype, sm_count, st);
}

cudaError_t launch_combine_bwd_rows(const void* dy, int dy_dtype, const PeerRows& yrows, const int* pos, const float* score,
                                    const int* pad_seg, const int* pad_kept, int n_pad, int64_t T, int d, int k,
                                    const PeerRows& dyrows, void* dypad, float* dscore, cudaStream_t st) {
    const int nb = static_cast<int>((T + 7) / 8);
    auto db = static_cast<__nv_bfloat16*>(dypad);
    if (dy_dtype == MOE_DTYPE_F32)
        combine_bwd_kernel<float><<<nb + n_pad, 256, 0, st>>>(static_cast<const float*>(dy), yrows, pos, score, pad_seg, pad_kept,
                                                              T, d, k, nb, dyrows, db, dscore);
    else
        combine_bwd_kernel<__nv_bfloat16><<<nb + n_pad, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), yrows, pos, score,
                                                                      pad_seg, pad_kept, T, d, k, nb, dyrows, db, dscore);
    return cudaGetLastError();
}

cudaError_t launch_combine_bwd(const void* dy, int dy_dtype, const void* ybuf, const int* pos, const float* score,
                               const int* seg_start, const int* kept, int64_t T, int d, int k, int E, void* dybuf,
                               float* dscore, cudaStream_t st) {
    return launch_combine_bwd_rows(dy, dy_dtype, local_rows(ybuf), pos, score, seg_start, kept, E, T, d, k, local_rows(dybuf), dybuf,
                                   dscore, st);
}

cudaError_t launch_gate_bwd(const float* logits, const int* idx, const float* score, const float* dscore,
                            const float* dpsum, int64_t T, int E, int k, int score_mode, float* dlogits,
                            cudaStream_t st) {
    const int grid = static_cast<int>((T + 255) / 256);
    gate_bwd_kernel<<<grid, 256, 0, st>>>(logits, idx, score, dscore, dpsum, T, E, k, score_mode, dlogits);
    return cudaGetLastError();
}

cudaError_t launch_dispatch_bwd(const void* dxbuf, const int* pos, const float* dlogits, const int* idx, const float* Wg,
                                int64_t T, int d, int E, int k, int dense_dlogits, void* dx, int dx_dtype, int sm_count,
                                cudaStream_t st) {
    const int grid = grid_for(T * (d / 8), sm_count);
    auto xb = static_cast<const __nv_bfloat16*>(dxbuf);
    if (dx_dtype == MOE_DTYPE_F32)
        dispatch_bwd_kernel<float><<<grid, 256, 0, st>>>(xb, pos, dlogits, idx, Wg, T, d, E, k, dense_dlogits,
                                                         static_cast<float*>(dx));
    else
        dispatch_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(xb, pos, dlogits, idx, Wg, T, d, E, k, dense_dlogits,
                                                                 static_cast<__nv_bfloat16*>(dx));
    return cudaGetLastError();
}

cudaError_t launch_gate_dispatch_bwd(const void* dxbuf, const int* pos, const float* logits, const int* idx, const float* score,
                                     const float* dscore, const float* dpsum, const float* Wg, int64_t T, int d, int E, int k,
                                     int score_mode, float* dlogits, void* dx, int dx_dtype, cudaStream_t st) {
    if (gate_dispatch_bwd_mma_supported(d, E, k))   // E <= 64: dlogits Wg on the tensor cores (csrc/gate_bwd_mma.cu)
        return launch_gate_dispatch_bwd_mma(local_rows(dxbuf), pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode,
                                            dlogits, dx, dx_dtype, st);
    const int ntiles = static_cast<int>((T + kTokTile - 1) / kTokTile);
    // staged gather: the largest sub-tile (tokens) whose k rows each fit 64 KB, so that 2-3 CTAs stay resident per SM
    int ts = 0;
    if (dxbuf != nullptr && d % 8 == 0 && E % 4 == 0) {
        ts = kTokTile;
        while (ts > 4 && static_cast<size_t>(ts) * k * d * 2 > 65536) ts >>= 1;
        if (static_cast<size_t>(ts) * k * d * 2 > 98304) ts = 0;
    }
    const size_t smem = static_cast<size_t>(kTokTile) * (E + k) * 4 + static_cast<size_t>(ts) * k * d * 2;
    auto xb = static_cast<const __nv_bfloat16*>(dxbuf);
    cudaError_t err;
    if (dx_dtype == MOE_DTYPE_F32) {
        auto kfn = gate_dispatch_bwd_kernel<float>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles, 256, smem, st>>>(xb, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits,
                                       static_cast<float*>(dx), ts);
    } else {
        auto kfn = gate_dispatch_bwd_kernel<__nv_bfloat16>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles, 256, smem, st>>>(xb, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits,
                                       static_cast<__nv_bfloat16*>(dx), ts);
    }
    return cudaGetLastError();
}

static int gate_wgrad_blocks(int64_t T, int sm_count) {
    const int64_t ntiles = (T + kWgTile - 1) / kWgTile;
    return static_cast<int>(ntiles < 2LL * sm_count ? ntiles : 2LL * sm_count);
}

size_t gate_wgrad_workspace_bytes(int64_t T, int d, int E) {
    const size_t nb = static_cast<size_t>(gate_wgrad_blocks(T, sm_count()));
    return nb * (static_cast<size_t>(E) * d + E) * 4;
}

template <int MT>
static cudaError_t launch_gate_wgrad_mma(const float* dlogits, const __nv_bfloat16* x, int64_t T, int d, int E, int ntiles, int nb,
                                         int slices, float* part_w, float* part_b, cudaStream_t st) {
    const int ds = d / slices;
    const size_t smem = 2 * static_cast<size_t>(kWgSub) * (ds + 8) * 2 + 2 * static_cast<size_t>(kWgSub) * (16 * MT + 8) * 4;
    auto kfn = gate_wgrad_partial_mma_kernel<MT>;
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kfn<<<dim3(nb, slices), 256, smem, st>>>(dlogits, x, T, d, E, ntiles, part_w, part_b);
    return cudaGetLastError();
}

cudaError_t launch_gate_wgrad(const float* dlogits, const void* x, int x_dtype, int64_t T, int d, int E, void* workspace,
                              float* dWg, float* dbg, cudaStream_t st) {
    const int ntiles = static_cast<int>((T + kWgTile - 1) / kWgTile);
    int nb = gate_wgrad_blocks(T, sm_count());
    float* part_w = static_cast<float*>(workspace);
    cudaError_t err;
    if (x_dtype == MOE_DTYPE_F32) {
        float* part_b = part_w + static_cast<size_t>(nb) * E * d;
        const int CG = d / 4, TG = CG + 1 > 256 ? 1 : 256 / (CG + 1);
        const size_t smem = (static_cast<size_t>(kWgTile) * E + (TG > 1 ? static_cast<size_t>(TG) * kWgEG * (d + 4) : 0)) * 4;
        auto kfn = gate_wgrad_partial_kernel<float>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<nb, 256, smem, st>>>(dlogits, static_cast<const float*>(x), T, d, E, ntiles, part_w, part_b);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        const int n = E * d + E;
        gate_wgrad_reduce_kernel<<<(n + 63) / 64, 64 * kWgRedSlices, 0, st>>>(part_w, part_b, nb, d, E, dWg, dbg);
        return cudaGetLastError();
    }
    // bf16 activations: tensor cores, every expert in one pass over x.  MT 16-expert m-tiles x (d / 64 / slices) n-tiles per
    // warp must fit 16 accumulator tiles: wide layers are cut into feature slices (grid.y), and the token blocks shrink by
    // the same factor so that the grid stays at two CTAs per SM and the partials to reduce get fewer.
    const int mt = E <= 16 ? 1 : E <= 32 ? 2 : 4;
    const int nt_full = d / 64;
    int slices = 1;
    while (slices < nt_full && (nt_full % slices != 0 || mt * (nt_full / slices) > 16)) ++slices;
    nb = nb / slices > 0 ? nb / slices : 1;
    float* part_b = part_w + static_cast<size_t>(nb) * E * d;
    auto xb = static_cast<const __nv_bfloat16*>(x);
    err = mt == 1 ? launch_gate_wgrad_mma<1>(dlogits, xb, T, d, E, ntiles, nb, slices, part_w, part_b, st)
        : mt == 2 ? launch_gate_wgrad_mma<2>(dlogits, xb, T, d, E, ntiles, nb, slices, part_w, part_b, st)
                  : launch_gate_wgrad_mma<4>(dlogits, xb, T, d, E, ntiles, nb, slices, part_w, part_b, st);
    if (err != cudaSuccess) return err;
    const int n = E * d + E;
    gate_wgrad_reduce_kernel<<<(n + 63) / 64, 64 * kWgRedSlices, 0, st>>>(part_w, part_b, nb, d, E, dWg, dbg);
    return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* src, void* dst, int64_t n, int sm_count, cudaStream_t st) {
    const int64_t n8 = n / 8;
    cast_bf16_kernel<<<grid_for(n8, sm_count), 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n8);
    return cudaGetLastError();
}

cudaError_t launch_cast_bf16_pair(const float* src0, void* dst0, int64_t n0, const float* src1, void* dst1, int64_t n1,
                                  int sm_count, cudaStream_t st) {
    const int64_t n8_0 = n0 / 8, n8_1 = n1 / 8;
    cast_bf16_pair_kernel<<<grid_for((n8_0 + n8_1 + 3) / 4, sm_count), 256, 0, st>>>(src0, static_cast<__nv_bfloat16*>(dst0), n8_0, src1,
                                                                                 static_cast<__nv_bfloat16*>(dst1), n8_1);
    return cudaGetLastError();
}

size_t segment_colsum_workspace_bytes(int64_t rows_cap, int cols) {
    return static_cast<size_t>(rows_cap / 128) * cols * 4;
}

cudaError_t launch_segment_colsum(const void* buf, const int* seg_start, int64_t rows_cap, int E, int cols, void* workspace,
                                  float* out, cudaStream_t st) {
    float* part = static_cast<float*>(workspace);
    dim3 g1((cols + 255) / 256, static_cast<unsigned>(rows_cap / 128));
    segment_colsum_partial_kernel<<<g1, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(buf), seg_start, E, cols, part);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    dim3 g2((cols + 63) / 64, E);
    segment_colsum_final_kernel<<<g2, 64 * kSegFinSlices, 0, st>>>(part, seg_start, cols, out, 128);
    return cudaGetLastError();
}

cudaError_t launch_slab_colsum_final(const float* part, const int* seg_start, int E, int cols, float* out, cudaStream_t st) {
    dim3 g2((cols + 63) / 64, E);
    segment_colsum_final_kernel<<<g2, 64 * kSegFinSlices, 0, st>>>(part, seg_start, cols, out, 32);
    return cudaGetLastError();
}

cudaError_t launch_ep_tables(const int* kept_recv, int W, int El, int* slab_dst, int* kept_loc, int* seg_start,
                             int* tile_expert, int* num_mtiles, int max_mtiles, cudaStream_t st) {
    ep_tables_kernel<<<1, 256, (El + 1) * sizeof(int), st>>>(kept_recv, W, El, slab_dst, kept_loc, seg_start, tile_expert,
                                                            num_mtiles, max_mtiles);
    return cudaGetLastError();
}

cudaError_t launch_ep_repack(const void* src, void* dst, const int* kept_recv, const int* slab_dst, const int* seg_start,
                             const int* kept_loc, int W, int El, long long slab_rows, int d, int to_packed, cudaStream_t st) {
    const int row_blocks = static_cast<int>((slab_rows + 63) / 64);
    const int grid = W * El * row_blocks + (to_packed ? El * row_blocks : 0);
    ep_repack_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), kept_recv,
                                           slab_dst, seg_start, kept_loc, W, El, slab_rows, d, to_packed, row_blocks);
    return cudaGetLastError();
}

}  // namespace moe
