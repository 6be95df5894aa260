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
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace moe {

constexpr int kTokTile = 256;  // tokens per routing tile (one CTA)
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
                const float* __restrict__ noise, int64_t T, int d, int E, int k, int score_mode, int want_psum, float* __restrict__ logits, int* __restrict__ idx,
                float* __restrict__ score, int* __restrict__ tile_hist, float* __restrict__ tile_psum) {
    constexpr int NT = 64 / EG;  // tokens per warp pass: NT*EG = 64 accumulators per lane
    extern __shared__ float smem_f[];
    float* wg_s = smem_f;                          // [EG][d]
    float* lg_s = wg_s + EG * d;                   // [256][E+1]
    float* m_s = lg_s + kTokTile * (E + 1);        // [256] row max
    float* rz_s = m_s + kTokTile;                  // [256] 1/Z
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

        for (int it = 0; it < 32 / NT; ++it) {
            const int64_t tok0 = t_base + warp * 32 + it * NT;
            if (tok0 >= T) break;  // warp-uniform
            float acc[NT * EG];
#pragma unroll
            for (int i = 0; i < NT * EG; ++i) acc[i] = 0.0f;
            for (int c = 0; c < n_chunks; ++c) {
                const int i0 = c * 128 + lane * 4;
                if (i0 < d) {
                    float4 xv[NT];
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        xv[n] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tok0 + n < T) xv[n] = load_x4(x + (tok0 + n) * d + i0);
                    }
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
        if (tok < T) {
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
    for (int e = tid; e < E; e += 256) tile_hist[static_cast<size_t>(blockIdx.x) * E + e] = hist_s[e];

    if (want_psum) {
        const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTokTile), T - t_base));
        for (int e = warp; e < E; e += 8) {
            float part = 0.0f;
            for (int t = lane; t < n_tok; t += 32) part += expf(lg_s[t * ldl + e] - m_s[t]) / rz_s[t];
            part = warp_sum_xor(part);
            if (lane == 0) tile_psum[static_cast<size_t>(blockIdx.x) * E + e] = part;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: scan of the per-tile histograms -> tile bases, counts, capacity-clamped segments, GEMM tile table
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
route_scan_kernel(const int* __restrict__ tile_hist, const float* __restrict__ tile_psum, int ntiles, int E,
                  long long capacity, int* __restrict__ tile_base, int* __restrict__ count, int* __restrict__ kept,
                  int* __restrict__ seg_start, int* __restrict__ tile_expert, int* __restrict__ num_mtiles,
                  int max_mtiles, float* __restrict__ psum) {
    extern __shared__ int smem_i[];
    int* cnt_s = smem_i;          // [E]
    int* seg_s = smem_i + E;      // [E+1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = warp; e < E; e += 32) {
        int running = 0;
        for (int b0 = 0; b0 < ntiles; b0 += 32) {
            const int b = b0 + lane;
            const int v = b < ntiles ? tile_hist[static_cast<size_t>(b) * E + e] : 0;
            int incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            if (b < ntiles) tile_base[static_cast<size_t>(b) * E + e] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) cnt_s[e] = running;
        if (tile_psum != nullptr && psum != nullptr) {
            float part = 0.0f;
            for (int b = lane; b < ntiles; b += 32) part += tile_psum[static_cast<size_t>(b) * E + e];
            part = warp_sum_xor(part);
            if (lane == 0) psum[e] = part;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int start = 0;
        for (int e = 0; e < E; ++e) {
            const int c = cnt_s[e];
            const int kp = static_cast<long long>(c) < capacity ? c : static_cast<int>(capacity);
            count[e] = c;
            kept[e] = kp;
            seg_s[e] = start;
            seg_start[e] = start;
            start += (kp + MOE_ROW_ALIGN - 1) / MOE_ROW_ALIGN * MOE_ROW_ALIGN;
        }
        seg_s[E] = start;
        seg_start[E] = start;
        *num_mtiles = start / MOE_ROW_ALIGN;
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
template <typename XT>
__global__ void __launch_bounds__(256)
dispatch_fwd_kernel(const XT* __restrict__ x, const int* __restrict__ idx, const int* __restrict__ tile_base,
                    const int* __restrict__ seg_start, const int* __restrict__ kept, int64_t T, int d, int E, int k,
                    long long capacity, int ntiles, int* __restrict__ pos, int* __restrict__ row_src,
                    __nv_bfloat16* __restrict__ xbuf) {
    if (static_cast<int>(blockIdx.x) >= ntiles) {
        zero_pad_rows(xbuf, d, seg_start, kept, blockIdx.x - ntiles, row_src);
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
        const bool valid = il < n_ent;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int e = idx[i_base + il];
            const unsigned peers = __match_any_sync(act, e);
            const int rk = __popc(peers & ((1u << lane) - 1u));
            if (rk == 0) chunk_cnt[ch * E + e] = __popc(peers);
            ent_row[il] = rk;  // temporarily: rank inside chunk
        }
    }
    __syncthreads();
    // B: exclusive scan over chunks per expert, seeded with this tile's global base
    for (int e = tid; e < E; e += 256) {
        int running = tile_base[static_cast<size_t>(blockIdx.x) * E + e];
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
        const long long rank = static_cast<long long>(chunk_cnt[(il >> 5) * E + e]) + ent_row[il];
        const int row = rank < capacity ? seg_start[e] + static_cast<int>(rank) : -1;
        pos[i_base + il] = row;
        if (row >= 0) row_src[row] = static_cast<int>(i_base + il);
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
        store8(xbuf + static_cast<size_t>(row) * d + c, v);
    }
}

// ------------------------------------------------------------------------------------------------
// K4: combine forward   out[t] = sum_j score[t,j] * Y[pos[t,j]]
// ------------------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const __nv_bfloat16* __restrict__ ybuf, const int* __restrict__ pos, const float* __restrict__ score,
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
            load8(ybuf + static_cast<size_t>(row) * d + c, y);
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
combine_bwd_kernel(const GT* __restrict__ dy, const __nv_bfloat16* __restrict__ ybuf, const int* __restrict__ pos,
                   const float* __restrict__ score, const int* __restrict__ seg_start, const int* __restrict__ kept,
                   int64_t T, int d, int k, int E, int n_tok_blocks, __nv_bfloat16* __restrict__ dybuf,
                   float* __restrict__ dscore) {
    if (static_cast<int>(blockIdx.x) >= n_tok_blocks) {
        zero_pad_rows(dybuf, d, seg_start, kept, blockIdx.x - n_tok_blocks, nullptr);
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
            for (int c = lane * 8; c < d; c += 256) {
                float g[8], y[8], o[8];
                load8(dy + t * d + c, g);
                load8(ybuf + static_cast<size_t>(row) * d + c, y);
#pragma unroll
                for (int i = 0; i < 8; ++i) { dot = fmaf(g[i], y[i], dot); o[i] = s * g[i]; }
                store8(dybuf + static_cast<size_t>(row) * d + c, o);
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
    if (need_p) {
        m = lr[pk[0]];
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
        dl[e] = v;
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
// K8: gate weight gradient   dWg = dlogits^T x, dbg = colsum(dlogits)   (two deterministic stages)
// ------------------------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(256)
gate_wgrad_partial_kernel(const float* __restrict__ dlogits, const XT* __restrict__ x, int64_t T, int d, int E,
                          float* __restrict__ part_w, float* __restrict__ part_b) {
    extern __shared__ float smem_f[];
    float* dl_s = smem_f;  // [256][E]
    const int tid = threadIdx.x;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kTokTile;
    const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTokTile), T - t_base));
    for (int i = tid; i < n_tok * E; i += 256) dl_s[i] = dlogits[t_base * E + i];
    __syncthreads();
    float* pw = part_w + static_cast<size_t>(blockIdx.x) * E * d;
    for (int e0 = 0; e0 < E; e0 += 16) {
        for (int c = tid; c < d; c += 256) {
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
            for (int t = 0; t < n_tok; ++t) {
                const float xv = static_cast<float>(x[(t_base + t) * d + c]);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (e0 + i < E) acc[i] = fmaf(dl_s[t * E + e0 + i], xv, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (e0 + i < E) pw[static_cast<size_t>(e0 + i) * d + c] = acc[i];
        }
    }
    for (int e = tid; e < E; e += 256) {
        float s = 0.0f;
        for (int t = 0; t < n_tok; ++t) s += dl_s[t * E + e];
        part_b[static_cast<size_t>(blockIdx.x) * E + e] = s;
    }
}

__global__ void __launch_bounds__(256)
gate_wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int ntiles, int d, int E,
                         float* __restrict__ dWg, float* __restrict__ dbg) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = E * d;
    if (i < n) {
        float s = 0.0f;
        for (int b = 0; b < ntiles; ++b) s += part_w[static_cast<size_t>(b) * n + i];
        dWg[i] = s;
    } else if (i < n + E && dbg != nullptr) {
        const int e = i - n;
        float s = 0.0f;
        for (int b = 0; b < ntiles; ++b) s += part_b[static_cast<size_t>(b) * E + e];
        dbg[e] = s;
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

// src[E,R,C] fp32 -> dst[E,R,C] bf16 (optional) + dst_t[E,C,R] bf16; 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
cast_bf16_transposed_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                            __nv_bfloat16* __restrict__ dst_t, int R, int C) {
    __shared__ float tile[32][33];
    const int e = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t base = static_cast<size_t>(e) * R * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8;
        const float v = src[base + static_cast<size_t>(r) * C + c0 + tx];
        tile[ty + i * 8][tx] = v;
        if (dst != nullptr) dst[base + static_cast<size_t>(r) * C + c0 + tx] = __float2bfloat16_rn(v);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8;
        dst_t[base + static_cast<size_t>(c) * R + r0 + tx] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
    }
}

// out[e, c] = sum over rows r in [seg_start[e], seg_start[e+1]) of buf[r, c]; block = 32 x 8 threads,
// each lane owns 2 adjacent columns, the 8 row-lanes stride the segment; fixed-order smem reduction.
__global__ void __launch_bounds__(256)
segment_colsum_kernel(const __nv_bfloat16* __restrict__ buf, const int* __restrict__ seg_start, int cols,
                      float* __restrict__ out) {
    __shared__ float red[8][64];
    const int e = blockIdx.y;
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + lane * 2;
    const int r0 = seg_start[e], r1 = seg_start[e + 1];
    float a0 = 0.0f, a1 = 0.0f;
    if (c < cols) {
        for (int r = r0 + ry; r < r1; r += 8) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(buf + static_cast<size_t>(r) * cols + c);
            a0 += __low2float(v);
            a1 += __high2float(v);
        }
    }
    red[ry][lane * 2] = a0;
    red[ry][lane * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        const int cc = blockIdx.x * 64 + threadIdx.x;
        if (cc < cols) out[static_cast<size_t>(e) * cols + cc] = s;
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
static cudaError_t launch_gate_fwd_t(const XT* x, const float* Wg, const float* bg, const float* noise, int64_t T, int d, int E, int k,
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
        kfn<<<ntiles, 256, smem, st>>>(x, Wg, bg, noise, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist, \
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

cudaError_t launch_gate_fwd(const void* x, int x_dtype, const float* Wg, const float* bg, const float* noise, int64_t T, int d, int E, int k,
                            int score_mode, int want_psum, float* logits, int* idx, float* score, int* tile_hist,
                            float* tile_psum, cudaStream_t st) {
    if (x_dtype == MOE_DTYPE_F32)
        return launch_gate_fwd_t(static_cast<const float*>(x), Wg, bg, noise, T, d, E, k, score_mode, want_psum, logits, idx,
                                 score, tile_hist, tile_psum, st);
    return launch_gate_fwd_t(static_cast<const __nv_bfloat16*>(x), Wg, bg, noise, T, d, E, k, score_mode, want_psum, logits,
                             idx, score, tile_hist, tile_psum, st);
}

cudaError_t launch_route_scan(const int* tile_hist, const float* tile_psum, int ntiles, int E, long long capacity,
                              int* tile_base, int* count, int* kept, int* seg_start, int* tile_expert, int* num_mtiles,
                              int max_mtiles, float* psum, cudaStream_t st) {
    const size_t smem = (2 * static_cast<size_t>(E) + 1) * 4;
    route_scan_kernel<<<1, 1024, smem, st>>>(tile_hist, tile_psum, ntiles, E, capacity, tile_base, count, kept,
                                             seg_start, tile_expert, num_mtiles, max_mtiles, psum);
    return cudaGetLastError();
}

cudaError_t launch_dispatch_fwd(const void* x, int x_dtype, const int* idx, const int* tile_base, const int* seg_start,
                                const int* kept, int64_t T, int d, int E, int k, long long capacity, int* pos,
                                int* row_src, void* xbuf, cudaStream_t st) {
    const int ntiles = static_cast<int>((T + kTokTile - 1) / kTokTile);
    const size_t smem = (static_cast<size_t>(kTokTile * k / 32) * E + static_cast<size_t>(kTokTile) * k) * 4;
    auto xb = static_cast<__nv_bfloat16*>(xbuf);
    cudaError_t err;
    if (x_dtype == MOE_DTYPE_F32) {
        auto kfn = dispatch_fwd_kernel<float>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles + E, 256, smem, st>>>(static_cast<const float*>(x), idx, tile_base, seg_start, kept, T, d, E, k,
                                           capacity, ntiles, pos, row_src, xb);
    } else {
        auto kfn = dispatch_fwd_kernel<__nv_bfloat16>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles + E, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(x), idx, tile_base, seg_start, kept, T, d,
                                           E, k, capacity, ntiles, pos, row_src, xb);
    }
    return cudaGetLastError();
}

static int grid_for(int64_t items, int sm_count) {
    int64_t blocks = (items + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count) * 8;  // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

cudaError_t launch_combine_fwd(const void* ybuf, const int* pos, const float* score, int64_t T, int d, int k, void* out,
                               int out_dtype, int sm_count, cudaStream_t st) {
    const int grid = grid_for(T * (d / 8), sm_count);
    auto yb = static_cast<const __nv_bfloat16*>(ybuf);
    if (out_dtype == MOE_DTYPE_F32)
        combine_fwd_kernel<float><<<grid, 256, 0, st>>>(yb, pos, score, T, d, k, static_cast<float*>(out));
    else
        combine_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(yb, pos, score, T, d, k, static_cast<__nv_bfloat16*>(out));
    return cudaGetLastError();
}

cudaError_t launch_combine_bwd(const void* dy, int dy_dtype, const void* ybuf, const int* pos, const float* score,
                               const int* seg_start, const int* kept, int64_t T, int d, int k, int E, void* dybuf,
                               float* dscore, cudaStream_t st) {
    const int nb = static_cast<int>((T + 7) / 8);
    auto yb = static_cast<const __nv_bfloat16*>(ybuf);
    auto db = static_cast<__nv_bfloat16*>(dybuf);
    if (dy_dtype == MOE_DTYPE_F32)
        combine_bwd_kernel<float><<<nb + E, 256, 0, st>>>(static_cast<const float*>(dy), yb, pos, score, seg_start, kept,
                                                          T, d, k, E, nb, db, dscore);
    else
        combine_bwd_kernel<__nv_bfloat16><<<nb + E, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), yb, pos, score,
                                                                  seg_start, kept, T, d, k, E, nb, db, dscore);
    return cudaGetLastError();
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

size_t gate_wgrad_workspace_bytes(int64_t T, int d, int E) {
    const size_t ntiles = static_cast<size_t>((T + kTokTile - 1) / kTokTile);
    return ntiles * (static_cast<size_t>(E) * d + E) * 4;
}

cudaError_t launch_gate_wgrad(const float* dlogits, const void* x, int x_dtype, int64_t T, int d, int E, void* workspace,
                              float* dWg, float* dbg, cudaStream_t st) {
    const int ntiles = static_cast<int>((T + kTokTile - 1) / kTokTile);
    float* part_w = static_cast<float*>(workspace);
    float* part_b = part_w + static_cast<size_t>(ntiles) * E * d;
    const size_t smem = static_cast<size_t>(kTokTile) * E * 4;
    cudaError_t err;
    if (x_dtype == MOE_DTYPE_F32) {
        auto kfn = gate_wgrad_partial_kernel<float>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles, 256, smem, st>>>(dlogits, static_cast<const float*>(x), T, d, E, part_w, part_b);
    } else {
        auto kfn = gate_wgrad_partial_kernel<__nv_bfloat16>;
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kfn<<<ntiles, 256, smem, st>>>(dlogits, static_cast<const __nv_bfloat16*>(x), T, d, E, part_w, part_b);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const int n = E * d + E;
    gate_wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part_w, part_b, ntiles, d, E, dWg, dbg);
    return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* src, void* dst, int64_t n, int sm_count, cudaStream_t st) {
    const int64_t n8 = n / 8;
    cast_bf16_kernel<<<grid_for(n8, sm_count), 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n8);
    return cudaGetLastError();
}

cudaError_t launch_cast_bf16_transposed(const float* src, void* dst, void* dst_t, int E, int R, int C, cudaStream_t st) {
    dim3 grid(C / 32, R / 32, E);
    cast_bf16_transposed_kernel<<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst),
                                                      static_cast<__nv_bfloat16*>(dst_t), R, C);
    return cudaGetLastError();
}

cudaError_t launch_segment_colsum(const void* buf, const int* seg_start, int E, int cols, float* out, cudaStream_t st) {
    dim3 grid((cols + 63) / 64, E);
    segment_colsum_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(buf), seg_start, cols, out);
    return cudaGetLastError();
}

}  // namespace moe
