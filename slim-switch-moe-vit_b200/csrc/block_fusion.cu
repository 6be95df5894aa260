// block_fusion.cu — the two HBM passes around the MoE layer in the hosting transformer block, fused:
//   x_out = x_in + delta ;  n = LayerNorm(x_out) * gamma + beta          (one kernel instead of add + LN + cast)
// and its backward.  In the reference block (/root/reference/models/vision_transformer.py:319-322) these are
// `x + drop_path(self.mlp(self.norm2(x)))` -> the residual add that consumes the layer's output and the
// pre-norm that produces the next layer's input.  The residual stream stays fp32
// (what autocast does in the reference, engine.py:52); the normalised output is written directly in bf16,
// the dtype the MoE dispatch / attention projections read, so no separate cast kernel runs.
//
// One warp per row, the whole row in registers (d <= 1024, d % 4 == 0), two-pass mean / variance.
// Backward is persistent: every lane owns fixed columns, so the dgamma / dbeta partial sums live in
// registers across all rows of the CTA, are combined over the CTA's 8 warps in shared memory in warp order
// and written as one partial per CTA; a second kernel adds the partials in CTA order (deterministic).
// Its operands come through a per-warp ring of row slots in shared memory filled by cp.async.bulk
// (addln_bwd_bulk_kernel: 5.9 TB/s at d = 384 against 4.2 TB/s for the register-staged version, which
// remains for d % 8 != 0).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace moe {

constexpr int kLnMaxV = 8;  // float4 per lane: d <= 32 * 4 * 8 = 1024 (kernels are instantiated for NV = 2, 3, 6, 8)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_f4(const __nv_bfloat16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_f4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// delta (nullable): x_out = x_in + delta is written to x_out; otherwise x_out is not touched and x_in is normalised.
template <typename DT, typename NT, int NV>
__global__ void __launch_bounds__(256)
addln_fwd_kernel(const float* __restrict__ x_in, const DT* __restrict__ delta, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int64_t T, int d, float* __restrict__ x_out, NT* __restrict__ n,
                 float* __restrict__ mean, float* __restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= T) return;
    const int nv = d / 4;
    float4 v[NV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (i * 32 + lane < nv) {
            v[i] = ld_f4(x_in + row * d + c);
            if (delta != nullptr) {
                const float4 dl = ld_f4(delta + row * d + c);
                v[i].x += dl.x; v[i].y += dl.y; v[i].z += dl.z; v[i].w += dl.w;
                st_f4(x_out + row * d + c, v[i]);
            }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mu = warp_sum(s) / static_cast<float>(d);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i * 32 + lane < nv) {
            const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, e = v[i].w - mu;
            q += (a * a + b * b) + (c * c + e * e);
        }
    }
    const float rs = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (i * 32 + lane < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
            o.x = (v[i].x - mu) * rs * g.x + b.x; o.y = (v[i].y - mu) * rs * g.y + b.y;
            o.z = (v[i].z - mu) * rs * g.z + b.z; o.w = (v[i].w - mu) * rs * g.w + b.w;
            st_f4(n + row * d + c, o);
        }
    }
}

// dx_in = (dx_out or 0) + LN'(dn);  d_delta (nullable) = dx_in in the delta dtype;  partial dgamma / dbeta per CTA.
template <typename NT, typename DT, int NV>
__global__ void __launch_bounds__(256, NV <= 6 ? 2 : 1)
addln_bwd_kernel(const NT* __restrict__ dn, const float* __restrict__ dx_out, const float* __restrict__ x, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma, int64_t T, int d, float* __restrict__ dx_in,
                 DT* __restrict__ d_delta, float* __restrict__ part /* [grid][2][d] */) {
    extern __shared__ float red[];  // [8][2*d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = d / 4;
    float4 gam[NV], dg[NV], db[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        gam[i] = (i * 32 + lane < nv) ? __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_d = 1.0f / static_cast<float>(d);
    // One row per warp at a time, the NEXT row's operands (x, dn, the incoming residual gradient and the row
    // statistics) already in flight while this one is reduced and stored: every load of a row is issued at once,
    // two rows deep (ping-pong register sets a / b, no copies).  With the loads in two dependent phases per row
    // and a single row in flight the kernel ran at 3.6-3.9 TB/s (round 1c-1e).
    struct RowRegs { float4 x[NV], g[NV], r[NV]; float mu, rs; };
    auto load_row = [&](RowRegs& R, int64_t row) {
        R.mu = mean[row]; R.rs = rstd[row];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (i * 32 + lane < nv) {
                R.x[i] = ld_f4(x + row * d + c);
                R.g[i] = ld_f4(dn + row * d + c);
                if constexpr (NV <= 3) R.r[i] = dx_out != nullptr ? ld_f4(dx_out + row * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto do_row = [&](RowRegs& R, int64_t row) {
        const float mu = R.mu, rs = R.rs;
        float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i * 32 + lane < nv) {
                const float4 xv = R.x[i], gv = R.g[i];
                const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                dg[i].x += gv.x * xh.x; dg[i].y += gv.y * xh.y; dg[i].z += gv.z * xh.z; dg[i].w += gv.w * xh.w;
                db[i].x += gv.x; db[i].y += gv.y; db[i].z += gv.z; db[i].w += gv.w;
                const float4 g = make_float4(gv.x * gam[i].x, gv.y * gam[i].y, gv.z * gam[i].z, gv.w * gam[i].w);
                c1 += (g.x + g.y) + (g.z + g.w);
                c2 += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
                R.x[i] = xh; R.g[i] = g;
            }
        }
        c1 = warp_sum(c1) * inv_d;
        c2 = warp_sum(c2) * inv_d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (i * 32 + lane < nv) {
                const float4 xh = R.x[i], g = R.g[i];
                float4 r;   // wide rows: no registers left to hold it across the reduction
                if constexpr (NV <= 3) r = R.r[i];
                else r = dx_out != nullptr ? ld_f4(dx_out + row * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 o;
                o.x = rs * (g.x - c1 - xh.x * c2); o.y = rs * (g.y - c1 - xh.y * c2);
                o.z = rs * (g.z - c1 - xh.z * c2); o.w = rs * (g.w - c1 - xh.w * c2);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                st_f4(dx_in + row * d + c, o);
                if (d_delta != nullptr) st_f4(d_delta + row * d + c, o);
            }
        }
    };
    const int64_t stride = static_cast<int64_t>(gridDim.x) * 8;
    int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    if constexpr (NV <= 3) {
        RowRegs a, b;
        if (row < T) load_row(a, row);
        while (row < T) {
            if (row + stride < T) load_row(b, row + stride);
            do_row(a, row);
            row += stride;
            if (row >= T) break;
            if (row + stride < T) load_row(a, row + stride);
            do_row(b, row);
            row += stride;
        }
    } else {
        RowRegs a;
        for (; row < T; row += stride) {
            load_row(a, row);
            do_row(a, row);
        }
    }
    // combine the 8 warps of the CTA in warp order
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (i * 32 + lane < nv) {
            st_f4(red + warp * 2 * d + c, dg[i]);
            st_f4(red + warp * 2 * d + d + c, db[i]);
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * d; o += 256) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w * 2 * d + o];
        part[static_cast<size_t>(blockIdx.x) * 2 * d + o] = sum;
    }
}

// Same arithmetic, operands staged by the copy engine (d % 8 == 0).  The register version above holds at most two rows per
// warp in flight (its 128 registers are the per-lane dgamma / dbeta accumulators plus two rows of operands): ~80 KB of loads
// per SM, short of what 6.5 TB/s needs under load.  Here every warp owns a ring of `stages` row slots in shared memory
// (x | dx_out | dn of one row, 10 d bytes with a bf16 dn); lane 0 refills a slot with three `cp.async.bulk` copies that
// complete on the slot's mbarrier as soon as the warp has read it, so stages - 1 rows per warp are always in flight without
// costing a register: one CTA of 8 warps per SM, 6 stages at d = 384 = 150 KB of loads in flight per SM.  Row statistics
// are fetched 32 rows at a time, one batch ahead (lane l holds the row of iteration 32 b + l).  Wide rows (NV > 3) read
// their slot twice instead of keeping the row in registers.
template <typename NT, typename DT, int NV>
__global__ void __launch_bounds__(256, 1)
addln_bwd_bulk_kernel(const NT* __restrict__ dn, const float* __restrict__ dx_out, const float* __restrict__ x, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, int64_t T, int d, int stages, float* __restrict__ dx_in,
                      DT* __restrict__ d_delta, float* __restrict__ part /* [grid][2][d] */) {
    extern __shared__ __align__(128) uint8_t ln_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = d / 4;
    const uint32_t xb = static_cast<uint32_t>(d) * 4u, gb = static_cast<uint32_t>(d) * static_cast<uint32_t>(sizeof(NT));
    const uint32_t slot_bytes = 2 * xb + gb;   // x | dx_out | dn
    uint8_t* const ring = ln_smem + static_cast<size_t>(warp) * stages * slot_bytes;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(ln_smem + static_cast<size_t>(8) * stages * slot_bytes) + warp * stages;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int64_t stride = static_cast<int64_t>(gridDim.x) * 8;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    const int nrows = row0 < T ? static_cast<int>((T - 1 - row0) / stride) + 1 : 0;
    const uint32_t tx = xb + gb + (dx_out != nullptr ? xb : 0u);
    auto issue = [&](int it, int s) {   // lane 0 only
        const int64_t row = row0 + it * stride;
        const uint32_t dst = smem_u32(ring + static_cast<size_t>(s) * slot_bytes);
        mbar_arrive_expect_tx(bars + s, tx);
        bulk_load_1d(dst, x + row * d, xb, bars + s);
        if (dx_out != nullptr) bulk_load_1d(dst + xb, dx_out + row * d, xb, bars + s);
        bulk_load_1d(dst + 2 * xb, dn + row * d, gb, bars + s);
    };
    if (lane == 0)
        for (int it = 0; it < stages && it < nrows; ++it) issue(it, it);

    float4 gam[NV], dg[NV], db[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        gam[i] = (i * 32 + lane < nv) ? __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_d = 1.0f / static_cast<float>(d);
    auto ld_stats = [&](int it0, float& mu, float& rs) {
        const int it = it0 + lane;
        mu = 0.0f; rs = 0.0f;
        if (it < nrows) {
            const int64_t r = row0 + it * stride;
            mu = __ldg(mean + r);
            rs = __ldg(rstd + r);
        }
    };
    float mu_c = 0.0f, rs_c = 0.0f, mu_n, rs_n;
    ld_stats(0, mu_n, rs_n);
    constexpr bool kKeep = NV <= 3;   // the row stays in registers between the two passes
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nrows; ++it) {
        if ((it & 31) == 0) {
            mu_c = mu_n; rs_c = rs_n;
            ld_stats(it + 32, mu_n, rs_n);
        }
        const float mu = __shfl_sync(0xffffffffu, mu_c, it & 31), rs = __shfl_sync(0xffffffffu, rs_c, it & 31);
        const int64_t row = row0 + it * stride;
        mbar_wait(bars + s, ph);
        const uint8_t* const slot = ring + static_cast<size_t>(s) * slot_bytes;
        const float* const sx = reinterpret_cast<const float*>(slot);
        const float* const sr = reinterpret_cast<const float*>(slot + xb);
        const NT* const sg = reinterpret_cast<const NT*>(slot + 2 * xb);
        float c1 = 0.0f, c2 = 0.0f;
        [[maybe_unused]] float4 kxh[kKeep ? NV : 1], kg[kKeep ? NV : 1];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i * 32 + lane < nv) {
                const int c = (i * 32 + lane) * 4;
                const float4 xv = ld_f4(sx + c), gv = ld_f4(sg + c);
                const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                dg[i].x += gv.x * xh.x; dg[i].y += gv.y * xh.y; dg[i].z += gv.z * xh.z; dg[i].w += gv.w * xh.w;
                db[i].x += gv.x; db[i].y += gv.y; db[i].z += gv.z; db[i].w += gv.w;
                const float4 g = make_float4(gv.x * gam[i].x, gv.y * gam[i].y, gv.z * gam[i].z, gv.w * gam[i].w);
                c1 += (g.x + g.y) + (g.z + g.w);
                c2 += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
                if constexpr (kKeep) { kxh[i] = xh; kg[i] = g; }
            }
        }
        c1 = warp_sum(c1) * inv_d;
        c2 = warp_sum(c2) * inv_d;
        float4 o[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i * 32 + lane < nv) {
                const int c = (i * 32 + lane) * 4;
                float4 xh, g;
                if constexpr (kKeep) {
                    xh = kxh[i]; g = kg[i];
                } else {
                    const float4 xv = ld_f4(sx + c), gv = ld_f4(sg + c);
                    xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                    g = make_float4(gv.x * gam[i].x, gv.y * gam[i].y, gv.z * gam[i].z, gv.w * gam[i].w);
                }
                const float4 r = dx_out != nullptr ? ld_f4(sr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                o[i].x = rs * (g.x - c1 - xh.x * c2); o[i].y = rs * (g.y - c1 - xh.y * c2);
                o[i].z = rs * (g.z - c1 - xh.z * c2); o[i].w = rs * (g.w - c1 - xh.w * c2);
                o[i].x += r.x; o[i].y += r.y; o[i].z += r.z; o[i].w += r.w;
            }
        }
        // every lane has read the slot: refill it (generic-proxy reads ordered before the async-proxy write)
        __syncwarp();
        if (lane == 0 && it + stages < nrows) {
            fence_proxy_async_smem();
            issue(it + stages, s);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i * 32 + lane < nv) {
                const int c = (i * 32 + lane) * 4;
                st_f4(dx_in + row * d + c, o[i]);
                if (d_delta != nullptr) st_f4(d_delta + row * d + c, o[i]);
            }
        }
        if (++s == stages) { s = 0; ph ^= 1; }
    }
    // every copy a warp issued has been waited for: the ring is free, the warps are combined in warp order through it
    __syncthreads();
    float* const red = reinterpret_cast<float*>(ln_smem);   // [8][2*d]
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (i * 32 + lane < nv) {
            st_f4(red + warp * 2 * d + c, dg[i]);
            st_f4(red + warp * 2 * d + d + c, db[i]);
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * d; o += 256) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w * 2 * d + o];
        part[static_cast<size_t>(blockIdx.x) * 2 * d + o] = sum;
    }
}

// out[o] = sum_b part[b][o]: one warp per output, lanes stride the partials (independent loads), xor butterfly
__global__ void __launch_bounds__(256)
addln_bwd_reduce_kernel(const float* __restrict__ part, int nparts, int d, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= 2 * d) return;
    float v = 0.0f;
#pragma unroll 4
    for (int b = lane; b < nparts; b += 32) v += part[static_cast<size_t>(b) * 2 * d + o];
    v = warp_sum(v);
    if (lane == 0) {
        if (o < d) dgamma[o] = v;
        else dbeta[o - d] = v;
    }
}

// persistent grid sized to what is resident at once (3 / 2 / 1 CTAs per SM for NV <= 3 / <= 6 / 8), so no partial last wave
// ------------------------------------------------------------------------------------------------
// Plain column sum out[c] = sum_r buf[r, c] of a [rows, cols] bf16 / fp32 matrix (bias gradient of the dense
// projections of the hosting block), same two deterministic stages as segment_colsum: 128 rows x 256 columns per
// CTA with every load in flight, warps combined in order, then one warp per output column over the row blocks.
// ------------------------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const XT* __restrict__ buf, int64_t rows, int cols, float* __restrict__ part) {
    __shared__ float red[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = static_cast<int64_t>(blockIdx.y) * 128 + warp * 16;
    const int c = blockIdx.x * 256 + lane * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (c < cols) {
        float4 lo[16], hi[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int64_t rr = row0 + r < rows ? row0 + r : rows - 1;   // clamped load, masked below
            lo[r] = ld_f4(buf + rr * cols + c);
            hi[r] = ld_f4(buf + rr * cols + c + 4);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float m = row0 + r < rows ? 1.0f : 0.0f;
            acc[0] = fmaf(m, lo[r].x, acc[0]); acc[1] = fmaf(m, lo[r].y, acc[1]); acc[2] = fmaf(m, lo[r].z, acc[2]); acc[3] = fmaf(m, lo[r].w, acc[3]);
            acc[4] = fmaf(m, hi[r].x, acc[4]); acc[5] = fmaf(m, hi[r].y, acc[5]); acc[6] = fmaf(m, hi[r].z, acc[6]); acc[7] = fmaf(m, hi[r].w, acc[7]);
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
        part[static_cast<size_t>(blockIdx.y) * cols + cc] = sum;
    }
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int nparts, int cols, float* __restrict__ out) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= cols) return;
    float v = 0.0f;
#pragma unroll 4
    for (int b = lane; b < nparts; b += 32) v += part[static_cast<size_t>(b) * cols + o];
    v = warp_sum(v);
    if (lane == 0) out[o] = v;
}

size_t colsum_workspace_bytes(int64_t rows, int cols) { return static_cast<size_t>((rows + 127) / 128) * cols * 4; }

cudaError_t launch_colsum(const void* buf, int dtype, int64_t rows, int cols, void* workspace, float* out, cudaStream_t st) {
    const int nb = static_cast<int>((rows + 127) / 128);
    dim3 g1((cols + 255) / 256, nb);
    float* part = static_cast<float*>(workspace);
    if (dtype == MOE_DTYPE_BF16) colsum_partial_kernel<__nv_bfloat16><<<g1, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(buf), rows, cols, part);
    else colsum_partial_kernel<float><<<g1, 256, 0, st>>>(static_cast<const float*>(buf), rows, cols, part);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    colsum_final_kernel<<<(cols + 7) / 8, 256, 0, st>>>(part, nb, cols, out);
    return cudaGetLastError();
}

// The staged kernel (addln_bwd_bulk_kernel) runs one 8-warp CTA per SM with as many row slots per warp as fit
// kLnRingBudget; rows it cannot take (d % 8 != 0, fewer than two stages) go to the register kernel.
constexpr int kLnRingBudget = 200 * 1024;
struct LnBwdPlan { int bulk, stages, grid; size_t smem; };
static LnBwdPlan addln_bwd_plan(int64_t T, int d, int n_dtype) {
    LnBwdPlan p{0, 0, 0, 0};
    const int64_t want = (T + 7) / 8;
    const size_t slot = static_cast<size_t>(d) * (8 + (n_dtype == MOE_DTYPE_BF16 ? 2 : 4));
    // narrow rows are latency-bound with 8 warps per SM (two warp reductions per 2 KB row): two CTAs per SM, half the ring each
    // (measured at T = 50 432, profiles/r02_addln_bwd_staged.md: d = 192 42.5 -> 31.5 us with two CTAs, d = 384 52.7 -> 56.9 us)
    int ctas = d <= 256 ? 2 : 1, budget = kLnRingBudget, stages_max = 8, mode = 1;
#ifdef MOE_EXPERIMENT_HOOKS   // tools/build_variant.sh only: the shipped library never reads the environment
    if (getenv("MOE_LN_BWD_MODE")) mode = atoi(getenv("MOE_LN_BWD_MODE"));
    if (getenv("MOE_LN_CTAS")) ctas = atoi(getenv("MOE_LN_CTAS"));
    if (getenv("MOE_LN_BUDGET_KB")) budget = atoi(getenv("MOE_LN_BUDGET_KB")) * 1024;
    if (getenv("MOE_LN_STAGES")) stages_max = atoi(getenv("MOE_LN_STAGES"));
#endif
    int stages = static_cast<int>(budget / ctas / (8 * slot));
    if (stages > stages_max) stages = stages_max;
    if (mode == 1 && d % 8 == 0 && stages >= 2) {
        p.bulk = 1;
        p.stages = stages;
        const int64_t cap = static_cast<int64_t>(sm_count()) * ctas;
        p.grid = static_cast<int>(want < cap ? want : cap);
        p.smem = static_cast<size_t>(8) * stages * slot + static_cast<size_t>(8) * stages * 8;
        return p;
    }
    const int per_sm = d <= 768 ? 2 : 1;   // matches the register kernel's __launch_bounds__ residency
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    p.grid = static_cast<int>(want < cap ? want : cap);
    p.smem = static_cast<size_t>(8) * 2 * d * 4;
    return p;
}
// partials per CTA: sized for the larger of the two grids so the workspace does not depend on the operand dtype
static int addln_bwd_blocks(int64_t T, int d) {
    const int a = addln_bwd_plan(T, d, MOE_DTYPE_BF16).grid, b = addln_bwd_plan(T, d, MOE_DTYPE_F32).grid;
    return a > b ? a : b;
}

size_t addln_bwd_workspace_bytes(int64_t T, int d) { return static_cast<size_t>(addln_bwd_blocks(T, d)) * 2 * d * 4; }

cudaError_t launch_addln_fwd(const float* x_in, const void* delta, int delta_dtype, const float* gamma, const float* beta, float eps,
                             int64_t T, int d, float* x_out, void* n, int n_dtype, float* mean, float* rstd, cudaStream_t st) {
    const int grid = static_cast<int>((T + 7) / 8);
#define MOE_LN_FWD_NV(DT, NT, NVV)                                                                                      \
    addln_fwd_kernel<DT, NT, NVV><<<grid, 256, 0, st>>>(x_in, static_cast<const DT*>(delta), gamma, beta, eps, T, d, x_out, \
                                                        static_cast<NT*>(n), mean, rstd)
#define MOE_LN_FWD(DT, NT)                                                                                              \
    {                                                                                                                   \
        if (d <= 256) MOE_LN_FWD_NV(DT, NT, 2);                                                                         \
        else if (d <= 384) MOE_LN_FWD_NV(DT, NT, 3);                                                                    \
        else if (d <= 768) MOE_LN_FWD_NV(DT, NT, 6);                                                                    \
        else MOE_LN_FWD_NV(DT, NT, 8);                                                                                  \
    }
    const bool dbf = delta != nullptr && delta_dtype == MOE_DTYPE_BF16, nbf = n_dtype == MOE_DTYPE_BF16;
    if (dbf && nbf) MOE_LN_FWD(__nv_bfloat16, __nv_bfloat16)
    else if (dbf) MOE_LN_FWD(__nv_bfloat16, float)
    else if (nbf) MOE_LN_FWD(float, __nv_bfloat16)
    else MOE_LN_FWD(float, float)
#undef MOE_LN_FWD
#undef MOE_LN_FWD_NV
    return cudaGetLastError();
}

cudaError_t launch_addln_bwd(const void* dn, int n_dtype, const float* dx_out, const float* x, const float* mean, const float* rstd,
                             const float* gamma, int64_t T, int d, float* dx_in, void* d_delta, int delta_dtype, void* workspace,
                             float* dgamma, float* dbeta, cudaStream_t st) {
    const bool nbf = n_dtype == MOE_DTYPE_BF16, dbf = d_delta != nullptr && delta_dtype == MOE_DTYPE_BF16;
    LnBwdPlan plan = addln_bwd_plan(T, d, n_dtype);
    const auto misaligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
    if (plan.bulk && (misaligned(dn) || misaligned(x) || misaligned(dx_out))) {   // cp.async.bulk needs 16-byte aligned rows
        plan.bulk = 0;
        plan.smem = static_cast<size_t>(8) * 2 * d * 4;
    }
    const int grid = plan.grid;
    float* part = static_cast<float*>(workspace);
    const size_t smem = plan.smem;
    cudaError_t err = cudaSuccess;
#define MOE_LN_BWD_NV(NT, DT, NVV)                                                                                          \
    if (plan.bulk) {                                                                                                        \
        auto kfn = addln_bwd_bulk_kernel<NT, DT, NVV>;                                                                      \
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                            \
        if (err != cudaSuccess) return err;                                                                                 \
        kfn<<<grid, 256, smem, st>>>(static_cast<const NT*>(dn), dx_out, x, mean, rstd, gamma, T, d, plan.stages, dx_in,   \
                                     static_cast<DT*>(d_delta), part);                                                      \
    } else {                                                                                                                \
        auto kfn = addln_bwd_kernel<NT, DT, NVV>;                                                                           \
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                            \
        if (err != cudaSuccess) return err;                                                                                 \
        kfn<<<grid, 256, smem, st>>>(static_cast<const NT*>(dn), dx_out, x, mean, rstd, gamma, T, d, dx_in,                \
                                     static_cast<DT*>(d_delta), part);                                                      \
    }
#define MOE_LN_BWD(NT, DT)                                                                                                  \
    {                                                                                                                       \
        if (d <= 256) MOE_LN_BWD_NV(NT, DT, 2)                                                                              \
        else if (d <= 384) MOE_LN_BWD_NV(NT, DT, 3)                                                                         \
        else if (d <= 768) MOE_LN_BWD_NV(NT, DT, 6)                                                                         \
        else MOE_LN_BWD_NV(NT, DT, 8)                                                                                       \
    }
    if (nbf && dbf) MOE_LN_BWD(__nv_bfloat16, __nv_bfloat16)
    else if (nbf) MOE_LN_BWD(__nv_bfloat16, float)
    else if (dbf) MOE_LN_BWD(float, __nv_bfloat16)
    else MOE_LN_BWD(float, float)
#undef MOE_LN_BWD
#undef MOE_LN_BWD_NV
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    addln_bwd_reduce_kernel<<<(2 * d + 7) / 8, 256, 0, st>>>(part, grid, d, dgamma, dbeta);
    return cudaGetLastError();
}

}  // namespace moe
