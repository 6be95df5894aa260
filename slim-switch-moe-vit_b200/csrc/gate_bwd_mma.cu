// gate_bwd_mma.cu — gate backward + dispatch backward in one pass, the gate term on the tensor cores (sm_100a).
//
//   dlogits[t, :]  from dscore (+ dpsum of the load-balancing loss)      -> written out for the gate weight gradient
//   dx[t]        = sum_j dXbuf[pos[t, j]]  +  dlogits[t, :] Wg            (autograd of MOEScatter + NaiveGate's Linear;
//                                                                          FastMoE, reached from /root/reference/models/resMoE.py:27-29)
//
// Round 1 computed dlogits Wg with T d E fp32 FMAs on the CUDA cores: 48 us at the config-2 layer shape and 4x that
// at E = 64.  Here the product is mma.sync m16n8k8 (tf32 operands, fp32 accumulate; no bit-exactness constraint on
// this path — the tolerance is the bf16 one of the surrounding tensors), so the kernel is a gather-add stream
// whatever E is:
//   * one CTA per 64-token tile; its dlogits rows are computed by 4 threads per token into shared memory (A operand);
//   * the packed dXbuf rows of a sub-tile (16-64 tokens) are requested with 16-byte cp.async up front (possibly from a
//     peer GPU under expert parallelism: PeerRows — then through L1, so that the requests cross NVLink as whole 128-byte
//     lines), rows padded by 16 bytes so that fragment-shaped accesses are conflict-free;
//   * warp w owns d / 8 output columns: Wg fragments (B operand, tf32) are loaded once per 32-column chunk and reused
//     by every 16-token m-tile; the accumulator fragment is added to the staged rows IN PLACE (bf16 output) and the
//     finished rows leave with coalesced 16-byte stores — or straight from the fragments for fp32 output (a quad
//     covers one 32-byte sector).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace moe {

namespace {

constexpr int kTile = MOE_TOKEN_TILE;   // 64 tokens per CTA
constexpr int kMaxPick = 8;

__device__ __forceinline__ uint32_t f2tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void stsm_x2(uint32_t saddr, uint32_t r0, uint32_t r1) {   // row addresses from lanes 0-15
    asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1, %2};" ::"r"(saddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void stsm_x4(uint32_t saddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// NKB: k-steps of 8 experts (E rounded up to 8 NKB <= 64).  KT: compile-time top-k (0 = runtime k <= 8).
// The gather-add itself also runs on the tensor cores: the staged rows of a 16-token m-tile are the B operand of an
// m16n8k16 bf16 MMA whose A operand is the 16 x 16 identity (ldmatrix.trans turns the token-major rows into the k-major
// fragment), so  acc = dlogits Wg (tf32)  +  I rows_slot0 (+ I rows_slot1 ...)  exactly, in fp32, with two shared-memory
// instructions per 16 x 16 block instead of a load / unpack / add per element pair.
template <typename OT, int NKB, int KT>
__global__ void __launch_bounds__(256, (NKB <= 2 ? 4 : NKB == 4 ? 3 : 2))   // resident CTAs hide each other's phase latencies
gate_dispatch_bwd_mma_kernel(PeerRows dxrows, const int* __restrict__ pos, const float* __restrict__ logits,
                             const int* __restrict__ idx, const float* __restrict__ score, const float* __restrict__ dscore,
                             const float* __restrict__ dpsum, const float* __restrict__ Wg, int64_t T, int d, int E, int k_rt,
                             int score_mode, float* __restrict__ dlogits, OT* __restrict__ dx, int ts) {
    constexpr int NC = 4;
    constexpr int DLS = 8 * NKB + 4;           // floats per staged dlogits row (+4: conflict-free A-fragment reads)
    constexpr int KP = KT > 0 ? KT : kMaxPick;
    constexpr int EPT = 2 * NKB;               // experts per thread in the dlogits phase (4 threads per token)
    const int k = KT > 0 ? KT : k_rt;
    extern __shared__ __align__(16) uint8_t smem_gb[];
    float* dl_s = reinterpret_cast<float*>(smem_gb);                      // [kTile][DLS]
    int* pos_s = reinterpret_cast<int*>(dl_s + kTile * DLS);              // [kTile * k]
    const int RS = d * 2 + 16;                                            // bytes per staged row (bank-skewed)
    uint8_t* rows_s = reinterpret_cast<uint8_t*>(pos_s + ((kTile * k + 3) & ~3));   // [ts * k][RS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t t_base = static_cast<int64_t>(blockIdx.x) * kTile;
    const int n_tok = static_cast<int>(min(static_cast<int64_t>(kTile), T - t_base));
    const bool dense = (score_mode == 1) || (dpsum != nullptr);
    const bool have_rows = dxrows.base[0] != nullptr;
    const bool remote = dxrows.n > 1;
    const uint32_t rows_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(rows_s));
    const int c16 = d >> 3;   // 16-byte chunks per row

    for (int i = tid; i < kTile * k; i += 256) pos_s[i] = i < n_tok * k ? pos[t_base * k + i] : -1;
    __syncthreads();
    // staged gather of one sub-tile: warp w requests the packed rows w, w + 8, ... with 16-byte cp.async, all up front
    auto stage = [&](int tl_begin) {
        for (int pr = warp; pr < ts * k; pr += 8) {
            const int row = have_rows ? pos_s[tl_begin * k + pr] : -1;
            const uint32_t dst = rows_u32 + pr * RS;
            if (row >= 0) {
                const __nv_bfloat16* src = peer_row<__nv_bfloat16>(dxrows, row, d);
                for (int c = lane; c < c16; c += 32)
                    // Rows on a peer GPU go through L1 (.ca): the LSU then sends a warp's 32 x 16 bytes over NVLink as whole
                    // 128-byte requests.  L2-only copies (.cg) measured 84 us against 57 us for this kernel with half the rows
                    // remote (profiles/r02_cpasync_peer_ab.log); local rows keep .cg (nothing is re-read, no L1 allocation).
                    if (remote)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c * 8) : "memory");
                    else
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c * 8) : "memory");
            } else {
                for (int c = lane; c < c16; c += 32)
                    *reinterpret_cast<uint4*>(rows_s + pr * RS + c * 16) = make_uint4(0u, 0u, 0u, 0u);   // dropped / skipped / past T
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);

    {   // dlogits rows of the tile: 4 threads per token, experts strided over the 4 (same formulas as gate_bwd_kernel)
        const int tl = tid >> 2, part = tid & 3;
        const bool live = tl < n_tok;
        const int64_t t = t_base + (live ? tl : 0);
        const float* lr = logits + t * E;
        int pk[KP];
        float sc[KP], gr[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) {   // fixed trip count + predicate: the arrays stay in registers
            const bool on = j < k;
            pk[j] = on ? idx[t * k + j] : -2;
            sc[j] = on ? score[t * k + j] : 0.0f;
            gr[j] = on ? dscore[t * k + j] : 0.0f;
        }
        const bool masked = pk[0] < 0;   // token-skip mask: no gate gradient at all
        float pe[EPT];                   // softmax probabilities of this thread's experts (one expf each)
        float pdot = 0.0f;
        if (dense) {
            const float m = lr[max(pk[0], 0)];
            float z = 0.0f;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int e = part + 4 * i;
                pe[i] = e < E ? expf(lr[e] - m) : 0.0f;
                z += pe[i];
            }
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            const float rz = 1.0f / z;
#pragma unroll
            for (int i = 0; i < EPT; ++i) pe[i] *= rz;
            if (dpsum != nullptr) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int e = part + 4 * i;
                    if (e < E) pdot += pe[i] * dpsum[e];
                }
                pdot += __shfl_xor_sync(0xffffffffu, pdot, 1);
                pdot += __shfl_xor_sync(0xffffffffu, pdot, 2);
            }
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) pe[i] = 0.0f;
        }
        float inner = 0.0f;
#pragma unroll
        for (int j = 0; j < KP; ++j) inner += sc[j] * gr[j];
        float* dl = dl_s + tl * DLS;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int e = part + 4 * i;
            float v = 0.0f;
            if (live && !masked && e < E) {
                if (score_mode == 0) {
#pragma unroll
                    for (int j = 0; j < KP; ++j)
                        if (pk[j] == e) v += sc[j] * (gr[j] - inner);
                } else {
#pragma unroll
                    for (int j = 0; j < KP; ++j)
                        if (pk[j] == e) v += gr[j] * sc[j];
                    v -= inner * pe[i];
                }
                if (dpsum != nullptr) v += pe[i] * (dpsum[e] - pdot);
            }
            dl[e] = v;   // rows past the end of the batch and expert columns past E are zero: they feed the MMA
            if (live && e < E) dlogits[t * E + e] = v;   // the quad of a token covers 4 consecutive experts: 16-byte segments
        }
    }
    __syncthreads();

    const int CW = d >> 3;          // output columns of this warp
    const int NTW = CW >> 3;        // its 8-column n-tiles
    // identity A fragment (bf16 1.0 = 0x3F80): row g has its one at k = g
    const uint32_t ident = (g == 2 * t4 ? 0x00003F80u : 0u) | (g == 2 * t4 + 1 ? 0x3F800000u : 0u);
    // ldmatrix / stmatrix lane address inside a 16-token x 16-column block: matrices (tok 0-7, col 0-7), (tok 8-15, col 0-7),
    // (tok 0-7, col 8-15), (tok 8-15, col 8-15)
    const int lm_tok = ((lane >> 3) & 1) * 8 + (lane & 7);
    const int lm_col = (lane >> 4) * 8;
    for (int sub = 0; sub < n_tok; sub += ts) {
        if (sub > 0) {
            __syncthreads();        // the previous sub-tile's rows have left
            stage(sub);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int n_mt = (min(ts, n_tok - sub) + 15) >> 4;
        for (int c0 = 0; c0 < NTW; c0 += NC) {
            const int nc = min(NC, NTW - c0);
            uint32_t bf[NKB][NC][2];
#pragma unroll
            for (int ks = 0; ks < NKB; ++ks) {
#pragma unroll
                for (int n = 0; n < NC; ++n) {
                    // expert rows past E are clamped (their dlogits columns are zero), n-tiles past the warp's range re-read
                    // its last one (their MMAs are skipped): no predicates on the loads
                    const int col = warp * CW + (c0 + min(n, nc - 1)) * 8 + g;
                    const int e0 = min(ks * 8 + t4, E - 1), e1 = min(ks * 8 + t4 + 4, E - 1);
                    bf[ks][n][0] = f2tf32(__ldg(Wg + static_cast<size_t>(e0) * d + col));
                    bf[ks][n][1] = f2tf32(__ldg(Wg + static_cast<size_t>(e1) * d + col));
                }
            }
            for (int mt = 0; mt < n_mt; ++mt) {
                float acc[NC][4];
#pragma unroll
                for (int n = 0; n < NC; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
                const float* a0p = dl_s + (sub + mt * 16 + g) * DLS + t4;
#pragma unroll
                for (int ks = 0; ks < NKB; ++ks) {
                    uint32_t a[4];
                    // raw fp32 bits: the tensor core reads the tf32 part (truncation, 2^-10 relative — the tolerance of this
                    // term is the bf16 one; cvt.rna costs three instructions per element here, once per m-tile and chunk)
                    a[0] = __float_as_uint(a0p[ks * 8]);
                    a[1] = __float_as_uint(a0p[ks * 8 + 8 * DLS]);
                    a[2] = __float_as_uint(a0p[ks * 8 + 4]);
                    a[3] = __float_as_uint(a0p[ks * 8 + 8 * DLS + 4]);
#pragma unroll
                    for (int n = 0; n < NC; ++n)
                        if (n < nc) mma_tf32(acc[n], a, bf[ks][n][0], bf[ks][n][1]);
                }
                // + the staged rows (identity MMA), two n-tiles per ldmatrix
                const uint32_t blk = rows_u32 + static_cast<uint32_t>((mt * 16 + lm_tok) * k) * RS + (warp * CW + c0 * 8 + lm_col) * 2;
#pragma unroll
                for (int np = 0; np < NC / 2; ++np) {
                    if (2 * np < nc) {
                        const bool pair = 2 * np + 1 < nc;   // an odd tail (d / 64 odd) reads its neighbour's 8 columns (or the row pad) and drops them
                        for (int j = 0; j < k; ++j) {
                            uint32_t b[4];
                            ldsm_x4_trans(blk + j * RS + np * 32, b);
                            mma_bf16(acc[2 * np], ident, 0u, 0u, ident, b[0], b[1]);
                            if (pair) mma_bf16(acc[2 * np + 1], ident, 0u, 0u, ident, b[2], b[3]);
                        }
                        if constexpr (sizeof(OT) == 2) {   // slot j = 0 of the block now holds dx (bf16), row-major again
                            if (pair)
                                stsm_x4(blk + np * 32, pack_bf162(acc[2 * np][0], acc[2 * np][1]), pack_bf162(acc[2 * np][2], acc[2 * np][3]),
                                        pack_bf162(acc[2 * np + 1][0], acc[2 * np + 1][1]), pack_bf162(acc[2 * np + 1][2], acc[2 * np + 1][3]));
                            else
                                stsm_x2(blk + np * 32, pack_bf162(acc[2 * np][0], acc[2 * np][1]), pack_bf162(acc[2 * np][2], acc[2 * np][3]));
                        }
                    }
                }
                if constexpr (sizeof(OT) == 4) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int tls = mt * 16 + g + rr * 8;
                        if (sub + tls < n_tok) {
#pragma unroll
                            for (int n = 0; n < NC; ++n)
                                if (n < nc)
                                    *reinterpret_cast<float2*>(reinterpret_cast<float*>(dx) + (t_base + sub + tls) * d + warp * CW + (c0 + n) * 8 + 2 * t4) =
                                        make_float2(acc[n][rr * 2], acc[n][rr * 2 + 1]);
                        }
                    }
                }
            }
        }
        if constexpr (sizeof(OT) == 2) {
            __syncthreads();
            const int n_sub = min(ts, n_tok - sub);
            for (int r = warp; r < n_sub; r += 8) {   // warp per row, 16 bytes per lane
                const uint8_t* src = rows_s + static_cast<size_t>(r) * k * RS;
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(dx) + (t_base + sub + r) * d;
                for (int c = lane; c < c16; c += 32) *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(src + c * 16);
            }
        }
    }
}

template <typename OT, int NKB, int KT>
cudaError_t launch_t(const PeerRows& rows, const int* pos, const float* logits, const int* idx, const float* score,
                     const float* dscore, const float* dpsum, const float* Wg, int64_t T, int d, int E, int k, int score_mode,
                     float* dlogits, void* dx, int ts, size_t smem, cudaStream_t st) {
    auto kfn = gate_dispatch_bwd_mma_kernel<OT, NKB, KT>;
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (err != cudaSuccess) return err;
    const int ntiles = static_cast<int>((T + kTile - 1) / kTile);
    kfn<<<ntiles, 256, smem, st>>>(rows, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits,
                                   static_cast<OT*>(dx), ts);
    return cudaGetLastError();
}

template <typename OT, int NKB>
cudaError_t launch_k(const PeerRows& rows, const int* pos, const float* logits, const int* idx, const float* score,
                     const float* dscore, const float* dpsum, const float* Wg, int64_t T, int d, int E, int k, int score_mode,
                     float* dlogits, void* dx, int ts, size_t smem, cudaStream_t st) {
    if (k == 1) return launch_t<OT, NKB, 1>(rows, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx, ts, smem, st);
    if (k == 2) return launch_t<OT, NKB, 2>(rows, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx, ts, smem, st);
    return launch_t<OT, NKB, 0>(rows, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx, ts, smem, st);
}

}  // namespace

// sub-tile (tokens per staged gather) of the tensor-core kernel, 0 = shape not supported (the CUDA-core kernel runs)
static size_t gdb_fixed_smem(int nkb, int k) { return static_cast<size_t>(kTile) * (8 * nkb + 4) * 4 + ((kTile * k + 3) & ~3) * 4; }

static int gdb_mma_subtile(int d, int E, int k) {
    if (E > 64 || d % 64 != 0 || k > kMaxPick) return 0;
    const size_t rs = static_cast<size_t>(d) * 2 + 16;
    const int nkb = E <= 8 ? 1 : E <= 16 ? 2 : E <= 32 ? 4 : 8;
    const int ctas = nkb <= 2 ? 4 : nkb == 4 ? 3 : 2;       // the kernel's launch bounds
    const size_t budget = (226 * 1024) / ctas - 1024 - gdb_fixed_smem(nkb, k);
    int ts = kTile;
    while (ts > 16 && static_cast<size_t>(ts) * k * rs > budget) ts >>= 1;
    if (static_cast<size_t>(ts) * k * rs <= budget) return ts;
    return gdb_fixed_smem(nkb, k) + static_cast<size_t>(ts) * k * rs <= 100 * 1024 ? ts : 0;   // fewer resident CTAs, still one kernel
}

bool gate_dispatch_bwd_mma_supported(int d, int E, int k) { return gdb_mma_subtile(d, E, k) > 0; }

cudaError_t launch_gate_dispatch_bwd_mma(const PeerRows& rows, const int* pos, const float* logits, const int* idx,
                                         const float* score, const float* dscore, const float* dpsum, const float* Wg, int64_t T,
                                         int d, int E, int k, int score_mode, float* dlogits, void* dx, int dx_dtype,
                                         cudaStream_t st) {
    const int ts = gdb_mma_subtile(d, E, k);
    const int nkb = E <= 8 ? 1 : E <= 16 ? 2 : E <= 32 ? 4 : 8;
    const size_t smem = gdb_fixed_smem(nkb, k) + static_cast<size_t>(ts) * k * (d * 2 + 16);
#define MOE_GDB_LAUNCH(OT_, NKB_) \
    return launch_k<OT_, NKB_>(rows, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx, ts, smem, st)
    if (dx_dtype == MOE_DTYPE_F32) {
        if (nkb == 1) MOE_GDB_LAUNCH(float, 1);
        if (nkb == 2) MOE_GDB_LAUNCH(float, 2);
        if (nkb == 4) MOE_GDB_LAUNCH(float, 4);
        MOE_GDB_LAUNCH(float, 8);
    }
    if (nkb == 1) MOE_GDB_LAUNCH(__nv_bfloat16, 1);
    if (nkb == 2) MOE_GDB_LAUNCH(__nv_bfloat16, 2);
    if (nkb == 4) MOE_GDB_LAUNCH(__nv_bfloat16, 4);
    MOE_GDB_LAUNCH(__nv_bfloat16, 8);
#undef MOE_GDB_LAUNCH
}

}  // namespace moe
