// gemm_launch.cu — host side of the grouped tcgen05 GEMM: TMA tensor-map encoding and launch.
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "gemm.cuh"

namespace moe {

namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 2-D row-major tensor [outer, inner] of `esz`-byte elements; 128-byte swizzle unless told otherwise
bool encode_2d(CUtensorMap* m, CUtensorMapDataType dt, int esz, const void* base, uint64_t inner, uint64_t outer,
               uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    auto fn = get_encode_fn();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * static_cast<uint64_t>(esz)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d (inner=%llu outer=%llu box=%u x %u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer);
        return false;
    }
    return true;
}

// 3-D tensor [d2, d1, d0] (d0 innermost), fp32
bool encode_3d_f32(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                   CUtensorMapSwizzle swz) {
    auto fn = get_encode_fn();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r); return false; }
    return true;
}

}  // namespace

// shared with the other translation units (gate_mma.cu): [outer, inner] bf16 row-major, 128-byte swizzle
bool encode_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
    return encode_2d(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, inner, outer, box_inner, box_outer);
}

namespace {

template <int BN, int EPI, bool WGRAD>
int launch_one(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tO0, const CUtensorMap& tO1,
               const CUtensorMap& tAux, const GemmParams& p, int grid, cudaStream_t st) {
    using Cfg = GemmCfg<BN, EPI>;
    auto kfn = grouped_gemm_kernel<BN, EPI, WGRAD>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    // the kernel carries __cluster_dims__(2,1,1): the grid is a whole number of CTA pairs
    kfn<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tA, tB, tO0, tO1, tAux, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("grouped_gemm launch: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

// ROWS mode: N must be tiled exactly (the outputs are dense row buffers).  fc2 / dgrad (`wide_ok`) take 384-wide
// single-accumulator tiles when N allows it and K is long enough to amortise the un-overlapped epilogue.
int pick_bn_rows(int N, int K, bool wide_ok) {
#ifdef MOE_EXPERIMENT_HOOKS   // tools/build_variant.sh only: the shipped library never reads the environment
    if (wide_ok && getenv("MOE_ROWS_BN")) return atoi(getenv("MOE_ROWS_BN"));
#endif
    if (wide_ok && N % 384 == 0 && N % 256 != 0 && K >= 768) return 384;
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    if (N % 64 == 0) return 64;
    return 0;
}
// WGRAD mode: BN in {128, 192, 256}; ragged N is handled by TMA zero-fill / clipping.  Exact tilings first
// (N = 384 -> two 192-wide tiles, B read through 64-byte-swizzle atoms); otherwise 256 wins whenever N > 128 even
// when part of the last tile is padding (N = 384: measured 75 us vs 86 us with three 128-wide tiles): the A tile
// is re-read once per N tile, and operand feed is the limit.
int pick_bn_wgrad(int N) {
#ifdef MOE_EXPERIMENT_HOOKS
    if (getenv("MOE_WGRAD_BN")) return atoi(getenv("MOE_WGRAD_BN"));
#endif
    if (N % 384 == 0 && N % 256 != 0) return 384;   // one 384-column accumulator: A is read once per 384 columns
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    return N > 128 ? 256 : 128;
}

}  // namespace

int launch_grouped_gemm(int op, const void* A, const void* B, void* out0, void* out1, const float* bias, const void* aux,
                        const int* tile_expert, const int* num_mtiles, const int* seg_start, int64_t rows_cap, int E,
                        int M, int N, int K, int sm_count, cudaStream_t st) {
    const bool wgrad = op == MOE_GEMM_WGRAD || op == MOE_GEMM_WGRAD_T;
    if (op < MOE_GEMM_FC1 || op > MOE_GEMM_WGRAD_T) { set_error("grouped gemm: unknown op %d", op); return 1; }
    if (N <= 0 || N % 64 != 0) { set_error("grouped gemm: N=%d must be a positive multiple of 64", N); return 1; }
    if (!wgrad && (K % 64 != 0 || K <= 0)) { set_error("grouped gemm: K=%d must be a positive multiple of 64", K); return 1; }
    if (wgrad && (M % 64 != 0 || M <= 0)) { set_error("grouped gemm: M=%d must be a positive multiple of 64", M); return 1; }
    if (rows_cap % MOE_ROW_ALIGN != 0) { set_error("grouped gemm: rows_cap=%lld must be a multiple of %d", (long long)rows_cap, MOE_ROW_ALIGN); return 1; }
    if ((op == MOE_GEMM_FC1 || op == MOE_GEMM_FC2) && bias == nullptr) { set_error("grouped gemm: fc1 / fc2 need a bias vector"); return 1; }
    if (op == MOE_GEMM_DGELU && aux == nullptr) { set_error("grouped gemm: dgelu needs the pre-activation (aux)"); return 1; }
    const bool wide_ok = op == MOE_GEMM_FC2 || op == MOE_GEMM_DGRAD;
    const bool infer_fc1 = op == MOE_GEMM_FC1 && out0 == nullptr && out1 != nullptr;
    const int bn = wgrad ? pick_bn_wgrad(N) : pick_bn_rows(N, K, wide_ok);
    if (!wgrad && (bn <= 0 || N % bn != 0 || (bn == 384 && !wide_ok))) { set_error("grouped gemm: BN=%d does not tile N=%d for op %d", bn, N, op); return 1; }

    GemmParams p{};
    p.tile_expert = tile_expert;
    p.num_mtiles = num_mtiles;
    p.seg_start = seg_start;
    p.bias = bias;
    p.aux = wgrad ? nullptr : static_cast<const __nv_bfloat16*>(aux);
    // WGRAD: `aux` is the optional stream-K flag workspace (moe_wgrad_flags_bytes(E, M, N) bytes, zero-filled once; the
    // kernel leaves it zero)
    p.colsum = op == MOE_GEMM_DGELU ? static_cast<float*>(out1) : nullptr;   // DGELU: out1 = optional slab column sums
    p.flags = wgrad ? static_cast<int*>(const_cast<void*>(aux)) : nullptr;
    p.streamk = 0;
    p.E = E; p.M = M; p.N = N; p.K = K;

    CUtensorMap tA, tB, tO0, tO1, tAux;
    const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const uint64_t R = static_cast<uint64_t>(rows_cap);
    bool ok = true;
    if (!wgrad) {
        // A [rows, K] K-major: each CTA of a pair loads its 128 A rows per k-block.  B: the forward ops read the weights
        // [E*N, K] K-major (bn/2 rows per CTA and k-block); dgelu / dgrad read the SAME weights as [E*K, N] MN-major, in
        // 64 x 64 boxes (32-column boxes with 64-byte swizzle when the CTA's half of the tile is not whole 64-column atoms)
        ok = ok && encode_2d(&tA, BF, 2, A, K, R, 64, 128);
        const bool b_mn = op == MOE_GEMM_DGELU || op == MOE_GEMM_DGRAD;
        if (!b_mn) ok = ok && encode_2d(&tB, BF, 2, B, K, static_cast<uint64_t>(E) * N, 64, bn <= 256 ? bn / 2 : 64);
        else if ((bn / 2) % 64 == 0) ok = ok && encode_2d(&tB, BF, 2, B, N, static_cast<uint64_t>(E) * K, 64, 64);
        else ok = ok && encode_2d(&tB, BF, 2, B, N, static_cast<uint64_t>(E) * K, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
        // each epilogue warp stores (and, for dgelu, loads its rows of the pre-activation as) 32-row x 32-column slabs
        const auto S64 = CU_TENSOR_MAP_SWIZZLE_64B;
        // fc1 without out0 (G = gelu'): forward-only pass, the single output H goes through the first descriptor
        if (infer_fc1) out0 = out1;
        ok = ok && encode_2d(&tO0, BF, 2, out0, N, R, 32, 32, S64);
        ok = ok && encode_2d(&tO1, BF, 2, op == MOE_GEMM_FC1 ? out1 : out0, N, R, 32, 32, S64);
        ok = ok && encode_2d(&tAux, BF, 2, op == MOE_GEMM_DGELU ? aux : out0, N, R, 32, 32, S64);
    } else {
        // A [rows, M], B [rows, N] read MN-major in 64 x 64 boxes (B: 32 x 64 boxes, 64-byte swizzle, when the
        // CTA's half of the tile is not whole 64-column atoms); out [E, M, N] (WGRAD_T: [E, N, M]) fp32, 32 x 32 boxes
        if (bn != 128 && bn != 192 && bn != 256 && bn != 384) { set_error("grouped gemm: wgrad BN=%d not in {128,192,256,384}", bn); return 1; }
        ok = ok && encode_2d(&tA, BF, 2, A, M, R, 64, 64);
        if ((bn / 2) % 64 == 0) ok = ok && encode_2d(&tB, BF, 2, B, N, R, 64, 64);
        else ok = ok && encode_2d(&tB, BF, 2, B, N, R, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
        // half-slab boxes: 16 columns x 32 rows (64-byte rows, 64-byte swizzle) / transposed: 32 x 16 (128-byte rows)
        if (op == MOE_GEMM_WGRAD) ok = ok && encode_3d_f32(&tO0, out0, N, M, E, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        else ok = ok && encode_3d_f32(&tO0, out0, M, N, E, 32, 16, CU_TENSOR_MAP_SWIZZLE_128B);
        tO1 = tO0;
        tAux = tO0;
    }
    if (!ok) return 1;
    const int grid = (sm_count / 2) * 2;
    p.ksplit = 1;
    if (wgrad && aux != nullptr) {
        // How the K ranges of the output tiles are spread over the CTA pairs when whole tiles do not fill the grid evenly
        // (measured, profiles/r02_wgrad_schedules.md):
        //  * three or more rounds of tiles: whole tiles round-robin (balanced to a few percent, one epilogue per tile);
        //  * stream-K (gemm.cuh) when a tile is cut in at most two or three pieces (tiles >= pairs / 2), the tile spans
        //    all of N (the A columns of a tile are not shared with another tile) and B fits L2 with room to spare: pairs
        //    work at unrelated K offsets, so every operand that is shared between tiles must come from L2, not from being
        //    read at the same time.  Config 2: 96 tiles on 74 pairs, 76 -> 67 us;
        //  * otherwise equal split-K: S = 2 parts per tile, or S = pairs / tiles when there are fewer tiles than half the
        //    pairs (12 tiles at E_local = 2 under 8-way expert parallelism), as long as a part keeps ~16 k-blocks.  Parts
        //    of sibling tiles run in lock-step, so shared operands are read once.
        const int pairs = grid / 2;
        const int n_nt = (N + bn - 1) / bn;
        const int64_t ntile = static_cast<int64_t>(E) * ((M + 255) / 256) * n_nt;
        const int64_t avg_kb = rows_cap / (64LL * (E > 0 ? E : 1));
        const bool b_in_l2 = static_cast<int64_t>(rows_cap) * N * 2 <= (48LL << 20);
        if (ntile >= 3LL * pairs) {
            p.ksplit = 1;
        } else if (n_nt == 1 && b_in_l2 && 2 * ntile >= pairs && E <= kSkMaxE && pairs <= kSkMaxPairs) {
            p.streamk = 1;
        } else {
            int S = 2;
            if (ntile * 2 <= pairs) {
                S = static_cast<int>(pairs / ntile);
                if (S > 8) S = 8;
                while (S > 2 && avg_kb / S < 16) --S;
            }
            p.ksplit = S;
        }
#ifdef MOE_EXPERIMENT_HOOKS
        if (getenv("MOE_WGRAD_NO_SPLIT") != nullptr) { p.streamk = 0; p.ksplit = 1; }
        if (getenv("MOE_WGRAD_STREAMK") != nullptr) { p.streamk = atoi(getenv("MOE_WGRAD_STREAMK")); if (p.streamk) p.ksplit = 1; }
#endif
    }

#define MOE_BN_ROWS_WIDE(EPI)                                                             \
    if (bn == 384) return launch_one<384, EPI, false>(tA, tB, tO0, tO1, tAux, p, grid, st); \
    MOE_BN_ROWS(EPI)
#define MOE_BN_ROWS(EPI)                                                                  \
    switch (bn) {                                                                         \
        case 256: return launch_one<256, EPI, false>(tA, tB, tO0, tO1, tAux, p, grid, st);      \
        case 192: return launch_one<192, EPI, false>(tA, tB, tO0, tO1, tAux, p, grid, st);      \
        case 128: return launch_one<128, EPI, false>(tA, tB, tO0, tO1, tAux, p, grid, st);      \
        default: return launch_one<64, EPI, false>(tA, tB, tO0, tO1, tAux, p, grid, st);        \
    }
    switch (op) {
        case MOE_GEMM_FC1:
            if (infer_fc1) { MOE_BN_ROWS(EPI_BIAS_GELU) }
            MOE_BN_ROWS(EPI_BIAS_GELU_DUAL)
        case MOE_GEMM_FC2: MOE_BN_ROWS_WIDE(EPI_BIAS)
        case MOE_GEMM_DGELU: MOE_BN_ROWS(EPI_DGELU)
        case MOE_GEMM_DGRAD: MOE_BN_ROWS_WIDE(EPI_PLAIN)
        case MOE_GEMM_WGRAD:
            if (bn == 384) return launch_one<384, EPI_F32, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            if (bn == 256) return launch_one<256, EPI_F32, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            if (bn == 192) return launch_one<192, EPI_F32, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            return launch_one<128, EPI_F32, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
        default:
            if (bn == 384) return launch_one<384, EPI_F32_T, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            if (bn == 256) return launch_one<256, EPI_F32_T, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            if (bn == 192) return launch_one<192, EPI_F32_T, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
            return launch_one<128, EPI_F32_T, true>(tA, tB, tO0, tO1, tAux, p, grid, st);
    }
#undef MOE_BN_ROWS
#undef MOE_BN_ROWS_WIDE
}

}  // namespace moe

#ifdef MOE_DBG_TIMELINE
// kernel experiments only: copy the timeline stamps of the last launch to the host (2*4*64*4 long longs)
extern "C" int moe_debug_timeline(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, moe::g_tl, sizeof(moe::g_tl)) == cudaSuccess ? 0 : 1;
}
#endif
