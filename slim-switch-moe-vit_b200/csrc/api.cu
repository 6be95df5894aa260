// api.cu — the extern "C" surface declared in include/moe_b200.h.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace moe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

static int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
}

static bool dims_ok(const char* fn, int64_t T, int d, int E, int k) {
    if (T <= 0 || d <= 0 || d % 64 != 0 || E <= 0 || E > 1024 || k < 1 || k > 8 || k > E) {
        set_error("%s: unsupported shape T=%lld d=%d E=%d k=%d (need d %% 64 == 0, 1 <= k <= min(8,E), E <= 1024)", fn,
                  (long long)T, d, E, k);
        return false;
    }
    return true;
}
static bool dtype_ok(const char* fn, int dt) {
    if (dt != MOE_DTYPE_F32 && dt != MOE_DTYPE_BF16) { set_error("%s: unsupported dtype code %d", fn, dt); return false; }
    return true;
}

}  // namespace moe

using namespace moe;

extern "C" {

const char* moe_last_error(void) { return g_err; }
int moe_version(void) { return 100; }

int64_t moe_rows_cap(int64_t T, int k, int E, int64_t capacity) {
    int64_t pairs = T * k;
    if (capacity > 0 && capacity < pairs && capacity * E < pairs) pairs = capacity * E;
    return (pairs + MOE_ROW_ALIGN - 1) / MOE_ROW_ALIGN * MOE_ROW_ALIGN + static_cast<int64_t>(MOE_ROW_ALIGN) * E;
}

size_t moe_gate_fwd_workspace_bytes(int d, int E) { return gate_fwd_workspace_bytes(d, E); }

int moe_gate_fwd(const void* x, int x_dtype, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask, int64_t T, int d, int E, int k,
                 int score_mode, int want_psum, float* logits, int32_t* idx, float* score, int32_t* tile_hist,
                 float* tile_psum, void* workspace, void* stream) {
    if (!dims_ok("moe_gate_fwd", T, d, E, k) || !dtype_ok("moe_gate_fwd", x_dtype)) return 1;
    if (score_mode != MOE_SCORE_TOPK_SOFTMAX && score_mode != MOE_SCORE_FULL_SOFTMAX) { set_error("moe_gate_fwd: bad score_mode %d", score_mode); return 1; }
    if (want_psum && tile_psum == nullptr) { set_error("moe_gate_fwd: want_psum needs tile_psum"); return 1; }
    // tensor-core path (certified routing) when the caller provides its workspace and the shape allows it; otherwise the
    // CUDA-core kernel, whose logits are bit-identical to the oracle's (LOGIT ORDER v1)
    if (workspace != nullptr && gate_mma_supported(x_dtype, d, E))
        return launch_gate_fwd_mma(x, Wg, bg, noise, token_mask, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist,
                                   tile_psum, workspace, static_cast<cudaStream_t>(stream));
    return check(launch_gate_fwd(x, x_dtype, Wg, bg, noise, token_mask, T, d, E, k, score_mode, want_psum, logits, idx, score, tile_hist,
                                 tile_psum, static_cast<cudaStream_t>(stream)),
                 "moe_gate_fwd");
}

int moe_route_scan(const int32_t* tile_hist, const float* tile_psum, int ntiles, int E, int64_t capacity,
                   int32_t* tile_base, int32_t* count, int32_t* kept, int32_t* seg_start, int32_t* tile_expert,
                   int32_t* num_mtiles, int max_mtiles, float* psum, int aux_mode, int64_t T, int k, float* aux_loss,
                   float* aux_coef, int64_t slab_rows, void* stream) {
    if (slab_rows < 0 || slab_rows % MOE_ROW_ALIGN != 0 || (slab_rows > 0 && slab_rows < capacity)) {
        set_error("moe_route_scan: slab_rows=%lld must be 0 (packed) or a multiple of %d that is >= capacity=%lld", (long long)slab_rows,
                  MOE_ROW_ALIGN, (long long)capacity);
        return 1;
    }
    if (ntiles <= 0 || E <= 0 || E > 1024 || capacity <= 0) { set_error("moe_route_scan: bad arguments ntiles=%d E=%d capacity=%lld", ntiles, E, (long long)capacity); return 1; }
    if (aux_mode != MOE_AUX_NONE) {
        if ((aux_mode != MOE_AUX_SWITCH && aux_mode != MOE_AUX_GSHARD) || tile_psum == nullptr || psum == nullptr ||
            aux_loss == nullptr || aux_coef == nullptr || T <= 0 || k < 1) {
            set_error("moe_route_scan: aux_mode %d needs tile_psum, psum, aux_loss, aux_coef, T > 0 and k >= 1", aux_mode);
            return 1;
        }
    }
    return check(launch_route_scan(tile_hist, tile_psum, ntiles, E, capacity, tile_base, count, kept, seg_start,
                                   tile_expert, num_mtiles, max_mtiles, psum, aux_mode, T, k, aux_loss, aux_coef, slab_rows,
                                   static_cast<cudaStream_t>(stream)),
                 "moe_route_scan");
}

// ---- expert parallelism over peer memory (csrc/ep_peer.cu)
static bool peers_ok(const char* fn, void* const* ptrs, int W, int rank, PeerRows* out, int rows_per_rank) {
    if (W < 1 || W > kMaxPeers || rank < 0 || rank >= W || ptrs == nullptr) {
        set_error("%s: expert-parallel group of %d ranks (rank %d): 1..%d ranks of one NVLink domain are supported", fn, W, rank, kMaxPeers);
        return false;
    }
    *out = PeerRows{};
    for (int i = 0; i < W; ++i) {
        if (ptrs[i] == nullptr) { set_error("%s: peer pointer %d is NULL", fn, i); return false; }
        out->base[i] = ptrs[i];
    }
    out->rows_per_rank = rows_per_rank;
    out->n = W;
    return true;
}

int moe_ep_heap_alloc(size_t bytes, void** ptr, void* handle_out) {
    if (bytes == 0 || ptr == nullptr || handle_out == nullptr) { set_error("moe_ep_heap_alloc: bad arguments"); return 1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == MOE_IPC_HANDLE_BYTES, "IPC handle size");
    if (check(cudaMalloc(ptr, bytes), "moe_ep_heap_alloc: cudaMalloc")) return 1;
    if (check(cudaMemset(*ptr, 0, bytes), "moe_ep_heap_alloc: cudaMemset") ||
        check(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), *ptr), "moe_ep_heap_alloc: cudaIpcGetMemHandle") ||
        check(cudaDeviceSynchronize(), "moe_ep_heap_alloc: sync")) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return 1;
    }
    return 0;
}
int moe_ep_heap_open(const void* handle, void** ptr) {
    if (handle == nullptr || ptr == nullptr) { set_error("moe_ep_heap_open: bad arguments"); return 1; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return check(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "moe_ep_heap_open: cudaIpcOpenMemHandle");
}
int moe_ep_heap_close(void* ptr) { return check(cudaIpcCloseMemHandle(ptr), "moe_ep_heap_close"); }
int moe_ep_heap_free(void* ptr) { return check(cudaFree(ptr), "moe_ep_heap_free"); }

int moe_ep_barrier(void* const* flags, int32_t* epoch, int rank, int W, int32_t* status, void* stream) {
    PeerRows fl;
    if (!peers_ok("moe_ep_barrier", flags, W, rank, &fl, 0)) return 1;
    return check(launch_ep_barrier(fl, epoch, rank, W, status, static_cast<cudaStream_t>(stream)), "moe_ep_barrier");
}

int moe_ep_exchange_counts(const int32_t* kept, void* const* kept_all, void* const* flags, int32_t* epoch, int rank, int W,
                           int E_local, int64_t rows_per_rank, int32_t* dst_row, int32_t* kept_local, int32_t* seg_start,
                           int32_t* tile_expert, int32_t* num_mtiles, int max_mtiles, int32_t* status, void* stream) {
    PeerRows ka, fl;
    if (!peers_ok("moe_ep_exchange_counts", kept_all, W, rank, &ka, 0) || !peers_ok("moe_ep_exchange_counts", flags, W, rank, &fl, 0)) return 1;
    if (E_local < 1 || W * E_local > 1024 || rows_per_rank <= 0 || rows_per_rank % MOE_ROW_ALIGN != 0 ||
        rows_per_rank * W > 0x7fffffffLL || max_mtiles < 1) {
        set_error("moe_ep_exchange_counts: bad arguments E_local=%d rows_per_rank=%lld", E_local, (long long)rows_per_rank);
        return 1;
    }
    return check(launch_ep_exchange_counts(kept, ka, fl, epoch, rank, W, E_local, static_cast<int>(rows_per_rank), dst_row, kept_local,
                                           seg_start, tile_expert, num_mtiles, max_mtiles, status, static_cast<cudaStream_t>(stream)),
                 "moe_ep_exchange_counts");
}

int moe_dispatch_fwd_peer(const void* x, int x_dtype, const int32_t* idx, const int32_t* tile_base, const int32_t* dst_row, int64_t T,
                          int d, int E, int k, int64_t capacity, void* const* xbuf_peers, int rank, int W, int64_t rows_per_rank,
                          const int32_t* seg_start_local, const int32_t* kept_local, int32_t* pos, void* stream) {
    PeerRows xr;
    if (!dims_ok("moe_dispatch_fwd_peer", T, d, E, k) || !dtype_ok("moe_dispatch_fwd_peer", x_dtype) ||
        !peers_ok("moe_dispatch_fwd_peer", xbuf_peers, W, rank, &xr, static_cast<int>(rows_per_rank))) return 1;
    if (E % W != 0) { set_error("moe_dispatch_fwd_peer: E=%d is not a multiple of the group size %d", E, W); return 1; }
    return check(launch_dispatch_fwd_rows(x, x_dtype, idx, tile_base, dst_row, T, d, E, k, capacity, pos, nullptr, xr, xbuf_peers[rank],
                                          seg_start_local, kept_local, E / W, static_cast<cudaStream_t>(stream)),
                 "moe_dispatch_fwd_peer");
}

int moe_combine_fwd_peer(void* const* ybuf_peers, int rank, int W, int64_t rows_per_rank, const int32_t* pos, const float* score,
                         int64_t T, int d, int k, void* out, int out_dtype, void* stream) {
    PeerRows yr;
    if (!dims_ok("moe_combine_fwd_peer", T, d, k, k) || !dtype_ok("moe_combine_fwd_peer", out_dtype) ||
        !peers_ok("moe_combine_fwd_peer", ybuf_peers, W, rank, &yr, static_cast<int>(rows_per_rank))) return 1;
    return check(launch_combine_fwd_rows(yr, pos, score, T, d, k, out, out_dtype, sm_count(), static_cast<cudaStream_t>(stream)),
                 "moe_combine_fwd_peer");
}

int moe_combine_bwd_peer(const void* dy, int dy_dtype, void* const* ybuf_peers, void* const* dybuf_peers, int rank, int W,
                         int64_t rows_per_rank, const int32_t* pos, const float* score, const int32_t* seg_start_local,
                         const int32_t* kept_local, int E_local, int64_t T, int d, int k, float* dscore, void* stream) {
    PeerRows yr, dr;
    if (!dims_ok("moe_combine_bwd_peer", T, d, k, k) || !dtype_ok("moe_combine_bwd_peer", dy_dtype) ||
        !peers_ok("moe_combine_bwd_peer", ybuf_peers, W, rank, &yr, static_cast<int>(rows_per_rank)) ||
        !peers_ok("moe_combine_bwd_peer", dybuf_peers, W, rank, &dr, static_cast<int>(rows_per_rank))) return 1;
    return check(launch_combine_bwd_rows(dy, dy_dtype, yr, pos, score, seg_start_local, kept_local, E_local, T, d, k, dr, dybuf_peers[rank],
                                         dscore, static_cast<cudaStream_t>(stream)),
                 "moe_combine_bwd_peer");
}

int moe_gate_dispatch_bwd_peer(void* const* dxbuf_peers, int rank, int W, int64_t rows_per_rank, const int32_t* pos, const float* logits,
                               const int32_t* idx, const float* score, const float* dscore, const float* dpsum, const float* Wg,
                               int64_t T, int d, int E, int k, int score_mode, float* dlogits, void* dx, int dx_dtype, void* stream) {
    PeerRows xr;
    if (!dims_ok("moe_gate_dispatch_bwd_peer", T, d, E, k) || !dtype_ok("moe_gate_dispatch_bwd_peer", dx_dtype) ||
        !peers_ok("moe_gate_dispatch_bwd_peer", dxbuf_peers, W, rank, &xr, static_cast<int>(rows_per_rank))) return 1;
    if (!gate_dispatch_bwd_mma_supported(d, E, k)) {
        set_error("moe_gate_dispatch_bwd_peer: shape d=%d E=%d k=%d is outside the peer-memory kernel (E <= 64, staged rows must fit shared memory)", d, E, k);
        return 1;
    }
    return check(launch_gate_dispatch_bwd_mma(xr, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx, dx_dtype,
                                              static_cast<cudaStream_t>(stream)),
                 "moe_gate_dispatch_bwd_peer");
}

int moe_ep_tables(const int32_t* kept_recv, int W, int E_local, int32_t* slab_dst, int32_t* kept_local, int32_t* seg_start,
                  int32_t* tile_expert, int32_t* num_mtiles, int max_mtiles, void* stream) {
    if (W < 1 || E_local < 1 || E_local > 1024 || max_mtiles < 1) { set_error("moe_ep_tables: bad arguments W=%d E_local=%d", W, E_local); return 1; }
    return check(launch_ep_tables(kept_recv, W, E_local, slab_dst, kept_local, seg_start, tile_expert, num_mtiles, max_mtiles,
                                  static_cast<cudaStream_t>(stream)),
                 "moe_ep_tables");
}

int moe_ep_repack(const void* src, void* dst, const int32_t* kept_recv, const int32_t* slab_dst, const int32_t* seg_start,
                  const int32_t* kept_local, int W, int E_local, int64_t slab_rows, int d, int to_packed, void* stream) {
    if (W < 1 || E_local < 1 || slab_rows <= 0 || slab_rows % MOE_ROW_ALIGN != 0 || d <= 0 || d % 8 != 0) {
        set_error("moe_ep_repack: bad arguments W=%d E_local=%d slab_rows=%lld d=%d", W, E_local, (long long)slab_rows, d);
        return 1;
    }
    return check(launch_ep_repack(src, dst, kept_recv, slab_dst, seg_start, kept_local, W, E_local, slab_rows, d, to_packed,
                                  static_cast<cudaStream_t>(stream)),
                 "moe_ep_repack");
}

int moe_dispatch_fwd(const void* x, int x_dtype, const int32_t* idx, const int32_t* tile_base, const int32_t* seg_start,
                     const int32_t* kept, int64_t T, int d, int E, int k, int64_t capacity, int32_t* pos,
                     int32_t* row_src, void* xbuf, void* stream) {
    if (!dims_ok("moe_dispatch_fwd", T, d, E, k) || !dtype_ok("moe_dispatch_fwd", x_dtype)) return 1;
    return check(launch_dispatch_fwd(x, x_dtype, idx, tile_base, seg_start, kept, T, d, E, k, capacity, pos, row_src,
                                     xbuf, static_cast<cudaStream_t>(stream)),
                 "moe_dispatch_fwd");
}

int moe_combine_fwd(const void* ybuf, const int32_t* pos, const float* score, int64_t T, int d, int k, void* out,
                    int out_dtype, void* stream) {
    if (!dims_ok("moe_combine_fwd", T, d, 8, k) || !dtype_ok("moe_combine_fwd", out_dtype)) return 1;
    return check(launch_combine_fwd(ybuf, pos, score, T, d, k, out, out_dtype, sm_count(), static_cast<cudaStream_t>(stream)),
                 "moe_combine_fwd");
}

int moe_combine_bwd(const void* dy, int dy_dtype, const void* ybuf, const int32_t* pos, const float* score,
                    const int32_t* seg_start, const int32_t* kept, int64_t T, int d, int k, int E, void* dybuf,
                    float* dscore, void* stream) {
    if (!dims_ok("moe_combine_bwd", T, d, E, k) || !dtype_ok("moe_combine_bwd", dy_dtype)) return 1;
    return check(launch_combine_bwd(dy, dy_dtype, ybuf, pos, score, seg_start, kept, T, d, k, E, dybuf, dscore,
                                    static_cast<cudaStream_t>(stream)),
                 "moe_combine_bwd");
}

int moe_gate_bwd(const float* logits, const int32_t* idx, const float* score, const float* dscore, const float* dpsum,
                 int64_t T, int E, int k, int score_mode, float* dlogits, void* stream) {
    if (T <= 0 || E <= 0 || k < 1 || k > 8) { set_error("moe_gate_bwd: bad shape"); return 1; }
    return check(launch_gate_bwd(logits, idx, score, dscore, dpsum, T, E, k, score_mode, dlogits, static_cast<cudaStream_t>(stream)),
                 "moe_gate_bwd");
}

int moe_dispatch_bwd(const void* dxbuf, const int32_t* pos, const float* dlogits, const int32_t* idx, const float* Wg,
                     int64_t T, int d, int E, int k, int dense_dlogits, void* dx, int dx_dtype, void* stream) {
    if (!dims_ok("moe_dispatch_bwd", T, d, E, k) || !dtype_ok("moe_dispatch_bwd", dx_dtype)) return 1;
    return check(launch_dispatch_bwd(dxbuf, pos, dlogits, idx, Wg, T, d, E, k, dense_dlogits, dx, dx_dtype, sm_count(),
                                     static_cast<cudaStream_t>(stream)),
                 "moe_dispatch_bwd");
}

size_t moe_gate_wgrad_workspace_bytes(int64_t T, int d, int E) { return gate_wgrad_workspace_bytes(T, d, E); }

int moe_gate_wgrad(const float* dlogits, const void* x, int x_dtype, int64_t T, int d, int E, void* workspace,
                   float* dWg, float* dbg, void* stream) {
    if (!dims_ok("moe_gate_wgrad", T, d, E, 1) || !dtype_ok("moe_gate_wgrad", x_dtype)) return 1;
    return check(launch_gate_wgrad(dlogits, x, x_dtype, T, d, E, workspace, dWg, dbg, static_cast<cudaStream_t>(stream)),
                 "moe_gate_wgrad");
}

int moe_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
    if (n <= 0 || n % 8 != 0) { set_error("moe_cast_bf16: n=%lld must be a positive multiple of 8", (long long)n); return 1; }
    return check(launch_cast_bf16(src, dst, n, sm_count(), static_cast<cudaStream_t>(stream)), "moe_cast_bf16");
}

int moe_cast_bf16_pair(const float* src0, void* dst0, int64_t n0, const float* src1, void* dst1, int64_t n1, void* stream) {
    if (n0 <= 0 || n0 % 8 != 0 || n1 <= 0 || n1 % 8 != 0 || src0 == nullptr || dst0 == nullptr || src1 == nullptr || dst1 == nullptr) {
        set_error("moe_cast_bf16_pair: two non-null ranges of positive multiples of 8 elements required (n0=%lld n1=%lld)", (long long)n0, (long long)n1);
        return 1;
    }
    return check(launch_cast_bf16_pair(src0, dst0, n0, src1, dst1, n1, sm_count(), static_cast<cudaStream_t>(stream)), "moe_cast_bf16_pair");
}

size_t moe_segment_colsum_workspace_bytes(int64_t rows_cap, int cols) { return segment_colsum_workspace_bytes(rows_cap, cols); }

int moe_segment_colsum(const void* buf, const int32_t* seg_start, int64_t rows_cap, int E, int cols, void* workspace, float* out,
                       void* stream) {
    if (E <= 0 || cols <= 0 || cols % 8 != 0 || rows_cap <= 0 || rows_cap % MOE_ROW_ALIGN != 0 || workspace == nullptr) {
        set_error("moe_segment_colsum: bad arguments (E=%d cols=%d rows_cap=%lld; cols %% 8 == 0, rows_cap %% %d == 0, workspace required)",
                  E, cols, (long long)rows_cap, MOE_ROW_ALIGN);
        return 1;
    }
    return check(launch_segment_colsum(buf, seg_start, rows_cap, E, cols, workspace, out, static_cast<cudaStream_t>(stream)),
                 "moe_segment_colsum");
}

int moe_gate_dispatch_bwd(const void* dxbuf, const int32_t* pos, const float* logits, const int32_t* idx, const float* score,
                          const float* dscore, const float* dpsum, const float* Wg, int64_t T, int d, int E, int k, int score_mode,
                          float* dlogits, void* dx, int dx_dtype, void* stream) {
    if (!dims_ok("moe_gate_dispatch_bwd", T, d, E, k) || !dtype_ok("moe_gate_dispatch_bwd", dx_dtype)) return 1;
    return check(launch_gate_dispatch_bwd(dxbuf, pos, logits, idx, score, dscore, dpsum, Wg, T, d, E, k, score_mode, dlogits, dx,
                                          dx_dtype, static_cast<cudaStream_t>(stream)),
                 "moe_gate_dispatch_bwd");
}

size_t moe_addln_bwd_workspace_bytes(int64_t T, int d) { return addln_bwd_workspace_bytes(T, d); }

static bool addln_ok(const char* fn, int64_t T, int d) {
    if (T <= 0 || d <= 0 || d % 4 != 0 || d > 1024) { set_error("%s: unsupported shape T=%lld d=%d (need d %% 4 == 0, d <= 1024)", fn, (long long)T, d); return false; }
    return true;
}

int moe_addln_fwd(const float* x_in, const void* delta, int delta_dtype, const float* gamma, const float* beta, float eps, int64_t T,
                  int d, float* x_out, void* n, int n_dtype, float* mean, float* rstd, void* stream) {
    if (!addln_ok("moe_addln_fwd", T, d) || !dtype_ok("moe_addln_fwd", n_dtype) || (delta != nullptr && !dtype_ok("moe_addln_fwd", delta_dtype))) return 1;
    if (delta != nullptr && x_out == nullptr) { set_error("moe_addln_fwd: x_out is required when delta is given"); return 1; }
    return check(launch_addln_fwd(x_in, delta, delta_dtype, gamma, beta, eps, T, d, x_out, n, n_dtype, mean, rstd,
                                  static_cast<cudaStream_t>(stream)),
                 "moe_addln_fwd");
}

int moe_addln_bwd(const void* dn, int n_dtype, const float* dx_out, const float* x, const float* mean, const float* rstd,
                  const float* gamma, int64_t T, int d, float* dx_in, void* d_delta, int delta_dtype, void* workspace, float* dgamma,
                  float* dbeta, void* stream) {
    if (!addln_ok("moe_addln_bwd", T, d) || !dtype_ok("moe_addln_bwd", n_dtype) || (d_delta != nullptr && !dtype_ok("moe_addln_bwd", delta_dtype))) return 1;
    if (workspace == nullptr) { set_error("moe_addln_bwd: workspace (moe_addln_bwd_workspace_bytes) required"); return 1; }
    return check(launch_addln_bwd(dn, n_dtype, dx_out, x, mean, rstd, gamma, T, d, dx_in, d_delta, delta_dtype, workspace, dgamma, dbeta,
                                  static_cast<cudaStream_t>(stream)),
                 "moe_addln_bwd");
}

size_t moe_colsum_workspace_bytes(int64_t rows, int cols) { return colsum_workspace_bytes(rows, cols); }

int moe_colsum(const void* buf, int dtype, int64_t rows, int cols, void* workspace, float* out, void* stream) {
    if (rows <= 0 || cols <= 0 || cols % 8 != 0 || workspace == nullptr || !dtype_ok("moe_colsum", dtype)) {
        if (rows <= 0 || cols <= 0 || cols % 8 != 0 || workspace == nullptr) set_error("moe_colsum: need rows > 0, cols %% 8 == 0 and a workspace (rows=%lld cols=%d)", (long long)rows, cols);
        return 1;
    }
    return check(launch_colsum(buf, dtype, rows, cols, workspace, out, static_cast<cudaStream_t>(stream)), "moe_colsum");
}

size_t moe_slab_colsum_bytes(int64_t rows_cap, int cols) { return static_cast<size_t>(rows_cap / 32) * cols * sizeof(float); }

int moe_slab_colsum_final(const float* part, const int32_t* seg_start, int E, int cols, float* out, void* stream) {
    if (part == nullptr || seg_start == nullptr || out == nullptr || E <= 0 || cols <= 0) { set_error("moe_slab_colsum_final: bad arguments"); return 1; }
    return check(launch_slab_colsum_final(part, seg_start, E, cols, out, static_cast<cudaStream_t>(stream)), "moe_slab_colsum_final");
}

size_t moe_wgrad_flags_bytes(int E, int M, int N) {
    // one int per (tile, CTA rank, epilogue warp); tiles counted for the narrowest tile the launcher may pick (128)
    return static_cast<size_t>(E) * ((M + 255) / 256) * ((N + 127) / 128) * 2 * 16 * sizeof(int32_t);
}

int moe_grouped_gemm(int op, const void* A, const void* B, void* out0, void* out1, const float* bias, const void* aux,
                     const int32_t* tile_expert, const int32_t* num_mtiles, const int32_t* seg_start, int64_t rows_cap,
                     int E, int M, int N, int K, void* stream) {
    return launch_grouped_gemm(op, A, B, out0, out1, bias, aux, tile_expert, num_mtiles, seg_start, rows_cap, E, M, N, K,
                               sm_count(), static_cast<cudaStream_t>(stream));
}

int moe_expert_ffn_fwd(const void* xbuf, const void* W1b, const float* b1, const void* W2b, const float* b2,
                       const int32_t* tile_expert, const int32_t* num_mtiles, int64_t rows_cap, int d, int h, int E,
                       void* G, void* H, void* Y, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = launch_grouped_gemm(MOE_GEMM_FC1, xbuf, W1b, G, H, b1, nullptr, tile_expert, num_mtiles, nullptr, rows_cap,
                                 E, 0, h, d, sm_count(), st);
    if (rc) return rc;
    return launch_grouped_gemm(MOE_GEMM_FC2, H, W2b, Y, nullptr, b2, nullptr, tile_expert, num_mtiles, nullptr, rows_cap,
                               E, 0, d, h, sm_count(), st);
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t moe_expert_ffn_bwd_workspace_bytes(int64_t rows_cap, int d, int h, int E) {
    // [slab column sums of dU (db1)] [row-block column sums of dY (db2)] [split-K flags of the two weight gradients]
    return align256(moe_slab_colsum_bytes(rows_cap, h)) + align256(segment_colsum_workspace_bytes(rows_cap, d)) +
           align256(moe_wgrad_flags_bytes(E, h, d));
}

size_t moe_workspace_bytes(int64_t T, int d, int h, int E, int k, int64_t capacity) {
    // one scratch buffer that covers every workspace argument of one layer's forward + backward (they are never live
    // at the same time: the expert-FFN backward finishes before the gate weight gradient starts)
    const int64_t rows_cap = moe_rows_cap(T, k, E, capacity);
    const size_t a = moe_expert_ffn_bwd_workspace_bytes(rows_cap, d, h, E);
    const size_t b = gate_wgrad_workspace_bytes(T, d, E);
    return a > b ? a : b;
}

int moe_expert_ffn_bwd(const void* dybuf, const void* xbuf, const void* G, const void* H, const void* W1b,
                       const void* W2b, const int32_t* tile_expert, const int32_t* num_mtiles,
                       const int32_t* seg_start, int64_t rows_cap, int d, int h, int E, void* dU, void* dxbuf,
                       float* dW1, float* db1, float* dW2, float* db2, void* workspace, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int sms = sm_count();
    int rc;
    if (workspace == nullptr) { set_error("moe_expert_ffn_bwd: workspace (moe_expert_ffn_bwd_workspace_bytes(rows_cap, d, h, E)) required"); return 1; }
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* slab_sums = reinterpret_cast<float*>(ws);
    void* colsum_ws = ws + align256(moe_slab_colsum_bytes(rows_cap, h));
    void* flags = static_cast<uint8_t*>(colsum_ws) + align256(segment_colsum_workspace_bytes(rows_cap, d));
    // dU = (dY W2) * G, G = gelu'(U)                   [rows, h]   K = d, B = W2 [E, d, h] read MN-major
    // (the epilogue leaves the column sums of every 32-row slab of dU behind: db1 without a second pass over dU)
    rc = launch_grouped_gemm(MOE_GEMM_DGELU, dybuf, W2b, dU, slab_sums, nullptr, G, tile_expert, num_mtiles, nullptr,
                             rows_cap, E, 0, h, d, sms, st);
    if (rc) return rc;
    // split-K flags of the two weight gradients (cleared here; the kernels leave them clear)
    if (check(cudaMemsetAsync(flags, 0, moe_wgrad_flags_bytes(E, h, d), st), "wgrad flags memset")) return 1;
    // dW2[e] = dY_e^T H_e = (H_e^T dY_e)^T      [d, h]   computed with M = h, stored transposed
    rc = launch_grouped_gemm(MOE_GEMM_WGRAD_T, H, dybuf, dW2, nullptr, nullptr, flags, nullptr, nullptr, seg_start,
                             rows_cap, E, h, d, 0, sms, st);
    if (rc) return rc;
    // dW1[e] = dU_e^T X_e                          [h, d]
    rc = launch_grouped_gemm(MOE_GEMM_WGRAD, dU, xbuf, dW1, nullptr, nullptr, flags, nullptr, nullptr, seg_start,
                             rows_cap, E, h, d, 0, sms, st);
    if (rc) return rc;
    // dX = dU W1                                   [rows, d]   K = h, B = W1 [E, h, d] read MN-major
    rc = launch_grouped_gemm(MOE_GEMM_DGRAD, dU, W1b, dxbuf, nullptr, nullptr, nullptr, tile_expert, num_mtiles, nullptr,
                             rows_cap, E, 0, d, h, sms, st);
    if (rc) return rc;
    if (check(launch_segment_colsum(dybuf, seg_start, rows_cap, E, d, colsum_ws, db2, st), "db2 colsum")) return 1;
    return check(launch_slab_colsum_final(slab_sums, seg_start, E, h, db1, st), "db1 slab sums");
}

}  // extern "C"
