// common.cuh — declarations shared by the translation units of libmoe_b200.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/moe_b200.h"

namespace moe {

// A packed row buffer addressed through `pos` / row indices that may be spread over the GPUs of an expert-parallel
// group (NVLink peer memory, csrc/ep_peer.cu): global row r is row r % rows_per_rank of rank r / rows_per_rank.
// A local buffer is the n == 1 case (base[0], rows_per_rank = INT_MAX).  base[0] == nullptr: no buffer.
constexpr int kMaxPeers = 8;
struct PeerRows {
    void* base[kMaxPeers];
    int rows_per_rank;
    int n;
};
inline PeerRows local_rows(const void* p) {
    PeerRows r{};
    r.base[0] = const_cast<void*>(p);
    r.rows_per_rank = 0x7fffffff;
    r.n = 1;
    return r;
}
#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ T* peer_row(const PeerRows& pr, int row, int d) {
    if (pr.n == 1) return static_cast<T*>(pr.base[0]) + static_cast<size_t>(row) * d;
    const int rk = row / pr.rows_per_rank;
    void* b = pr.base[0];   // select chain: a dynamic index would make the compiler copy the parameter struct to local memory
#pragma unroll
    for (int i = 1; i < kMaxPeers; ++i)
        if (rk == i) b = pr.base[i];
    return static_cast<T*>(b) + static_cast<size_t>(row - rk * pr.rows_per_rank) * d;
}
#endif

// routing.cu
cudaError_t launch_gate_fwd(const void* x, int x_dtype, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask, int64_t T, int d, int E, int k,
                            int score_mode, int want_psum, float* logits, int* idx, float* score, int* tile_hist,
                            float* tile_psum, cudaStream_t st);
cudaError_t launch_route_scan(const int* tile_hist, const float* tile_psum, int ntiles, int E, long long capacity,
                              int* tile_base, int* count, int* kept, int* seg_start, int* tile_expert, int* num_mtiles,
                              int max_mtiles, float* psum, int aux_mode, long long tokens, int k, float* aux_loss,
                              float* aux_coef, long long slab_rows, cudaStream_t st);
cudaError_t launch_ep_tables(const int* kept_recv, int W, int El, int* slab_dst, int* kept_loc, int* seg_start,
                             int* tile_expert, int* num_mtiles, int max_mtiles, cudaStream_t st);
cudaError_t launch_ep_repack(const void* src, void* dst, const int* kept_recv, const int* slab_dst, const int* seg_start,
                             const int* kept_loc, int W, int El, long long slab_rows, int d, int to_packed, cudaStream_t st);
cudaError_t launch_dispatch_fwd(const void* x, int x_dtype, const int* idx, const int* tile_base, const int* seg_start,
                                const int* kept, int64_t T, int d, int E, int k, long long capacity, int* pos,
                                int* row_src, void* xbuf, cudaStream_t st);
cudaError_t launch_combine_fwd(const void* ybuf, const int* pos, const float* score, int64_t T, int d, int k, void* out,
                               int out_dtype, int sm_count, cudaStream_t st);
cudaError_t launch_combine_bwd(const void* dy, int dy_dtype, const void* ybuf, const int* pos, const float* score,
                               const int* seg_start, const int* kept, int64_t T, int d, int k, int E, void* dybuf,
                               float* dscore, cudaStream_t st);
cudaError_t launch_gate_bwd(const float* logits, const int* idx, const float* score, const float* dscore,
                            const float* dpsum, int64_t T, int E, int k, int score_mode, float* dlogits,
                            cudaStream_t st);
cudaError_t launch_dispatch_bwd(const void* dxbuf, const int* pos, const float* dlogits, const int* idx, const float* Wg,
                                int64_t T, int d, int E, int k, int dense_dlogits, void* dx, int dx_dtype, int sm_count,
                                cudaStream_t st);
size_t gate_wgrad_workspace_bytes(int64_t T, int d, int E);
cudaError_t launch_gate_wgrad(const float* dlogits, const void* x, int x_dtype, int64_t T, int d, int E, void* workspace,
                              float* dWg, float* dbg, cudaStream_t st);
cudaError_t launch_cast_bf16(const float* src, void* dst, int64_t n, int sm_count, cudaStream_t st);
cudaError_t launch_cast_bf16_pair(const float* src0, void* dst0, int64_t n0, const float* src1, void* dst1, int64_t n1,
                                  int sm_count, cudaStream_t st);
size_t segment_colsum_workspace_bytes(int64_t rows_cap, int cols);
cudaError_t launch_slab_colsum_final(const float* part, const int* seg_start, int E, int cols, float* out, cudaStream_t st);
cudaError_t launch_segment_colsum(const void* buf, const int* seg_start, int64_t rows_cap, int E, int cols, void* workspace,
                                  float* out, cudaStream_t st);
cudaError_t launch_gate_dispatch_bwd(const void* dxbuf, const int* pos, const float* logits, const int* idx, const float* score,
                                     const float* dscore, const float* dpsum, const float* Wg, int64_t T, int d, int E, int k,
                                     int score_mode, float* dlogits, void* dx, int dx_dtype, cudaStream_t st);

// the same three with the packed rows addressed through PeerRows (expert parallelism over peer memory); pad rows are
// zeroed in the LOCAL buffer (xpad / dypad) for the n_pad segments of pad_seg / pad_kept
cudaError_t launch_dispatch_fwd_rows(const void* x, int x_dtype, const int* idx, const int* tile_base, const int* seg_start,
                                     int64_t T, int d, int E, int k, long long capacity, int* pos, int* row_src,
                                     const PeerRows& xrows, void* xpad, const int* pad_seg, const int* pad_kept, int n_pad,
                                     cudaStream_t st);
cudaError_t launch_combine_fwd_rows(const PeerRows& yrows, const int* pos, const float* score, int64_t T, int d, int k, void* out,
                                    int out_dtype, int sm_count, cudaStream_t st);
cudaError_t launch_combine_bwd_rows(const void* dy, int dy_dtype, const PeerRows& yrows, const int* pos, const float* score,
                                    const int* pad_seg, const int* pad_kept, int n_pad, int64_t T, int d, int k,
                                    const PeerRows& dyrows, void* dypad, float* dscore, cudaStream_t st);

// ep_peer.cu — expert parallelism over NVLink peer memory: counts exchange + packed layout, device barrier
cudaError_t launch_ep_exchange_counts(const int* kept, const PeerRows& kept_all, const PeerRows& flags, int* epoch, int rank, int W,
                                      int El, int rows_per_rank, int* dst_row, int* kept_local, int* seg_start, int* tile_expert,
                                      int* num_mtiles, int max_mtiles, int* status, cudaStream_t st);
cudaError_t launch_ep_barrier(const PeerRows& flags, int* epoch, int rank, int W, int* status, cudaStream_t st);

// gate_bwd_mma.cu — the same pass with dlogits Wg on the tensor cores (tf32 mma.sync); dXbuf rows may live on peer GPUs
bool gate_dispatch_bwd_mma_supported(int d, int E, int k);
cudaError_t launch_gate_dispatch_bwd_mma(const PeerRows& rows, const int* pos, const float* logits, const int* idx,
                                         const float* score, const float* dscore, const float* dpsum, const float* Wg, int64_t T,
                                         int d, int E, int k, int score_mode, float* dlogits, void* dx, int dx_dtype,
                                         cudaStream_t st);

// block_fusion.cu
size_t addln_bwd_workspace_bytes(int64_t T, int d);
cudaError_t launch_addln_fwd(const float* x_in, const void* delta, int delta_dtype, const float* gamma, const float* beta, float eps,
                             int64_t T, int d, float* x_out, void* n, int n_dtype, float* mean, float* rstd, cudaStream_t st);
cudaError_t launch_addln_bwd(const void* dn, int n_dtype, const float* dx_out, const float* x, const float* mean, const float* rstd,
                             const float* gamma, int64_t T, int d, float* dx_in, void* d_delta, int delta_dtype, void* workspace,
                             float* dgamma, float* dbeta, cudaStream_t st);

size_t colsum_workspace_bytes(int64_t rows, int cols);
cudaError_t launch_colsum(const void* buf, int dtype, int64_t rows, int cols, void* workspace, float* out, cudaStream_t st);

// gate_mma.cu — tensor-core gate with certified routing (bf16 activations, E <= 64); returns 0 on success
bool gate_mma_supported(int x_dtype, int d, int E);
size_t gate_fwd_workspace_bytes(int d, int E);
int launch_gate_fwd_mma(const void* x, const float* Wg, const float* bg, const float* noise, const uint8_t* token_mask,
                        int64_t T, int d, int E, int k, int score_mode, int want_psum, float* logits, int* idx, float* score,
                        int* tile_hist, float* tile_psum, void* workspace, cudaStream_t st);

// gemm_launch.cu — TMA descriptor of a [outer, inner] bf16 row-major tensor, 128-byte swizzle (false + set_error on failure)
bool encode_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer);

// gemm_launch.cu — returns 0 on success, otherwise sets the error string via set_error()
int launch_grouped_gemm(int op, const void* A, const void* B, void* out0, void* out1, const float* bias, const void* aux,
                        const int* tile_expert, const int* num_mtiles, const int* seg_start, int64_t rows_cap, int E,
                        int M, int N, int K, int sm_count, cudaStream_t st);

// api.cu
void set_error(const char* fmt, ...);
int sm_count();

}  // namespace moe
