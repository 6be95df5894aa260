/*
 * moe_b200.h — C ABI of libmoe_b200.so, the sm_100a implementation of the Switch-style MoE layer
 * that d0-rb/slim-switch-moe-vit obtains from FastMoE (`fmoe.FMoETransformerMLP`, imported at
 * /root/reference/models/resMoE.py:6, wrapped at models/resMoE.py:15-29, invoked at
 * models/vision_transformer.py:321 and models/resMoE.py:121,143).
 *
 * The reference has no FFI of its own for this path: its native half is FastMoE's `fmoe_cuda`
 * torch extension (un-vendored).  Each entry point below names the fmoe_cuda op / Python step it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory unless stated; caller allocates all
 *     outputs and workspaces; `stream` is a cudaStream_t passed as void*; all work is enqueued
 *     asynchronously, nothing synchronises the host.
 *   - return value: 0 = OK, non-zero = error; moe_last_error() (thread-local) describes it.
 *   - dtype codes: MOE_DTYPE_F32 = 0, MOE_DTYPE_BF16 = 1.
 *   - packed row buffers ("xbuf", "G", "H", "Y", ...) are [rows_cap, cols] bf16 row-major;
 *     expert e owns rows [seg_start[e], seg_start[e+1]), each segment start is a multiple of 256
 *     (one CTA-pair MMA tile; MOE_ROW_ALIGN),
 *     rows [seg_start[e] + kept[e], seg_start[e+1]) are padding.
 *     rows_cap >= moe_rows_cap(T, k, E, capacity).
 *   - constraints: d % 64 == 0, h % 64 == 0, 1 <= k <= 8, k <= E, E <= 1024.
 */
#ifndef MOE_B200_H
#define MOE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOE_DTYPE_F32 0
#define MOE_DTYPE_BF16 1

#define MOE_SCORE_TOPK_SOFTMAX 0 /* NaiveGate / GShardGate: softmax over the k selected logits      */
#define MOE_SCORE_FULL_SOFTMAX 1 /* SwitchGate: softmax over all experts, score = prob of selection */

#define MOE_AUX_NONE 0   /* no load-balancing loss (NaiveGate)                                       */
#define MOE_AUX_SWITCH 1 /* E * sum_e f_e P_e, f_e = share of KEPT pairs, P_e = mean softmax prob      */
#define MOE_AUX_GSHARD 2 /* mean_e(c_e m_e) * E^2, c_e = share of routed pairs, m_e = mean softmax prob */

#define MOE_TOKEN_TILE 64 /* tokens per routing tile; ntiles = ceil(T / 64) */
#define MOE_ROW_ALIGN 256  /* segment alignment of the packed buffers = rows of one CTA-pair MMA tile */

/* grouped GEMM ops (moe_grouped_gemm) */
#define MOE_GEMM_FC1 0   /* U = A W^T + b: out0 = gelu_erf'(U), out1 = gelu_erf(U)   A[rows,K] B[E,N,K] */
#define MOE_GEMM_FC2 1   /* out0 = A W^T + b                            A[rows,K] B[E,N,K]            */
#define MOE_GEMM_DGELU 2 /* out0 = (A W) * aux  (aux = FC1's out0)          A[rows,K] B[E,K,N] aux[rows,N]; out1: optional slab column sums */
#define MOE_GEMM_DGRAD 3 /* out0 = A W                                  A[rows,K] B[E,K,N]            */
/* (the backward contractions read the forward weights as they are — W2 [E,d,h] is [E,K,N] for DGELU, W1 [E,h,d] for
 *  DGRAD — through MN-major UMMA descriptors: one bf16 copy per weight matrix, no transposes) */
#define MOE_GEMM_WGRAD 4 /* out0[e] (fp32 [E,M,N]) = A_e^T B_e          A[rows,M] B[rows,N]           */
#define MOE_GEMM_WGRAD_T 5 /* out0[e] (fp32 [E,N,M]) = (A_e^T B_e)^T = B_e^T A_e, same tiles as WGRAD, transposed store:
                              lets the wide dimension be M (256-row tiles) whatever the parameter's layout */

const char *moe_last_error(void);
int moe_version(void);

/* rows the packed buffers must hold: min(T*k, E*capacity) rounded up to 256, + 256*E */
int64_t moe_rows_cap(int64_t T, int k, int E, int64_t capacity);

/* ---- gate: replaces NaiveGate's nn.Linear + torch.topk + F.softmax and fmoe_cuda.expert_count.
 * logits[T,E] fp32, idx[T,k] i32, score[T,k] fp32, tile_hist[E,ntiles] i32, tile_psum[E,ntiles] fp32 (only if
 * want_psum); ntiles = ceil(T / MOE_TOKEN_TILE).
 * Two arithmetic paths, both with routing integers (idx and everything derived from it) bit-identical to
 * oracle/gate_ref.c:
 *   workspace == NULL (or fp32 activations, or E > 64): fp32 FMA chains on the CUDA cores in LOGIT ORDER v1 — the
 *     logits themselves are bit-identical to the oracle's;
 *   workspace != NULL (moe_gate_fwd_workspace_bytes(d, E) bytes), bf16 activations, E <= 64: the projection runs on
 *     the tensor cores (mma.sync, Wg split into two bf16 planes); a token whose top-(k+1) logits are further apart
 *     than a rigorous bound on the difference to the oracle's logits keeps the tensor-core logits (certified: the
 *     oracle cannot route it differently), every other token is recomputed in LOGIT ORDER v1.  Emitted logits are
 *     within kappa(d) * ||x_t|| * max_e ||Wg_e|| of the oracle's (kappa(384) = 1.6e-5), exact for recomputed tokens.
 * token_mask (nullable, uint8 [T]): the token-skip mask of the residual-MoE block
 * (/root/reference/models/resMoE.py:126-145, `tk = x * mask[:, :, 1]`).  A token with mask 0 is not routed at
 * all: idx = -1 and score = 0 in every slot, no histogram entry (it takes no capacity), no share in psum; every
 * downstream kernel treats idx / pos = -1 as "no row" (zero output, zero gradient). */
int moe_gate_fwd(const void *x, int x_dtype, const float *Wg, const float *bg /* nullable */,
                 const float *noise /* nullable [T,E], added to the logits (SwitchGate jitter) */,
                 const uint8_t *token_mask /* nullable [T] */, int64_t T, int d, int E,
                 int k, int score_mode, int want_psum, float *logits, int32_t *idx, float *score, int32_t *tile_hist,
                 float *tile_psum, void *workspace /* nullable, see above */, void *stream);
size_t moe_gate_fwd_workspace_bytes(int d, int E);

/* ---- scan: replaces torch.cumsum + .item() + limit_by_capacity (no host sync).
 * tile_base[E,ntiles], count[E], kept[E] = min(count, capacity), seg_start[E+1],
 * tile_expert[max_mtiles] (expert of each 256-row tile, -1 past the end), num_mtiles[1],
 * psum[E] = sum over tiles of tile_psum (nullable together with tile_psum).  With aux_mode != 0 the
 * load-balancing loss of the gate and its gradient w.r.t. psum are produced here too, so that no
 * framework-level reduction kernels run per layer. */
int moe_route_scan(const int32_t *tile_hist, const float *tile_psum, int ntiles, int E, int64_t capacity,
                   int32_t *tile_base, int32_t *count, int32_t *kept, int32_t *seg_start, int32_t *tile_expert,
                   int32_t *num_mtiles, int max_mtiles, float *psum,
                   int aux_mode /* MOE_AUX_* */, int64_t T, int k, float *aux_loss /* [1] */,
                   float *aux_coef /* [E] = d aux_loss / d psum */,
                   int64_t slab_rows /* 0: packed segments; > 0: every expert gets a fixed slab of this many rows
                                        (expert-parallel send layout, multiple of MOE_ROW_ALIGN, >= capacity) */,
                   void *stream);

/* ---- expert parallelism, receive side: replaces fmoe_cuda.expert_exchange + the receive half of
 * global_scatter / global_gather.  The all-to-all itself is issued by the host (NCCL) on fixed slabs
 * [W, E_local, slab_rows, d]; kept_recv[W, E_local] are the live rows of each received slab.
 * moe_ep_tables: packed layout of the local experts (sources in rank order inside each segment):
 *   slab_dst[W,E_local] first packed row of each slab, kept_local[E_local], seg_start[E_local+1],
 *   tile_expert[max_mtiles], num_mtiles[1].
 * moe_ep_repack: to_packed = 1 slabs -> packed rows (pad rows zeroed); to_packed = 0 packed rows -> slabs. */
int moe_ep_tables(const int32_t *kept_recv, int W, int E_local, int32_t *slab_dst, int32_t *kept_local, int32_t *seg_start,
                  int32_t *tile_expert, int32_t *num_mtiles, int max_mtiles, void *stream);
int moe_ep_repack(const void *src, void *dst, const int32_t *kept_recv, const int32_t *slab_dst, const int32_t *seg_start,
                  const int32_t *kept_local, int W, int E_local, int64_t slab_rows, int d, int to_packed, void *stream);

/* ---- expert parallelism over NVLink / NVSwitch peer memory: replaces fmoe_cuda.expert_exchange + global_scatter /
 * global_gather (NCCL grouped send/recv around a host round trip for the counts) AND the NCCL all_to_all_single +
 * moe_ep_repack path above.  One process per GPU, all ranks of the expert-parallel group on one NVLink domain (W <= 8).
 * Every rank allocates one symmetric heap (moe_ep_heap_alloc: cudaMalloc, zero-filled, + a 64-byte CUDA IPC handle that
 * the host exchanges by any means, e.g. torch.distributed.all_gather_object), opens its peers' heaps
 * (moe_ep_heap_open) and carves the same regions out of each: flags int32[W], kept_all int32[W][E], and the packed row
 * buffers xbuf / ybuf / dybuf / dxbuf [rows_per_rank, d] bf16 of its LOCAL experts.  `void *const *x_peers` arguments
 * are HOST arrays of W device pointers, entry r = that region in rank r's heap (entry `rank` = the local one).
 * "Global rows": row index g = owner_rank * rows_per_rank + row inside the owner's buffer; `pos` holds global rows.
 *   moe_ep_exchange_counts  writes this rank's kept[E] (moe_route_scan) into every rank's kept_all, meets the other ranks
 *                           (device-side barrier through `flags`, epoch counter in `epoch`, both graph-replay safe), and
 *                           lays out the packed buffers: dst_row[E] = global row of this rank's first pair of each expert
 *                           (inside an expert segment the sources follow each other in rank order), and for the local
 *                           experts kept_local[E_local], seg_start[E_local + 1], tile_expert[max_mtiles], num_mtiles[1]
 *                           exactly as moe_route_scan produces them on one GPU.
 *   moe_dispatch_fwd_peer   moe_dispatch_fwd with the kept rows written straight into the owners' packed segments
 *                           (seg_start := dst_row); zeroes the pad rows of the local xbuf.
 *   moe_ep_barrier          all ranks' previous writes to this rank's heap are complete and visible afterwards.
 *   moe_combine_fwd_peer / moe_combine_bwd_peer / moe_gate_dispatch_bwd_peer   the single-GPU kernels reading Y / dX rows
 *                           from, and writing dY rows to, the owners' buffers in place.
 * status[1] (device int32, caller-zeroed): set to 1 if a barrier spin exceeded its bound (a rank is missing), 2 if a
 * packed layout does not fit rows_per_rank; never hangs the GPU. */
#define MOE_IPC_HANDLE_BYTES 64
int moe_ep_heap_alloc(size_t bytes, void **ptr, void *handle_out /* host, MOE_IPC_HANDLE_BYTES */);
int moe_ep_heap_open(const void *handle /* host, MOE_IPC_HANDLE_BYTES */, void **ptr);
int moe_ep_heap_close(void *ptr);
int moe_ep_heap_free(void *ptr);
int moe_ep_barrier(void *const *flags, int32_t *epoch, int rank, int W, int32_t *status, void *stream);
int moe_ep_exchange_counts(const int32_t *kept, void *const *kept_all, void *const *flags, int32_t *epoch, int rank, int W,
                           int E_local, int64_t rows_per_rank, int32_t *dst_row, int32_t *kept_local, int32_t *seg_start,
                           int32_t *tile_expert, int32_t *num_mtiles, int max_mtiles, int32_t *status, void *stream);
int moe_dispatch_fwd_peer(const void *x, int x_dtype, const int32_t *idx, const int32_t *tile_base, const int32_t *dst_row, int64_t T,
                          int d, int E, int k, int64_t capacity, void *const *xbuf_peers, int rank, int W, int64_t rows_per_rank,
                          const int32_t *seg_start_local, const int32_t *kept_local, int32_t *pos, void *stream);
int moe_combine_fwd_peer(void *const *ybuf_peers, int rank, int W, int64_t rows_per_rank, const int32_t *pos, const float *score,
                         int64_t T, int d, int k, void *out, int out_dtype, void *stream);
int moe_combine_bwd_peer(const void *dy, int dy_dtype, void *const *ybuf_peers, void *const *dybuf_peers, int rank, int W,
                         int64_t rows_per_rank, const int32_t *pos, const float *score, const int32_t *seg_start_local,
                         const int32_t *kept_local, int E_local, int64_t T, int d, int k, float *dscore, void *stream);
int moe_gate_dispatch_bwd_peer(void *const *dxbuf_peers, int rank, int W, int64_t rows_per_rank, const int32_t *pos, const float *logits,
                               const int32_t *idx, const float *score, const float *dscore, const float *dpsum, const float *Wg,
                               int64_t T, int d, int E, int k, int score_mode, float *dlogits, void *dx, int dx_dtype, void *stream);

/* ---- dispatch: replaces fmoe_cuda.assign_pos + MOEScatter (deterministic, token order).
 * pos[T,k] row of each (token,slot) or -1 if dropped; row_src[rows_cap] flattened pair index of
 * each row (-1 on padding); xbuf[rows_cap,d] bf16 with padding rows zeroed. */
int moe_dispatch_fwd(const void *x, int x_dtype, const int32_t *idx, const int32_t *tile_base, const int32_t *seg_start,
                     const int32_t *kept, int64_t T, int d, int E, int k, int64_t capacity, int32_t *pos,
                     int32_t *row_src, void *xbuf, void *stream);

/* ---- expert FFN forward: replaces _Expert.forward = fmoe_cuda.linear_forward x2 + GELU.
 * W1b[E,h,d], W2b[E,d,h] bf16 copies of the fp32 parameters; b1[E,h], b2[E,d] fp32.
 * With U = X W1^T + b1 (fp32, never stored): writes G = gelu_erf'(U) — all that backward needs of U —,
 * H = gelu_erf(U) and Y = H W2^T + b2, all bf16 [rows_cap, .]. */
int moe_expert_ffn_fwd(const void *xbuf, const void *W1b, const float *b1, const void *W2b, const float *b2,
                       const int32_t *tile_expert, const int32_t *num_mtiles, int64_t rows_cap, int d, int h, int E,
                       void *G, void *H, void *Y, void *stream);

/* ---- combine: replaces MOEGather + torch.bmm(gate_score, expert_out). out[T,d] in out_dtype. */
int moe_combine_fwd(const void *ybuf, const int32_t *pos, const float *score, int64_t T, int d, int k, void *out,
                    int out_dtype, void *stream);

/* ---- backward of combine: dybuf[rows_cap,d] bf16 (padding zeroed), dscore[T,k] fp32. */
int moe_combine_bwd(const void *dy, int dy_dtype, const void *ybuf, const int32_t *pos, const float *score,
                    const int32_t *seg_start, const int32_t *kept, int64_t T, int d, int k, int E, void *dybuf,
                    float *dscore, void *stream);

/* ---- expert FFN backward: replaces fmoe_cuda.linear_backward x2 + GELU' + column_reduce.
 * W1b[E,h,d] and W2b[E,d,h] are the SAME bf16 weight copies the forward reads (moe_cast_bf16): the two backward
 * contractions read them MN-major.
 * dU[rows_cap,h] and dxbuf[rows_cap,d] are bf16 outputs (dU doubles as workspace);
 * dW1[E,h,d], db1[E,h], dW2[E,d,h], db2[E,d] are fp32 and are overwritten. */
size_t moe_expert_ffn_bwd_workspace_bytes(int64_t rows_cap, int d, int h, int E);
int moe_expert_ffn_bwd(const void *dybuf, const void *xbuf, const void *G, const void *H, const void *W1b,
                       const void *W2b, const int32_t *tile_expert, const int32_t *num_mtiles,
                       const int32_t *seg_start, int64_t rows_cap, int d, int h, int E, void *dU, void *dxbuf,
                       float *dW1, float *db1, float *dW2, float *db2,
                       void *workspace /* moe_expert_ffn_bwd_workspace_bytes(rows_cap, d, h, E) bytes */, void *stream);

/* One scratch size that covers every `workspace` argument of one layer's forward + backward (the expert-FFN backward
 * and the gate weight gradient never hold theirs at the same time): max over moe_expert_ffn_bwd_workspace_bytes and
 * moe_gate_wgrad_workspace_bytes at rows_cap = moe_rows_cap(T, k, E, capacity). */
size_t moe_workspace_bytes(int64_t T, int d, int h, int E, int k, int64_t capacity);

/* ---- gate backward: dlogits[T,E] from dscore[T,k] and (nullable) dpsum[E]. */
int moe_gate_bwd(const float *logits, const int32_t *idx, const float *score, const float *dscore, const float *dpsum,
                 int64_t T, int E, int k, int score_mode, float *dlogits, void *stream);

/* ---- dispatch backward (+ gate input gradient): dx[t] = sum_j dxbuf[pos[t,j]] + dlogits[t] Wg.
 * dxbuf or dlogits may be NULL (term skipped).  dense_dlogits = 0 means only the k selected
 * entries of each dlogits row are non-zero (NaiveGate without aux loss). */
int moe_dispatch_bwd(const void *dxbuf, const int32_t *pos, const float *dlogits, const int32_t *idx, const float *Wg,
                     int64_t T, int d, int E, int k, int dense_dlogits, void *dx, int dx_dtype, void *stream);

/* ---- gate backward + dispatch backward in one pass (what the layer's autograd node calls):
 * dlogits[T,E] (output, consumed by moe_gate_wgrad) from dscore[T,k] and the nullable dpsum[E];
 * dx[t] = sum_j dxbuf[pos[t,j]] + dlogits[t] Wg.  dxbuf may be NULL. */
int moe_gate_dispatch_bwd(const void *dxbuf, const int32_t *pos, const float *logits, const int32_t *idx, const float *score,
                          const float *dscore, const float *dpsum, const float *Wg, int64_t T, int d, int E, int k,
                          int score_mode, float *dlogits, void *dx, int dx_dtype, void *stream);

/* ---- gate weight gradient: dWg[E,d] = dlogits^T x, dbg[E] (nullable) = colsum(dlogits).
 * workspace: moe_gate_wgrad_workspace_bytes(T, d, E) bytes. */
size_t moe_gate_wgrad_workspace_bytes(int64_t T, int d, int E);
int moe_gate_wgrad(const float *dlogits, const void *x, int x_dtype, int64_t T, int d, int E, void *workspace,
                   float *dWg, float *dbg, void *stream);

/* ---- block-level fusion around the layer (reference models/vision_transformer.py:319-322):
 * x_out = x_in + delta (delta nullable: x_out untouched), n = LayerNorm(x_out) * gamma + beta.
 * The residual stream x is fp32; delta and n are fp32 or bf16; mean / rstd [T] fp32 are saved for backward.
 * Backward: dx_in = dx_out (nullable) + LN'(dn), d_delta (nullable) = dx_in in delta's dtype, dgamma / dbeta [d]
 * through a two-stage deterministic reduction (workspace: moe_addln_bwd_workspace_bytes).  d % 4 == 0, d <= 1024. */
int moe_addln_fwd(const float *x_in, const void *delta, int delta_dtype, const float *gamma, const float *beta, float eps,
                  int64_t T, int d, float *x_out, void *n, int n_dtype, float *mean, float *rstd, void *stream);
size_t moe_addln_bwd_workspace_bytes(int64_t T, int d);
int moe_addln_bwd(const void *dn, int n_dtype, const float *dx_out, const float *x, const float *mean, const float *rstd,
                  const float *gamma, int64_t T, int d, float *dx_in, void *d_delta, int delta_dtype, void *workspace,
                  float *dgamma, float *dbeta, void *stream);

/* out[cols] (fp32) = column sums of buf[rows, cols] (fp32 or bf16): bias gradient of the block's dense projections,
 * two deterministic stages (workspace: moe_colsum_workspace_bytes).  cols % 8 == 0. */
size_t moe_colsum_workspace_bytes(int64_t rows, int cols);
int moe_colsum(const void *buf, int dtype, int64_t rows, int cols, void *workspace, float *out, void *stream);

/* ---- utilities */
int moe_cast_bf16(const float *src, void *dst, int64_t n /* % 8 == 0 */, void *stream);
/* two ranges in one launch: the per-step cast of a layer's two expert weight matrices (fp32 master -> bf16 operand copies) */
int moe_cast_bf16_pair(const float *src0, void *dst0, int64_t n0, const float *src1, void *dst1, int64_t n1, void *stream);
/* out[E,cols] = per-segment column sums of the packed bf16 buffer buf[rows_cap,cols] (bias gradients;
 * replaces fmoe_cuda's column_reduce).  Two deterministic stages through `workspace`. */
size_t moe_segment_colsum_workspace_bytes(int64_t rows_cap, int cols);
int moe_segment_colsum(const void *buf, const int32_t *seg_start, int64_t rows_cap, int E, int cols, void *workspace,
                       float *out, void *stream);

/* ---- the grouped tcgen05 GEMM itself (building block of the two FFN entry points; exported so
 * each contraction can be tested and timed on its own).  See MOE_GEMM_* for operand shapes. */
/* MOE_GEMM_DGELU only: when `out1` is not NULL it receives the column sums of every 32-row slab of out0
 * ([rows_cap / 32, N] fp32, moe_slab_colsum_bytes(rows_cap, N) bytes; slabs of tiles that are not live are not written).
 * moe_slab_colsum_final adds the slabs of each expert segment in row order: out[E, cols] = the bias gradient db1 of the
 * first expert projection, without a second pass over dU (replaces fmoe_cuda's column_reduce for that tensor). */
size_t moe_slab_colsum_bytes(int64_t rows_cap, int cols);
int moe_slab_colsum_final(const float *part, const int32_t *seg_start, int E, int cols, float *out, void *stream);
/* MOE_GEMM_WGRAD / MOE_GEMM_WGRAD_T only: bytes of the optional flag workspace passed as `aux` (int32, zero-filled once by
 * the caller; every launch leaves it zero again; one workspace per stream).  With it the K range of an output tile may be
 * computed in several pieces on several CTA pairs whenever whole tiles would not fill the grid evenly: stream-K (every
 * pair takes an equal share of the linearised (tile, k-block) space; config 2: 96 tiles on 74 pairs) or S equal parts per
 * tile (S = 2, or pairs / tiles when there are fewer tiles than half the pairs — few local experts under expert
 * parallelism).  The piece that starts a tile's K range stores, the others are added in order, chained through the flags:
 * bit-reproducible.  All CTAs of the launch must be resident at once (one per SM): do not run it next to another kernel
 * that occupies whole SMs for its entire duration. */
size_t moe_wgrad_flags_bytes(int E, int M, int N);
int moe_grouped_gemm(int op, const void *A, const void *B, void *out0, void *out1, const float *bias, const void *aux,
                     const int32_t *tile_expert, const int32_t *num_mtiles, const int32_t *seg_start, int64_t rows_cap,
                     int E, int M, int N, int K, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MOE_B200_H */
