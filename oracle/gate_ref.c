/*
 * oracle/gate_ref.c — CPU restatement of the MoE gate + routing arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
 * product path (slim-switch-moe-vit_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * PARITY UNPINNED: the reference (d0-rb/slim-switch-moe-vit) gets this arithmetic from the
 * third-party package FastMoE (`from fmoe import FMoETransformerMLP`,
 * /root/reference/models/resMoE.py:6), which is neither vendored, version-pinned nor
 * installed, and the reference ships no tests / golden vectors for it.  This file restates
 * FastMoE's published algorithm (fmoe/gates/naive_gate.py: Linear -> topk -> softmax over the
 * selected logits; fmoe/gates/switch_gate.py: full softmax -> top-1 -> capacity limit;
 * fmoe/functions.py: count_by_gate / assign_pos) with the two places where upstream is
 * non-deterministic (torch.topk tie order, atomics in assign_pos / prune_gate_by_capacity)
 * replaced by the canonical choice documented in DESIGN.md:
 *   - ties in top-k resolve to the LOWEST expert index;
 *   - (token,slot) pairs are ranked inside their expert in ascending flattened index t*k+j
 *     (the Switch-Transformer cumsum order), and the first C of them are kept.
 * The reference call sites this anchors to: models/resMoE.py:15-29 (constructor mapping),
 * models/vision_transformer.py:319-322 (caller), models/resmoe_flop_hook.py:7 (gate = Linear).
 *
 * LOGIT ORDER v1 (the fixed reduction order the CUDA gate kernel also uses, so that routing
 * integers are bit-exact between this file and the GPU):
 *   lane(i) = (i / 4) % 32 for feature index i;
 *   partial[l] = fma-chain over { i : lane(i) == l } in ascending i, starting from +0.0f;
 *   for off in 16,8,4,2,1: partial[l] = partial[l] + partial[l ^ off]   (all l at once);
 *   logit = partial[0] + bias.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* logits[T,E] from x[T,d] (fp32 values; bf16 inputs are widened exactly by the caller),
 * Wg[E,d], bg[E] (may be NULL), optional additive noise[T,E] (SwitchGate jitter). */
void moe_oracle_gate_logits(const float *x, int64_t T, int d, const float *Wg, const float *bg,
                            const float *noise, int E, float *logits)
{
    for (int64_t t = 0; t < T; ++t) {
        const float *xr = x + t * (int64_t)d;
        for (int e = 0; e < E; ++e) {
            const float *wr = Wg + (int64_t)e * d;
            float part[32];
            for (int l = 0; l < 32; ++l) part[l] = 0.0f;
            for (int i = 0; i < d; ++i) {
                int l = (i >> 2) & 31;
                part[l] = fmaf(xr[i], wr[i], part[l]);
            }
            for (int off = 16; off >= 1; off >>= 1) {
                float nxt[32];
                for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
                memcpy(part, nxt, sizeof(part));
            }
            float v = part[0] + (bg ? bg[e] : 0.0f);
            if (noise) v += noise[t * E + e];
            logits[t * E + e] = v;
        }
    }
}

/*
 * Canonical top-k + scores + capacity positions.
 *   score_mode 0: softmax over the k selected logits (FastMoE NaiveGate / GShardGate)
 *   score_mode 1: full-softmax probability of the selected expert (FastMoE SwitchGate)
 *   capacity    : per-expert row limit C (pairs with rank >= C are dropped, pos = -1)
 *   align       : every expert's segment in the packed buffer starts at a multiple of `align`
 *   token_mask  : NULL, or uint8[T]: a token with mask 0 is skipped (the residual-MoE block's token gate,
 *                 /root/reference/models/resMoE.py:126-145): idx = -1, score = 0, pos = -1 in every slot, it is
 *                 not counted, takes no capacity and adds nothing to psum
 * Outputs: idx[T*k] i32, score[T*k] f32, count[E], kept[E], seg_start[E+1], pos[T*k] (-1 = dropped),
 *          psum[E] = sum_t softmax(logits[t,:])[e] (double accumulated; may be NULL).
 */
void moe_oracle_route(const float *logits, int64_t T, int E, int k, int score_mode,
                      int64_t capacity, int align, int32_t *idx, float *score, int32_t *count,
                      int32_t *kept, int32_t *seg_start, int32_t *pos, double *psum, const uint8_t *token_mask)
{
    int64_t *rank = (int64_t *)calloc((size_t)E, sizeof(int64_t));
    for (int e = 0; e < E; ++e) count[e] = 0;
    if (psum) for (int e = 0; e < E; ++e) psum[e] = 0.0;
    int32_t *rk = (int32_t *)malloc(sizeof(int32_t) * (size_t)(T * k));

    for (int64_t t = 0; t < T; ++t) {
        const float *lr = logits + t * E;
        if (token_mask && token_mask[t] == 0) {
            for (int j = 0; j < k; ++j) { idx[t * k + j] = -1; score[t * k + j] = 0.0f; rk[t * k + j] = 0; }
            continue;
        }
        int picked[64];
        float pv[64];
        for (int j = 0; j < k; ++j) {
            int besti = -1;
            float best = 0.0f;
            for (int e = 0; e < E; ++e) {
                int used = 0;
                for (int q = 0; q < j; ++q) used |= (picked[q] == e);
                if (used) continue;
                if (besti < 0 || lr[e] > best) { besti = e; best = lr[e]; }
            }
            picked[j] = besti;
            pv[j] = best;
            idx[t * k + j] = besti;
        }
        float m = pv[0];
        if (score_mode == 0) {
            float w[64], s = 0.0f;
            for (int j = 0; j < k; ++j) { w[j] = expf(pv[j] - m); s += w[j]; }
            for (int j = 0; j < k; ++j) score[t * k + j] = w[j] / s;
        }
        if (score_mode == 1 || psum) {
            float z = 0.0f;
            for (int e = 0; e < E; ++e) z += expf(lr[e] - m);
            if (score_mode == 1)
                for (int j = 0; j < k; ++j) score[t * k + j] = expf(pv[j] - m) / z;
            if (psum)
                for (int e = 0; e < E; ++e) psum[e] += (double)(expf(lr[e] - m) / z);
        }
        for (int j = 0; j < k; ++j) {
            int e = picked[j];
            rk[t * k + j] = (int32_t)rank[e];
            rank[e] += 1;
            count[e] += 1;
        }
    }
    int64_t start = 0;
    for (int e = 0; e < E; ++e) {
        int64_t kp = count[e] < capacity ? count[e] : capacity;
        kept[e] = (int32_t)kp;
        seg_start[e] = (int32_t)start;
        start += (kp + align - 1) / align * align;
    }
    seg_start[E] = (int32_t)start;
    for (int64_t i = 0; i < T * k; ++i) {
        int e = idx[i];
        pos[i] = (e >= 0 && rk[i] < capacity) ? seg_start[e] + rk[i] : -1;
    }
    free(rk);
    free(rank);
}
