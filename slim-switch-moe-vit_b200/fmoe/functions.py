"""Autograd functions of the B200 MoE layer.

Mirrors the role of FastMoE's `fmoe/functions.py` (prepare_forward / MOEScatter / MOELinear /
MOEGather — upstream, un-vendored; reached from /root/reference/models/resMoE.py:27-29), but the
whole layer is ONE autograd node whose forward and backward are sequences of C-ABI kernel calls on
the current CUDA stream, with no host synchronisation anywhere (upstream syncs twice per layer).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _cabi as C


@dataclass(frozen=True)
class RouteSpec:
    """What the gate asks of the fused path."""
    top_k: int
    score_mode: int          # C.SCORE_TOPK_SOFTMAX | C.SCORE_FULL_SOFTMAX
    capacity: int            # rows per expert (>= T*k means unlimited)
    aux_mode: int            # C.AUX_NONE | C.AUX_SWITCH | C.AUX_GSHARD: load-balancing loss computed by the scan kernel

    @property
    def want_psum(self) -> bool:   # sum_t softmax(logits)[e] is needed exactly when there is an aux loss
        return int(self.aux_mode) != C.AUX_NONE


class Bf16WeightCache:
    """bf16 operand copies of the fp32 expert weights — W1b [E,h,d], W2b [E,d,h].  One copy per matrix: the forward
    contractions read them K-major, dgelu / dgrad read the same buffers MN-major (csrc/gemm.cuh), so no transposes.

    A training forward (`fresh=True`: autograd is recording and the weights require grad) ALWAYS re-casts: the
    weights move once per optimizer step anyway, and neither `p.data.add_` (many timm / apex optimizers) nor a CUDA-graph
    replay of the optimizer bumps `p._version`, so no host-side key can be trusted there.  Inference forwards
    (`torch.no_grad()`, frozen weights) reuse the copies keyed on (data_ptr, _version); `invalidate()` drops them
    (call it after writing weights through `.data` outside of training).  A training forward leaves NO key behind: the
    optimizer step that follows it may be one of those invisible writes, so the first inference forward after training
    (the reference's `evaluate` after `train_one_epoch`, main.py:930-940) casts once more and only then starts reusing."""

    def __init__(self):
        self._key = None
        self._val = None
        self._graph_seen = False   # a CUDA graph that updates the weights may be replayed behind Python's back: never trust the key again

    def invalidate(self):
        self._key, self._val = None, None

    def get(self, W1: torch.Tensor, W2: torch.Tensor, fresh: bool = False):
        key = (W1.data_ptr(), W1._version, W2.data_ptr(), W2._version, W1.device)
        capturing = torch.cuda.is_current_stream_capturing()
        self._graph_seen = self._graph_seen or capturing
        if fresh or self._graph_seen or key != self._key:
            bf = torch.bfloat16
            E, h, d = W1.shape
            W1b, W2b = torch.empty((E, h, d), dtype=bf, device=W1.device), torch.empty((E, d, h), dtype=bf, device=W1.device)
            st = C.stream_ptr()
            C.call("moe_cast_bf16_pair", C.ptr(W1.detach()), C.ptr(W1b), W1.numel(), C.ptr(W2.detach()), C.ptr(W2b), W2.numel(), st)
            if capturing:
                return W1b, W2b      # copies live in the graph's pool and are refreshed by every replay: not cached
            self._key, self._val = (None if fresh else key), (W1b, W2b)
        return self._val

    def __deepcopy__(self, memo):  # ModelEma deep-copies the model (reference main.py:602-607)
        return Bf16WeightCache()


def _as_kernel_input(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise C.MoeB200Error("fmoe (B200) has no CPU path: input must be a CUDA tensor")
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()  # fp16 activations (reference engine.py:52 autocast) are widened, never narrowed
    return t.contiguous()


# Gate arithmetic: False (default) = tensor-core projection with certified routing wherever the kernel supports it (bf16
# activations, E <= 64); True = always the CUDA-core kernel whose logits are bit-identical to the CPU oracle's (LOGIT
# ORDER v1).  The routing integers are bit-identical to the oracle's either way.
GATE_EXACT_LOGITS = False


def _gate_workspace(x, E):
    if GATE_EXACT_LOGITS or x.dtype != torch.bfloat16 or E > 64:
        return None
    return torch.empty(int(C.lib.moe_gate_fwd_workspace_bytes(x.shape[1], E)), dtype=torch.uint8, device=x.device)


def _i32(n, dev):
    return torch.empty(n, dtype=torch.int32, device=dev)


def _f32(n, dev):
    return torch.empty(n, dtype=torch.float32, device=dev)


def route(x, Wg, bg, spec: RouteSpec, noise=None, slab_rows: int = 0, token_mask=None):
    """gate + scan + dispatch.  Returns a dict of device tensors (no host sync).
    slab_rows > 0 selects the expert-parallel send layout: expert e owns rows [e*slab_rows, (e+1)*slab_rows).
    token_mask (uint8 [T], optional): tokens with mask 0 are not routed (idx = pos = -1, no capacity taken)."""
    T, d = x.shape
    E = Wg.shape[0]
    k = spec.top_k
    dev = x.device
    st = C.stream_ptr()
    ntiles = (T + C.TOKEN_TILE - 1) // C.TOKEN_TILE
    rows_cap = E * slab_rows if slab_rows else C.rows_cap(T, k, E, spec.capacity)
    max_mtiles = rows_cap // C.ROW_ALIGN
    r = dict(
        logits=_f32((T, E), dev), idx=_i32((T, k), dev), score=_f32((T, k), dev),
        tile_hist=_i32((E, ntiles), dev), tile_base=_i32((E, ntiles), dev),
        count=_i32(E, dev), kept=_i32(E, dev), seg_start=_i32(E + 1, dev),
        tile_expert=_i32(max_mtiles, dev), num_mtiles=_i32(1, dev),
        pos=_i32((T, k), dev), row_src=_i32(rows_cap, dev),
        xbuf=torch.empty((rows_cap, d), dtype=torch.bfloat16, device=dev),
        rows_cap=rows_cap,
    )
    tile_psum = _f32((E, ntiles), dev) if spec.want_psum else None
    r["psum"] = _f32(E, dev) if spec.want_psum else None
    r["aux_loss"] = _f32(1, dev) if spec.want_psum else None
    r["aux_coef"] = _f32(E, dev) if spec.want_psum else None
    gws = _gate_workspace(x, E)
    C.call("moe_gate_fwd", C.ptr(x), C.dtype_code(x), C.ptr(Wg), C.ptr(bg), C.ptr(noise), C.ptr(token_mask), T, d, E, k,
           spec.score_mode, int(spec.want_psum), C.ptr(r["logits"]), C.ptr(r["idx"]), C.ptr(r["score"]),
           C.ptr(r["tile_hist"]), C.ptr(tile_psum), C.ptr(gws), st)
    C.call("moe_route_scan", C.ptr(r["tile_hist"]), C.ptr(tile_psum), ntiles, E, spec.capacity,
           C.ptr(r["tile_base"]), C.ptr(r["count"]), C.ptr(r["kept"]), C.ptr(r["seg_start"]),
           C.ptr(r["tile_expert"]), C.ptr(r["num_mtiles"]), max_mtiles, C.ptr(r["psum"]), int(spec.aux_mode), T, k,
           C.ptr(r["aux_loss"]), C.ptr(r["aux_coef"]), slab_rows, st)
    C.call("moe_dispatch_fwd", C.ptr(x), C.dtype_code(x), C.ptr(r["idx"]), C.ptr(r["tile_base"]),
           C.ptr(r["seg_start"]), C.ptr(r["kept"]), T, d, E, k, spec.capacity, C.ptr(r["pos"]),
           C.ptr(r["row_src"]), C.ptr(r["xbuf"]), st)
    return r


class MoEFunction(torch.autograd.Function):
    """y, aux_loss, count, kept = MoE(x; Wg, bg, W1, b1, W2, b2).

    x [T,d] fp32|bf16; Wg [E,d], bg [E]|None, W1 [E,h,d], b1 [E,h], W2 [E,d,h], b2 [E,d] fp32.
    y has x's dtype.  aux_loss is the gate's load-balancing loss (fp32 scalar, differentiable w.r.t.
    x / Wg / bg; an empty tensor when spec.aux_mode is AUX_NONE).
    count/kept [E] int32 are the per-expert routed / kept pair counts (not differentiable).
    """

    @staticmethod
    def forward(ctx, x, Wg, bg, W1, b1, W2, b2, spec: RouteSpec, cache: Bf16WeightCache, noise, token_mask=None,
                fresh: bool = True, infer: bool = False):
        """infer=True: forward-only pass (torch.no_grad(): the reference's `evaluate`, /root/reference/engine.py:88-121) —
        fc1 runs its single-output epilogue and G = gelu'(U), which only backward reads, is neither computed nor written."""
        x = _as_kernel_input(x)
        T, d = x.shape
        E, h = W1.shape[0], W1.shape[1]
        k = spec.top_k
        dev = x.device
        st = C.stream_ptr()
        Wg_c, W1_c, W2_c = Wg.detach().contiguous(), W1.detach().contiguous(), W2.detach().contiguous()
        for w in (Wg_c, W1_c, W2_c, b1, b2):
            if w.dtype != torch.float32:
                raise C.MoeB200Error("expert / gate parameters must be fp32 (master weights)")
        bg_c = None if bg is None else bg.detach().contiguous()
        r = route(x, Wg_c, bg_c, spec, noise, token_mask=token_mask)
        rows_cap = r["rows_cap"]
        W1b, W2b = cache.get(W1_c, W2_c, fresh)
        G = None if infer else torch.empty((rows_cap, h), dtype=torch.bfloat16, device=dev)
        H = torch.empty((rows_cap, h), dtype=torch.bfloat16, device=dev)
        Y = torch.empty((rows_cap, d), dtype=torch.bfloat16, device=dev)
        # the two forward GEMMs (same kernels as the bundled moe_expert_ffn_fwd entry point)
        b1_c, b2_c = b1.detach().contiguous(), b2.detach().contiguous()
        te, nm = C.ptr(r["tile_expert"]), C.ptr(r["num_mtiles"])
        C.call("moe_grouped_gemm", C.GEMM_FC1, C.ptr(r["xbuf"]), C.ptr(W1b), C.ptr(G), C.ptr(H), C.ptr(b1_c), None,
               te, nm, None, rows_cap, E, 0, h, d, st, tag="gemm_fc1")
        C.call("moe_grouped_gemm", C.GEMM_FC2, C.ptr(H), C.ptr(W2b), C.ptr(Y), None, C.ptr(b2_c), None,
               te, nm, None, rows_cap, E, 0, d, h, st, tag="gemm_fc2")
        y = torch.empty_like(x)
        C.call("moe_combine_fwd", C.ptr(Y), C.ptr(r["pos"]), C.ptr(r["score"]), T, d, k, C.ptr(y),
               C.dtype_code(y), st)

        ctx.spec = spec
        ctx.has_bg = bg is not None
        ctx.rows_cap = rows_cap
        coef = r["aux_coef"] if spec.want_psum else torch.empty(0, dtype=torch.float32, device=dev)
        if infer:
            aux = r["aux_loss"].reshape(()) if spec.want_psum else torch.empty(0, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(y, aux, r["count"], r["kept"])
            return y, aux, r["count"], r["kept"]
        ctx.save_for_backward(x, Wg_c, r["logits"], r["idx"], r["score"], r["pos"], r["seg_start"], r["kept"],
                              r["tile_expert"], r["num_mtiles"], r["xbuf"], G, H, Y, W1b, W2b, coef)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(r["count"], r["kept"])
        if spec.want_psum:
            aux = r["aux_loss"].reshape(())
        else:
            aux = torch.empty(0, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(aux)
        return y, aux, r["count"], r["kept"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy, daux, _dcount, _dkept):
        (x, Wg, logits, idx, score, pos, seg_start, kept, tile_expert, num_mtiles, xbuf, G, H, Y, W1b,
         W2b, coef) = ctx.saved_tensors
        spec: RouteSpec = ctx.spec
        T, d = x.shape
        E, h = W1b.shape[0], W1b.shape[1]
        k = spec.top_k
        dev = x.device
        st = C.stream_ptr()
        rows_cap = ctx.rows_cap
        if dy is None:
            dy = torch.zeros_like(x)
        dy = _as_kernel_input(dy)
        # aux_loss = sum_e coef_e psum_e  =>  d aux / d psum = coef (the shares inside coef are integers)
        dpsum = (coef * daux.float()).contiguous() if (spec.want_psum and daux is not None) else None

        dybuf = torch.empty((rows_cap, d), dtype=torch.bfloat16, device=dev)
        dscore = _f32((T, k), dev)
        C.call("moe_combine_bwd", C.ptr(dy), C.dtype_code(dy), C.ptr(Y), C.ptr(pos), C.ptr(score),
               C.ptr(seg_start), C.ptr(kept), T, d, k, E, C.ptr(dybuf), C.ptr(dscore), st)

        dU = torch.empty((rows_cap, h), dtype=torch.bfloat16, device=dev)
        dxbuf = torch.empty((rows_cap, d), dtype=torch.bfloat16, device=dev)
        dW1, db1 = _f32((E, h, d), dev), _f32((E, h), dev)
        dW2, db2 = _f32((E, d, h), dev), _f32((E, d), dev)
        # same kernel sequence as the bundled moe_expert_ffn_bwd entry point
        te, nm, sg = C.ptr(tile_expert), C.ptr(num_mtiles), C.ptr(seg_start)
        # db1 = column sums of dU per expert: the dgelu epilogue leaves the sums of every 32-row slab behind
        slab_sums = torch.empty(C.lib.moe_slab_colsum_bytes(rows_cap, h) // 4, dtype=torch.float32, device=dev)
        C.call("moe_grouped_gemm", C.GEMM_DGELU, C.ptr(dybuf), C.ptr(W2b), C.ptr(dU), C.ptr(slab_sums), None, C.ptr(G),
               te, nm, None, rows_cap, E, 0, h, d, st, tag="gemm_dgelu")
        # dW2 = (H^T dY)^T: the wide dimension h is M (256-row tiles), the store is transposed
        wfl = C.ptr(C.wgrad_flags(E, h, d, dev))   # split-K flags: each tile's K range runs as two halves
        C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, C.ptr(H), C.ptr(dybuf), C.ptr(dW2), None, None, wfl,
               None, None, sg, rows_cap, E, h, d, 0, st, tag="gemm_wgrad2")
        C.call("moe_grouped_gemm", C.GEMM_WGRAD, C.ptr(dU), C.ptr(xbuf), C.ptr(dW1), None, None, wfl,
               None, None, sg, rows_cap, E, h, d, 0, st, tag="gemm_wgrad1")
        C.call("moe_grouped_gemm", C.GEMM_DGRAD, C.ptr(dU), C.ptr(W1b), C.ptr(dxbuf), None, None, None,
               te, nm, None, rows_cap, E, 0, d, h, st, tag="gemm_dgrad")
        cws = torch.empty(C.lib.moe_segment_colsum_workspace_bytes(rows_cap, d), dtype=torch.uint8, device=dev)
        C.call("moe_segment_colsum", C.ptr(dybuf), sg, rows_cap, E, d, C.ptr(cws), C.ptr(db2), st, tag="colsum_db2")
        C.call("moe_slab_colsum_final", C.ptr(slab_sums), sg, E, h, C.ptr(db1), st, tag="colsum_db1")
        # gate backward (dlogits) + un-permute of dX + dlogits Wg in one pass
        dlogits = _f32((T, E), dev)
        dx = torch.empty_like(x)
        C.call("moe_gate_dispatch_bwd", C.ptr(dxbuf), C.ptr(pos), C.ptr(logits), C.ptr(idx), C.ptr(score), C.ptr(dscore),
               C.ptr(dpsum), C.ptr(Wg), T, d, E, k, spec.score_mode, C.ptr(dlogits), C.ptr(dx), C.dtype_code(dx), st)
        ws = torch.empty(C.lib.moe_gate_wgrad_workspace_bytes(T, d, E), dtype=torch.uint8, device=dev)
        dWg = _f32((E, d), dev)
        dbg = _f32(E, dev) if ctx.has_bg else None
        C.call("moe_gate_wgrad", C.ptr(dlogits), C.ptr(x), C.dtype_code(x), T, d, E, C.ptr(ws), C.ptr(dWg),
               C.ptr(dbg), st)
        return dx, dWg, dbg, dW1, db1, dW2, db2, None, None, None, None, None, None


class SkipFill(torch.autograd.Function):
    """Token-skip semantics of the reference's residual-MoE block (/root/reference/models/resMoE.py:126-145):
    there a skipped token enters the layer as `x * 0`, so its output is the constant c = mlp(0) and the gradient
    with respect to its (zero) input row is J0^T dy, J0 = d mlp / d z at z = 0.  The fused layer never routes such
    tokens; this node restores both terms:  out = keep ? y : c,  d inp[skipped] = dy J0,  dc = sum over skipped dy."""

    @staticmethod
    def forward(ctx, y, inp, keep, c, J0):
        ctx.save_for_backward(keep, J0)
        ctx.inp_dtype = inp.dtype
        return torch.where(keep.bool().unsqueeze(1), y, c.to(y.dtype).unsqueeze(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        keep, J0 = ctx.saved_tensors
        kb = keep.bool().unsqueeze(1)
        dskip = torch.where(kb, torch.zeros((), dtype=torch.float32, device=dout.device), dout.float())   # rows of skipped tokens
        dy = torch.where(kb, dout, torch.zeros((), dtype=dout.dtype, device=dout.device))
        dinp = (dskip @ J0).to(ctx.inp_dtype) if ctx.needs_input_grad[1] else None
        dc = dskip.sum(0) if ctx.needs_input_grad[3] else None
        return dy, dinp, None, dc, None


def zero_token_path(gate_w, gate_b, W1, b1, W2, b2, top_k: int, score_mode: int):
    """c = mlp(0) [d] (differentiable w.r.t. the parameters) and J0 = d mlp(z) / d z at z = 0 [d_out, d_in] (detached)
    for one all-zero token — a handful of small fp32 ops on the expert(s) the gate bias selects; no host sync.
    Selection follows the layer's canonical rule (descending logit, ties -> lowest index)."""
    E, d = gate_w.shape
    bg = gate_b if gate_b is not None else torch.zeros(E, dtype=torch.float32, device=gate_w.device)
    vals, e_idx = torch.sort(bg.float(), descending=True, stable=True)
    vals, e_idx = vals[:top_k], e_idx[:top_k]
    if score_mode == C.SCORE_TOPK_SOFTMAX:
        s = torch.softmax(vals, dim=0)                                         # [k]
    else:
        p = torch.softmax(bg.float(), dim=0)
        s = p[e_idx]
    W1s, b1s, W2s, b2s = W1[e_idx], b1[e_idx], W2[e_idx], b2[e_idx]           # [k,h,d] [k,h] [k,d,h] [k,d]
    act = torch.nn.functional.gelu(b1s)                                        # exact erf
    E0 = torch.einsum("kdh,kh->kd", W2s, act) + b2s                            # expert outputs at z = 0
    c = (s.unsqueeze(1) * E0).sum(0)
    with torch.no_grad():
        u = b1s.float()
        dact = 0.5 * (1.0 + torch.erf(u * 0.7071067811865476)) + u * torch.exp(-0.5 * u * u) * 0.3989422804014327
        J0 = torch.einsum("k,kdh,khi->di", s, W2s.float() * dact.unsqueeze(1), W1s.float())   # sum_j s_j W2 diag(gelu') W1
        Wg_sel = gate_w.float()[e_idx]                                                       # [k, d_in]
        if score_mode == C.SCORE_TOPK_SOFTMAX:
            ds = s.unsqueeze(1) * (Wg_sel - (s.unsqueeze(1) * Wg_sel).sum(0, keepdim=True))  # d s_j / d z
        else:
            ds = s.unsqueeze(1) * (Wg_sel - (p.unsqueeze(1) * gate_w.float()).sum(0, keepdim=True))
        J0 = J0 + E0.float().t() @ ds
    return c, J0


class GateFunction(torch.autograd.Function):
    """Stand-alone gate (what `gate.forward(x)` returns in FastMoE): idx [T,k] int64, score [T,k]."""

    @staticmethod
    def forward(ctx, x, Wg, bg, spec: RouteSpec, noise):
        x = _as_kernel_input(x)
        T, d = x.shape
        E = Wg.shape[0]
        dev = x.device
        st = C.stream_ptr()
        ntiles = (T + C.TOKEN_TILE - 1) // C.TOKEN_TILE
        logits, idx, score = _f32((T, E), dev), _i32((T, spec.top_k), dev), _f32((T, spec.top_k), dev)
        tile_hist = _i32((E, ntiles), dev)
        tile_psum = _f32((E, ntiles), dev) if spec.want_psum else None
        Wg_c = Wg.detach().contiguous()
        bg_c = None if bg is None else bg.detach().contiguous()
        gws = _gate_workspace(x, E)
        C.call("moe_gate_fwd", C.ptr(x), C.dtype_code(x), C.ptr(Wg_c), C.ptr(bg_c), C.ptr(noise), None, T, d, E,
               spec.top_k, spec.score_mode, int(spec.want_psum), C.ptr(logits), C.ptr(idx), C.ptr(score),
               C.ptr(tile_hist), C.ptr(tile_psum), C.ptr(gws), st)
        ctx.spec, ctx.has_bg = spec, bg is not None
        ctx.save_for_backward(x, Wg_c, logits, idx, score)
        idx64 = idx.long()
        ctx.mark_non_differentiable(idx64, logits)
        return idx64, score, logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _didx, dscore, _dlogits):
        x, Wg, logits, idx, score = ctx.saved_tensors
        spec: RouteSpec = ctx.spec
        T, d = x.shape
        E = Wg.shape[0]
        dev = x.device
        st = C.stream_ptr()
        dscore = torch.zeros_like(score) if dscore is None else dscore.float().contiguous()
        dlogits = _f32((T, E), dev)
        C.call("moe_gate_bwd", C.ptr(logits), C.ptr(idx), C.ptr(score), C.ptr(dscore), None, T, E, spec.top_k,
               spec.score_mode, C.ptr(dlogits), st)
        dx = torch.empty_like(x)
        C.call("moe_dispatch_bwd", None, None, C.ptr(dlogits), C.ptr(idx), C.ptr(Wg), T, d, E, spec.top_k,
               int(spec.score_mode == C.SCORE_FULL_SOFTMAX), C.ptr(dx), C.dtype_code(dx), st)
        ws = torch.empty(C.lib.moe_gate_wgrad_workspace_bytes(T, d, E), dtype=torch.uint8, device=dev)
        dWg = _f32((E, d), dev)
        dbg = _f32(E, dev) if ctx.has_bg else None
        C.call("moe_gate_wgrad", C.ptr(dlogits), C.ptr(x), C.dtype_code(x), T, d, E, C.ptr(ws), C.ptr(dWg),
               C.ptr(dbg), st)
        return dx, dWg, dbg, None, None
