"""The dense MLP of the hosting block (fc1 -> GELU -> fc2) on the MoE layer's own grouped GEMM, as a single
"expert" that owns every token (block-level fusion; reference block
/root/reference/models/vision_transformer.py:319-322, its `Mlp` comes from timm).

With E = 1 the packed row buffer is the token matrix itself, so there is no routing, dispatch or combine:
six grouped-GEMM launches and two column sums replace the framework's four GEMMs, two GELU passes and
two bias reductions, and the hidden activation is written once (gelu and gelu' come out of the fc1 epilogue).
Parameters and state_dict keys are those of the stock `Mlp` (`fc1.weight [h,d]`, `fc1.bias`, `fc2.weight [d,h]`,
`fc2.bias`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi as C
from .functions import Bf16WeightCache

_TABLES: dict = {}


def _tables(rows_cap: int, dev):
    """tile -> expert table (all zero), live tile count and the single segment [0, rows_cap)."""
    key = (torch.device(dev).index, rows_cap)
    t = _TABLES.get(key)
    if t is None:
        n = rows_cap // C.ROW_ALIGN
        t = (torch.zeros(n, dtype=torch.int32, device=dev), torch.tensor([n], dtype=torch.int32, device=dev),
             torch.tensor([0, rows_cap], dtype=torch.int32, device=dev))
        if torch.cuda.is_current_stream_capturing():
            return t
        _TABLES[key] = t
    return t


def _rows(t: torch.Tensor, rows_cap: int) -> torch.Tensor:
    """[T, c] -> contiguous bf16 [rows_cap, c]; rows past T are zero (they are operands of the weight gradients)."""
    T = t.shape[0]
    if t.dtype == torch.bfloat16 and t.is_contiguous() and T == rows_cap:
        return t
    out = torch.zeros((rows_cap, t.shape[1]), dtype=torch.bfloat16, device=t.device) if T != rows_cap else \
        torch.empty((rows_cap, t.shape[1]), dtype=torch.bfloat16, device=t.device)
    out[:T].copy_(t)
    return out


class _DenseFFN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, cache: Bf16WeightCache):
        if not x.is_cuda:
            raise C.MoeB200Error("fmoe (B200) has no CPU path: DenseFFN needs CUDA tensors")
        shape = x.shape
        d, h = W1.shape[1], W1.shape[0]
        x2 = x.detach().reshape(-1, d)
        T = x2.shape[0]
        rows_cap = (T + C.ROW_ALIGN - 1) // C.ROW_ALIGN * C.ROW_ALIGN
        dev, st, bf = x2.device, C.stream_ptr(), torch.bfloat16
        for w in (W1, b1, W2, b2):
            if w.dtype != torch.float32:
                raise C.MoeB200Error("DenseFFN parameters must be fp32 (master weights)")
        xb = _rows(x2, rows_cap)
        te, nm, sg = _tables(rows_cap, dev)
        W1b, W2b = cache.get(W1.detach().view(1, h, d), W2.detach().view(1, d, h), fresh=True)
        G = torch.empty((rows_cap, h), dtype=bf, device=dev)
        H = torch.empty((rows_cap, h), dtype=bf, device=dev)
        Y = torch.empty((rows_cap, d), dtype=bf, device=dev)
        C.call("moe_grouped_gemm", C.GEMM_FC1, C.ptr(xb), C.ptr(W1b), C.ptr(G), C.ptr(H), C.ptr(b1.detach().contiguous()), None,
               C.ptr(te), C.ptr(nm), None, rows_cap, 1, 0, h, d, st, tag="dense_fc1")
        C.call("moe_grouped_gemm", C.GEMM_FC2, C.ptr(H), C.ptr(W2b), C.ptr(Y), None, C.ptr(b2.detach().contiguous()), None,
               C.ptr(te), C.ptr(nm), None, rows_cap, 1, 0, d, h, st, tag="dense_fc2")
        ctx.save_for_backward(xb, G, H, W1b, W2b)
        ctx.meta = (shape, x.dtype, T, rows_cap, d, h)
        y = Y[:T].view(*shape[:-1], d)
        return y if x.dtype == bf else y.to(x.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xb, G, H, W1b, W2b = ctx.saved_tensors
        shape, x_dtype, T, rows_cap, d, h = ctx.meta
        dev, st, bf = xb.device, C.stream_ptr(), torch.bfloat16
        te, nm, sg = _tables(rows_cap, dev)
        dyb = _rows(dy.reshape(-1, d), rows_cap)
        dU = torch.empty((rows_cap, h), dtype=bf, device=dev)
        dxb = torch.empty((rows_cap, d), dtype=bf, device=dev)
        dW1, db1 = torch.empty((1, h, d), device=dev), torch.empty((1, h), device=dev)
        dW2, db2 = torch.empty((1, d, h), device=dev), torch.empty((1, d), device=dev)
        pte, pnm, psg = C.ptr(te), C.ptr(nm), C.ptr(sg)
        slab_sums = torch.empty(C.lib.moe_slab_colsum_bytes(rows_cap, h) // 4, dtype=torch.float32, device=dev)
        C.call("moe_grouped_gemm", C.GEMM_DGELU, C.ptr(dyb), C.ptr(W2b), C.ptr(dU), C.ptr(slab_sums), None, C.ptr(G),
               pte, pnm, None, rows_cap, 1, 0, h, d, st, tag="dense_dgelu")
        wfl = C.ptr(C.wgrad_flags(1, h, d, dev))
        C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, C.ptr(H), C.ptr(dyb), C.ptr(dW2), None, None, wfl,
               None, None, psg, rows_cap, 1, h, d, 0, st, tag="dense_wgrad2")
        C.call("moe_grouped_gemm", C.GEMM_WGRAD, C.ptr(dU), C.ptr(xb), C.ptr(dW1), None, None, wfl,
               None, None, psg, rows_cap, 1, h, d, 0, st, tag="dense_wgrad1")
        C.call("moe_grouped_gemm", C.GEMM_DGRAD, C.ptr(dU), C.ptr(W1b), C.ptr(dxb), None, None, None,
               pte, pnm, None, rows_cap, 1, 0, d, h, st, tag="dense_dgrad")
        cws = torch.empty(C.lib.moe_segment_colsum_workspace_bytes(rows_cap, d), dtype=torch.uint8, device=dev)
        C.call("moe_segment_colsum", C.ptr(dyb), psg, rows_cap, 1, d, C.ptr(cws), C.ptr(db2), st, tag="dense_db2")
        C.call("moe_slab_colsum_final", C.ptr(slab_sums), psg, 1, h, C.ptr(db1), st, tag="dense_db1")
        dx = dxb[:T].view(shape)
        if x_dtype != bf:
            dx = dx.to(x_dtype)
        return dx, dW1.view(h, d), db1.view(h), dW2.view(d, h), db2.view(d), None


class DenseFFN(nn.Module):
    """Drop-in for the block's dense `Mlp(dim, hidden)` with exact-erf GELU and no dropout."""

    def __init__(self, dim: int, hidden: int):
        super().__init__()
        if dim % 64 or hidden % 64:
            raise ValueError("DenseFFN: dim and hidden must be multiples of 64 (grouped GEMM tiling)")
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self._cache = Bf16WeightCache()

    def forward(self, x):
        if x.dtype != torch.bfloat16:   # full-precision callers keep full precision (library GEMMs); bf16 = the autocast path
            return self.fc2(self.act(self.fc1(x)))
        return _DenseFFN.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self._cache)
