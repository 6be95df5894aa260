"""Block-level fusion around the MoE layer: the residual add that consumes a
sub-layer's output and the pre-norm that feeds the next sub-layer, in one kernel each way
(reference block: /root/reference/models/vision_transformer.py:319-322).

    x_out, n = add_layer_norm(x_in, delta, weight, bias)     # x_out = x_in + delta ; n = LN(x_out)

The residual stream stays fp32 (as under the reference's autocast, engine.py:52); `n` is produced directly
in `out_dtype` (bf16 under autocast), the dtype the attention projections and the MoE dispatch read.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi as C


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, delta, weight, bias, eps, out_dtype):
        if not x_in.is_cuda:
            raise C.MoeB200Error("fmoe (B200) has no CPU path: add_layer_norm needs CUDA tensors")
        shape = x_in.shape
        d = shape[-1]
        x2 = x_in.detach().reshape(-1, d)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        x2 = x2.contiguous()
        T = x2.shape[0]
        dev, st = x2.device, C.stream_ptr()
        dl = None
        if delta is not None:
            dl = delta.detach().reshape(-1, d).contiguous()
            if dl.dtype not in (torch.float32, torch.bfloat16):
                dl = dl.float()
        n = torch.empty((T, d), dtype=out_dtype, device=dev)
        x_out = torch.empty_like(x2) if dl is not None else x2
        mean = torch.empty(T, dtype=torch.float32, device=dev)
        rstd = torch.empty(T, dtype=torch.float32, device=dev)
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        C.call("moe_addln_fwd", C.ptr(x2), C.ptr(dl), C.dtype_code(dl) if dl is not None else 0, C.ptr(w), C.ptr(b), float(eps),
               T, d, C.ptr(x_out) if dl is not None else None, C.ptr(n), C.dtype_code(n), C.ptr(mean), C.ptr(rstd), st)
        ctx.save_for_backward(x_out, mean, rstd, w)
        ctx.has_delta = dl is not None
        ctx.delta_dtype = dl.dtype if dl is not None else None
        ctx.shape = shape
        ctx.set_materialize_grads(False)
        if dl is not None:
            return x_out.view(shape), n.view(shape)
        return n.view(shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        x, mean, rstd, w = ctx.saved_tensors
        T, d = x.shape
        dev, st = x.device, C.stream_ptr()
        if ctx.has_delta:
            dx_out, dn = grads
        else:
            dx_out, dn = None, grads[0]
        if dn is None:
            dn = torch.zeros((T, d), dtype=torch.bfloat16, device=dev)
        dn = dn.reshape(T, d)
        if dn.dtype not in (torch.float32, torch.bfloat16):
            dn = dn.float()
        dn = dn.contiguous()
        if dx_out is not None:
            dx_out = dx_out.reshape(T, d).float().contiguous()
        dx_in = torch.empty((T, d), dtype=torch.float32, device=dev)
        d_delta = torch.empty((T, d), dtype=ctx.delta_dtype, device=dev) if (ctx.has_delta and ctx.delta_dtype != torch.float32) else None
        ws = torch.empty(C.lib.moe_addln_bwd_workspace_bytes(T, d), dtype=torch.uint8, device=dev)
        dgamma = torch.empty(d, dtype=torch.float32, device=dev)
        dbeta = torch.empty(d, dtype=torch.float32, device=dev)
        C.call("moe_addln_bwd", C.ptr(dn), C.dtype_code(dn), C.ptr(dx_out), C.ptr(x), C.ptr(mean), C.ptr(rstd), C.ptr(w), T, d,
               C.ptr(dx_in), C.ptr(d_delta), C.dtype_code(d_delta) if d_delta is not None else 0, C.ptr(ws), C.ptr(dgamma),
               C.ptr(dbeta), st)
        dx_in = dx_in.view(ctx.shape)
        if ctx.has_delta:
            dd = d_delta.view(ctx.shape) if d_delta is not None else dx_in
            return dx_in, dd, dgamma, dbeta, None, None
        return dx_in, None, dgamma, dbeta, None, None


def add_layer_norm(x_in, delta, weight, bias, eps=1e-6, out_dtype=None):
    """Returns (x_out, n) when `delta` is given, else n = LN(x_in) alone."""
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
        if out_dtype not in (torch.float32, torch.bfloat16):
            out_dtype = torch.float32     # fp16 autocast (reference engine.py:52): keep the normalised output fp32
    return _AddLayerNorm.apply(x_in, delta, weight, bias, eps, out_dtype)


class AddLayerNorm(nn.LayerNorm):
    """`nn.LayerNorm` parameters (same state_dict keys) with the fused residual-add forward:
    `x_out, n = m(x_in, delta)`; `n = m(x_in)` when there is nothing to add."""

    def forward(self, x_in, delta=None):
        return add_layer_norm(x_in, delta, self.weight, self.bias, self.eps)


class _LinearFastBiasGrad(torch.autograd.Function):
    """y = x W^T + b with the library GEMMs (cuBLAS) for y, dx and dW, and the two-stage deterministic column-sum
    kernel for db — the framework's generic reduction takes 60 us per projection at 50 K rows (11 % of the step)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.bfloat16)
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.reshape(-1, N)
        if dy2.dtype not in (torch.float32, torch.bfloat16):
            dy2 = dy2.float()
        dy2 = dy2.contiguous()
        dx = (dy2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = dy2.t() @ x.reshape(-1, K)
        db = None
        if ctx.has_bias:
            rows = dy2.shape[0]
            ws = torch.empty(C.lib.moe_colsum_workspace_bytes(rows, N), dtype=torch.uint8, device=dy2.device)
            db32 = torch.empty(N, dtype=torch.float32, device=dy2.device)
            C.call("moe_colsum", C.ptr(dy2), C.dtype_code(dy2), rows, N, C.ptr(ws), C.ptr(db32), C.stream_ptr())
            db = db32.to(w.dtype)
        return dx, dw, db


def fast_linear(x, weight, bias):
    """Drop-in for F.linear on CUDA tensors whose output features are a multiple of 8."""
    if x.is_cuda and weight.shape[0] % 8 == 0:
        return _LinearFastBiasGrad.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


class Linear(nn.Linear):
    """`nn.Linear` (same parameters / state_dict) with the fast bias-gradient backward."""

    def forward(self, x):
        return fast_linear(x, self.weight, self.bias)
