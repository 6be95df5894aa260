"""Residual-MoE block forward with *real* token skipping — the drop-in for the reference's
`forward_residule_moe` (/root/reference/models/resMoE.py:126-145), which the factory binds onto every
`Block` (models/resMoE.py:178-186).

The reference multiplies skipped tokens by 0 and still pushes them through the whole MoE layer
(`self.mlp(tk)`), so the "slim" gate saves nothing.  Here the 0/1 keep-mask of `moe_gate` goes into the fused
layer (`FMoETransformerMLP.forward(..., token_mask=)`): the gate kernel never routes a skipped token, so dispatch,
the grouped expert GEMMs and combine only see the kept ones.  Results are those of the reference expression:
skipped rows receive mlp(0) and the straight-through gradient of the mask includes the J0^T dy term
(`fmoe.functions.SkipFill`).

Usage, mirroring models/resMoE.py:184-185:

    block.forward = fmoe.residual.forward_residual_moe.__get__(block, block.__class__)

`block` needs the attributes the reference block has: norm1, attn, drop_path, norm2, mlp (an
`fmoe.FMoETransformerMLP`), dense_gate and moe_gate (the reference's `Gate`, models/resMoE.py:32-85, or anything
returning a [B, N, 2] tensor of (skip, keep) weights).
"""
from __future__ import annotations

import torch


def _is_hard(gate) -> bool:
    """The mask is 0/1 unless the gate is in its soft training mode (models/resMoE.py:72-74)."""
    return not (getattr(gate, "training", False) and not getattr(gate, "is_hard", True))


def _split(x: torch.Tensor, weights: torch.Tensor):
    """(kept part, skipped part) of x under a [B, N, 2] = (skip, keep) weight tensor."""
    return x * weights[..., 1:2], x * weights[..., 0:1]


def moe_with_skip(mlp, x: torch.Tensor, weights: torch.Tensor, hard: bool = True) -> torch.Tensor:
    """`mlp(x * keep) + x * keep + x * skip` (models/resMoE.py:139-143 without drop_path) with skipped tokens left out
    of the expert computation when the weights are a hard 0/1 mask."""
    tk, skip_tk = _split(x, weights)
    if hard and not getattr(mlp, "world_size", 1) > 1:
        return mlp(tk, token_mask=weights[..., 1]), tk, skip_tk
    return mlp(tk), tk, skip_tk


def forward_residual_moe(self, x: torch.Tensor) -> torch.Tensor:
    x = self.norm1(x)
    tk, skip_tk = _split(x, self.dense_gate(x))
    x = self.norm2(self.drop_path(self.attn(tk)) + tk + skip_tk)
    gate = self.moe_gate
    w = gate(x)
    hard = _is_hard(gate) and not getattr(gate, "disable", False)   # a disabled gate keeps every token: nothing to skip
    y, tk, skip_tk = moe_with_skip(self.mlp, x, w, hard)
    return self.drop_path(y) + tk + skip_tk
