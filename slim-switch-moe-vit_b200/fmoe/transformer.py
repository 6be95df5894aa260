"""`FMoETransformerMLP` — the drop-in for the class the reference imports at
/root/reference/models/resMoE.py:6 and subclasses at models/resMoE.py:15-29
(`super().__init__(moe_num_experts, in_features, hidden_features, activation, top_k=moe_top_k)`).
Constructor signature, parameter names and forward contract follow FastMoE's `fmoe/transformer.py`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import FMoE
from .linear import FMoELinear


def _check_activation(activation) -> None:
    """The fc1 epilogue hard-wires exact-erf GELU followed by nothing.  The reference builds
    `nn.Sequential(GELU(), Dropout(p=drop))` with drop=0.0 in every factory
    (/root/reference/models/resMoE.py:25,161,183,198,207); anything else is refused (no fallback)."""
    mods = list(activation) if isinstance(activation, nn.Sequential) else [activation]
    seen_gelu = False
    for m in mods:
        if isinstance(m, nn.GELU):
            if getattr(m, "approximate", "none") != "none" or seen_gelu:
                raise NotImplementedError("only exact-erf nn.GELU() is fused into the expert GEMM epilogue")
            seen_gelu = True
        elif isinstance(m, nn.Dropout):
            if m.p != 0.0:
                raise NotImplementedError("activation Dropout(p>0) inside the experts is not supported (reference uses p=0)")
        elif isinstance(m, nn.Identity):
            pass
        else:
            raise NotImplementedError(f"unsupported expert activation {type(m).__name__}: the B200 layer fuses nn.GELU() only")
    if not seen_gelu:
        raise NotImplementedError("expert activation must contain nn.GELU()")


class _Expert(nn.Module):
    """Parameter holder with upstream's attribute names: htoh4 (d -> h) and h4toh (h -> d)."""

    def __init__(self, num_expert, d_model, d_hidden, activation, rank=0):
        super().__init__()
        self.htoh4 = FMoELinear(num_expert, d_model, d_hidden, bias=True, rank=rank)
        self.h4toh = FMoELinear(num_expert, d_hidden, d_model, bias=True, rank=rank)
        self.activation = activation

    def forward(self, inp, fwd_expert_count):
        raise NotImplementedError("_Expert is evaluated inside the fused layer; call FMoETransformerMLP.forward")


class FMoETransformerMLP(FMoE):
    def __init__(self, num_expert=32, d_model=1024, d_hidden=4096, activation=torch.nn.GELU(),
                 expert_dp_comm="none", expert_rank=0, **kwargs):
        super().__init__(num_expert=num_expert, d_model=d_model, **kwargs)
        if d_model % 64 != 0 or d_hidden % 64 != 0:
            raise ValueError(f"d_model={d_model} and d_hidden={d_hidden} must be multiples of 64 (UMMA/TMA tile granularity)")
        _check_activation(activation)
        self.d_hidden = d_hidden
        self.experts = _Expert(num_expert, d_model, d_hidden, activation, rank=expert_rank)
        self.mark_parallel_comm(expert_dp_comm)

    def _expert_params(self):
        e = self.experts
        return e.htoh4.weight, e.htoh4.bias, e.h4toh.weight, e.h4toh.bias

    def forward(self, inp: torch.Tensor, token_mask: torch.Tensor | None = None):
        """inp [..., d_model] -> same shape.  token_mask [...] (optional): see `FMoE.forward` and `fmoe.residual`."""
        original_shape = inp.shape
        inp = inp.reshape(-1, self.d_model)
        output = super().forward(inp, None if token_mask is None else token_mask.reshape(-1))
        if output.dtype != inp.dtype:   # fp16 activations are widened to fp32 inside the layer; the caller gets its dtype back
            output = output.to(inp.dtype)
        return output.reshape(original_shape)
