"""Expert parameter containers with FastMoE's names and shapes (`fmoe/linear.py`, `fmoe/transformer.py`
upstream): `experts.htoh4.{weight[E,h,d], bias[E,h]}`, `experts.h4toh.{weight[E,d,h], bias[E,d]}` so
FastMoE checkpoints load unchanged."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class FMoELinear(nn.Module):
    """num_expert independent linear maps held as one [E, out, in] parameter (fp32 master weights).
    The grouped tcgen05 GEMM consumes a bf16 copy; this module itself is only a parameter holder."""

    def __init__(self, num_expert: int, in_feat: int, out_feat: int, bias: bool = True, rank: int = 0):
        super().__init__()
        self.num_expert = num_expert
        self.in_feat = in_feat
        self.out_feat = out_feat
        self.rank = rank
        self.weight = nn.Parameter(torch.empty(num_expert, out_feat, in_feat))
        if bias:
            self.bias = nn.Parameter(torch.zeros(num_expert, out_feat))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        # same initialisation as upstream FMoELinear: kaiming_uniform(a=sqrt(5)) on the 3-D weight, zero bias
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def extra_repr(self) -> str:
        return "num_expert={}, in_features={}, out_features={}, bias={}, rank={}".format(
            self.num_expert, self.in_feat, self.out_feat, self.bias is not None, self.rank)

    def forward(self, inp, fwd_expert_count):
        raise NotImplementedError(
            "FMoELinear is a parameter holder in the B200 path; the fused layer (FMoETransformerMLP.forward) "
            "runs both expert GEMMs in the grouped tcgen05 kernels")
