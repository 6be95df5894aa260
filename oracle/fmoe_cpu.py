"""oracle/fmoe_cpu.py — plain-PyTorch CPU restatement of `fmoe.FMoETransformerMLP` with FastMoE's
module structure (gate = NaiveGate{gate: nn.Linear}, experts = {htoh4, h4toh}), so that the
UNMODIFIED reference models (/root/reference/models/resMoE.py:15-29, :151-209) can be built on it.

TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle/moe_oracle.py).  Two uses:
  * tests/golden/make_golden.py installs it as `sys.modules['fmoe']`, imports the reference's
    `CustomizedMoEMLP` and records golden input/output vectors;
  * bench.py's `cpu_baseline` / `--impl reference` legs time it on the host cores — FastMoE itself
    has no CPU kernels, so this restated path is the only CPU form of the reference's layer.

Arithmetic: fp32 throughout (what FastMoE's SGEMM path computes), autograd by PyTorch.  Routing is
the canonical one (lowest index on ties, token-order ranks).  With `exact_logit_order=True` the
gate logits come from oracle/gate_ref.c (bit-identical to the CUDA kernel); otherwise `F.linear`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import moe_oracle as O


class _Linear3(nn.Module):
    def __init__(self, num_expert, in_feat, out_feat):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(num_expert, out_feat, in_feat))
        self.bias = nn.Parameter(torch.zeros(num_expert, out_feat))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _Expert(nn.Module):
    def __init__(self, num_expert, d_model, d_hidden, activation):
        super().__init__()
        self.htoh4 = _Linear3(num_expert, d_model, d_hidden)
        self.h4toh = _Linear3(num_expert, d_hidden, d_model)
        self.activation = activation


class NaiveGate(nn.Module):
    def __init__(self, d_model, num_expert, world_size, top_k=2):
        super().__init__()
        self.gate = nn.Linear(d_model, num_expert * world_size)
        self.top_k = top_k
        self.loss = None

    def set_loss(self, loss):
        self.loss = loss

    def get_loss(self, clear=True):
        loss, self.loss = self.loss, (None if clear else self.loss)
        return loss

    @property
    def has_loss(self):
        return self.loss is not None


class FMoETransformerMLP(nn.Module):
    exact_logit_order = False  # class-level switch used by the golden generator

    def __init__(self, num_expert=32, d_model=1024, d_hidden=4096, activation=nn.GELU(), expert_dp_comm="none",
                 expert_rank=0, world_size=1, top_k=2, gate=NaiveGate, score_mode=O.SCORE_TOPK_SOFTMAX,
                 capacity_factor=0.0, **kwargs):
        super().__init__()
        assert world_size == 1
        self.num_expert, self.d_model, self.d_hidden, self.top_k = num_expert, d_model, d_hidden, top_k
        self.world_size = world_size
        self.score_mode, self.capacity_factor = score_mode, capacity_factor
        self.gate = gate(d_model, num_expert, world_size, top_k)
        self.experts = _Expert(num_expert, d_model, d_hidden, activation)

    def forward(self, inp):
        shape = inp.shape
        x = inp.reshape(-1, self.d_model)
        T, E, k = x.shape[0], self.num_expert, self.top_k
        g = self.gate.gate
        logits = F.linear(x, g.weight, g.bias)
        if self.exact_logit_order:
            exact = O.gate_logits(x, g.weight, g.bias)
            logits = logits + (exact - logits).detach()   # exact values, autograd of F.linear
        # canonical top-k: descending value, ties -> lowest index (stable sort on -logits)
        order = torch.sort(-logits.detach(), dim=-1, stable=True).indices[:, :k]
        picked = logits.gather(1, order)
        if self.score_mode == O.SCORE_TOPK_SOFTMAX:
            score = torch.softmax(picked, dim=-1)
        else:
            score = torch.softmax(logits, dim=-1).gather(1, order)
        flat_e = order.reshape(-1)
        perm = torch.sort(flat_e, stable=True).indices            # token-order inside each expert
        counts = torch.bincount(flat_e, minlength=E)
        cap = O.capacity_from_factor(self.capacity_factor, T, k, E)
        starts = torch.cumsum(counts, 0) - counts
        rank = torch.empty_like(perm)
        rank[perm] = torch.arange(perm.numel()) - starts[flat_e[perm]]
        keep = rank < cap
        if self.score_mode == O.SCORE_FULL_SOFTMAX:
            # Switch load-balancing loss E * sum_e f_e P_e: f_e = share of kept pairs, P_e = mean prob
            kept_e = torch.bincount(flat_e[keep], minlength=E).to(logits.dtype)
            f = kept_e / kept_e.sum().clamp(min=1)
            self.gate.set_loss(E * (f * torch.softmax(logits, dim=-1).mean(0)).sum())
        else:
            self.gate.set_loss(torch.zeros(1, requires_grad=True))   # NaiveGate: dummy zero loss
        y = x.new_zeros(T, self.d_model)
        W1, b1, W2, b2 = self.experts.htoh4.weight, self.experts.htoh4.bias, self.experts.h4toh.weight, self.experts.h4toh.bias
        flat_s = score.reshape(-1)
        off = 0
        for e in range(E):
            n = int(counts[e])
            sel = perm[off:off + n]
            off += n
            sel = sel[keep[sel]]
            if sel.numel() == 0:
                continue
            tok = sel // k
            hmid = self.experts.activation(F.linear(x[tok], W1[e], b1[e]))
            out = F.linear(hmid, W2[e], b2[e])
            y = y.index_add(0, tok, out * flat_s[sel].unsqueeze(1))
        return y.reshape(shape)
