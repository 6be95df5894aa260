"""Gates of the B200 MoE layer — same classes, constructor arguments and loss hooks as FastMoE's
`fmoe/gates/{base_gate,naive_gate,switch_gate,gshard_gate}.py` (upstream, un-vendored).  The
reference uses only NaiveGate (the FastMoE default picked at /root/reference/models/resMoE.py:27-29);
`layer.gate.gate` must stay an `nn.Linear(d_model, tot_expert)` because
/root/reference/models/resmoe_flop_hook.py:7 reads `.in_features/.out_features` from it.

A gate here is a parameter holder plus a routing specification: inside `FMoE.forward` the
projection, top-k, scores and capacity positions are computed by the fused CUDA gate/scan/dispatch
kernels, not by the gate module.  Calling a gate on its own (`gate(x)`) still works and returns
`(top_k_idx int64 [T,k], top_k_score [T,k])` like upstream, through the same CUDA gate kernel.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _cabi as C
from .functions import GateFunction, RouteSpec


class BaseGate(nn.Module):
    def __init__(self, num_expert, world_size):
        super().__init__()
        self.world_size = world_size
        self.num_expert = num_expert
        self.tot_expert = world_size * num_expert
        self.loss = None

    def forward(self, x):
        raise NotImplementedError("Base gate cannot be directly used for fwd")

    def set_loss(self, loss):
        self.loss = loss

    def get_loss(self, clear=True):
        loss = self.loss
        if clear:
            self.loss = None
        return loss

    @property
    def has_loss(self):
        return self.loss is not None

    # ---- B200 fused-path protocol -------------------------------------------------------------
    def route_spec(self, num_tokens: int) -> RouteSpec:
        raise NotImplementedError

    def make_noise(self, x):  # optional additive logit jitter [T, tot_expert] fp32
        return None

    def finish(self, aux_loss):  # called after the fused forward with the kernel-computed load-balancing loss
        raise NotImplementedError


class NaiveGate(BaseGate):
    """Linear -> top-k -> softmax over the k selected logits; no capacity; dummy zero loss."""

    def __init__(self, d_model, num_expert, world_size, top_k=2, gate_bias=True):
        super().__init__(num_expert, world_size)
        self.gate = nn.Linear(d_model, self.tot_expert, bias=gate_bias)
        self.top_k = top_k
        if not 1 <= top_k <= min(8, self.tot_expert):
            raise ValueError(f"top_k={top_k} unsupported (1 <= top_k <= min(8, experts))")

    def route_spec(self, num_tokens):
        return RouteSpec(self.top_k, C.SCORE_TOPK_SOFTMAX, num_tokens * self.top_k, C.AUX_NONE)

    def finish(self, aux_loss):
        self.set_loss(torch.zeros(1, requires_grad=True, device=aux_loss.device))

    def forward(self, inp, return_all_scores=False):
        spec = self.route_spec(inp.shape[0])
        idx, score, logits = GateFunction.apply(inp, self.gate.weight, self.gate.bias, spec, self.make_noise(inp))
        self.set_loss(torch.zeros(1, requires_grad=True, device=inp.device))
        if return_all_scores:
            return idx, score, logits
        return idx, score


def _capacity(cf: float, num_tokens: int, top_k: int, tot_expert: int) -> int:
    """C = ceil(cf * T * k / E) — the Switch/GShard convention adopted by SURVEY.md §8a (upstream
    FastMoE uses ceil(cf * T) per expert, which never binds on one worker)."""
    if cf is None or cf <= 0:
        return num_tokens * top_k
    return max(1, min(num_tokens * top_k, int(math.ceil(cf * num_tokens * top_k / tot_expert))))


class SwitchGate(NaiveGate):
    """Top-1 Switch-Transformer gate: softmax over all experts, score = probability of the chosen
    expert, tokens beyond the per-expert capacity are dropped (zero layer output), aux loss
    E * sum_e f_e P_e with f_e = fraction of kept tokens on e and P_e = mean softmax prob of e.

    Deliberate differences from upstream FastMoE's SwitchGate (SURVEY.md §8a defines the north-star semantics; tuned
    aux-loss coefficients do not carry over unchanged): capacity is ceil(cf * T * k / E) per expert (upstream: ceil(cf * T),
    which never binds on one worker); P_e averages the softmax probability over ALL T tokens (upstream: over the kept
    tokens only: upstream's loss == this one * T / sum(kept)); dropped tokens are the LAST ones of an expert in token order
    (upstream: whichever lose an atomics race)."""

    def __init__(self, d_model, num_expert, world_size, topk=1, switch_eps=0.1, capacity=(1.2, 2.4), gate_bias=True):
        assert topk == 1, "topk should be 1 in switch"
        super().__init__(d_model, num_expert, world_size, top_k=1, gate_bias=gate_bias)
        self.switch_eps = switch_eps
        self.capacity = capacity

    def _cf(self):
        return self.capacity[0 if self.training else 1]

    def route_spec(self, num_tokens):
        cap = _capacity(self._cf(), num_tokens, 1, self.tot_expert)
        return RouteSpec(1, C.SCORE_FULL_SOFTMAX, cap, C.AUX_SWITCH)

    def make_noise(self, x):
        if not self.training or not self.switch_eps:
            return None
        noise = torch.rand(x.shape[0], self.tot_expert, device=x.device, dtype=torch.float32)
        return noise * (2 * self.switch_eps) + (1.0 - self.switch_eps)

    def finish(self, aux_loss):
        self.set_loss(aux_loss)   # E * sum_e f_e P_e, computed (with its gradient) by the scan kernel

    def forward(self, inp):
        spec = self.route_spec(inp.shape[0])
        idx, score, _ = GateFunction.apply(inp, self.gate.weight, self.gate.bias, spec, self.make_noise(inp))
        return idx, score


class GShardGate(NaiveGate):
    """Top-2 GShard gate: NaiveGate scores, per-expert capacity, aux loss mean(c_e * m_e) * E^2 with
    c_e = fraction of pairs routed to e (before capacity) and m_e = mean softmax prob of e.
    Upstream's `random_routing` of the second expert is not implemented (raises if requested).

    Deliberate differences from upstream FastMoE's GShardGate: c_e counts all k picks of a token divided by T * k (upstream:
    the top-1 pick only, divided by T) and the scale is (global expert count)^2 (upstream: the local `num_expert`^2): on one
    worker both losses are E * sum_e c_e m_e and differ only in which picks c_e counts; capacity ceil(cf * T * k / E) equals
    upstream's ceil(cf * T) * k // (W * E) up to rounding (tests/test_oracle_cpu.py pins these relations as identities)."""

    def __init__(self, d_model, num_expert, world_size, topk=2, capacity=(1.2, 2.4), random_routing=False,
                 gate_bias=True):
        assert topk == 2, "topk should be 2 in gshard"
        if random_routing:
            raise NotImplementedError("GShardGate(random_routing=True) is not supported by the B200 path")
        super().__init__(d_model, num_expert, world_size, top_k=2, gate_bias=gate_bias)
        self.capacity = capacity

    def route_spec(self, num_tokens):
        cap = _capacity(self.capacity[0 if self.training else 1], num_tokens, 2, self.tot_expert)
        return RouteSpec(2, C.SCORE_TOPK_SOFTMAX, cap, C.AUX_GSHARD)

    def finish(self, aux_loss):
        self.set_loss(aux_loss)   # mean_e(c_e m_e) * E^2, computed (with its gradient) by the scan kernel

    def forward(self, inp):
        spec = self.route_spec(inp.shape[0])
        idx, score, _ = GateFunction.apply(inp, self.gate.weight, self.gate.bias, spec, None)
        return idx, score
