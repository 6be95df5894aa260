"""`FMoE` — the sparse-MoE module, with FastMoE's constructor signature (`fmoe/layers.py` upstream,
un-vendored; the reference reaches it through `FMoETransformerMLP`, /root/reference/models/resMoE.py:27-29).

Differences from upstream that a user can observe:
  * no CPU / eager fallback: inputs must be CUDA tensors on an sm_100a device;
  * routing is deterministic (token-order positions, lowest-index tie-break) where upstream's
    atomics make the intra-expert order and the capacity victims run-dependent;
  * options whose semantics need a Python-level expert or hook (`expert=`, `gate_hook=`, `mask=`,
    `mask_dict=`, `mp_group=` / `slice_group=`) raise NotImplementedError at construction.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .functions import Bf16WeightCache, MoEFunction, SkipFill, zero_token_path
from .gates import BaseGate, NaiveGate


def mark_module_parallel_comm(module, comm):
    """Tag every parameter with the data-parallel group it is synchronised in (FastMoE convention,
    consumed by `fmoe.distributed.DistributedGroupedDataParallel`)."""
    for p in module.parameters():
        setattr(p, "dp_comm", comm)


class FMoE(nn.Module):
    def __init__(self, num_expert=32, d_model=1024, world_size=1, mp_group=None, slice_group=None, moe_group=None,
                 top_k=2, gate=NaiveGate, expert=None, gate_hook=None, mask=None, mask_dict=None):
        super().__init__()
        for name, val in (("mp_group", mp_group), ("slice_group", slice_group), ("expert", expert),
                          ("gate_hook", gate_hook), ("mask", mask), ("mask_dict", mask_dict)):
            if val is not None:
                raise NotImplementedError(f"FMoE({name}=...) is not supported by the B200 fused layer")
        self.num_expert = num_expert
        self.d_model = d_model
        self.world_size = world_size
        self.slice_group = None
        self.slice_size = 1
        self.slice_rank = 0
        self.top_k = top_k
        self.moe_group = moe_group
        # a gate CLASS as upstream, or any callable with the same (d_model, num_expert, world_size, top_k) signature
        # (e.g. functools.partial(SwitchGate, capacity=(1.25, 1.25)), see fmoe.integration.make_gate)
        self.gate = gate(d_model, num_expert, world_size, top_k) if callable(gate) else None
        if not isinstance(self.gate, BaseGate):
            raise NotImplementedError("gate must build one of the fmoe.gates classes (NaiveGate, SwitchGate, GShardGate)")
        if self.gate.top_k != top_k:
            raise ValueError(f"gate {type(self.gate).__name__} fixes top_k={self.gate.top_k}, layer was given top_k={top_k}")
        self.experts = None            # set by the subclass (FMoETransformerMLP)
        self.experts_fused = True
        self._bf16_cache = Bf16WeightCache()

    def mark_parallel_comm(self, expert_dp_comm="none"):
        if self.experts is not None:
            mark_module_parallel_comm(self.experts, expert_dp_comm)
        mark_module_parallel_comm(self.gate, "gate")

    def _expert_params(self):
        raise NotImplementedError

    def invalidate_weight_cache(self):
        """Drop the cached bf16 copies of the expert weights (only inference forwards reuse them; call this after
        writing weights through `.data` or by any other route that does not bump the tensor version)."""
        self._bf16_cache.invalidate()

    def forward(self, moe_inp: torch.Tensor, token_mask: torch.Tensor | None = None) -> torch.Tensor:
        """moe_inp [T, d_model] -> [T, d_model]; sets `self.gate`'s aux loss as a side effect.

        token_mask [T] (optional; non-zero = keep): the token-skip mask of the reference's residual-MoE block
        (/root/reference/models/resMoE.py:126-145 calls `self.mlp(x * mask)`).  `layer(x * m, token_mask=m)` returns
        what `layer(x * m)` returns for a 0/1 mask — skipped rows get the constant mlp(0), their input gradient is
        J0^T dy — but skipped tokens are never routed: they cost no dispatch, expert-FFN or combine work and take no
        expert capacity (with a capacity-limited gate that is where the two differ: upstream would let them compete
        for the slots of the expert the gate bias favours).  Rows of moe_inp under a zero mask are not read."""
        if moe_inp.dim() != 2 or moe_inp.shape[1] != self.d_model:
            raise ValueError(f"expected [tokens, {self.d_model}] input, got {tuple(moe_inp.shape)}")
        if self.world_size > 1:
            if token_mask is not None:
                raise NotImplementedError("token_mask under expert parallelism: mlp(0) of a skipped token may live on a remote rank")
            from .distributed import ep_forward
            return ep_forward(self, moe_inp)
        T = moe_inp.shape[0]
        W1, b1, W2, b2 = self._expert_params()
        gate = self.gate
        spec = gate.route_spec(T)
        keep = None
        if token_mask is not None:
            if token_mask.numel() != T:
                raise ValueError(f"token_mask has {token_mask.numel()} entries for {T} tokens")
            keep = (token_mask.reshape(T) != 0).to(torch.uint8)
        # a forward that autograd records re-casts the bf16 weight copies (see Bf16WeightCache); inference reuses them
        fresh = torch.is_grad_enabled() and (W1.requires_grad or W2.requires_grad)
        infer = not torch.is_grad_enabled()      # evaluate() / no_grad: nothing is kept for a backward pass
        y, aux, count, kept = MoEFunction.apply(moe_inp, gate.gate.weight, gate.gate.bias, W1, b1, W2, b2, spec,
                                                self._bf16_cache, gate.make_noise(moe_inp), keep, fresh, infer)
        if keep is not None:
            c, J0 = zero_token_path(gate.gate.weight, gate.gate.bias, W1, b1, W2, b2, spec.top_k, spec.score_mode)
            y = SkipFill.apply(y, moe_inp, keep, c, J0)
        gate.finish(aux)
        self.last_count, self.last_kept = count, kept   # load-balance statistics (device tensors, no sync)
        return y
