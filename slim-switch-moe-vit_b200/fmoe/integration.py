"""Glue between the B200 MoE layer and the reference's model zoo / training loop.

Nothing here is on the hot path; it is what a maintainer of d0-rb/slim-switch-moe-vit calls from the files that stay
unchanged:

* `install_switch_moe(model, ...)`  — the swap the reference's factories perform by hand
  (/root/reference/models/resMoE.py:200-208: `module.mlp = CustomizedMoEMLP(...)` for every `Block`), generalised to
  the north-star configs (Switch top-1 with capacity, GShard top-2, MoE every `moe_stride`-th block, expert
  parallelism).  Under expert parallelism it also lists the expert parameters in
  `model._ddp_params_and_buffers_to_ignore`, so the unchanged `DistributedDataParallel(model, device_ids=[args.gpu])`
  at /root/reference/main.py:610-612 leaves them out of its all-reduce.
* `MoEAuxCriterion(criterion, model, coef)` — /root/reference/engine.py:52-54 computes `criterion(samples, outputs,
  targets)` and never reads the gates' load-balancing loss; wrapping the criterion adds `coef * sum(gate.get_loss())`
  without touching engine.py.
* `load_balance_stats(model)` / `log_load_balance(...)` — per-layer expert counts, drop rate and imbalance for the
  logging block the reference leaves commented out (/root/reference/main.py:945-951).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .gates import GShardGate, NaiveGate, SwitchGate
from .layers import FMoE
from .transformer import FMoETransformerMLP


def make_gate(kind: str, capacity_factor: float = 1.25, switch_eps: float = 0.0):
    """Gate constructor with the (d_model, num_expert, world_size, top_k) call signature `FMoE` uses.
    kind: "naive" (what the reference configures), "switch" (top-1 + capacity + aux loss), "gshard" (top-2 + capacity)."""
    cf = (capacity_factor, capacity_factor)
    if kind == "naive":
        return NaiveGate
    if kind == "switch":
        return partial(SwitchGate, switch_eps=switch_eps, capacity=cf)
    if kind == "gshard":
        return partial(GShardGate, capacity=cf)
    raise ValueError(f"unknown gate kind {kind!r} (naive | switch | gshard)")


def build_moe_mlp(dim: int, hidden: int, *, num_experts: int, top_k: int, gate: str = "switch", capacity_factor: float = 1.25,
                  world_size: int = 1, moe_group=None, drop: float = 0.0) -> FMoETransformerMLP:
    """The module `CustomizedMoEMLP(dim, hidden, moe_num_experts=, moe_top_k=, drop=)` builds
    (/root/reference/models/resMoE.py:15-29), with the gate and the expert-parallel sharding as options.
    `num_experts` is the GLOBAL expert count; every rank holds `num_experts // world_size` of them."""
    if num_experts % world_size:
        raise ValueError("num_experts must be divisible by the expert-parallel world size")
    act = nn.Sequential(nn.GELU(), nn.Dropout(p=drop))      # reference models/resMoE.py:25
    return FMoETransformerMLP(num_experts // world_size, dim, hidden, act, top_k=top_k, gate=make_gate(gate, capacity_factor),
                              world_size=world_size, moe_group=moe_group)


def _default_world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def install_switch_moe(model: nn.Module, block_type, embed_dim: int, *, num_experts: int, top_k: int, gate: str = "switch",
                       capacity_factor: float = 1.25, moe_stride: int = 1, mlp_ratio: int = 4, expert_parallel: bool = False,
                       moe_group=None, drop: float = 0.0):
    """Replace `block.mlp` of every `moe_stride`-th `block_type` module of `model` (in module order; stride 2 = blocks
    1, 3, 5, ...) by the B200 MoE layer.  Returns the list of installed layers.  With `expert_parallel=True` the experts
    are sharded over the (already initialised) default process group and kept out of DDP's all-reduce."""
    world = _default_world_size() if expert_parallel else 1
    layers, i = [], 0
    for _, module in model.named_modules():
        if isinstance(module, block_type):
            if moe_stride > 0 and i % moe_stride == moe_stride - 1:
                module.mlp = build_moe_mlp(embed_dim, embed_dim * mlp_ratio, num_experts=num_experts, top_k=top_k, gate=gate,
                                           capacity_factor=capacity_factor, world_size=world, moe_group=moe_group, drop=drop)
                layers.append(module.mlp)
            i += 1
    if world > 1:
        from .distributed import mark_expert_parallel
        mark_expert_parallel(model, world)
    return layers


def moe_layers(model: nn.Module):
    return [m for m in model.modules() if isinstance(m, FMoE)]


class MoEAuxCriterion(nn.Module):
    """`criterion(samples, outputs, targets) + coef * sum over MoE layers of gate.get_loss()`.

    Same call signature as the reference's `DistillationLoss` (/root/reference/losses.py:13-73, called at
    /root/reference/engine.py:53-54), so `criterion = MoEAuxCriterion(criterion, model_without_ddp, 0.01)` after
    /root/reference/main.py:700 is the whole integration.  The gates' losses are consumed (cleared) on every call."""

    def __init__(self, criterion, model: nn.Module, coef: float = 0.01):
        super().__init__()
        self.criterion = criterion
        self.coef = float(coef)
        self._layers = moe_layers(model)      # plain list: the model is not registered as a sub-module of the criterion
        self.last_aux = None

    def forward(self, *args, **kwargs):
        loss = self.criterion(*args, **kwargs)
        aux = None
        for layer in self._layers:
            if layer.gate.has_loss:
                g = layer.gate.get_loss(clear=True).sum()
                aux = g if aux is None else aux + g
        self.last_aux = None if aux is None else aux.detach()
        return loss if aux is None else loss + self.coef * aux


@torch.no_grad()
def load_balance_stats(model: nn.Module) -> dict:
    """Per-MoE-layer routing statistics of the LAST forward (one device -> host copy per layer; call it where the
    reference already synchronises, e.g. next to `loss.item()` at /root/reference/engine.py:56 or once per epoch):
    routed / kept pairs per expert, drop rate, and max-over-mean load."""
    out = {}
    for name, m in model.named_modules():
        if isinstance(m, FMoE) and getattr(m, "last_count", None) is not None:
            both = torch.stack((m.last_count, m.last_kept)).cpu()
            count, kept = both[0].double(), both[1].double()
            routed = float(count.sum())
            out[name] = {
                "routed_pairs": int(routed), "kept_pairs": int(kept.sum()),
                "drop_rate": 0.0 if routed == 0 else float(1.0 - kept.sum() / routed),
                "max_over_mean_load": 0.0 if routed == 0 else float(count.max() / count.mean()),
                "count": [int(v) for v in count.tolist()], "kept": [int(v) for v in kept.tolist()],
            }
    return out


def log_load_balance(model: nn.Module, metric_logger=None, writer=None, step: int | None = None) -> dict:
    """Feed `load_balance_stats` into the reference's loggers: `metric_logger.update(**scalars)`
    (/root/reference/utils.py:118-211) and / or `writer.log_scalar(name, value, step)` (utils.py:299-319 — the call the
    commented-out block at /root/reference/main.py:945-951 was going to make).  Returns the scalars."""
    stats = load_balance_stats(model)
    scalars = {}
    if stats:
        scalars["moe_drop_rate"] = sum(s["drop_rate"] for s in stats.values()) / len(stats)
        scalars["moe_max_load"] = max(s["max_over_mean_load"] for s in stats.values())
    if metric_logger is not None and scalars:
        metric_logger.update(**scalars)
    if writer is not None:
        for lname, s in stats.items():
            writer.log_scalar(f"moe/{lname}/drop_rate", s["drop_rate"], step)
            writer.log_scalar(f"moe/{lname}/max_over_mean_load", s["max_over_mean_load"], step)
    return scalars
