"""Expert parallelism for the B200 MoE layer (FastMoE's `world_size > 1` mode, which the
reference never switches on — /root/reference/models/resMoE.py:27-29 leaves world_size at 1 — and which
BASELINE.json configs[2..3] ask for).

Layout: global expert g lives on rank g // E_local.  Every rank routes its own tokens over all
E = W * E_local experts with a per-source-rank capacity C = ceil(cf * T_local * k / E), so routing needs no
cross-rank traffic and stays bit-exact against a single-process oracle run on the rank's token shard.
The dispatch kernel writes FIXED slabs [E, slab_rows, d] (slab_rows = C rounded up to 256), which makes
both all-to-alls static-size — no `.item()` / `.cpu()` round trip for the counts as in upstream:

    forward   gate+scan+dispatch -> a2a(counts) + a2a(slabs) -> repack -> expert FFN -> unpack -> a2a -> combine
    backward  combine_bwd -> a2a -> repack -> expert FFN backward -> unpack -> a2a -> gate/dispatch backward

Gates without a capacity (NaiveGate, what the reference configures) would need slabs of T_local*k rows
per expert; they are refused under expert parallelism — use SwitchGate / GShardGate.
The collectives are `torch.distributed.all_to_all_single` on the layer's `moe_group` (NCCL over NVLink on
the GPU box; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _cabi as C
from .functions import Bf16WeightCache, RouteSpec, _as_kernel_input, _f32, _i32, route


def slab_rows_for(capacity: int) -> int:
    """Rows of one (source rank, expert) slab: the capacity rounded up to the GEMM tile height."""
    return (int(capacity) + C.ROW_ALIGN - 1) // C.ROW_ALIGN * C.ROW_ALIGN


def all_to_all_slabs(send: torch.Tensor, group=None, tag: str = "a2a") -> torch.Tensor:
    """send[W, ...] -> recv[W, ...]: chunk j of rank r becomes chunk r of rank j (equal, static sizes)."""
    recv = torch.empty_like(send)
    if C.PROF.enabled and send.is_cuda:   # same per-call CUDA-event timing as the C-ABI launches (bench.py)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_to_all_single(recv, send, group=group)
        b.record()
        C.PROF.events.setdefault(tag, []).append((a, b))
    else:
        dist.all_to_all_single(recv, send, group=group)
    return recv


def expert_parameters(model: torch.nn.Module):
    """(name, parameter) of every expert weight of every expert-parallel MoE layer in `model`."""
    from .layers import FMoE
    out = []
    for mod_name, mod in model.named_modules():
        if isinstance(mod, FMoE) and mod.world_size > 1 and mod.experts is not None:
            for n, p in mod.experts.named_parameters():
                out.append((f"{mod_name}.experts.{n}" if mod_name else f"experts.{n}", p))
    return out


def _check_group(layer) -> None:
    """Experts are sharded over the layer's `moe_group`, and their gradients are scaled by 1 / (global world size)
    without any all-reduce: that is only right when the expert-parallel group IS the whole world.  A strict
    subgroup (experts replicated across data-parallel groups, FastMoE's `expert_dp_comm='dp'`) would leave the
    replicas unsynchronised — refuse it instead of diverging silently."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("expert parallelism (world_size > 1) needs an initialised torch.distributed process group")
    g, w = dist.get_world_size(layer.moe_group), dist.get_world_size()
    if not (g == layer.world_size == w):
        raise NotImplementedError(
            f"expert-parallel group of {g} ranks, layer.world_size={layer.world_size}, global world size {w}: the B200 layer "
            "shards experts over the whole world only (no expert replicas across data-parallel groups)")


def mark_expert_parallel(model: torch.nn.Module, world_size: int | None = None) -> list[str]:
    """Prepares `model` for `torch.nn.parallel.DistributedDataParallel` (reference main.py:610-612):
    expert parameters are sharded, not replicated, so they are excluded from DDP's all-reduce
    (`_ddp_params_and_buffers_to_ignore`) and their gradients — which already hold the contributions of
    every rank's tokens after the backward all-to-all — are divided by W, the same 1/W DDP applies to
    the replicated parameters (so the optimised objective is the mean of the ranks' losses everywhere).
    Returns the ignored parameter names."""
    from .layers import FMoE
    for mod in model.modules():
        if isinstance(mod, FMoE) and mod.world_size > 1:
            _check_group(mod)
    names = []
    for name, p in expert_parameters(model):
        names.append(name)
        if not getattr(p, "_moe_ep_hooked", False):
            w = float(world_size if world_size is not None else dist.get_world_size())
            p.register_hook(lambda g, w=w: g / w)
            p._moe_ep_hooked = True
    ignored = list(getattr(model, "_ddp_params_and_buffers_to_ignore", []))
    model._ddp_params_and_buffers_to_ignore = sorted(set(ignored) | set(names))
    return names


def wrap_ddp(model: torch.nn.Module, local_rank: int, **kw):
    mark_expert_parallel(model)
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], **kw)


class DistributedGroupedDataParallel(torch.nn.parallel.DistributedDataParallel):
    """Name-compatible stand-in for `fmoe.DistributedGroupedDataParallel`: plain DDP over the data-parallel
    group with the expert parameters left out of the all-reduce."""

    def __init__(self, module, **kw):
        mark_expert_parallel(module)
        super().__init__(module, **kw)


def full_state_dict(model: torch.nn.Module, group=None) -> dict:
    """EP-aware `state_dict()`: expert parameters are all-gathered over the expert-parallel group
    into their global `[W * E_local, ...]` shape, everything else is taken as is.  The result is what a
    single-process (world_size = 1) model — or a FastMoE checkpoint of the same architecture — holds, so the
    reference's rank-0 `utils.save_on_master` (/root/reference/utils.py:264-266) no longer loses remote experts.
    Collective: every rank of the group must call it; every rank gets the full dict."""
    sd = model.state_dict()
    expert_names = {n for n, _ in expert_parameters(model)}
    W = dist.get_world_size(group)
    out = {}
    for name, t in sd.items():
        if name in expert_names and W > 1:
            parts = [torch.empty_like(t) for _ in range(W)]
            dist.all_gather(parts, t.contiguous(), group=group)
            out[name] = torch.cat(parts, dim=0)
        else:
            out[name] = t
    return out


def load_full_state_dict(model: torch.nn.Module, full: dict, group=None, strict: bool = True):
    """Inverse of `full_state_dict`: every rank keeps rows [r * E_local, (r + 1) * E_local) of the expert tensors."""
    expert_names = {n for n, _ in expert_parameters(model)}
    W, r = dist.get_world_size(group), dist.get_rank(group)
    local = {}
    for name, t in full.items():
        if name in expert_names and W > 1:
            El = t.shape[0] // W
            local[name] = t[r * El:(r + 1) * El]
        else:
            local[name] = t
    return model.load_state_dict(local, strict=strict)


class EPMoEFunction(torch.autograd.Function):
    """y, aux_loss, count, kept = expert-parallel MoE(x; Wg, bg, local W1, b1, W2, b2)."""

    @staticmethod
    def forward(ctx, x, Wg, bg, W1, b1, W2, b2, spec: RouteSpec, cache: Bf16WeightCache, noise, group, world,
                fresh: bool = True):
        x = _as_kernel_input(x)
        T, d = x.shape
        El, h = W1.shape[0], W1.shape[1]
        E, k, W = Wg.shape[0], spec.top_k, world
        assert E == El * W, "gate must score world_size * num_expert experts"
        dev, st = x.device, C.stream_ptr()
        slab = slab_rows_for(spec.capacity)
        Wg_c, W1_c, W2_c = Wg.detach().contiguous(), W1.detach().contiguous(), W2.detach().contiguous()
        bg_c = None if bg is None else bg.detach().contiguous()

        r = route(x, Wg_c, bg_c, spec, noise, slab_rows=slab)                       # send slabs [E, slab, d]
        # both exchanges run on NCCL's stream while this stream refreshes the bf16 weight copies (after an optimizer step)
        kept_recv = torch.empty((W, El), dtype=torch.int32, device=dev)
        recv_x = torch.empty((W, El * slab * d), dtype=torch.bfloat16, device=dev)
        w1 = dist.all_to_all_single(kept_recv, r["kept"].view(W, El), group=group, async_op=True)     # [W(src), El]
        w2 = dist.all_to_all_single(recv_x, r["xbuf"].view(W, El * slab * d), group=group, async_op=True)  # [W(src), El, slab, d]
        W1b, W2b = cache.get(W1_c, W2_c, fresh)
        w1.wait()
        w2.wait()

        rows_cap = C.rows_cap(W * El * slab, 1, El, W * El * slab)                   # every received row could be live
        max_mtiles = rows_cap // C.ROW_ALIGN
        tb = dict(slab_dst=_i32((W, El), dev), kept=_i32(El, dev), seg_start=_i32(El + 1, dev),
                  tile_expert=_i32(max_mtiles, dev), num_mtiles=_i32(1, dev))
        C.call("moe_ep_tables", C.ptr(kept_recv), W, El, C.ptr(tb["slab_dst"]), C.ptr(tb["kept"]), C.ptr(tb["seg_start"]),
               C.ptr(tb["tile_expert"]), C.ptr(tb["num_mtiles"]), max_mtiles, st)
        bf = torch.bfloat16
        xbuf = torch.empty((rows_cap, d), dtype=bf, device=dev)
        C.call("moe_ep_repack", C.ptr(recv_x), C.ptr(xbuf), C.ptr(kept_recv), C.ptr(tb["slab_dst"]), C.ptr(tb["seg_start"]),
               C.ptr(tb["kept"]), W, El, slab, d, 1, st)

        G = torch.empty((rows_cap, h), dtype=bf, device=dev)
        H = torch.empty((rows_cap, h), dtype=bf, device=dev)
        Y = torch.empty((rows_cap, d), dtype=bf, device=dev)
        b1_c, b2_c = b1.detach().contiguous(), b2.detach().contiguous()
        te, nm = C.ptr(tb["tile_expert"]), C.ptr(tb["num_mtiles"])
        C.call("moe_grouped_gemm", C.GEMM_FC1, C.ptr(xbuf), C.ptr(W1b), C.ptr(G), C.ptr(H), C.ptr(b1_c), None,
               te, nm, None, rows_cap, El, 0, h, d, st, tag="gemm_fc1")
        C.call("moe_grouped_gemm", C.GEMM_FC2, C.ptr(H), C.ptr(W2b), C.ptr(Y), None, C.ptr(b2_c), None,
               te, nm, None, rows_cap, El, 0, d, h, st, tag="gemm_fc2")

        send_y = torch.empty((W, El * slab * d), dtype=bf, device=dev)
        C.call("moe_ep_repack", C.ptr(Y), C.ptr(send_y), C.ptr(kept_recv), C.ptr(tb["slab_dst"]), C.ptr(tb["seg_start"]),
               C.ptr(tb["kept"]), W, El, slab, d, 0, st)
        ybuf = all_to_all_slabs(send_y, group, "a2a_combine_fwd").view(E * slab, d)                    # this rank's pairs, slab layout
        y = torch.empty_like(x)
        C.call("moe_combine_fwd", C.ptr(ybuf), C.ptr(r["pos"]), C.ptr(r["score"]), T, d, k, C.ptr(y), C.dtype_code(y), st)

        ctx.spec, ctx.has_bg, ctx.group, ctx.world, ctx.slab, ctx.rows_cap = spec, bg is not None, group, W, slab, rows_cap
        coef = r["aux_coef"] if spec.want_psum else torch.empty(0, dtype=torch.float32, device=dev)
        ctx.save_for_backward(x, Wg_c, r["logits"], r["idx"], r["score"], r["pos"], r["seg_start"], r["kept"], kept_recv,
                              tb["slab_dst"], tb["seg_start"], tb["kept"], tb["tile_expert"], tb["num_mtiles"], xbuf, G, H,
                              ybuf, W1b, W2b, coef)
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
        (x, Wg, logits, idx, score, pos, seg_send, kept_send, kept_recv, slab_dst, seg_loc, kept_loc, tile_expert,
         num_mtiles, xbuf, G, H, ybuf, W1b, W2b, coef) = ctx.saved_tensors
        spec, W, slab, rows_cap, group = ctx.spec, ctx.world, ctx.slab, ctx.rows_cap, ctx.group
        T, d = x.shape
        El, h = W1b.shape[0], W1b.shape[1]
        E, k = El * W, spec.top_k
        dev, st, bf = x.device, C.stream_ptr(), torch.bfloat16
        if dy is None:
            dy = torch.zeros_like(x)
        dy = _as_kernel_input(dy)
        dpsum = (coef * daux.float()).contiguous() if (spec.want_psum and daux is not None) else None

        send_dy = torch.empty((E * slab, d), dtype=bf, device=dev)                   # slab layout, pads zeroed
        dscore = _f32((T, k), dev)
        C.call("moe_combine_bwd", C.ptr(dy), C.dtype_code(dy), C.ptr(ybuf), C.ptr(pos), C.ptr(score), C.ptr(seg_send),
               C.ptr(kept_send), T, d, k, E, C.ptr(send_dy), C.ptr(dscore), st)
        recv_dy = all_to_all_slabs(send_dy.view(W, El * slab * d), group, "a2a_combine_bwd")
        dybuf = torch.empty((rows_cap, d), dtype=bf, device=dev)
        C.call("moe_ep_repack", C.ptr(recv_dy), C.ptr(dybuf), C.ptr(kept_recv), C.ptr(slab_dst), C.ptr(seg_loc),
               C.ptr(kept_loc), W, El, slab, d, 1, st)

        dU = torch.empty((rows_cap, h), dtype=bf, device=dev)
        dxbuf = torch.empty((rows_cap, d), dtype=bf, device=dev)
        dW1, db1 = _f32((El, h, d), dev), _f32((El, h), dev)
        dW2, db2 = _f32((El, d, h), dev), _f32((El, d), dev)
        te, nm, sg = C.ptr(tile_expert), C.ptr(num_mtiles), C.ptr(seg_loc)
        slab_sums = torch.empty(C.lib.moe_slab_colsum_bytes(rows_cap, h) // 4, dtype=torch.float32, device=dev)
        C.call("moe_grouped_gemm", C.GEMM_DGELU, C.ptr(dybuf), C.ptr(W2b), C.ptr(dU), C.ptr(slab_sums), None, C.ptr(G),
               te, nm, None, rows_cap, El, 0, h, d, st, tag="gemm_dgelu")
        C.call("moe_grouped_gemm", C.GEMM_DGRAD, C.ptr(dU), C.ptr(W1b), C.ptr(dxbuf), None, None, None,
               te, nm, None, rows_cap, El, 0, d, h, st, tag="gemm_dgrad")
        # dX goes back to the token owners while the weight / bias gradients (which nobody waits for) are computed:
        # the all-to-all runs on NCCL's stream, the four kernels below on ours
        send_dx = torch.empty((W, El * slab * d), dtype=bf, device=dev)
        C.call("moe_ep_repack", C.ptr(dxbuf), C.ptr(send_dx), C.ptr(kept_recv), C.ptr(slab_dst), C.ptr(seg_loc),
               C.ptr(kept_loc), W, El, slab, d, 0, st)
        dx_recv = torch.empty_like(send_dx)
        work = dist.all_to_all_single(dx_recv, send_dx, group=group, async_op=True)
        # dW2 = (H^T dY)^T: the wide dimension h is M (256-row tiles), the store is transposed
        wfl = C.ptr(C.wgrad_flags(El, h, d, dev))   # split-K flags: each tile's K range runs as two halves
        C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, C.ptr(H), C.ptr(dybuf), C.ptr(dW2), None, None, wfl,
               None, None, sg, rows_cap, El, h, d, 0, st, tag="gemm_wgrad2")
        C.call("moe_grouped_gemm", C.GEMM_WGRAD, C.ptr(dU), C.ptr(xbuf), C.ptr(dW1), None, None, wfl,
               None, None, sg, rows_cap, El, h, d, 0, st, tag="gemm_wgrad1")
        cws = torch.empty(C.lib.moe_segment_colsum_workspace_bytes(rows_cap, d), dtype=torch.uint8, device=dev)
        C.call("moe_segment_colsum", C.ptr(dybuf), sg, rows_cap, El, d, C.ptr(cws), C.ptr(db2), st, tag="colsum_db2")
        C.call("moe_slab_colsum_final", C.ptr(slab_sums), sg, El, h, C.ptr(db1), st, tag="colsum_db1")
        work.wait()
        dx_slabs = dx_recv.view(E * slab, d)

        dlogits = _f32((T, E), dev)
        dx = torch.empty_like(x)
        C.call("moe_gate_dispatch_bwd", C.ptr(dx_slabs), C.ptr(pos), C.ptr(logits), C.ptr(idx), C.ptr(score), C.ptr(dscore),
               C.ptr(dpsum), C.ptr(Wg), T, d, E, k, spec.score_mode, C.ptr(dlogits), C.ptr(dx), C.dtype_code(dx), st)
        ws = torch.empty(C.lib.moe_gate_wgrad_workspace_bytes(T, d, E), dtype=torch.uint8, device=dev)
        dWg = _f32((E, d), dev)
        dbg = _f32(E, dev) if ctx.has_bg else None
        C.call("moe_gate_wgrad", C.ptr(dlogits), C.ptr(x), C.dtype_code(x), T, d, E, C.ptr(ws), C.ptr(dWg), C.ptr(dbg), st)
        return dx, dWg, dbg, dW1, db1, dW2, db2, None, None, None, None, None, None


# How the rows travel between the ranks: "peer" = NVLink / NVSwitch peer memory, kernels write and read the owners'
# packed buffers in place (fmoe/peer.py); "nccl" = all_to_all_single on fixed slabs (this file); "auto" = peer when every
# rank sits on this node with peer access and the shape fits its kernels (E <= 64), else nccl.
TRANSPORT = "auto"


def _pick_transport(layer, x: torch.Tensor, spec: RouteSpec) -> str:
    if TRANSPORT in ("peer", "nccl"):
        return TRANSPORT
    if not x.is_cuda:
        return "nccl"
    from .peer import peer_transport_available
    E, d, k = layer.num_expert * layer.world_size, layer.d_model, spec.top_k
    fits = E <= 64 and 16 * k * (2 * d + 16) + 64 * (E + 12) * 4 <= 100 * 1024     # gate/dispatch backward stages k rows per token
    return "peer" if (fits and peer_transport_available(layer.moe_group)) else "nccl"


def ep_forward(layer, moe_inp: torch.Tensor) -> torch.Tensor:
    """`FMoE.forward` for world_size > 1."""
    gate = layer.gate
    T = moe_inp.shape[0]
    spec = gate.route_spec(T)
    unlimited = spec.capacity >= T * spec.top_k
    if not getattr(layer, "_ep_group_checked", False):
        _check_group(layer)
        layer._ep_group_checked = True
    W1, b1, W2, b2 = layer._expert_params()
    fresh = torch.is_grad_enabled() and (W1.requires_grad or W2.requires_grad)
    transport = getattr(layer, "_ep_transport", None)
    if transport is None or TRANSPORT != "auto":
        transport = layer._ep_transport = _pick_transport(layer, moe_inp, spec)
    if unlimited and transport != "peer":
        # NCCL slabs are [E, ceil256(capacity), d] per rank: without a capacity that is E times the token buffer.  The peer
        # transport packs live rows only and sizes its buffers for the worst case W * T * k rows per rank.
        raise NotImplementedError(
            f"{type(gate).__name__} has no per-expert capacity: under expert parallelism it runs on the NVLink peer-memory "
            "transport only (one node, peer access); the NCCL slab exchange needs a capacity-limited gate (SwitchGate / GShardGate)")
    if transport == "peer":
        from .peer import EPPeerMoEFunction, peer_buffers
        try:
            pb = peer_buffers(layer, T, layer.d_model, layer.num_expert * layer.world_size, layer.num_expert, spec.top_k, spec.capacity,
                              moe_inp.device)
        except C.MoeB200Error as e:
            if TRANSPORT != "auto" or unlimited:
                raise
            # the heap set-up fails on every rank together (fmoe/peer.py): all of them fall back to the NCCL exchange
            import warnings
            warnings.warn(f"fmoe: NVLink peer-memory exchange unavailable ({e}); expert parallelism falls back to NCCL all-to-all")
            transport = layer._ep_transport = "nccl"
    if transport == "peer":
        y, aux, count, kept = EPPeerMoEFunction.apply(moe_inp, gate.gate.weight, gate.gate.bias, W1, b1, W2, b2, spec,
                                                      layer._bf16_cache, gate.make_noise(moe_inp), pb, fresh, not torch.is_grad_enabled())
    else:
        y, aux, count, kept = EPMoEFunction.apply(moe_inp, gate.gate.weight, gate.gate.bias, W1, b1, W2, b2, spec,
                                                  layer._bf16_cache, gate.make_noise(moe_inp), layer.moe_group, layer.world_size, fresh)
    gate.finish(aux)
    layer.last_count, layer.last_kept = count, kept
    return y
