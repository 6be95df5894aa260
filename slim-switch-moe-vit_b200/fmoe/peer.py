"""Expert parallelism over NVLink / NVSwitch peer memory (replaces FastMoE's `expert_exchange` +
`global_scatter` / `global_gather`, the `world_size > 1` mode of the layer the reference imports at
/root/reference/models/resMoE.py:6, and round 1's NCCL all-to-all on padded slabs in `fmoe.distributed`).

Every rank of the expert-parallel group owns a symmetric heap (`PeerHeap`: `moe_ep_heap_alloc` = cudaMalloc + a CUDA
IPC handle, exchanged once with `all_gather_object`); the packed row buffers of a layer's LOCAL experts live there and
the dispatch / combine kernels of the other ranks write and read their rows in place over NVLink:

    forward   gate + scan -> counts exchange (+ barrier + layout, one kernel) -> dispatch writes kept rows straight into
              the owners' packed segments -> barrier -> expert FFN -> barrier -> combine reads remote Y rows in place
    backward  combine_bwd writes dY rows into the owners' buffers -> barrier -> expert FFN backward -> barrier ->
              gate/dispatch backward gathers remote dX rows in place

No send / receive staging, no repack passes, only live rows cross the links, nothing synchronises the host (the
barriers are device-side flag exchanges with epoch counters in device memory, so the whole step stays capturable as
one CUDA graph).  Routing is the single-GPU routing of each rank's token shard (per-source-rank capacity), so every
integer stays bit-exact against the oracle run on that shard.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _cabi as C
from .functions import Bf16WeightCache, RouteSpec, _as_kernel_input, _f32, _gate_workspace, _i32

IPC_BYTES = 64
_ALIGN = 1024


def _round_up(n: int, a: int = _ALIGN) -> int:
    return (n + a - 1) // a * a


class PeerHeap:
    """One symmetric allocation per rank, mapped into every peer.  Collective over `group`: every rank must construct
    it at the same point of the program with the same size."""

    def __init__(self, nbytes: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = int(nbytes)
        self.local, self.bases = 0, []
        # every step is followed by an agreement round, so that a failure on ONE rank (no IPC permission, out of memory,
        # no peer access) raises on EVERY rank instead of leaving the others inside a collective
        base, handle, err = ctypes.c_void_p(), ctypes.create_string_buffer(IPC_BYTES), None
        try:
            C.call("moe_ep_heap_alloc", self.nbytes, ctypes.byref(base), handle)
            self.local = int(base.value)
        except C.MoeB200Error as e:
            err = str(e)
        got = [None] * self.world
        dist.all_gather_object(got, (err, bytes(handle.raw)), group=group)
        if any(g[0] for g in got):
            self.close()
            raise C.MoeB200Error("peer heap allocation failed on rank(s) " + ", ".join(f"{r}: {g[0]}" for r, g in enumerate(got) if g[0]))
        bases, err = [], None
        for r, (_, h) in enumerate(got):
            if r == self.rank:
                bases.append(self.local)
                continue
            p = ctypes.c_void_p()
            try:
                C.call("moe_ep_heap_open", ctypes.create_string_buffer(h, IPC_BYTES), ctypes.byref(p))
                bases.append(int(p.value))
            except C.MoeB200Error as e:
                err = str(e)
                bases.append(0)
        self.bases = bases
        got = [None] * self.world
        dist.all_gather_object(got, err, group=group)   # doubles as the barrier: nobody touches a peer's heap before every mapping exists
        if any(got):
            self.close()
            raise C.MoeB200Error("peer heap mapping failed on rank(s) " + ", ".join(f"{r}: {g}" for r, g in enumerate(got) if g))

    def region(self, offset: int):
        """HOST array of W device pointers: the same region in every rank's heap (entry `rank` is the local one)."""
        arr = (ctypes.c_void_p * self.world)(*[b + offset for b in self.bases])
        return arr

    def close(self):
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                C.lib.moe_ep_heap_close(ctypes.c_void_p(b))
        if self.local:
            C.lib.moe_ep_heap_free(ctypes.c_void_p(self.local))
        self.bases, self.local = [], 0


def peer_rows_per_rank(W: int, El: int, T: int, k: int, capacity: int) -> int:
    """Rows of a rank's packed buffers: every received row could be live — W sources x El local experts x capacity rows, but
    never more than every pair of every source (W * T * k), which is the bound of a gate without a capacity (NaiveGate, the
    reference's gate: /root/reference/models/resMoE.py:27-29; worst case = every token of every rank picks this rank's
    experts).  Each segment is padded to a GEMM tile."""
    total = min(W * El * capacity, W * T * k)
    return C.rows_cap(total, 1, El, total)


class PeerBuffers:
    """The per-layer regions of a heap for one problem shape (T_local, d, E, k, capacity)."""

    def __init__(self, T: int, d: int, E: int, El: int, k: int, capacity: int, group, device):
        W = dist.get_world_size(group)
        self.W, self.rank, self.El, self.E, self.d = W, dist.get_rank(group), El, E, d
        self.rows_per_rank = peer_rows_per_rank(W, El, T, k, capacity)
        row_bytes = _round_up(self.rows_per_rank * d * 2)
        off, self.off = 0, {}
        for name, nbytes in (("flags", _round_up(W * 4)), ("kept_all", _round_up(W * E * 4)), ("xbuf", row_bytes), ("ybuf", row_bytes),
                             ("dybuf", row_bytes), ("dxbuf", row_bytes)):
            self.off[name] = off
            off += nbytes
        self.heap = PeerHeap(off, group)
        self.ptrs = {name: self.heap.region(o) for name, o in self.off.items()}
        self.local = {name: self.heap.local + o for name, o in self.off.items()}
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.key = (T, d, E, k, capacity)
        self.max_mtiles = self.rows_per_rank // C.ROW_ALIGN

    def barrier(self, st):
        C.call("moe_ep_barrier", self.ptrs["flags"], C.ptr(self.epoch), self.rank, self.W, C.ptr(self.status), st)

    def check(self):
        """Host-side check of the device status word (a sync: tests / debugging only)."""
        s = int(self.status.item())
        if s:
            raise C.MoeB200Error("expert-parallel peer exchange failed: " + {1: "a barrier timed out (a rank is missing)",
                                                                            2: "packed layout exceeds rows_per_rank"}.get(s, str(s)))


class PeerState:
    """Holder of a layer's peer buffers.  Deep copies (ModelEma, reference main.py:602-607) start empty: a heap is a
    process-local CUDA IPC mapping, never copied."""

    def __init__(self):
        self.pb = None

    def __deepcopy__(self, memo):
        return PeerState()


def peer_buffers(layer, T: int, d: int, E: int, El: int, k: int, capacity: int, device) -> PeerBuffers:
    """The layer's buffers for this shape (allocated — collectively — on first use and on a shape change)."""
    state = getattr(layer, "_peer_state", None)
    if state is None:
        state = PeerState()
        layer._peer_state = state
    pb, key = state.pb, (T, d, E, k, capacity)
    if pb is None or pb.key != key:
        if torch.cuda.is_current_stream_capturing():
            raise C.MoeB200Error("expert-parallel peer buffers must be allocated before CUDA-graph capture: run one eager step first")
        if pb is not None:
            torch.cuda.synchronize()
            dist.barrier(group=layer.moe_group)
            pb.heap.close()
        pb = state.pb = PeerBuffers(T, d, E, El, k, capacity, layer.moe_group, device)
    return pb


class EPPeerMoEFunction(torch.autograd.Function):
    """y, aux_loss, count, kept = expert-parallel MoE(x; Wg, bg, local W1, b1, W2, b2), rows exchanged through peer memory."""

    @staticmethod
    def forward(ctx, x, Wg, bg, W1, b1, W2, b2, spec: RouteSpec, cache: Bf16WeightCache, noise, pb: PeerBuffers, fresh: bool = True,
                infer: bool = False):
        x = _as_kernel_input(x)
        T, d = x.shape
        El, h = W1.shape[0], W1.shape[1]
        E, k, W = Wg.shape[0], spec.top_k, pb.W
        assert E == El * W, "gate must score world_size * num_expert experts"
        dev, st, bf = x.device, C.stream_ptr(), torch.bfloat16
        Wg_c, W1_c, W2_c = Wg.detach().contiguous(), W1.detach().contiguous(), W2.detach().contiguous()
        bg_c = None if bg is None else bg.detach().contiguous()
        rpr = pb.rows_per_rank

        # gate + scan on this rank's tokens (identical to the single-GPU routing of the shard)
        ntiles = (T + C.TOKEN_TILE - 1) // C.TOKEN_TILE
        logits, idx, score = _f32((T, E), dev), _i32((T, k), dev), _f32((T, k), dev)
        tile_hist, tile_base = _i32((E, ntiles), dev), _i32((E, ntiles), dev)
        count, kept, seg_send = _i32(E, dev), _i32(E, dev), _i32(E + 1, dev)
        mt_send = C.rows_cap(T, k, E, spec.capacity) // C.ROW_ALIGN
        te_send, nm_send = _i32(mt_send, dev), _i32(1, dev)
        tile_psum = _f32((E, ntiles), dev) if spec.want_psum else None
        psum = _f32(E, dev) if spec.want_psum else None
        aux_loss = _f32(1, dev) if spec.want_psum else None
        aux_coef = _f32(E, dev) if spec.want_psum else None
        C.call("moe_gate_fwd", C.ptr(x), C.dtype_code(x), C.ptr(Wg_c), C.ptr(bg_c), C.ptr(noise), None, T, d, E, k,
               spec.score_mode, int(spec.want_psum), C.ptr(logits), C.ptr(idx), C.ptr(score), C.ptr(tile_hist), C.ptr(tile_psum),
               C.ptr(_gate_workspace(x, E)), st)
        C.call("moe_route_scan", C.ptr(tile_hist), C.ptr(tile_psum), ntiles, E, spec.capacity, C.ptr(tile_base), C.ptr(count),
               C.ptr(kept), C.ptr(seg_send), C.ptr(te_send), C.ptr(nm_send), mt_send, C.ptr(psum), int(spec.aux_mode), T, k,
               C.ptr(aux_loss), C.ptr(aux_coef), 0, st)
        # counts to every rank, meet, lay out the owners' packed buffers
        dst_row, kept_loc, seg_loc = _i32(E, dev), _i32(El, dev), _i32(El + 1, dev)
        tile_expert, num_mtiles = _i32(pb.max_mtiles, dev), _i32(1, dev)
        C.call("moe_ep_exchange_counts", C.ptr(kept), pb.ptrs["kept_all"], pb.ptrs["flags"], C.ptr(pb.epoch), pb.rank, W, El, rpr,
               C.ptr(dst_row), C.ptr(kept_loc), C.ptr(seg_loc), C.ptr(tile_expert), C.ptr(num_mtiles), pb.max_mtiles,
               C.ptr(pb.status), st)
        pos = _i32((T, k), dev)
        C.call("moe_dispatch_fwd_peer", C.ptr(x), C.dtype_code(x), C.ptr(idx), C.ptr(tile_base), C.ptr(dst_row), T, d, E, k,
               spec.capacity, pb.ptrs["xbuf"], pb.rank, W, rpr, C.ptr(seg_loc), C.ptr(kept_loc), C.ptr(pos), st)
        W1b, W2b = cache.get(W1_c, W2_c, fresh)      # the weight casts overlap the peers' pushes
        pb.barrier(st)                                           # every source's rows have landed in my xbuf

        G = None if infer else torch.empty((rpr, h), dtype=bf, device=dev)   # forward-only pass: fc1 writes no gelu'
        H = torch.empty((rpr, h), dtype=bf, device=dev)
        b1_c, b2_c = b1.detach().contiguous(), b2.detach().contiguous()
        te, nm = C.ptr(tile_expert), C.ptr(num_mtiles)
        C.call("moe_grouped_gemm", C.GEMM_FC1, pb.local["xbuf"], C.ptr(W1b), C.ptr(G), C.ptr(H), C.ptr(b1_c), None,
               te, nm, None, rpr, El, 0, h, d, st, tag="gemm_fc1")
        C.call("moe_grouped_gemm", C.GEMM_FC2, C.ptr(H), C.ptr(W2b), pb.local["ybuf"], None, C.ptr(b2_c), None,
               te, nm, None, rpr, El, 0, d, h, st, tag="gemm_fc2")
        pb.barrier(st)                                           # every owner's Y is complete
        y = torch.empty_like(x)
        C.call("moe_combine_fwd_peer", pb.ptrs["ybuf"], pb.rank, W, rpr, C.ptr(pos), C.ptr(score), T, d, k, C.ptr(y),
               C.dtype_code(y), st)

        ctx.spec, ctx.has_bg, ctx.pb = spec, bg is not None, pb
        coef = aux_coef if spec.want_psum else torch.empty(0, dtype=torch.float32, device=dev)
        if infer:
            aux = aux_loss.reshape(()) if spec.want_psum else torch.empty(0, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(y, aux, count, kept)
            return y, aux, count, kept
        ctx.save_for_backward(x, Wg_c, logits, idx, score, pos, seg_loc, kept_loc, tile_expert, num_mtiles, G, H, W1b, W2b, coef)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(count, kept)
        if spec.want_psum:
            aux = aux_loss.reshape(())
        else:
            aux = torch.empty(0, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(aux)
        return y, aux, count, kept

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy, daux, _dcount, _dkept):
        (x, Wg, logits, idx, score, pos, seg_loc, kept_loc, tile_expert, num_mtiles, G, H, W1b, W2b, coef) = ctx.saved_tensors
        spec, pb = ctx.spec, ctx.pb
        T, d = x.shape
        El, h = W1b.shape[0], W1b.shape[1]
        W, rpr = pb.W, pb.rows_per_rank
        E, k = El * W, spec.top_k
        dev, st, bf = x.device, C.stream_ptr(), torch.bfloat16
        if dy is None:
            dy = torch.zeros_like(x)
        dy = _as_kernel_input(dy)
        dpsum = (coef * daux.float()).contiguous() if (spec.want_psum and daux is not None) else None

        dscore = _f32((T, k), dev)
        C.call("moe_combine_bwd_peer", C.ptr(dy), C.dtype_code(dy), pb.ptrs["ybuf"], pb.ptrs["dybuf"], pb.rank, W, rpr, C.ptr(pos),
               C.ptr(score), C.ptr(seg_loc), C.ptr(kept_loc), El, T, d, k, C.ptr(dscore), st)
        pb.barrier(st)                                           # every source's dY rows have landed in my dybuf

        dU = torch.empty((rpr, h), dtype=bf, device=dev)
        dW1, db1 = _f32((El, h, d), dev), _f32((El, h), dev)
        dW2, db2 = _f32((El, d, h), dev), _f32((El, d), dev)
        te, nm, sg = C.ptr(tile_expert), C.ptr(num_mtiles), C.ptr(seg_loc)
        dyb, xb, dxb = pb.local["dybuf"], pb.local["xbuf"], pb.local["dxbuf"]
        slab_sums = torch.empty(C.lib.moe_slab_colsum_bytes(rpr, h) // 4, dtype=torch.float32, device=dev)
        C.call("moe_grouped_gemm", C.GEMM_DGELU, dyb, C.ptr(W2b), C.ptr(dU), C.ptr(slab_sums), None, C.ptr(G),
               te, nm, None, rpr, El, 0, h, d, st, tag="gemm_dgelu")
        C.call("moe_grouped_gemm", C.GEMM_DGRAD, C.ptr(dU), C.ptr(W1b), dxb, None, None, None,
               te, nm, None, rpr, El, 0, d, h, st, tag="gemm_dgrad")
        pb.barrier(st)                                           # every owner's dX is complete
        # token side first: the gather of remote dX rows runs while nothing else needs the links
        dlogits = _f32((T, E), dev)
        dx = torch.empty_like(x)
        C.call("moe_gate_dispatch_bwd_peer", pb.ptrs["dxbuf"], pb.rank, W, rpr, C.ptr(pos), C.ptr(logits), C.ptr(idx), C.ptr(score),
               C.ptr(dscore), C.ptr(dpsum), C.ptr(Wg), T, d, E, k, spec.score_mode, C.ptr(dlogits), C.ptr(dx), C.dtype_code(dx), st)
        # weight / bias gradients of the local experts
        wfl = C.ptr(C.wgrad_flags(El, h, d, dev))
        C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, C.ptr(H), dyb, C.ptr(dW2), None, None, wfl,
               None, None, sg, rpr, El, h, d, 0, st, tag="gemm_wgrad2")
        C.call("moe_grouped_gemm", C.GEMM_WGRAD, C.ptr(dU), xb, C.ptr(dW1), None, None, wfl,
               None, None, sg, rpr, El, h, d, 0, st, tag="gemm_wgrad1")
        cws = torch.empty(C.lib.moe_segment_colsum_workspace_bytes(rpr, d), dtype=torch.uint8, device=dev)
        C.call("moe_segment_colsum", dyb, sg, rpr, El, d, C.ptr(cws), C.ptr(db2), st, tag="colsum_db2")
        C.call("moe_slab_colsum_final", C.ptr(slab_sums), sg, El, h, C.ptr(db1), st, tag="colsum_db1")
        ws = torch.empty(C.lib.moe_gate_wgrad_workspace_bytes(T, d, E), dtype=torch.uint8, device=dev)
        dWg = _f32((E, d), dev)
        dbg = _f32(E, dev) if ctx.has_bg else None
        C.call("moe_gate_wgrad", C.ptr(dlogits), C.ptr(x), C.dtype_code(x), T, d, E, C.ptr(ws), C.ptr(dWg), C.ptr(dbg), st)
        return dx, dWg, dbg, dW1, db1, dW2, db2, None, None, None, None, None, None


def peer_transport_available(group=None) -> bool:
    """True when every rank of the group sits on this node with peer access to every other GPU (NVLink / NVSwitch)."""
    if not torch.cuda.is_available():
        return False
    W = dist.get_world_size(group)
    if W > 8 or W > torch.cuda.device_count():
        return False
    me = torch.cuda.current_device()
    return all(o == me or torch.cuda.can_device_access_peer(me, o) for o in range(torch.cuda.device_count()))
