"""Expert-parallel parity worker: run under torch.distributed.run with W ranks (one GPU each, NCCL).
Every rank checks its own outputs / gradients against the single-process CPU oracle evaluated with
ALL experts on that rank's token shard (per-source-rank capacity keeps the routing identical).
usage: ep_worker.py [peer|nccl]  — how the rows travel: NVLink peer memory (fmoe/peer.py, default) or NCCL all-to-all on
slabs (fmoe/distributed.py).  Prints `EP_OK rank=<r>` on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "slim-switch-moe-vit_b200"), os.path.dirname(os.path.abspath(__file__))]
import torch
import torch.distributed as dist

from _util import make_problem, rel_err
from oracle import moe_oracle as O


def main():
    rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from fmoe import _cabi as C
    from fmoe import functions as Fn
    from fmoe.distributed import EPMoEFunction
    from fmoe.peer import EPPeerMoEFunction, PeerBuffers, peer_transport_available
    transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
    if transport == "peer":
        assert peer_transport_available(None), "peer transport needs all ranks on one NVLink node"

    cases = [(1000, 128, 256, 8, 1, 1, 1.25, C.AUX_SWITCH), (777, 192, 768, 4 * W, 2, 0, 1.0, C.AUX_GSHARD),
             (2500, 384, 1536, 16, 1, 1, 1.25, C.AUX_SWITCH)]
    if transport == "peer":   # NaiveGate (no capacity, top-2: the reference's own gate, models/resMoE.py:27-29), skewed routing
        cases.append((900, 192, 768, 8, 2, 0, 0.0, C.AUX_NONE))
    for (T, d, h, E, k, mode, cf, aux_mode) in cases:
        El = E // W
        # the same global problem on every rank (seeded), each rank's tokens seeded by its rank
        _, Wg, bg, W1, b1, W2, b2 = make_problem(T, d, h, E, seed=11, skew=1.0)
        xs, dys = [], []
        for r in range(W):
            g = torch.Generator().manual_seed(100 + r)
            xs.append(torch.randn(T, d, generator=g))
            dys.append(torch.randn(T, d, generator=g))
        cap = O.capacity_from_factor(cf, T, k, E)
        spec = Fn.RouteSpec(k, mode, cap, aux_mode)
        sl = slice(rank * El, (rank + 1) * El)
        dev = [t.cuda().requires_grad_() for t in (xs[rank], Wg, bg, W1[sl].contiguous(), b1[sl].contiguous(),
                                                   W2[sl].contiguous(), b2[sl].contiguous())]
        if transport == "peer":
            pb = PeerBuffers(T, d, E, El, k, cap, None, torch.device("cuda", torch.cuda.current_device()))
            y, aux, count, kept = EPPeerMoEFunction.apply(*dev, spec, Fn.Bf16WeightCache(), None, pb)
        else:
            y, aux, count, kept = EPMoEFunction.apply(*dev, spec, Fn.Bf16WeightCache(), None, None, W)
        aux_w = 0.37
        has_aux = aux_mode != C.AUX_NONE     # NaiveGate: no load-balancing loss (an empty tensor comes back)
        ((y * dys[rank].cuda()).sum() + (aux_w * aux if has_aux else 0.0)).backward()
        torch.cuda.synchronize()
        if transport == "peer":
            pb.check()          # no barrier timed out, every packed layout fitted
            # a second forward + backward on the same buffers (epochs keep counting; what a training loop / graph replay does)
            for t_ in dev:
                t_.grad = None
            y2, aux2, _, _ = EPPeerMoEFunction.apply(*dev, spec, Fn.Bf16WeightCache(), None, pb)
            ((y2 * dys[rank].cuda()).sum() + (aux_w * aux2 if has_aux else 0.0)).backward()
            torch.cuda.synchronize()
            pb.check()
            assert torch.equal(y2, y), "second pass over the same peer buffers must reproduce the first bit for bit"

        exp = {}
        for r in range(W):   # oracle of every rank's shard (expert gradients sum over the source ranks)
            ym, sv = O.forward_model(xs[r], Wg, bg, W1, b1, W2, b2, k, mode, cap)
            coef = O.aux_coef(sv.r, T, aux_mode)
            gm = O.backward_model(sv, dys[r], Wg, aux_w * coef)
            for n in ("dW1", "db1", "dW2", "db2"):
                exp[n] = exp.get(n, 0) + gm[n]
            if r == rank:
                mine, mine_y, mine_sv, mine_coef = gm, ym, sv, coef
        assert torch.equal(count.cpu(), mine_sv.r.count) and torch.equal(kept.cpu(), mine_sv.r.kept), "routing counts"
        assert rel_err(y, mine_y) <= 3e-3, f"y {rel_err(y, mine_y)}"
        assert not has_aux or abs(float(aux) - float((mine_coef * mine_sv.r.psum).sum())) <= 1e-5
        for name, t, want in (("dx", dev[0], mine["dx"]), ("dWg", dev[1], mine["dWg"]), ("dbg", dev[2], mine["dbg"]),
                              ("dW1", dev[3], exp["dW1"][sl]), ("db1", dev[4], exp["db1"][sl]),
                              ("dW2", dev[5], exp["dW2"][sl]), ("db2", dev[6], exp["db2"][sl])):
            e = rel_err(t.grad, want)
            assert e <= 6e-3, f"{name}: {e} (T={T} E={E} k={k})"
        dropped = int((mine_sv.r.pos < 0).sum())
        if rank == 0:
            print(f"case T={T} d={d} E={E} k={k} transport={transport} W={W}: ok (dropped pairs on rank 0: {dropped})", flush=True)
        if transport == "peer":
            dist.barrier()
            pb.heap.close()
    dist.barrier()
    print(f"EP_OK rank={rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
