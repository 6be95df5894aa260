"""BASELINE configs[4]: isolated MoE-layer sweep, tokens x experts x hidden, phases timed separately.

    python tools/layer_sweep.py [--quick] > profiles/rNN_layer_sweep.md

Per point: fwd+bwd time of the layer replayed as a CUDA graph (pure device time), tokens/s, the expert-FFN
TFLOP/s (12*R*d*h over the summed GEMM launches) and the per-phase CUDA-event times of an eager pass
(gate / scan / dispatch / GEMMs / combine / backward kernels)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "slim-switch-moe-vit_b200")]
import torch
import fmoe
from fmoe import _cabi as C


def run_point(T, d, E, k, cf, iters):
    h = 4 * d

    class Sw(fmoe.SwitchGate):
        def __init__(self, d_model, num_expert, world_size, top_k):
            super().__init__(d_model, num_expert, world_size, topk=top_k, switch_eps=0.0, capacity=(cf, cf))

    class Gs(fmoe.GShardGate):
        def __init__(self, d_model, num_expert, world_size, top_k):
            super().__init__(d_model, num_expert, world_size, topk=top_k, capacity=(cf, cf))

    torch.manual_seed(0)
    layer = fmoe.FMoETransformerMLP(E, d, h, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0)), top_k=k,
                                    gate=Sw if k == 1 else Gs).cuda()
    x = torch.randn(T, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(T, d, device="cuda", dtype=torch.bfloat16)
    one = torch.full((), 0.01, device="cuda")

    def it():
        y = layer(x)
        aux = layer.gate.get_loss()
        torch.autograd.backward([y, aux], [dy, one])

    def clear():
        x.grad = None
        for p in layer.parameters():
            p.grad = None

    for _ in range(3):
        it(); clear()
    torch.cuda.synchronize()
    C.PROF.reset(); C.PROF.enabled = True
    for _ in range(5):
        torch.cuda._sleep(int(6e-3 * 1.9e9))   # the device starts ~6 ms behind the host: event pairs bracket kernels, not launch gaps
        it(); clear()
    torch.cuda.synchronize()
    C.PROF.enabled = False
    ph = {t: ms for t, (n, ms) in C.PROF.summary_ms().items()}
    R = int(layer.last_kept.sum())
    # graph replay for the total
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        it()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        it()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    gemm = sum(v for t, v in ph.items() if t.startswith("gemm_"))
    grp = lambda *names: sum(ph.get(n, 0.0) for n in names) * 1e3
    return dict(T=T, d=d, E=E, k=k, R=R, ms=ms, mtok=T / ms / 1e3, ffn_tf=12.0 * R * d * h / (gemm * 1e-3) / 1e12,
                gate=grp("moe_gate_fwd", "moe_route_scan"), dispatch=grp("moe_dispatch_fwd"), combine=grp("moe_combine_fwd"),
                gemm_fwd=grp("gemm_fc1", "gemm_fc2"), gemm_bwd=grp("gemm_dgelu", "gemm_dgrad", "gemm_wgrad1", "gemm_wgrad2"),
                bwd_other=grp("moe_combine_bwd", "colsum_db1", "colsum_db2", "moe_gate_dispatch_bwd", "moe_gate_wgrad"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    pts = []
    Ts = [4096, 16384, 65536, 262144]
    for d in (384, 768, 1024):
        for E in (8, 16, 32, 64):
            for T in Ts:
                if a.quick and not ((T, E) in ((16384, 16), (65536, 16), (65536, 64), (262144, 32))):
                    continue
                if T * d * 4 * 2 * 12 > 60e9:     # keep the footprint of the saved activations bounded
                    continue
                pts.append((T, d, E, 1))
    pts += [(65536, 768, 32, 2), (16384, 768, 32, 2)]
    print("| T | d | E | k | kept R | fwd+bwd ms (graph) | Mtok/s | FFN TFLOP/s | gate+scan us | dispatch us | GEMM fwd us | combine us | GEMM bwd us | other bwd us |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (T, d, E, k) in pts:
        try:
            r = run_point(T, d, E, k, 1.25, 10)
            print(f"| {r['T']} | {r['d']} | {r['E']} | {r['k']} | {r['R']} | {r['ms']:.3f} | {r['mtok']:.1f} | {r['ffn_tf']:.0f} | {r['gate']:.0f} | "
                  f"{r['dispatch']:.0f} | {r['gemm_fwd']:.0f} | {r['combine']:.0f} | {r['gemm_bwd']:.0f} | {r['bwd_other']:.0f} |", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"| {T} | {d} | {E} | {k} | error: {type(e).__name__}: {str(e)[:80]} |", flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
