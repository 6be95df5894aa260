"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): one forward+backward of the
layer for each gate kind at sizes that exercise ragged segments, empty experts, drops and the pad-row zeroing,
plus the fused add+LayerNorm kernels.  usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "slim-switch-moe-vit_b200")]
import torch
import fmoe


def gate_cls(kind, cf):
    if kind == "naive":
        return fmoe.NaiveGate
    base = fmoe.SwitchGate if kind == "switch" else fmoe.GShardGate

    class G(base):
        def __init__(self, d_model, num_expert, world_size, top_k):
            kw = dict(switch_eps=0.0) if kind == "switch" else {}
            super().__init__(d_model, num_expert, world_size, topk=top_k, capacity=(cf, cf), **kw)
    return G


torch.manual_seed(0)
for (T, d, E, k, kind, cf, dt) in [(333, 64, 5, 2, "naive", 0.0, torch.float32), (515, 192, 8, 1, "switch", 1.0, torch.bfloat16),
                                   (260, 128, 12, 2, "gshard", 1.0, torch.float32)]:
    layer = fmoe.FMoETransformerMLP(E, d, 4 * d, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0)), top_k=k,
                                    gate=gate_cls(kind, cf)).cuda()
    with torch.no_grad():
        layer.gate.gate.bias[0] += 3.0          # overload expert 0 -> drops; the last expert mostly empty
        layer.gate.gate.bias[-1] -= 10.0
    ln = fmoe.AddLayerNorm(d).cuda()
    x = torch.randn(T, d, device="cuda", requires_grad=True)
    delta = torch.randn(T, d, device="cuda", dtype=dt, requires_grad=True)
    xo, n = fmoe.add_layer_norm(x, delta, ln.weight, ln.bias, 1e-6, out_dtype=dt)
    y = layer(n)
    loss = y.float().square().mean() + xo.mean() + layer.gate.get_loss().sum()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(p.grad).all() for p in layer.parameters())
    print(f"ok T={T} d={d} E={E} k={k} {kind}", flush=True)
print("SANITIZE_CASE_OK")
