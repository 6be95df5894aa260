"""Device time of the fused add+LayerNorm kernels at the config-2 shape (CUDA-graph replay of 20 calls)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch, fmoe
T, d = 256 * 197, int(sys.argv[1]) if len(sys.argv) > 1 else 384
_main = torch.cuda.Stream()          # keep every tensor (and AccumulateGrad node) off the legacy default stream: capture needs it
torch.cuda.set_stream(_main)
ln = fmoe.AddLayerNorm(d, eps=1e-6).cuda()
x = torch.randn(T, d, device="cuda", requires_grad=True)
delta = torch.randn(T, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
gx, gn = torch.randn(T, d, device="cuda"), torch.randn(T, d, device="cuda", dtype=torch.bfloat16)
def fwd():
    return fmoe.add_layer_norm(x, delta, ln.weight, ln.bias, 1e-6, out_dtype=torch.bfloat16)
def time_graph(fn, n=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
t_f = time_graph(fwd)
xo, n_ = fwd()
def both():
    xo, n_ = fwd()
    torch.autograd.backward([xo, n_], [gx, gn])
    x.grad = None; delta.grad = None; ln.weight.grad = None; ln.bias.grad = None
t_b = time_graph(both) - t_f
fb, bb = T * d * 12, T * d * 16
print(f"d={d}: fwd {t_f:.1f} us ({fb / t_f / 1e3:.0f} GB/s)   bwd (incl. reduce) {t_b:.1f} us ({bb / t_b / 1e3:.0f} GB/s)")
