"""Sweep of the fused add+LayerNorm backward variants in ONE process (needs a library built with -DMOE_EXPERIMENT_HOOKS:
tools/build_variant.sh --all lnhooks -DMOE_EXPERIMENT_HOOKS; MOE_B200_LIB=tools/variants/libmoe_lnhooks.so python tools/ln_sweep.py).
The hooks are read at every launch, so the environment is switched between measurements.  Device time per call from a CUDA-graph
replay of 20 forward+backward pairs minus 20 forwards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch, fmoe

_main = torch.cuda.Stream()
torch.cuda.set_stream(_main)


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
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n * 1e3)
    return best


def run(T, d, configs):
    ln = fmoe.AddLayerNorm(d, eps=1e-6).cuda()
    x = torch.randn(T, d, device="cuda", requires_grad=True)
    delta = torch.randn(T, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    gx, gn = torch.randn(T, d, device="cuda"), torch.randn(T, d, device="cuda", dtype=torch.bfloat16)

    def fwd():
        return fmoe.add_layer_norm(x, delta, ln.weight, ln.bias, 1e-6, out_dtype=torch.bfloat16)

    def both():
        xo, n_ = fwd()
        torch.autograd.backward([xo, n_], [gx, gn])
        x.grad = None; delta.grad = None; ln.weight.grad = None; ln.bias.grad = None

    t_f = time_graph(fwd)
    print(f"T={T} d={d}: fwd {t_f:.1f} us ({T * d * 12 / t_f / 1e3:.0f} GB/s)", flush=True)
    for env in configs:
        for k in ("MOE_LN_BWD_MODE", "MOE_LN_CTAS", "MOE_LN_BUDGET_KB", "MOE_LN_STAGES"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t_b = time_graph(both) - t_f
        print(f"    {str(env):70s} bwd (incl. reduce) {t_b:6.1f} us  {T * d * 16 / t_b / 1e3:5.0f} GB/s", flush=True)


if __name__ == "__main__":
    reg = {"MOE_LN_BWD_MODE": "0"}
    run(256 * 197, 384, [reg, {}, {"MOE_LN_STAGES": "4"}, {"MOE_LN_STAGES": "3"}, {"MOE_LN_STAGES": "2"},
                         {"MOE_LN_CTAS": "2"}, {"MOE_LN_CTAS": "2", "MOE_LN_STAGES": "2"}, {"MOE_LN_BUDGET_KB": "220"}, reg, {}])
    run(128 * 197, 768, [reg, {}, {"MOE_LN_STAGES": "2"}])
    run(128 * 197, 1024, [reg, {}])
    run(256 * 197, 192, [reg, {}, {"MOE_LN_CTAS": "2"}])
