"""One fused add+LayerNorm forward + backward per backward variant (register kernel, staged kernel) for an ncu capture:
ncu --set full -k regex:addln_bwd python tools/ln_prof.py [d]   (library with -DMOE_EXPERIMENT_HOOKS, see tools/ln_sweep.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch, fmoe
T, d = 256 * 197, int(sys.argv[1]) if len(sys.argv) > 1 else 384
ln = fmoe.AddLayerNorm(d, eps=1e-6).cuda()
x = torch.randn(T, d, device="cuda", requires_grad=True)
delta = torch.randn(T, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
gx, gn = torch.randn(T, d, device="cuda"), torch.randn(T, d, device="cuda", dtype=torch.bfloat16)
for mode in ("0", "1"):
    os.environ["MOE_LN_BWD_MODE"] = mode
    xo, n_ = fmoe.add_layer_norm(x, delta, ln.weight, ln.bias, 1e-6, out_dtype=torch.bfloat16)
    torch.autograd.backward([xo, n_], [gx, gn])
    torch.cuda.synchronize()
print("ok")
