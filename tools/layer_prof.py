"""One forward+backward of the isolated MoE layer at the BASELINE configs[1] layer shape (for ncu).
usage: python tools/layer_prof.py [T] [d] [E] [k] [dtype: f32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch
import fmoe

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50432
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
E = int(sys.argv[3]) if len(sys.argv) > 3 else 16
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dt = torch.bfloat16 if (len(sys.argv) > 5 and sys.argv[5] == "bf16") else torch.float32


class Gate(fmoe.SwitchGate):
    def __init__(self, d_model, num_expert, world_size, top_k):
        super().__init__(d_model, num_expert, world_size, topk=top_k, switch_eps=0.0, capacity=(1.25, 1.25))


torch.manual_seed(0)
gate = Gate if k == 1 else fmoe.NaiveGate
layer = fmoe.FMoETransformerMLP(E, d, 4 * d, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0)), top_k=k, gate=gate).cuda()
x = torch.randn(T, d, device="cuda", dtype=dt, requires_grad=True)
dy = torch.randn(T, d, device="cuda", dtype=dt)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    y = layer(x)
    aux = layer.gate.get_loss()
    torch.autograd.backward([y, aux], [dy, torch.ones_like(aux) * 0.01])
    x.grad = None
    for p in layer.parameters():
        p.grad = None
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
