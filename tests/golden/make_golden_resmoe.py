"""Generates tests/golden/ref_resmoe_skip.npz.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_resmoe.py

What is pinned: the reference's OWN `forward_residule_moe`, `Gate` and `CustomizedMoEMLP`
(/root/reference/models/resMoE.py:126-145, :32-85, :15-29), imported unmodified, run on a block whose attention half
is the identity (norm1 = norm2 = Identity, attn = 0, dense_gate disabled, so the function's first half returns its
input) and whose `moe_gate` skips roughly half of the tokens (training-time threshold stepped down to 0.5, the state
`Gate.step` reaches over training).  FastMoE underneath is the CPU restatement oracle/fmoe_cpu.py (PARITY UNPINNED).
Recorded in fp32: input, upstream gradient, the (skip, keep) mask, output, dx, and the gradients of the MoE
parameters and of the gate head.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "oracle", "stubs"), "/root/reference"]

from oracle import fmoe_cpu  # noqa: E402

sys.modules["fmoe"] = fmoe_cpu
fmoe_cpu.FMoETransformerMLP.exact_logit_order = True
from models.resMoE import CustomizedMoEMLP, Gate, forward_residule_moe  # noqa: E402  (the reference's code, unmodified)
from make_golden_params import build_params, weights_digest  # noqa: E402

CASES = {"ref_resmoe_skip": (64, 256, 4, 2, 3, 50, 0.5), "ref_resmoe_skip_top1": (128, 512, 8, 1, 2, 197, 0.45)}


class _Zero(nn.Module):
    def forward(self, x):
        return torch.zeros_like(x)


def gate_head_params(d, seed=99):
    g = torch.Generator().manual_seed(seed)
    return {"head.1.weight": (torch.rand(1, d, generator=g) * 2 - 1) * d ** -0.5 * 3.0, "head.1.bias": torch.zeros(1)}


def main():
    for name, (d, hid, E, k, B, N, thr) in CASES.items():
        blk = nn.Module()
        blk.norm1, blk.norm2, blk.drop_path, blk.attn = nn.Identity(), nn.Identity(), nn.Identity(), _Zero()
        blk.dense_gate = Gate(d, 1.0)
        blk.dense_gate.disable = True
        blk.moe_gate = Gate(d, 1.0)
        blk.moe_gate.load_state_dict(gate_head_params(d), strict=False)
        blk.moe_gate._threshold.fill_(thr)
        blk.mlp = CustomizedMoEMLP(d, hid, moe_num_experts=E, moe_top_k=k, drop=0.0)
        blk.mlp.load_state_dict(build_params(d, hid, E))
        blk.train()
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, N, d, generator=g, requires_grad=True)
        dy = torch.randn(B, N, d, generator=g)
        out = forward_residule_moe(blk, x)
        (out * dy).sum().backward()
        with torch.no_grad():
            blk.moe_gate._total_tokens, blk.moe_gate._skipped_tokens = 0, 0
            mask = blk.moe_gate(x)
        rec = dict(x=x.detach().numpy(), dy=dy.numpy(), out=out.detach().numpy(), dx=x.grad.numpy(), mask=mask.numpy(),
                   meta=np.array([d, hid, E, k, B, N], dtype=np.int64), threshold=np.array(thr, dtype=np.float32),
                   weights_sha256=np.array(weights_digest({kk: v for kk, v in blk.mlp.state_dict().items()})))
        for p_name, p in blk.mlp.named_parameters():
            if p_name.startswith("experts") and p_name.endswith("weight") and E > 4:   # keep the fixture small
                rec["gradnorm.mlp." + p_name] = np.array([float(p.grad[e].norm()) for e in range(E)], dtype=np.float64)
            else:
                rec["grad.mlp." + p_name] = p.grad.numpy()
        for p_name, p in blk.moe_gate.named_parameters():
            rec["param.moe_gate." + p_name] = p.detach().numpy()
            rec["grad.moe_gate." + p_name] = p.grad.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, os.path.getsize(path) // 1024, "KiB; kept", int(mask[..., 1].detach().sum()), "of", B * N, "tokens; |out|max", float(out.detach().abs().max()))


if __name__ == "__main__":
    main()
