"""Generates tests/golden/*.npz.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is pinned: the reference's OWN wrapper class `CustomizedMoEMLP`
(/root/reference/models/resMoE.py:15-29 — constructor mapping, activation composition) is imported
unmodified and instantiated exactly as the model factories do (models/resMoE.py:200-208), on top of
the CPU restatement of FastMoE in oracle/fmoe_cpu.py (FastMoE itself is absent: PARITY UNPINNED).
Inputs, parameters, routing integers, outputs and gradients are recorded in fp32.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "oracle", "stubs"), "/root/reference"]

from oracle import fmoe_cpu, moe_oracle as O  # noqa: E402

sys.modules["fmoe"] = fmoe_cpu
fmoe_cpu.FMoETransformerMLP.exact_logit_order = True
from models.resMoE import CustomizedMoEMLP  # noqa: E402  (the reference's class, unmodified)

CASES = {
    # name: (in_features, hidden, experts, top_k, batch, tokens, store_weights)
    "ref_wrapper_small": (64, 256, 4, 2, 3, 50, True),
    "ref_wrapper_tiny_e8_top2": (192, 768, 8, 2, 2, 197, False),   # the configuration the reference ships
    "ref_wrapper_tiny_e8_top1": (192, 768, 8, 1, 8, 197, False),   # BASELINE config 1 layer shape
}
from make_golden_params import build_params, weights_digest  # noqa: E402


def main():
    for name, (d, hid, E, k, B, N, store_w) in CASES.items():
        layer = CustomizedMoEMLP(d, hid, moe_num_experts=E, moe_top_k=k, drop=0.0)
        layer.load_state_dict(build_params(d, hid, E))
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, N, d, generator=g, requires_grad=True)
        dy = torch.randn(B, N, d, generator=g)
        y = layer(x)
        (y * dy).sum().backward()
        sd = {kk: v.detach().clone() for kk, v in layer.state_dict().items()}
        logits = O.gate_logits(x.detach().reshape(-1, d), sd["gate.gate.weight"], sd["gate.gate.bias"])
        r = O.route(logits, k, O.SCORE_TOPK_SOFTMAX, B * N * k)
        out = dict(x=x.detach().numpy(), dy=dy.numpy(), y=y.detach().numpy(), dx=x.grad.numpy(),
                   logits=logits.numpy(), idx=r.idx.numpy(), score=r.score.numpy(), pos=r.pos.numpy(),
                   count=r.count.numpy(), meta=np.array([d, hid, E, k, B, N], dtype=np.int64),
                   weights_sha256=np.array(weights_digest(sd)))
        for p_name, p in layer.named_parameters():
            big = p_name.startswith("experts") and p_name.endswith("weight") and not store_w
            if big:   # keep the fixture small: per-expert Frobenius norms instead of the full gradient
                out["gradnorm." + p_name] = np.array([float(p.grad[e].norm()) for e in range(E)], dtype=np.float64)
            else:
                out["grad." + p_name] = p.grad.numpy()
            if store_w:
                out["param." + p_name] = p.detach().numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", "y_absmax", float(y.abs().max()))


if __name__ == "__main__":
    main()
