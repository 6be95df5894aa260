"""Shared helpers for the parity tests: seeded problem generators and thin wrappers that call the
C ABI (through fmoe._cabi) on torch CUDA tensors."""
import math

import torch


def make_problem(T, d, h, E, seed=0, x_dtype=torch.float32, skew=0.0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, d, generator=g)
    Wg = (torch.rand(E, d, generator=g) * 2 - 1) / math.sqrt(d)
    bg = (torch.rand(E, generator=g) * 2 - 1) / math.sqrt(d)
    if skew:
        bg[: max(1, E // 8)] += skew
    W1 = (torch.rand(E, h, d, generator=g) * 2 - 1) / math.sqrt(d)
    b1 = (torch.rand(E, h, generator=g) * 2 - 1) * 0.1
    W2 = (torch.rand(E, d, h, generator=g) * 2 - 1) / math.sqrt(h)
    b2 = (torch.rand(E, d, generator=g) * 2 - 1) * 0.1
    x = x.to(x_dtype)
    return tuple(t.to(device) for t in (x, Wg, bg, W1, b1, W2, b2))


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def max_abs(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


class TokenGate(torch.nn.Module):
    """Test-side restatement of the reference's token-skip `Gate` in its hard mode
    (/root/reference/models/resMoE.py:60-84): returns [B, N, 2] = (skip, keep) 0/1 weights with the straight-through
    terms `+ p.detach() - p`.  Pinned by the `mask` array of tests/golden/ref_resmoe_skip*.npz."""

    def __init__(self, d, threshold):
        super().__init__()
        self.head = torch.nn.Sequential(torch.nn.Dropout(p=0.0), torch.nn.Linear(d, 1))
        self.threshold = float(threshold)
        self.is_hard, self.disable = True, False

    def forward(self, x):
        prob = torch.sigmoid(self.head(x))
        inv = 1 - prob
        skip = (prob > self.threshold).float() + inv.detach() - inv
        keep = (prob <= self.threshold).float() + prob.detach() - prob
        return torch.cat([skip, keep], dim=-1)
