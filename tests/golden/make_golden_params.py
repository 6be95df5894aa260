"""Deterministic parameter builder shared by make_golden.py and the tests (no reference needed)."""
import hashlib

import torch

SEED = 20261018


def weights_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(sd[k].detach().numpy().tobytes())
    return h.hexdigest()


def build_params(d, hid, E, seed=SEED):
    """Deterministic parameters from torch's CPU generator (shared with the tests, which rebuild
    them and verify the sha256 recorded in the fixture)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, a: (torch.rand(*s, generator=g) * 2 - 1) * a
    return {
        "gate.gate.weight": u(E, d, a=d ** -0.5), "gate.gate.bias": u(E, a=d ** -0.5),
        "experts.htoh4.weight": u(E, hid, d, a=d ** -0.5), "experts.htoh4.bias": u(E, hid, a=0.1),
        "experts.h4toh.weight": u(E, d, hid, a=hid ** -0.5), "experts.h4toh.bias": u(E, d, a=0.1),
    }
