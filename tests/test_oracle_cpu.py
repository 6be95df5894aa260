"""CPU tests (no GPU): the oracle against the committed golden vectors and against itself.

The golden fixtures were produced by the reference's own `CustomizedMoEMLP`
(/root/reference/models/resMoE.py:15-29) on top of oracle/fmoe_cpu.py — see tests/golden/make_golden.py.
"""
import glob
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from _util import make_problem, rel_err
from oracle import moe_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
from make_golden_params import build_params, weights_digest  # noqa: E402


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, hid, E, k, B, N = [int(v) for v in z["meta"]]
    if "param.gate.gate.weight" in z.files:
        sd = {n[6:]: torch.from_numpy(z[n]) for n in z.files if n.startswith("param.")}
    else:
        sd = build_params(d, hid, E)
    assert weights_digest(sd) == str(z["weights_sha256"]), "torch CPU RNG drifted: regenerate the fixtures"
    return z, sd, (d, hid, E, k, B, N)


GOLDEN_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_wrapper_*.npz")))
RESMOE_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_resmoe_*.npz")))


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_golden(name):
    z, sd, (d, hid, E, k, B, N) = load_golden(name)
    T = B * N
    x = torch.from_numpy(z["x"]).reshape(T, d)
    logits = O.gate_logits(x, sd["gate.gate.weight"], sd["gate.gate.bias"])
    assert np.array_equal(logits.numpy(), z["logits"])            # bit-exact
    r = O.route(logits, k, O.SCORE_TOPK_SOFTMAX, T * k)
    for f in ("idx", "pos", "count"):
        assert np.array_equal(getattr(r, f).numpy(), z[f]), f      # bit-exact integers
    assert np.abs(r.score.numpy() - z["score"]).max() <= 1e-6
    args = (x, sd["gate.gate.weight"], sd["gate.gate.bias"], sd["experts.htoh4.weight"], sd["experts.htoh4.bias"],
            sd["experts.h4toh.weight"], sd["experts.h4toh.bias"])
    # fp64 ideal vs the fp32 golden: tight
    xs = [t.clone().double().requires_grad_() for t in args]
    yi, _ = O.ideal_forward(*xs, r, O.SCORE_TOPK_SOFTMAX)
    (yi * torch.from_numpy(z["dy"]).reshape(T, d).double()).sum().backward()
    assert rel_err(yi, torch.from_numpy(z["y"]).reshape(T, d)) <= 1e-5
    assert rel_err(xs[0].grad, torch.from_numpy(z["dx"]).reshape(T, d)) <= 1e-4
    assert rel_err(xs[1].grad, torch.from_numpy(z["grad.gate.gate.weight"])) <= 1e-4
    assert rel_err(xs[4].grad, torch.from_numpy(z["grad.experts.htoh4.bias"])) <= 1e-4
    # bf16 arithmetic model vs the fp32 golden: bf16-level tolerance (2e-2, as for the CUDA path)
    ym, sv = O.forward_model(*args, k, O.SCORE_TOPK_SOFTMAX, T * k)
    assert rel_err(ym, torch.from_numpy(z["y"]).reshape(T, d)) <= 2e-2
    gm = O.backward_model(sv, torch.from_numpy(z["dy"]).reshape(T, d), sd["gate.gate.weight"])
    assert rel_err(gm["dx"], torch.from_numpy(z["dx"]).reshape(T, d)) <= 2e-2
    if "gradnorm.experts.htoh4.weight" in z.files:
        got = torch.stack([gm["dW1"][e].norm() for e in range(E)]).double()
        assert rel_err(got, torch.from_numpy(z["gradnorm.experts.htoh4.weight"])) <= 2e-2
    else:
        assert rel_err(gm["dW1"], torch.from_numpy(z["grad.experts.htoh4.weight"])) <= 2e-2
        assert rel_err(gm["dW2"], torch.from_numpy(z["grad.experts.h4toh.weight"])) <= 2e-2


@pytest.mark.parametrize("k,mode,cf", [(1, 0, 0.0), (2, 0, 0.0), (1, 1, 1.25), (2, 0, 1.0), (2, 1, 0.5)])
def test_c_routing_matches_python_restatement(k, mode, cf):
    T, d, E = 333, 64, 8
    x, Wg, bg, *_ = make_problem(T, d, 128, E, seed=11, skew=1.0)
    logits = O.gate_logits(x, Wg, bg)
    cap = O.capacity_from_factor(cf, T, k, E)
    a, b = O.route(logits, k, mode, cap), O.route_python(logits, k, mode, cap)
    for f in ("idx", "count", "kept", "seg_start", "pos"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert (a.score - b.score).abs().max() <= 1e-6
    assert (a.psum - b.psum).abs().max() <= 1e-4
    if cf:
        assert int(a.kept.max()) <= cap and (a.pos < 0).any()


@pytest.mark.parametrize("k,mode,cf", [(1, 0, 0.0), (2, 0, 0.0), (1, 1, 1.25), (2, 1, 0.5)])
def test_c_routing_with_token_mask_matches_python_restatement(k, mode, cf):
    """Token-skip mask: skipped tokens are never routed (idx = pos = -1), take no capacity and add nothing to psum."""
    T, d, E = 333, 64, 8
    x, Wg, bg, *_ = make_problem(T, d, 128, E, seed=12, skew=1.0)
    mask = (torch.rand(T, generator=torch.Generator().manual_seed(5)) < 0.6)
    mask[64:128] = False                       # a whole 64-token routing tile without live tokens
    logits = O.gate_logits(x, Wg, bg)
    cap = O.capacity_from_factor(cf, T, k, E)
    a, b = O.route(logits, k, mode, cap, token_mask=mask), O.route_python(logits, k, mode, cap, token_mask=mask)
    for f in ("idx", "count", "kept", "seg_start", "pos"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert (a.score - b.score).abs().max() <= 1e-6 and (a.psum - b.psum).abs().max() <= 1e-4
    assert (a.idx[~mask] == -1).all() and (a.pos[~mask] == -1).all() and (a.score[~mask] == 0).all()
    assert int(a.count.sum()) == int(mask.sum()) * k
    # the kept tokens route exactly as they would in a batch that only holds them (capacity aside)
    sub = O.route(logits[mask], k, mode, T * k)
    assert torch.equal(a.idx[mask], sub.idx) and torch.equal(a.count, sub.count)


def load_resmoe(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, hid, E, k, B, N = [int(v) for v in z["meta"]]
    sd = build_params(d, hid, E)
    assert weights_digest(sd) == str(z["weights_sha256"]), "torch CPU RNG drifted: regenerate the fixtures"
    return z, sd, (d, hid, E, k, B, N)


@pytest.mark.parametrize("name", RESMOE_NAMES)
def test_skip_formulation_matches_reference_residual_moe(name):
    """The product never routes skipped tokens and restores mlp(0) / the J0^T dy input gradient analytically
    (fmoe.functions.zero_token_path + SkipFill, pure torch).  Here that formulation is evaluated on the CPU with the
    oracle layer standing in for the CUDA kernels, and compared with what the reference's own
    `forward_residule_moe` + `Gate` produced (tests/golden/make_golden_resmoe.py)."""
    from _util import TokenGate
    from fmoe.functions import SkipFill, zero_token_path
    from oracle import fmoe_cpu
    z, sd, (d, hid, E, k, B, N) = load_resmoe(name)
    T = B * N
    layer = fmoe_cpu.FMoETransformerMLP(E, d, hid, torch.nn.GELU(), top_k=k)
    layer.load_state_dict(sd)
    gate = TokenGate(d, float(z["threshold"]))
    gate.head[1].weight.data.copy_(torch.from_numpy(z["param.moe_gate.head.1.weight"]))
    gate.head[1].bias.data.copy_(torch.from_numpy(z["param.moe_gate.head.1.bias"]))
    x = torch.from_numpy(z["x"]).clone().requires_grad_()
    w = gate(x)
    assert np.array_equal(w.detach().numpy(), z["mask"])                         # the Gate restatement is exact
    keep = (w[..., 1] != 0).reshape(T)
    assert 0 < int(keep.sum()) < T
    tk, skip_tk = x * w[..., 1:2], x * w[..., 0:1]
    tk2 = tk.reshape(T, d)
    y = torch.zeros(T, d).index_put((keep.nonzero().squeeze(1),), layer(tk2[keep]))   # kept tokens only
    g, e = layer.gate.gate, layer.experts
    c, J0 = zero_token_path(g.weight, g.bias, e.htoh4.weight, e.htoh4.bias, e.h4toh.weight, e.h4toh.bias, k, 0)
    out = SkipFill.apply(y, tk2, keep.to(torch.uint8), c, J0).reshape(B, N, d) + tk + skip_tk
    (out * torch.from_numpy(z["dy"])).sum().backward()
    assert rel_err(out, torch.from_numpy(z["out"])) <= 1e-5
    assert rel_err(x.grad, torch.from_numpy(z["dx"])) <= 1e-4
    assert rel_err(gate.head[1].weight.grad, torch.from_numpy(z["grad.moe_gate.head.1.weight"])) <= 1e-4
    assert rel_err(gate.head[1].bias.grad, torch.from_numpy(z["grad.moe_gate.head.1.bias"])) <= 1e-4
    for pn, p in layer.named_parameters():
        if "grad.mlp." + pn in z.files:
            assert rel_err(p.grad, torch.from_numpy(z["grad.mlp." + pn])) <= 1e-4, pn
        else:
            got = torch.stack([p.grad[i].norm() for i in range(E)]).double()
            assert rel_err(got, torch.from_numpy(z["gradnorm.mlp." + pn])) <= 1e-4, pn


def test_ties_resolve_to_lowest_index_and_edge_cases():
    logits = torch.tensor([[1.0, 3.0, 3.0, 3.0], [0.0, 0.0, 0.0, 0.0], [-1.0, -1.0, 5.0, 5.0]])
    r = O.route(logits, 2, 0, 6)
    assert r.idx.tolist() == [[1, 2], [0, 1], [2, 3]]
    assert torch.allclose(r.score, torch.full((3, 2), 0.5))
    # capacity 1: only the first pair of each expert (token order) survives
    r = O.route(logits, 2, 0, 1)
    A = O.ALIGN
    assert r.pos.tolist() == [[A, 2 * A], [0, -1], [-1, 3 * A]]
    assert r.kept.tolist() == [1, 1, 1, 1] and r.seg_start.tolist() == [0, A, 2 * A, 3 * A, 4 * A]
    # one token, one expert
    r = O.route(torch.zeros(1, 1), 1, 1, 1)
    assert r.idx.tolist() == [[0]] and r.score.tolist() == [[1.0]] and r.rows == O.ALIGN


def test_logit_order_is_what_the_header_says():
    """LOGIT ORDER v1 restated a third time in numpy (float32 ops; exact because x,w are bf16-valued)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 192, generator=g).bfloat16().float()
    w = torch.randn(5, 192, generator=g).bfloat16().float()
    b = torch.randn(5, generator=g)
    got = O.gate_logits(x, w, b).numpy()
    xn, wn = x.numpy(), w.numpy()
    want = np.zeros((37, 5), np.float32)
    for t in range(37):
        for e in range(5):
            part = np.zeros(32, np.float32)
            for i in range(192):
                l = (i // 4) % 32
                part[l] = np.float32(part[l] + np.float32(xn[t, i] * wn[e, i]))   # product of bf16 values is exact
            for off in (16, 8, 4, 2, 1):
                part = (part + part[np.arange(32) ^ off]).astype(np.float32)
            want[t, e] = np.float32(part[0] + b[e].item())
    assert np.array_equal(got, want)


def test_model_ideal_and_bruteforce_agree():
    for (k, mode, cf) in [(2, 0, 0.0), (1, 1, 1.25), (2, 0, 0.75)]:
        T, d, h, E = 150, 64, 128, 4
        prob = make_problem(T, d, h, E, seed=21, skew=0.5)
        cap = O.capacity_from_factor(cf, T, k, E)
        ym, sv = O.forward_model(*prob, k, mode, cap)
        yi, _ = O.ideal_forward(*prob, sv.r, mode)
        yb = O.bruteforce_forward(*prob, k, mode, cap)
        assert rel_err(yi, yb) <= 1e-5          # two independent fp64 restatements
        assert rel_err(ym, yi) <= 1e-2          # bf16 arithmetic model vs ideal


@pytest.mark.parametrize("T,E,cf,seed", [(1576, 8, 1.25, 0), (5000, 16, 1.25, 1), (777, 32, 1.0, 2), (4096, 64, 2.0, 3)])
def test_gate_losses_against_upstream_fastmoe_formulas(T, E, cf, seed):
    """What exactly differs from upstream FastMoE's capacity-limited gates (un-vendored; formulas restated from its
    published `fmoe/gates/switch_gate.py` and `fmoe/gates/gshard_gate.py`, v1.1.0), pinned as identities on the same routing:

    * SwitchGate, upstream: `fraction_expert = kept_e / sum(kept)`, `prob_expert = sum_t p[t, e] / sum(kept)`,
      `loss = E * sum(fraction_expert * prob_expert)`; here P_e divides by T:  upstream == ours * T / sum(kept).
    * SwitchGate capacity, upstream `ceil(cf * T)` PER EXPERT: with cf >= 1 it can never bind on one worker; ours is
      `ceil(cf * T / E)`.
    * GShardGate, upstream: `c_e` = histogram of the FIRST pick / T, `m_e = mean_t softmax(logits)[t, e]`,
      `loss = mean(c_e * m_e) * num_expert^2` = E * sum(c_e m_e) on one worker; ours counts both picks / (2 T):
      ours evaluated on the first-pick histogram == upstream.
    * GShardGate capacity, upstream `ceil(cf * T) * k // (W * E)`; ours `ceil(cf * T * k / E)`: equal up to rounding (<= k)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, E, generator=g) + torch.linspace(0.0, 1.5, E)   # skewed: real drops at cf 1.0 / 1.25
    p = torch.softmax(logits.double(), dim=-1)
    # ---- Switch (top-1, full-softmax score)
    cap = O.capacity_from_factor(cf, T, 1, E)
    r = O.route(logits, 1, 1, cap)
    ours = O.switch_aux_loss(r, r.psum, T)
    assert abs(float((O.aux_coef(r, T, O.AUX_SWITCH) * r.psum).sum()) - float(ours)) <= 1e-12
    kept = r.kept.double()
    valid = kept.sum()
    upstream = E * ((kept / valid) * (p.sum(0) / valid)).sum()
    assert abs(float(upstream) - float(ours) * T / float(valid)) <= 1e-6 * float(upstream)   # psum of the C oracle: fp32 exp
    assert int(np.ceil(cf * T)) >= int(r.count.max())            # upstream's per-expert capacity never binds (cf >= 1)
    assert cap == int(np.ceil(cf * T / E))
    # ---- GShard (top-2, scores = softmax over the two picks), before capacity
    cap2 = O.capacity_from_factor(cf, T, 2, E)
    r2 = O.route(logits, 2, 0, cap2)
    ours2 = float((O.aux_coef(r2, T, O.AUX_GSHARD) * r2.psum).sum())
    assert abs(ours2 - float(E * ((r2.count.double() / (2 * T)) * (r2.psum / T)).sum())) <= 1e-12
    top1 = torch.bincount(r2.idx[:, 0].long(), minlength=E).double()
    c_e, m_e = top1 / T, p.mean(0)
    upstream2 = float((c_e * m_e).mean() * E ** 2)
    ours_on_first_picks = float(E * ((top1 / T) * (r2.psum / T)).sum())
    assert abs(upstream2 - ours_on_first_picks) <= 1e-6 * upstream2
    assert abs(int(np.ceil(cf * T)) * 2 // E - cap2) <= 2
