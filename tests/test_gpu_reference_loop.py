"""GPU parity tests (`-m gpu`) at the reference's own boundary and under the reference's own training-loop conditions:

  * the drop-in `fmoe.FMoETransformerMLP`, instantiated with the constructor mapping of the reference's
    `CustomizedMoEMLP` (/root/reference/models/resMoE.py:15-29), against the fixtures that class produced on the CPU
    (tests/golden/ref_wrapper_*.npz, written by tests/golden/make_golden.py with the reference imported unmodified);
  * the layer inside `torch.cuda.amp.autocast()` (fp16) with a `GradScaler`-scaled backward, exactly what
    /root/reference/engine.py:52-54,68-74 does (timm's `NativeScaler` = scale -> backward -> unscale_ -> clip -> step);
  * the full per-GPU MoE-layer shapes of BASELINE configs[1..3], forward AND backward, against the oracle's arithmetic model;
  * a restated `train_one_epoch` body (engine.py:36-80) + `ModelEma` deepcopy (main.py:602-607) driving the CUDA layer.

Tolerances are written at each assertion: integers bit-exact; fp32 logits bit-exact; bf16-operand tensors vs the fp32
golden / fp64 ideal: relative Frobenius error <= 2e-2; vs the arithmetic model (same rounding points) <= 6e-3.
"""
import copy
import glob
import os
import sys

import numpy as np
import pytest
import torch

from _util import make_problem, rel_err
from oracle import moe_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
from make_golden_params import build_params, weights_digest  # noqa: E402

IDEAL_REL = 2e-2
MODEL_GRAD_REL = 6e-3
WRAPPER_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_wrapper_*.npz")))


def _act():
    return torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0))   # reference models/resMoE.py:25


@pytest.mark.parametrize("name", WRAPPER_NAMES)
def test_layer_vs_reference_wrapper_golden(name):
    """Same inputs and parameters as the reference's `CustomizedMoEMLP` run that wrote the fixture; the CUDA layer is
    built the way that class builds its parent: FMoETransformerMLP(E, d, hidden, Sequential(GELU, Dropout(0)), top_k=k)."""
    import fmoe
    from fmoe import _cabi as C
    from fmoe import functions as Fn
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, hid, E, k, B, N = [int(v) for v in z["meta"]]
    if "param.gate.gate.weight" in z.files:
        sd = {n[6:]: torch.from_numpy(z[n]) for n in z.files if n.startswith("param.")}
    else:
        sd = build_params(d, hid, E)
    assert weights_digest(sd) == str(z["weights_sha256"])
    layer = fmoe.FMoETransformerMLP(E, d, hid, _act(), top_k=k)
    layer.load_state_dict(sd)
    layer = layer.cuda().train()
    x = torch.from_numpy(z["x"]).cuda().requires_grad_()
    y = layer(x)
    assert y.shape == x.shape and y.dtype == x.dtype
    (y * torch.from_numpy(z["dy"]).cuda()).sum().backward()
    torch.cuda.synchronize()
    # routing of the same tokens through the C ABI: integers and fp32 logits bit-exact against the fixture
    T = B * N
    spec = Fn.RouteSpec(k, C.SCORE_TOPK_SOFTMAX, T * k, C.AUX_NONE)
    r = Fn.route(x.detach().reshape(T, d), layer.gate.gate.weight.detach(), layer.gate.gate.bias.detach(), spec)
    assert np.array_equal(r["logits"].cpu().numpy(), z["logits"])
    for f in ("idx", "pos", "count"):
        assert np.array_equal(r[f].cpu().numpy(), z[f]), f
    assert np.abs(r["score"].cpu().numpy() - z["score"]).max() <= 2e-6
    # floats: the fixture is fp32 arithmetic, the CUDA path runs bf16 operands with fp32 accumulation
    assert rel_err(y, torch.from_numpy(z["y"])) <= IDEAL_REL
    assert rel_err(x.grad, torch.from_numpy(z["dx"])) <= IDEAL_REL
    for pn, p in layer.named_parameters():
        if "grad." + pn in z.files:
            assert rel_err(p.grad, torch.from_numpy(z["grad." + pn])) <= IDEAL_REL, pn
        else:
            got = torch.stack([p.grad[e].norm() for e in range(E)]).double()
            assert rel_err(got, torch.from_numpy(z["gradnorm." + pn])) <= IDEAL_REL, pn


@pytest.mark.parametrize("gate_kind,k", [("naive", 2), ("switch", 1)])
def test_fp16_autocast_and_gradscaler(gate_kind, k):
    """engine.py:52 runs the model under fp16 autocast; engine.py:68-74 hands the loss to timm's NativeScaler:
    `scaler.scale(loss).backward(); scaler.unscale_(opt); clip_grad_norm_; scaler.step(opt); scaler.update()`.
    The layer receives the fp32 output of LayerNorm (autocast keeps layer_norm in fp32) and, in backward, a dy that
    carries the 2^16 loss scale.  Unscaled gradients must match the fp64 ideal at the usual bf16 tolerance, and the
    scaler must not see an inf (no step skipped)."""
    import fmoe
    T, d, h, E = 788, 192, 768, 8
    torch.manual_seed(3)
    layer = fmoe.build_moe_mlp(d, h, num_experts=E, top_k=k, gate=gate_kind).cuda().train()
    norm = torch.nn.LayerNorm(d).cuda()
    head = torch.nn.Linear(d, 10).cuda()
    params = list(layer.parameters()) + list(norm.parameters()) + list(head.parameters())
    opt = torch.optim.SGD(params, lr=0.0)            # lr 0: the step is exercised, the parameters stay comparable
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 16)
    x0 = torch.randn(4, 197, d, device="cuda", requires_grad=True)
    target = torch.randint(0, 10, (4,), device="cuda")
    with torch.autocast("cuda", dtype=torch.float16):
        n = norm(x0)
        assert n.dtype == torch.float32               # what the layer sees under the reference's autocast
        y = layer(n)
        assert y.dtype == n.dtype and y.shape == n.shape
        out = head((x0 + y).mean(1))
        loss = torch.nn.functional.cross_entropy(out.float(), target)
    scale_before = scaler.get_scale()
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    gnorm = torch.nn.utils.clip_grad_norm_(params, 1e9)
    assert torch.isfinite(gnorm)
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() == scale_before, "GradScaler backed off: the scaled backward produced an inf/nan"
    # fp64 ideal of the same graph on the CPU (same routing: taken from the oracle on the same fp32 LayerNorm output)
    sd = {kk: v.detach().cpu() for kk, v in layer.state_dict().items()}
    names = ["gate.gate.weight", "gate.gate.bias", "experts.htoh4.weight", "experts.htoh4.bias", "experts.h4toh.weight", "experts.h4toh.bias"]
    n_cpu = n.detach().cpu().reshape(T, d)
    logits = O.gate_logits(n_cpu, sd[names[0]], sd[names[1]])
    mode = O.SCORE_TOPK_SOFTMAX if gate_kind == "naive" else 1
    cap = T * k if gate_kind == "naive" else O.capacity_from_factor(1.25, T, k, E)
    r = O.route(logits, k, mode, cap)
    xs64 = x0.detach().cpu().double().requires_grad_()
    nw, nb = norm.weight.detach().cpu().double(), norm.bias.detach().cpu().double()
    ps = [sd[nm].double().requires_grad_() for nm in names]
    n64 = torch.nn.functional.layer_norm(xs64, (d,), nw, nb, norm.eps)
    y64, _ = O.ideal_forward(n64.reshape(T, d), *ps, r, mode)
    out64 = torch.nn.functional.linear((xs64 + y64.reshape(4, 197, d)).mean(1), head.weight.detach().cpu().double(), head.bias.detach().cpu().double())
    torch.nn.functional.cross_entropy(out64, target.cpu()).backward()
    assert rel_err(y.reshape(T, d), y64) <= IDEAL_REL
    assert rel_err(x0.grad / scale_before, xs64.grad) <= IDEAL_REL      # x0 is not an optimizer parameter: still scaled
    for nm, p64 in zip(names, ps):
        got = dict(layer.named_parameters())[nm].grad
        if p64.grad is None or float(p64.grad.abs().max()) == 0.0:     # top-1 NaiveGate-style scores carry no gate gradient
            assert got is None or float(got.abs().max()) <= 1e-6, nm
        else:
            assert rel_err(got, p64.grad) <= IDEAL_REL, nm


def test_fp16_input_is_widened_and_returned_in_fp16():
    """A caller that hands the layer fp16 activations directly (no LayerNorm in front) gets fp16 back; inside, the
    row is widened to fp32 (never narrowed through fp16's range)."""
    import fmoe
    torch.manual_seed(0)
    layer = fmoe.FMoETransformerMLP(8, 192, 768, _act(), top_k=2).cuda()
    x = torch.randn(2, 197, 192, device="cuda")
    y32 = layer(x.half().float() * 64.0)          # same values; x64 keeps the outputs clear of fp16's subnormal range
    y16 = layer(x.half() * 64.0)
    assert y16.dtype == torch.float16 and y16.shape == x.shape
    assert rel_err(y16, y32) <= 2e-3                # fp16 rounding of the returned tensor only (2^-11)


FULL_SIZE = {
    # name: (T, d, E, k, score mode, aux mode name)  — the per-GPU MoE-layer shapes of BASELINE configs[1..3]
    "c2": (256 * 197, 384, 16, 1, 1, "switch"),      # ViT-S, Switch top-1
    "c3": (128 * 197, 768, 32, 2, 0, "gshard"),      # ViT-B, GShard top-2
    "c4": (128 * 197, 1024, 64, 1, 1, "switch"),     # ViT-L, Switch top-1, E = 64
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE))
def test_full_size_forward_backward_vs_model(name):
    """The MoE-layer shapes of BASELINE configs[1..3] at FULL size (bf16 activations, cf 1.25, skewed gate bias -> real
    drops): every output and every gradient of the CUDA layer against the oracle's arithmetic model (fp32 BLAS on the
    host: seconds), routing counts bit-exact."""
    from fmoe import _cabi as C
    from fmoe import functions as Fn
    T, d, E, k, mode, aux_name = FULL_SIZE[name]
    h = 4 * d
    aux_mode = C.AUX_SWITCH if aux_name == "switch" else C.AUX_GSHARD
    x, Wg, bg, W1, b1, W2, b2 = make_problem(T, d, h, E, seed=9, skew=0.5, x_dtype=torch.bfloat16)
    cap = O.capacity_from_factor(1.25, T, k, E)
    spec = Fn.RouteSpec(k, mode, cap, aux_mode)
    dev = [t.cuda().requires_grad_() for t in (x, Wg, bg, W1, b1, W2, b2)]
    dy = torch.randn(T, d, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    y, aux, count, kept = Fn.MoEFunction.apply(*dev, spec, Fn.Bf16WeightCache(), None)
    torch.autograd.backward([y, aux], [dy.cuda(), torch.tensor(0.01, device="cuda")])
    torch.cuda.synchronize()
    ym, sv = O.forward_model(x, Wg, bg, W1, b1, W2, b2, k, mode, cap)
    assert torch.equal(count.cpu(), sv.r.count) and torch.equal(kept.cpu(), sv.r.kept)
    assert int(sv.r.count.sum() - sv.r.kept.sum()) > 0
    dpsum = O.aux_coef(sv.r, T, aux_mode) * 0.01
    gm = O.backward_model(sv, dy, Wg, dpsum=dpsum)
    assert rel_err(y, ym.float()) <= 3e-3
    got = dict(zip(("dx", "dWg", "dbg", "dW1", "db1", "dW2", "db2"), (t.grad for t in dev)))
    for nme in ("dx", "dW1", "db1", "dW2", "db2", "dWg", "dbg"):
        assert rel_err(got[nme], gm[nme]) <= MODEL_GRAD_REL, (nme, rel_err(got[nme], gm[nme]))


def test_train_loop_body_on_the_cuda_layer():
    """Four iterations of the reference's `train_one_epoch` body (/root/reference/engine.py:36-80: fp16 autocast,
    criterion, `loss_scaler(loss, optimizer, clip_grad=..., parameters=model.parameters())`, `model_ema.update`) with
    the model deep-copied for the EMA first (main.py:602-607), on a small ViT whose MoE blocks are the CUDA layer.
    The reference's own `models/` and timm are absent on the GPU box, so the host model is this repo's `MoEViT`; the
    loop body is restated line for line.  Loss must be finite, go down on a repeated batch, and the EMA must move."""
    import fmoe
    from moe_vit import MoEViT, MoEViTConfig
    torch.manual_seed(0)
    cfg = MoEViTConfig(size="tiny", num_experts=8, top_k=1, gate="switch", capacity_factor=1.25, moe_stride=2, num_classes=10)
    model = MoEViT(cfg, fused_norm=False).cuda()      # stock LayerNorm / Linear / attention blocks, as in the reference; MoE blocks = the CUDA layer
    ema = copy.deepcopy(model).eval()                      # ModelEma(model, decay, device='', resume='')
    for p in ema.parameters():
        p.requires_grad_(False)
    decay = 0.9
    criterion = fmoe.MoEAuxCriterion(torch.nn.CrossEntropyLoss(), model, coef=0.01)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    scaler = torch.amp.GradScaler("cuda")                 # timm NativeScaler wraps exactly this
    samples = torch.randn(8, 3, 224, 224, device="cuda")
    targets = torch.randint(0, 10, (8,), device="cuda")
    model.train()
    losses = []
    for _ in range(4):
        with torch.autocast("cuda", dtype=torch.float16):  # engine.py:52
            outputs = model(samples)
            loss = criterion(outputs.float(), targets)      # engine.py:53-54 (MoEAuxCriterion adds the gates' losses)
        loss_value = loss.item()                            # engine.py:56
        assert np.isfinite(loss_value)                      # engine.py:58-60
        losses.append(loss_value)
        opt.zero_grad()
        scaler.scale(loss).backward()                       # NativeScaler.__call__ (engine.py:68-74)
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        torch.cuda.synchronize()                            # engine.py:76
        with torch.no_grad():                               # model_ema.update(model), engine.py:77-78
            for pe, pm in zip(ema.state_dict().values(), model.state_dict().values()):
                if pe.dtype.is_floating_point:
                    pe.mul_(decay).add_(pm.detach(), alpha=1 - decay)
    assert min(losses[1:]) < losses[0], losses          # the repeated batch is being fitted (single steps may overshoot)
    stats = fmoe.load_balance_stats(model)
    assert len(stats) == 6 and all(s["routed_pairs"] == 8 * 197 for s in stats.values())
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):   # evaluate(), engine.py:88-121, on the EMA copy
        out = ema(samples)
    assert torch.isfinite(out).all()


def test_expert_weight_copies_follow_invisible_weight_updates():
    """The bf16 operand copies of the expert weights must follow updates that do not bump the tensor version: many
    optimizers reachable through the reference's `--opt` write through `p.data`.  (a) train forward,
    `.data` update, evaluation forward: the first no_grad forward after training re-casts; (b) two training forwards
    around a `.data` update; (c) between two evaluation forwards `.data` writes need `invalidate_weight_cache()`."""
    import fmoe
    torch.manual_seed(0)
    d, h, E = 128, 256, 4
    layer = fmoe.FMoETransformerMLP(E, d, h, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(0.0)), top_k=2).cuda()
    x = torch.randn(600, d, device="cuda")

    def fresh_eval():   # what a layer without any cached state returns for the current weights
        twin = copy.deepcopy(layer)
        with torch.no_grad():
            return twin(x)

    y0 = layer(x)                                   # training forward (autograd records, weights require grad)
    y0.square().mean().backward()
    with torch.no_grad():
        assert rel_err(layer(x), y0.detach()) <= 1e-6
    layer.experts.htoh4.weight.data.mul_(1.5)       # invisible to `_version`
    layer.experts.h4toh.weight.data.add_(0.01)
    # (b) the next training forward sees the new weights
    y1 = layer(x)
    want = fresh_eval()
    assert rel_err(y1.detach(), want) <= 1e-6 and rel_err(y1.detach(), y0.detach()) > 1e-2
    # (a) train forward, invisible update, evaluation forward
    layer.experts.htoh4.weight.data.mul_(0.5)
    with torch.no_grad():
        y2 = layer(x)
        assert rel_err(y2, fresh_eval()) <= 1e-6 and rel_err(y2, y1.detach()) > 1e-2
        # (c) evaluation forwards reuse their copies until told otherwise
        layer.experts.h4toh.weight.data.mul_(2.0)
        layer.invalidate_weight_cache()
        assert rel_err(layer(x), fresh_eval()) <= 1e-6
