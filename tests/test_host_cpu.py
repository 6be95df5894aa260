"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares,
and the `fmoe` drop-in mirrors the interface the reference relies on.  No compute
calls are made here — there is no GPU and, by design, no CPU fallback."""
import copy
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "moe_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(moe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fmoe import _cabi as C
    names = header_functions()
    assert len(names) >= 17
    assert set(names) == set(C.SIGNATURES), "ctypes table and header must list the same entry points"
    out = subprocess.run(["nm", "-D", "--defined-only", C.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (moe_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert C.lib.moe_version() >= 100
    assert C.ROW_ALIGN == 256
    assert C.lib.moe_rows_cap(1000, 2, 8, 2000) == 2048 + 256 * 8
    assert C.lib.moe_rows_cap(1000, 1, 8, 100) == 1024 + 2048    # min(T*k, E*C) rounded up to 256, + 256 per expert


def test_op_codes_and_workspace_sizes_match_the_header():
    """The ctypes side mirrors the header's op codes; the size functions are pure host arithmetic (no GPU needed)."""
    from fmoe import _cabi as C
    src = open(HEADER).read()
    for name in ("FC1", "FC2", "DGELU", "DGRAD", "WGRAD", "WGRAD_T"):
        m = re.search(rf"#define MOE_GEMM_{name} (\d+)", src)
        assert m and int(m.group(1)) == getattr(C, f"GEMM_{name}"), name
    # split-K flags: one int per (tile, CTA of the pair, epilogue warp), tiles counted for 128-wide tiles
    assert C.lib.moe_wgrad_flags_bytes(16, 1536, 384) == 16 * 6 * 3 * 2 * 16 * 4
    # slab column sums of the dgelu epilogue: one fp32 row per 32 packed rows
    assert C.lib.moe_slab_colsum_bytes(53504, 1536) == 53504 // 32 * 1536 * 4
    assert C.lib.moe_segment_colsum_workspace_bytes(53504, 1536) == 53504 // 128 * 1536 * 4


def test_library_is_sm100a_tcgen05_code():
    """The shipped kernels are Blackwell-native: tcgen05.mma (UTCHMMA), TMA (UTMALDG/UTMASTG), TMEM loads (LDTM)."""
    from fmoe import _cabi as C
    sass = subprocess.run(["cuobjdump", "-sass", C.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM"):
        assert mnemonic in sass.stdout, mnemonic


def test_argument_validation_without_gpu():
    from fmoe import _cabi as C
    with pytest.raises(C.MoeB200Error, match="multiple of 8"):
        C.call("moe_cast_bf16", None, None, 12, None)
    with pytest.raises(C.MoeB200Error, match="unsupported shape"):
        C.call("moe_gate_fwd", None, 0, None, None, None, None, 10, 100, 4, 1, 0, 0, None, None, None, None, None, None, None)
    with pytest.raises(C.MoeB200Error, match="multiple of 64"):
        C.call("moe_grouped_gemm", C.GEMM_FC2, None, None, None, None, None, None, None, None, None, 256, 2, 0, 100, 64, None)


def test_gate_kernels_are_tensor_core_code():
    """The gate forward runs on tcgen05 in its own single-CTA flavour (UTCHMMA without .2CTA next to the GEMM's .2CTA), the
    gate / dispatch backward on mma.sync with ldmatrix.trans / stmatrix (HMMA, LDSM, STSM)."""
    from fmoe import _cabi as C
    sass = subprocess.run(["cuobjdump", "-sass", C.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    funcs = sass.stdout.split("Function : ")
    gate = [f for f in funcs if f.startswith("_ZN3moe20gate_fwd_umma_kernel")]
    gdb = [f for f in funcs if "gate_dispatch_bwd_mma_kernel" in f.split("\n", 1)[0]]
    assert len(gate) >= 9 and len(gdb) >= 12
    assert all("UTCHMMA" in f and "UTMALDG" in f and "LDTM" in f for f in gate)
    assert all("HMMA.1688.F32.TF32" in f and "HMMA.16816.F32.BF16" in f and "LDSM" in f for f in gdb)
    assert any("STSM" in f for f in gdb)


def test_layer_norm_backward_is_bulk_copy_staged():
    """The fused add+LayerNorm backward that ships stages its rows with non-tensor bulk copies on mbarriers
    (cp.async.bulk -> UBLKCP, mbarrier try_wait -> SYNCS) and spills nothing; its workspace is one partial per CTA of
    the larger of the two grids (one 8-warp CTA per SM for d > 256, two below)."""
    from fmoe import _cabi as C
    sass = subprocess.run(["cuobjdump", "-sass", C.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    funcs = [f for f in sass.stdout.split("Function : ") if f.startswith("_ZN3moe21addln_bwd_bulk_kernel")]
    assert len(funcs) == 16                                    # {f32, bf16} dn x {f32, bf16} d_delta x four row widths
    assert all("UBLKCP" in f and "SYNCS.PHASECHK" in f and "LDS.128" in f for f in funcs)
    assert not any("STL" in f or "LDL" in f for f in funcs)    # no local-memory spills
    ws = C.lib.moe_addln_bwd_workspace_bytes(50432, 384)
    assert ws % (2 * 384 * 4) == 0 and ws // (2 * 384 * 4) >= 1


def test_peer_entry_points_validate_their_arguments_without_gpu():
    """Expert parallelism over peer memory: group size, rank and pointer-table checks happen before anything touches a GPU."""
    import ctypes
    from fmoe import _cabi as C
    arr = (ctypes.c_void_p * 2)(None, None)
    with pytest.raises(C.MoeB200Error, match="ranks"):
        C.call("moe_ep_barrier", arr, None, 0, 9, None, None)            # more ranks than one NVLink node holds
    with pytest.raises(C.MoeB200Error, match="rank"):
        C.call("moe_ep_barrier", arr, None, 2, 2, None, None)            # rank outside the group
    with pytest.raises(C.MoeB200Error, match="NULL"):
        C.call("moe_ep_barrier", arr, None, 0, 2, None, None)            # unmapped peer
    with pytest.raises(C.MoeB200Error, match="bad arguments"):
        C.call("moe_ep_heap_alloc", 0, None, None)


def test_expert_parallel_transport_selection():
    from fmoe import distributed as D
    from fmoe.functions import RouteSpec
    lyr = _layer(top_k=2)
    lyr.world_size, lyr.moe_group = 2, None
    spec = RouteSpec(2, 0, 100, 0)
    x = torch.zeros(4, 192)
    assert D.TRANSPORT == "auto" and D._pick_transport(lyr, x, spec) == "nccl"     # CPU tensors (gloo tests): never the peer path
    D.TRANSPORT = "peer"
    try:
        assert D._pick_transport(lyr, x, spec) == "peer"                            # an explicit choice is honoured (and fails loudly later)
    finally:
        D.TRANSPORT = "auto"


def test_peer_buffer_row_bound():
    """Static size of a rank's packed buffers under peer-memory expert parallelism: W * E_local * capacity rows for a
    capacity-limited gate, W * T * k (every pair of every rank) for NaiveGate, + one 256-row tile of padding per local expert."""
    from fmoe.peer import peer_rows_per_rank
    W, El, T, k = 8, 2, 50432, 1
    cap = 3940                                                    # ceil(1.25 * T / 16)
    assert peer_rows_per_rank(W, El, T, k, cap) == (W * El * cap + 255) // 256 * 256 + 256 * El
    assert peer_rows_per_rank(W, El, T, k, T * k) == (W * T * k + 255) // 256 * 256 + 256 * El      # no capacity: bounded by the pairs
    assert peer_rows_per_rank(2, 4, 900, 2, 900 * 2) == (2 * 900 * 2 + 255) // 256 * 256 + 256 * 4


def _layer(**kw):
    import fmoe
    act = torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0))
    return fmoe.FMoETransformerMLP(8, 192, 768, act, **kw)


def test_module_contract():
    import fmoe
    layer = _layer(top_k=2)
    sd = layer.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "gate.gate.weight": (8, 192), "gate.gate.bias": (8,),
        "experts.htoh4.weight": (8, 768, 192), "experts.htoh4.bias": (8, 768),
        "experts.h4toh.weight": (8, 192, 768), "experts.h4toh.bias": (8, 192)}
    assert sum(p.numel() for p in layer.parameters()) == 2_368_520
    assert isinstance(layer.gate.gate, torch.nn.Linear)                   # resmoe_flop_hook.py:7
    assert (layer.gate.gate.in_features, layer.gate.gate.out_features) == (192, 8)
    assert (layer.top_k, layer.num_expert, layer.d_model, layer.world_size) == (2, 8, 192, 1)
    assert all("moe_gate" not in n and "dense_gate" not in n for n, _ in layer.named_parameters())  # main.py:619-631
    assert float(layer.experts.htoh4.bias.abs().sum()) == 0.0            # upstream init: zero expert bias
    clone = copy.deepcopy(layer)                                          # ModelEma, main.py:602-607
    clone.load_state_dict(sd)
    assert not layer.gate.has_loss and layer.gate.get_loss() is None
    layer.train(); layer.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.randn(2, 5, 192))
    with pytest.raises(ValueError):
        fmoe.FMoE.forward(layer, torch.randn(10, 100))     # wrong feature size


def test_unsupported_configs_raise_at_construction():
    import fmoe
    with pytest.raises(NotImplementedError):
        fmoe.FMoETransformerMLP(4, 64, 256, torch.nn.ReLU())
    with pytest.raises(NotImplementedError):
        fmoe.FMoETransformerMLP(4, 64, 256, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.1)))
    with pytest.raises(NotImplementedError):
        fmoe.FMoETransformerMLP(4, 64, 256, torch.nn.GELU(approximate="tanh"))
    with pytest.raises(ValueError):
        fmoe.FMoETransformerMLP(4, 100, 256, torch.nn.GELU())
    with pytest.raises(NotImplementedError):
        fmoe.FMoETransformerMLP(4, 64, 256, torch.nn.GELU(), gate_hook=lambda *a: None)
    with pytest.raises(AssertionError):   # same assertion as upstream SwitchGate
        fmoe.FMoETransformerMLP(4, 64, 256, torch.nn.GELU(), top_k=2, gate=fmoe.SwitchGate)
    sw = fmoe.FMoETransformerMLP(16, 384, 1536, torch.nn.GELU(), top_k=1, gate=fmoe.SwitchGate)
    spec = sw.gate.route_spec(50432)
    assert (spec.top_k, spec.score_mode, spec.want_psum) == (1, 1, True)
    assert spec.capacity == 3783      # ceil(1.2 * 50432 / 16) in training mode
    sw.eval()
    assert sw.gate.route_spec(50432).capacity == 7565


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not mounted (GPU box)")
def test_unmodified_reference_models_build_on_the_drop_in():
    """`import models` from /root/reference (through tiny timm stubs) must pick up OUR fmoe."""
    code = r"""
import sys
sys.path[:0] = [%r, %r, '/root/reference']
import torch, fmoe, models
from timm.models import create_model
kw = dict(num_classes=10, drop_rate=0., drop_path_rate=0., drop_block_rate=None, img_size=224)
m = create_model('moe_tiny_patch16_224_expert8', **kw)
assert sum(p.numel() for p in m.parameters()) == 30398122
blocks = [b for b in m.modules() if type(b).__name__ == 'Block']
assert len(blocks) == 12 and all(isinstance(b.mlp, fmoe.FMoETransformerMLP) for b in blocks)
assert blocks[0].mlp.gate.gate.out_features == 8 and blocks[0].mlp.top_k == 2
m2 = create_model('resmoe_tiny_patch16_224_expert8', starting_threshold=1., target_threshold=.9, **kw)
assert sum(p.numel() for p in m2.parameters()) == 30402754
import copy; copy.deepcopy(m)
print('OK')
""" % (os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "slim-switch-moe-vit_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not mounted (GPU box)")
def test_reference_flop_hook_reads_the_drop_in_layer():
    """The reference's own flop counter for the MoE layer (/root/reference/models/resmoe_flop_hook.py:4-10) walks
    `mlp.gate.gate.in_features / .out_features`: run the unmodified function on the drop-in layer of each reference model."""
    code = r"""
import sys
sys.path[:0] = [%r, %r, '/root/reference']
import torch, fmoe, models
from models.resmoe_flop_hook import moe_flops
from timm.models import create_model
kw = dict(num_classes=10, drop_rate=0., drop_path_rate=0., drop_block_rate=None, img_size=224)
for name, extra in (('moe_tiny_patch16_224_expert8', {}), ('resmoe_tiny_patch16_224_expert8', dict(starting_threshold=1., target_threshold=.9))):
    m = create_model(name, **kw, **extra)
    mlp = [b for b in m.modules() if type(b).__name__ == 'Block'][0].mlp
    shape = (8, 197, 192)
    assert int(moe_flops(mlp, shape)) == 8 * 197 * 192 * 8 + 8 * 197 * (3 * 192 - 1)
print('OK')
""" % (os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "slim-switch-moe-vit_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "slim-switch-moe-vit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "moe_oracle" not in text, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not mounted (GPU box)")
def test_north_star_factories_register_and_build_on_the_reference_backbones():
    """switch_moe_models.py: the four BASELINE configs as timm-registered factories over the reference's own dense
    backbones (pattern of /root/reference/models/resMoE.py:190-209), selected exactly as main.py:520-530 does.
    Tiny / Small are built for real; the MoE layers of Base / Large on the meta device (their expert weights are GBs)."""
    code = r"""
import sys
sys.path[:0] = [%r, %r, '/root/reference']
import torch, fmoe, models, switch_moe_models as Z
from timm.models import create_model
kw = dict(num_classes=10, drop_rate=0., drop_path_rate=0., drop_block_rate=None, img_size=224, starting_threshold=1., target_threshold=.9)
want = {'switch_moe_tiny_patch16_224_e8_top1': (192, 12, 6, 8, 1, fmoe.SwitchGate),
        'switch_moe_small_patch16_224_e16_top1': (384, 12, 6, 16, 1, fmoe.SwitchGate),
        'gshard_moe_base_patch16_224_e32_top2': (768, 12, 6, 32, 2, fmoe.GShardGate),
        'switch_moe_large_patch16_224_e64_top1': (1024, 24, 12, 64, 1, fmoe.SwitchGate)}
import fmoe.integration as I
_real = I.build_moe_mlp
def _on_meta(dim, hidden, **k):      # Base / Large expert weights are GBs: build those layers on the meta device
    with torch.device('meta' if dim >= 768 else 'cpu'):
        return _real(dim, hidden, **k)
I.build_moe_mlp = _on_meta
for name, (d, depth, n_moe, E, k, gate_cls) in want.items():
    big = d >= 768
    m = create_model(name, **kw)
    blocks = [b for b in m.modules() if type(b).__name__ == 'Block']
    moe = [b.mlp for b in blocks if isinstance(b.mlp, fmoe.FMoETransformerMLP)]
    assert len(blocks) == depth and len(moe) == n_moe, (name, len(blocks), len(moe))
    assert [isinstance(b.mlp, fmoe.FMoETransformerMLP) for b in blocks[:4]] == [False, True, False, True]   # every other block
    l = moe[0]
    assert (l.d_model, l.d_hidden, l.num_expert, l.top_k, l.world_size) == (d, 4 * d, E, k, 1)
    assert isinstance(l.gate, gate_cls) and l.gate.capacity == (1.25, 1.25) and l.gate.gate.out_features == E
    assert not getattr(m, '_ddp_params_and_buffers_to_ignore', [])
    # the optimizer grouping of main.py:619-631 keeps every MoE parameter in the base group
    assert all('moe_gate' not in n and 'dense_gate' not in n for n, _ in m.named_parameters())
    if not big:
        import copy; copy.deepcopy(m)               # ModelEma (main.py:602-607)
        crit = fmoe.MoEAuxCriterion(lambda s, o, t: o.sum() * 0, m, 0.01)
        assert len(crit._layers) == n_moe
print('OK')
""" % (os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "slim-switch-moe-vit_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-3000:]


def test_aux_criterion_and_load_balance_logging_host_logic():
    """MoEAuxCriterion adds coef * sum(gate losses) and clears them; log_load_balance feeds a MetricLogger-like object
    (reference utils.py:118-211) and a TensorboardXTracker-like writer (utils.py:299-319).  Pure host logic: the gate
    losses / counts are planted by hand (no kernels run on CPU)."""
    import fmoe
    act = torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0))
    model = torch.nn.ModuleList([fmoe.FMoETransformerMLP(4, 64, 256, act, top_k=1, gate=fmoe.make_gate("switch", 1.25)) for _ in range(2)])
    for i, l in enumerate(model):
        l.gate.set_loss(torch.tensor(1.5 + i, requires_grad=True))
        l.last_count = torch.tensor([10, 30, 0, 40], dtype=torch.int32)
        l.last_kept = torch.tensor([10, 25, 0, 25], dtype=torch.int32)
    crit = fmoe.MoEAuxCriterion(lambda s, o, t: o.sum(), model, coef=0.1)
    out = torch.ones(3, requires_grad=True)
    loss = crit(None, out, None)
    assert abs(float(loss) - (3.0 + 0.1 * (1.5 + 2.5))) < 1e-6 and abs(float(crit.last_aux) - 4.0) < 1e-6
    assert not any(l.gate.has_loss for l in model), "gate losses are consumed"
    assert float(crit(None, out, None)) == 3.0          # nothing pending: plain criterion

    class Logger:
        def __init__(self): self.kw = {}
        def update(self, **kw): self.kw.update(kw)

    class Writer:
        def __init__(self): self.rows = []
        def log_scalar(self, name, value, step): self.rows.append((name, value, step))

    lg, wr = Logger(), Writer()
    scalars = fmoe.log_load_balance(model, lg, wr, step=7)
    assert abs(scalars["moe_drop_rate"] - 0.25) < 1e-12 and abs(scalars["moe_max_load"] - 2.0) < 1e-12
    assert lg.kw == scalars and len(wr.rows) == 4 and wr.rows[0] == ("moe/0/drop_rate", 0.25, 7)
    st = fmoe.load_balance_stats(model)["1"]
    assert st["routed_pairs"] == 80 and st["kept_pairs"] == 60 and st["count"] == [10, 30, 0, 40]
