#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the B200 Switch-MoE layer and the MoE-ViT built on it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A *step* is one full training step (forward, loss + Switch aux
loss, backward, fused AdamW) of BASELINE.json configs[1] — ViT-Small/16 Switch-MoE, 16 experts, top-1,
capacity 1.25, MoE every other block, bf16 autocast, synthetic 224x224 batch 256 PER GPU — so every
step makes 6 forward+backward passes through the hot path at T = 50 432 tokens.  With N > 1 the
experts are sharded N ways (expert parallel, all-to-all dispatch/combine) and the dense blocks are
data parallel: weak scaling.

  value   images/s, inputs resident in HBM (CUDA events, max over ranks)
  e2e     images/s through the public module API with HOST inputs: per step a pinned-host -> device
          copy of the batch and a device -> host read of the loss are inside the timed region (the loss is read
          one step behind through a pinned buffer; `e2e.sync_readback_value` = the same loop with loss.item()
          after every step)
  roofline   the grouped tcgen05 expert GEMM (the dominant kernel of the hot path): algorithmic flops
             12*R*d*h per layer fwd+bwd over the CUDA-event time of its launches inside the timed region
  moe_layer  the isolated layer at the same shape: tokens/s fwd+bwd and per-kernel times / HBM fractions
  cpu_baseline / --impl reference   the CPU restatement of the reference's layer (oracle/fmoe_cpu.py —
             FastMoE has no CPU kernels and is not installable here) inside the same host model, fp32,
             on the box's host cores, on a bounded sample (batch 8) of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "slim-switch-moe-vit_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# stdout carries exactly ONE line (the JSON): everything libraries print (NCCL's version banner, warnings) goes to stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

UNIT = "images/s"
CPU_SAMPLE_BATCH = 8
AUX_COEF = 0.01
# BASELINE.json configs[1..3] (SURVEY.md §8 cfg table).  c2 is the default: the configuration the metric is quoted on
# that fits one GPU; c3 / c4 are the expert-parallel configurations (run them under torchrun with --gpus 2/4/8).
CONFIGS = {
    "c1": dict(size="tiny", num_experts=8, top_k=1, gate="switch", batch=8,
               metric="MoE-ViT training throughput (ViT-Ti/16 Switch-MoE E8 top-1 cf1.25, fp32, 224px, batch 8, CPU)"),
    "c2": dict(size="small", num_experts=16, top_k=1, gate="switch", batch=256,
               metric="MoE-ViT training throughput (ViT-S/16 Switch-MoE E16 top-1 cf1.25, bf16, 224px, batch 256/GPU)"),
    "c3": dict(size="base", num_experts=32, top_k=2, gate="gshard", batch=128,
               metric="MoE-ViT training throughput (ViT-B/16 GShard-MoE E32 top-2 cf1.25, bf16, 224px, batch 128/GPU, expert parallel)"),
    "c4": dict(size="large", num_experts=64, top_k=1, gate="switch", batch=128,
               metric="MoE-ViT training throughput (ViT-L/16 Switch-MoE E64 top-1 cf1.25, bf16, 224px, batch 128/GPU = 1024 on 8 GPUs, expert parallel)"),
}
BACKLOG_CYCLES_STEP = int(60e-3 * 1.9e9)    # torch.cuda._sleep spin before each profiled training step (~60 ms)
BACKLOG_CYCLES_LAYER = int(4e-3 * 1.9e9)    # ... before each profiled isolated-layer iteration (~4 ms)
CONFIG = "c2"       # set by --config
METRIC = CONFIGS["c2"]["metric"]
PER_GPU_BATCH = CONFIGS["c2"]["batch"]


def select_config(name: str) -> None:
    global CONFIG, METRIC, PER_GPU_BATCH
    CONFIG, METRIC, PER_GPU_BATCH = name, CONFIGS[name]["metric"], CONFIGS[name]["batch"]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = self.proc.communicate()[0]
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the workload
# ------------------------------------------------------------------------------------------------
def make_cfg(world_size: int, name: str | None = None):
    from moe_vit import MoEViTConfig
    c = CONFIGS[name or CONFIG]
    return MoEViTConfig(size=c["size"], num_experts=c["num_experts"], top_k=c["top_k"], capacity_factor=1.25, moe_stride=2,
                        gate=c["gate"], num_classes=1000, world_size=world_size)


def synthetic_batch(batch: int, seed: int, pin: bool):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, 224, 224, generator=g)
    lab = torch.randint(0, 1000, (batch,), generator=g)
    if pin:
        img, lab = img.pin_memory(), lab.pin_memory()
    return img, lab


def train_step(model, opt, img, lab, autocast_dtype, zero_grad=True):
    if autocast_dtype is not None:
        with torch.autocast("cuda", dtype=autocast_dtype):
            logits = model(img)
            loss = F.cross_entropy(logits.float(), lab)
    else:
        logits = model(img)
        loss = F.cross_entropy(logits, lab)
    inner = model.module if hasattr(model, "module") else model
    aux = inner.aux_loss()
    total = loss if aux is None else loss + AUX_COEF * aux
    total.backward()
    opt.step()
    if zero_grad:   # (a captured step keeps its static .grad tensors: backward overwrites them on every replay)
        opt.zero_grad(set_to_none=True)
    return loss


# ------------------------------------------------------------------------------------------------
# CPU arm (reported baseline): same host model around the CPU restatement of the reference's layer
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float | None, name: str | None = None):
    """Times `steps` training steps (after `warmup`) of the config on the host cores at batch
    CPU_SAMPLE_BATCH, fp32.  With `budget_s` the step count is cut so the leg stays inside the budget.
    All host cores are used (torchrun exports OMP_NUM_THREADS=1, which would leave this arm on one thread)."""
    from moe_vit import MoEViT
    from oracle import fmoe_cpu, moe_oracle as O

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = make_cfg(1, name)
    act = torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0))
    score_mode = O.SCORE_FULL_SOFTMAX if cfg.gate == "switch" else O.SCORE_TOPK_SOFTMAX

    def moe_mlp(dim, hidden):
        return fmoe_cpu.FMoETransformerMLP(cfg.num_experts, dim, hidden, act, top_k=cfg.top_k,
                                           score_mode=score_mode, capacity_factor=cfg.capacity_factor)

    torch.manual_seed(0)
    model = MoEViT(cfg, moe_mlp=moe_mlp)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    img, lab = synthetic_batch(CPU_SAMPLE_BATCH, 0, pin=False)
    t0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        train_step(model, opt, img, lab, None)
    per = (time.perf_counter() - t0) / max(1, warmup)
    if budget_s is not None:
        steps = max(2, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(model, opt, img, lab, None)
    dt = time.perf_counter() - t0
    T = CPU_SAMPLE_BATCH * 197
    n_moe = len(model.moe_layers)
    return dict(images_per_s=CPU_SAMPLE_BATCH * steps / dt, ms_per_step=1e3 * dt / steps, steps=steps,
                cores=torch.get_num_threads(),
                sample=f"{steps} training steps at batch {CPU_SAMPLE_BATCH} ({T} tokens x {n_moe} MoE layers per step), fp32, "
                       f"{cfg.describe()}; oracle/fmoe_cpu.py (CPU restatement of FastMoE; FastMoE has no CPU kernels)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, budget_s=120.0)
    # BASELINE.md §3's CPU-runnable case (configs[0]: ViT-Ti/16, E = 8, top-1, MoE every other block, batch 8), reported beside it
    r1 = cpu_reference_run(10, 3, budget_s=30.0, name="c1")
    cfg = make_cfg(1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["images_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": max(1, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg.describe() + f", training step at batch {CPU_SAMPLE_BATCH} (bounded sample) on host cores",
                   "global_batch": CPU_SAMPLE_BATCH, "parallelism": "cpu"},
        "cpu_baseline": {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["images_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "config1_cpu": {"value": r1["images_per_s"], "unit": UNIT, "ms_per_step": r1["ms_per_step"], "cores": r1["cores"],
                        "moe_layer_tokens_per_s": r1["images_per_s"] * 197, "sample": r1["sample"]},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# isolated layer (configs[1] layer shape): tokens/s and per-kernel breakdown
# ------------------------------------------------------------------------------------------------
PHASE_BYTES = {
    # algorithmic HBM bytes, x and y in bf16 under autocast (s = 2), R = kept pairs
    "moe_gate_fwd": lambda T, d, h, E, k, R: T * d * 2 + E * d * 4 + T * k * 8 + T * E * 4,
    "moe_dispatch_fwd": lambda T, d, h, E, k, R: T * d * 2 + R * d * 2 + T * k * 4,
    "moe_combine_fwd": lambda T, d, h, E, k, R: R * d * 2 + R * 4 + T * d * 2,
    "moe_combine_bwd": lambda T, d, h, E, k, R: T * d * 2 + 2 * R * d * 2 + 2 * R * 4,
    "moe_gate_dispatch_bwd": lambda T, d, h, E, k, R: R * d * 2 + T * d * 2 + 2 * T * E * 4 + T * k * 12,
    "colsum_db1": lambda T, d, h, E, k, R: (R // 32 + E) * h * 4,   # reads the dgelu epilogue's slab sums, not dU
    "colsum_db2": lambda T, d, h, E, k, R: R * d * 2,
    "moe_gate_wgrad": lambda T, d, h, E, k, R: T * d * 2 + T * E * 4,
}
GEMM_FLOPS = {
    "gemm_fc1": lambda R, d, h: 2 * R * d * h, "gemm_fc2": lambda R, d, h: 2 * R * d * h,
    "gemm_dgelu": lambda R, d, h: 2 * R * d * h, "gemm_dgrad": lambda R, d, h: 2 * R * d * h,
    "gemm_wgrad1": lambda R, d, h: 2 * R * d * h, "gemm_wgrad2": lambda R, d, h: 2 * R * d * h,
}


def layer_bench(peaks, iters=20, warmup=5, T=None, d=384, E=16, k=1, cf=1.25, gate="switch"):
    import fmoe
    from fmoe import _cabi as C

    T = PER_GPU_BATCH * 197 if T is None else T
    torch.manual_seed(0)
    h = 4 * d
    layer = fmoe.FMoETransformerMLP(E, d, h, torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(p=0.0)), top_k=k,
                                    gate=fmoe.make_gate(gate, cf)).cuda()
    x = torch.randn(T, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(T, d, device="cuda", dtype=torch.bfloat16)

    def it():
        y = layer(x)
        aux = layer.gate.get_loss()
        torch.autograd.backward([y, aux], [dy, torch.ones_like(aux) * AUX_COEF])
        x.grad = None
        for p in layer.parameters():
            p.grad = None

    for _ in range(warmup):
        it()
    torch.cuda.synchronize()
    # per-kernel times: CUDA events around every launch of an eager pass
    C.PROF.reset()
    C.PROF.enabled = True
    for _ in range(iters):
        torch.cuda._sleep(BACKLOG_CYCLES_LAYER)   # the device starts behind the host: event pairs bracket kernels, not launch gaps
        it()
    torch.cuda.synchronize()
    C.PROF.enabled = False
    prof = C.PROF.summary_ms()
    # host cost of one eager fwd+bwd (Python + ctypes + allocator), profiling off: what an eager training loop pays
    torch.cuda.synchronize()
    t_h = time.perf_counter()
    for _ in range(iters):
        it()
    host_ms = (time.perf_counter() - t_h) * 1e3 / iters
    torch.cuda.synchronize()
    # total: the same iteration captured as one CUDA graph and replayed (pure device time, no launch gaps)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        it()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        it()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    R = int(layer.last_kept.sum())
    phases = {}
    for tag, (n, mean_ms) in sorted(prof.items()):
        ent = {"ms": round(mean_ms, 4), "calls_per_iter": n // iters}
        if tag in PHASE_BYTES:
            gbs = PHASE_BYTES[tag](T, d, h, E, k, R) / (mean_ms * 1e-3) / 1e9
            ent.update(gbs=round(gbs, 1), frac_hbm=round(gbs / peaks["hbm"], 3))
        if tag in GEMM_FLOPS:
            tf = GEMM_FLOPS[tag](R, d, h) / (mean_ms * 1e-3) / 1e12
            ent.update(tflops=round(tf, 1), frac_tensor=round(tf / peaks["tf_burst"], 3))
        phases[tag] = ent
    gemm_ms = sum(v["ms"] for t, v in phases.items() if t in GEMM_FLOPS)
    flops = 12.0 * R * d * h
    return {
        "shape": {"T": T, "d": d, "h": h, "E": E, "k": k, "gate": gate, "capacity_factor": cf, "kept_pairs": R},
        "tokens_per_s_fwd_bwd": T / (ms * 1e-3), "ms_fwd_bwd": round(ms, 4), "host_enqueue_ms_fwd_bwd_eager": round(host_ms, 4),
        "timing": "CUDA-graph replay of fwd+bwd (total); CUDA events per launch, eager (kernels)",
        "ffn_tflops_fwd_bwd": round(flops / (gemm_ms * 1e-3) / 1e12, 1),
        "ffn_frac_of_burst_peak": round(flops / (gemm_ms * 1e-3) / 1e12 / peaks["tf_burst"], 3),
        "layer_tflops": round(flops / (ms * 1e-3) / 1e12, 1),
        "kernels": phases,
    }


def routing_self_check(inner, B: int, rank: int, world: int):
    """Bit-exactness of the routing integers on every rank at the benchmark's own shape (T = B * 197 tokens per rank):
    expert indices, per-expert counts / kept counts, the rank of every pair inside its expert and the drop mask, against
    oracle/gate_ref.c run on the same seeded shard with the first MoE layer's gate parameters."""
    import torch.distributed as dist
    from fmoe import functions as Fn
    from fmoe.distributed import slab_rows_for
    from oracle import moe_oracle as O

    layer = inner.moe_layers[0]
    T, d = B * 197, layer.d_model
    g = torch.Generator().manual_seed(4242 + rank)
    x = torch.randn(T, d, generator=g).to(torch.bfloat16)
    Wg = layer.gate.gate.weight.detach().float().cpu()
    bg = layer.gate.gate.bias.detach().float().cpu()
    spec = layer.gate.route_spec(T)
    slab = slab_rows_for(spec.capacity) if world > 1 else 0
    r = Fn.route(x.cuda(), Wg.cuda(), bg.cuda(), spec, slab_rows=slab)
    torch.cuda.synchronize()
    logits = O.gate_logits(x, Wg, bg)
    ref = O.route(logits, spec.top_k, spec.score_mode, spec.capacity)
    idx, pos, seg = r["idx"].cpu(), r["pos"].cpu(), r["seg_start"].cpu()
    live, live_ref = pos >= 0, ref.pos >= 0
    ok = torch.equal(idx, ref.idx) and torch.equal(r["count"].cpu(), ref.count) and torch.equal(r["kept"].cpu(), ref.kept)
    ok = ok and torch.equal(live, live_ref)
    if ok:   # rank of every kept pair inside its expert (the packed and the slab layouts differ only in the segment starts)
        e = idx[live].long()
        ok = torch.equal(pos[live] - seg[e], ref.pos[live_ref] - ref.seg_start[e])
    mine = torch.tensor([1.0 if ok else 0.0, float((~live).sum()), float((r["logits"].cpu() - logits).abs().max())],
                        device="cuda", dtype=torch.float64)
    rows = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(rows, mine)
    return {"what": "idx / count / kept / rank-in-expert / drop mask of a seeded [T, d] bf16 shard per rank vs oracle/gate_ref.c",
            "tokens_per_rank": T, "bit_exact": all(bool(t[0] == 1.0) for t in rows), "ranks_ok": [bool(t[0] == 1.0) for t in rows],
            "dropped_pairs_per_rank": [int(t[1]) for t in rows], "logits_max_abs_err_per_rank": [float(t[2]) for t in rows]}


# ------------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4"],
                    help="BASELINE.json configs[1..3]; c2 (default) is the single-GPU configuration the metric is quoted on")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the per-rank routing self-check against the C oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-layer", action="store_true", help="skip the isolated-layer breakdown (profiling runs)")
    ap.add_argument("--ep-transport", choices=["auto", "peer", "nccl"], default="auto",
                    help="expert-parallel row exchange: NVLink peer memory (fmoe/peer.py) or NCCL all-to-all on slabs (fmoe/distributed.py)")
    ap.add_argument("--e2e-breakdown", action="store_true",
                    help="diagnostic: after the timed regions, time the same K steps again device-only, with the staging copy only and "
                         "with the host->device prefetch only (e2e.breakdown_ms_per_step)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of replaying the captured training step")
    ap.add_argument("--profile-window", action="store_true",
                    help="cudaProfilerStart/Stop around the device-resident timed region (ncu --profile-from-start off)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # whole-step CUDA graph capture with NCCL inside needs the watchdog's async error handling off
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fmoe import _cabi as C
    from fmoe import distributed as fmoe_dist
    from moe_vit import MoEViT

    fmoe_dist.TRANSPORT = args.ep_transport
    peaks = load_peaks()
    cfg = make_cfg(world)
    torch.manual_seed(0)           # same dense/gate init on every rank; experts differ per rank below
    model = MoEViT(cfg).cuda()
    use_graph = not args.no_graph
    if world > 1:
        from fmoe.distributed import wrap_ddp
        torch.manual_seed(1 + rank)    # experts are sharded, not replicated: every rank initialises its own
        for lyr in model.moe_layers:
            lyr.experts.htoh4.reset_parameters()
            lyr.experts.h4toh.reset_parameters()
        side0 = torch.cuda.Stream()    # DDP built on a side stream so that its collectives can be graph-captured later
        side0.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side0):
            model = wrap_ddp(model, local_rank)
        torch.cuda.current_stream().wait_stream(side0)
        args.warmup = max(args.warmup, 11) if use_graph else args.warmup   # DDP needs 11 eager iterations before capture
    inner = model.module if hasattr(model, "module") else model
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05, fused=True, capturable=use_graph)

    B = PER_GPU_BATCH
    img_h, lab_h = synthetic_batch(B, 1000 + rank, pin=True)
    img_d, lab_d = img_h.cuda(), lab_h.cuda()
    bf16 = torch.bfloat16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        train_step(model, opt, img_d, lab_d, bf16)
    barrier()

    # ---- per-kernel timing pass (eager): CUDA events around every C-ABI call of the same training step.
    # (Events cannot be recorded inside a captured graph, so this pass directly precedes the timed region.)
    prof_steps = min(args.steps, 5)
    C.PROF.reset()
    C.PROF.enabled = True
    for _ in range(prof_steps):
        # eager launches are host-bound (the host needs longer to enqueue a step than the device to run it), so an event pair
        # around a launch would also time the idle gap before it; a spin kernel first puts the device ~60 ms behind the host
        torch.cuda._sleep(BACKLOG_CYCLES_STEP)
        train_step(model, opt, img_d, lab_d, bf16)
    barrier()
    C.PROF.enabled = False
    launches_per_step = C.PROF.launches // prof_steps
    kern = C.PROF.summary_ms()
    kept = [int(m.last_kept.sum()) for m in inner.moe_layers]
    # load-balance statistics of the last step (what /root/reference/main.py:945-951 would log)
    import fmoe
    lb_full = fmoe.load_balance_stats(inner)
    lb = {"drop_rate_mean": round(sum(v["drop_rate"] for v in lb_full.values()) / max(1, len(lb_full)), 5),
          "max_over_mean_load": round(max(v["max_over_mean_load"] for v in lb_full.values()), 3),
          "per_layer_drop_rate": {k_: round(v["drop_rate"], 5) for k_, v in lb_full.items()}}
    # parity evidence that travels with the number: every rank routes a seeded token shard of the benchmark's shape through
    # the CUDA gate / scan / dispatch and compares every routing integer with the C oracle (the checker, never the thing timed)
    parity = None if args.no_parity_check else routing_self_check(inner, B, rank, world)

    # ---- the whole training step as ONE CUDA graph (single GPU): the layer never synchronises the host and all
    # its buffers are static functions of the shapes, so forward + backward + fused AdamW capture as they are.
    graph, static_loss, mode = None, None, "eager"
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                train_step(model, opt, img_d, lab_d, bf16)        # one step on the capture stream (allocator warm-up)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = train_step(model, opt, img_d, lab_d, bf16, zero_grad=False)
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
            if not bool(torch.isfinite(static_loss)):
                raise RuntimeError("non-finite loss after graph replay")
            mode = "cuda_graph"
        except Exception as e:  # noqa: BLE001 — fall back to eager launches, say so in the JSON line
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
            return static_loss
        return train_step(model, opt, img_d, lab_d, bf16)

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile_window:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        run_step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # ~= ms_per_step means the host is the bottleneck
    ev1.record()
    barrier()
    if args.profile_window:
        torch.cuda.cudart().cudaProfilerStop()
    launches = launches_per_step * args.steps
    ms_total = ev0.elapsed_time(ev1)

    # ---- timed region 2: end to end from host buffers.  Every step's batch is copied from pinned host memory inside
    # the timed region; the copy of batch i+1 runs on a copy stream while step i computes (a prefetching loader),
    # lands in a staging buffer and is moved into the step's static input by a device-to-device copy.
    copy_stream = torch.cuda.Stream()
    img_stage, lab_stage = torch.empty_like(img_d), torch.empty_like(lab_d)
    staged = torch.cuda.Event()

    def prefetch():
        with torch.cuda.stream(copy_stream):
            img_stage.copy_(img_h, non_blocking=True)
            lab_stage.copy_(lab_h, non_blocking=True)
            staged.record(copy_stream)

    def e2e_loop(pipelined: bool) -> tuple[float, float]:
        """K steps from host buffers; returns (seconds, last loss).  Every step's loss is read on the host inside the timed
        region.  pipelined: the loss of step i lands in a pinned buffer through an asynchronous copy and is read after step
        i + 1 has been enqueued (a loop that logs one step behind), so the device never waits for the host round trip;
        otherwise `loss.item()` right after every step, the way /root/reference/engine.py:56-60 reads it."""
        loss_pin = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        loss_host = float("nan")
        barrier()
        t0 = time.perf_counter()
        prefetch()
        for i in range(args.steps):
            torch.cuda.current_stream().wait_event(staged)
            img_d.copy_(img_stage, non_blocking=True)
            lab_d.copy_(lab_stage, non_blocking=True)
            copy_stream.wait_stream(torch.cuda.current_stream())   # staging buffers are free again
            if i + 1 < args.steps:
                prefetch()
            loss = run_step()
            if pipelined:
                loss_pin[i & 1].copy_(loss.detach().float(), non_blocking=True)   # device -> host read of the step's result
                loss_ev[i & 1].record()
                if i > 0:
                    loss_ev[(i - 1) & 1].synchronize()
                    loss_host = float(loss_pin[(i - 1) & 1])
                    if not (loss_host == loss_host):
                        raise SystemExit("non-finite loss in the timed region")
            else:
                loss_host = loss.item()
        torch.cuda.synchronize()
        if pipelined:
            loss_host = float(loss_pin[(args.steps - 1) & 1])
        return time.perf_counter() - t0, loss_host

    # one batch's host -> device copy alone (device idle): what the first step of each loop below exposes
    h2d_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(copy_stream):
        h2d_ev[0].record(copy_stream)
        img_stage.copy_(img_h, non_blocking=True)
        lab_stage.copy_(lab_h, non_blocking=True)
        h2d_ev[1].record(copy_stream)
    copy_stream.synchronize()
    h2d_ms = h2d_ev[0].elapsed_time(h2d_ev[1])

    e2e_sync_s, loss_host = e2e_loop(pipelined=False)
    e2e_s, loss_host2 = e2e_loop(pipelined=True)
    if loss_host2 != loss_host2:
        loss_host = loss_host2
    clocks = sampler.stop() if sampler is not None else None
    if not (loss_host == loss_host):
        raise SystemExit("non-finite loss in the timed region")

    breakdown = None
    if args.e2e_breakdown:   # where the difference between `e2e` and `value` comes from: one ingredient of the e2e loop at a time
        def timed(body):
            barrier()
            t0_ = time.perf_counter()
            for i in range(args.steps):
                body(i)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0_) * 1e3 / args.steps

        def only_staging_copy(i):
            img_d.copy_(img_stage, non_blocking=True)
            lab_d.copy_(lab_stage, non_blocking=True)
            run_step()

        def only_prefetch(i):
            copy_stream.wait_stream(torch.cuda.current_stream())
            prefetch()
            run_step()

        done_ev, rec_ev = torch.cuda.Event(), torch.cuda.Event()
        done_ev.record()
        done_ev.synchronize()
        pin1 = torch.empty((), dtype=torch.float32).pin_memory()

        def only_event_record(i):
            rec_ev.record()
            run_step()

        def only_event_wait(i):
            torch.cuda.current_stream().wait_event(done_ev)   # completed long ago
            run_step()

        def only_free_running_h2d(i):   # the copy engine busy beside the step, no dependency on the main stream at all
            with torch.cuda.stream(copy_stream):
                img_stage.copy_(img_h, non_blocking=True)
            run_step()

        def only_loss_d2h(i):
            pin1.copy_(run_step().detach().float(), non_blocking=True)

        breakdown = {"device_only_again": round(timed(lambda i: run_step()), 3),
                     "plus_staging_copy": round(timed(only_staging_copy), 3),
                     "plus_h2d_prefetch_overlapped": round(timed(only_prefetch), 3),
                     "device_only_2": round(timed(lambda i: run_step()), 3),
                     "plus_event_record": round(timed(only_event_record), 3),
                     "plus_event_wait": round(timed(only_event_wait), 3),
                     "plus_free_running_h2d": round(timed(only_free_running_h2d), 3),
                     "plus_loss_d2h": round(timed(only_loss_d2h), 3),
                     "device_only_last": round(timed(lambda i: run_step()), 3)}

    t = torch.tensor([ms_total, e2e_s * 1e3, e2e_sync_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2])
    global_batch = B * world
    value = global_batch * args.steps / (ms_total * 1e-3)
    e2e_value = global_batch * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        d, depth, _ = cfg.dims
        h = 4 * d
        n_moe = len(inner.moe_layers)
        # roofline of the dominant hot-path kernel family: the grouped tcgen05 GEMM (6 launches per layer fwd+bwd)
        gemm_tags = [t_ for t_ in kern if t_.startswith("gemm_")]
        gemm_calls = sum(kern[t_][0] for t_ in gemm_tags)
        gemm_ms = sum(kern[t_][0] * kern[t_][1] for t_ in gemm_tags)         # total over the timed region
        flops_per_launch = 2.0 * (sum(kept) / max(1, n_moe)) * d * h           # every launch is one R x d x h contraction
        achieved = flops_per_launch * gemm_calls / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "moe::grouped_gemm_kernel (fc1+gelu, fc2, dgelu, dgrad, wgrad x2)",
                    "achieved": round(achieved, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": round(achieved / peaks["tf_sustained"], 4), "traffic": None,
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "timing": f"CUDA events around each launch in {prof_steps} eager steps of the same loop directly before the timed region, device kept backlogged behind the host by a spin kernel (no launch gaps inside the event pairs)",
                    "launches_timed": gemm_calls, "avg_launch_ms": round(gemm_ms / max(1, gemm_calls), 4),
                    "flops_per_launch": flops_per_launch,
                    "share_of_step": round(gemm_ms / prof_steps / (ms_total / args.steps), 4),
                    "per_op_ms": {t_: round(kern[t_][1], 4) for t_ in sorted(gemm_tags)}}
        # each launch against ITS roofline: min(tensor peak, arithmetic intensity x HBM bandwidth).  Algorithmic bytes
        # (DESIGN.md §4): operands read once, outputs written once, bf16 activations / weights, fp32 weight gradients.
        Rk, El_ = sum(kept) / max(1, n_moe), cfg.num_experts // world
        act_d, act_h, wts = Rk * d * 2, Rk * h * 2, El_ * d * h * 2
        op_bytes = {"gemm_fc1": act_d + wts + 2 * act_h, "gemm_fc2": act_h + wts + act_d, "gemm_dgelu": act_d + wts + 2 * act_h,
                    "gemm_dgrad": act_h + wts + act_d, "gemm_wgrad1": act_h + act_d + 2 * wts, "gemm_wgrad2": act_d + act_h + 2 * wts}
        per_op, ideal_ms, actual_ms = {}, 0.0, 0.0
        for t_ in sorted(gemm_tags):
            ms_ = kern[t_][1]
            t_tensor = flops_per_launch / (peaks["tf_sustained"] * 1e12) * 1e3
            t_hbm = op_bytes.get(t_, 0.0) / (peaks["hbm"] * 1e9) * 1e3
            floor = max(t_tensor, t_hbm)
            per_op[t_] = {"ms": round(ms_, 4), "tflops": round(flops_per_launch / (ms_ * 1e-3) / 1e12, 1),
                          "gbs": round(op_bytes.get(t_, 0.0) / (ms_ * 1e-3) / 1e9, 1), "bound": "hbm" if t_hbm > t_tensor else "tensor",
                          "floor_ms": round(floor, 4), "frac_of_its_roofline": round(floor / ms_, 3)}
            ideal_ms += floor
            actual_ms += ms_
        roofline["per_op"] = per_op
        roofline["frac_vs_per_op_roofline"] = round(ideal_ms / actual_ms, 4) if actual_ms > 0 else None
        # DRAM bytes per launch from the committed `ncu --set full` capture of the same six launches at this layer shape
        # (tools/gemm_traffic.py); the capture is of this shape only, so other configs report null
        roofline["algorithmic_bytes"] = round(sum(op_bytes.get(t_, 0.0) for t_ in gemm_tags) / max(1, len(gemm_tags)))
        tr_path = os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")
        if os.path.exists(tr_path) and CONFIG == "c2" and world == 1:
            tr = json.load(open(tr_path))
            roofline["traffic"] = tr["mean_dram_bytes_per_launch"]
            roofline["traffic_per_op"] = {k_: v_["dram_bytes"] for k_, v_ in tr["per_op"].items()}
            roofline["traffic_source"] = "profiles/gemm_dram_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, " + tr["source"] + ")"
        moe_ms = sum(n * m for n, m in kern.values())
        flops_img = inner.train_flops_per_image(kept_fraction=sum(kept) / max(1, n_moe) / (B * 197 * cfg.top_k))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "host_enqueue_ms_per_step": round(host_enqueue_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg.describe() + ", full training step (fwd + CE + aux loss + bwd + fused AdamW)",
                       "global_batch": global_batch, "per_gpu_batch": B, "tokens_per_moe_layer_per_gpu": B * 197,
                       "launch": mode,
                       "parallelism": "1 GPU" if world == 1 else f"dp{world} dense blocks + ep{world} experts (all-to-all)",
                       "cache": "inputs and activations larger than L2 (154 MB fp32 image batch, >1 GB activations per step)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": (img_h.numel() * img_h.element_size() + lab_h.numel() * lab_h.element_size()) * world,
                    "d2h_bytes_per_step": 4 * world,
                    "h2d_ms_per_batch_alone": round(h2d_ms, 3),
                    **({"breakdown_ms_per_step": breakdown} if breakdown is not None else {}),
                    "readback": "every step's loss is copied to pinned host memory and read (and checked finite) on the host one step "
                                "behind, inside the timed region; the first batch's host -> device copy is exposed, the others are "
                                "prefetched on a copy stream",
                    "sync_readback_value": global_batch * args.steps / (e2e_sync_ms * 1e-3),
                    "sync_readback_ms_per_step": e2e_sync_ms / args.steps,
                    "sync_readback": "loss.item() after every step (the device drains before the next step is enqueued), "
                                     "as /root/reference/engine.py:56-60 does"},
            "gpu_launches": launches,
            "roofline": roofline,
            "model_roofline": {"train_flops_per_image": flops_img, "achieved_tflops": round(flops_img * value / 1e12, 1),
                               "frac_of_sustained_bf16": round(flops_img * value / 1e12 / (peaks["tf_sustained"] * world), 4),
                               "moe_kernels_share_of_step": round(moe_ms / prof_steps / (ms_total / args.steps), 4)},
        }
        if world > 1:   # expert-parallel exchange: per-call CUDA-event times of the eager profiling pass (rank 0)
            transport = getattr(inner.moe_layers[0], "_ep_transport", "nccl")
            peer_tags = ("moe_ep_exchange_counts", "moe_ep_barrier", "moe_dispatch_fwd_peer", "moe_combine_fwd_peer", "moe_combine_bwd_peer",
                         "moe_gate_dispatch_bwd_peer")
            ep_tags = [t_ for t_ in kern if t_.startswith("a2a_") or t_ in ("moe_ep_repack", "moe_ep_tables") or t_ in peer_tags]
            sync_tags = [t_ for t_ in ep_tags if t_ in ("moe_ep_exchange_counts", "moe_ep_barrier") or t_.startswith("a2a_")]
            line["expert_parallel"] = {
                "transport": transport,
                "per_call_ms": {t_: round(kern[t_][1], 4) for t_ in sorted(ep_tags)},
                "calls_per_step": {t_: kern[t_][0] // prof_steps for t_ in sorted(ep_tags)},
                "ms_per_step": round(sum(kern[t_][0] * kern[t_][1] for t_ in ep_tags) / prof_steps, 3),
                "sync_ms_per_step": round(sum(kern[t_][0] * kern[t_][1] for t_ in sync_tags) / prof_steps, 3),
                "note": ("NVLink peer memory: dispatch / combine kernels write and read the owners' packed buffers in place, device-side "
                         "barriers (eager-mode times include the inter-rank skew they absorb); the peer kernels' times include their local work"
                         if transport == "peer" else
                         "NCCL all_to_all_single on fixed slabs [W, E_local, slab_rows, d] bf16 + repack passes")}
        line["config"]["name"] = CONFIG
        line["load_balance"] = lb
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_layer:
            line["moe_layer"] = layer_bench(peaks, d=d, E=cfg.num_experts, k=cfg.top_k, gate=cfg.gate)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=1000, warmup=1, budget_s=20.0)
            line["cpu_baseline"] = {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        emit(line)
    if world > 1:
        # no barrier / destroy_process_group here: tearing NCCL down after its collectives were captured into a CUDA
        # graph can block forever (seen once on 2 GPUs, round 1); every rank is already past its last collective.
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
