"""Expert parallelism: host-side logic on CPU (gloo, world_size 2 and 4) and the 2-GPU parity run (NCCL)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path[:0] = [ROOT, os.path.join(ROOT, "slim-switch-moe-vit_b200")]
        import fmoe
        from fmoe import distributed as D
        from moe_vit import MoEViT, MoEViTConfig

        # 1. slab exchange: chunk j of rank r lands as chunk r of rank j
        El0, slab, d = 3, 4, 2
        send = torch.zeros(world, El0 * slab * d)
        for j in range(world):
            send[j] = 1000 * rank + 10 * j + torch.arange(El0 * slab * d) * 0.001
        recv = D.all_to_all_slabs(send)
        for s in range(world):
            assert torch.equal(recv[s], 1000 * s + 10 * rank + torch.arange(El0 * slab * d) * 0.001)
        kept = torch.arange(world * El0, dtype=torch.int32).view(world, El0) + 100 * rank
        kr = D.all_to_all_slabs(kept)
        assert kr.tolist() == [[100 * s + rank * El0 + e for e in range(El0)] for s in range(world)]
        assert D.slab_rows_for(3940) == 4096 and D.slab_rows_for(256) == 256 and D.slab_rows_for(1) == 256

        # 2. a model with sharded experts: expert parameters are kept out of DDP and carry the 1/W hook
        cfg = MoEViTConfig(size="tiny", num_experts=8, top_k=1, capacity_factor=1.25, moe_stride=2, world_size=world)
        torch.manual_seed(0)
        model = MoEViT(cfg)
        layer = model.moe_layers[0]
        El = 8 // world
        assert layer.num_expert == El and layer.world_size == world and layer.gate.gate.out_features == 8
        names = D.mark_expert_parallel(model)
        assert len(names) == 4 * len(model.moe_layers) and all(".experts." in n for n in names)
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        assert set(names) <= set(ddp.parameters_to_ignore)
        p = layer.experts.htoh4.weight
        (p.sum() * 3.0).backward()
        assert torch.allclose(p.grad, torch.full_like(p, 3.0 / world))
        # replicated parameters still all-reduce (mean over ranks)
        g = model.head.weight
        (g.sum() * float(rank + 1)).backward()

        # 2b. EP-aware checkpointing: the gathered state dict has global expert shapes and round-trips
        full = D.full_state_dict(model)
        w_name = "blocks.1.mlp.experts.htoh4.weight"
        assert full[w_name].shape[0] == 8 and torch.equal(full[w_name][rank * El:(rank + 1) * El], layer.experts.htoh4.weight)
        single = MoEViT(MoEViTConfig(size="tiny", num_experts=8, top_k=1, capacity_factor=1.25, moe_stride=2, world_size=1))
        single.load_state_dict(full)            # what rank 0 saves loads into a single-process model unchanged
        with torch.no_grad():
            layer.experts.htoh4.weight.zero_()
        D.load_full_state_dict(model, full)
        assert torch.equal(layer.experts.htoh4.weight, full[w_name][rank * El:(rank + 1) * El])

        # 3. gates without a capacity are refused under expert parallelism, before any kernel is touched
        naive = fmoe.FMoETransformerMLP(2, 64, 256, torch.nn.GELU(), top_k=2, world_size=world)
        try:
            naive(torch.randn(5, 64))
            raise AssertionError("NaiveGate under EP must raise")
        except NotImplementedError:
            pass
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 4])
def test_ep_host_logic_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_ep_parity_multi_gpu(transport):
    """Expert-parallel layer on every visible GPU (2, 4 or 8; `gpurun --gpus N`) against the single-process oracle of each
    rank's token shard — rows exchanged through NVLink peer memory (fmoe/peer.py) and through NCCL slabs."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2|4|8)")
    W = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={W}", "--master-addr", "127.0.0.1",
           "--master-port", "29631" if transport == "peer" else "29633", os.path.join(ROOT, "tests", "ep_worker.py"), transport]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    log_dir = os.path.join(ROOT, "gpurun_out")      # the workers' per-case lines, kept as evidence next to the pytest log
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"ep_worker_{transport}_w{W}.log"), "w") as f:
            f.write(out.stdout)
    assert out.returncode == 0 and out.stdout.count("EP_OK") == W, out.stdout[-3000:] + out.stderr[-3000:]
