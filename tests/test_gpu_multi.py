"""Multi-GPU row (SURVEY.md 8e / BASELINE config 4): a mixed-size image list scored on N ranks over NCCL with
area-balanced shards and ONE all-gather must return, per image, exactly the bits one rank alone returns.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SHAPES = [(16, 16), (16, 24), (24, 16), (8, 8), (16, 16), (32, 16), (8, 24), (16, 24), (24, 24), (8, 8), (16, 32), (9, 13)]
N_DRAWS = 3


def _inputs():
    g = torch.Generator().manual_seed(404)
    lat = [torch.randn(1, 4, h, w, generator=g) for h, w in SHAPES]
    ctx = [torch.randn(77, 768, generator=g) for _ in range(2)]
    return lat, ctx


def _score(eng, lat, idx):
    """T map of every image in idx (same-shape images batched, (eps, t) re-seeded per shape as D.draws does)"""
    out = {}
    by_shape = {}
    for i in idx:
        by_shape.setdefault(SHAPES[i], []).append(i)
    for (h, w), ii in by_shape.items():
        torch.manual_seed(42)
        x = torch.empty(1, 4, h, w, device=eng.device)
        ns, ts = zip(*[(torch.randn_like(x), torch.randint(0, 1000, (1,), device=eng.device)) for _ in range(N_DRAWS)])
        _, T = eng.typicality(torch.cat([lat[i] for i in ii]), torch.cat(ns), torch.cat(ts).long(), [1, 0])
        for k, i in enumerate(ii):
            out[i] = T[k, 0].clone()
    return [out[i] for i in idx]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from diff_mining_b200 import parallel
        from diff_mining_b200.engine import Engine
        from oracle import sd15

        eng = Engine(rank)
        eng.load_state_dict(sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0), "unet.")
        eng.finalize()
        eng.set_schedule(*sd15.schedule_tables())
        lat, ctx = _inputs()
        for s, c in enumerate(ctx):
            eng.set_context(s, c)
        maps = parallel.run_sharded_mixed(SHAPES, lambda idx: _score(eng, lat, idx), device=eng.device)
        ok = True
        if rank == 0:
            alone = _score(eng, lat, list(range(len(SHAPES))))   # the 1-rank answer
            ok = all(torch.equal(a, b) for a, b in zip(alone, maps))
        shards = parallel.shard_by_area([h * w for h, w in SHAPES], world)
        q.put((rank, bool(ok), len(shards[rank])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_mixed_size_sharding_is_bit_identical_over_nccl():
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == len(SHAPES)
