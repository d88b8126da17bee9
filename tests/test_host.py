"""CPU tests of the host-side logic: sharding + the single all-gather (gloo, world_size 2), the reference-surface
helpers of typicality.D, prompt templating, the T(x|c) reduction."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diff_mining_b200 import parallel
from diff_mining_b200 import typicality as ty
from oracle import sd15


def test_shard_indices_round_robin():
    # compute.py:339 -> lines[i::sub_split]
    assert parallel.shard_indices(10, 4, 0) == [0, 4, 8]
    assert parallel.shard_indices(10, 4, 3) == [3, 7]
    assert parallel.shard_indices(3, 8, 5) == []
    allidx = sorted(i for r in range(8) for i in parallel.shard_indices(10000, 8, r))
    assert allidx == list(range(10000))


def _gather_worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def compute_local(idx):
            return torch.stack([torch.full((3, 5), float(i)) + torch.arange(5.0) for i in idx]) if idx else torch.zeros(0, 3, 5)

        out = parallel.run_sharded(n_total, compute_local)
        ref = torch.stack([torch.full((3, 5), float(i)) + torch.arange(5.0) for i in range(n_total)])
        q.put((rank, bool(torch.equal(out, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8, 1])
def test_gather_tmaps_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_schedule_matches_oracle():
    a, b = ty.scaled_linear_schedule()
    oa, ob = sd15.schedule_tables()
    assert torch.equal(a, oa) and torch.equal(b, ob)


def test_prompt_templates():
    # compute.py:41-48
    assert ty.prompt_for("cars", "1975") == "A car at the 1975's."
    assert ty.prompt_for("cars", "") == "A car."
    assert ty.prompt_for("places", "art_gallery") == "Image of art gallery."
    assert ty.prompt_for("geo", "France") == "France"
    assert ty.prompt_for("geo", "") == ""
    assert ty.prompt_for("faces", "") == "Portrait."


def test_typicality_map_reduction():
    # cluster.py:112-123: channel mean -> bilinear -> uncond - cond -> mean over N
    g = torch.randn(5, 2, 4, 8, 8).abs().half()
    T = ty.typicality_map(g)
    ref = (g.float()[:, 1].mean(1) - g.float()[:, 0].mean(1)).mean(0)
    torch.testing.assert_close(T, ref)
    T2 = ty.typicality_map(g, size=(64, 64))
    assert T2.shape == (64, 64)
    # linear ops commute: upsampling the reduced map equals reducing the upsampled maps
    up = torch.nn.functional.interpolate(ref[None, None], (64, 64), mode="bilinear")[0, 0]
    torch.testing.assert_close(T2, up, atol=1e-5, rtol=1e-5)


class _FakeSD:
    device = torch.device("cpu")
    scheduler = type("S", (), {"num_train_timesteps": 1000})()


def test_D_helpers_follow_reference():
    from PIL import Image

    d = ty.D(_FakeSD(), "/tmp/typ", "cars", seed=42, N=3, t_min=0.1, t_max=0.7)
    assert d.get_path("/data/cars/1975__a.jpg") == "/tmp/typ/1975__a.npy"     # compute.py:162-163
    assert d.get_path("x/b.png") == "/tmp/typ/b.npy"
    img = Image.new("RGB", (640, 480))
    assert d.rescale(img).size == (341, 256)                                     # compute.py:166-174
    assert ty.D(_FakeSD(), "", "places").rescale(Image.new("RGB", (300, 600))).size == (512, 1024)
    x = d.load_image(Image.new("RGB", (16, 8), (255, 0, 127)))
    assert x.shape == (1, 3, 8, 16) and x[0, 0, 0, 0] == 1.0 and x[0, 1, 0, 0] == -1.0
    lat = torch.zeros(1, 4, 4, 4)
    n1, t1 = d.draws(lat)
    n2, t2 = d.draws(lat)
    assert n1.shape == (3, 4, 4, 4) and t1.shape == (3,) and t1.dtype == torch.int64
    assert torch.equal(n1, n2) and torch.equal(t1, t2)          # re-seeded per image (compute.py:139)
    assert int(t1.min()) >= 100 and int(t1.max()) < 700
    # identical to the reference's call sequence
    torch.manual_seed(42)
    ref = [(torch.randn_like(lat), torch.randint(100, 700, (1,))) for _ in range(3)]
    assert torch.equal(n1, torch.cat([r[0] for r in ref])) and torch.equal(t1, torch.cat([r[1] for r in ref]))


def test_exists_and_call_roundtrip():
    with tempfile.TemporaryDirectory() as td:
        d = ty.D(_FakeSD(), td, "geo")
        assert not d.exists("a/France__1.jpg")
        arr = np.random.rand(2, 2, 4, 3, 3).astype(np.float16)
        np.save(open(d.get_path("a/France__1.jpg"), "wb"), arr)
        assert d.exists("a/France__1.jpg")
        np.testing.assert_array_equal(d("a/France__1.jpg"), arr)


# ---------------------------------------------------------------------------- mixed-size sharding (config 4 / geo)
def test_shard_by_area_is_balanced_and_deterministic():
    rng = np.random.RandomState(0)
    areas = [int(h * w) for h, w in zip(rng.randint(24, 96, 1000), rng.randint(24, 96, 1000))]
    shards = parallel.shard_by_area(areas, 8)
    assert sorted(i for s in shards for i in s) == list(range(1000))
    loads = [sum(areas[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(areas)            # LPT bound
    assert max(loads) / (sum(areas) / 8) < 1.01
    assert shards == parallel.shard_by_area(list(areas), 8)  # no dependence on anything but the areas
    assert parallel.shard_by_area([5, 5, 5], 1) == [[0, 1, 2]]
    assert parallel.shard_by_area([], 4) == [[], [], [], []]


def _shapes_mixed(n):
    rng = np.random.RandomState(3)
    return [(int(h), int(w)) for h, w in zip(rng.randint(2, 9, n), rng.randint(2, 9, n))]


def _map_of(i, shape):
    return (torch.arange(shape[0] * shape[1], dtype=torch.float32).view(shape) * 0.5 + i).contiguous()


def _mixed_worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = _shapes_mixed(n_total)
        out = parallel.run_sharded_mixed(shapes, lambda idx: [_map_of(i, shapes[i]) for i in idx])
        ok = all(torch.equal(out[i], _map_of(i, shapes[i])) for i in range(n_total))
        # async gather of equal-size maps completes on result()
        idx = parallel.shard_indices(n_total, world, rank)
        loc = torch.stack([torch.full((2, 3), float(i)) for i in idx]) if idx else torch.zeros(0, 2, 3)
        h = parallel.gather_tmaps_async(loc, n_total)
        full = h.result()
        ok = ok and torch.equal(full, torch.stack([torch.full((2, 3), float(i)) for i in range(n_total)]))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [13, 2])
def test_run_sharded_mixed_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_mixed_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_run_sharded_mixed_single_process():
    shapes = _shapes_mixed(9)
    out = parallel.run_sharded_mixed(shapes, lambda idx: [_map_of(i, shapes[i]) for i in idx])
    assert all(torch.equal(out[i], _map_of(i, shapes[i])) for i in range(9))


# ---------------------------------------------------------------------------- context-slot allocator
class _FakeEngine:
    def __init__(self):
        self.uploaded = []

    def _upload_context(self, slot, ctx):
        self.uploaded.append((slot, float(ctx[0, 0])))


def _ctx(i):
    return torch.full((77, 768), float(i))


def test_context_slots_lru_pinning_and_overflow():
    from diff_mining_b200.engine import ContextSlots

    fe = _FakeEngine()
    cs = ContextSlots(fe, n_slots=4)
    assert cs.acquire([_ctx(1), _ctx(2), _ctx(1)]) == [0, 1, 0]          # content-keyed: equal values share a slot
    assert cs.acquire([_ctx(2).clone()]) == [1] and len(fe.uploaded) == 2  # other storage, same content: no upload
    assert cs.acquire([_ctx(3), _ctx(4)]) == [2, 3]
    # all four slots hold 1..4; a call needing {5, 2} must keep 2 and recycle the least recently used of the others (1)
    assert cs.acquire([_ctx(5), _ctx(2)]) == [0, 1]
    assert cs.slot_of[ContextSlots.key(_ctx(3))] == 2 and ContextSlots.key(_ctx(1)) not in cs.slot_of
    # one call may never evict a context it needs itself: 4 new contexts replace everything, 5 is too many
    assert sorted(cs.acquire([_ctx(i) for i in (10, 11, 12, 13)])) == [0, 1, 2, 3]
    with pytest.raises(RuntimeError, match="distinct text contexts"):
        cs.acquire([_ctx(i) for i in range(20, 25)])
    # manual slots leave the pool
    cs.mark_manual(0)
    got = cs.acquire([_ctx(30), _ctx(31), _ctx(32)])
    assert 0 not in got and len(set(got)) == 3
    with pytest.raises(RuntimeError):
        cs.acquire([_ctx(i) for i in range(40, 44)])


def test_context_slots_many_categories():
    """the reference builds SD with every category (365 for places): a (category, "") pair per call must always fit"""
    from diff_mining_b200.engine import ContextSlots

    fe = _FakeEngine()
    cs = ContextSlots(fe, n_slots=64)
    uncond = _ctx(0)
    for c in range(1, 366):
        s = cs.acquire([_ctx(c), uncond])
        assert s[0] != s[1]
    assert len(fe.uploaded) == 366   # the unconditional context stayed resident the whole time (most recently used)


def test_groupnorm_shifted_sum_algebra_two_sources():
    """the algebra the GroupNorm kernels fold with (norm.cuh: S_c = sum(x - k_c), Q_c = sum((x - k_c)^2) per channel with an
    ARBITRARY per-(image, channel) shift k_c, partials over disjoint pixel sets, groups that straddle two concatenated
    sources) restated in numpy float64 and checked against torch.group_norm on cat(src0, src1) -- large channel means
    included, where E[x^2] - E[x]^2 in fp32 would cancel"""
    g = np.random.default_rng(11)
    HW, C0, C1, groups = 256, 80, 48, 32      # cpg = 4: group boundaries do not care about the source boundary
    x0 = g.normal(size=(HW, C0)) + g.normal(size=C0) * 50.0
    x1 = g.normal(size=(HW, C1)) * 3.0 - 20.0
    xs, out = (x0, x1), []
    cpg = (C0 + C1) // groups
    # per-source records: 8 partial entries (disjoint pixel blocks) of (S, Q) about a shift that is NOT the mean
    S, Q, K = [], [], []
    for x in xs:
        k = x[0] + g.normal(size=x.shape[1])                     # "value at the image's first pixel" + anything
        parts = np.split(np.arange(HW), 8)
        s = np.stack([(x[p] - k).sum(0) for p in parts])
        q = np.stack([((x[p] - k) ** 2).sum(0) for p in parts])
        S.append(s.astype(np.float32).sum(0, dtype=np.float32))   # fp32 fold in index order, like the kernel
        Q.append(q.astype(np.float32).sum(0, dtype=np.float32))
        K.append(k.astype(np.float32))
    S, Q, K = (np.concatenate(v).astype(np.float64) for v in (S, Q, K))
    n = float(HW)
    mean = np.empty(groups)
    rstd = np.empty(groups)
    for gi in range(groups):
        c = slice(gi * cpg, (gi + 1) * cpg)
        mean[gi] = (S[c] + n * K[c]).sum() / (n * cpg)
        d = mean[gi] - K[c]
        m2 = (Q[c] - 2.0 * d * S[c] + n * d * d).sum()
        rstd[gi] = 1.0 / np.sqrt(max(m2 / (n * cpg), 0.0) + 1e-5)
    xcat = np.concatenate(xs, axis=1)
    got = (xcat - np.repeat(mean, cpg)) * np.repeat(rstd, cpg)
    ref = torch.nn.functional.group_norm(torch.from_numpy(xcat.T[None].copy()), groups, eps=1e-5)[0].T.numpy()
    assert np.abs(got - ref).max() < 2e-4
