"""CPU, world_size 2 over gloo: scene sharding + the result gather (the only collective of the inference path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cagroup3d_b200 import dist as D


def test_shard_indices_matches_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    for n in (1, 5, 8, 13):
        for world in (1, 2, 4, 8):
            for r in range(world):
                want = list(DistributedSampler(range(n), num_replicas=world, rank=r, shuffle=False))
                assert D.shard_indices(n, r, world) == want


def _fake_dets(scene: int):
    g = torch.Generator().manual_seed(scene)
    n = (scene * 7) % 5                                   # some scenes have no detections
    return {"pred_boxes": torch.rand((n, 6 if scene % 2 else 7), generator=g), "pred_scores": torch.rand((n,), generator=g),
            "pred_labels": torch.randint(0, 18, (n,), generator=g)}


def _worker(rank, world, port, n_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [_fake_dets(i) for i in D.shard_indices(n_scenes, rank, world)]
    D._GATHER_CAP["rows"] = 2                             # the fixed-capacity buffer overflows: every rank grows it alike
    merged = D.gather_detections(mine, n_scenes)
    ok0 = D._GATHER_CAP["rows"] >= max(sum(len(d["pred_boxes"]) for d in mine), 2)
    again = D.gather_detections(mine, n_scenes, to_host=True)           # second call: one collective, no retry
    ok0 &= all(torch.equal(a["pred_boxes"], b["pred_boxes"]) for a, b in zip(merged, again))
    ok = ok0 and len(merged) == n_scenes
    for i, d in enumerate(merged):
        w = _fake_dets(i)
        wb = w["pred_boxes"] if w["pred_boxes"].shape[1] == 7 else torch.cat([w["pred_boxes"], torch.zeros((len(w["pred_boxes"]), 1))], 1)
        ok &= torch.equal(d["pred_boxes"], wb) and torch.equal(d["pred_scores"], w["pred_scores"]) \
            and torch.equal(d["pred_labels"], w["pred_labels"])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    for n_scenes in (5, 12):
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_scenes, q)) for r in range(2)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=120) for _ in ps)
        [p.join(30) for p in ps]
        assert res == [(0, True), (1, True)]


def test_gather_single_process_normalises_like_the_collective():
    """world size 1: no collective, but the same output format as with N ranks (7-wide boxes, int64 labels)."""
    d = [_fake_dets(i) for i in range(4)]
    out = D.gather_detections(d, 4)
    assert len(out) == 4 and len(D.gather_detections(d, 2)) == 2
    for a, b in zip(out, d):
        assert a["pred_boxes"].shape == (len(b["pred_boxes"]), 7) and a["pred_labels"].dtype == torch.int64
        assert torch.equal(a["pred_boxes"][:, :b["pred_boxes"].shape[1]], b["pred_boxes"]) and torch.equal(a["pred_scores"], b["pred_scores"])
    host = D.detections_to_host(d)
    assert all(torch.equal(h["pred_labels"], b["pred_labels"]) for h, b in zip(host, d))


# ---- training collectives: gradient all-reduce in flat buckets, reduce_mean -------------------------------------
def _toy_model(seed=0):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.ReLU(), torch.nn.Linear(33, 5, bias=False),
                            torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    m[2].weight.requires_grad_(False)            # a frozen parameter stays out of the buckets
    return m


def _toy_batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn((16, 7), generator=g), torch.randn((16, 3), generator=g)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _toy_model()
    red = D.GradientAllReducer(m.parameters(), bucket_mb=0.0005)      # ~131 floats per bucket -> more than one bucket
    ok = len(red.buckets) >= 2
    for step in range(2):
        red.zero_grad()
        x, y = _toy_batch(rank)
        ((m(x) - y) ** 2).mean().backward()
        n_coll = red.reduce()
        ok &= n_coll == len(red.buckets)
        # the average of the per-rank gradients, computed independently on every rank
        want = None
        for r in range(world):
            mr = _toy_model()
            xr, yr = _toy_batch(r)
            ((mr(xr) - yr) ** 2).mean().backward()
            gs = [p.grad for p in mr.parameters() if p.requires_grad]
            want = gs if want is None else [a + b for a, b in zip(want, gs)]
        got = [p.grad for p in m.parameters() if p.requires_grad]
        ok &= all(torch.allclose(a, b / world, rtol=1e-6, atol=1e-7) for a, b in zip(got, want))
        ok &= all(p.grad.data_ptr() == red._view(p).data_ptr() for p in red.params)
    # a gradient replaced behind the reducer's back (optimizer.zero_grad(set_to_none=True)) is still reduced
    for p in red.params:
        p.grad = None
    next(iter(red.params)).grad = torch.full_like(next(iter(red.params)), float(rank + 1))
    red.reduce()
    ok &= torch.allclose(next(iter(red.params)).grad, torch.full_like(next(iter(red.params)), (1 + world) / 2))
    ok &= all(float(p.grad.abs().max()) == 0 for p in list(red.params)[1:])
    # reduce_mean: cagroup_utils.py:6-12
    t = torch.tensor([float(rank), 10.0 * (rank + 1)])
    rm = D.reduce_mean(t)
    ok &= torch.allclose(rm, torch.tensor([(world - 1) / 2, 10.0 * (world + 1) / 2])) and float(t[0]) == float(rank)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_allreduce_and_reduce_mean_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert res == [(0, True), (1, True)]


def test_gradient_reducer_single_process():
    m = _toy_model()
    red = D.GradientAllReducer(m.parameters(), bucket_mb=64)
    assert len(red.buckets) == 1 and red.nbytes >= 4 * sum(p.numel() for p in m.parameters() if p.requires_grad)
    x, y = _toy_batch(0)
    ((m(x) - y) ** 2).mean().backward()
    ref = _toy_model()
    ((ref(x) - y) ** 2).mean().backward()
    assert red.reduce() == 0
    for a, b in zip(m.parameters(), ref.parameters()):
        if a.requires_grad:
            assert torch.equal(a.grad, b.grad)
    t = torch.tensor([3.0])
    assert D.reduce_mean(t) is t
    red.zero_grad()
    assert all(float(p.grad.abs().max()) == 0 for p in red.params)


# ---- the training step under two ranks: every rank its own scene, gradients averaged through the flat buckets -------------
def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from cagroup3d_b200 import model_init, ops, sparse as S, synthetic, train_step as TS, train_targets as TT
    from oracle import me_cpu as me
    from tests import cabi_emulator as E
    E.install()                                                   # C ABI emulated on the CPU (tests/cabi_emulator.py)
    TT._require_cuda = lambda t: None
    ops._chk = lambda *ts: None

    def voxelize_cpu(points, voxel_size):
        c = points[:, :4].clone()
        c[:, 1:] /= voxel_size
        ox = me.from_points(c, points[:, 4:])
        mgr = S.Manager()
        cm = E.cpu_map(ox.C, 1, mgr)
        mgr.by_stride[1] = cm
        return S.SparseTensor(ox.F.float().contiguous(), cm, mgr)
    TS.voxelize = voxelize_cpu

    def batch_of(scene_index):
        p, b, s, m = synthetic.make_scene(1000 * 7 + scene_index, 250, n_classes=18, return_masks=True)
        bt = synthetic.collate_batch([(p, b)])
        return {"points": torch.from_numpy(bt["points"]).clone(), "batch_size": 1, "gt_boxes": torch.from_numpy(bt["gt_boxes"]).float(),
                "semantic_mask": [s], "instance_mask": [m]}

    def grads_of(scene_index, reducer_on):
        model = model_init.seeded_model(18, False, seed=4).train()
        params = [p for n, p in model.named_parameters() if n.startswith(("backbone_3d.", "dense_head.semantic_conv", "dense_head.offset_block"))]
        red = D.GradientAllReducer(params, bucket_mb=16) if reducer_on else None
        tb = TS.partial_training_step(model, batch_of(scene_index), None, red, impl="simt")
        return [p.grad.clone() for p in params], tb

    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    got, tb = grads_of(rank, True)                                # my scene, gradients averaged over the ranks
    mine, _ = grads_of(rank, False)                               # the same step without the reducer ...
    flat = torch.cat([g.reshape(-1) for g in mine])
    parts = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)                                  # ... averaged by the test itself
    want = sum(parts) / world
    have = torch.cat([g.reshape(-1) for g in got])
    worst = float((have - want).norm()) / (float(want.norm()) + 1e-12)
    differs = float((have - flat).norm()) / (float(flat.norm()) + 1e-12)          # and it is not just my own gradient
    ok = bool(np.isfinite(tb["loss"])) and worst < 1e-4 and differs > 1e-2
    q.put((rank, bool(ok), worst, differs))
    dist.destroy_process_group()


def test_training_step_gradients_averaged_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=600) for _ in ps)
    [p.join(60) for p in ps]
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
