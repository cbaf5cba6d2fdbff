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
    merged = D.gather_detections(mine, n_scenes)
    ok = len(merged) == n_scenes
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
    for n_scenes in (5,):
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_scenes, q)) for r in range(2)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=120) for _ in ps)
        [p.join(30) for p in ps]
        assert res == [(0, True), (1, True)]


def test_gather_single_process_is_identity():
    d = [_fake_dets(i) for i in range(3)]
    assert D.gather_detections(d, 3) is not None and len(D.gather_detections(d, 2)) == 2
