"""Multi-GPU plumbing of the inference path: scenes shard by index, one process per GPU, and ONE collective
-- the gather of the per-scene detections at the end.

Replaces, with the same ordering semantics,
  * pcdet/datasets/__init__.py:28-48  DistributedSampler(shuffle=False): index list padded by wrap-around to a
    multiple of world_size, rank r takes indices r, r + W, r + 2W, ...
  * pcdet/utils/common_utils.py:202-223  merge_results_dist: every rank pickles its list to a shared tmpdir,
    two barriers, rank 0 re-interleaves (zip over ranks) and truncates to the dataset size.
Here the detections travel as fixed-width rows [box(7) | score | label] in one padded all_gather (NCCL over
NVLink on the GPU box, gloo in the CPU tests); no files, no pickling, every rank gets the result.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

ROW = 9     # x y z dx dy dz yaw score label


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """indices rank `rank` evaluates (DistributedSampler(shuffle=False) semantics, with wrap-around padding)."""
    if n == 0:
        return []
    total = -(-n // world) * world
    idx = list(range(n))
    while len(idx) < total:
        idx += idx[:total - len(idx)]
    return idx[rank:total:world]


def pack_detections(pred_dicts: Sequence[dict], device=None) -> tuple:
    """list of {'pred_boxes' (n,6|7), 'pred_scores' (n), 'pred_labels' (n)} -> (rows (sum n, 9) f32, counts (S,) i64)."""
    rows, counts = [], []
    for d in pred_dicts:
        b = d["pred_boxes"].float()
        if b.shape[1] == 6:
            b = torch.cat([b, b.new_zeros((len(b), 1))], 1)
        rows.append(torch.cat([b[:, :7], d["pred_scores"].float()[:, None], d["pred_labels"].float()[:, None]], 1))
        counts.append(len(b))
    dev = device if device is not None else (rows[0].device if rows else torch.device("cpu"))
    r = torch.cat(rows) if rows else torch.zeros((0, ROW), device=dev)
    return r.to(dev), torch.tensor(counts, dtype=torch.int64, device=dev)


def unpack_detections(rows: torch.Tensor, counts: torch.Tensor) -> List[dict]:
    out, o = [], 0
    for n in counts.tolist():
        r = rows[o:o + n]
        out.append({"pred_boxes": r[:, :7], "pred_scores": r[:, 7], "pred_labels": r[:, 8].long()})
        o += n
    return out


def gather_detections(pred_dicts: Sequence[dict], n_total: int, group=None) -> List[dict]:
    """all ranks -> the detections of scenes 0..n_total-1 in dataset order (merge_results_dist semantics:
    interleave the ranks' lists, drop the wrap-around padding).  Single process: identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(pred_dicts)[:n_total]
    world = dist.get_world_size(group)
    rows, counts = pack_detections(pred_dicts)
    dev = rows.device
    # sizes first (scenes per rank are equal by construction; detections are not)
    meta = torch.tensor([rows.shape[0], counts.shape[0]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_rows = max(int(m[0]) for m in metas)
    max_scn = max(int(m[1]) for m in metas)
    buf = torch.zeros((max_rows * ROW + max_scn,), dtype=torch.float32, device=dev)
    buf[:rows.numel()] = rows.reshape(-1)
    buf[max_rows * ROW:max_rows * ROW + counts.numel()] = counts.float()
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    per_rank = []
    for m, b in zip(metas, bufs):
        nr, ns = int(m[0]), int(m[1])
        c = b[max_rows * ROW:max_rows * ROW + ns].long()
        per_rank.append(unpack_detections(b[:nr * ROW].view(nr, ROW), c))
    ordered = []
    for group_of_w in zip(*per_rank):
        ordered.extend(group_of_w)
    return ordered[:n_total]
