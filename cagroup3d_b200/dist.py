"""Multi-GPU plumbing of the inference path: scenes shard by index, one process per GPU, and ONE collective
-- the gather of the per-scene detections at the end.

Replaces, with the same ordering semantics,
  * pcdet/datasets/__init__.py:28-48  DistributedSampler(shuffle=False): index list padded by wrap-around to a
    multiple of world_size, rank r takes indices r, r + W, r + 2W, ...
  * pcdet/utils/common_utils.py:202-223  merge_results_dist: every rank pickles its list to a shared tmpdir,
    two barriers, rank 0 re-interleaves (zip over ranks) and truncates to the dataset size.
Here the detections travel as fixed-width rows [box(7) | score | label] in one padded all_gather (NCCL over
NVLink on the GPU box, gloo in the CPU tests); no files, no pickling, every rank gets the result.

Training side (SURVEY.md 8e "Training", config 4): the two collectives of the reference's DDP step --
  * tools/train.py:144  nn.parallel.DistributedDataParallel(model): gradients averaged over ranks after backward.
    Here `GradientAllReducer`: parameters are laid out once into a few large flat fp32 buckets (reverse registration
    order, i.e. roughly the order backward produces them), `.grad` of every parameter is a VIEW into its bucket, so a
    step is one in-place all_reduce per bucket with no pack / unpack copies; the buckets are sized for NVSwitch
    (launch latency, not link count: 64 MB default -> 8 collectives for the 506 MB of CAGroup3D gradients), issued
    asynchronously in backward order and awaited together.
  * pcdet/models/model_utils/cagroup_utils.py:6-12  reduce_mean (avg_factor of the loss terms): `reduce_mean`.
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


# ---- training collectives ---------------------------------------------------------------------------------
def reduce_mean(tensor: torch.Tensor, group=None) -> torch.Tensor:
    """mean of `tensor` over the ranks (cagroup_utils.py:6-12: clone, divide by world size, all_reduce SUM);
    the input itself when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    t = tensor.clone()
    dist.all_reduce(t.div_(dist.get_world_size(group)), op=dist.ReduceOp.SUM, group=group)
    return t


class GradientAllReducer:
    """Averages the gradients of `params` over the ranks, DDP-style, through flat buckets the gradients live in.

    reducer = GradientAllReducer(model.parameters()); ...; loss.backward(); reducer.reduce(); optimizer.step()

    After construction p.grad is a view into a bucket for every parameter (zero-filled), so autograd accumulates
    straight into the bucket; `zero_grad()` clears the buckets and keeps the views (use it instead of
    optimizer.zero_grad(set_to_none=True), which would detach them; `reduce()` re-attaches a detached or replaced
    gradient by copying it in, so that case stays correct, just slower).  Parameters without a gradient in a step
    contribute zeros, as with DDP(find_unused_parameters=False) they must not exist; frozen parameters
    (requires_grad=False) are left out."""

    def __init__(self, params, bucket_mb: float = 64.0, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        cap = max(1, int(bucket_mb * (1 << 20)) // 4)
        self.buckets: List[torch.Tensor] = []
        self.slots = []                      # per parameter: (bucket index, offset, numel)
        cur, cur_n = [], 0
        plan = []
        for p in reversed(self.params):      # backward produces gradients roughly last-registered first
            assert p.dtype == torch.float32, "gradients are reduced in fp32 (BASELINE config 4)"
            if cur and (cur_n + p.numel() > cap or p.device != cur[0].device):
                plan.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            plan.append(cur)
        self._slot_of = {}
        for bi, ps in enumerate(plan):
            n = sum(-(-p.numel() // 4) * 4 for p in ps)          # 16-byte aligned slots
            buf = torch.zeros((n,), dtype=torch.float32, device=ps[0].device)
            self.buckets.append(buf)
            o = 0
            for p in ps:
                self._slot_of[id(p)] = (bi, o, p.numel())
                o += -(-p.numel() // 4) * 4
        self._attach()

    def _view(self, p):
        bi, o, n = self._slot_of[id(p)]
        return self.buckets[bi][o:o + n].view(p.shape)

    def _attach(self):
        for p in self.params:
            p.grad = self._view(p)

    def zero_grad(self):
        for b in self.buckets:
            b.zero_()
        self._attach()

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)

    def reduce(self):
        """in-place average of all buckets over the ranks; returns the number of collectives issued."""
        for p in self.params:                # a gradient that no longer lives in its bucket is copied back in
            v = self._view(p)
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        world = dist.get_world_size(self.group)
        works = []
        for b in self.buckets:
            b.div_(world)
            works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()
        return len(works)
