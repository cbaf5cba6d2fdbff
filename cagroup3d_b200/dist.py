"""Multi-GPU plumbing of the inference path: scenes shard by index, one process per GPU, and ONE collective
-- the gather of the per-scene detections at the end.

Replaces, with the same ordering semantics,
  * pcdet/datasets/__init__.py:28-48  DistributedSampler(shuffle=False): index list padded by wrap-around to a
    multiple of world_size, rank r takes indices r, r + W, r + 2W, ...
  * pcdet/utils/common_utils.py:202-223  merge_results_dist: every rank pickles its list to a shared tmpdir,
    two barriers, rank 0 re-interleaves (zip over ranks) and truncates to the dataset size.
Here the detections travel as fixed-width rows [box(7) | score | label] in ONE fixed-capacity all_gather_into_tensor (NCCL
over NVLink on the GPU box, gloo in the CPU tests) that carries its own sizes; no files, no pickling, every rank gets the result.

Training side (SURVEY.md 8e "Training", config 4): the two collectives of the reference's DDP step --
  * tools/train.py:144  nn.parallel.DistributedDataParallel(model): gradients averaged over ranks after backward.
    Here `GradientAllReducer`: parameters are laid out once into a few large flat fp32 buckets (reverse registration
    order, i.e. roughly the order backward produces them), `.grad` of every parameter is a VIEW into its bucket, so a
    step is one in-place all_reduce per bucket with no pack / unpack copies; the buckets are sized for NVSwitch
    (launch latency, not link count: 64 MB default -> 8 collectives for the 506 MB of CAGroup3D gradients), each
    launched from the autograd hook of the bucket's last gradient, so the collectives overlap the rest of backward, and
    awaited together in reduce().
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


def detections_to_host(pred_dicts: Sequence[dict]) -> List[dict]:
    """the per-scene detections as CPU tensors through ONE device -> host copy of the packed rows (instead of three
    synchronising copies per scene)."""
    rows, counts = pack_detections(pred_dicts)
    return unpack_detections(rows.cpu(), torch.tensor([len(d["pred_boxes"]) for d in pred_dicts], dtype=torch.int64))


_GATHER_CAP = {"rows": 4096}        # rows per rank of the fixed-capacity buffer; grows (on every rank alike) after an overflow


def gather_detections(pred_dicts: Sequence[dict], n_total: int, group=None, to_host: bool = False) -> List[dict]:
    """all ranks -> the detections of scenes 0..n_total-1 in dataset order (merge_results_dist semantics:
    interleave the ranks' lists, drop the wrap-around padding).  Boxes come back 7 wide, labels int64, in every launch
    mode (single process included).

    ONE collective: every rank sends a fixed-capacity buffer [n_rows, n_scenes, counts[S] (int32 bit patterns) | rows
    (cap x 9 fp32)] through all_gather_into_tensor; the sizes travel inside it, so there is no size exchange with a host
    read-back in front of the payload.  The result is read once on the host (the per-scene lists need the counts
    anyway).  A rank with more rows than the capacity reports its true count; then every rank sees the overflow in the
    gathered headers, the capacity doubles everywhere and the collective is repeated (rare: 4096 rows per step)."""
    rows, counts = pack_detections(pred_dicts)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = unpack_detections(rows.cpu() if to_host else rows, counts.cpu())
        return out[:n_total]
    world = dist.get_world_size(group)
    dev = rows.device
    S = counts.shape[0]
    while True:
        cap = _GATHER_CAP["rows"]
        hdr = 2 + S
        buf = torch.zeros((hdr + cap * ROW,), dtype=torch.float32, device=dev)
        h = buf[:hdr].view(torch.int32)
        h[0], h[1] = rows.shape[0], S
        h[2:] = counts.to(torch.int32)
        n_send = min(rows.shape[0], cap)
        buf[hdr:hdr + n_send * ROW] = rows[:n_send].reshape(-1)
        flat = torch.empty((world * (hdr + cap * ROW),), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(flat, buf, group=group)
        allb = flat.view(world, hdr + cap * ROW)
        host = allb.cpu()                                                    # the one host read of the step
        heads = host[:, :hdr].contiguous().view(torch.int32)
        if int(heads[:, 0].max()) <= cap:
            break
        while _GATHER_CAP["rows"] < int(heads[:, 0].max()):
            _GATHER_CAP["rows"] *= 2
    src = host if to_host else allb
    per_rank = []
    for r in range(world):
        nr, ns = int(heads[r, 0]), int(heads[r, 1])
        per_rank.append(unpack_detections(src[r, hdr:hdr + nr * ROW].view(nr, ROW), heads[r, 2:2 + ns].long()))
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

    def __init__(self, params, bucket_mb: float = 64.0, group=None, overlap: bool = True):
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
        # Overlap (DDP's behaviour, tools/train.py:144): every parameter's post-accumulate hook counts its bucket down; the
        # hook of a bucket's LAST gradient launches that bucket's all-reduce at once, while autograd is still producing the
        # earlier layers' gradients.  reduce() launches whatever is left (parameters without a gradient this step) and
        # waits.  Without an initialised process group the hooks only count.
        self.overlap = overlap and hasattr(torch.Tensor, "register_post_accumulate_grad_hook")
        self._bucket_size = [len(ps) for ps in plan]
        self._pending = list(self._bucket_size)
        self._works = {}
        self.launched_in_backward = 0
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, bi: int):
        b = self.buckets[bi]
        b.div_(dist.get_world_size(self.group))
        self._works[bi] = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p):
        bi, _, _ = self._slot_of[id(p)]
        if p.grad is None or p.grad.data_ptr() != self._view(p).data_ptr():
            return                            # detached gradient: reduce() copies it back in and reduces the bucket then
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and bi not in self._works and self._distributed():
            self._launch(bi)
            self.launched_in_backward += 1

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
        self._pending = list(self._bucket_size)
        self._works = {}

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)

    def reduce(self):
        """in-place average of all buckets over the ranks; returns the number of collectives issued (in the hooks during
        backward + here).  Call zero_grad() before the next backward."""
        dirty = set()
        for p in self.params:                # a gradient that no longer lives in its bucket is copied back in
            v = self._view(p)
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
                dirty.add(self._slot_of[id(p)][0])
        if not self._distributed():
            self._pending = list(self._bucket_size)
            return 0
        assert not (dirty & set(self._works)), "a bucket was reduced before a detached gradient was copied back"
        for bi in range(len(self.buckets)):
            if bi not in self._works:
                self._launch(bi)
        n = len(self._works)
        for w in self._works.values():
            w.wait()
        self._works = {}
        self._pending = list(self._bucket_size)
        return n


def reduce_mean_many(values, group=None):
    """mean over the ranks of a LIST of scalars / small tensors through ONE all_reduce (the reference issues one
    reduce_mean per loss normaliser: 3 per class branch, cagroup_head.py:523,530,538); -> list of tensors."""
    vals = [v if torch.is_tensor(v) else torch.tensor(float(v)) for v in values]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not vals:
        return vals
    dev = vals[0].device
    flat = torch.cat([v.reshape(-1).float().to(dev) for v in vals])
    dist.all_reduce(flat.div_(dist.get_world_size(group)), op=dist.ReduceOp.SUM, group=group)
    out, o = [], 0
    for v in vals:
        n = v.numel()
        out.append(flat[o:o + n].reshape(v.shape))
        o += n
    return out
