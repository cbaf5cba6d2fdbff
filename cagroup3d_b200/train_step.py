"""Training steps of what the CUDA path can differentiate so far (BASELINE config 4):
tools/train_utils/train_utils.py:42-75 (train_one_epoch's inner loop: forward, loss, backward, optimizer step), with
the DDP gradient average done by dist.GradientAllReducer.

  training_step               the reference's whole step: both stages' losses (`loss_all`), ScanNet and SUN RGB-D (WITH_YAW:
                              3 votes per seed, yaw code + rotated IoU loss in the first stage; (cos, sin) residual code and
                              IoU loss on the decoded RoIs in the second)
  first_stage_training_step   BiResNet in training mode + the whole CAGroup3DHead in training mode + all five terms of
                              CAGroup3DHead.loss (`one_stage_loss` of the reference's tb_dict)
  partial_training_step       the same restricted to the semantic and vote terms (no per-class branch)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import backbone_train as BT
from . import head_train as HT
from .detector import voxelize


def _targets_of(batch_dict: dict, pts: torch.Tensor, B: int):
    gtb, gtl, scene, semm, insm = [], [], [], [], []
    for b in range(B):
        g = batch_dict["gt_boxes"][b]
        g = g[~(g == 0.).all(1)]                                           # zero padding rows (cagroup_head.py:305-308)
        gtb.append(g[:, :7].float().contiguous())
        gtl.append(g[:, 7].long())
        scene.append(pts[pts[:, 0] == b][:, 1:4].contiguous())
        has_masks = "semantic_mask" in batch_dict and "instance_mask" in batch_dict      # SUN RGB-D has none (sunrgbd_dataset.py)
        semm.append(torch.as_tensor(batch_dict["semantic_mask"][b], device=pts.device).long() if has_masks else None)
        insm.append(torch.as_tensor(batch_dict["instance_mask"][b], device=pts.device).long() if has_masks else None)
    return gtb, gtl, scene, semm, insm


def first_stage_training_step(model, batch_dict: dict, optimizer: Optional[torch.optim.Optimizer] = None, reducer=None,
                              impl: Optional[str] = None) -> dict:
    """batch_dict as for partial_training_step plus `cur_epoch` (the semantic threshold schedule, cagroup3d.py:29-31).
    Returns the reference's first-stage tb_dict (loss_centerness, loss_bbox, loss_cls, loss_sem, loss_vote, one_stage_loss)."""
    B = batch_dict["batch_size"]
    pts = batch_dict["points"]
    pts[:, -3:] = pts[:, -3:] / 255.
    head = model.dense_head
    head.semantic_threshold = max(model.semantic_value - int(batch_dict["cur_epoch"]) * model.semantic_iter_value,
                                  model.semantic_min_threshold)
    if reducer is not None:
        reducer.zero_grad()
    elif optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    out = BT.run_train(model.backbone_3d, voxelize(pts, model.voxel_size), impl=impl)
    loss, tb = HT.first_stage_loss(head, out, B, *_targets_of(batch_dict, pts, B), impl=impl)
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    if optimizer is not None:
        optimizer.step()
    if hasattr(model, "update_global_step"):
        model.update_global_step()
    return tb


def two_stage_loss(model, batch_dict: dict, impl: Optional[str] = None, dropout: bool = True):
    """CAGroup3D.get_training_loss (cagroup3d.py:99-158): first-stage loss + RoI-stage loss of a batch.
    -> (loss_all, tb_dict with the reference's keys).  batch_dict["points"] must already be normalised / on the device."""
    from . import roi_train as RT
    B = batch_dict["batch_size"]
    pts = batch_dict["points"]
    head = model.dense_head
    out = BT.run_train(model.backbone_3d, voxelize(pts, model.voxel_size), impl=impl)
    gtb, gtl, scene, semm, insm = _targets_of(batch_dict, pts, B)
    loss1, tb, br = HT.first_stage_loss(head, out, B, gtb, gtl, scene, semm, insm, impl=impl, return_branch=True)
    if model.roi_head is None:
        return loss1, {"loss_all": float(loss1.detach()), **tb}
    proposals = HT.stage1_proposals(head, br, B)
    loss2, tb2, _ = RT.roi_stage_loss(model.roi_head, out, proposals, gtb, gtl, impl=impl, dropout=dropout,
                                      cfg=model.model_cfg.get("ROI_HEAD", None))
    tb.update(tb2)
    loss = loss1 + loss2
    return loss, {"loss_all": float(loss.detach()), **tb}


def training_step(model, batch_dict: dict, optimizer: Optional[torch.optim.Optimizer] = None, reducer=None,
                  impl: Optional[str] = None, grad_norm_clip: Optional[float] = None) -> dict:
    """train_utils.py:48-72 for one batch: zero_grad, forward + both losses, backward, gradient average over the ranks,
    clip_grad_norm_ (OPTIMIZATION.GRAD_NORM_CLIP), optimizer step.  Returns the reference's tb_dict."""
    pts = batch_dict["points"]
    pts[:, -3:] = pts[:, -3:] / 255.
    model.dense_head.semantic_threshold = max(model.semantic_value - int(batch_dict["cur_epoch"]) * model.semantic_iter_value,
                                              model.semantic_min_threshold)
    if reducer is not None:
        reducer.zero_grad()
    elif optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    loss, tb = two_stage_loss(model, batch_dict, impl=impl)
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    if grad_norm_clip:
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], grad_norm_clip)
    if optimizer is not None:
        optimizer.step()
    if hasattr(model, "update_global_step"):
        model.update_global_step()
    return tb


def partial_training_step(model, batch_dict: dict, optimizer: Optional[torch.optim.Optimizer] = None, reducer=None,
                          impl: Optional[str] = None) -> dict:
    """batch_dict as the reference's training loader builds it (cagroup3d.py:27-40, scannet_dataset.py:68-76):
    `points` (N, 7) [b, x, y, z, r, g, b] with colours 0..255 (divided by 255 IN PLACE, like the reference), `batch_size`,
    `gt_boxes` (B, M, 8) zero-padded, class id in the last column, `semantic_mask` / `instance_mask` lists of
    per-point int64 arrays.  reducer: dist.GradientAllReducer (or None in a single process).  Returns the tb_dict."""
    B = batch_dict["batch_size"]
    pts = batch_dict["points"]
    pts[:, -3:] = pts[:, -3:] / 255.
    head = model.dense_head
    if reducer is not None:
        reducer.zero_grad()
    elif optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    x = voxelize(pts, model.voxel_size)
    out = BT.run_train(model.backbone_3d, x, impl=impl)
    sem, offs, _ = HT.shared_part(head, out, impl=impl)
    loss_sem, loss_vote = HT.semantic_and_vote_loss(head, out, sem, offs, B, *_targets_of(batch_dict, pts, B))
    loss = loss_sem + loss_vote
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    if optimizer is not None:
        optimizer.step()
    if hasattr(model, "update_global_step"):
        model.update_global_step()
    return {"loss": float(loss.detach()), "loss_sem": float(loss_sem.detach()), "loss_vote": float(loss_vote.detach())}
