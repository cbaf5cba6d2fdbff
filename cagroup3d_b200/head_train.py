"""CAGroup3DHead in TRAINING mode, first part (cagroup_head.py:200-225 under model.train(), and the two loss terms that
depend on it alone: semantic focal loss and vote loss, cagroup_head.py:505-517).  SURVEY.md 8f rank 1.

`shared_part` is the differentiable counterpart of the first half of head.CAGroup3DHead.class_maps: semantic logits
(1x1 conv + bias), the offset block (conv-BN-ELU x 2 + conv) and the offset features (3^3 conv-BN-ELU), on the same
module / parameters, through the autograd bricks of autograd.py with batch-statistics BatchNorm.

`semantic_and_vote_loss` evaluates the two terms per sample and averages them over the batch like CAGroup3DHead.loss
(cagroup_head.py:374-398).  Together with backbone_train.run_train this is a complete, if partial, training step: it
trains the backbone, the semantic branch and the vote branch (the per-class grouping branch -- centerness / box / class
terms -- is not differentiable on the CUDA path yet).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import autograd as A
from . import backbone_train as BT
from . import sparse as S
from . import train_targets as TT


def shared_part(head, out: S.SparseTensor, impl: Optional[str] = None):
    """-> (semantic logits (N, n_classes), vote offsets (N, 3 nv), offset features (N, C nv)), all with autograd graphs."""
    sem = A.add_bias(A.conv(out, head.semantic_conv.kernel, 1, 1, impl=impl).F, head.semantic_conv.bias)
    ob = head.offset_block
    h = BT.conv_bn(out, ob[0], ob[1], "elu", impl=impl)
    h = BT.conv_bn(h, ob[3], ob[4], "elu", impl=impl)
    offs = A.conv(h, ob[6].kernel, 1, 1, impl=impl).F
    offF = BT.conv_bn(out, head.feature_offset[0], head.feature_offset[1], "elu", impl=impl).F
    head.fold.clear()                                   # running statistics changed
    return sem, offs, offF


def semantic_and_vote_loss(head, out: S.SparseTensor, sem: torch.Tensor, offs: torch.Tensor, batch_size: int,
                           gt_bboxes: Sequence[torch.Tensor], gt_labels: Sequence[torch.Tensor],
                           scene_points: Sequence[torch.Tensor], pts_semantic_mask: Sequence[torch.Tensor],
                           pts_instance_mask: Sequence[torch.Tensor]):
    """(loss_sem, loss_vote) averaged over the samples of the batch (WITH_YAW False)."""
    assert not head.with_yaw, "the SUN RGB-D vote targets (3 votes per seed) are not on the CUDA training path yet"
    from .dist import reduce_mean
    focal = TT.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
    smooth = TT.SmoothL1Loss(beta=0.04, reduction="sum", loss_weight=1.0)
    C = out.C
    sems, votes = [], []
    for b in range(batch_size):
        rows = torch.nonzero(C[:, 0] == b).squeeze(1)
        vox = C[rows, 1:].float() * head.voxel_size
        with torch.no_grad():
            sem_labels, _ = TT.CAGroup3DAssigner.assign_semantic(vox, gt_bboxes[b], gt_labels[b], head.n_classes)
            off_t, off_m = TT.vote_targets(scene_points[b], vox, gt_bboxes[b], pts_semantic_mask[b], pts_instance_mask[b],
                                           head.n_classes)
        w = (off_m / torch.ones_like(off_m).sum() + 1e-6).unsqueeze(1).repeat(1, 3)
        votes.append(smooth(offs[rows], off_t, weight=w))
        n_pos = max(float(reduce_mean((sem_labels >= 0).sum().float())), 1.)
        sems.append(focal(sem[rows], sem_labels, avg_factor=n_pos))
    return torch.mean(torch.stack(sems)), torch.mean(torch.stack(votes))
