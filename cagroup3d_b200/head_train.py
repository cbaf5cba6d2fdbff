"""CAGroup3DHead in TRAINING mode, first part (cagroup_head.py:200-225 under model.train(), and the two loss terms that
depend on it alone: semantic focal loss and vote loss, cagroup_head.py:505-517).  SURVEY.md 8f rank 1.

`shared_part` is the differentiable counterpart of the first half of head.CAGroup3DHead.class_maps: semantic logits
(1x1 conv + bias), the offset block (conv-BN-ELU x 2 + conv) and the offset features (3^3 conv-BN-ELU), on the same
module / parameters, through the autograd bricks of autograd.py with batch-statistics BatchNorm.

`semantic_and_vote_loss` evaluates the two terms per sample and averages them over the batch like CAGroup3DHead.loss
(cagroup_head.py:374-398).  Together with backbone_train.run_train this is a complete, if partial, training step: it
trains the backbone, the semantic branch and the vote branch; `first_stage_loss` below adds the per-class grouping branch
(centerness / box / class terms) for both the ScanNet and the SUN RGB-D (WITH_YAW) configuration.
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
    """(loss_sem, loss_vote) averaged over the samples of the batch (both branches: per-point-mask votes for ScanNet, up to
    three in-box votes per voxel for WITH_YAW, cagroup_head.py:418-451,505-517)."""
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
            if head.with_yaw:
                off_t, off_m = TT.vote_targets_yaw(vox, gt_bboxes[b], 3)
            else:
                off_t, off_m = TT.vote_targets(scene_points[b], vox, gt_bboxes[b], pts_semantic_mask[b], pts_instance_mask[b],
                                               head.n_classes)
        if head.with_yaw:
            w = (off_m.float() / (off_m.float().sum() + 1e-6)).unsqueeze(1).repeat(1, 9)
            base = vox.repeat(1, 3)
            votes.append(smooth(base + offs[rows], base + off_t, weight=w))
        else:
            w = (off_m / torch.ones_like(off_m).sum() + 1e-6).unsqueeze(1).repeat(1, 3)
            votes.append(smooth(offs[rows], off_t, weight=w))
        n_pos = max(float(reduce_mean((sem_labels >= 0).sum().float())), 1.)
        sems.append(focal(sem[rows], sem_labels, avg_factor=n_pos))
    return torch.mean(torch.stack(sems)), torch.mean(torch.stack(votes))


# ---- the per-class grouping branch (cagroup_head.py:227-282, 627-652) in training mode -------------------------------------
def coordinate_phase(head, out: S.SparseTensor, sem: torch.Tensor, offs: torch.Tensor, B: int) -> dict:
    """Everything of head.CAGroup3DHead.class_maps that depends on coordinates and on the (detached) semantic scores /
    vote offsets only: threshold selection, voted points, class voxels at both sizes with their point -> voxel maps.
    No gradient flows through it (voxel indices are floors).  Same kernels, same class-batched layout (row batch index =
    cls * B + b, rows class-major) as the inference plan."""
    dev = out.F.device
    N, ncls, nv = out.cmap.n, head.n_classes, (3 if head.with_yaw else 1)
    vsA = torch.tensor(head.voxel_size_list, dtype=torch.float32, device=dev)
    vsE = (vsA * head.expand).contiguous()
    semd, offd = sem.detach().contiguous(), offs.detach().contiguous()
    i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
    pad_rows = i32(B)
    S._call("cg3d_first_rows", out.C, N, B, pad_rows)
    mm = i32(6)
    S._call("cg3d_coord_bounds", out.C, N, mm)
    voted = torch.empty((N, nv, 3), dtype=torch.float32, device=dev)
    S._call("cg3d_vote_points", out.C, offd, N, nv, float(head.voxel_size), out.cmap.stride, mm, voted)
    flags = i32(ncls * N)
    S._call("cg3d_semantic_flags", semd, N, ncls, float(head.semantic_threshold), flags)
    pos, total = S.exclusive_scan(flags)
    bounds = torch.cat([pos[::N][:ncls], total]).cpu().tolist()
    sel_rows = i32(max(bounds[-1], 1))
    S._call("cg3d_compact_rows", flags, pos, N, ncls, sel_rows)
    fused = [0]
    for c in range(ncls):
        fused.append(fused[-1] + (nv + 1) * (bounds[c + 1] - bounds[c] + B))
    nf = fused[-1]
    meta = torch.tensor([bounds, fused], dtype=torch.int32).to(dev)
    coordsA, coordsE, ref = i32(nf, 4), i32(nf, 4), i32(nf, 2)
    S._call("cg3d_class_points", out.C, voted, sel_rows, meta[0], meta[1], pad_rows, vsA, vsE, ncls, B, nv, head.expand, nf,
            float(head.voxel_size), coordsA, coordsE, ref)
    mgr = S.Manager(batch_bits=max(1, (ncls * B - 1).bit_length()))
    mapA, _, invA = S.unique_first(coordsA, 1, mgr, want_inverse=True)
    mapE, _, invE = S.unique_first(coordsE, head.expand, mgr, want_inverse=True)
    starts = meta[1][:ncls].long()
    offA = invA[starts].cpu().tolist() + [mapA.n]
    offE = invE[starts].cpu().tolist() + [mapE.n]
    return dict(ref=ref, invA=invA, invE=invE, mapA=mapA, mapE=mapE, offA=offA, offE=offE, mgr=mgr, vsA=vsA)


def _class_bn_elu(F: torch.Tensor, bns, off) -> torch.Tensor:
    """every class has its own BatchNorm: batch statistics over the class's row range, then ELU"""
    parts = [BT._bn(bn, F[off[c]:off[c + 1]]) for c, bn in enumerate(bns)]
    return A.elu(torch.cat(parts))


def class_branch(head, out: S.SparseTensor, offF: torch.Tensor, art: dict, B: int, impl: Optional[str] = None) -> dict:
    """cagroup_head.py:227-282 + forward_single for all classes at once, differentiable.  -> dict(coords (V, 4) int32 with
    batch index cls * B + b, class_off, feat (V, C), centerness (V, 1), cls (V, n_classes), reg (V, n_reg), bbox_pred (V, n_reg)
    = exp(scale_c * reg[:, :6]) | reg[:, 6:])."""
    C, ncls, nv = head.out_channels, head.n_classes, (3 if head.with_yaw else 1)
    mapA, mapE, offA, offE, mgr = art["mapA"], art["mapE"], art["offA"], art["offE"], art["mgr"]
    N = out.cmap.n
    row, kind = art["ref"][:, 0].long(), art["ref"][:, 1].long()
    # point features: the voted copy (kind >= 0) carries its slice of the offset features, the original copy the backbone's
    P = torch.where((kind >= 0).unsqueeze(1), offF.view(N, nv, C)[row, kind.clamp(min=0)], out.F[row])
    FA = A.segment_mean(P, art["invA"], mapA.n)
    FE = A.segment_mean(P, art["invE"], mapE.n)
    st = lambda ts: torch.stack(list(ts))
    nbrE, ordE = S.neighbor_table(mapE, mapE, 5, mgr, ordered=True, group_div=B)
    EF = A.grouped_conv(FE, st(m[0].kernel for m in head.cls_individual_expand_out), nbrE, ordE, mapE.n, 125, offE, offE, impl)
    EF = _class_bn_elu(EF, [m[1] for m in head.cls_individual_expand_out], offE)
    nbrU, ordU = S.transpose_table(mapE, mapA, head.expand, mgr, ordered=True, group_div=B)
    UP = A.grouped_conv(EF, st(m[0].kernel for m in head.cls_individual_up), nbrU, ordU, mapA.n, head.expand ** 3, offA, offE, impl)
    UP = _class_bn_elu(UP, [m[1][0] for m in head.cls_individual_up], offA)
    nbrA, ordA = S.neighbor_table(mapA, mapA, head.cls_kernel, mgr, ordered=True, group_div=B)
    OA = A.grouped_conv(FA, st(m[0].kernel for m in head.cls_individual_out), nbrA, ordA, mapA.n, head.cls_kernel ** 3, offA, offA, impl)
    OA = _class_bn_elu(OA, [m[1] for m in head.cls_individual_out], offA)
    Wf = st(m[0].kernel.unsqueeze(0) for m in head.cls_individual_fuse)                       # [G, 1, 2C, C]
    O = A.grouped_conv(torch.cat([UP, OA], 1), Wf, None, None, mapA.n, 1, offA, offA, impl)
    O = _class_bn_elu(O, [m[1] for m in head.cls_individual_fuse], offA)
    x = S.SparseTensor(O, mapA, mgr)
    ctr = A.conv(x, head.centerness_conv.kernel, 1, 1, impl=impl).F
    cls = A.add_bias(A.conv(x, head.cls_conv.kernel, 1, 1, impl=impl).F, head.cls_conv.bias)
    reg = A.conv(x, head.reg_conv.kernel, 1, 1, impl=impl).F
    # Scale + exp (cagroup_head.py:639-645): row-wise scale of the row's class; glue on (V, 6) values
    scale_rows = torch.cat([head.scales[c].scale.reshape(1).expand(offA[c + 1] - offA[c]) for c in range(ncls)])
    bbox_pred = torch.cat([torch.exp(reg[:, :6] * scale_rows.unsqueeze(1)), reg[:, 6:]], 1)
    head.fold.clear()
    return dict(coords=mapA.coords, class_off=offA, feat=O, centerness=ctr, cls=cls, reg=reg, bbox_pred=bbox_pred)


# ---- the whole first-stage loss (CAGroup3DHead.loss, cagroup_head.py:322-398) -----------------------------------------------
def first_stage_loss(head, out: S.SparseTensor, batch_size: int, gt_bboxes, gt_labels, scene_points, pts_semantic_mask,
                     pts_instance_mask, impl: Optional[str] = None, art: Optional[dict] = None, return_branch: bool = False):
    """shared part -> coordinate phase -> per-class branch -> the five loss terms per sample -> batch means.
    Returns (loss, tb_dict) like the reference (`one_stage_loss` = the sum).  `art`: precomputed coordinate artifacts
    (tests teacher-force them).  WITH_YAW (SUN RGB-D): three votes per voxel, yaw code in the box prediction, rotated IoU
    loss (train_targets.FirstStageLoss(with_yaw=True)); the per-point masks are not used there and may be None."""
    B = batch_size
    sem, offs, offF = shared_part(head, out, impl=impl)
    if art is None:
        with torch.no_grad():
            art = coordinate_phase(head, out, sem, offs, B)
    br = class_branch(head, out, offF, art, B, impl=impl)
    crit = TT.FirstStageLoss(head.n_classes, with_yaw=head.with_yaw, yaw_parametrization=head.yaw_parametrization)
    coords, vs = br["coords"], art["vsA"]
    C = out.C
    # Rows of the class-batched maps carry (class * B + sample) in column 0.  The reference walks samples and classes and
    # picks each (sample, class) block with a boolean mask (B x n_classes selections, each a host sync here); one stable
    # sort by (sample, class) puts every block into one contiguous slice with its rows in their original order, and ONE
    # read-back of the block sizes gives the slice bounds.
    ncls = head.n_classes
    key = coords[:, 0].long()
    blk = (key % B) * ncls + key // B
    perm = torch.sort(blk, stable=True)[1]
    cnt = torch.bincount(blk, minlength=B * ncls)
    vs_rows = vs[key // B]                              # per-class voxel size: a scalar or a per-axis triple
    if vs_rows.dim() == 1:
        vs_rows = vs_rows.unsqueeze(1)
    ctr_s, box_s, cls_s = br["centerness"][perm], br["bbox_pred"][perm], br["cls"][perm]
    pts_s = (coords[:, 1:].float() * vs_rows)[perm]
    Cb = C[:, 0].long()
    vperm = torch.sort(Cb, stable=True)[1]
    bounds = torch.cat([cnt, torch.bincount(Cb, minlength=B)]).cumsum(0).cpu().tolist()
    off = [0] + bounds[:B * ncls]
    voff = [0] + [v - bounds[B * ncls - 1] for v in bounds[B * ncls:]]
    vox_all = C[:, 1:].float() * head.voxel_size
    samples = []
    for b in range(B):
        sl = [slice(off[b * ncls + c], off[b * ncls + c + 1]) for c in range(ncls)]
        rows = vperm[voff[b]:voff[b + 1]]
        vox = vox_all[rows]
        samples.append(dict(centernesses=[ctr_s[q] for q in sl], bbox_preds=[box_s[q] for q in sl], cls_scores=[cls_s[q] for q in sl],
                            points=[pts_s[q] for q in sl], voxel_offset_preds=offs[rows],
                            original_points=vox, semantic_scores=sem[rows], semantic_points=vox, gt_bboxes=gt_bboxes[b],
                            gt_labels=gt_labels[b], scene_points=scene_points[b], pts_semantic_mask=pts_semantic_mask[b],
                            pts_instance_mask=pts_instance_mask[b]))
    terms = crit.loss_batch(samples)                 # one rank-average for all 3 B normalisers
    names = ("loss_centerness", "loss_bbox", "loss_cls", "loss_sem", "loss_vote")
    means = [torch.mean(torch.stack([t[i] for t in terms])) for i in range(5)]
    loss = sum(means)
    tb = {n: float(m.detach()) for n, m in zip(names, means)}
    tb["one_stage_loss"] = float(loss.detach())
    if return_branch:
        return loss, tb, br
    return loss, tb


def stage1_proposals(head, br: dict, B: int):
    """get_bboxes on the training-mode predictions (cagroup_head.py:296-297, detached): the inference plan's device-resident
    decode / top-k / NMS (head.CAGroup3DHead.proposals) on [centerness | cls | reg] -> per-sample (boxes, scores, labels)."""
    with torch.no_grad():
        cm = dict(pred=torch.cat([br["centerness"], br["cls"], br["reg"]], 1).detach().contiguous(), coords=br["coords"])
        det_boxes, det_scores, det_labels, off, _ = head.proposals(cm, B)
    return [(det_boxes[off[b]:off[b + 1]], det_scores[off[b]:off[b + 1]], det_labels[off[b]:off[b + 1]].long()) for b in range(B)]
