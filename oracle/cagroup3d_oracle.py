"""CPU restatement of CAGroup3D eval-mode forward on top of ``me_cpu``.

ORACLE / TEST INFRASTRUCTURE -- never imported by the product path.
"parity unpinned" for the MinkowskiEngine arithmetic (see oracle/__init__.py).

Follows, function by function:
  detector  pcdet/models/detectors/cagroup3d.py:18-50
  backbone  pcdet/models/backbones_3d/biresnet.py:8-406
  head      pcdet/models/dense_heads/cagroup_head.py:200-320, 557-797
  roi head  pcdet/models/roi_heads/cagroup_roi_head.py:14-93, 199-261, 328-510
  coder     pcdet/models/model_utils/cagroup_utils.py:147-197
  nms       pcdet/ops/iou3d_nms (through oracle/iou3d_oracle)

Weights come in as a flat state dict with the reference's parameter names
(SURVEY.md Appendix C).  ``dtype`` selects fp32 (the parity target) or fp64
(a sanity bound on fp32 accumulation error).
"""
from __future__ import annotations

import numpy as np
import torch

from . import me_cpu as me
from . import iou3d_oracle

SCANNET_SIZES = [[0.2309, 0.2435, 0.2777], [0.5631, 0.5528, 0.3579], [0.1840, 0.1845, 0.2155],
                 [0.4187, 0.4536, 0.2503], [0.2938, 0.3203, 0.1899], [0.1595, 0.1787, 0.5250],
                 [0.2887, 0.2174, 0.3445], [0.2497, 0.3147, 0.5063], [0.0634, 0.1262, 0.1612],
                 [0.4332, 0.5691, 0.0810], [0.3088, 0.4212, 0.2627], [0.4130, 0.1966, 0.5044],
                 [0.1995, 0.2133, 0.3897], [0.1260, 0.1137, 0.5254], [0.1781, 0.1774, 0.2218],
                 [0.1526, 0.1520, 0.0904], [0.3453, 0.3164, 0.1491], [0.1426, 0.1477, 0.1741]]
SUNRGBD_SIZES = [[0.6343, 0.4861, 0.2782], [0.2373, 0.3839, 0.2155], [0.2771, 0.5602, 0.2536],
                 [0.1776, 0.1659, 0.2482], [0.2097, 0.1363, 0.2269], [0.2086, 0.4039, 0.2209],
                 [0.1586, 0.3008, 0.3519], [0.1502, 0.1896, 0.2050], [0.1214, 0.3213, 0.5067],
                 [0.2298, 0.4195, 0.1418]]


def class_voxel_sizes(n_classes: int):
    """cagroup_head.py:75-106."""
    sizes = SCANNET_SIZES if n_classes == 18 else SUNRGBD_SIZES
    return np.clip(np.array(sizes) / 2., 0.04, 1.0).tolist()


def default_cfg(n_classes=18, with_yaw=False):
    return dict(n_classes=n_classes, with_yaw=with_yaw, voxel_size=0.02,
                semantic_thr=0.15, semantic_iter=0.02, semantic_min=0.05,
                n_reg_outs=8 if with_yaw else 6, expand=3, cls_kernel=9,
                score_thr=0.01, nms_pre=1000, iou_thr=0.5,
                code_size=7 if with_yaw else 6, sincos=with_yaw, grid=7, coord_key=2,
                roi_kernel=5, test_score_thr=0.01, test_iou_thr=0.5)


class Oracle:
    def __init__(self, params: dict, cfg: dict, dtype=torch.float32):
        self.p = {k: v.detach().cpu() for k, v in params.items()}
        self.cfg = cfg
        self.dt = dtype
        self.train_bn = False          # True: batch statistics in every BatchNorm (the reference's model.train())

    # ---- small helpers ----------------------------------------------------------------
    def W(self, name):
        return self.p[name].to(self.dt)

    def bn(self, F, prefix):
        return me.batchnorm(F, self.p, prefix + ".bn.", train=self.train_bn)

    def conv_bn(self, x, cname, bname, k, stride=1, act=None):
        y = me.conv(x, self.W(cname + ".kernel"), k, stride)
        y.F = self.bn(y.F, bname)
        if act == "relu":
            y.F = torch.relu(y.F)
        elif act == "elu":
            y.F = torch.nn.functional.elu(y.F)
        return y

    # ---- backbone ---------------------------------------------------------------------
    def basic_block(self, x, pre, stride, downsample, no_relu):
        out = self.conv_bn(x, pre + ".conv1", pre + ".norm1", 3, stride, "relu")
        out = self.conv_bn(out, pre + ".conv2", pre + ".norm2", 3, 1)
        res = x
        if downsample:
            res = self.conv_bn(x, pre + ".downsample.0", pre + ".downsample.1", 1, stride)
        out.F = out.F + res.F
        if not no_relu:
            out.F = torch.relu(out.F)
        return out

    def bottleneck(self, x, pre, stride):
        out = self.conv_bn(x, pre + ".conv1", pre + ".norm1", 1, 1, "relu")
        out = self.conv_bn(out, pre + ".conv2", pre + ".norm2", 3, stride, "relu")
        out = self.conv_bn(out, pre + ".conv3", pre + ".norm3", 1, 1)
        res = self.conv_bn(x, pre + ".downsample.0", pre + ".downsample.1", 1, stride)
        out.F = out.F + res.F          # Bottleneck default no_relu=True (biresnet.py:61)
        return out

    def layer(self, x, pre, stride, downsample):
        """_make_layer with 2 BasicBlocks (biresnet.py:336-354)."""
        x = self.basic_block(x, pre + ".0", stride, downsample, no_relu=False)
        return self.basic_block(x, pre + ".1", 1, False, no_relu=True)

    def pre_act_conv(self, x, pre, bn_i, conv_i, k):
        """BN -> ReLU -> conv blocks of DAPPM (biresnet.py:109-174)."""
        F = torch.relu(self.bn(x.F, f"{pre}.{bn_i}"))
        return me.conv(x.with_F(F), self.W(f"{pre}.{conv_i}.kernel"), k)

    def dappm(self, x, pre="backbone_3d.spp"):
        q = x.C
        xs = [self.pre_act_conv(x, pre + ".scale0", 0, 2, 1)]
        for i, (k, s) in enumerate([(5, 2), (9, 4), (17, 8), (33, 16)], start=1):
            pooled = me.avg_pool(x, k, s)
            y = self.pre_act_conv(pooled, f"{pre}.scale{i}", 1, 3, 1)
            up = me.features_at(y, q)
            xs.append(self.pre_act_conv(x.with_F(up + xs[i - 1].F), f"{pre}.process{i}", 0, 2, 3))
        cat = x.with_F(torch.cat([t.F for t in xs], 1))
        out = self.pre_act_conv(cat, pre + ".compression", 0, 2, 1)
        sc = self.pre_act_conv(x, pre + ".shortcut", 0, 2, 1)
        return x.with_F(out.F + sc.F)

    def backbone(self, x):
        b = "backbone_3d."
        relu = lambda t: t.with_F(torch.relu(t.F))
        x = self.conv_bn(x, b + "conv1.0", b + "conv1.1", 3, 1, "relu")
        x = self.conv_bn(x, b + "conv1.3", b + "conv1.4", 3, 1, "relu")
        x = self.layer(x, b + "layer1", 2, True)
        l1 = self.layer(relu(x), b + "layer2", 2, True)                 # stride 4
        l2 = self.layer(relu(l1), b + "layer3", 2, True)                # stride 8
        x_ = self.layer(relu(l1), b + "layer3_", 1, False)              # stride 4
        x = l2.with_F(l2.F + self.conv_bn(relu(x_), b + "down3.0", b + "down3.1", 3, 2).F)
        c3 = self.conv_bn(relu(l2), b + "compression3.0", b + "compression3.1", 1)
        x_ = x_.with_F(x_.F + me.features_at(c3, x_.C))
        l3 = self.layer(relu(x), b + "layer4", 2, True)                 # stride 16
        x_ = self.layer(relu(x_), b + "layer4_", 1, False)
        d = self.conv_bn(relu(x_), b + "down4.0", b + "down4.1", 3, 2, "relu")
        d = self.conv_bn(d, b + "down4.3", b + "down4.4", 3, 2)
        x = l3.with_F(l3.F + d.F)
        c4 = self.conv_bn(relu(l3), b + "compression4.0", b + "compression4.1", 1)
        x_ = x_.with_F(x_.F + me.features_at(c4, x_.C))
        x_ = self.bottleneck(relu(x_), b + "layer5_.0", 1)
        x5 = self.bottleneck(relu(x), b + "layer5.0", 2)                # stride 32
        x_ = x_.with_F(x_.F + me.features_at(self.dappm(x5), x_.C))
        up = me.conv_transpose_k2s2(x_, self.W(b + "out.0.kernel"))     # stride 2
        up.F = torch.relu(self.bn(up.F, b + "out.1"))
        return self.conv_bn(up, b + "out.3", b + "out.4", 1, 1, "relu")

    # ---- dense head -------------------------------------------------------------------
    def head(self, out, sem_thr, batch_size, force=None):
        """force: optional {"sem": (N,ncls), "offsets": (N,3nv)} tensors used IN PLACE of the computed ones at
        the two discontinuities (threshold select, floor into class voxels) -- teacher forcing for tests that
        compare against another fp32 implementation whose last-bit differences would flip a voxel."""
        force = force or {}
        cfg, h = self.cfg, "dense_head."
        vs = cfg["voxel_size"]
        ncls = cfg["n_classes"]
        f32 = torch.float32
        sem = me.conv(out, self.W(h + "semantic_conv.kernel"), 1, 1, self.W(h + "semantic_conv.bias"))
        pad_id = np.array([r[0] for r in out.batch_rows()], dtype=np.int64)
        C = torch.from_numpy(out.C)
        ts = out.cmap.stride
        # scene bounds over the whole batch (cagroup_head.py:209-211), fp32 like the reference
        max_b = ((C[:, 1:].max(0)[0] + ts) * vs).to(f32)
        min_b = ((C[:, 1:].min(0)[0] - ts) * vs).to(f32)
        o = out
        for i, act in ((0, "elu"), (3, "elu")):
            o = self.conv_bn(o, h + f"offset_block.{i}", h + f"offset_block.{i + 1}", 1, 1, act)
        offs = me.conv(o, self.W(h + "offset_block.6.kernel"), 1)
        offF = self.conv_bn(out, h + "feature_offset.0", h + "feature_offset.1", 3, 1, "elu").F
        nv = 3 if cfg["with_yaw"] else 1
        # voted coordinates are formed in fp32 (they get floored into voxel indices)
        base = (C[:, 1:].to(f32) * vs).view(-1, 1, 3)
        voted = base + force.get("offsets", offs.F).to(f32).view(-1, nv, 3)
        voted = torch.maximum(torch.minimum(voted, max_b.view(1, 1, 3)), min_b.view(1, 1, 3))
        offF = offF.view(offF.shape[0], nv, -1)
        sizes = class_voxel_sizes(ncls)
        Cf = C.to(f32)
        per_class, maps = [], []
        for cls in range(ncls):
            s = torch.sigmoid(force.get("sem", sem.F)[:, cls])
            sel = torch.cat([torch.nonzero(s > sem_thr).squeeze(1), torch.from_numpy(pad_id)])
            vc = Cf[sel].view(-1, 1, 4).repeat(1, nv, 1)
            vc[:, :, 1:4] = voted[sel]
            oc = Cf[sel].clone()
            oc[:, 1:4] *= vs
            fuse_c = torch.cat([vc.reshape(-1, 4), oc], 0)
            fuse_f = torch.cat([offF[sel].reshape(-1, offF.shape[-1]), out.F[sel]], 0)
            vsz = torch.tensor(sizes[cls], dtype=f32)
            qa = fuse_c.clone()
            qa[:, 1:] = torch.floor(fuse_c[:, 1:] / vsz)
            A = me.from_points(qa, fuse_f, average=True)
            Aout = self.conv_bn(A, h + f"cls_individual_out.{cls}.0", h + f"cls_individual_out.{cls}.1",
                                cfg["cls_kernel"], 1, "elu")
            ex = cfg["expand"]
            qe = fuse_c.clone()
            qe[:, 1:] = torch.floor(fuse_c[:, 1:] / (vsz * ex)) * ex
            E = me.from_points(qe, fuse_f, average=True, stride=ex)
            E = self.conv_bn(E, h + f"cls_individual_expand_out.{cls}.0",
                             h + f"cls_individual_expand_out.{cls}.1", 5, 1, "elu")
            U = me.generative_transpose_k3s3(E, self.W(h + f"cls_individual_up.{cls}.0.kernel"), A.cmap)
            U = torch.nn.functional.elu(self.bn(U, h + f"cls_individual_up.{cls}.1.0"))
            O = A.with_F(torch.cat([U, Aout.F], 1))
            O = self.conv_bn(O, h + f"cls_individual_fuse.{cls}.0", h + f"cls_individual_fuse.{cls}.1", 1, 1, "elu")
            # forward_single (cagroup_head.py:627-652)
            ctr = O.F @ self.W(h + "centerness_conv.kernel")
            cls_score = O.F @ self.W(h + "cls_conv.kernel") + self.W(h + "cls_conv.bias").reshape(1, -1)
            reg = O.F @ self.W(h + "reg_conv.kernel")
            dist = torch.exp(reg[:, :6] * self.W(h + f"scales.{cls}.scale"))
            bbox = torch.cat([dist, reg[:, 6:]], 1)
            rows = O.batch_rows()
            assert len(rows) == batch_size
            pts = [torch.from_numpy(O.C[r, 1:]).to(f32) * vsz for r in rows]
            per_class.append(([ctr[r] for r in rows], [bbox[r] for r in rows],
                              [cls_score[r] for r in rows], pts))
            maps.append(dict(coords=O.C, ctr=ctr, bbox=bbox, cls=cls_score, reg=reg, feat=O.F,
                             n_sel=int(len(sel))))
        # get_bboxes (cagroup_head.py:557-624)
        results = []
        for b in range(batch_size):
            bbs, scs = [], []
            for cls in range(ncls):
                ctr, bbox, cs, pts = (per_class[cls][j][b] for j in range(4))
                scores = torch.sigmoid(cs) * torch.sigmoid(ctr)
                if len(scores) > cfg["nms_pre"] > 0:
                    mx = scores.max(1)[0]
                    ids = torch.sort(mx, descending=True, stable=True)[1][:cfg["nms_pre"]]
                    bbox, scores, pts = bbox[ids], scores[ids], pts[ids]
                bbs.append(bbox_pred_to_bbox(pts.to(self.dt), bbox))
                scs.append(scores)
            results.append(stage1_nms(torch.cat(bbs), torch.cat(scs), cfg))
        inter = dict(sem=sem.F, offsets=offs.F, offset_feat=offF, voted=voted, maps=maps)
        return results, inter

    # ---- roi head ---------------------------------------------------------------------
    def roi_head(self, out, pred_list, batch_size):
        cfg, r = self.cfg, "roi_head."
        f32 = torch.float32
        nmax = max(1, max(len(p[0]) for p in pred_list))
        rois = torch.zeros((batch_size, nmax, 7), dtype=self.dt)
        roi_scores = torch.zeros((batch_size, nmax), dtype=self.dt)
        roi_labels = torch.zeros((batch_size, nmax), dtype=torch.long)
        for b, (bx, sc, lb) in enumerate(pred_list):
            rois[b, :len(bx)] = bx
            roi_scores[b, :len(bx)] = sc
            roi_labels[b, :len(bx)] = lb
        rois[..., 6] *= -1
        # 7^3 grid points per RoI (cagroup_roi_head.py:199-224)
        g = cfg["grid"]
        flat = rois.view(-1, 7).to(f32)
        idx = torch.nonzero(torch.ones(g, g, g)).to(f32)
        size = flat[:, 3:6].unsqueeze(1)
        local = (idx.unsqueeze(0) + 0.5) / g * size - size / 2
        if cfg["code_size"] > 6:
            local = rotate_z(local, flat[:, 6])
        pts = (local + flat[:, 0:3].unsqueeze(1)).reshape(batch_size, -1, 3)
        bidx = torch.arange(batch_size, dtype=f32).view(-1, 1, 1).expand(-1, pts.shape[1], 1)
        gp = torch.cat([bidx, pts], -1).reshape(-1, 4)
        # SimplePoolingLayer (cagroup_roi_head.py:46-93)
        vsz = cfg["voxel_size"] * cfg["coord_key"]
        gc = gp.long()
        gc[:, 1:4] = torch.floor(gp[:, 1:4] / vsz).long()
        half = 768 // 2
        gc[:, 1:4] = torch.clamp(gc[:, 1:4], min=-half + 1, max=half - 1)
        uq, inv, _ = me.unique_first(gc.numpy())
        uq = uq.copy()
        uq[:, 1:] *= cfg["coord_key"]
        pre = r + "roi_grid_pool_layers.0."
        y = me.conv_at(out, self.W(pre + "grid_conv.kernel"), cfg["roi_kernel"], uq)
        yF = torch.nn.functional.elu(self.bn(y.F, pre + "grid_bn"))
        feat = yF[torch.from_numpy(inv)].view(-1, g * g * g, yF.shape[1])      # (B*R, 343, 128)
        # pooling conv == dense contraction over a permuted tap axis (A20)
        Wp = self.W(pre + "pooling_conv.kernel")                               # (343,128,128)
        ii, jj, kk = np.meshgrid(np.arange(g), np.arange(g), np.arange(g), indexing="ij")
        tap = torch.from_numpy((ii + g * jj + g * g * kk).ravel())              # grid point i*49+j*7+k -> tap
        pooled = torch.einsum("rgc,gcd->rd", feat, Wp[tap])
        pooled = self.bn(pooled, pre + "pooling_bn")
        # reg MLP (cagroup_roi_head.py:168-184); Dropout is identity in eval
        x = pooled
        for li, bi in ((0, 1), (4, 5)):
            x = x @ self.W(r + f"reg_fc_layers.{li}.weight").t()
            x = torch.relu(me.batchnorm(x, self.p, r + f"reg_fc_layers.{bi}.", train=self.train_bn))
        reg = x @ self.W(r + "reg_pred_layer.weight").t() + self.W(r + "reg_pred_layer.bias")
        # decode (cagroup_roi_head.py:477-510)
        cs = cfg["code_size"]
        enc = reg.view(batch_size, -1, cs + (1 if cfg["sincos"] else 0))
        local_rois = rois.clone()[..., :cs]
        local_rois[..., 0:3] = 0
        dec = residual_decode(enc, local_rois, cs, cfg["sincos"]).view(-1, cs)
        if cs > 6:
            dec = torch.cat([rotate_z(dec[:, None, 0:3], rois[..., 6].reshape(-1))[:, 0], dec[:, 3:]], 1)
        dec[:, 0:3] += rois[..., 0:3].reshape(-1, 3)
        dec = dec.view(batch_size, -1, cs)
        final = [roi_nms(dec[b], roi_scores[b], roi_labels[b], cfg) for b in range(batch_size)]
        inter = dict(rois=rois, roi_scores=roi_scores, roi_labels=roi_labels, grid_coords=gc.numpy(),
                     uniq=uq, pooled=pooled, rcnn_reg=reg, decoded=dec)
        return final, inter

    # ---- detector ---------------------------------------------------------------------
    def forward(self, points: torch.Tensor, batch_size: int, cur_epoch: int = 10, stages="all", force=None):
        """points: (N,7) fp32 [b,x,y,z,r,g,b] with colours in 0..255 (cagroup3d.py:27-50)."""
        cfg = self.cfg
        thr = max(cfg["semantic_thr"] - int(cur_epoch) * cfg["semantic_iter"], cfg["semantic_min"])
        pts = points.detach().cpu().to(torch.float32).clone()
        pts[:, -3:] = pts[:, -3:] / 255.
        coords = pts[:, :4].clone()
        coords[:, 1:] /= cfg["voxel_size"]
        x = me.from_points(coords, pts[:, 4:].to(self.dt))
        res = dict(vox_coords=x.C, vox_feats=x.F)
        if stages == "voxelize":
            return res
        out = self.backbone(x)
        res.update(bb_coords=out.C, bb_feats=out.F, maps={s: m.coords for s, m in x.mgr.by_stride.items()})
        if stages == "backbone":
            return res
        pred_list, hi = self.head(out, thr, batch_size, force)
        res.update(head=hi, stage1=pred_list)
        if stages == "head":
            return res
        final, ri = self.roi_head(out, pred_list, batch_size)
        res.update(roi=ri, final=final)
        return res


# ---- box helpers -------------------------------------------------------------------------
def rotate_z(points, angle):
    """common_utils.rotate_points_along_z (pcdet/utils/common_utils.py:35-57)."""
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot = torch.stack((c, s, z, -s, c, z, z, z, o), 1).view(-1, 3, 3).to(points.dtype)
    return torch.matmul(points[:, :, 0:3], rot)


def bbox_pred_to_bbox(points, bbox_pred):
    """cagroup_head.py:654-703 ('fcaf3d' yaw parametrisation for 8 outputs)."""
    if bbox_pred.shape[0] == 0:
        return bbox_pred
    xc = points[:, 0] + (bbox_pred[:, 1] - bbox_pred[:, 0]) / 2
    yc = points[:, 1] + (bbox_pred[:, 3] - bbox_pred[:, 2]) / 2
    zc = points[:, 2] + (bbox_pred[:, 5] - bbox_pred[:, 4]) / 2
    if bbox_pred.shape[1] == 6:
        return torch.stack([xc, yc, zc, bbox_pred[:, 0] + bbox_pred[:, 1],
                            bbox_pred[:, 2] + bbox_pred[:, 3], bbox_pred[:, 4] + bbox_pred[:, 5]], -1)
    scale = bbox_pred[:, 0] + bbox_pred[:, 1] + bbox_pred[:, 2] + bbox_pred[:, 3]
    q = torch.exp(torch.sqrt(torch.pow(bbox_pred[:, 6], 2) + torch.pow(bbox_pred[:, 7], 2)))
    alpha = 0.5 * torch.atan2(bbox_pred[:, 6], bbox_pred[:, 7])
    return torch.stack((xc, yc, zc, scale / (1 + q), scale / (1 + q) * q,
                        bbox_pred[:, 5] + bbox_pred[:, 4], alpha), dim=-1)


def residual_decode(enc, anchors, code_size, sincos):
    """CAGroupResidualCoder.decode_torch (cagroup_utils.py:147-197)."""
    if code_size > 6:
        xa, ya, za, dxa, dya, dza, ra = torch.split(anchors, 1, dim=-1)
        if sincos:
            xt, yt, zt, dxt, dyt, dzt, cost, sint = torch.split(enc, 1, dim=-1)
        else:
            xt, yt, zt, dxt, dyt, dzt, rt = torch.split(enc, 1, dim=-1)
    else:
        xa, ya, za, dxa, dya, dza = torch.split(anchors, 1, dim=-1)
        xt, yt, zt, dxt, dyt, dzt = torch.split(enc, 1, dim=-1)
    diag = torch.sqrt(dxa ** 2 + dya ** 2)
    out = [xt * diag + xa, yt * diag + ya, zt * dza + za,
           torch.exp(dxt) * dxa, torch.exp(dyt) * dya, torch.exp(dzt) * dza]
    if code_size > 6:
        out.append((torch.atan2(sint, cost) if sincos else rt) + ra)
    return torch.cat(out, -1)


def stage1_nms(bboxes, scores, cfg):
    """CAGroup3DHead._nms (cagroup_head.py:747-797)."""
    yaw = bboxes.shape[1] == 7
    ob, os_, ol = [], [], []
    for i in range(scores.shape[1]):
        ids = scores[:, i] > cfg["score_thr"]
        if not ids.any():
            continue
        cs, cb = scores[ids, i], bboxes[ids]
        if not yaw:
            cb = torch.cat((cb, torch.zeros_like(cb[:, :1])), 1)
        corr = cb.clone()
        if yaw:
            corr[:, 6] *= -1
        keep = iou3d_oracle.nms(corr, cs, cfg["iou_thr"], rotated=yaw)
        ob.append(cb[keep]); os_.append(cs[keep]); ol.append(torch.full((len(keep),), i, dtype=torch.long))
    if ob:
        ob, os_, ol = torch.cat(ob), torch.cat(os_), torch.cat(ol)
    else:
        ob, os_, ol = bboxes.new_zeros((0, 7)), bboxes.new_zeros((0,)), torch.zeros((0,), dtype=torch.long)
    if not yaw:
        ob = torch.cat([ob[:, :6], ob.new_zeros(ob.shape[0], 1)], 1)
    return ob, os_, ol


def roi_nms(bboxes, scores, labels, cfg):
    """CAGroup3DRoIHead._nms (cagroup_roi_head.py:433-475)."""
    yaw = bboxes.shape[1] == 7
    nz = bool(bboxes.sum() != 0)
    ob, os_, ol = [], [], []
    for i in range(cfg["n_classes"]):
        ids = (labels == i) & (scores > cfg["test_score_thr"]) & nz
        if not ids.any():
            continue
        cs, cb = scores[ids], bboxes[ids]
        if not yaw:
            cb = torch.cat((cb, torch.zeros_like(cb[:, :1])), 1)
        keep = iou3d_oracle.nms(cb, cs, cfg["test_iou_thr"], rotated=yaw)
        ob.append(cb[keep]); os_.append(cs[keep]); ol.append(torch.full((len(keep),), i, dtype=torch.long))
    if ob:
        ob, os_, ol = torch.cat(ob), torch.cat(os_), torch.cat(ol)
    else:
        ob, os_, ol = bboxes.new_zeros((0, 7)), bboxes.new_zeros((0,)), torch.zeros((0,), dtype=torch.long)
    if yaw:
        ob = ob.clone()
        ob[:, 6] *= -1
    else:
        ob = torch.cat([ob[:, :6], ob.new_zeros(ob.shape[0], 1)], 1)
    return ob, os_, ol
