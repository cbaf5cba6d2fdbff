"""CAGroup3DHead inference path on the CUDA C-ABI ops (mirror of pcdet/models/dense_heads/cagroup_head.py).

The reference loops over the classes in Python (cagroup_head.py:227-282), building two coordinate
managers and running three sparse convs per class.  Here the class index is folded into the batch
index of the re-voxelised coordinates (row batch = cls * B + b), so ONE hash-unique, ONE rule-map
build and ONE grouped-conv launch serve all classes; proposal top-k, score filtering, sorting, NMS
and packing stay on the device (proposal.cu, sort.cu, nms.cu).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from . import _lib
from . import sparse as S
from .backbone import FoldCache, conv_bn
from .me_compat import (MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiELU,
                        MinkowskiGenerativeConvolutionTranspose)

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


def bias_init_with_prob(p: float) -> float:
    """cagroup_utils.py: bias so that sigmoid(bias) == p."""
    return float(-math.log((1 - p) / p))


class Scale(nn.Module):
    """cagroup_utils.py:69-84."""

    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))


def _i32(*shape, device):
    return torch.empty(shape, dtype=torch.int32, device=device)


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _u64(n, device):
    return torch.empty((max(n, 1),), dtype=torch.int64, device=device)


sort_pairs = S.sort_pairs


def seg_bits(nseg: int) -> int:
    """number of key bits a sort on (segment << 32 | score) has to touch."""
    return 32 + max(1, int(nseg - 1).bit_length())


def nms_and_pack(keys, src_row, n, boxes, box_dim, nseg, ncls, B, thr, with_yaw, gather_flip):
    """sort pairs by (segment, score desc) -> greedy NMS per segment -> packed detections.

    Returns (det_boxes [m,7], det_scores [m], det_labels [m] int32, sample_off list[B+1]).
    One host sync (kept count + per-sample counts)."""
    dev = boxes.device
    if n == 0:
        return _f32(0, 7, device=dev), _f32(0, device=dev), _i32(0, device=dev), [0] * (B + 1)
    sort_pairs(keys, src_row, n, seg_bits(nseg))
    sorted_boxes = _f32(n, 7, device=dev)
    S._call("cg3d_gather_boxes", boxes, box_dim, src_row, n, gather_flip, sorted_boxes)
    seg = _i32(n, device=dev)
    S._call("cg3d_key_segments", keys, n, seg)
    counts = _i32(nseg + 1, device=dev)
    counts[nseg:].zero_()
    S._call("cg3d_histogram_i32", seg, n, nseg, counts)
    seg_off, _ = S.exclusive_scan(counts)
    max_len = n          # upper bound of any segment (avoids a sync for the true maximum; it only sizes a bitset)
    keep = _i32(n, device=dev)
    S._call("cg3d_nms_segments", sorted_boxes, n, seg_off, nseg, max_len, float(thr), int(with_yaw), keep, None)
    pos, total = S.exclusive_scan(keep)
    # per-sample kept counts = kept-prefix at the sample boundaries of seg_off
    m = int(total.item())
    det_boxes, det_scores = _f32(m, 7, device=dev), _f32(m, device=dev)
    det_labels, det_sample = _i32(m, device=dev), _i32(m, device=dev)
    if m:
        S._call("cg3d_emit_detections", sorted_boxes, keys, keep, pos, n, ncls, int(with_yaw), 1, det_boxes,
                det_scores, det_labels, det_sample)
        cnt = _i32(B + 1, device=dev)
        cnt[B:].zero_()
        S._call("cg3d_histogram_i32", det_sample, m, B, cnt)
        off, _ = S.exclusive_scan(cnt)
        sample_off = off.cpu().tolist()
    else:
        sample_off = [0] * (B + 1)
    return det_boxes, det_scores, det_labels, sample_off


class CAGroup3DHead(nn.Module):
    def __init__(self, model_cfg, yaw_parametrization="fcaf3d", predict_boxes=True, **kwargs):
        super().__init__()
        g = model_cfg.get
        self.n_classes = g("N_CLASSES")
        out_channels = g("OUT_CHANNELS")
        self.n_reg_outs = g("N_REG_OUTS")
        self.voxel_size = g("VOXEL_SIZE")
        self.semantic_threshold = g("SEMANTIC_THR")
        self.expand = g("EXPAND_RATIO")
        self.with_yaw = g("WITH_YAW")
        self.use_sem_score = g("USE_SEM_SCORE", False)
        self.cls_kernel = g("CLS_KERNEL")
        nms = g("NMS_CONFIG", None) or {}
        self.score_thr = nms.get("SCORE_THR", 0.01)
        self.nms_pre = nms.get("NMS_PRE", 1000)
        self.iou_thr = nms.get("IOU_THR", 0.5)
        if self.use_sem_score:
            raise NotImplementedError("USE_SEM_SCORE=True is not used by the shipped CAGroup3D configs")
        self.yaw_parametrization = yaw_parametrization
        self.predict_boxes = predict_boxes
        sizes = SCANNET_SIZES if self.n_classes == 18 else SUNRGBD_SIZES
        if len(sizes) != self.n_classes:
            sizes = (sizes * self.n_classes)[:self.n_classes]
        self.voxel_size_list = np.clip(np.array(sizes) / 2., 0.04, 1.0).tolist()          # cagroup_head.py:75-106
        self.out_channels = out_channels
        self._init_layers(out_channels, self.n_reg_outs, self.n_classes)
        self.fold = FoldCache()
        self.init_weights()

    # ---- parameters (names as in cagroup_head.py:116-187) -------------------------------------------
    @staticmethod
    def _block(cin, cout, k):
        return nn.Sequential(MinkowskiConvolution(cin, cout, kernel_size=k), MinkowskiBatchNorm(cout), MinkowskiELU())

    def _init_layers(self, c, n_reg_outs, n_classes):
        nv = 3 if self.with_yaw else 1
        self.offset_block = nn.Sequential(
            MinkowskiConvolution(c, c, kernel_size=1), MinkowskiBatchNorm(c), MinkowskiELU(),
            MinkowskiConvolution(c, c, kernel_size=1), MinkowskiBatchNorm(c), MinkowskiELU(),
            MinkowskiConvolution(c, 3 * nv, kernel_size=1))
        self.feature_offset = self._block(c, c * nv, 3)
        self.semantic_conv = MinkowskiConvolution(c, n_classes, kernel_size=1, bias=True)
        self.centerness_conv = MinkowskiConvolution(c, 1, kernel_size=1)
        self.reg_conv = MinkowskiConvolution(c, n_reg_outs, kernel_size=1)
        self.cls_conv = MinkowskiConvolution(c, n_classes, kernel_size=1, bias=True)
        self.scales = nn.ModuleList([Scale(1.) for _ in range(n_classes)])
        self.cls_individual_out = nn.ModuleList([self._block(c, c, self.cls_kernel) for _ in range(n_classes)])
        self.cls_individual_up = nn.ModuleList([nn.ModuleList([
            MinkowskiGenerativeConvolutionTranspose(c, c, kernel_size=self.expand, stride=self.expand),
            nn.Sequential(MinkowskiBatchNorm(c), MinkowskiELU())]) for _ in range(n_classes)])
        self.cls_individual_fuse = nn.ModuleList([self._block(c * 2, c, 1) for _ in range(n_classes)])
        self.cls_individual_expand_out = nn.ModuleList([self._block(c, c, 5) for _ in range(n_classes)])

    def init_weights(self):
        """cagroup_head.py:190-198."""
        nn.init.normal_(self.centerness_conv.kernel, std=.01)
        nn.init.normal_(self.reg_conv.kernel, std=.01)
        nn.init.normal_(self.cls_conv.kernel, std=.01)
        nn.init.constant_(self.cls_conv.bias, bias_init_with_prob(.01))
        nn.init.normal_(self.semantic_conv.kernel, std=.01)
        nn.init.constant_(self.semantic_conv.bias, bias_init_with_prob(.01))
        for cls_id in range(self.n_classes):
            nn.init.normal_(self.cls_individual_out[cls_id][0].kernel, std=.01)

    def _load_from_state_dict(self, *a, **k):
        self.fold.clear()
        return super()._load_from_state_dict(*a, **k)

    def train(self, mode: bool = True):
        """switching between training and evaluation drops the folded / stacked copies of the parameters (an optimizer
        step or new running statistics made them stale)"""
        self.fold.clear()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.fold.clear()
        return super()._apply(fn, *a, **k)

    # ---- stacked per-class parameters for the grouped launches --------------------------------------
    def _stacked(self, dev):
        def build():
            with torch.no_grad():
                st = lambda ts: torch.stack([t.detach().float() for t in ts]).contiguous().to(dev)
                bn = lambda ms: [st(x) for x in zip(*[self.fold.bn(m) for m in ms])]
                d = dict(
                    W_out=st([m[0].kernel for m in self.cls_individual_out]),
                    bn_out=bn([m[1] for m in self.cls_individual_out]),
                    W_exp=st([m[0].kernel for m in self.cls_individual_expand_out]),
                    bn_exp=bn([m[1] for m in self.cls_individual_expand_out]),
                    W_up=st([m[0].kernel for m in self.cls_individual_up]),
                    bn_up=bn([m[1][0] for m in self.cls_individual_up]),
                    W_fuse=st([m[0].kernel.unsqueeze(0) for m in self.cls_individual_fuse]),
                    bn_fuse=bn([m[1] for m in self.cls_individual_fuse]),
                    scales=st([s.scale for s in self.scales]),
                    vsA=torch.tensor(self.voxel_size_list, dtype=torch.float32, device=dev),
                )
                d["vsE"] = (d["vsA"] * self.expand).contiguous()                       # fp32 product, :266
                # one GEMM for centerness | cls | reg (forward_single, :627-636)
                d["W_pred"] = torch.cat([self.centerness_conv.kernel, self.cls_conv.kernel, self.reg_conv.kernel],
                                        1).detach().float().contiguous().to(dev)
                b = torch.zeros((d["W_pred"].shape[1],), dtype=torch.float32, device=dev)
                b[1:1 + self.n_classes] = self.cls_conv.bias.detach().float().reshape(-1).to(dev)
                d["b_pred"] = b
                d["b_sem"] = self.semantic_conv.bias.detach().float().reshape(-1).contiguous().to(dev)
            return d
        return self.fold.get(("stacked", str(dev)), build)

    # ---- forward ------------------------------------------------------------------------------------
    def class_maps(self, out: S.SparseTensor, B: int):
        """cagroup_head.py:205-282 for all classes at once.  Returns a dict of device tensors."""
        dev, fc, P = out.F.device, self.fold, self._stacked(out.F.device)
        N, ncls, nv, C = out.cmap.n, self.n_classes, (3 if self.with_yaw else 1), self.out_channels
        sem = S.gemm_rows(out.F, None, self.semantic_conv.kernel, N, 1, shift=P["b_sem"])
        pad_rows = _i32(B, device=dev)
        S._call("cg3d_first_rows", out.C, N, B, pad_rows)
        mm = _i32(6, device=dev)
        S._call("cg3d_coord_bounds", out.C, N, mm)
        ob = self.offset_block
        h = conv_bn(out, ob[0], ob[1], fc, act="elu", split_out="none")      # + the split copy the next conv gathers from
        h = conv_bn(h, ob[3], ob[4], fc, act="elu")
        offs = conv_bn(h, ob[6], None, fc).F
        offF = conv_bn(out, self.feature_offset[0], self.feature_offset[1], fc, act="elu").F
        voted = _f32(N, nv, 3, device=dev)
        S._call("cg3d_vote_points", out.C, offs, N, nv, float(self.voxel_size), out.cmap.stride, mm, voted)
        flags = _i32(ncls * N, device=dev)
        S._call("cg3d_semantic_flags", sem, N, ncls, float(self.semantic_threshold), flags)
        pos, total = S.exclusive_scan(flags)
        # host sync 1: per-class selection counts (sizes of everything below depend on them)
        bounds = torch.cat([pos[::N][:ncls], total]).cpu().tolist()
        sel_rows = _i32(max(bounds[-1], 1), device=dev)
        S._call("cg3d_compact_rows", flags, pos, N, ncls, sel_rows)
        fused = [0]
        for c in range(ncls):
            fused.append(fused[-1] + (nv + 1) * (bounds[c + 1] - bounds[c] + B))
        nf = fused[-1]
        meta = torch.tensor([bounds, fused], dtype=torch.int32).to(dev)
        coordsA, coordsE, ref = _i32(nf, 4, device=dev), _i32(nf, 4, device=dev), _i32(nf, 2, device=dev)
        S._call("cg3d_class_points", out.C, voted, sel_rows, meta[0], meta[1], pad_rows, P["vsA"], P["vsE"], ncls, B, nv,
                self.expand, nf, float(self.voxel_size), coordsA, coordsE, ref)
        mgr = S.Manager(batch_bits=max(1, (ncls * B - 1).bit_length()))
        mapA, _, invA = S.unique_first(coordsA, 1, mgr, want_inverse=True)                  # sync 2
        mapE, _, invE = S.unique_first(coordsE, self.expand, mgr, want_inverse=True)        # sync 3
        # class row ranges: the first fused point of a class is a first occurrence, so its unique row starts the class (rows
        # are in first-occurrence order).  Read back HERE (sync 4, one copy for both maps), right after the syncs of the two
        # unique passes (the GPU is idle there anyway): everything below is then queued without a host wait, and the tile tables /
        # rule-map launches are prepared while the GPU averages the features (this read used to sit after the segment means:
        # a 0.35 ms idle gap)
        starts = meta[1][:ncls].long()
        offs2 = torch.stack([invA[starts], invE[starts]]).cpu().tolist()
        offA, offE = offs2[0] + [mapA.n], offs2[1] + [mapE.n]
        # the 9^3 rule map of the class voxels (1.2 ms of hash probes) depends on coordinates only: it is built on the
        # coordinate stream while this stream averages the features and runs the 5^3 / transposed branch
        from .backbone import _COORD_STREAM, _side_stream
        if _COORD_STREAM["on"]:
            main = torch.cuda.current_stream()
            mgr.stream = _side_stream(dev, "coord")
            mgr.stream.wait_stream(main)
            S.neighbor_table(mapA, mapA, self.cls_kernel, mgr, ordered=True, group_div=B, wait=False)
            mgr.stream = None
        FA = S.segment_mean(offF, offF.shape[1], out.F, out.F.shape[1], ref, invA, nf, mapA.n, C)
        FE = S.segment_mean(offF, offF.shape[1], out.F, out.F.shape[1], ref, invE, nf, mapE.n, C)
        # (host work from here on runs while the GPU is busy with the launches above)
        tile = 128 if S.get_conv_impl() == "tc" else 64
        tilesA, tilesE = S.make_tiles(offA, dev, tile), S.make_tiles(offE, dev, tile)
        cat = _f32(mapA.n, 2 * C, device=dev)                                               # [up | out] (:276-277)
        # tile order keeps rows class-major (the class is the high part of the batch index), so the per-class
        # position ranges equal the per-class row ranges
        nbrE, ordE = S.neighbor_table(mapE, mapE, 5, mgr, ordered=True, group_div=B)
        EF = S.gemm_rows(FE, nbrE, P["W_exp"], mapE.n, 125, scale=P["bn_exp"][0], shift=P["bn_exp"][1], act="elu",
                         tiles=tilesE, out_rows=ordE, split_out="none")
        nbrU, ordU = S.transpose_table(mapE, mapA, self.expand, mgr, ordered=True, group_div=B)
        S.gemm_rows(EF, nbrU, P["W_up"], mapA.n, self.expand ** 3, scale=P["bn_up"][0], shift=P["bn_up"][1],
                    act="elu", tiles=tilesA, out=cat[:, :C], out_rows=ordU)
        nbrA, ordA = S.neighbor_table(mapA, mapA, self.cls_kernel, mgr, ordered=True, group_div=B)
        S.gemm_rows(FA, nbrA, P["W_out"], mapA.n, self.cls_kernel ** 3, scale=P["bn_out"][0], shift=P["bn_out"][1],
                    act="elu", tiles=tilesA, out=cat[:, C:], out_rows=ordA)
        O = S.gemm_rows(cat, None, P["W_fuse"], mapA.n, 1, scale=P["bn_fuse"][0], shift=P["bn_fuse"][1], act="elu",
                        tiles=tilesA)
        pred = S.gemm_rows(O, None, P["W_pred"], mapA.n, 1, shift=P["b_pred"])
        return dict(sem=sem, offsets=offs, offset_feat=offF, voted=voted, coords=mapA.coords, feat=O, pred=pred,
                    class_off=offA, n_sel=[bounds[c + 1] - bounds[c] + B for c in range(ncls)], mapA=mapA, mapE=mapE)

    def proposals(self, cm: dict, B: int):
        """get_bboxes + _nms (cagroup_head.py:557-624, 747-797), device resident."""
        dev, ncls, P = cm["pred"].device, self.n_classes, self._stacked(cm["pred"].device)
        V = cm["pred"].shape[0]
        scores, maxscore, boxes = _f32(V, ncls, device=dev), _f32(V, device=dev), _f32(V, 7, device=dev)
        S._call("cg3d_head_decode", cm["pred"], cm["pred"].shape[1], cm["coords"], V, ncls, self.n_reg_outs, B, P["vsA"],
                P["scales"], scores, maxscore, boxes, 7)
        nseg = B * ncls
        seg = _i32(V, device=dev)
        S._call("cg3d_map_segments", cm["coords"], V, B, ncls, seg)
        counts = _i32(nseg + 1, device=dev)
        counts[nseg:].zero_()
        S._call("cg3d_histogram_i32", seg, V, nseg, counts)
        seg_off, _ = S.exclusive_scan(counts)
        keys, rows = _u64(V, dev), _i32(V, device=dev)
        S._call("cg3d_topk_keys", seg, maxscore, V, counts, int(self.nms_pre), keys, rows)
        sort_pairs(keys, rows, V, seg_bits(nseg))
        flags = _i32(V, device=dev)
        S._call("cg3d_rank_filter", keys, V, seg_off, int(self.nms_pre), flags)
        pos, total = S.exclusive_scan(flags)
        cand = _i32(V, device=dev)
        S._call("cg3d_compact_i32", flags, pos, V, rows, cand)
        # candidates <= V; the pair list is sized by its upper bound to avoid a sync
        nc = int(total.item())                                                               # sync
        pflags = _i32(max(nc * ncls, 1), device=dev)
        S._call("cg3d_pair_flags", scores, cand, nc, ncls, float(self.score_thr), pflags)
        ppos, ptotal = S.exclusive_scan(pflags[:nc * ncls])
        n = int(ptotal.item())                                                               # sync
        pkeys, prow = _u64(n, dev), _i32(max(n, 1), device=dev)
        S._call("cg3d_pair_keys", scores, cand, seg, nc, ncls, pflags, ppos, pkeys, prow)
        out = nms_and_pack(pkeys, prow, n, boxes, 7, nseg, ncls, B, self.iou_thr, self.with_yaw,
                           gather_flip=int(self.with_yaw))
        return out + (dict(scores=scores, boxes=boxes, cand=cand[:nc]),)

    def run(self, out: S.SparseTensor, B: int):
        cm = self.class_maps(out, B)
        det_boxes, det_scores, det_labels, sample_off, extra = self.proposals(cm, B)
        cm.update(extra)
        return det_boxes, det_scores, det_labels, sample_off, cm

    def forward(self, input_dict, return_middle_feature=True):
        B = input_dict["batch_size"]
        out = input_dict["sp_tensor"]
        det_boxes, det_scores, det_labels, off, cm = self.run(out, B)
        labels64 = det_labels.long()                                    # one conversion, B views
        bbox_list = [(det_boxes[off[b]:off[b + 1]], det_scores[off[b]:off[b + 1]], labels64[off[b]:off[b + 1]])
                     for b in range(B)]
        return {
            "one_stage_results": (cm, cm["sem"], cm["offsets"]),
            "middle_feature_list": [None, None, None, out] if return_middle_feature else None,
            "pred_bbox_list": bbox_list,
            "_packed_proposals": (det_boxes, det_scores, det_labels, off),
        }
