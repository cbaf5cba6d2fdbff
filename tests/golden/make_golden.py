"""Generates tests/golden/*.npz by running the REFERENCE's own Python on the CPU (TEST INFRASTRUCTURE).

    python tests/golden/make_golden.py            # needs /root/reference; run in the build container only

What runs here is the unmodified reference code imported from /root/reference:
  * pcdet.config.cfg_from_yaml_file on tools/cfgs/{scannet,sunrgbd}_models/CAGroup3D.yaml,
  * pcdet.models.build_network -> CAGroup3D / BiResNet / CAGroup3DHead / CAGroup3DRoIHead,
  * model.load_state_dict(strict=True) of OUR seeded state dict (so the parameter names and shapes of
    SURVEY.md Appendix C are checked against the reference's real module tree),
  * model.eval(); model(batch_dict) -- the whole forward of cagroup3d.py:27-50,
  * pcdet.models.model_utils.cagroup_utils / pcdet.utils.loss_utils / pcdet.ops.rotated_iou helpers for
    stand-alone known-answer vectors,
  * the reference's compiled CPU IoU (oracle/_ref/iou3d_nms_cuda.so, boxes_iou_bev_cpu).
What does NOT exist offline and is substituted:
  * MinkowskiEngine -> tests/golden/me_shim.py (the oracle's restatement of its semantics);
  * the CUDA-only NMS entry points nms_gpu / nms_normal_gpu -> the same greedy suppression
    (iou3d_nms.cpp:103-132) on the CPU, rotated IoU from the reference's boxes_iou_bev_cpu, axis-aligned
    IoU per iou3d_nms_kernel.cu:314-325;
  * every other missing third-party / compiled module (spconv, easydict, SharedArray, pointnet2 ops ...)
    -> inert stubs; they are import-time dependencies only.
The fixtures hold inputs (seeds, the two calibrated bias vectors) and the reference's outputs; weights are
regenerated from the seed by cagroup3d_b200.model_init.seeded_model, so the files stay small.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CG3D_REFERENCE", "/root/reference")
# the reference's `pcdet` must win over this repo's drop-in `pcdet` package; everything else comes from the repo
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
for _m in [m for m in sys.modules if m == "pcdet" or m.startswith("pcdet.")]:
    del sys.modules[_m]

STUB_TOP = {"spconv", "cumm", "SharedArray", "tensorboardX", "skimage", "open3d", "terminaltables", "kornia",
            "mayavi", "av2", "nuscenes", "waymo_open_dataset", "tensorflow", "lyft_dataset_sdk", "pandaset", "cv2",
            "matplotlib", "numba", "easydict", "sort_vertices", "KNN_OP", "tkinter", "turtle"}


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (torch.nn.Module,), {"__init__": lambda self, *a, **k: torch.nn.Module.__init__(self)})
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: inert modules for absent third-party packages and un-built pcdet extensions."""

    def find_spec(self, name, path=None, target=None):
        top = name.split(".")[0]
        if top in STUB_TOP or (top == "pcdet" and (name.endswith("_cuda") or name.endswith("KNN_OP") or
                                                   name.endswith("sort_vertices"))):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


class EasyDict(dict):
    """the slice of easydict.EasyDict that pcdet/config.py uses."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, d=None, **kw):
        for k, v in dict(d or {}, **kw).items():
            self[k] = v


def install():
    from tests.golden import me_shim
    me_shim.install(sys.modules)
    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed
    sys.meta_path.append(_StubFinder())
    for alias, t in (("int", int), ("float", float), ("bool", bool), ("object", object)):
        if alias not in np.__dict__:
            setattr(np, alias, t)                 # aliases the reference (numpy < 1.24 era) still uses
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    # the reference's compiled iou3d op (CPU IoU real; the CUDA entry points are replaced below)
    from oracle import build_ref
    build_ref.build()
    ref_iou = build_ref.load("iou3d_nms_cuda")
    sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_cuda"] = ref_iou
    return ref_iou


def cpu_nms_factory(ref_iou):
    def greedy(iou, n):
        keep, dead = [], np.zeros((n,), bool)
        for i in range(n):
            if dead[i]:
                continue
            keep.append(i)
            dead |= iou[i] > greedy.thr
        return keep

    def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kw):
        order = torch.sort(scores, dim=0, descending=True, stable=True)[1]
        if pre_maxsize is not None:
            order = order[:pre_maxsize]
        b = boxes[order].contiguous().float()
        iou = torch.zeros((len(b), len(b)))
        if len(b):
            ref_iou.boxes_iou_bev_cpu(b, b, iou)
        greedy.thr = thresh
        keep = greedy(iou.numpy(), len(b))
        return order[torch.tensor(keep, dtype=torch.long)].contiguous(), None

    def nms_normal_gpu(boxes, scores, thresh, **kw):
        order = torch.sort(scores, dim=0, descending=True, stable=True)[1]
        b = boxes[order].contiguous().float()
        # iou_normal (iou3d_nms_kernel.cu:314-325), fp32
        x1, y1 = b[:, 0] - b[:, 3] / 2, b[:, 1] - b[:, 4] / 2
        x2, y2 = b[:, 0] + b[:, 3] / 2, b[:, 1] + b[:, 4] / 2
        l, r = torch.max(x1[:, None], x1[None]), torch.min(x2[:, None], x2[None])
        t, bt = torch.max(y1[:, None], y1[None]), torch.min(y2[:, None], y2[None])
        inter = (r - l).clamp(min=0) * (bt - t).clamp(min=0)
        area = b[:, 3] * b[:, 4]
        iou = inter / torch.clamp(area[:, None] + area[None] - inter, min=1e-8)
        greedy.thr = thresh
        keep = greedy(iou.numpy(), len(b))
        return order[torch.tensor(keep, dtype=torch.long)].contiguous(), None

    return nms_gpu, nms_normal_gpu


class _Dataset:
    """what Detector3DTemplate.build_networks reads from the dataset (detector3d_template.py:35-45)."""

    def __init__(self, class_names):
        self.class_names = class_names
        self.point_feature_encoder = types.SimpleNamespace(num_point_features=6)
        self.grid_size = np.array([1, 1, 1])
        self.point_cloud_range = np.array([-40, -40, -10, 40, 40, 10], dtype=np.float32)
        self.voxel_size = [0.02, 0.02, 0.02]
        self.depth_downsample_factor = None


def reference_model(dataset: str):
    """build the reference CAGroup3D through its own config + registry path."""
    from pcdet.config import cfg, cfg_from_yaml_file
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "tools"))                     # _BASE_CONFIG_ is relative to tools/
    try:
        for k in list(cfg.keys()):
            if k not in ("ROOT_DIR", "LOCAL_RANK"):
                del cfg[k]
        cfg_from_yaml_file(f"cfgs/{dataset}_models/CAGroup3D.yaml", cfg)
    finally:
        os.chdir(cwd)
    from pcdet.models import build_network
    import pcdet.models.dense_heads.cagroup_head as H
    import pcdet.models.roi_heads.cagroup_roi_head as R
    return build_network(model_cfg=cfg.MODEL, num_class=len(cfg.CLASS_NAMES), dataset=_Dataset(cfg.CLASS_NAMES)), cfg, H, R


def t2n(x):
    return x.detach().cpu().numpy()


def run_model_case(name, dataset, n_classes, with_yaw, seed, voxels, ref_iou, p_sel=0.08, p_box=0.01):
    from cagroup3d_b200 import model_init, synthetic
    from oracle import cagroup3d_oracle as O
    B = 2
    batch = synthetic.make_batch(B, target_voxels=voxels, config=7, n_classes=n_classes, sunrgbd=with_yaw)
    pts = torch.from_numpy(batch["points"])
    ours = model_init.seeded_model(n_classes, with_yaw, seed=seed)
    # declared head-occupancy knobs (model_init.py), computed with the oracle; only the two bias vectors
    # are stored in the fixture
    orc = O.Oracle(ours.state_dict(), O.default_cfg(n_classes, with_yaw))
    bb = orc.forward(pts, B, stages="backbone")
    model_init.calibrate_semantic_bias(ours, bb["bb_feats"], p_sel)
    orc = O.Oracle(ours.state_dict(), O.default_cfg(n_classes, with_yaw))
    mid = orc.forward(pts, B, cur_epoch=10, stages="head")
    pred_all = torch.cat([torch.cat([m["ctr"], m["cls"], m["reg"]], 1) for m in mid["head"]["maps"]])
    model_init.calibrate_cls_bias(ours, pred_all, p_box)

    model, cfg, H, R = reference_model(dataset)
    nms_gpu, nms_normal_gpu = cpu_nms_factory(ref_iou)
    H.nms_gpu, H.nms_normal_gpu, R.nms_gpu, R.nms_normal_gpu = nms_gpu, nms_normal_gpu, nms_gpu, nms_normal_gpu
    sd = ours.state_dict()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("num_batches_tracked") or "code_weights" in k for k in missing), missing
    assert set(model.state_dict().keys()) - set(sd.keys()) <= set(missing)
    model.eval()
    bd = {"points": pts.clone(), "batch_size": B, "cur_epoch": 10}
    with torch.no_grad():
        pred_dicts, _ = model(bd)
    out = {
        "seed": seed, "voxels": voxels, "n_classes": n_classes, "with_yaw": int(with_yaw), "config": 7, "batch": B,
        "p_sel": p_sel, "p_box": p_box,
        "semantic_bias": t2n(ours.dense_head.semantic_conv.bias), "cls_bias": t2n(ours.dense_head.cls_conv.bias),
        "bb_coords": t2n(bd["sp_tensor"].C) if hasattr(bd["sp_tensor"], "C") else None,
    }
    sp = bd["middle_feature_list"][3]
    out["bb_coords"], out["bb_feats"] = t2n(sp.C), t2n(sp.F)
    (ctr, bbox, cls, pts_l), sem, offs = bd["one_stage_results"]
    out["sem"], out["offsets"] = t2n(sem.F), t2n(offs.F)
    for c in range(n_classes):
        for b in range(B):
            out[f"map_c{c}_b{b}"] = np.concatenate([t2n(pts_l[c][b]), t2n(ctr[c][b]), t2n(cls[c][b]), t2n(bbox[c][b])], 1)
    for b in range(B):
        bx, sc, lb = bd["pred_bbox_list"][b][:3]
        out[f"stage1_b{b}"] = np.concatenate([t2n(bx), t2n(sc)[:, None], t2n(lb).astype(np.float32)[:, None]], 1)
        pd = pred_dicts[b]
        out[f"final_b{b}"] = np.concatenate([t2n(pd["pred_boxes"]), t2n(pd["pred_scores"])[:, None],
                                              t2n(pd["pred_labels"]).astype(np.float32)[:, None]], 1)
    out["rois"], out["rcnn_reg"] = t2n(bd["rois"]), t2n(bd["rcnn_reg"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "voxels", len(out["bb_coords"]), "stage1", [len(out[f"stage1_b{b}"]) for b in range(B)],
          "final", [len(out[f"final_b{b}"]) for b in range(B)])


def run_helper_vectors(ref_iou):
    """stand-alone known-answer vectors from the reference's pure-torch helpers and its compiled CPU IoU."""
    from pcdet.models.model_utils.cagroup_utils import CAGroupResidualCoder, rotation_3d_in_axis
    from pcdet.utils.loss_utils import axis_aligned_bbox_overlaps_3d
    from pcdet.utils import common_utils
    import pcdet.models.dense_heads.cagroup_head as H
    import pcdet.models.roi_heads.cagroup_roi_head as R
    g = torch.Generator().manual_seed(123)
    out = {}
    # rotated / axis-aligned BEV IoU
    n = 96
    boxes = torch.cat([(torch.rand((n, 3), generator=g) - 0.5) * 4, torch.rand((n, 3), generator=g) * 2 + 0.2,
                       (torch.rand((n, 1), generator=g) - 0.5) * 6.3], 1)
    boxes[:8, 6] = 0.0
    boxes[8:16] = boxes[:8]                      # identical pairs
    boxes[16:24, :2] = boxes[:8, :2] + 0.5       # shifted copies
    boxes[16:24, 2:] = boxes[:8, 2:]
    iou = torch.zeros((n, n))
    ref_iou.boxes_iou_bev_cpu(boxes.contiguous(), boxes.contiguous(), iou)
    out["iou_boxes"], out["iou_bev"] = t2n(boxes), t2n(iou)
    kat = torch.tensor([[0, 0, 0, 2, 2, 1, 0.], [0.5, 0, 0, 2, 2, 1, 0.]])
    k2 = torch.zeros((2, 2))
    ref_iou.boxes_iou_bev_cpu(kat, kat, k2)
    out["iou_kat"] = t2n(k2)                      # [[1, .6], [.6, 1]]
    # residual coder, both code sizes
    for cs, sincos in ((6, False), (7, True)):
        coder = CAGroupResidualCoder(code_size=cs, encode_angle_by_sincos=sincos)
        anchors = torch.cat([(torch.rand((64, 3), generator=g) - 0.5) * 4, torch.rand((64, 3), generator=g) * 2 + 0.1,
                             (torch.rand((64, 1), generator=g) - 0.5) * 6], 1)[:, :cs]
        enc = torch.randn((64, cs + (1 if sincos else 0)), generator=g) * 0.3
        out[f"coder{cs}_anchors"], out[f"coder{cs}_enc"] = t2n(anchors), t2n(enc)
        out[f"coder{cs}_dec"] = t2n(coder.decode_torch(enc.clone(), anchors.clone()))
    # rotate_points_along_z / rotation_3d_in_axis
    p = torch.randn((16, 5, 3), generator=g)
    a = (torch.rand((16,), generator=g) - 0.5) * 6
    out["rot_pts"], out["rot_ang"] = t2n(p), t2n(a)
    out["rot_along_z"] = t2n(common_utils.rotate_points_along_z(p.clone(), a))
    out["rot_axis2"] = t2n(rotation_3d_in_axis(p.clone(), a, axis=2))
    # FCOS distance -> box decode, all three parametrisations that reach it (cagroup_head.py:654-703)
    pts = torch.randn((128, 3), generator=g)
    pred6 = torch.rand((128, 6), generator=g) * 2
    pred8 = torch.cat([pred6, torch.randn((128, 2), generator=g)], 1)
    dummy = type("D", (), {"yaw_parametrization": "fcaf3d"})()
    out["dec_pts"], out["dec_pred8"] = t2n(pts), t2n(pred8)
    out["dec_box6"] = t2n(H.CAGroup3DHead._bbox_pred_to_bbox(dummy, pts, pred6))
    out["dec_box8"] = t2n(H.CAGroup3DHead._bbox_pred_to_bbox(dummy, pts, pred8))
    # RoI grid points (cagroup_roi_head.py:199-224) with and without yaw
    rois = torch.cat([(torch.rand((10, 3), generator=g) - 0.5) * 6, torch.rand((10, 3), generator=g) * 2 + 0.1,
                      (torch.rand((10, 1), generator=g) - 0.5) * 6], 1)
    for cs in (6, 7):
        d = type("D", (), {"code_size": cs, "get_dense_grid_points": staticmethod(R.CAGroup3DRoIHead.get_dense_grid_points)})()
        gpts, _ = R.CAGroup3DRoIHead.get_global_grid_points_of_roi(d, rois[:, :7].clone(), 7)
        out[f"grid_pts{cs}"] = t2n(gpts)
    out["grid_rois"] = t2n(rois)
    # SimplePoolingLayer coordinate arithmetic (cagroup_roi_head.py:54-68): the unique grid voxels the 5^3 conv is
    # evaluated at and the inverse map, captured by a recording stand-in for grid_conv
    class _Stop(Exception):
        pass
    rec = {}

    class _Recorder(torch.nn.Module):
        def forward(self, sp, coords):
            rec["coords"] = coords.clone()
            raise _Stop()
    layer = R.SimplePoolingLayer(channels=[64, 128, 128], grid_kernel_size=5, grid_num=7, voxel_size=0.04, coord_key=2,
                                 pooling=True)
    layer.grid_conv = _Recorder()
    gp = torch.cat([torch.zeros((10 * 343, 1)), out_t(out["grid_pts6"]).reshape(-1, 3)], 1)
    gp[5 * 343:, 0] = 1
    gp[:7, 1] = 100.0                               # exercise the +-191 clamp
    try:
        layer(None, grid_points=gp)
    except _Stop:
        pass
    out["pool_grid_points"], out["pool_unique_coords"] = t2n(gp), t2n(rec["coords"])
    # axis-aligned 3D IoU (loss_utils.py:389-540) on corner-format boxes
    c1 = torch.rand((32, 3), generator=g)
    b1 = torch.cat([c1, c1 + torch.rand((32, 3), generator=g) + 0.05], 1)
    c2 = torch.rand((32, 3), generator=g)
    b2 = torch.cat([c2, c2 + torch.rand((32, 3), generator=g) + 0.05], 1)
    out["aa_b1"], out["aa_b2"] = t2n(b1), t2n(b2)
    out["aa_iou_aligned"] = t2n(axis_aligned_bbox_overlaps_3d(b1, b2, is_aligned=True))
    # differentiable rotated IoU, torch part (box_intersection_2d.py): candidate vertices + mask; sort_vertices input
    from pcdet.ops.rotated_iou.box_intersection_2d import box_intersection_th, box_in_box_th, build_vertices
    from pcdet.ops.rotated_iou.oriented_iou_loss import box2corners_th
    bx1 = torch.cat([(torch.rand((1, 48, 2), generator=g) - 0.5) * 2, torch.rand((1, 48, 2), generator=g) + 0.3,
                     (torch.rand((1, 48, 1), generator=g) - 0.5) * 3], 2)
    bx2 = bx1 + torch.randn((1, 48, 5), generator=g) * 0.15
    bx2[..., 2:4] = bx2[..., 2:4].abs() + 0.1
    cr1, cr2 = box2corners_th(bx1), box2corners_th(bx2)
    inters, mask_inter = box_intersection_th(cr1, cr2)
    c12, c21 = box_in_box_th(cr1, cr2)
    vertices, mask = build_vertices(cr1, cr2, c12, c21, inters, mask_inter)
    out["sv_boxes1"], out["sv_boxes2"] = t2n(bx1), t2n(bx2)
    out["sv_vertices"], out["sv_mask"] = t2n(vertices), t2n(mask)
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), **out)
    print("helpers.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def out_t(a):
    return torch.from_numpy(a)


def main():
    ref_iou = install()
    run_helper_vectors(ref_iou)
    run_model_case("scannet_small", "scannet", 18, False, seed=3, voxels=1500, ref_iou=ref_iou)
    run_model_case("sunrgbd_small", "sunrgbd", 10, True, seed=4, voxels=1500, ref_iou=ref_iou)


if __name__ == "__main__":
    main()
