"""Generates tests/golden/scannet_train_small.npz: the REFERENCE's training step on the CPU (TEST INFRASTRUCTURE for the
next coverage row, SURVEY.md 8f rank 1 -- losses, assigner, gradients; run in the build container only).

    python tests/golden/make_train_golden.py

Everything of tests/golden/make_golden.py applies (the reference's unmodified Python from /root/reference, MinkowskiEngine
served by tests/golden/me_shim.py over oracle/me_cpu.py -- whose ops are differentiable torch ops, so the reference's
loss.backward() runs through them).  Additional CUDA-only entry points and their CPU stand-ins:
  * pcdet.ops.knn.knn (knn_cuda.cu:26-94, k = 1)  -> torch.cdist + topk (same strict-< / first-index tie rule for k = 1);
  * iou3d_nms_utils.boxes_iou3d_gpu (ProposalTargetLayer) -> BEV IoU from the reference's compiled boxes_iou_bev_cpu x
    height overlap (iou3d_nms_utils.py:48-81);
  * nms_gpu / nms_normal_gpu -> make_golden.cpu_nms_factory on detached tensors.
Inputs: two synthetic ScanNet-shaped scenes with per-point semantic / instance masks (cagroup3d_b200.synthetic,
return_masks=True), seed-3 weights with the two calibrated bias vectors.  Stored: every entry of the reference's tb_dict,
the total loss, the L2 norm of the gradient of EVERY parameter and a few small gradients in full.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import make_golden as MG          # noqa: E402  (puts /root/reference first on sys.path)


def main():
    ref_iou = MG.install()
    from cagroup3d_b200 import model_init, synthetic
    from oracle import cagroup3d_oracle as O
    B, ncls, yaw, seed, voxels, p_sel, p_box = 2, 18, False, 3, 2500, 0.08, 0.02
    scenes = [synthetic.make_scene(1000 * 7 + i, voxels, n_classes=ncls, sunrgbd=yaw, return_masks=True) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    pts = torch.from_numpy(batch["points"])
    ours = model_init.seeded_model(ncls, yaw, seed=seed)
    orc = O.Oracle(ours.state_dict(), O.default_cfg(ncls, yaw))
    bb = orc.forward(pts, B, stages="backbone")
    model_init.calibrate_semantic_bias(ours, bb["bb_feats"], p_sel)
    orc = O.Oracle(ours.state_dict(), O.default_cfg(ncls, yaw))
    mid = orc.forward(pts, B, cur_epoch=10, stages="head")
    pred_all = torch.cat([torch.cat([m["ctr"], m["cls"], m["reg"]], 1) for m in mid["head"]["maps"]])
    model_init.calibrate_cls_bias(ours, pred_all, p_box)

    model, cfg, H, R = MG.reference_model("scannet")
    _ng, _nn = MG.cpu_nms_factory(ref_iou)
    nms_gpu = lambda b, s, t, *a, **k: _ng(b.detach(), s.detach(), t, *a, **k)
    nms_normal_gpu = lambda b, s, t, **k: _nn(b.detach(), s.detach(), t, **k)
    H.nms_gpu, H.nms_normal_gpu, R.nms_gpu, R.nms_normal_gpu = nms_gpu, nms_normal_gpu, nms_gpu, nms_normal_gpu

    def knn_cpu(k, xyz, center_xyz=None, transposed=False):
        d = torch.cdist(center_xyz, xyz)                       # (B, M, N)
        return d.topk(k, dim=2, largest=False)[1].transpose(1, 2).int().contiguous()
    H.knn = knn_cpu
    import pcdet.models.roi_heads.target_assigner.cagroup_proposal_target_layer as PT

    def boxes_iou3d_cpu(a, b):
        iou_bev = torch.zeros((len(a), len(b)))
        if len(a) and len(b):
            ref_iou.boxes_iou_bev_cpu(a[:, :7].detach().contiguous().float(), b[:, :7].detach().contiguous().float(), iou_bev)
        area = (a[:, 3] * a[:, 4])[:, None] + (b[:, 3] * b[:, 4])[None]
        inter_bev = iou_bev * area / (1 + iou_bev)
        top = torch.min((a[:, 2] + a[:, 5] / 2)[:, None], (b[:, 2] + b[:, 5] / 2)[None])
        bot = torch.max((a[:, 2] - a[:, 5] / 2)[:, None], (b[:, 2] - b[:, 5] / 2)[None])
        inter = inter_bev * (top - bot).clamp(min=0)
        vol = (a[:, 3] * a[:, 4] * a[:, 5])[:, None] + (b[:, 3] * b[:, 4] * b[:, 5])[None]
        return (inter / torch.clamp(vol - inter, min=1e-6)).detach()
    PT.boxes_iou3d_gpu = boxes_iou3d_cpu

    model.load_state_dict(ours.state_dict(), strict=False)
    model.train()
    torch.manual_seed(0)
    np.random.seed(0)
    bd = {"points": pts.clone(), "batch_size": B, "cur_epoch": 10, "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float(),
          "semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
    ret, tb, disp = model(bd)
    ret["loss"].backward()
    out = {"seed": seed, "voxels": voxels, "n_classes": ncls, "config": 7, "batch": B, "p_sel": p_sel, "p_box": p_box, "cur_epoch": 10,
           "semantic_bias": MG.t2n(ours.dense_head.semantic_conv.bias), "cls_bias": MG.t2n(ours.dense_head.cls_conv.bias),
           "loss": float(ret["loss"].item())}
    for k, v in tb.items():
        out["tb_" + k] = float(v)
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.norm().item()) if p.grad is not None else -1.0)
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms, np.float64)
    for n in ("dense_head.semantic_conv.bias", "dense_head.cls_conv.bias", "dense_head.centerness_conv.kernel",
              "dense_head.scales.0.scale", "backbone_3d.conv1.1.bn.weight"):
        p = dict(model.named_parameters())[n]
        out["grad__" + n] = MG.t2n(p.grad) if p.grad is not None else np.zeros(0)
    out["n_rois"] = int(bd["rois"].shape[1]) if "rois" in bd else 0
    np.savez_compressed(os.path.join(HERE, "scannet_train_small.npz"), **out)
    print({k: round(v, 5) for k, v in tb.items()}, "rois", out["n_rois"], "params with grad", int((out["grad_norms"] >= 0).sum()), "/", len(names))


if __name__ == "__main__":
    main()
