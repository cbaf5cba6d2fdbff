"""CPU: the oracle against golden vectors produced by the reference's own code (tests/golden/make_golden.py).

helpers.npz      reference pure-torch helpers + the reference's compiled CPU IoU (boxes_iou_bev_cpu)
*_small.npz      the reference's whole CAGroup3D forward (its unmodified Python, MinkowskiEngine calls served by
                 tests/golden/me_shim.py) on two small synthetic scenes, ScanNet (18 cls) and SUN RGB-D (10 cls, yaw)
Bars: indices / coordinates exact; floats 1e-4 absolute here (both sides are fp32 CPU, only the summation order
differs), detections matched one-to-one.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cagroup3d_oracle as O
from oracle import iou3d_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def helpers():
    return np.load(os.path.join(GOLD, "helpers.npz"))


def test_iou_bev_cpu_reference(helpers):
    b = torch.from_numpy(helpers["iou_boxes"])
    got = iou3d_oracle.pairwise(b, b, "iou").numpy()
    assert np.abs(got - helpers["iou_bev"]).max() <= 1e-6
    assert np.allclose(helpers["iou_kat"], [[1.0, 0.6], [0.6, 1.0]], atol=1e-6)      # SURVEY 8c known answers


def test_live_reference_iou_if_built(helpers):
    """when oracle/_ref holds the reference's compiled op, compare against it directly as well."""
    from oracle import build_ref
    mod = build_ref.load("iou3d_nms_cuda")
    if mod is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    g = torch.Generator().manual_seed(5)
    b = torch.cat([(torch.rand((200, 3), generator=g) - 0.5) * 3, torch.rand((200, 3), generator=g) + 0.2,
                   (torch.rand((200, 1), generator=g) - 0.5) * 7], 1)
    want = torch.zeros((200, 200))
    mod.boxes_iou_bev_cpu(b, b, want)
    got = iou3d_oracle.pairwise(b, b, "iou")
    assert (got - want).abs().max().item() <= 1e-6


def test_residual_coder_and_rotation(helpers):
    for cs, sincos in ((6, False), (7, True)):
        dec = O.residual_decode(torch.from_numpy(helpers[f"coder{cs}_enc"]), torch.from_numpy(helpers[f"coder{cs}_anchors"]),
                                cs, sincos)
        assert np.abs(dec.numpy() - helpers[f"coder{cs}_dec"]).max() <= 1e-6
    got = O.rotate_z(torch.from_numpy(helpers["rot_pts"]), torch.from_numpy(helpers["rot_ang"]))
    assert np.abs(got.numpy() - helpers["rot_along_z"]).max() <= 1e-6


def test_bbox_pred_to_bbox(helpers):
    pts, p8 = torch.from_numpy(helpers["dec_pts"]), torch.from_numpy(helpers["dec_pred8"])
    assert np.abs(O.bbox_pred_to_bbox(pts, p8[:, :6]).numpy() - helpers["dec_box6"]).max() <= 1e-6
    assert np.abs(O.bbox_pred_to_bbox(pts, p8).numpy() - helpers["dec_box8"]).max() <= 1e-5


def _oracle_for(gold):
    from cagroup3d_b200 import model_init, synthetic
    ncls, yaw, B = int(gold["n_classes"]), bool(gold["with_yaw"]), int(gold["batch"])
    batch = synthetic.make_batch(B, target_voxels=int(gold["voxels"]), config=int(gold["config"]), n_classes=ncls,
                                 sunrgbd=yaw)
    model = model_init.seeded_model(ncls, yaw, seed=int(gold["seed"]))
    with torch.no_grad():
        model.dense_head.semantic_conv.bias.copy_(torch.from_numpy(gold["semantic_bias"]))
        model.dense_head.cls_conv.bias.copy_(torch.from_numpy(gold["cls_bias"]))
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    # teacher-force the two discontinuities with the reference's own values (compared separately below)
    force = {"sem": torch.from_numpy(gold["sem"]), "offsets": torch.from_numpy(gold["offsets"])}
    return orc.forward(torch.from_numpy(batch["points"]), B, cur_epoch=10, force=force), ncls, B


def match_detections(got, want, tol):
    """one-to-one match of [box(7) | score | label] rows; returns the fraction of `want` matched."""
    if len(want) == 0:
        return 1.0 if len(got) == 0 else 0.0
    if len(got) == 0:
        return 0.0
    d = torch.cdist(torch.as_tensor(got[:, :8]).double(), torch.as_tensor(want[:, :8]).double(), p=float("inf"))
    same = torch.as_tensor(got[:, 8:9]).double() == torch.as_tensor(want[:, 8:9]).double().T
    d = torch.where(same, d, torch.full_like(d, 1e9))
    return (d.min(0).values <= tol).double().mean().item()


@pytest.mark.parametrize("name", ["scannet_small", "sunrgbd_small"])
def test_oracle_forward_vs_reference_python(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    res, ncls, B = _oracle_for(gold)
    # backbone: same stride-2 coordinate rows in the same order, features within 1e-4
    assert (res["bb_coords"] == gold["bb_coords"]).all()
    assert np.abs(res["bb_feats"].numpy() - gold["bb_feats"]).max() <= 1e-4
    assert np.abs(res["head"]["sem"].numpy() - gold["sem"]).max() <= 1e-4
    assert np.abs(res["head"]["offsets"].numpy() - gold["offsets"]).max() <= 1e-4
    # per-class maps: voxel points (exact rows) and raw predictions
    for c in range(ncls):
        m = res["head"]["maps"][c]
        for b in range(B):
            want = gold[f"map_c{c}_b{b}"]
            rows = np.nonzero(m["coords"][:, 0] == b)[0]
            assert len(rows) == len(want), (c, b, len(rows), len(want))
            vs = np.asarray(O.class_voxel_sizes(ncls)[c], dtype=np.float32)
            pts = m["coords"][rows, 1:].astype(np.float32) * vs
            assert np.abs(pts - want[:, :3]).max() <= 1e-6
            got = np.concatenate([m["ctr"][rows].numpy(), m["cls"][rows].numpy(), m["bbox"][rows].numpy()], 1)
            assert np.abs(got - want[:, 3:]).max() <= 2e-4, (c, b, np.abs(got - want[:, 3:]).max())
    for b in range(B):
        bx, sc, lb = res["stage1"][b]
        got = np.concatenate([bx.numpy(), sc.numpy()[:, None], lb.numpy()[:, None].astype(np.float32)], 1)
        want = gold[f"stage1_b{b}"]
        if want.shape[1] == 8:                      # 6-dof boxes: reference keeps (n,6)+score+label
            got = np.concatenate([got[:, :6], got[:, 7:]], 1) if got.shape[1] == 9 else got
            pad = lambda a: np.concatenate([a[:, :6], np.zeros((len(a), 1), np.float32), a[:, 6:]], 1)
            got, want = pad(got), pad(want)
        assert len(got) == len(want)
        assert match_detections(got, want, 1e-4) == 1.0
        fb, fs, fl = res["final"][b]
        gotf = np.concatenate([fb.numpy(), fs.numpy()[:, None], fl.numpy()[:, None].astype(np.float32)], 1)
        wantf = gold[f"final_b{b}"]
        if wantf.shape[1] == 8:
            pad = lambda a: np.concatenate([a[:, :6], np.zeros((len(a), 1), np.float32), a[:, 6:]], 1)
            gotf = pad(np.concatenate([gotf[:, :6], gotf[:, -2:]], 1)) if gotf.shape[1] != 8 else pad(gotf)
            wantf = pad(wantf)
        assert len(gotf) == len(wantf)
        assert match_detections(gotf, wantf, 2e-4) == 1.0
    assert np.abs(res["roi"]["rcnn_reg"].numpy() - gold["rcnn_reg"]).max() <= 2e-4
