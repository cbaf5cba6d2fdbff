"""Indoor mAP / mAR evaluation (pcdet/datasets/indoor_eval.py) against the reference's own eval.py
(tests/golden/indoor_eval.json, made by tests/golden/make_eval_golden.py) and known answers."""
import json
import os

import numpy as np
import pytest

from pcdet.datasets import indoor_eval as IE
from tests.golden.make_eval_golden import make_case

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "indoor_eval.json")))


def _close(got, want, tol):
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    for k, w in want.items():
        g = got[k]
        assert (np.isnan(w) and np.isnan(g)) or abs(g - w) <= tol, (k, g, w)


def test_matches_reference_eval_on_heading_free_boxes():
    c = GOLD["scannet"]
    gts, dts = make_case(c["seed"], c["n_scenes"], c["n_cls"], False)
    got = IE.indoor_eval(gts, dts, [0.25, 0.5], {i: f"c{i}" for i in range(6)}, bev_overlap_fn=IE.axis_aligned_bev_overlap)
    _close(got, c["result"], 1e-6)
    assert 0 < got["mAR_0.50"] < got["mAR_0.25"] <= 1 or np.isnan(got["mAR_0.25"])


def test_known_answers():
    rng = np.random.default_rng(0)
    boxes = np.concatenate([rng.uniform(-3, 3, (6, 3)), rng.uniform(0.5, 1, (6, 3)), np.zeros((6, 1))], 1).astype(np.float32)
    gts = [{"gt_num": 6, "gt_boxes_upright_depth": boxes[:, :6], "class": np.arange(6) % 3}]
    perfect = [{"labels_3d": np.arange(6) % 3, "boxes_3d": boxes, "scores_3d": np.linspace(0.9, 0.4, 6).astype(np.float32)}]
    r = IE.indoor_eval(gts, perfect, [0.25, 0.5], {0: "a", 1: "b", 2: "c"}, bev_overlap_fn=IE.axis_aligned_bev_overlap)
    assert r["mAP_0.25"] == r["mAP_0.50"] == r["mAR_0.50"] == 1.0
    shifted = [dict(perfect[0], boxes_3d=boxes + np.array([0.2 * 0.75, 0, 0, 0, 0, 0, 0], np.float32))]
    r = IE.indoor_eval(gts, shifted, [0.25, 0.5], {0: "a", 1: "b", 2: "c"}, bev_overlap_fn=IE.axis_aligned_bev_overlap)
    assert r["mAP_0.25"] == 1.0                               # a 0.15 m shift of >= 0.5 m boxes keeps IoU > 0.25 ...
    none = [{"labels_3d": np.zeros(0, np.int64), "boxes_3d": np.zeros((0, 7), np.float32), "scores_3d": np.zeros(0, np.float32)}]
    r = IE.indoor_eval(gts, none, [0.25], {0: "a", 1: "b", 2: "c"}, bev_overlap_fn=IE.axis_aligned_bev_overlap)
    assert r["mAP_0.25"] == 0.0 and r["mAR_0.25"] == 0.0
    # duplicates of a matched box are false positives: AP stays 1 only if they score lower than every true positive
    dup = [{"labels_3d": np.concatenate([perfect[0]["labels_3d"], [0]]), "boxes_3d": np.concatenate([boxes, boxes[:1]]),
            "scores_3d": np.concatenate([perfect[0]["scores_3d"], [0.95]]).astype(np.float32)}]
    r = IE.indoor_eval(gts, dup, [0.25], {0: "a", 1: "b", 2: "c"}, bev_overlap_fn=IE.axis_aligned_bev_overlap)
    assert r["a_AP_0.25"] < 1.0 and r["b_AP_0.25"] == 1.0
    iou = IE.d3_box_overlap(boxes[:2], boxes[:2], IE.axis_aligned_bev_overlap)
    assert np.allclose(np.diag(iou), 1.0)


def test_synthetic_dataset_evaluation_uses_it():
    from pcdet.datasets import SyntheticIndoorDataset
    ds = SyntheticIndoorDataset({"SYNTHETIC": {"NUM_SCENES": 3, "VOXELS": 2000}}, [f"c{i}" for i in range(18)])
    gts = ds.gt_annos()
    assert len(gts) == 3 and gts[0]["gt_boxes_upright_depth"].shape == (12, 6)
    assert np.array_equal(ds[1]["gt_boxes"][:, :6], gts[1]["gt_boxes_upright_depth"])          # boxes_only == full scene
    dets = [{"frame_id": i, "labels_3d": g["class"], "scores_3d": np.full(12, 0.9, np.float32),
             "boxes_3d": np.concatenate([g["gt_boxes_upright_depth"], np.zeros((12, 1), np.float32)], 1)} for i, g in enumerate(gts)]
    ret, _ = ds.evaluation(dets[::-1], ds.class_names)
    assert ret["mAP_0.50"] == 1.0 and ret["mAR_0.25"] == 1.0


@pytest.mark.gpu
def test_rotated_boxes_on_the_cuda_overlap_op(lib):
    c = GOLD["rotated"]
    gts, dts = make_case(c["seed"], c["n_scenes"], c["n_cls"], True)
    got = IE.indoor_eval(gts, dts, [0.25, 0.5], {i: f"c{i}" for i in range(6)})
    _close(got, c["result"], 1e-4)


def test_rotate_iou_oracle_known_answers():
    """oracle/rotate_iou_oracle.py (restatement of the reference's rotate_iou.py) on answers derivable by hand, and the
    pair on which the clockwise convention of rbbox_to_corners and the counter-clockwise one of iou3d_nms disagree."""
    from oracle import rotate_iou_oracle as R
    sq = np.array([[0, 0, 2, 2, 0]], np.float32)
    assert R.rotate_iou_eval(sq, sq, 2)[0, 0] == 4.0 and R.rotate_iou_eval(sq, sq, -1)[0, 0] == 1.0
    assert abs(R.rotate_iou_eval(sq, np.array([[0.5, 0, 2, 2, 0]], np.float32), -1)[0, 0] - 0.6) < 1e-6
    bar = np.array([[0, 0, 2, 1, 0]], np.float32)
    assert abs(R.rotate_iou_eval(bar, np.array([[0, 0, 2, 1, np.pi / 2]], np.float32), 2)[0, 0] - 1.0) < 1e-5
    A, B = np.array([[0, 0, 2, 1, .5]], np.float32), np.array([[.3, .4, 1.8, 1.1, .3]], np.float32)
    cw = R.rotate_iou_eval(A, B, -1)[0, 0]
    An, Bn = A.copy(), B.copy()
    An[:, 4], Bn[:, 4] = -A[:, 4], -B[:, 4]
    ccw = R.rotate_iou_eval(An, Bn, -1)[0, 0]
    assert abs(cw - 0.331) < 2e-3 and abs(ccw - 0.431) < 2e-3                # distinct centres: the sign matters
    # a heading turns the long axis of a bar at the origin towards -y for x > 0 (clockwise)
    c = R.rbbox_to_corners(np.array([0, 0, 2, 0.2, 0.3], np.float32)).reshape(4, 2)
    assert c[2, 0] > 0 and c[2, 1] < 0.2


@pytest.mark.gpu
def test_cuda_bev_overlap_follows_rotate_iou_convention(lib):
    """the product's BEV intersection == the rotate_iou restatement on random yawed boxes with distinct centres."""
    from oracle import rotate_iou_oracle as R
    rng = np.random.default_rng(3)
    mk = lambda n: np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(0.4, 2.0, (n, 3)), rng.uniform(-1.5, 1.5, (n, 1))], 1).astype(np.float32)
    a, b = mk(40), mk(30)
    got = IE._gpu_bev_overlap(a, b)
    want = R.rotate_iou_eval(a[:, [0, 1, 3, 4, 6]], b[:, [0, 1, 3, 4, 6]], 2)
    an, bn = a.copy(), b.copy()
    an[:, 6], bn[:, 6] = -a[:, 6], -b[:, 6]
    wrong = R.rotate_iou_eval(an[:, [0, 1, 3, 4, 6]], bn[:, [0, 1, 3, 4, 6]], 2)          # the counter-clockwise reading
    err, err_wrong = np.abs(got - want), np.abs(got - wrong)
    print("max / mean |CUDA - rotate_iou|: %.2e / %.2e   against the other sign: %.2e / %.2e" % (err.max(), err.mean(), err_wrong.max(), err_wrong.mean()))
    # the two implementations clip differently at the edges (iou3d_nms tests corners with a 1e-2 margin,
    # iou3d_nms_kernel.cu:35-49), so areas agree to ~1e-2, far below the effect of the sign (tenths of a square metre)
    assert err.max() <= 3e-2 and err.mean() <= 2e-3, (err.max(), err.mean())
    assert err_wrong.mean() > 20 * err.mean() and err_wrong.max() > 0.2
    assert (want > 0.05).sum() > 50
