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
