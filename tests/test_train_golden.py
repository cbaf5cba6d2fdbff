"""The reference's TRAINING step on two synthetic ScanNet-shaped scenes (tests/golden/scannet_train_small.npz, made by
tests/golden/make_train_golden.py running the reference's own Python + autograd on the CPU): the fixture that pins the next
coverage row (losses / assigner / gradients).  Until the CUDA training path exists these tests keep
the fixture and its inputs honest: the synthetic masks the generator feeds, and the fixture's internal consistency."""
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scannet_train_small.npz"), allow_pickle=True)


def test_fixture_is_consistent():
    parts = sum(float(G["tb_" + k]) for k in ("loss_centerness", "loss_bbox", "loss_cls", "loss_sem", "loss_vote"))
    assert abs(parts - float(G["tb_one_stage_loss"])) <= 1e-4 * parts
    assert abs(float(G["tb_one_stage_loss"]) + float(G["tb_loss_two_stage"]) - float(G["loss"])) <= 1e-4 * float(G["loss"])
    norms, names = G["grad_norms"], G["grad_names"]
    assert len(names) == len(norms) == 432 and np.isfinite(norms).all() and (norms >= 0).all()      # every parameter got a gradient
    by = dict(zip(names.tolist(), norms.tolist()))
    assert by["backbone_3d.conv1.0.kernel"] > 0 and by["dense_head.semantic_conv.kernel"] > 0
    assert by["dense_head.cls_individual_out.0.0.kernel"] > 0                                      # the per-class 9^3 convs train
    assert G["grad__dense_head.semantic_conv.bias"].shape[-1] == 18


def test_synthetic_masks_are_what_the_fixture_was_made_from():
    from cagroup3d_b200 import synthetic
    pts, boxes, sem, ins = synthetic.make_scene(1000 * int(G["config"]), int(G["voxels"]), n_classes=int(G["n_classes"]), return_masks=True)
    p2, b2 = synthetic.make_scene(1000 * int(G["config"]), int(G["voxels"]), n_classes=int(G["n_classes"]))
    assert np.array_equal(pts, p2) and np.array_equal(boxes, b2)               # masks do not perturb the random stream
    assert sem.shape == ins.shape == (len(pts),) and ins.min() == 0 and ins.max() == 16
    on_box = ins >= 5
    assert (sem[~on_box] == int(G["n_classes"])).all()                         # floor / walls: background class
    assert np.array_equal(sem[on_box], boxes[ins[on_box] - 5, 7].astype(np.int64))
    for i in (0, 5, 11):                                                       # points of box i lie on its surface (4 mm noise)
        q = pts[ins == 5 + i, :3] - boxes[i, :3]
        assert (np.abs(q) <= boxes[i, 3:6] / 2 + 0.03).all()
