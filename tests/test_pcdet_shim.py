"""CPU: the pcdet-shaped boundary -- config parsing of the shipped YAMLs (and, where present, the reference's own),
registries, state-dict naming, checkpoint loading, collate layout."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def load_cfg(root, name):
    from pcdet.config import EasyDict, cfg_from_yaml_file
    cwd = os.getcwd()
    os.chdir(os.path.join(root, "tools"))
    try:
        return cfg_from_yaml_file(f"cfgs/{name}_models/CAGroup3D.yaml", EasyDict())
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("name,ncls,yaw", [("scannet", 18, False), ("sunrgbd", 10, True)])
def test_shipped_yaml_builds_the_model(name, ncls, yaw):
    from pcdet.models import build_network
    from cagroup3d_b200 import model_init
    cfg = load_cfg(ROOT, name)
    assert len(cfg.CLASS_NAMES) == ncls and cfg.MODEL.DENSE_HEAD.WITH_YAW is yaw
    assert cfg.DATA_CONFIG.DATASET in ("ScannetDataset", "SunrgbdDataset")          # _BASE_CONFIG_ merged
    ours = model_init.default_model_cfg(ncls, yaw)
    for sec in ("BACKBONE_3D", "DENSE_HEAD", "ROI_HEAD"):
        for k, v in ours[sec].items():
            assert cfg.MODEL[sec][k] == v, (sec, k)
    if name == "scannet":
        ds = type("D", (), {"class_names": cfg.CLASS_NAMES})()
        model = build_network(cfg.MODEL, ncls, ds)
        assert model.backbone_3d.num_point_features == 64 and model.class_names == cfg.CLASS_NAMES


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not on this machine")
@pytest.mark.parametrize("name", ["scannet", "sunrgbd"])
def test_reference_yaml_parses_to_the_same_model_section(name):
    a, b = load_cfg(ROOT, name), load_cfg(REF, name)
    assert a.CLASS_NAMES == b.CLASS_NAMES
    def same(x, y, path):
        for k, v in y.items():
            assert k in x, path + k
            if isinstance(v, dict):
                same(x[k], v, path + k + ".")
            else:
                assert x[k] == v, (path + k, x[k], v)
    same(a.MODEL, b.MODEL, "MODEL.")
    for k in ("NUM_EPOCHS", "LR", "DECAY_STEP_LIST", "GRAD_NORM_CLIP", "OPTIMIZER", "WEIGHT_DECAY", "LR_DECAY"):
        assert a.OPTIMIZATION[k] == b.OPTIMIZATION[k]


def test_cfg_from_list_overrides():
    from pcdet.config import cfg_from_list
    cfg = load_cfg(ROOT, "scannet")
    cfg_from_list(["MODEL.DENSE_HEAD.NMS_CONFIG.IOU_THR", "0.25", "MODEL.ROI_HEAD.MIDDLE_FEATURE_SOURCE", "2",
                   "DATA_CONFIG.DATA_SPLIT", "test:train"], cfg)
    assert cfg.MODEL.DENSE_HEAD.NMS_CONFIG.IOU_THR == 0.25 and cfg.MODEL.ROI_HEAD.MIDDLE_FEATURE_SOURCE == [2]
    assert cfg.DATA_CONFIG.DATA_SPLIT.test == "train"
    with pytest.raises(AssertionError):
        cfg_from_list(["MODEL.NOPE", "1"], cfg)


def test_state_dict_names_and_shapes_appendix_c():
    from cagroup3d_b200 import model_init
    sd = model_init.seeded_model(18, False, seed=0).state_dict()
    want = {
        "backbone_3d.conv1.0.kernel": (27, 3, 64), "backbone_3d.conv1.1.bn.running_var": (64,),
        "backbone_3d.layer1.0.downsample.0.kernel": (1, 64, 64), "backbone_3d.compression3.0.kernel": (256, 128),
        "backbone_3d.spp.scale1.3.kernel": (1024, 128), "backbone_3d.out.0.kernel": (8, 256, 256),
        "dense_head.semantic_conv.kernel": (64, 18), "dense_head.semantic_conv.bias": (1, 18),
        "dense_head.feature_offset.0.kernel": (27, 64, 64), "dense_head.scales.17.scale": (),
        "dense_head.cls_individual_out.0.0.kernel": (729, 64, 64), "dense_head.cls_individual_up.3.0.kernel": (27, 64, 64),
        "dense_head.cls_individual_up.3.1.0.bn.weight": (64,), "dense_head.cls_individual_fuse.5.0.kernel": (128, 64),
        "dense_head.cls_individual_expand_out.9.0.kernel": (125, 64, 64), "dense_head.cls_conv.bias": (1, 18),
        "roi_head.roi_grid_pool_layers.0.grid_conv.kernel": (125, 64, 128),
        "roi_head.roi_grid_pool_layers.0.pooling_conv.kernel": (343, 128, 128),
        "roi_head.roi_grid_pool_layers.0.pooling_bn.bn.running_mean": (128,), "roi_head.reg_fc_layers.0.weight": (256, 128),
        "roi_head.reg_fc_layers.5.running_var": (256,), "roi_head.reg_pred_layer.weight": (6, 256), "global_step": (1,),
    }
    for k, shp in want.items():
        assert k in sd and tuple(sd[k].shape) == shp, (k, tuple(sd[k].shape) if k in sd else None)
    n = sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k)
    assert 120e6 < n < 130e6                                  # ~126.5 M parameters (ScanNet)


def test_load_params_from_file_matches_by_key_and_shape(tmp_path):
    from cagroup3d_b200 import model_init
    a = model_init.seeded_model(10, True, seed=1)
    b = model_init.seeded_model(10, True, seed=2)
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    sd["dense_head.cls_conv.kernel"] = torch.zeros((64, 3))           # wrong shape -> skipped, reported
    sd["not.a.key"] = torch.zeros(1)
    path = tmp_path / "checkpoint_epoch_12.pth"
    torch.save({"model_state": sd, "epoch": 12}, path)
    missing = b.load_params_from_file(str(path), to_cpu=True)
    assert missing == ["dense_head.cls_conv.kernel"]
    assert torch.equal(b.state_dict()["backbone_3d.conv1.0.kernel"], a.state_dict()["backbone_3d.conv1.0.kernel"])


def test_collate_layout_and_prediction_dicts():
    from pcdet.datasets import build_dataloader
    cfg = load_cfg(ROOT, "scannet")
    cfg.DATA_CONFIG.SYNTHETIC.VOXELS = 1500
    cfg.DATA_CONFIG.SYNTHETIC.NUM_SCENES = 3
    ds, loader, sampler = build_dataloader(cfg.DATA_CONFIG, cfg.CLASS_NAMES, batch_size=2, dist=False, workers=0, training=False)
    batch = next(iter(loader))
    p = batch["points"]
    assert p.shape[1] == 7 and set(np.unique(p[:, 0])) == {0.0, 1.0} and batch["batch_size"] == 2
    assert p[:, 4:].max() > 1.5 and p.dtype == np.float32                      # colours still 0..255 (divided in forward)
    pred = [{"pred_boxes": torch.rand(4, 7), "pred_scores": torch.rand(4), "pred_labels": torch.tensor([0, 3, 17, 2])},
            {"pred_boxes": torch.zeros(0, 7), "pred_scores": torch.zeros(0), "pred_labels": torch.zeros(0, dtype=torch.long)}]
    annos = ds.generate_prediction_dicts(batch, pred, cfg.CLASS_NAMES)
    assert list(annos[0]["name"]) == ["cabinet", "sofa", "garbagebin", "chair"] and annos[0]["boxes_3d"].shape == (4, 7)
    assert len(annos[1]["name"]) == 0 and annos[1]["frame_id"] == 1


def test_forward_refuses_cpu_tensors_in_both_modes():
    from cagroup3d_b200 import model_init
    m = model_init.seeded_model(18, False)
    with pytest.raises(AssertionError):
        m({"points": torch.zeros((10, 7)), "batch_size": 1, "cur_epoch": 10})     # CPU tensor: no CPU fallback
    m.train()
    with pytest.raises(AssertionError):                                            # training mode: CUDA tensors only as well
        m({"points": torch.zeros((10, 7)), "batch_size": 1, "cur_epoch": 10, "gt_boxes": torch.zeros((1, 1, 8))})


def test_training_loader_items_carry_masks():
    """build_dataloader(training=True): batches in the layout CAGroup3D.get_training_loss reads (cagroup3d.py:99-135) --
    zero-padded gt_boxes (B, M, 8) and per-sample semantic / instance mask lists aligned with the points."""
    from pcdet.datasets import build_dataloader
    cfg = load_cfg(ROOT, "scannet")
    cfg.DATA_CONFIG["SYNTHETIC"] = {"NUM_SCENES": 3, "VOXELS": 800}
    ds, loader, _ = build_dataloader(cfg.DATA_CONFIG, cfg.CLASS_NAMES, batch_size=2, dist=False, workers=0, training=True)
    b = next(iter(loader))
    assert b["batch_size"] == 2 and b["gt_boxes"].shape[0] == 2 and b["gt_boxes"].shape[2] == 8
    assert len(b["semantic_mask"]) == len(b["instance_mask"]) == 2
    for i in range(2):
        n = int((b["points"][:, 0] == i).sum())
        assert len(b["semantic_mask"][i]) == len(b["instance_mask"][i]) == n
        fg = b["semantic_mask"][i] < len(cfg.CLASS_NAMES)
        assert fg.any() and (b["instance_mask"][i][fg] >= 5).all() and (b["instance_mask"][i][~fg] < 5).all()


def test_mode_switch_drops_folded_parameter_copies():
    """model.eval() after training steps must not evaluate with folded BatchNorms / stacked class weights made before the
    last optimizer step."""
    from cagroup3d_b200 import model_init
    m = model_init.seeded_model(18, False)
    for mod in (m.backbone_3d, m.dense_head, m.roi_head):
        mod.fold.get("probe", lambda: 1)
        assert mod.fold._c
    m.train()
    assert not m.backbone_3d.fold._c and not m.dense_head.fold._c and not m.roi_head.fold._c
    m.dense_head.fold.get("probe", lambda: 1)
    m.eval()
    assert not m.dense_head.fold._c
