"""Inference stages on the CPU: the product's own host code (cagroup3d_b200/head.py) driving the product's own CUDA sources
compiled for the CPU (tests/cuda_on_cpu, routed by tests/cabi_emulator.install(compiled=True): raw pointers, sizes and strides
exactly as _lib.call hands them to the GPU library), teacher-forced with the ORACLE's class maps like the GPU test
tests/test_gpu_model.py::test_stage1_proposals_teacher_forced.  What it adds to the `-m "not gpu"` suite: the stage-1 proposal
stage -- decode, per-(sample, class-map) top-k, per-(sample, class) score threshold, sort, NMS (axis-aligned and rotated),
packing -- is checked against the oracle end to end without a GPU, 20 kernels and the Python between them; with
CG3D_SLOW_TESTS=1 so is the whole RoI stage (hash-unique grid voxels, tap-pattern-ordered rule map, 5^3 conv, 7^3 pooling
contraction, MLP, decode, NMS)."""
import numpy as np
import pytest
import torch

from oracle import cagroup3d_oracle as O
from tests import cabi_emulator as E

TOL = 1e-3


@pytest.mark.parametrize("ncls,yaw", [pytest.param(18, False, marks=pytest.mark.slow), (10, True)], ids=["scannet18", "sunrgbd10"])
def test_stage1_proposals_through_the_cuda_sources_on_cpu(monkeypatch, ncls, yaw):
    from cagroup3d_b200 import model_init, synthetic
    E.install(monkeypatch, compiled=True)
    B = 2
    batch = synthetic.make_batch(B, target_voxels=700, n_classes=ncls, sunrgbd=yaw, config=7)
    model = model_init.seeded_model(ncls, yaw, seed=3)
    pts = torch.from_numpy(batch["points"])
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    model_init.calibrate_semantic_bias(model, orc.forward(pts, B, stages="backbone")["bb_feats"], 0.10)
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    res = orc.forward(pts, B, cur_epoch=10, stages="head")
    pred = torch.cat([torch.cat([m["ctr"], m["cls"], m["reg"]], 1) for m in res["head"]["maps"]])
    model_init.calibrate_cls_bias(model, pred, 0.05)             # enough boxes above the score threshold for the NMS to work on
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    res = orc.forward(pts, B, cur_epoch=10)
    coords, rows = [], []
    for c, m in enumerate(res["head"]["maps"]):                  # class maps -> the `cm` dict `proposals` consumes
        cc = m["coords"].copy()
        cc[:, 0] += c * B
        coords.append(cc)
        rows.append(torch.cat([m["ctr"], m["cls"], m["reg"]], 1))
    cm = dict(coords=torch.from_numpy(np.concatenate(coords).astype(np.int32)), pred=torch.cat(rows).float().contiguous())
    db, ds, dl, off, _ = model.dense_head.proposals(cm, B)
    assert off[-1] > 20
    for b in range(B):
        wb, ws, wl = res["stage1"][b]
        gb, gs, gl = db[off[b]:off[b + 1]], ds[off[b]:off[b + 1]], dl[off[b]:off[b + 1]]
        assert len(gb) == len(wb), (len(gb), len(wb))
        assert torch.equal(gl.long(), wl.long())
        assert (gs - ws.float()).abs().max().item() <= TOL and (gb - wb.float()).abs().max().item() <= TOL


# (one minute each on CPU threads, both pass: CG3D_SLOW_TESTS=1)
@pytest.mark.slow
@pytest.mark.parametrize("ncls,yaw", [(18, False), (10, True)], ids=["scannet18", "sunrgbd10"])
def test_roi_stage_through_the_cuda_sources_on_cpu(monkeypatch, ncls, yaw):
    """roi_head.CAGroup3DRoIHead.run -- pad the detections, 7^3 grid points per RoI -> hash-unique voxels, the 5^3 conv at
    those voxels over a rule map with the tap-pattern tile order, the 7^3 pooling contraction, the regression MLP, the box
    decode and the per-class NMS -- with coordinate maps, rule maps, sorts and the exact-fp32 conv all served by the product's
    CUDA sources on the CPU, against the oracle's RoI head on the same (oracle) stage-1 detections and backbone features:
    RoI grid voxels exact, pooled features / regression / decoded boxes <= 1e-3, the same final detections."""
    from cagroup3d_b200 import model_init, synthetic, sparse as S
    from tests.util import assert_same_coord_set
    E.install(monkeypatch, compiled=True, native_maps=True)
    monkeypatch.setitem(S._CONV_IMPL, "name", "simt")
    monkeypatch.setattr(S, "MASK_MIN_ROWS", 256)                 # small scene: still take the tap-pattern tile order
    B = 2
    batch = synthetic.make_batch(B, target_voxels=300, n_classes=ncls, sunrgbd=yaw, config=7)
    model = model_init.seeded_model(ncls, yaw, seed=3)
    pts = torch.from_numpy(batch["points"])
    cfg = O.default_cfg(ncls, yaw)
    orc = O.Oracle(model.state_dict(), cfg)
    bb = orc.forward(pts, B, stages="backbone")
    # two stage-1 detections per sample (jittered ground-truth boxes): the RoI stage's input
    g = torch.Generator().manual_seed(1)
    st1 = []
    for b in range(B):
        gt = torch.from_numpy(batch["gt_boxes"][b][:2, :7]).float()
        bx = gt.clone()
        bx[:, :3] += torch.randn((len(gt), 3), generator=g) * 0.05
        if not yaw:
            bx[:, 6] = 0
        st1.append((bx, torch.rand((len(gt),), generator=g) * 0.5 + 0.3, torch.from_numpy(batch["gt_boxes"][b][:2, 7]).long() % ncls))
    omgr, ocm = O.me.Manager(), O.me.CoordMap(bb["bb_coords"], 2)
    omgr.by_stride[2] = ocm
    want_final, inter_o = orc.roi_head(O.me.SparseTensor(bb["bb_feats"], ocm, omgr), st1, B)
    # product side, everything on CPU tensors
    mgr = S.Manager()
    cmap = S.build_map(torch.from_numpy(np.ascontiguousarray(bb["bb_coords"], dtype=np.int32)), 2, mgr)
    mgr.by_stride[2] = cmap
    sp = S.SparseTensor(bb["bb_feats"].detach().float().contiguous(), cmap, mgr)
    db = torch.cat([x[0] for x in st1]).float().contiguous()
    ds = torch.cat([x[1] for x in st1]).float().contiguous()
    dl = torch.cat([x[2] for x in st1]).int().contiguous()
    off = np.cumsum([0] + [len(x[0]) for x in st1]).tolist()
    fb, fs, fl, foff, inter = model.roi_head.run(sp, db, ds, dl, off, B)
    assert (inter["rois"] - inter_o["rois"].float()).abs().max().item() == 0
    want_gc = inter_o["grid_coords"].copy()
    want_gc[:, 1:] *= 2
    assert (inter["grid_coords"].numpy().astype(np.int64) == want_gc).all()
    assert_same_coord_set(inter["uniq"].numpy(), inter_o["uniq"])
    assert (inter["pooled"] - inter_o["pooled"]).abs().max().item() <= TOL
    assert (inter["rcnn_reg"] - inter_o["rcnn_reg"]).abs().max().item() <= TOL
    assert (inter["decoded"] - inter_o["decoded"].float()).abs().max().item() <= TOL
    for b in range(B):
        wb, ws, wl = want_final[b]
        gb, gs, gl = fb[foff[b]:foff[b + 1]], fs[foff[b]:foff[b + 1]], fl[foff[b]:foff[b + 1]]
        assert len(gb) == len(wb) > 0 and torch.equal(gl.long(), wl.long())
        assert (gs - ws.float()).abs().max().item() <= TOL and (gb - wb.float()).abs().max().item() <= TOL


def test_voxelise_and_coordinate_pyramid_on_cpu(monkeypatch):
    """sparse.quantize + sparse.voxel_pyramid (the stride-1 map and the nine strided maps of BiResNet / DAPPM built with
    DEVICE-side row counts and one size read-back, cg3d_unique_first_dev) over the CUDA sources on the CPU == the oracle's
    ME-CPU maps: the same rows in the same (first-occurrence) order at every stride, each derived from the rows of the level
    the plan names; the first-point colours of the voxels equal the oracle's RANDOM_SUBSAMPLE-made-deterministic features."""
    from cagroup3d_b200 import sparse as S, synthetic
    from oracle import me_cpu as me
    E.install(monkeypatch, compiled=True, native_maps=True)
    batch = synthetic.make_batch(2, target_voxels=1500, n_classes=18, config=7)
    pts = torch.from_numpy(batch["points"]).float().contiguous()
    pts[:, 4:] /= 255.
    n, ld = pts.shape
    coords, err = S.quantize(pts, ld, n, (0.02,) * 3)
    mgr = S.Manager()
    cmap, first, n_err = S.voxel_pyramid(coords, mgr, err=err)
    assert n_err == 0 and sorted(mgr.by_stride) == [1] + [p[0] for p in S.BACKBONE_PYRAMID]
    c = pts[:, :4].clone()
    c[:, 1:] /= 0.02
    ox = me.from_points(c, pts[:, 4:])
    assert np.array_equal(cmap.coords.numpy(), ox.C) and cmap.n == len(ox.C) < n
    F = S.gather_rows(pts, 4, first, cmap.n, ld - 4)
    assert torch.equal(F, ox.F.float())
    omaps = {1: ox.cmap}
    for ts, src in S.BACKBONE_PYRAMID:
        cc = omaps[src].coords.copy()
        cc[:, 1:] = np.floor_divide(cc[:, 1:], ts) * ts
        omaps[ts] = me.CoordMap(me.unique_first(cc)[0], ts)
        got = mgr.by_stride[ts]
        assert got.stride == ts and got.n == len(omaps[ts]) and np.array_equal(got.coords.numpy()[:got.n], omaps[ts].coords), ts
    assert mgr.by_stride[512].n <= mgr.by_stride[64].n <= mgr.by_stride[2].n < cmap.n


@pytest.mark.parametrize("ncls,yaw", [pytest.param(18, False, marks=pytest.mark.slow), (10, True)], ids=["scannet18", "sunrgbd10"])
def test_class_grouping_coordinate_phase_on_cpu(monkeypatch, ncls, yaw):
    """head_train.coordinate_phase -- the coordinate half of the class-aware grouping (cagroup_head.py:227-271: per-class
    threshold selection, voted points clamped to the scene, fused voted + original points quantised at the class's voxel
    size and at 3 x that, hash-unique class maps with the point -> voxel inverse; the same kernels and class-batched layout
    as the inference plan, 1 or 3 votes per seed) over the CUDA sources on the CPU == the oracle's per-class loop: the same
    class voxels in the same order, the same inverse maps, the same per-class row ranges."""
    from cagroup3d_b200 import head_train as HT, model_init, sparse as S, synthetic
    from tests.test_train_wiring_cpu import _oracle_class_artifacts
    E.install(monkeypatch, compiled=True, native_maps=True)
    B = 2
    batch = synthetic.make_batch(B, target_voxels=600, n_classes=ncls, sunrgbd=yaw, config=7)
    model = model_init.seeded_model(ncls, yaw, seed=3)
    pts = torch.from_numpy(batch["points"])
    cfg = O.default_cfg(ncls, yaw)
    orc = O.Oracle(model.state_dict(), cfg)
    model_init.calibrate_semantic_bias(model, orc.forward(pts, B, stages="backbone")["bb_feats"], 0.10)
    orc = O.Oracle(model.state_dict(), cfg)
    res = orc.forward(pts, B, cur_epoch=10, stages="head")
    hi = res["head"]
    head = model.dense_head
    head.semantic_threshold = 0.05
    mgr = S.Manager()
    cmap = S.build_map(torch.from_numpy(np.ascontiguousarray(res["bb_coords"], dtype=np.int32)), 2, mgr)
    mgr.by_stride[2] = cmap
    out = S.SparseTensor(res["bb_feats"].detach().float().contiguous(), cmap, mgr)
    art = HT.coordinate_phase(head, out, hi["sem"].float().contiguous(), hi["offsets"].float().contiguous(), B)
    # the oracle-backed builder the training wiring tests use (itself == the oracle's per-class maps)
    E.install(monkeypatch)                                       # (its coordinate maps are oracle objects)
    want = _oracle_class_artifacts(res, 0.05, B, cfg, S.Manager())
    assert art["offA"] == want["offA"] and art["offE"] == want["offE"] and art["offA"][-1] > 50
    assert np.array_equal(art["mapA"].coords.numpy(), want["mapA"].coords.numpy())
    assert np.array_equal(art["mapE"].coords.numpy(), want["mapE"].coords.numpy())
    assert torch.equal(art["invA"].int(), want["invA"]) and torch.equal(art["invE"].int(), want["invE"])
    assert torch.equal(art["ref"].int(), want["ref"])
    for c, m in enumerate(hi["maps"]):                           # and directly: the oracle's class voxels
        rows = art["mapA"].coords[art["offA"][c]:art["offA"][c + 1]].numpy().astype(np.int64)
        rows[:, 0] -= c * B
        assert np.array_equal(rows, m["coords"]), c
