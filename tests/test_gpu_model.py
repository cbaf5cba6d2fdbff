"""GPU parity of the CAGroup3D stages against the CPU oracle (seeded synthetic scenes, seeded weights).

Stage tests are teacher-forced: the CUDA stage gets the ORACLE's input for that stage, so a discrete
decision (threshold, floor, top-k, NMS) can only differ where the oracle's own value sits within
rounding distance of the decision boundary.  Tolerance: 1e-3 on features / logits / boxes (north_star),
exact on coordinates and index sets."""
import numpy as np
import pytest
import torch

from oracle import cagroup3d_oracle as O
from oracle import me_cpu as me
from tests.util import assert_same_coord_set, sort_rows, to_gpu_sparse

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3


# third case: ONE scene of the bench generator at the BASELINE density (config 2, ~50 k voxels): every stage test below
# then also runs at the size the headline number is quoted on (~40 s of CPU oracle)
@pytest.fixture(scope="module", params=[(18, False, 2500, 2, 7), (10, True, 2500, 2, 7), (18, False, 50000, 1, 2)],
                ids=["scannet18", "sunrgbd10", "scannet18-50k"])
def setup(request, lib):
    from cagroup3d_b200 import model_init, synthetic, sparse as S
    ncls, yaw, voxels, B, config = request.param
    # the test scenes are small: lower the row threshold so that the tap-pattern / coarse-block tile orders of the bench
    # configuration are part of what is compared with the oracle
    old_min = S.MASK_MIN_ROWS
    S.MASK_MIN_ROWS = 256
    request.addfinalizer(lambda: setattr(S, "MASK_MIN_ROWS", old_min))
    batch = synthetic.make_batch(B, target_voxels=voxels, n_classes=ncls, sunrgbd=yaw, config=config)
    model = model_init.seeded_model(ncls, yaw, seed=3)
    pts = torch.from_numpy(batch["points"])
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    bb = orc.forward(pts, B, stages="backbone")
    model_init.calibrate_semantic_bias(model, bb["bb_feats"], 0.08 if voxels < 10000 else 1.0 / ncls)   # bench: p_sel = 1 / n_cls
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    res = orc.forward(pts, B, cur_epoch=10)
    pred = torch.cat([torch.cat([m["ctr"], m["cls"], m["reg"]], 1) for m in res["head"]["maps"]])
    model_init.calibrate_cls_bias(model, pred, 0.02 if voxels < 10000 else 0.002)                        # bench: p_box
    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, yaw))
    res = orc.forward(pts, B, cur_epoch=10)
    return dict(model=model.to(DEV), orc=orc, res=res, pts=pts, B=B, ncls=ncls, yaw=yaw)


def test_backbone_features(setup):
    from cagroup3d_b200.detector import voxelize
    s = setup
    p = s["pts"].clone()
    p[:, -3:] /= 255.
    x = voxelize(p.to(DEV), 0.02)
    out = s["model"].backbone_3d.run(x)
    for ts, oc in s["res"]["maps"].items():
        assert_same_coord_set(x.mgr.by_stride[ts].coords.cpu().numpy(), oc)
    pa, pb = assert_same_coord_set(out.C.cpu().numpy(), s["res"]["bb_coords"])
    d = (out.F.cpu()[pa] - s["res"]["bb_feats"][pb]).abs().max().item()
    assert d <= TOL, d


def test_backbone_two_streams_equals_one_stream(setup, monkeypatch):
    """the two-stream issue order of the backbone branches changes no bit of the result (and repeats exactly)."""
    from cagroup3d_b200 import backbone as BB
    from cagroup3d_b200.detector import voxelize
    s = setup
    p = s["pts"].clone()
    p[:, -3:] /= 255.
    outs = []
    for on, coord in ((True, True), (False, False), (True, False), (False, True), (True, True)):
        monkeypatch.setitem(BB._TWO_STREAMS, "on", on)
        monkeypatch.setitem(BB._COORD_STREAM, "on", coord)          # rule maps / strided maps on their own stream
        out = s["model"].backbone_3d.run(voxelize(p.to(DEV), 0.02))
        torch.cuda.synchronize()
        outs.append((out.C.clone(), out.F.clone()))
    for c, f in outs[1:]:
        assert torch.equal(c, outs[0][0]) and torch.equal(f, outs[0][1])


def _oracle_out(s):
    r = s["res"]
    return to_gpu_sparse(r["bb_coords"], r["bb_feats"], 2)


def test_head_class_maps_teacher_forced(setup):
    s = setup
    head = s["model"].dense_head
    head.semantic_threshold = 0.05
    cm = head.class_maps(_oracle_out(s), s["B"])
    h = s["res"]["head"]
    assert (cm["sem"].cpu() - h["sem"]).abs().max().item() <= TOL
    assert (cm["offsets"].cpu() - h["offsets"]).abs().max().item() <= TOL
    assert (cm["voted"].cpu() - h["voted"]).abs().max().item() <= TOL
    off = cm["class_off"]
    C = cm["coords"].cpu().numpy()
    worst = 0.0
    for c, m in enumerate(h["maps"]):
        assert cm["n_sel"][c] == m["n_sel"], f"class {c}: selection differs"
        rows = C[off[c]:off[c + 1]].copy()
        assert (rows[:, 0] // s["B"] == c).all()
        rows[:, 0] -= c * s["B"]
        pa, pb = assert_same_coord_set(rows, m["coords"])
        want = torch.cat([m["ctr"], m["cls"], m["reg"]], 1)[pb]
        got = cm["pred"].cpu()[off[c]:off[c + 1]][pa]
        worst = max(worst, (got - want).abs().max().item(), (cm["feat"].cpu()[off[c]:off[c + 1]][pa] - m["feat"][pb]).abs().max().item())
    assert worst <= TOL, worst


def _pack_oracle_maps(s):
    """oracle class maps -> the `cm` dict `proposals` consumes (rows class-major, batch index c*B+b)."""
    coords, pred = [], []
    for c, m in enumerate(s["res"]["head"]["maps"]):
        cc = m["coords"].copy()
        cc[:, 0] += c * s["B"]
        coords.append(cc)
        pred.append(torch.cat([m["ctr"], m["cls"], m["reg"]], 1))
    return dict(coords=torch.from_numpy(np.concatenate(coords).astype(np.int32)).to(DEV),
                pred=torch.cat(pred).float().contiguous().to(DEV))


def _cmp_dets(got, want, tol=TOL):
    gb, gs, gl = got
    wb, ws, wl = want
    assert len(gb) == len(wb), (len(gb), len(wb))
    if len(wb) == 0:
        return
    assert torch.equal(gl.cpu().long(), wl.long())
    assert (gs.cpu() - ws.float()).abs().max().item() <= tol
    assert (gb.cpu() - wb.float()).abs().max().item() <= tol


def test_stage1_proposals_teacher_forced(setup):
    s = setup
    head = s["model"].dense_head
    db, ds, dl, off, _ = head.proposals(_pack_oracle_maps(s), s["B"])
    assert off[-1] > 0
    for b in range(s["B"]):
        _cmp_dets((db[off[b]:off[b + 1]], ds[off[b]:off[b + 1]], dl[off[b]:off[b + 1]]), s["res"]["stage1"][b])


def test_roi_head_teacher_forced(setup):
    s = setup
    roi = s["model"].roi_head
    st1 = s["res"]["stage1"]
    db = torch.cat([x[0] for x in st1]).float().contiguous().to(DEV)
    ds = torch.cat([x[1] for x in st1]).float().contiguous().to(DEV)
    dl = torch.cat([x[2] for x in st1]).int().contiguous().to(DEV)
    off = np.cumsum([0] + [len(x[0]) for x in st1]).tolist()
    fb, fs, fl, foff, inter = roi.run(_oracle_out(s), db, ds, dl, off, s["B"])
    r = s["res"]["roi"]
    assert (inter["rois"].cpu() - r["rois"].float()).abs().max().item() == 0
    gq = inter["grid_coords"].cpu().numpy().astype(np.int64)
    want = r["grid_coords"].copy()
    want[:, 1:] *= 2
    assert (gq == want).all()                                           # RoI grid voxels, bit exact
    assert_same_coord_set(inter["uniq"].cpu().numpy(), r["uniq"])
    assert (inter["pooled"].cpu() - r["pooled"]).abs().max().item() <= TOL
    assert (inter["rcnn_reg"].cpu() - r["rcnn_reg"]).abs().max().item() <= TOL
    assert (inter["decoded"].cpu() - r["decoded"].float()).abs().max().item() <= TOL
    for b in range(s["B"]):
        _cmp_dets((fb[foff[b]:foff[b + 1]], fs[foff[b]:foff[b + 1]], fl[foff[b]:foff[b + 1]]), s["res"]["final"][b])


def _match(pred, want, tag):
    """one-to-one matching of detection lists: fraction of the oracle's (box | score) rows reproduced within TOL."""
    gb, gs = pred["pred_boxes"].cpu(), pred["pred_scores"].cpu()
    wb, ws, _ = want
    if len(wb) == 0:
        return 1.0, len(gb), 0
    d = torch.cdist(torch.cat([gb, gs[:, None]], 1), torch.cat([wb.float(), ws.float()[:, None]], 1), p=float("inf"))
    near = d.min(0).values
    matched = (near <= TOL).float().mean().item()
    print("%s: %d detections (oracle %d), matched within %g: %.4f, max |delta| on the matched ones %.2e, worst %.2e"
          % (tag, len(gb), len(wb), TOL, matched, float(near[near <= TOL].max()) if matched > 0 else float("nan"), float(near.max())))
    return matched, len(gb), len(wb)


def test_end_to_end_forward(setup):
    """Whole forward through the pcdet-style API; every stage runs on its own (CUDA) inputs.

    Two comparisons, both printed.  FREE-RUNNING against the oracle's own forward: the two fp32 implementations differ in
    the last bits of the vote offsets, and a voted point that sits within that distance of a class-voxel boundary is
    floored into the neighbouring voxel (cagroup_head.py:251) -- which moves the few detections fed by that voxel by
    0.05-0.3 m.  First hardware run: 97.2 % ... 100 % of the detections matched within 1e-3 (profiles/r2_gpu_parity_e2e.log).
    TEACHER-FORCED at the two discontinuities: the oracle re-run with the CUDA path's semantic logits and vote offsets
    (its `force` argument: threshold select and floor() then take the same decisions) must reproduce the detections."""
    s = setup
    pts = s["pts"].clone().to(DEV)
    model = s["model"]
    pred_dicts, recall = model({"points": pts, "batch_size": s["B"], "cur_epoch": 10})
    assert torch.allclose(pts[:, -3:].cpu(), s["pts"][:, -3:] / 255.)    # the in-place /255 of the reference
    assert len(pred_dicts) == s["B"] and "gt" in recall
    for b in range(s["B"]):
        gl, gb = pred_dicts[b]["pred_labels"], pred_dicts[b]["pred_boxes"]
        assert gl.dtype == torch.int64 and gb.shape[1] == 7
        m, ng, nw = _match(pred_dicts[b], s["res"]["final"][b], "free-running sample %d" % b)
        assert abs(ng - nw) <= max(2, nw // 50), (ng, nw)
        assert m >= 0.95, m
    # the oracle forced with the CUDA path's values at the two discontinuities
    from cagroup3d_b200.detector import voxelize
    p = s["pts"].clone().to(DEV)
    p[:, -3:] /= 255.
    out = model.backbone_3d.run(voxelize(p, 0.02))
    model.dense_head.semantic_threshold = 0.05
    cm = model.dense_head.class_maps(out, s["B"])
    pa, pb = assert_same_coord_set(out.C.cpu().numpy(), s["res"]["bb_coords"])
    n = len(pa)
    sem, offs = torch.empty((n, cm["sem"].shape[1])), torch.empty((n, cm["offsets"].shape[1]))
    sem[pb], offs[pb] = cm["sem"].cpu()[pa], cm["offsets"].cpu()[pa]
    forced = s["orc"].forward(s["pts"], s["B"], cur_epoch=10, force={"sem": sem, "offsets": offs})
    for b in range(s["B"]):
        m, ng, nw = _match(pred_dicts[b], forced["final"][b], "teacher-forced sample %d" % b)
        assert abs(ng - nw) <= 1, (ng, nw)
        assert m >= 0.995, m
