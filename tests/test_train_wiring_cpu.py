"""Host logic of the TRAINING path on the CPU: cagroup3d_b200/autograd.py, backbone_train.py and train_targets.py run
against tests/cabi_emulator.py (a torch restatement of each C-ABI entry point's contract, arguments bound by the
parameter names of include/cagroup3d_b200.h) and are compared with autograd through the fp64 oracle.  Checks argument
order, shapes, which gradient goes where and the layer wiring -- not the CUDA code (that is tests/test_zz_gpu_*.py)."""
import numpy as np
import pytest
import torch

from oracle import cagroup3d_oracle as O
from oracle import me_cpu as me
from oracle import train_oracle as T
from tests import cabi_emulator as E


def _rel(got, want):
    return float((got.double() - want).norm()) / (float(want.norm()) + 1e-30)


@pytest.mark.parametrize("compiled", [False, True])
@pytest.mark.parametrize("cin,cout,k,stride", [(16, 24, 3, 1), (16, 8, 3, 2), (8, 8, 1, 1), (8, 12, 1, 2)])
def test_conv_autograd_wiring(monkeypatch, cin, cout, k, stride, compiled):
    from cagroup3d_b200 import autograd as A, sparse as S
    E.install(monkeypatch, compiled=compiled)
    rng = np.random.default_rng(k * 10 + stride)
    c = me.unique_first(np.concatenate([rng.integers(0, 2, (900, 1)), rng.integers(-12, 12, (900, 3))], 1))[0]
    ox = me.SparseTensor(torch.from_numpy(rng.standard_normal((len(c), cin))), me.CoordMap(c, 1), me.Manager())
    ox.mgr.by_stride[1] = ox.cmap
    Wd = torch.from_numpy(rng.standard_normal((k ** 3, cin, cout)) / 5).requires_grad_(True)
    Xd = ox.F.clone().requires_grad_(True)
    y = me.conv(ox.with_F(Xd), Wd[0] if (k == 1 and stride == 1) else Wd, k, stride)
    dY = torch.from_numpy(rng.standard_normal(tuple(y.F.shape)))
    (y.F * dY).sum().backward()
    mgr = S.Manager()
    cm = E.cpu_map(c, 1, mgr)
    mgr.by_stride[1] = cm
    X = ox.F.float().requires_grad_(True)
    W = (Wd.detach()[0] if (k == 1 and stride == 1) else Wd.detach()).float().contiguous().requires_grad_(True)
    out = A.conv(S.SparseTensor(X, cm, mgr), W, k, stride, impl="simt")
    assert np.array_equal(out.C.numpy(), y.C)
    assert _rel(out.F.detach(), y.F.detach()) < 1e-5
    out.F.backward(dY.float())
    assert _rel(X.grad, Xd.grad) < 1e-5 and _rel(W.grad.reshape(Wd.shape), Wd.grad) < 1e-5


def test_backbone_training_wiring_vs_oracle(monkeypatch):
    """run_train on the emulated C ABI == the oracle's training-mode BiResNet: features and the gradient of all 168
    backbone parameters (fp32 emulation vs fp64 oracle)."""
    from cagroup3d_b200 import backbone_train as BT, model_init, sparse as S, synthetic
    calls = E.install(monkeypatch)
    B = 2
    batch = synthetic.make_batch(B, target_voxels=700, config=11)
    model = model_init.seeded_model(18, False, seed=2)
    pts = torch.from_numpy(batch["points"])
    orc = O.Oracle(model.state_dict(), O.default_cfg(18, False), dtype=torch.float64)
    names = [k for k in orc.p if k.startswith("backbone_3d.") and k.endswith(("kernel", "bn.weight", "bn.bias"))]
    for k in names:
        orc.p[k] = orc.p[k].double().requires_grad_(True)
    orc.train_bn = True
    res = orc.forward(pts, B, stages="backbone")
    dY = torch.randn(tuple(res["bb_feats"].shape), generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    (res["bb_feats"] * dY).sum().backward()

    bb = model.backbone_3d.train()
    mgr = S.Manager()
    cm = E.cpu_map(res["vox_coords"], 1, mgr)
    mgr.by_stride[1] = cm
    out = BT.run_train(bb, S.SparseTensor(res["vox_feats"].float().contiguous(), cm, mgr), impl="simt")
    assert np.array_equal(out.C.numpy(), res["bb_coords"])
    assert (out.F.detach().double() - res["bb_feats"].detach()).abs().max().item() < 1e-3
    out.F.backward(dY.float())
    params = dict(model.named_parameters())
    G = max(float(orc.p[k].grad.norm()) for k in names)
    worst = max((float((params[k].grad.double().reshape(orc.p[k].grad.shape) - orc.p[k].grad).norm())
                 / (float(orc.p[k].grad.norm()) + 1e-5 * G), k) for k in names)
    # bound derived from the measured forward error: ReLU-mask flips, see the GPU twin of this test
    # (tests/test_zz_gpu_spconv_backward.py::test_backbone_training_forward_backward_vs_oracle)
    ref = res["bb_feats"].detach()
    eps_f = float(((out.F.detach().double() - ref).abs() / ref.std(0, keepdim=True))[ref > 0].median())
    bound = 2.0 * float(np.sqrt(48 * 0.8 * eps_f))
    print("eps_f %.3e bound %.3e worst %.3e %s" % (eps_f, bound, worst[0], worst[1]))
    assert eps_f < 2e-6 and worst[0] < bound, (worst, eps_f, bound)
    assert int(bb.conv1[1].bn.num_batches_tracked) == 1
    used = set(calls)
    assert {"cg3d_spconv_simt", "cg3d_spconv_wgrad", "cg3d_table_transpose", "cg3d_transpose_weights", "cg3d_bn_train_stats",
            "cg3d_bn_train_backward", "cg3d_affine_act", "cg3d_interp_trilinear", "cg3d_interp_trilinear_backward",
            "cg3d_avgpool_window", "cg3d_avgpool_window_backward", "cg3d_act_backward"} <= used, used


@pytest.mark.parametrize("compiled", [False, True])
def test_targets_and_losses_wiring(monkeypatch, compiled):
    from cagroup3d_b200 import train_targets as TT
    with pytest.raises(RuntimeError, match="no CPU"):
        TT.FocalLoss()(torch.zeros((2, 3)), torch.zeros((2,), dtype=torch.long), avg_factor=1.0)
    E.install(monkeypatch, compiled=compiled)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    g = torch.Generator().manual_seed(3)
    ncls, m = 4, 9
    boxes = torch.cat([(torch.rand((m, 3), generator=g) - 0.5) * 4, torch.rand((m, 3), generator=g) * 1.5 + 0.3, torch.zeros((m, 1))], 1)
    labels = torch.randint(0, ncls - 1, (m,), generator=g)
    pts = [(torch.rand((int(n), 3), generator=g) - 0.5) * 5 for n in (60, 200, 90, 33)]
    ct, bt, lb = T.assign(pts, boxes, labels, 18)
    gc, gb, gl = TT.CAGroup3DAssigner({"TOPK": 18}).assign(pts, boxes, labels)
    assert torch.equal(gl, lb) and torch.equal(gb, bt) and gc.shape == ct.shape
    sl, il = TT.CAGroup3DAssigner.assign_semantic(torch.cat(pts), boxes, labels, ncls)
    assert torch.equal(sl, T.assign_semantic(torch.cat(pts), boxes, labels)[0])
    N = len(lb)
    pred = torch.randn((N, ncls), generator=g).requires_grad_(True)
    loss = TT.FocalLoss()(pred, lb, avg_factor=7.0)
    (3.0 * loss).backward()
    pd = pred.detach().double().requires_grad_(True)
    want = T.focal_loss(pd, lb, 7.0)
    want.backward()
    assert abs(float(loss) - float(want)) < 1e-5 and _rel(pred.grad, 3.0 * pd.grad) < 1e-5
    # box loss through a column slice of a wider prediction (the head passes decoded boxes with a yaw column)
    P = 50
    tgt = torch.cat([torch.randn((P, 3), generator=g), torch.rand((P, 3), generator=g) + 0.3, torch.zeros((P, 1))], 1)
    pb = (tgt + torch.randn((P, 7), generator=g) * 0.2)
    pb[:, 3:6] = pb[:, 3:6].abs() + 0.05
    pb.requires_grad_(True)
    w = torch.rand((P,), generator=g)
    lbx = TT.IoU3DLoss(with_yaw=False)(pb, tgt, weight=w, avg_factor=float(w.sum()))
    lbx.backward()
    pdb = pb.detach().double().requires_grad_(True)
    T.axis_aligned_iou_loss(pdb[:, :6], tgt[:, :6].double(), w.double(), float(w.sum())).backward()
    assert _rel(pb.grad, pdb.grad) < 1e-5 and float(pb.grad[:, 6].abs().max()) == 0
    x = torch.randn((P, 1), generator=g, requires_grad=True)
    t = torch.rand((P, 1), generator=g)
    TT.CrossEntropy(use_sigmoid=True)(x, t, avg_factor=11.0).backward()
    xd = x.detach().double().requires_grad_(True)
    T.bce_loss(xd, t.double(), 11.0).backward()
    assert _rel(x.grad, xd.grad) < 1e-5
    p3, t3, w3 = torch.randn((P, 3), generator=g, requires_grad=True), torch.randn((P, 3), generator=g), torch.rand((P, 3), generator=g)
    TT.SmoothL1Loss(beta=0.04, reduction="sum")(p3, t3, weight=w3).backward()
    pd3 = p3.detach().double().requires_grad_(True)
    T.smooth_l1_sum(pd3, t3.double(), w3.double()).backward()
    assert _rel(p3.grad, pd3.grad) < 1e-5


@pytest.mark.parametrize("compiled", [False, True])
def test_first_stage_loss_single_wiring(monkeypatch, compiled):
    """train_targets.FirstStageLoss.loss_single (the reference's _loss_single argument list) == the oracle's five terms,
    with gradients reaching every prediction tensor."""
    from cagroup3d_b200 import ops, train_targets as TT
    E.install(monkeypatch, compiled=compiled)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    g = torch.Generator().manual_seed(12)
    ncls, m = 5, 8
    boxes = torch.cat([(torch.rand((m, 3), generator=g) - 0.5) * 4, torch.rand((m, 3), generator=g) * 1.5 + 0.4, torch.zeros((m, 1))], 1)
    gtl = torch.randint(0, ncls, (m,), generator=g)
    # scene points: some on every box (instance 5 + j, class of the box), some background (instances 0..2, class ncls)
    sp, sem, ins = [], [], []
    for j in range(m):
        q = boxes[j, :3] + (torch.rand((150, 3), generator=g) - 0.5) * boxes[j, 3:6]
        sp.append(q); sem.append(torch.full((150,), int(gtl[j]))); ins.append(torch.full((150,), 5 + j))
    q = (torch.rand((600, 3), generator=g) - 0.5) * 6
    sp.append(q); sem.append(torch.full((600,), ncls)); ins.append(torch.randint(0, 3, (600,), generator=g))
    sp, sem, ins = torch.cat(sp), torch.cat(sem), torch.cat(ins)
    vox = sp[torch.randperm(len(sp), generator=g)[:700]] + 0.01
    pts = [torch.cat([boxes[gtl == c][:, :3].repeat(12, 1) + (torch.rand((12 * int((gtl == c).sum()), 3), generator=g) - 0.5) * 0.6,
                      (torch.rand((40, 3), generator=g) - 0.5) * 6]) for c in range(ncls)]
    N = sum(len(p) for p in pts)
    mk = lambda *shape: (torch.randn(shape, generator=g) * 0.5).requires_grad_(True)
    ctr, cls, sems, offs = mk(N, 1), mk(N, ncls), mk(len(vox), ncls), (torch.randn((len(vox), 3), generator=g) * 0.1).requires_grad_(True)
    reg = (torch.rand((N, 6), generator=g) * 0.6 + 0.1).requires_grad_(True)
    sizes = [len(p) for p in pts]
    split = lambda t: list(torch.split(t, sizes))
    L = TT.FirstStageLoss(ncls)
    got = L.loss_single(split(ctr), split(reg), split(cls), pts, offs, vox, sems, vox, None, boxes, gtl, sp, sem, ins)
    sum(got).backward()
    # oracle
    ct_t, box_t, lab = T.assign(pts, boxes, gtl, 18)
    assert (lab >= 0).sum() > 10
    sem_l, _ = T.assign_semantic(vox, boxes, gtl)
    off_t, off_m = T.vote_targets_from_masks(sp, vox, boxes, sem, ins, ncls)
    d = {k: v.detach().double().requires_grad_(True) for k, v in dict(ctr=ctr, reg=reg, cls=cls, sems=sems, offs=offs).items()}
    want = T.head_loss_terms(d["ctr"], O.bbox_pred_to_bbox(torch.cat(pts).double(), d["reg"]), d["cls"], ct_t.double(), box_t.double(), lab,
                             d["sems"], sem_l, d["offs"], off_t.double(), off_m.double())
    sum(want).backward()
    for a, b in zip(got, want):
        assert abs(float(a.detach()) - float(b.detach())) < 2e-5 * max(1.0, abs(float(b.detach())))
    for k, t in dict(ctr=ctr, reg=reg, cls=cls, sems=sems, offs=offs).items():
        assert t.grad is not None and _rel(t.grad, d[k].grad) < 1e-4, k


@pytest.mark.slow       # (40 s; the semantic + vote terms it checks are two of the five of test_first_stage_loss_wiring_vs_training_oracle)
def test_partial_training_step_wiring_vs_oracle(monkeypatch):
    """backbone (training mode) -> shared part of the head -> semantic + vote loss -> backward, on the emulated C ABI,
    against the same two terms computed from the oracle's training-mode forward: loss values and the gradient of every
    parameter the two terms reach (backbone, semantic_conv, offset_block)."""
    from cagroup3d_b200 import backbone_train as BT, head_train as HT, model_init, ops, sparse as S, synthetic, train_targets as TT
    E.install(monkeypatch)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    B, ncls = 2, 18
    scenes = [synthetic.make_scene(1000 * 7 + i, 600, n_classes=ncls, return_masks=True) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    pts = torch.from_numpy(batch["points"])
    model = model_init.seeded_model(ncls, False, seed=4)
    with torch.no_grad():
        model.dense_head.semantic_conv.bias.fill_(-1.0)          # some foreground probability, so the focal term has signal
    gtb = [torch.from_numpy(b[:, :7]).float() for _, b, _, _ in scenes]
    gtl = [torch.from_numpy(b[:, 7]).long() for _, b, _, _ in scenes]
    sp = [torch.from_numpy(p[:, :3]).float() for p, _, _, _ in scenes]
    semm = [torch.from_numpy(s) for _, _, s, _ in scenes]
    insm = [torch.from_numpy(m) for _, _, _, m in scenes]

    orc = O.Oracle(model.state_dict(), O.default_cfg(ncls, False), dtype=torch.float64)
    reach = ("backbone_3d.", "dense_head.semantic_conv", "dense_head.offset_block")
    names = [k for k in orc.p if k.startswith(reach) and k.endswith(("kernel", "bn.weight", "bn.bias", "conv.bias"))]
    for k in names:
        orc.p[k] = orc.p[k].double().requires_grad_(True)
    orc.train_bn = True
    res = orc.forward(pts, B, cur_epoch=10, stages="head")
    hi, Cc = res["head"], torch.from_numpy(res["bb_coords"])
    want_sem, want_vote = [], []
    for b in range(B):
        rows = torch.nonzero(Cc[:, 0] == b).squeeze(1)
        vox = Cc[rows, 1:].float() * 0.02
        sl, _ = T.assign_semantic(vox, gtb[b], gtl[b])
        ot, om = T.vote_targets_from_masks(sp[b], vox, gtb[b], semm[b], insm[b], ncls)
        w = (om / torch.ones_like(om).sum() + 1e-6)[:, None].repeat(1, 3).double()
        want_vote.append(T.smooth_l1_sum(hi["offsets"][rows], ot.double(), w))
        want_sem.append(T.focal_loss(hi["sem"][rows], sl, max(float((sl >= 0).sum()), 1.0)))
    want = (torch.stack(want_sem).mean(), torch.stack(want_vote).mean())
    (want[0] + want[1]).backward()

    model.train()
    mgr = S.Manager()
    cm = E.cpu_map(res["vox_coords"], 1, mgr)
    mgr.by_stride[1] = cm
    out = BT.run_train(model.backbone_3d, S.SparseTensor(res["vox_feats"].float().contiguous(), cm, mgr), impl="simt")
    sem, offs, offF = HT.shared_part(model.dense_head, out, impl="simt")
    assert offF.shape == (out.cmap.n, 64)
    got = HT.semantic_and_vote_loss(model.dense_head, out, sem, offs, B, gtb, gtl, sp, semm, insm)
    (got[0] + got[1]).backward()
    for a, b in zip(got, want):
        assert abs(float(a.detach()) - float(b.detach())) < 1e-4 * max(1.0, abs(float(b.detach()))), (a, b)
    params = dict(model.named_parameters())
    G = max(float(orc.p[k].grad.norm()) for k in names)
    assert G > 0
    worst = max((float((params[k].grad.double().reshape(orc.p[k].grad.shape) - orc.p[k].grad).norm())
                 / (float(orc.p[k].grad.norm()) + 1e-4 * G), k) for k in names)
    assert worst[0] < 1e-2, worst


def test_partial_training_step_driver(monkeypatch):
    """train_step.partial_training_step: batch_dict in, gradients reduced, optimizer stepped, loss decreasing over a few
    steps on a fixed batch (emulated C ABI; the voxelisation comes from the oracle)."""
    from cagroup3d_b200 import dist as D, model_init, ops, sparse as S, synthetic, train_step as TS, train_targets as TT
    E.install(monkeypatch)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)

    def voxelize_cpu(points, voxel_size):
        c = points[:, :4].clone()
        c[:, 1:] /= voxel_size
        ox = me.from_points(c, points[:, 4:])
        mgr = S.Manager()
        cm = E.cpu_map(ox.C, 1, mgr)
        mgr.by_stride[1] = cm
        return S.SparseTensor(ox.F.float().contiguous(), cm, mgr)
    monkeypatch.setattr(TS, "voxelize", voxelize_cpu)
    B, ncls = 2, 18
    scenes = [synthetic.make_scene(1000 * 7 + i, 400, n_classes=ncls, return_masks=True) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    model = model_init.seeded_model(ncls, False, seed=4).train()
    params = [p for n, p in model.named_parameters() if n.startswith(("backbone_3d.", "dense_head.semantic_conv", "dense_head.offset_block"))]
    opt = torch.optim.AdamW(params, lr=2e-3)
    red = D.GradientAllReducer(params, bucket_mb=8)
    losses = []
    for _ in range(4):
        bd = {"points": torch.from_numpy(batch["points"]).clone(), "batch_size": B, "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float(),
              "semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
        tb = TS.partial_training_step(model, bd, opt, red, impl="simt")
        assert set(tb) == {"loss", "loss_sem", "loss_vote"} and np.isfinite(tb["loss"])
        losses.append(tb["loss"])
    assert losses[-1] < losses[0], losses
    assert int(model.global_step) == 4
    assert all(p.grad is not None and p.grad.data_ptr() == red._view(p).data_ptr() for p in params)


@pytest.mark.parametrize("compiled", [False, True])
def test_grouped_conv_and_segment_mean_wiring(monkeypatch, compiled):
    """GroupedConvFunction in the class-batched layout of the head (class folded into the batch index, rows class-major,
    one weight group per class) and SegmentMeanFunction, against autograd through per-class oracle convolutions."""
    from cagroup3d_b200 import autograd as A, sparse as S
    E.install(monkeypatch, compiled=compiled)
    rng = np.random.default_rng(8)
    G, Bs, Cin, Cout, k = 3, 2, 8, 12, 3
    coords, off = [], [0]
    for g_ in range(G):
        c = np.concatenate([rng.integers(0, Bs, (150 + 90 * g_, 1)) + g_ * Bs, rng.integers(-5, 5, (150 + 90 * g_, 3))], 1)
        c = me.unique_first(c)[0]
        coords.append(c)
        off.append(off[-1] + len(c))
    coords = np.concatenate(coords)
    n = len(coords)
    Wd = torch.from_numpy(rng.standard_normal((G, k ** 3, Cin, Cout)) / 4).requires_grad_(True)
    # point features -> quantise-average onto the voxels -> grouped conv
    inv = np.concatenate([np.arange(n), rng.integers(0, n, 2 * n)])
    Pd = torch.from_numpy(rng.standard_normal((len(inv), Cin))).requires_grad_(True)
    invt = torch.from_numpy(inv)
    Xd = torch.zeros((n, Cin), dtype=torch.float64).index_add_(0, invt, Pd) / torch.bincount(invt, minlength=n).double()[:, None]
    cm_o = me.CoordMap(coords, 1)
    rules = me.kernel_map(cm_o, coords, k, 1)
    Yd = torch.zeros((n, Cout), dtype=torch.float64)
    for g_ in range(G):
        sel = [(i[(o >= off[g_]) & (o < off[g_ + 1])], o[(o >= off[g_]) & (o < off[g_ + 1])]) for i, o in rules]
        Yd = Yd + me._apply_rules(Xd, Wd[g_], sel, n)
    dY = torch.from_numpy(rng.standard_normal((n, Cout)))
    (Yd * dY).sum().backward()

    mgr = S.Manager()
    cm = E.cpu_map(coords, 1, mgr)
    nbr, order = S.neighbor_table(cm, cm, k, mgr, ordered=True)
    # the head keeps positions class-major: order the emulator's random permutation by class
    cls_of = torch.from_numpy(np.searchsorted(np.array(off), order.numpy(), side="right") - 1)
    perm = torch.argsort(cls_of, stable=True)
    nbr, order = nbr[:, perm].contiguous(), order[perm].contiguous()
    P = Pd.detach().float().requires_grad_(True)
    W = Wd.detach().float().requires_grad_(True)
    X = A.segment_mean(P, invt.int(), n)
    Y = A.grouped_conv(X, W, nbr, order, n, k ** 3, off, off, impl="simt")
    assert _rel(Y.detach(), Yd.detach()) < 1e-5
    Y.backward(dY.float())
    assert _rel(P.grad, Pd.grad) < 1e-5 and _rel(W.grad, Wd.grad) < 1e-5


def _oracle_class_artifacts(res, thr, B, cfg, mgr):
    """the class-batched coordinate artifacts of head_train.coordinate_phase, rebuilt from the oracle's training-mode
    forward exactly as its per-class loop forms them (cagroup3d_oracle.Oracle.head), class folded into the batch index."""
    from cagroup3d_b200 import sparse as S
    hi = res["head"]
    ncls, vs, ex = cfg["n_classes"], cfg["voxel_size"], cfg["expand"]
    Cc = torch.from_numpy(res["bb_coords"])
    Cf = Cc.float()
    pad_id = np.array([np.nonzero(res["bb_coords"][:, 0] == b)[0][0] for b in range(B)], dtype=np.int64)
    sizes = O.class_voxel_sizes(ncls)
    qa_all, qe_all, ref = [], [], []
    for cls in range(ncls):
        s = torch.sigmoid(hi["sem"][:, cls])
        sel = torch.cat([torch.nonzero(s > thr).squeeze(1), torch.from_numpy(pad_id)])
        nv = hi["voted"].shape[1]                              # 3 votes per seed with yaw (SUN RGB-D), else 1
        vc = Cf[sel].view(-1, 1, 4).repeat(1, nv, 1)
        vc[:, :, 1:4] = hi["voted"][sel]
        oc = Cf[sel].clone()
        oc[:, 1:4] *= vs
        fuse_c = torch.cat([vc.reshape(-1, 4), oc], 0)
        vsz = torch.tensor(sizes[cls], dtype=torch.float32)
        qa, qe = fuse_c.clone(), fuse_c.clone()
        qa[:, 1:] = torch.floor(fuse_c[:, 1:] / vsz)
        qe[:, 1:] = torch.floor(fuse_c[:, 1:] / (vsz * ex)) * ex
        qa[:, 0] += cls * B
        qe[:, 0] += cls * B
        qa_all.append(qa.long().numpy()); qe_all.append(qe.long().numpy())
        ref.append(torch.cat([torch.stack([sel.repeat_interleave(nv), torch.arange(nv).repeat(len(sel))], 1),
                              torch.stack([sel, -torch.ones_like(sel)], 1)]))
    qa_all, qe_all, ref = np.concatenate(qa_all), np.concatenate(qe_all), torch.cat(ref).int()
    ua, inva, _ = me.unique_first(qa_all)
    ue, inve, _ = me.unique_first(qe_all)
    mapA, mapE = E.cpu_map(ua, 1, mgr), E.cpu_map(ue, ex, mgr)
    cls_rows = lambda u: [0] + np.cumsum(np.bincount(u[:, 0] // B, minlength=ncls)).tolist()
    return dict(ref=ref, invA=torch.from_numpy(inva).int(), invE=torch.from_numpy(inve).int(), mapA=mapA, mapE=mapE,
                offA=cls_rows(ua), offE=cls_rows(ue), mgr=mgr, vsA=torch.tensor(sizes))


def test_class_branch_training_wiring_vs_oracle(monkeypatch):
    """head_train.class_branch (per-class grouping branch, all classes in grouped launches, per-class batch-statistics
    BatchNorm) on the emulated C ABI against the oracle's per-class loop in training mode: features, the three prediction
    maps and the gradients of every parameter of the branch and of the backbone features."""
    from cagroup3d_b200 import head_train as HT, model_init, sparse as S, synthetic
    E.install(monkeypatch)
    B, ncls = 2, 18
    batch = synthetic.make_batch(B, target_voxels=500, config=12)
    model = model_init.seeded_model(ncls, False, seed=6)
    pts = torch.from_numpy(batch["points"])
    cfg = O.default_cfg(ncls, False)
    orc0 = O.Oracle(model.state_dict(), cfg)
    orc0.train_bn = True
    model_init.calibrate_semantic_bias(model, orc0.forward(pts, B, stages="backbone")["bb_feats"], 0.10)
    orc = O.Oracle(model.state_dict(), cfg, dtype=torch.float64)
    branch = ("dense_head.cls_individual", "dense_head.centerness_conv", "dense_head.cls_conv", "dense_head.reg_conv",
              "dense_head.scales", "dense_head.feature_offset")
    names = [k for k in orc.p if k.startswith(branch) and k.endswith(("kernel", "bn.weight", "bn.bias", "conv.bias", "scale"))]
    for k in names + ["backbone_3d.out.4.bn.bias"]:              # the last one puts the backbone features on the tape
        orc.p[k] = orc.p[k].double().requires_grad_(True)
    orc.train_bn = True
    res = orc.forward(pts, B, cur_epoch=10, stages="head")
    res["bb_feats"].retain_grad()
    maps = res["head"]["maps"]
    g = torch.Generator().manual_seed(1)
    loss, dys = 0.0, []
    for m in maps:
        d = [torch.randn(tuple(m[k].shape), generator=g, dtype=torch.float64) for k in ("ctr", "cls", "bbox")]
        dys.append(d)
        loss = loss + (m["ctr"] * d[0]).sum() + (m["cls"] * d[1]).sum() + (m["bbox"] * d[2]).sum()
    loss.backward()

    model.train()
    head = model.dense_head
    head.semantic_threshold = 0.05
    mgr = S.Manager()
    cm = E.cpu_map(res["bb_coords"], 2, mgr)
    mgr.by_stride[2] = cm
    outF = res["bb_feats"].detach().float().contiguous().requires_grad_(True)
    out = S.SparseTensor(outF, cm, mgr)
    sem, offs, offF = HT.shared_part(head, out, impl="simt")
    assert (offF.detach().double() - res["head"]["offset_feat"].detach().reshape(offF.shape)).abs().max().item() < 1e-3
    art = _oracle_class_artifacts(res, 0.05, B, cfg, S.Manager())
    got = HT.class_branch(head, out, offF, art, B, impl="simt")
    want_coords = np.concatenate([np.concatenate([m["coords"][:, :1] + c * B, m["coords"][:, 1:]], 1) for c, m in enumerate(maps)])
    assert np.array_equal(got["coords"].numpy(), want_coords)
    cat = lambda k: torch.cat([m[k] for m in maps]).detach()
    assert sum(len(m["coords"]) for m in maps) > 10 * ncls
    for k_got, k_want in (("feat", "feat"), ("centerness", "ctr"), ("cls", "cls"), ("reg", "reg"), ("bbox_pred", "bbox")):
        assert (got[k_got].detach().double() - cat(k_want)).abs().max().item() < 2e-3, k_got
    dctr, dcls, dbox = (torch.cat([d[i] for d in dys]).float() for i in range(3))
    ((got["centerness"] * dctr).sum() + (got["cls"] * dcls).sum() + (got["bbox_pred"] * dbox).sum()).backward()
    params = dict(model.named_parameters())
    G = max(float(orc.p[k].grad.norm()) for k in names)
    worst = max((float((params[k].grad.double().reshape(orc.p[k].grad.shape) - orc.p[k].grad).norm())
                 / (float(orc.p[k].grad.norm()) + 1e-4 * G), k) for k in names)
    assert worst[0] < 2e-3, worst
    # gradient reaching the backbone features: directly (the un-voted copy of every selected voxel) and through the
    # offset features
    assert _rel(outF.grad, res["bb_feats"].grad) < 2e-3


def test_first_stage_loss_wiring_vs_training_oracle(monkeypatch):
    """head_train.first_stage_loss (all five terms of CAGroup3DHead.loss) on the emulated C ABI == oracle/train_oracle.
    first_stage_loss (which reproduces the REFERENCE's training step, tests/test_train_oracle.py) on the same batch, and
    its backward reaches the backbone features and every head parameter."""
    from cagroup3d_b200 import head_train as HT, model_init, ops, sparse as S, synthetic, train_targets as TT
    E.install(monkeypatch)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    B, ncls = 2, 18
    scenes = [synthetic.make_scene(1000 * 7 + i, 500, n_classes=ncls, return_masks=True) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    pts = torch.from_numpy(batch["points"])
    model = model_init.seeded_model(ncls, False, seed=3)
    cfg = O.default_cfg(ncls, False)
    orc = O.Oracle(model.state_dict(), cfg)
    orc.train_bn = True
    model_init.calibrate_semantic_bias(model, orc.forward(pts, B, stages="backbone")["bb_feats"], 0.10)
    with torch.no_grad():
        model.dense_head.cls_conv.bias.fill_(-2.0)
    gtb = [torch.from_numpy(b[:, :7]).float() for _, b, _, _ in scenes]
    gtl = [torch.from_numpy(b[:, 7]).long() for _, b, _, _ in scenes]
    semm = [torch.from_numpy(s) for _, _, s, _ in scenes]
    insm = [torch.from_numpy(m) for _, _, _, m in scenes]
    orc = O.Oracle(model.state_dict(), cfg)
    want = T.first_stage_loss(orc, pts, B, gtb, gtl, semm, insm, cur_epoch=10)
    orc.train_bn = True
    res = orc.forward(pts, B, cur_epoch=10, stages="head")

    model.train()
    head = model.dense_head
    head.semantic_threshold = 0.05
    mgr = S.Manager()
    cm = E.cpu_map(res["bb_coords"], 2, mgr)
    mgr.by_stride[2] = cm
    outF = res["bb_feats"].detach().float().contiguous().requires_grad_(True)
    art = _oracle_class_artifacts(res, 0.05, B, cfg, S.Manager())
    sp = [pts[pts[:, 0] == b][:, 1:4].contiguous() for b in range(B)]
    loss, tb = HT.first_stage_loss(head, S.SparseTensor(outF, cm, mgr), B, gtb, gtl, sp, semm, insm, impl="simt", art=art)
    for k in ("loss_centerness", "loss_bbox", "loss_cls", "loss_sem", "loss_vote", "one_stage_loss"):
        assert abs(tb[k] - want[k]) <= 2e-4 * max(1.0, abs(want[k])), (k, tb[k], want[k])
    assert want["loss_bbox"] > 0 and want["loss_centerness"] > 0
    loss.backward()
    assert outF.grad is not None and float(outF.grad.abs().max()) > 0
    missing = [n for n, p in head.named_parameters() if p.grad is None]
    assert missing == [], missing


# (the ScanNet variant runs the same driver code as the SUN RGB-D one, whose branches are a superset: CG3D_SLOW_TESTS=1)
@pytest.mark.parametrize("yaw", [pytest.param(False, marks=pytest.mark.slow), True])
def test_first_stage_training_step_driver(monkeypatch, yaw):
    """train_step.first_stage_training_step end to end on the emulated C ABI, with the coordinate phase served by the
    oracle-backed artifact builder: tb_dict keys of the reference, gradients in the reducer's buckets, loss going down.
    yaw: the SUN RGB-D configuration (10 classes, 3 votes per seed, yaw code + rotated IoU loss, no per-point masks; RoI stage
    with code size 7, (cos, sin) heading code and the IoU loss)."""
    from cagroup3d_b200 import dist as D, head_train as HT, model_init, ops, sparse as S, synthetic, train_step as TS, train_targets as TT
    from oracle import sort_vertices_oracle as SVO
    E.install(monkeypatch)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    monkeypatch.setattr(ops, "sort_v", lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.numpy(), m.numpy(), nv.numpy())).int())
    B, ncls = 2, (10 if yaw else 18)
    cfg = O.default_cfg(ncls, yaw)

    def voxelize_cpu(points, voxel_size):
        c = points[:, :4].clone()
        c[:, 1:] /= voxel_size
        ox = me.from_points(c, points[:, 4:])
        mgr = S.Manager()
        cm = E.cpu_map(ox.C, 1, mgr)
        mgr.by_stride[1] = cm
        return S.SparseTensor(ox.F.float().contiguous(), cm, mgr)

    def coordinate_phase_cpu(head, out, sem, offs, Bn):
        Cc = out.C.long()
        ts, vs = out.cmap.stride, head.voxel_size
        mx = ((Cc[:, 1:].max(0)[0] + ts) * vs).float()
        mn = ((Cc[:, 1:].min(0)[0] - ts) * vs).float()
        nv = offs.shape[1] // 3
        voted = (Cc[:, 1:].float() * vs).view(-1, 1, 3) + offs.detach().float().view(-1, nv, 3)
        voted = torch.maximum(torch.minimum(voted, mx.view(1, 1, 3)), mn.view(1, 1, 3))
        res = dict(bb_coords=out.C.numpy().astype(np.int64), head=dict(sem=sem.detach(), voted=voted))
        return _oracle_class_artifacts(res, head.semantic_threshold, Bn, cfg, S.Manager())
    monkeypatch.setattr(TS, "voxelize", voxelize_cpu)
    monkeypatch.setattr(HT, "coordinate_phase", coordinate_phase_cpu)
    scenes = [synthetic.make_scene(1000 * 7 + i, 150, n_classes=ncls, return_masks=True, sunrgbd=yaw) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    model = model_init.seeded_model(ncls, yaw, seed=3).train()
    with torch.no_grad():
        model.dense_head.semantic_conv.bias.fill_(-2.5)
    params = [p for n, p in model.named_parameters() if n.startswith(("backbone_3d.", "dense_head."))]
    opt = torch.optim.AdamW(params, lr=1e-3)
    red = D.GradientAllReducer(params, bucket_mb=64)
    masks = {} if yaw else {"semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
    bd = {"points": torch.from_numpy(batch["points"]).clone(), "batch_size": B, "cur_epoch": 3,
          "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float(), **masks}
    tb = TS.first_stage_training_step(model, bd, opt, red, impl="simt")
    assert set(tb) == {"loss_centerness", "loss_bbox", "loss_cls", "loss_sem", "loss_vote", "one_stage_loss"}
    losses = [tb["one_stage_loss"]]
    assert all(p.grad.data_ptr() == red._view(p).data_ptr() for p in params)
    # the pcdet-style call of the train loop (train_utils.py:56-58): both stages, then a first-stage-only model
    from cagroup3d_b200 import detector as DT, roi_train as RT
    monkeypatch.setattr(DT, "voxelize", voxelize_cpu)
    monkeypatch.setitem(S._CONV_IMPL, "name", "simt")            # the emulation covers the fp32 conv entry point
    gtb = [torch.from_numpy(b[:, :7]).float() for _, b, _, _ in scenes]
    gtl = [torch.from_numpy(b[:, 7]).long() for _, b, _, _ in scenes]

    def proposals_cpu(head, br, Bn):                             # stage-1 detections near the gt boxes (the NMS kernels are not emulated)
        g = torch.Generator().manual_seed(0)
        return [(torch.cat([gtb[b][:, :3] + torch.randn((len(gtb[b]), 3), generator=g) * 0.05, gtb[b][:, 3:6], gtb[b][:, 6:7] + (0.1 if yaw else 0.0)], 1),
                 torch.rand((len(gtb[b]),), generator=g), gtl[b].clone()) for b in range(Bn)]

    def roi_coordinate_phase_cpu(roi_head, sp, rois, Bn, rmax):
        orc = O.Oracle(model.state_dict(), cfg)
        omgr, ocm = me.Manager(), me.CoordMap(sp.C.numpy().astype(np.int64), 2)
        omgr.by_stride[2] = ocm
        pl = [(rois[b].detach().clone() * torch.tensor([1, 1, 1, 1, 1, 1, -1.0]), torch.ones(rmax), torch.zeros(rmax, dtype=torch.long)) for b in range(Bn)]
        _, inter = orc.roi_head(me.SparseTensor(sp.F.detach().clone(), ocm, omgr), pl, Bn)
        return _roi_artifacts(inter, cfg, sp, Bn * rmax)
    monkeypatch.setattr(HT, "stage1_proposals", proposals_cpu)
    monkeypatch.setattr(RT, "coordinate_phase", roi_coordinate_phase_cpu)
    bd = {"points": torch.from_numpy(batch["points"]).clone(), "batch_size": B, "cur_epoch": 3,
          "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float(), **masks}
    np.random.seed(0)
    ret, tb, disp = model(dict(bd, points=bd["points"].clone()))
    assert {"loss_all", "one_stage_loss", "rcnn_loss_reg", "loss_two_stage"} <= set(tb) and tb["rcnn_loss_reg"] > 0
    assert ("rcnn_loss_iou" in tb and 0 < tb["rcnn_loss_iou"] <= 1.0) if yaw else "rcnn_loss_iou" not in tb
    assert abs(tb["loss_all"] - tb["one_stage_loss"] - tb["loss_two_stage"]) < 1e-4
    losses.append(tb["one_stage_loss"])                          # after one AdamW step on the same batch
    assert np.isfinite(losses).all() and losses[1] < losses[0], losses
    assert abs(disp.pop("cur_semantic_value") - max(0.15 - 3 * 0.02, 0.05)) < 1e-9 and set(disp) == set(tb) - {"loss_all"}
    ret["loss"].backward()
    assert [n for n, p in model.named_parameters() if p.grad is None] == []
    if not yaw:                                                  # (the same driver code for both configurations: run once)
        tb2 = TS.training_step(model, dict(bd, points=bd["points"].clone()), opt, red, impl="simt", grad_norm_clip=10.0)
        assert set(tb2) == set(tb)


def _roi_artifacts(inter, cfg, sp, nr):
    """coordinate artifacts of roi_train.coordinate_phase from the oracle's RoI-head intermediates"""
    from cagroup3d_b200 import sparse as S
    gsz = cfg["grid"]
    uq, inv, _ = me.unique_first(np.concatenate([inter["grid_coords"][:, :1], inter["grid_coords"][:, 1:] * cfg["coord_key"]], 1))
    ii, jj, kk = np.meshgrid(np.arange(gsz), np.arange(gsz), np.arange(gsz), indexing="ij")
    tap = (ii + gsz * jj + gsz * gsz * kk).ravel()
    ptab = np.zeros((gsz ** 3, nr), np.int32)
    ptab[tap[None, :].repeat(nr, 0), np.arange(nr)[:, None].repeat(gsz ** 3, 1)] = inv.reshape(nr, gsz ** 3)
    umap = E.cpu_map(uq, 2, sp.mgr)
    nbr, order = S.neighbor_table(sp.cmap, umap, cfg["roi_kernel"], sp.mgr, ordered=True)
    return dict(umap=umap, nbr=nbr, order=order, ptab=torch.from_numpy(ptab), nr=nr)


@pytest.mark.parametrize("compiled", [False, pytest.param(True, marks=pytest.mark.slow)])       # compiled: 4 minutes, passes
def test_roi_branch_training_wiring_vs_oracle(monkeypatch, compiled):
    """roi_train.roi_branch (5^3 grid conv at the RoI grid voxels, 7^3 pooling contraction with its non-injective table,
    batch-statistics BatchNorm, regression MLP) on the emulated C ABI against the oracle's RoI head run with batch
    statistics: pooled features, rcnn_reg, gradients of every RoI-head parameter and of the backbone features."""
    from cagroup3d_b200 import model_init, roi_train as RT, sparse as S, synthetic
    E.install(monkeypatch, compiled=compiled)
    B, ncls = 2, 18
    scenes = [synthetic.make_scene(1000 * 9 + i, 700, n_classes=ncls) for i in range(B)]
    batch = synthetic.collate_batch(scenes)
    pts = torch.from_numpy(batch["points"])
    model = model_init.seeded_model(ncls, False, seed=5)
    cfg = O.default_cfg(ncls, False)
    orc = O.Oracle(model.state_dict(), cfg, dtype=torch.float64)
    names = [k for k in orc.p if k.startswith("roi_head.") and k.endswith(("kernel", "weight", "bias")) and "running" not in k]
    for k in names:
        orc.p[k] = orc.p[k].double().requires_grad_(True)
    res = orc.forward(pts, B, stages="backbone")
    g = torch.Generator().manual_seed(2)
    R = 9
    pred_list = []
    for b in range(B):
        gt = torch.from_numpy(scenes[b][1][:R, :7]).double()
        bx = torch.cat([gt[:, :3] + torch.randn((R, 3), generator=g).double() * 0.05, gt[:, 3:6] * 1.1, torch.zeros((R, 1), dtype=torch.float64)], 1)
        bx[-1] = bx[0]                                            # two identical RoIs: the pooling table repeats (tap, voxel) pairs
        pred_list.append((bx, torch.rand((R,), generator=g).double(), torch.randint(0, ncls, (R,), generator=g)))
    Fb = res["bb_feats"].detach().double().requires_grad_(True)
    omgr = me.Manager()
    ocm = me.CoordMap(res["bb_coords"], 2)
    omgr.by_stride[2] = ocm
    orc.train_bn = True
    _, inter = orc.roi_head(me.SparseTensor(Fb, ocm, omgr), pred_list, B)
    d1 = torch.randn(tuple(inter["pooled"].shape), generator=g, dtype=torch.float64)
    d2 = torch.randn(tuple(inter["rcnn_reg"].shape), generator=g, dtype=torch.float64)
    ((inter["pooled"] * d1).sum() + (inter["rcnn_reg"] * d2).sum()).backward()

    # coordinate artifacts from the oracle's intermediate results
    gsz = cfg["grid"]
    uq, inv, _ = me.unique_first(np.concatenate([inter["grid_coords"][:, :1], inter["grid_coords"][:, 1:] * cfg["coord_key"]], 1))
    assert np.array_equal(uq, inter["uniq"])
    nr = B * R
    ii, jj, kk = np.meshgrid(np.arange(gsz), np.arange(gsz), np.arange(gsz), indexing="ij")
    tap = (ii + gsz * jj + gsz * gsz * kk).ravel()
    ptab = np.zeros((gsz ** 3, nr), np.int32)
    ptab[tap[None, :].repeat(nr, 0), np.arange(nr)[:, None].repeat(gsz ** 3, 1)] = inv.reshape(nr, gsz ** 3)
    assert len(np.unique(ptab[:, [0, R - 1]], axis=1).T) == 1                   # the duplicated RoI: a non-injective table
    mgr = S.Manager()
    cm = E.cpu_map(res["bb_coords"], 2, mgr)
    mgr.by_stride[2] = cm
    umap = E.cpu_map(uq, 2, mgr)
    nbr, order = S.neighbor_table(cm, umap, cfg["roi_kernel"], mgr, ordered=True)
    art = dict(umap=umap, nbr=nbr, order=order, ptab=torch.from_numpy(ptab), nr=nr)
    model.train()
    F = res["bb_feats"].detach().float().contiguous().requires_grad_(True)
    rois = inter["rois"].float()
    pooled, reg, _ = RT.roi_branch(model.roi_head, S.SparseTensor(F, cm, mgr), rois, B, R, impl="simt", dropout=False, art=art)
    assert (pooled.detach().double() - inter["pooled"].detach()).abs().max().item() < 1e-3
    assert (reg.detach().double() - inter["rcnn_reg"].detach()).abs().max().item() < 1e-3
    ((pooled * d1.float()).sum() + (reg * d2.float()).sum()).backward()
    assert _rel(F.grad, Fb.grad) < 2e-3
    params = dict(model.named_parameters())
    G = max(float(orc.p[k].grad.norm()) for k in names)
    worst = max((float((params[k].grad.double().reshape(orc.p[k].grad.shape) - orc.p[k].grad).norm())
                 / (float(orc.p[k].grad.norm()) + 1e-4 * G), k) for k in names)
    assert worst[0] < 5e-3, worst


@pytest.mark.parametrize("compiled", [False, True])
def test_roi_targets_and_loss_equal_the_reference(monkeypatch, compiled):
    """roi_train.reorder_rois / ProposalTargetLayer / assign_targets / roi_reg_loss on the seeded inputs of
    tests/golden/roi_train_parts.npz -- produced by the REFERENCE's own ProposalTargetLayer, CAGroup3DRoIHead.assign_targets
    and get_box_reg_layer_loss (tests/golden/make_roi_train_golden.py) -- with the same host generator seeds: the same
    sampled RoIs, targets, masks and loss."""
    import os
    from cagroup3d_b200 import ops, roi_train as RT, train_targets as TT
    from tests.golden import make_roi_train_golden as MK
    E.install(monkeypatch, compiled=compiled)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_train_parts.npz"))
    gtb, gtl, preds = MK.inputs()
    rois, scores, labels = RT.reorder_rois(preds)
    assert np.array_equal(rois.numpy(), z["padded_rois"]) and np.array_equal(labels.numpy(), z["padded_labels"])
    inp = dict(batch_size=2, rois=rois, roi_scores=scores, roi_labels=labels, gt_bboxes_3d=[b.clone() for b in gtb], gt_labels_3d=gtl)
    np.random.seed(0)
    torch.manual_seed(0)
    t = RT.assign_targets(RT.ProposalTargetLayer(roi_per_image=128, fg_ratio=0.9, reg_fg_thresh=0.3), inp, 6)
    assert np.array_equal(t["rois"].numpy(), z["rois"]) and np.array_equal(t["roi_labels"].numpy(), z["roi_labels"])
    assert np.array_equal(t["reg_valid_mask"].numpy(), z["reg_valid_mask"]) and int((t["reg_valid_mask"] > 0).sum()) == int(z["n_fg"])
    assert np.abs(t["gt_iou_of_rois"].numpy() - z["gt_iou_of_rois"]).max() < 1e-5
    for k in ("gt_of_rois", "gt_of_rois_src", "gt_label_of_rois", "rcnn_cls_labels", "roi_scores"):
        assert np.abs(t[k].numpy() - z[k]).max() < 1e-5, k
    reg = torch.from_numpy(z["rcnn_reg"]).requires_grad_(True)
    loss, tb = RT.roi_reg_loss(reg, t, 6, [1.0] * 6)
    assert abs(float(loss.detach()) - float(z["rcnn_loss_reg"])) < 1e-5 and tb["loss_two_stage"] == tb["rcnn_loss_reg"]
    loss.backward()
    fg = torch.from_numpy(z["reg_valid_mask"]).view(-1) > 0
    assert float(reg.grad[~fg].abs().max()) == 0 and float(reg.grad[fg].abs().max()) > 0


def test_autograd_bricks_through_the_compiled_kernels(monkeypatch):
    """BatchNorm (+ residual + ReLU) on a ROW SLICE of a wider matrix, ReLU / ELU, bias and the average pool, with the
    training kernels' own CUDA sources (compiled for the CPU) behind the Python glue: pointers, strides and sizes exactly as
    autograd.py hands them to the library."""
    from cagroup3d_b200 import autograd as A, sparse as S
    E.install(monkeypatch, compiled=True)
    g = torch.Generator().manual_seed(0)
    n, C = 700, 40
    big = (torch.randn((n + 300, C), generator=g) * 2 + 1)
    R, gamma, beta, dY = torch.randn((n, C), generator=g), torch.rand((C,), generator=g) + 0.5, torch.randn((C,), generator=g), torch.randn((n, C), generator=g)
    Xd, Rd, gd, bd = (t.double().requires_grad_(True) for t in (big, R, gamma, beta))
    want = torch.relu(torch.nn.functional.batch_norm(Xd[200:200 + n], None, None, gd, bd, training=True) + Rd)
    (want * dY.double()).sum().backward()
    X, Rg, gg, bg = (t.clone().requires_grad_(True) for t in (big, R, gamma, beta))
    rm, rv = torch.zeros(C), torch.ones(C)
    Y = A.batch_norm_train(X[200:200 + n], gg, bg, rm, rv, act="relu", residual=Rg)
    assert _rel(Y.detach(), want.detach()) < 1e-5
    Y.backward(dY)
    assert _rel(X.grad, Xd.grad) < 1e-4 and _rel(Rg.grad, Rd.grad) < 1e-6 and _rel(gg.grad, gd.grad) < 1e-4 and _rel(bg.grad, bd.grad) < 1e-4
    assert float(X.grad[:200].abs().max()) == 0 and float(X.grad[200 + n:].abs().max()) == 0
    assert _rel(rm, 0.1 * big[200:200 + n].double().mean(0)) < 1e-5
    for name, fn, op in (("relu", torch.relu, A.relu), ("elu", torch.nn.functional.elu, A.elu)):
        xd = R.double().requires_grad_(True)
        (fn(xd) * dY.double()).sum().backward()
        xg = R.clone().requires_grad_(True)
        op(xg).backward(dY)
        assert _rel(xg.grad, xd.grad) < 1e-6, name
    b = torch.randn((1, C), generator=g).requires_grad_(True)
    xg = R.clone().requires_grad_(True)
    A.add_bias(xg, b).backward(dY)
    assert _rel(b.grad, dY.double().sum(0, keepdim=True)) < 1e-5 and torch.equal(xg.grad, dY)
    rng = np.random.default_rng(1)
    c = me.unique_first(np.concatenate([rng.integers(0, 2, (300, 1)), rng.integers(-6, 6, (300, 3)) * 2], 1))[0]
    omgr = me.Manager()
    ox = me.SparseTensor(torch.from_numpy(rng.standard_normal((len(c), 16))).requires_grad_(True), me.CoordMap(c, 2), omgr)
    omgr.by_stride[2] = ox.cmap
    wantp = me.avg_pool(ox, 5, 2)
    dO = torch.from_numpy(rng.standard_normal(tuple(wantp.F.shape)))
    (wantp.F * dO).sum().backward()
    mgr = S.Manager()
    cm = E.cpu_map(c, 2, mgr)
    mgr.by_stride[2] = cm
    Fg = ox.F.detach().float().requires_grad_(True)
    y = A.avg_pool(S.SparseTensor(Fg, cm, mgr), 5, 2)
    assert np.array_equal(y.C.numpy(), wantp.C) and _rel(y.F.detach(), wantp.F.detach()) < 1e-5
    y.F.backward(dO.float())
    assert _rel(Fg.grad, ox.F.grad) < 1e-5
