"""GPU parity: every C-ABI op against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star / SURVEY.md 8c): coordinates, rule maps, index lists -- bit exact (as
sets where the reference leaves row order unspecified); features -- max |delta| <= 1e-3 fp32 (the
exact-fp32 SIMT path is held to 2e-5)."""
import numpy as np
import pytest
import torch

from oracle import iou3d_oracle, me_cpu as me
from tests.util import assert_same_coord_set, rules_from_oracle, rules_from_table, sort_rows, to_gpu_sparse

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand_cloud(seed, n=4000, batch=2, extent=3.0, surface=True):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand((n, 3), generator=g) - 0.5) * extent
    if surface:
        xyz[:, 2] = torch.round(xyz[:, 2] * 2) / 2 + 0.003 * torch.randn((n,), generator=g)
    b = torch.randint(0, batch, (n, 1), generator=g).float()
    rgb = torch.randint(0, 256, (n, 3), generator=g).float()
    return torch.cat([b, xyz, rgb], 1).contiguous()


def oracle_tensor(seed, C, n=4000, batch=2):
    pts = rand_cloud(seed, n, batch)
    c = pts[:, :4].clone()
    c[:, 1:] /= 0.02
    g = torch.Generator().manual_seed(seed + 7)
    x = me.from_points(c, torch.randn((n, C), generator=g))
    return x


def test_quantize_unique_first_exact(lib):
    from cagroup3d_b200 import sparse as S
    from cagroup3d_b200.detector import voxelize
    pts = rand_cloud(1, n=20000, batch=3)
    pts[:500] = pts[500:1000]            # forced duplicates
    pts[:, 4:] /= 255.
    x = voxelize(pts.to(DEV), 0.02)
    c = pts[:, :4].clone()
    c[:, 1:] /= 0.02
    ox = me.from_points(c, pts[:, 4:])
    assert (x.C.cpu().numpy() == ox.C).all()                 # same rows in the same (first occurrence) order
    assert torch.equal(x.F.cpu(), ox.F)                       # the first point's colour, bit exact


def test_voxel_pyramid_equals_on_demand_maps(lib):
    """sparse.voxel_pyramid (all strided maps of BiResNet / DAPPM with device-side row counts, ONE size read-back) gives
    exactly the maps the on-demand path builds one sync at a time: same rows in the same order, same table contents
    (probed through a rule map), and the oracle's row sets."""
    from cagroup3d_b200 import sparse as S
    from cagroup3d_b200.detector import voxelize
    pts = rand_cloud(5, n=60000, batch=3)
    pts[:, 1:4] *= 4.0                                       # spread: the coarse levels keep more than a handful of rows
    pts[:, 4:] /= 255.
    a = voxelize(pts.to(DEV), 0.02, pyramid=True)
    b = voxelize(pts.to(DEV), 0.02, pyramid=False)
    assert torch.equal(a.C, b.C) and torch.equal(a.F, b.F)
    assert sorted(a.mgr.by_stride) == [1] + [p[0] for p in S.BACKBONE_PYRAMID] and sorted(b.mgr.by_stride) == [1]
    for ts, src in S.BACKBONE_PYRAMID:
        want = S.strided_map(b.mgr.by_stride[src], b.mgr, ts // src)                 # one host sync each
        got = S.strided_map(a.mgr.by_stride[src], a.mgr, ts // src)                  # cached: no work
        assert got is a.mgr.by_stride[ts] and torch.equal(got.coords, want.coords), ts
        if got.n:
            ta = S.neighbor_table(a.mgr.by_stride[src], got, 3, a.mgr)
            tb = S.neighbor_table(b.mgr.by_stride[src], want, 3, b.mgr)
            assert torch.equal(ta, tb), ts
    c = pts[:, :4].clone()
    c[:, 1:] /= 0.02
    ox = me.from_points(c, pts[:, 4:])
    o2 = ox.mgr.strided(ox.cmap, 2)
    assert (a.mgr.by_stride[2].coords.cpu().numpy() == o2.coords).all() or \
        set(map(tuple, a.mgr.by_stride[2].coords.cpu().numpy())) == set(map(tuple, o2.coords))


def test_negative_and_empty_inputs(lib):
    from cagroup3d_b200 import sparse as S
    pts = torch.tensor([[0, -0.001, 0.0, 0.019, 1, 2, 3], [0, -0.02, -0.0200001, 0.02, 4, 5, 6]], dtype=torch.float32)
    q, err = S.quantize(pts.to(DEV), 7, 2, (0.02,) * 3)
    ref = torch.floor(pts[:, 1:4] / 0.02).int()
    assert (q.cpu()[:, 1:] == ref).all() and int(err.item()) == 0
    cm, first, inv = S.unique_first(torch.zeros((0, 4), dtype=torch.int32, device=DEV), 1, None, True, True)
    assert cm.n == 0
    far = torch.tensor([[0, 700.0, 0, 0, 0, 0, 0]], dtype=torch.float32)      # 35000 > 32767
    _, err = S.quantize(far.to(DEV), 7, 1, (0.02,) * 3)
    assert int(err.item()) == 1


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (1, 2), (5, 1)])
def test_rule_maps_exact(lib, k, stride):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(3, 4)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    omap_o = ox.cmap if stride == 1 else ox.mgr.strided(ox.cmap, stride)
    omap = x.cmap if stride == 1 else S.strided_map(x.cmap, x.mgr, stride)
    assert (omap.coords.cpu().numpy() == omap_o.coords).all()            # strided map: same first-occurrence order
    nbr = S.neighbor_table(x.cmap, omap, k, x.mgr)
    assert rules_from_table(nbr) == rules_from_oracle(me.kernel_map(ox.cmap, omap_o.coords, k, 1))
    assert S.count_rules(nbr) == len(rules_from_table(nbr))


@pytest.mark.parametrize("k", [3, 5, 9])
def test_symmetric_rule_map_equals_probed_rule_map(lib, k):
    """cg3d_neighbor_table_symmetric (half the probes + mirrored writes) == cg3d_neighbor_table, bit for bit."""
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(40 + k, 8, n=5000, batch=3)
    x = to_gpu_sparse(ox.C, ox.F, 2)
    n = x.cmap.n
    a = torch.empty((k ** 3, n), dtype=torch.int32, device=DEV)
    b = torch.full((k ** 3, n), 12345, dtype=torch.int32, device=DEV)
    S._call("cg3d_neighbor_table", x.cmap.coords, n, x.cmap.keys, x.cmap.vals, x.cmap.capacity, k, 2, a)
    S._call("cg3d_neighbor_table_symmetric", x.cmap.coords, n, x.cmap.keys, x.cmap.vals, x.cmap.capacity, k, 2, b)
    assert torch.equal(a, b)
    assert torch.equal(S.neighbor_table(x.cmap, x.cmap, k, x.mgr), a)          # the host routes same-map odd kernels to it


@pytest.mark.parametrize("cin,cout,k,stride", [(3, 64, 3, 1), (64, 64, 3, 2), (64, 128, 1, 2), (32, 48, 5, 1), (20, 7, 3, 1)])
def test_spconv_simt_vs_oracle(lib, cin, cout, k, stride):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(5, cin)
    g = torch.Generator().manual_seed(11)
    W = torch.randn((k ** 3, cin, cout), generator=g) / np.sqrt(cin * k ** 3)
    scale, shift = torch.rand((cout,), generator=g) + 0.5, torch.randn((cout,), generator=g)
    ref = me.conv(ox, W, k, stride)
    res = torch.randn((ref.F.shape[0], cout), generator=g)
    want = torch.relu(ref.F * scale + shift + res)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    y = S.conv(x, W.to(DEV), k, stride, scale=scale.to(DEV), shift=shift.to(DEV), residual=res.to(DEV), act="relu",
               impl="simt")
    assert (y.C.cpu().numpy() == ref.C).all()
    assert (y.F.cpu() - want).abs().max().item() <= 2e-5


def test_spconv_input_relu_and_column_slices(lib):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(6, 48)
    g = torch.Generator().manual_seed(2)
    W = torch.randn((27, 32, 16), generator=g) / 30
    ref = me.conv(ox.with_F(torch.relu(ox.F[:, 8:40])), W, 3, 1)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    out = torch.zeros((ox.F.shape[0], 40), device=DEV)
    nbr = S.neighbor_table(x.cmap, x.cmap, 3, x.mgr)
    S.gemm_rows(x.F[:, 8:40], nbr, W.to(DEV), x.cmap.n, 27, in_act="relu", out=out[:, 24:40], impl="simt")
    assert (out[:, 24:].cpu() - ref.F).abs().max().item() <= 2e-5
    assert out[:, :24].abs().max().item() == 0


def test_transpose_k2s2_and_generative_k3s3(lib):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(8, 16)
    coarse_o = me.conv(ox, torch.randn(27, 16, 16) / 20, 3, 2)
    g = torch.Generator().manual_seed(4)
    W = torch.randn((8, 16, 24), generator=g) / 4
    ref = me.conv_transpose_k2s2(coarse_o, W)
    x = to_gpu_sparse(ox.C, ox.F, 1, strided={2: coarse_o.C})
    cx = S.SparseTensor(coarse_o.F.to(DEV), x.mgr.by_stride[2], x.mgr)
    y = S.conv_transpose_k2s2(cx, W.to(DEV), impl="simt")
    assert (y.C.cpu().numpy() == ref.C).all()
    assert (y.F.cpu() - ref.F).abs().max().item() <= 2e-5
    # generative k3 s3 onto given coordinates (A13)
    pts = rand_cloud(9, 3000, 2)
    fine = pts[:, :4].clone(); fine[:, 1:] /= 0.05
    coarse = pts[:, :4].clone(); coarse[:, 1:] = torch.floor(pts[:, 1:4] / 0.15) * 3
    A = me.from_points(fine, torch.randn(3000, 8), average=True)
    E = me.from_points(coarse, torch.randn(3000, 8), average=True, stride=3)
    W3 = torch.randn((27, 8, 8), generator=g)
    ref = me.generative_transpose_k3s3(E, W3, A.cmap)
    gA, gE = to_gpu_sparse(A.C, A.F, 1), to_gpu_sparse(E.C, E.F, 3)
    nbr = S.transpose_table(gE.cmap, gA.cmap, 3, None)
    out = S.gemm_rows(gE.F, nbr, W3.to(DEV), gA.cmap.n, 27, impl="simt")
    assert (out.cpu() - ref).abs().max().item() <= 2e-5
    assert (nbr >= 0).sum(0).max().item() <= 1                         # exactly one candidate parent per fine voxel


def test_interp_avgpool_segment_mean(lib):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(12, 32, n=6000)
    coarse = me.conv(ox, torch.randn(27, 32, 32) / 30, 3, 2)
    coarse = me.conv(coarse, torch.randn(27, 32, 32) / 30, 3, 2)          # stride 4
    ref = me.features_at(coarse, ox.C)
    x = to_gpu_sparse(ox.C, ox.F, 1, strided={4: coarse.C})
    cx = S.SparseTensor(coarse.F.to(DEV), x.mgr.by_stride[4], x.mgr)
    base = torch.randn(ox.F.shape)
    got = S.interp(cx, x.C, base=base.to(DEV))
    assert (got.cpu() - (ref + base)).abs().max().item() <= 1e-5
    # non-zero average pooling k5 s2 on the stride-4 map
    pooled_o = me.avg_pool(coarse, 5, 2)
    pooled = S.avg_pool(cx, 5, 2)
    assert (pooled.C.cpu().numpy() == pooled_o.C).all()
    assert (pooled.F.cpu() - pooled_o.F).abs().max().item() <= 1e-5
    # UNWEIGHTED_AVERAGE quantisation
    pts = rand_cloud(13, 5000, 2)
    c = pts[:, :4].clone(); c[:, 1:] /= 0.1
    f = torch.randn(5000, 16)
    A = me.from_points(c, f, average=True)
    q, _ = S.quantize(pts.to(DEV), 7, 5000, (0.1,) * 3)
    cm, _, inv = S.unique_first(q, 1, None, want_inverse=True)
    got = S.segment_mean(f.to(DEV), 16, None, 0, None, inv, 5000, cm.n, 16)
    assert (cm.coords.cpu().numpy() == A.C).all()
    assert (got.cpu() - A.F).abs().max().item() <= 1e-5


def test_sort_pairs_stable(lib):
    from cagroup3d_b200.sparse import sort_pairs
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 255, 2049, 70001):
        keys = torch.randint(0, 1 << 40, (n,), generator=g, dtype=torch.int64)
        keys[::3] = keys[0]                                                # many ties
        vals = torch.arange(n, dtype=torch.int32)
        k, v = keys.to(DEV), vals.to(DEV)
        sort_pairs(k, v, n, 40)
        rk, ri = torch.sort(keys, stable=True)
        assert torch.equal(k.cpu(), rk) and torch.equal(v.cpu().long(), ri)


def random_boxes(n, seed, yaw=True):
    g = torch.Generator().manual_seed(seed)
    b = torch.zeros((n, 7))
    b[:, :3] = (torch.rand((n, 3), generator=g) - 0.5) * 4
    b[:, 3:6] = torch.rand((n, 3), generator=g) * 1.5 + 0.1
    if yaw:
        b[:, 6] = (torch.rand((n,), generator=g) - 0.5) * 6
    return b


def test_pairwise_iou_vs_oracle(lib):
    from cagroup3d_b200 import sparse as S
    a, b = random_boxes(200, 1), random_boxes(150, 2)
    a[:20] = b[:20]                                                        # identical boxes
    for mode, name, tol in ((0, "overlap", 2e-5), (1, "iou", 2e-5), (2, "iou_normal", 0.0)):
        out = torch.empty((200, 150), device=DEV)
        S._call("cg3d_boxes_pairwise_bev", a.to(DEV), 200, b.to(DEV), 150, mode, out)
        ref = iou3d_oracle.pairwise(a, b, name)
        assert (out.cpu() - ref).abs().max().item() <= tol, name
    # known answers: identical -> 1, 2x2 boxes shifted by 0.5 -> 0.6
    kat = torch.tensor([[0, 0, 0, 2, 2, 1, 0], [0.5, 0, 0, 2, 2, 1, 0.]])
    out = torch.empty((2, 2), device=DEV)
    S._call("cg3d_boxes_pairwise_bev", kat.to(DEV), 2, kat.to(DEV), 2, 1, out)
    assert abs(out[0, 0].item() - 1) < 1e-6 and abs(out[0, 1].item() - 0.6) < 1e-6


@pytest.mark.parametrize("path", ["auto", "serial", "blocked", "cluster"])
@pytest.mark.parametrize("rotated", [0, 1])
def test_nms_segments_vs_oracle(lib, rotated, path, monkeypatch):
    """path: the one-sweep-per-kept-box kernel, the blocked greedy kernel on one CTA or on a cluster of CTAs per segment, or
    the library's own choice (cluster here: 1500 > 64, 5 segments)"""
    from cagroup3d_b200 import sparse as S
    if path != "auto":
        monkeypatch.setenv("CG3D_NMS", path)
    segs, boxes = [0], []
    want = []
    for s, n in enumerate((300, 0, 1, 1500, 64)):
        b = random_boxes(n, 10 + s, yaw=bool(rotated))
        sc = torch.rand((n,), generator=torch.Generator().manual_seed(s))
        order = torch.sort(sc, descending=True, stable=True)[1]
        b = b[order]
        boxes.append(b)
        keep = iou3d_oracle.nms(b, torch.arange(n, 0, -1).float(), 0.5, bool(rotated))
        flags = torch.zeros(n, dtype=torch.int32)
        flags[keep] = 1
        want.append(flags)
        segs.append(segs[-1] + n)
    allb = torch.cat(boxes).to(DEV)
    keep = torch.empty((segs[-1],), dtype=torch.int32, device=DEV)
    cnt = torch.empty((5,), dtype=torch.int32, device=DEV)
    S._call("cg3d_nms_segments", allb, allb.shape[0], torch.tensor(segs, dtype=torch.int32, device=DEV), 5, 1500, 0.5, rotated, keep, cnt)
    assert torch.equal(keep.cpu(), torch.cat(want))
    assert cnt.cpu().tolist() == [int(w.sum()) for w in want]


def _clustered_boxes(n, seed, yaw):
    """n candidates jittered around 12 objects (what a trained head hands to the NMS: heavy suppression)"""
    g = torch.Generator().manual_seed(seed)
    obj = random_boxes(12, seed + 1000, yaw=yaw)
    b = obj[torch.randint(0, 12, (n,), generator=g)].clone()
    b[:, :3] += torch.randn((n, 3), generator=g) * 0.08
    b[:, 3:6] *= 1 + (torch.rand((n, 3), generator=g) - 0.5) * 0.3
    if yaw:
        b[:, 6] += torch.randn((n,), generator=g) * 0.1
    return b


@pytest.mark.parametrize("case", ["scattered-40x1000", "clustered-40x10000"])
@pytest.mark.parametrize("rotated", [0, 1])
def test_nms_long_segments_all_paths_agree(lib, rotated, case, monkeypatch):
    """40 (sample, class) segments as a training step hands them to the NMS -- up to nms_pre = 1000 scattered candidates, or
    up to n_classes * nms_pre = 10000 candidates clustered on 12 objects: the blocked greedy kernel (one CTA, and a cluster of
    3 CTAs per segment) == the one-sweep-per-kept-box kernel, flag for flag; prints the device times."""
    from cagroup3d_b200 import sparse as S
    g = torch.Generator().manual_seed(3)
    if case.startswith("scattered"):
        lens = torch.randint(600, 1001, (40,), generator=g).tolist()
        boxes = torch.cat([random_boxes(n, 50 + i, yaw=bool(rotated)) for i, n in enumerate(lens)]).to(DEV)
    else:
        lens = torch.randint(6000, 10001, (40,), generator=g).tolist()
        boxes = torch.cat([_clustered_boxes(n, 50 + i, bool(rotated)) for i, n in enumerate(lens)]).to(DEV)
    seg = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
    n = boxes.shape[0]
    res, ms = {}, {}
    for path in ("serial", "blocked", "cluster"):
        monkeypatch.setenv("CG3D_NMS", path)
        keep, cnt = torch.empty((n,), dtype=torch.int32, device=DEV), torch.empty((40,), dtype=torch.int32, device=DEV)
        for it in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            S._call("cg3d_nms_segments", boxes, n, seg, 40, max(lens), 0.5, rotated, keep, cnt)
            b.record()
            torch.cuda.synchronize()
        res[path], ms[path] = (keep.cpu(), cnt.cpu()), a.elapsed_time(b)
    print(f"\nNMS {case} rotated={rotated}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in ms.items()) + f", kept {int(res['serial'][0].sum())} of {n}")
    for path in ("blocked", "cluster"):
        assert torch.equal(res["serial"][0], res[path][0]) and torch.equal(res["serial"][1], res[path][1]), path
    assert 0 < int(res["serial"][0].sum()) < n
