"""The TRAINING kernels' CUDA sources executed on the CPU (tests/cuda_on_cpu: g++ build of the unmodified .cu files over a
thread-per-CUDA-thread shim with real barriers / warp exchanges) against the oracles the GPU tests use.  These kernels were
written without access to a GPU; this checks their source -- indexing, reductions, barrier placement, launch arithmetic --
until tests/test_zz_gpu_spconv_backward.py has run on hardware.  The product path never loads this library."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import backward_oracle as Bk
from oracle import me_cpu as me
from oracle import train_oracle as T

f32, i32, i64 = np.float32, np.int32, np.int64


@pytest.fixture(scope="module")
def K():
    from cagroup3d_b200 import _lib
    from tests.cuda_on_cpu import build
    lib = build.load()
    protos = _lib.parse_header()

    def call(name, *args):
        fn = getattr(lib, name)
        types = protos[name]
        stream = [] if name in _lib._host_only else [None]
        assert len(args) + len(stream) == len(types), (name, len(args), len(types))
        fn.argtypes, fn.restype = types, ctypes.c_int
        conv = [a.ctypes.data if isinstance(a, np.ndarray) else a for a in args]
        rc = fn(*conv, *stream)
        assert rc == 0 or name in _lib._host_only, (name, rc)
        return rc
    return call


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def test_wgrad_transpose_kernels(K):
    rng = np.random.default_rng(0)
    n_in, n, Kt, Cin, Cout = 900, 2500, 3, 40, 24
    nbr = rng.integers(-1, n_in, (Kt, n)).astype(i32)
    nbr[:, rng.random(n) < 0.3] = -1
    X, dY = rng.standard_normal((n_in, Cin)).astype(f32), rng.standard_normal((n, Cout)).astype(f32)
    perm = rng.permutation(n).astype(i32)
    for out_rows, c0, c1, act in ((None, 0, n, 0), (perm, 0, n, 1), (perm, 700, 1900, 0)):
        S = K("cg3d_spconv_wgrad_slabs", c1 - c0, Cin, Cout, Kt)
        slabs = np.zeros((S * Kt * Cin * Cout,), f32)
        dW = np.full((Kt, Cin, Cout), np.nan, f32)
        K("cg3d_spconv_wgrad", X, Cin, act, nbr, dY, Cout, n, c0, c1, Cin, Cout, Kt, out_rows, slabs if S > 1 else None, dW)
        masked = nbr.copy()
        masked[:, :c0] = -1
        masked[:, c1:] = -1
        Xa = np.maximum(X, 0) if act else X
        _, want = Bk.conv_backward(torch.from_numpy(Xa).double(), torch.zeros((Kt, Cin, Cout), dtype=torch.float64), masked,
                                   torch.from_numpy(dY).double(), out_rows=out_rows)
        assert _rel(dW, want.numpy()) < 1e-5, (S, c0, c1)
    assert K("cg3d_spconv_wgrad_slabs", 2500, Cin, Cout, Kt) > 1            # the slab path was exercised
    # identity rows (1x1 conv / Linear)
    dW1 = np.zeros((1, Cin, Cout), f32)
    K("cg3d_spconv_wgrad", X[:800], Cin, 0, None, dY[:800], Cout, 800, 0, 800, Cin, Cout, 1, None, None, dW1)
    assert _rel(dW1[0], X[:800].astype(np.float64).T @ dY[:800].astype(np.float64)) < 1e-5
    # transposed table of an injective rule map, positional
    inj = np.full((Kt, n), -1, i32)
    for k in range(Kt):
        cols = rng.permutation(n)[:n_in // 2]
        inj[k, cols] = rng.permutation(n_in)[:n_in // 2]
    T_ = np.zeros((Kt, n_in), i32)
    K("cg3d_table_transpose", inj, Kt, n, perm, n_in, T_)
    assert np.array_equal(T_, Bk.table_transpose(inj, n_in, perm))
    W = rng.standard_normal((5, Cin, Cout)).astype(f32)
    Wt = np.zeros((5, Cout, Cin), f32)
    K("cg3d_transpose_weights", W, 5, Cin, Cout, Wt)
    assert np.array_equal(Wt, W.transpose(0, 2, 1))


def test_batchnorm_and_rowwise_kernels(K):
    rng = np.random.default_rng(1)
    n, C = 3000, 40
    X = (rng.standard_normal((n, C)) * 3 + rng.standard_normal(C) * 5).astype(f32)
    gamma, beta, dY = (rng.random(C) + 0.5).astype(f32), rng.standard_normal(C).astype(f32), rng.standard_normal((n, C)).astype(f32)
    ws = np.zeros((K("cg3d_bn_train_workspace", n, C),), f32)
    mean, rstd, scale, shift = (np.zeros(C, f32) for _ in range(4))
    rm, rv = np.zeros(C, f32), np.ones(C, f32)
    K("cg3d_bn_train_stats", X, C, n, C, 1e-5, 0.1, gamma, beta, ws, mean, rstd, scale, shift, rm, rv)
    Xd = X.astype(np.float64)
    assert _rel(mean, Xd.mean(0)) < 1e-6 and _rel(rstd, 1 / np.sqrt(Xd.var(0) + 1e-5)) < 1e-5
    assert _rel(rm, 0.1 * Xd.mean(0)) < 1e-6 and _rel(rv, 0.9 + 0.1 * Xd.var(0, ddof=1)) < 1e-5
    Y = np.maximum(X * scale + shift, 0).astype(f32)
    want_y = np.maximum((Xd - Xd.mean(0)) / np.sqrt(Xd.var(0) + 1e-5) * gamma + beta, 0)
    assert _rel(Y, want_y) < 1e-5
    dx, dres, dg, db = np.zeros((n, C), f32), np.zeros((n, C), f32), np.zeros(C, f32), np.zeros(C, f32)
    K("cg3d_bn_train_backward", X, C, dY, C, Y, C, n, C, mean, rstd, gamma, ws, dx, C, dres, C, dg, db)
    dy_eff = np.where(Y > 0, dY, 0)
    wdx, wdg, wdb = Bk.batchnorm_train_backward(torch.from_numpy(Xd), torch.from_numpy(gamma).double(), torch.from_numpy(dy_eff).double())
    assert _rel(dx, wdx.numpy()) < 2e-5 and _rel(dg, wdg.numpy()) < 2e-5 and _rel(db, wdb.numpy()) < 2e-5
    assert np.array_equal(dres, dy_eff.astype(f32))
    cs = np.zeros(C, f32)
    K("cg3d_column_sum", dY, C, n, C, ws, cs)
    assert _rel(cs, dY.astype(np.float64).sum(0)) < 1e-5
    for act, fn in ((1, lambda v: np.maximum(v, 0)), (2, lambda v: np.where(v > 0, v, np.expm1(v)))):
        y = fn(X / 4).astype(f32)
        d = np.zeros((n, C), f32)
        K("cg3d_act_backward", dY, C, y, C, n, C, act, d, C)
        want = np.where(y > 0, dY, 0 if act == 1 else dY * (y + 1))
        assert _rel(d, want) < 1e-6
    U = 500
    inv = rng.integers(0, U, n).astype(i32)
    inv[:U] = np.arange(U)
    dO = rng.standard_normal((U, C)).astype(f32)
    cnt = np.bincount(inv, minlength=U).astype(f32)
    dIn = np.zeros((n, C), f32)
    K("cg3d_segment_mean_backward", dO, inv, cnt, n, C, dIn, C)
    assert _rel(dIn, Bk.segment_mean_backward(torch.from_numpy(dO).double(), inv, n).numpy()) < 1e-6
    order = np.argsort(inv, kind="stable").astype(i32)
    off = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=U))]).astype(i32)
    out = np.zeros((U, C), f32)
    K("cg3d_segment_sum_sorted", dY, order, off, U, C, out)
    want = np.zeros((U, C))
    np.add.at(want, inv, dY.astype(np.float64))
    assert _rel(out, want) < 1e-5


def _hash_table(coords):
    """the library's open-addressing table (common.cuh: cg3d_pack / cg3d_hash, linear probing)"""
    M = (1 << 64) - 1

    def h(k):
        k ^= k >> 33; k = (k * 0xff51afd7ed558ccd) & M; k ^= k >> 33; k = (k * 0xc4ceb9fe1a85ec53) & M; k ^= k >> 33
        return k & 0xFFFFFFFF
    cap = 1
    while cap < 2 * len(coords):
        cap *= 2
    keys = np.full(cap, M, np.uint64)
    vals = np.full(cap, -1, i32)
    for r, c in enumerate(coords):
        k = ((int(c[0]) & 0xFFFF) << 48) | (((int(c[1]) + 32768) & 0xFFFF) << 32) | (((int(c[2]) + 32768) & 0xFFFF) << 16) | ((int(c[3]) + 32768) & 0xFFFF)
        s = h(k) & (cap - 1)
        while keys[s] != M:
            s = (s + 1) & (cap - 1)
        keys[s], vals[s] = k, r
    return keys, vals, cap


def test_interp_and_avgpool_backward_kernels(K):
    rng = np.random.default_rng(2)
    ts, tq, C = 4, 1, 40
    q = np.unique(np.concatenate([rng.integers(0, 2, (700, 1)), rng.integers(-12, 12, (700, 3)) * tq], 1), axis=0)
    src = q.copy()
    src[:, 1:] = np.floor_divide(src[:, 1:], ts) * ts
    src = np.unique(src, axis=0)
    src = src[rng.random(len(src)) < 0.7]
    dOut = rng.standard_normal((len(q), C)).astype(f32)
    rows, w = Bk.interp_corners(me.CoordMap(src.astype(i64), ts), q.astype(i64))
    want = Bk.interp_backward(torch.from_numpy(dOut).double(), rows, w, len(src)).numpy()
    keys, vals, cap = _hash_table(q)
    dF = np.full((len(src), C), np.nan, f32)
    K("cg3d_interp_trilinear_backward", src.astype(i32), len(src), ts, keys, vals, cap, tq, dOut, C, dF)
    assert _rel(dF, want) < 1e-5
    # average pooling backward: coarse outputs over a window of inputs
    ic = np.unique(np.concatenate([rng.integers(0, 2, (300, 1)), rng.integers(-6, 6, (300, 3)) * 2], 1), axis=0).astype(i32)
    oc = ic.copy()
    oc[:, 1:] = np.floor_divide(oc[:, 1:], 4) * 4
    oc = np.unique(oc, axis=0).astype(i32)
    half = 4
    m = (np.abs(ic[None, :, 1:] - oc[:, None, 1:]).max(-1) <= half) & (ic[None, :, 0] == oc[:, None, 0])
    dO = rng.standard_normal((len(oc), C)).astype(f32)
    cnt, dIn = np.zeros(len(oc), f32), np.zeros((len(ic), C), f32)
    K("cg3d_avgpool_window_backward", oc, len(oc), ic, len(ic), half, dO, C, cnt, dIn)
    assert np.array_equal(cnt, m.sum(1).astype(f32))
    assert _rel(dIn, m.T.astype(np.float64) @ (dO / m.sum(1)[:, None])) < 1e-5


def test_assigner_and_loss_kernels(K):
    from tests.test_zz_gpu_spconv_backward import _assign_case
    for yaw in (False, True):
        pts, boxes, labels, ncls = _assign_case(yaw)
        pts = [p[:400] for p in pts]
        locs = np.ascontiguousarray(torch.cat(pts).numpy(), f32)
        offs = np.cumsum([0] + [len(p) for p in pts]).astype(i32)
        n, m = len(locs), len(boxes)
        b, l = np.ascontiguousarray(boxes.numpy(), f32), labels.numpy().astype(i32)
        for topk in (18, 3):
            ct, bt, lb = T.assign(pts, boxes, labels, topk)
            kth, c, bx, lab, idx = np.zeros(m, f32), np.zeros(n, f32), np.zeros((n, 7), f32), np.zeros(n, i64), np.zeros(n, i32)
            K("cg3d_assign", locs, n, offs, ncls, b, l, m, topk, kth, c, bx, lab, idx)
            same = lab == lb.numpy()
            assert (~same).mean() <= (0.0 if not yaw else 2e-3)
            pos = same & (lb.numpy() >= 0)
            assert pos.sum() > 10 and np.array_equal(bx[pos], bt.numpy()[pos]) and np.abs(c[pos] - ct.numpy()[pos]).max() < 1e-5
        sl, il = T.assign_semantic(torch.from_numpy(locs), boxes, labels)
        gs, gi = np.zeros(n, i64), np.zeros(n, i64)
        K("cg3d_assign_semantic", locs, n, b, l, m, gs, gi)
        assert ((gs != sl.numpy()) | (gi != il.numpy())).mean() <= (0.0 if not yaw else 2e-3)
    rng = np.random.default_rng(3)
    n, C = 700, 18
    pred, lab = (rng.standard_normal((n, C)) * 3).astype(f32), rng.integers(-1, C, n).astype(i64)
    pd = torch.from_numpy(pred).double().requires_grad_(True)
    want = T.focal_loss(pd, torch.from_numpy(lab), 37.0)
    want.backward()
    ws, loss, grad = np.zeros(max(K("cg3d_focal_loss_workspace", n, C), K("cg3d_loss_workspace", n * C)), f32), np.zeros(1, f32), np.zeros((n, C), f32)
    K("cg3d_focal_loss", pred, lab, n, C, 2.0, 0.25, 37.0, ws, loss, grad)
    assert abs(loss[0] - float(want)) < 1e-5 * max(1, abs(float(want))) and _rel(grad, pd.grad.numpy()) < 1e-5
    P = 600
    tgt = np.concatenate([rng.standard_normal((P, 3)), rng.random((P, 3)) + 0.3], 1).astype(f32)
    pb = (tgt + rng.standard_normal((P, 6)) * 0.3).astype(f32)
    pb[:, 3:] = np.abs(pb[:, 3:]) + 0.05
    pb[:30, :3] += 5
    w = rng.random(P).astype(f32)
    pdb = torch.from_numpy(pb).double().requires_grad_(True)
    wl = T.axis_aligned_iou_loss(pdb, torch.from_numpy(tgt).double(), torch.from_numpy(w).double(), float(w.sum()))
    wl.backward()
    g6 = np.zeros((P, 6), f32)
    K("cg3d_iou_loss_aa", pb, 6, tgt, 6, w, P, float(w.sum()), ws, loss, g6, 6)
    assert abs(loss[0] - float(wl)) < 1e-5 and _rel(g6, pdb.grad.numpy()) < 2e-5
    x, t = (rng.standard_normal(P) * 2).astype(f32), rng.random(P).astype(f32)
    xd = torch.from_numpy(x).double().requires_grad_(True)
    wb = T.bce_loss(xd, torch.from_numpy(t).double(), 37.0)
    wb.backward()
    g1 = np.zeros(P, f32)
    K("cg3d_bce_loss", x, t, P, 37.0, ws, loss, g1)
    assert abs(loss[0] - float(wb)) < 1e-5 * max(1, abs(float(wb))) and _rel(g1, xd.grad.numpy()) < 1e-6
    p3, t3, w3 = (rng.standard_normal((P, 3)) * 0.1).astype(f32), (rng.standard_normal((P, 3)) * 0.1).astype(f32), rng.random((P, 3)).astype(f32)
    pd3 = torch.from_numpy(p3).double().requires_grad_(True)
    wsl = T.smooth_l1_sum(pd3, torch.from_numpy(t3).double(), torch.from_numpy(w3).double())
    wsl.backward()
    g3 = np.zeros((P, 3), f32)
    K("cg3d_smooth_l1_loss", p3, t3, w3, P, 3, 0.04, ws, loss, g3)
    assert abs(loss[0] - float(wsl)) < 1e-5 * max(1, abs(float(wsl))) and _rel(g3, pd3.grad.numpy()) < 1e-6


def test_vote_targets_kernel(K):
    from cagroup3d_b200 import synthetic
    pts, boxes, sem, ins = synthetic.make_scene(1000 * 7 + 1, 600, n_classes=18, return_masks=True)
    sp = np.ascontiguousarray(pts[:, :3], f32)
    gtb = np.ascontiguousarray(boxes[:, :7], f32)
    vox = np.unique(np.floor(sp / 0.04), axis=0).astype(f32) * f32(0.04)
    spt, vt = torch.from_numpy(sp), torch.from_numpy(vox)
    want_t, want_m = T.vote_targets_from_masks(spt, vt, torch.from_numpy(gtb), torch.from_numpy(sem), torch.from_numpy(ins), 18)
    nearest = torch.cdist(vt, spt).argmin(1).numpy().astype(i32)
    n_inst = int(ins.max()) + 1
    ws, centers = np.zeros(n_inst * 8, i32), np.zeros((n_inst, 3), f32)
    tg, mk = np.zeros((len(vox), 3), f32), np.zeros(len(vox), f32)
    K("cg3d_vote_targets", sp, 3, sem.astype(i64), ins.astype(i64), len(sp), n_inst, 18, gtb, len(gtb), vox, nearest, len(vox), ws, centers,
      tg, mk)
    assert np.array_equal(mk, want_m.numpy()) and np.abs(tg - want_t.numpy()).max() < 1e-5


def _nms_boxes(rng, n, rotated):
    """clustered BEV boxes (many overlaps, ties in position) + a few far away, descending-score order is the row order"""
    centres = rng.uniform(-3, 3, (max(n // 6, 1), 2))
    b = np.zeros((n, 7), f32)
    b[:, :2] = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.25, (n, 2))
    b[:, 2] = rng.uniform(-1, 1, n)
    b[:, 3:6] = rng.uniform(0.3, 1.5, (n, 3))
    if rotated:
        b[:, 6] = rng.uniform(-3.2, 3.2, n)
    if n > 8:
        b[5] = b[2]                                              # an exact duplicate: IoU 1
        b[-3:, :2] += 40.0                                       # far apart: the early reject path
    return b


@pytest.mark.parametrize("rotated", [0, 1])
def test_nms_blocked_kernel_equals_the_sweep_kernel_and_the_oracle(K, rotated, monkeypatch):
    """csrc/nms.cu: the blocked greedy NMS (64-row blocks: pair tile, keep bits, later boxes against the block's kept boxes)
    gives the same keep flags and counts as the one-sweep-per-kept-box kernel and as oracle/iou3d_oracle.nms (pinned to the
    reference's iou3d_nms), on
    segments of 0, 1, 63, 64, 65, 130 and 300 boxes in one launch (block boundaries, partial last blocks, empty segments)."""
    from oracle import iou3d_oracle
    import torch
    rng = np.random.default_rng(5 + rotated)
    lens = [0, 1, 63, 64, 65, 130, 0, 300]
    boxes = np.concatenate([_nms_boxes(rng, n, rotated) for n in lens]).astype(f32)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(i32)
    n = len(boxes)
    out = {}
    for mode in ("serial", "blocked"):
        monkeypatch.setenv("CG3D_NMS", mode)
        keep, cnt = np.full(n, -1, i32), np.full(len(lens), -1, i32)
        K("cg3d_nms_segments", boxes, n, seg, len(lens), max(lens), 0.5, rotated, keep, cnt)
        out[mode] = (keep, cnt)
    for mode in ("blocked",):
        assert np.array_equal(out["serial"][0], out[mode][0]) and np.array_equal(out["serial"][1], out[mode][1]), mode
    keep, cnt = out["blocked"]
    assert set(np.unique(keep)) <= {0, 1} and np.array_equal(cnt, [keep[seg[s]:seg[s + 1]].sum() for s in range(len(lens))])
    assert 0 < keep.sum() < n
    for s, ln in enumerate(lens):                                # the oracle, segment by segment
        b = boxes[seg[s]:seg[s + 1]]
        want = iou3d_oracle.nms(torch.from_numpy(b), torch.arange(ln, 0, -1).float(), 0.5, bool(rotated)).numpy()
        got = np.nonzero(keep[seg[s]:seg[s + 1]])[0]
        if rotated:                                              # libm sin / cos differ in the last bit between the two builds
            assert len(np.setxor1d(got, want)) <= max(1, ln // 100), (s, ln)
        else:
            assert np.array_equal(got, want), (s, ln)
    # a loose bound (max_segment_len = n_boxes, what a caller without the true maximum passes) changes nothing
    monkeypatch.setenv("CG3D_NMS", "blocked")
    keep2 = np.full(n, -1, i32)
    K("cg3d_nms_segments", boxes, n, seg, len(lens), n, 0.5, rotated, keep2, None)
    assert np.array_equal(keep2, keep)


@pytest.mark.parametrize("n,begin,end", [(2, 0, 8), (300, 0, 27), (4096, 6, 39), (4097, 0, 27), (20000, 0, 64), (13000, 32, 64)])
def test_onesweep_radix_sort_is_a_stable_sort(K, n, begin, end):
    """csrc/sort.cu cg3d_sort_pairs (digit histograms in one launch, then one chained-look-back launch per 8-bit digit) ==
    numpy's stable argsort on the key bits [begin, end): keys with many duplicates, values = original index, so the
    order of equal keys is checked; several tiles (4096 keys each), a partial last tile, odd and even numbers of digits."""
    rng = np.random.default_rng(n + begin)
    width = end - begin
    base = rng.integers(0, 2 ** min(width, 62), n, dtype=np.uint64) if n > 300 else rng.integers(0, 7, n, dtype=np.uint64)
    base[::3] = base[0]                                           # heavy duplication
    noise = rng.integers(0, 2 ** 62, n, dtype=np.uint64)
    lowmask = np.uint64((1 << begin) - 1)
    keys = ((base << np.uint64(begin)) | (noise & lowmask)).astype(np.uint64)      # bits below begin_bit must be ignored
    if end < 64:
        keys &= np.uint64((1 << end) - 1)
    vals = np.arange(n, dtype=i32)
    k, v = keys.copy(), vals.copy()
    ws = np.full(K("cg3d_sort_workspace_ints", n), -1, i32)         # (the sort clears what it needs itself)
    K("cg3d_sort_pairs", k, v, n, begin, end, np.zeros(n, np.uint64), np.zeros(n, i32), ws)
    digits = (keys >> np.uint64(begin)) & (np.uint64((1 << (8 * ((width + 7) // 8))) - 1) if 8 * ((width + 7) // 8) < 64 else np.uint64(2 ** 64 - 1))
    order = np.argsort(digits, kind="stable")
    assert np.array_equal(v, order.astype(i32)) and np.array_equal(k, keys[order])


@pytest.mark.parametrize("n", [0, 1, 15, 4096, 4097, 70001])
def test_chained_exclusive_scan(K, n):
    """csrc/coords.cu cg3d_exclusive_scan_i32 (one kernel: tiles of 4096 ints in ticket order, the offset of a tile from the
    published sums of the tiles before it) == numpy cumsum; unaligned tails, one and many tiles, in place."""
    rng = np.random.default_rng(n)
    x = rng.integers(0, 5, max(n, 1)).astype(i32)
    x[::7] = 0
    out = np.full(max(n, 1), -7, i32)
    ws = np.full(K("cg3d_scan_workspace_ints", n), -1, i32)
    total = np.full(1, -1, i32)
    K("cg3d_exclusive_scan_i32", x, n, out, ws, total)
    want = np.concatenate([[0], np.cumsum(x[:n])]).astype(i32)
    assert np.array_equal(out[:n], want[:n]) and total[0] == want[n]
    if n:
        y = x.copy()                                             # in place, and from an address that is not 16-byte aligned
        buf = np.zeros(n + 1, i32)
        buf[1:] = x[:n]
        K("cg3d_exclusive_scan_i32", buf[1:], n, buf[1:], ws, total)
        assert np.array_equal(buf[1:], want[:n]) and total[0] == want[n]


# ---- inference-side kernels (detect.cu, pool_interp.cu, proposal.cu) ------------------------------------------------------------
@pytest.mark.parametrize("nreg", [6, 8])
def test_head_decode_kernel_vs_oracle(K, nreg):
    """csrc/detect.cu cg3d_head_decode == the oracle's forward_single tail (cagroup_head.py:636-649: Scale + exp on the six
    face distances; :590-593: sigmoid(cls) * sigmoid(centerness); :654-703: _bbox_pred_to_bbox incl. the 'fcaf3d' yaw code) on
    class-batched rows (batch index = class * B + sample, per-class voxel sizes and scales)."""
    import torch
    from oracle import cagroup3d_oracle as O
    rng = np.random.default_rng(nreg)
    n, ncls, B = 700, 5, 2
    coords = np.concatenate([rng.integers(0, ncls * B, (n, 1)), rng.integers(-40, 40, (n, 3))], 1).astype(i32)
    pred = rng.normal(0, 1.0, (n, 1 + ncls + nreg)).astype(f32)
    vsA = rng.uniform(0.02, 0.2, (ncls, 3)).astype(f32)
    scales = rng.uniform(0.5, 1.5, ncls).astype(f32)
    scores, mx, boxes = np.zeros((n, ncls), f32), np.zeros(n, f32), np.full((n, 7), -9, f32)
    K("cg3d_head_decode", pred, pred.shape[1], coords, n, ncls, nreg, B, vsA, scales, scores, mx, boxes, 7)
    cls_of = coords[:, 0] // B
    p = torch.from_numpy(pred).double()
    pts = torch.from_numpy(coords[:, 1:].astype(np.float64) * vsA[cls_of].astype(np.float64))
    reg = p[:, 1 + ncls:]
    bp = torch.cat([torch.exp(reg[:, :6] * torch.from_numpy(scales[cls_of].astype(np.float64))[:, None]), reg[:, 6:]], 1)
    want_boxes = O.bbox_pred_to_bbox(pts, bp).numpy()
    want_scores = (torch.sigmoid(p[:, 1:1 + ncls]) * torch.sigmoid(p[:, :1])).numpy()
    assert _rel(scores, want_scores) < 1e-6 and _rel(mx, want_scores.max(1)) < 1e-6
    assert _rel(boxes[:, :want_boxes.shape[1]], want_boxes) < 2e-6
    if nreg == 6:
        assert np.all(boxes[:, 6] == 0)


@pytest.mark.parametrize("code_size,sincos", [(6, 0), (7, 0), (7, 1)])
def test_roi_decode_kernel_vs_oracle(K, code_size, sincos):
    """csrc/detect.cu cg3d_roi_decode == CAGroupResidualCoder.decode_torch on the zero-centred RoI + rotation by the RoI heading +
    shift to its centre (cagroup_roi_head.py:477-510; oracle: residual_decode + rotate_z)."""
    import torch
    from oracle import cagroup3d_oracle as O
    rng = np.random.default_rng(10 * code_size + sincos)
    n = 333
    rois = np.concatenate([rng.uniform(-3, 3, (n, 3)), rng.uniform(0.3, 2, (n, 3)), rng.uniform(-3.1, 3.1, (n, 1)) * (code_size > 6)], 1).astype(f32)
    reg = rng.normal(0, 0.3, (n, code_size + sincos)).astype(f32)
    out = np.full((n, code_size), -9, f32)
    K("cg3d_roi_decode", rois, reg, n, code_size, sincos, out)
    r64 = torch.from_numpy(rois).double()
    local = r64[:, :code_size].clone()
    local[:, 0:3] = 0
    dec = O.residual_decode(torch.from_numpy(reg).double(), local, code_size, bool(sincos)).view(-1, code_size)
    if code_size > 6:
        dec = torch.cat([O.rotate_z(dec[:, None, 0:3], r64[:, 6])[:, 0], dec[:, 3:]], 1)
    dec[:, 0:3] += r64[:, 0:3]
    assert _rel(out, dec.numpy()) < 2e-6


@pytest.mark.parametrize("with_ref", [False, True])
def test_segment_mean_kernel(K, with_ref):
    """csrc/pool_interp.cu cg3d_segment_mean (histogram -> scan -> fill -> one warp per unique row, 2^-30 fixed-point sums) ==
    the per-segment mean; with `ref`: point p reads slice `kind` of a wide source row or a row of the second source
    (the head's voted / original copies, cagroup_head.py:257-271).  `inverse` comes from a unique pass in the product, so
    every segment holds at least one point (as here); bit-repeatable."""
    rng = np.random.default_rng(int(with_ref))
    n, U, C, nv = 1500, 400, 64, 3
    inv = rng.integers(0, U, n).astype(i32)
    inv[:U] = np.arange(U)                                       # every segment occurs
    inv[U:U + 50] = 7                                            # one long segment
    out, cnt = np.full((U, C), -9, f32), np.full(U, -9, f32)
    ws = np.zeros(K("cg3d_segment_mean_workspace", n, U), np.int64)
    if with_ref:
        rows = 300
        A, Bm = rng.normal(0, 2, (rows, nv * C)).astype(f32), rng.normal(0, 2, (rows, C)).astype(f32)
        ref = np.stack([rng.integers(0, rows, n), rng.integers(-1, nv, n)], 1).astype(i32)
        K("cg3d_segment_mean", A, nv * C, Bm, C, ref, inv, n, U, C, out, cnt, ws)
        feat = np.where((ref[:, 1] >= 0)[:, None], A.reshape(rows, nv, C)[ref[:, 0], np.maximum(ref[:, 1], 0)], Bm[ref[:, 0]])
    else:
        A = rng.normal(0, 2, (n, C)).astype(f32)
        K("cg3d_segment_mean", A, C, None, 0, None, inv, n, U, C, out, cnt, ws)
        feat = A
    sums = np.zeros((U, C), np.float64)
    np.add.at(sums, inv, feat.astype(np.float64))
    m = np.bincount(inv, minlength=U)
    want = sums / m[:, None]
    assert m.min() >= 1 and np.array_equal(cnt, m.astype(f32)) and _rel(out, want) < 1e-6
    out2 = np.zeros_like(out)
    K("cg3d_segment_mean", *( (A, nv * C, Bm, C, ref) if with_ref else (A, C, None, 0, None) ), inv.copy(), n, U, C, out2, cnt, ws)
    assert np.array_equal(out, out2)


def test_topk_keys_and_rank_filter_kernels(K):
    """csrc/proposal.cu cg3d_topk_keys -> cg3d_sort_pairs -> cg3d_rank_filter == `max_scores.topk(nms_pre)` per (sample, class
    map) segment (cagroup_head.py:595-599): segments longer than nms_pre keep their nms_pre best rows (ties: lower row
    first), shorter ones keep all rows in row order."""
    rng = np.random.default_rng(3)
    nseg, nms_pre = 6, 50
    lens = [0, 20, 50, 51, 300, 120]
    seg = np.repeat(np.arange(nseg), lens).astype(i32)
    perm = rng.permutation(len(seg))
    seg = seg[perm]                                              # rows of a segment are scattered over the array
    n = len(seg)
    score = rng.random(n).astype(f32)
    score[::9] = score[0]                                        # ties
    counts = np.bincount(seg, minlength=nseg).astype(i32)
    keys, vals = np.zeros(n, np.uint64), np.zeros(n, i32)
    K("cg3d_topk_keys", seg, score, n, counts, nms_pre, keys, vals)
    ws = np.full(K("cg3d_sort_workspace_ints", n), -1, i32)
    K("cg3d_sort_pairs", keys, vals, n, 0, 64, np.zeros(n, np.uint64), np.zeros(n, i32), ws)
    seg_off = np.concatenate([[0], np.cumsum(counts)]).astype(i32)
    flags = np.full(n, -1, i32)
    K("cg3d_rank_filter", keys, n, seg_off, nms_pre, flags)
    assert np.array_equal(seg[vals], np.sort(seg, kind="stable"))           # sorted by segment
    for s in range(nseg):
        rows = vals[seg_off[s]:seg_off[s + 1]]
        kept = rows[flags[seg_off[s]:seg_off[s + 1]] == 1]
        mine = np.nonzero(seg == s)[0]
        if len(mine) <= nms_pre:
            assert np.array_equal(kept, mine)                                # everything, in row order
        else:
            order = mine[np.argsort(-score[mine], kind="stable")][:nms_pre]  # best scores, lower row first on ties
            assert np.array_equal(kept, order)


def test_head_coordinate_kernels(K):
    """csrc/detect.cu: cg3d_coord_bounds / cg3d_first_rows / cg3d_vote_points / cg3d_semantic_flags -> scan -> cg3d_compact_rows
    == their numpy statements (cagroup_head.py:207-230: pad ids, scene bounds, clamped voted points, per-class selection in
    row order)."""
    rng = np.random.default_rng(8)
    n, B, ncls, nv, ts, vs = 1024, 3, 4, 3, 2, 0.02                # whole warps (first_rows groups lanes with __activemask)
    b = np.sort(rng.integers(0, B, n)).astype(i32)
    coords = np.concatenate([b[:, None], rng.integers(-60, 90, (n, 3)) * ts], 1).astype(i32)
    mm = np.zeros(6, i32)
    K("cg3d_coord_bounds", coords, n, mm)
    assert np.array_equal(mm, np.concatenate([coords[:, 1:].min(0), coords[:, 1:].max(0)]))
    first = np.full(B, 0x7F7F7F7F, i32)
    K("cg3d_first_rows", coords, n, B, first)
    assert np.array_equal(first, [np.nonzero(b == k)[0][0] for k in range(B)])
    off = rng.normal(0, 0.8, (n, nv, 3)).astype(f32)
    voted = np.zeros((n, nv, 3), f32)
    K("cg3d_vote_points", coords, off, n, nv, vs, ts, mm, voted)
    lo, hi = ((mm[:3] - ts).astype(f32) * f32(vs)), ((mm[3:] + ts).astype(f32) * f32(vs))
    want = np.maximum(np.minimum(coords[:, None, 1:].astype(f32) * f32(vs) + off, hi), lo)
    assert np.array_equal(voted, want) and (voted == hi).any() and (voted == lo).any()
    sem = rng.normal(-1.5, 1.5, (n, ncls)).astype(f32)
    thr = 0.15
    flags = np.zeros(ncls * n, i32)
    K("cg3d_semantic_flags", sem, n, ncls, thr, flags)
    sig = 1.0 / (1.0 + np.exp(-sem.astype(np.float64)))
    want_f = (sig > thr).T.reshape(-1)
    edge = np.abs(sig.T.reshape(-1) - thr) < 1e-6
    assert np.array_equal(flags[~edge] == 1, want_f[~edge])
    pos, total = np.zeros(ncls * n, i32), np.zeros(1, i32)
    K("cg3d_exclusive_scan_i32", flags, ncls * n, pos, np.zeros(K("cg3d_scan_workspace_ints", ncls * n), i32), total)
    sel = np.full(max(int(total[0]), 1), -1, i32)
    K("cg3d_compact_rows", flags, pos, n, ncls, sel)
    want_sel = np.concatenate([np.nonzero(flags[c * n:(c + 1) * n])[0] for c in range(ncls)])
    assert total[0] == len(want_sel) > 0 and np.array_equal(sel[:total[0]], want_sel)


def test_knn_kernels_vs_exhaustive_numpy(K):
    """csrc/train_ops.cu: cg3d_knn (k = 1 and k = 3) == an exhaustive scan in index order with the kernel's distance
    expression (knn_cuda.cu:58-94: strict `<`, first index wins); cg3d_knn_grid (counting sort into cells + ring search) ==
    cg3d_knn bit for bit, incl. duplicated points, queries outside the box and a far cluster."""
    rng = np.random.default_rng(4)
    n, m = 4500, 320
    xyz = rng.uniform(-2, 2, (1, n, 3)).astype(f32)
    xyz[0, 100:110] = xyz[0, 90:100]                             # duplicated points: ties at distance 0 for a query on them
    xyz[0, -40:] += 30.0                                         # a far cluster
    q = rng.uniform(-2.2, 2.2, (1, m, 3)).astype(f32)
    q[0, :10] = xyz[0, 100:110]
    q[0, 10:14] = [[9, 9, 9], [-9, 0, 0], [31, 31, 31], [0, 0, -50]]
    idx1, d1 = np.zeros((1, m, 1), i32), np.zeros((1, m, 1), f32)
    K("cg3d_knn", xyz, 1, n, q, m, 1, idx1, d1)
    # the kernel's expression (the reference binary's contraction): fmaf(dz, dz, fmaf(dx, dx, dy * dy)); an fp32 fma is the
    # exactly formed product + addend rounded once: float64 holds the 48-bit product exactly
    dx, dy, dz = ((q[0, :, None, a] - xyz[0, None, :, a]).astype(f32) for a in range(3))
    f64 = np.float64
    d = (dy * dy).astype(f32)
    d = (dx.astype(f64) * dx.astype(f64) + d.astype(f64)).astype(f32)
    d = (dz.astype(f64) * dz.astype(f64) + d.astype(f64)).astype(f32)
    want = d.argmin(1)
    assert np.array_equal(idx1[0, :, 0], want) and np.array_equal(d1[0, :, 0], d[np.arange(m), want])
    assert np.array_equal(idx1[0, :10, 0], np.arange(90, 100))   # the first of two identical points wins
    idx3, d3 = np.zeros((1, m, 3), i32), np.zeros((1, m, 3), f32)
    K("cg3d_knn", xyz, 1, n, q, m, 3, idx3, d3)
    order = np.argsort(d, 1, kind="stable")[:, :3]
    assert np.array_equal(d3[0], np.take_along_axis(d, order, 1))
    # equal distances (the duplicated points) leave the reference's heap in heap order: rows whose four smallest distances
    # are distinct must match exactly; for the others every returned index must carry the returned distance
    distinct = (np.diff(np.sort(d, 1)[:, :4], axis=1) > 0).all(1)
    assert distinct.sum() > 290 and np.array_equal(idx3[0][distinct], order[distinct])
    assert np.array_equal(np.take_along_axis(d, idx3[0].astype(np.int64), 1), d3[0])
    assert all(len(set(r)) == 3 for r in idx3[0].tolist())
    ig, dg = np.full((1, m), -1, i32), np.zeros((1, m), f32)
    ws = np.zeros(K("cg3d_knn_grid_workspace", n) + 4, i32)
    off = (-ws.ctypes.data // 4) % 4                             # 16-byte aligned start
    K("cg3d_knn_grid", xyz, 1, n, q, m, ig, dg, ws[off:])
    assert np.array_equal(ig[0], idx1[0, :, 0]) and np.array_equal(dg[0], d1[0, :, 0])


def test_sort_vertices_kernel_vs_oracle(K):
    """csrc/train_ops.cu cg3d_sort_vertices == oracle/sort_vertices_oracle.py (the numpy restatement of sort_vert_kernel.cu that
    served the reference's Python when the rotated-IoU goldens were made) on random polygons with masked-out vertices."""
    from oracle import sort_vertices_oracle as SVO
    rng = np.random.default_rng(2)
    b, n, m = 2, 96, 24
    mask = np.zeros((b, n, m), bool)                             # 0 .. 8 valid vertices (a convex intersection has at most 8)
    for bi in range(b):
        for r in range(n):
            mask[bi, r, rng.choice(m, rng.integers(0, 9), replace=False)] = True
    mask[0, 0] = False
    mask[0, 0, :3] = True                                        # a triangle
    mask[0, 1] = False                                           # no valid vertex at all
    nv = mask.sum(-1).astype(i32)
    verts = rng.normal(0, 1, (b, n, m, 2)).astype(f32)
    cen = (verts * mask[..., None]).sum(2, keepdims=True) / np.maximum(nv, 1)[..., None, None]
    verts = (verts - cen).astype(f32)
    idx = np.full((b, n, 9), -7, i32)
    K("cg3d_sort_vertices", verts, mask.astype(np.uint8), nv, b, n, m, idx)
    want = SVO.sort_vertices(verts, mask, nv)
    assert np.array_equal(idx, want)


@pytest.mark.parametrize("cin,cout,k,grouped", [(3, 16, 27, False), (20, 7, 8, False), (16, 24, 5, True)])
def test_spconv_simt_kernel_vs_numpy(K, cin, cout, k, grouped):
    """csrc/spconv_simt.cu cg3d_spconv_simt (exact-fp32 gather conv: the stem, the narrow heads, the training fallback) ==
    out = act((sum_k in_act(in[nbr[k]]) @ W[g][k]) * scale + shift + residual) in fp64: missing neighbours (-1), input ReLU,
    ELU, residual, column-sliced input / output, positional tables (out_rows) and the grouped (tile -> weight group) mode."""
    rng = np.random.default_rng(cin * 100 + k)
    n_in, n_out, G = 500, 300, (3 if grouped else 1)
    ldi, ldo = cin + 5, cout + 3
    X = rng.normal(0, 1, (n_in, ldi)).astype(f32)
    nbr = rng.integers(-1, n_in, (k, n_out)).astype(i32)
    nbr[rng.random((k, n_out)) < 0.5] = -1
    W = rng.normal(0, 0.3, (G, k, cin, cout)).astype(f32)
    scale, shift = rng.uniform(0.5, 1.5, (G, cout)).astype(f32), rng.normal(0, 0.2, (G, cout)).astype(f32)
    res = rng.normal(0, 1, (n_out, cout)).astype(f32)
    out = np.full((n_out, ldo), -9, f32)
    perm = rng.permutation(n_out).astype(i32)                    # position j of the table is output row perm[j]
    if grouped:
        bounds = [0, 100, 100, 300]                              # an empty group in the middle
        tiles = [(r0, min(64, bounds[g + 1] - r0), g) for g in range(G) for r0 in range(bounds[g], bounds[g + 1], 64)]
        t0, tn, tg = (np.array(c, i32) for c in zip(*tiles))
        K("cg3d_spconv_simt", X, ldi, 1, nbr, W, out, ldo, n_out, cin, cout, k, scale, shift, res, 2, t0, tn, tg, len(tiles), perm)
        grp = np.repeat(np.arange(G), np.diff(bounds))
    else:
        K("cg3d_spconv_simt", X, ldi, 1, nbr, W, out, ldo, n_out, cin, cout, k, scale, shift, res, 2, None, None, None, 0, perm)
        grp = np.zeros(n_out, np.int64)
    Xr = np.maximum(X[:, :cin].astype(np.float64), 0)
    acc = np.zeros((n_out, cout))
    for j in range(n_out):
        for t in range(k):
            if nbr[t, j] >= 0:
                acc[j] += Xr[nbr[t, j]] @ W[grp[j], t].astype(np.float64)
    y = acc * scale[grp] + shift[grp] + res[perm]
    y = np.where(y > 0, y, np.expm1(y))
    want = np.zeros((n_out, cout))
    want[perm] = y
    assert _rel(out[:, :cout], want) < 1e-5
    assert np.all(out[:, cout:] == -9)                           # the slice's neighbours are untouched


def test_voxelise_and_rule_map_kernels_vs_me_oracle(K):
    """csrc/coords.cu: cg3d_quantize -> cg3d_unique_first (hash insert with the min-row winner, scan, compact) == the oracle's
    ME-CPU first-occurrence unique (unique rows, inverse, first rows, all exact); cg3d_stride_coords + a second unique = the
    stride-2 map; cg3d_neighbor_table (3^3, strided 2^3 down-sampling map) and cg3d_neighbor_table_symmetric == the oracle's
    kernel maps entry for entry; cg3d_hash_lookup finds every row and rejects absent coordinates."""
    from oracle import me_cpu as me
    rng = np.random.default_rng(6)
    n = 3000
    pts = np.concatenate([rng.integers(0, 2, (n, 1)).astype(f32), rng.uniform(-0.6, 0.6, (n, 3)).astype(f32)], 1)
    pts[::11] = pts[5]                                           # many duplicates of one voxel
    coords, err = np.zeros((n, 4), i32), np.zeros(1, i32)
    K("cg3d_quantize", pts, 4, n, 0.02, 0.02, 0.02, 1, coords, err)
    want_c = np.concatenate([pts[:, :1], np.floor(pts[:, 1:] / f32(0.02))], 1).astype(np.int64)
    assert err[0] == 0 and np.array_equal(coords, want_c)

    def unique(c):
        m = len(c)
        cap = K("cg3d_hash_capacity", m)
        keys, vals = np.zeros(cap, np.uint64), np.zeros(cap, i32)
        out, first, inv, nu = np.zeros((m, 4), i32), np.zeros(m, i32), np.zeros(m, i32), np.zeros(1, i32)
        K("cg3d_unique_first", np.ascontiguousarray(c, i32), m, keys, vals, cap, out, first, inv, nu,
          np.zeros(3 * m + K("cg3d_scan_workspace_ints", m), i32))
        return out[:nu[0]].copy(), first[:nu[0]].copy(), inv, keys, vals, cap
    u, first, inv, keys, vals, cap = unique(coords)
    wu, winv, wfirst = me.unique_first(want_c)
    assert np.array_equal(u, wu) and np.array_equal(inv, winv) and np.array_equal(first, wfirst) and len(u) < n
    rows = np.zeros(len(u) + 3, i32)
    query = np.concatenate([u, [[0, 500, 0, 0], [1, -31, 2, 900], [5, 0, 0, 0]]]).astype(i32)
    K("cg3d_hash_lookup", query, len(query), keys, vals, cap, rows)
    assert np.array_equal(rows[:len(u)], np.arange(len(u))) and np.all(rows[len(u):] == -1)
    in_map = me.CoordMap(wu, 1)

    def table(rules, n_out):
        t = np.full((len(rules), n_out), -1, i32)
        for k_, (i_, o_) in enumerate(rules):
            t[k_, o_] = i_
        return t
    nbr = np.zeros((27, len(u)), i32)
    K("cg3d_neighbor_table", u, len(u), keys, vals, cap, 3, 1, nbr)
    want_t = table(me.kernel_map(in_map, wu, 3, 1), len(u))
    assert np.array_equal(nbr, want_t) and (nbr >= 0).sum() > len(u)
    nbr_s = np.full((27, len(u)), -5, i32)
    K("cg3d_neighbor_table_symmetric", u, len(u), keys, vals, cap, 3, 1, nbr_s)
    assert np.array_equal(nbr_s, want_t)
    # stride-2 map (A6) and the 2^3 / stride-2 down-sampling rule map onto it
    c2 = np.zeros((len(u), 4), i32)
    K("cg3d_stride_coords", u, len(u), 2, c2)
    assert np.array_equal(c2[:, 1:], (wu[:, 1:] // 2) * 2)
    u2, _, _, _, _, _ = unique(c2)
    assert np.array_equal(u2, me.unique_first(c2.astype(np.int64))[0])
    nbr2 = np.zeros((8, len(u2)), i32)
    K("cg3d_neighbor_table", u2, len(u2), keys, vals, cap, 2, 1, nbr2)
    assert np.array_equal(nbr2, table(me.kernel_map(in_map, u2.astype(np.int64), 2, 1), len(u2)))
    assert (nbr2 >= 0).sum() == len(u)                           # every fine voxel feeds exactly one coarse voxel
