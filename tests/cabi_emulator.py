"""CPU emulation of the C-ABI entry points the TRAINING path calls (TEST INFRASTRUCTURE, like oracle/: the product never
imports this).

Purpose: the host side of the training path (cagroup3d_b200/autograd.py, backbone_train.py, train_targets.py) is Python
that can only run against the CUDA library.  To check that host logic on a machine without a GPU -- argument order of every
call, shapes, the autograd wiring (which gradient goes where, how many `None`s), the order of the layers -- `install()`
replaces `_lib.call` by a dispatcher that binds the positional arguments to the PARAMETER NAMES parsed from
include/cagroup3d_b200.h and runs a plain torch restatement of each entry point's documented contract on CPU tensors.  A
call with the wrong number / order of arguments fails here exactly as it would corrupt memory on the GPU.

Coordinate maps are built with the oracle (oracle/me_cpu.py): `cpu_map()` makes a `sparse.CoordMap` whose "hash table" is
simply (sorted keys, rows), which the emulated lookups understand; `sparse.strided_map / neighbor_table /
transpose_table` are replaced by oracle-backed versions.  What this does NOT check is the CUDA code itself -- that is
what the `-m gpu` parity tests are for.
"""
from __future__ import annotations

import re

import numpy as np
import torch

from oracle import backward_oracle as Bk
from oracle import me_cpu as me

_ACT = {0: lambda v: v, 1: torch.relu, 2: torch.nn.functional.elu}


def _param_names():
    from cagroup3d_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+(cg3d_\w+)\s*\(([^)]*)\)\s*;", src):
        names = [a.strip().split()[-1].lstrip("*") for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = [n for n in names if n != "stream"]
    return out


# ---- coordinate maps ---------------------------------------------------------------------------------------------------
_ME = {}            # id(keys tensor) -> me.CoordMap


def cpu_map(coords: np.ndarray, stride: int, mgr):
    from cagroup3d_b200 import sparse as S
    cm = me.CoordMap(np.asarray(coords, np.int64), stride)
    keys = torch.from_numpy(me.pack(cm.coords)) if len(cm) else torch.zeros((0,), dtype=torch.int64)
    m = S.CoordMap(torch.from_numpy(cm.coords.astype(np.int32)), stride, keys, torch.arange(len(cm), dtype=torch.int32), mgr.new_uid())
    _ME[id(keys)] = cm
    m._me = cm
    return m


def _me_of_keys(keys) -> me.CoordMap:
    return _ME[id(keys)]


def _strided_map(x_map, mgr, s):
    ts = x_map.stride * s
    if ts not in mgr.by_stride:
        c = x_map._me.coords.copy()
        c[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
        mgr.by_stride[ts] = cpu_map(me.unique_first(c)[0], ts, mgr)
    return mgr.by_stride[ts]


def _neighbor_table(in_map, out_map, k, mgr, ordered=False, **kw):
    rules = me.kernel_map(in_map._me, out_map._me.coords, k, in_map.stride)
    nbr = torch.from_numpy(Bk.rules_to_table(rules, max(out_map.n, 1)).astype(np.int32))
    if not ordered:
        return nbr
    # a positional table in a non-trivial order, so the out_rows plumbing is exercised; with group_div (grouped convs:
    # the class is batch // group_div) positions stay class-major, as the real tile orders keep them
    perm = np.random.default_rng(out_map.n + k).permutation(out_map.n)
    if kw.get("group_div"):
        perm = perm[np.argsort(out_map._me.coords[perm, 0] // kw["group_div"], kind="stable")]
    perm = torch.from_numpy(perm.astype(np.int32))
    return nbr[:, perm.long()].contiguous(), perm


def _transpose_table(in_map, fine_map, k, mgr, ordered=False, **kw):
    if k == 3:          # generative k3 s3 onto given coordinates (A13), the oracle's child <-> tap assignment
        f = fine_map._me.coords
        r = np.mod(f[:, 1:], 3)
        off = np.where(r == 0, 0, np.where(r == 1, 1, -1))
        c = f.copy()
        c[:, 1:] = f[:, 1:] - off
        tap = (off[:, 0] + 1) + 3 * ((off[:, 1] + 1) + 3 * (off[:, 2] + 1))
        nbr = np.full((27, max(len(f), 1)), -1, np.int32)
        nbr[tap, np.arange(len(f))] = in_map._me.lookup(c)
        nbr = torch.from_numpy(nbr)
        return (nbr, None) if ordered else nbr
    assert k == 2
    f = fine_map._me.coords
    ts_c, ts_f = in_map.stride, fine_map.stride
    p = f.copy()
    p[:, 1:] = np.floor_divide(f[:, 1:], ts_c) * ts_c
    d = (f[:, 1:] - p[:, 1:]) // ts_f
    tap = d[:, 0] + 2 * (d[:, 1] + 2 * d[:, 2])
    rows = in_map._me.lookup(p)
    nbr = np.full((8, len(f)), -1, np.int32)
    nbr[tap, np.arange(len(f))] = rows
    nbr = torch.from_numpy(nbr)
    return (nbr, None) if ordered else nbr


# ---- entry points ------------------------------------------------------------------------------------------------------
def _rows(n, out_rows):
    return out_rows.long() if out_rows is not None else torch.arange(n)


def cg3d_spconv_simt(**a):
    x, W, out = a["in"], a["W"], a["out"]
    n_out, Cin, Cout, K = a["n_out"], a["Cin"], a["Cout"], a["K"]
    assert x.shape[1] == Cin and out.shape == (n_out, Cout) and a["ldi"] == x.stride(0) and a["ldo"] == out.stride(0)
    W = W.reshape(-1, K, Cin, Cout)
    xin = _ACT[a["in_act"]](x)
    if a["tile_row0"] is None:
        assert a["n_tiles"] == 0 and W.shape[0] == 1
        tiles = [(0, n_out, 0)]
    else:                                   # grouped mode: only the positions covered by a tile are written
        assert len(a["tile_row0"]) == len(a["tile_rows"]) == len(a["tile_group"]) == a["n_tiles"]
        assert int(a["tile_rows"].max()) <= 64, "the SIMT kernel's tile holds 64 rows"
        tiles = list(zip(a["tile_row0"].tolist(), a["tile_rows"].tolist(), a["tile_group"].tolist()))
    for p0, np_, g in tiles:
        pos = torch.arange(p0, p0 + np_)
        rows = a["out_rows"].long()[pos] if a["out_rows"] is not None else pos
        acc = torch.zeros((np_, Cout), dtype=x.dtype)
        for k in range(K):
            if a["nbr"] is None:
                acc += xin[rows] @ W[g, k]
                continue
            assert a["nbr"].shape == (K, max(n_out, 1))
            src = a["nbr"][k, pos].long()
            hit = src >= 0
            acc[hit] += xin[src[hit]] @ W[g, k]
        if a["scale"] is not None:
            acc = acc * a["scale"].reshape(-1, Cout)[g]
        if a["shift"] is not None:
            acc = acc + a["shift"].reshape(-1, Cout)[g]
        if a["residual"] is not None:
            acc = acc + a["residual"][rows]
        out[rows] = _ACT[a["act"]](acc)


def cg3d_affine_act(**a):
    x, out = a["x"], a["out"]
    assert x.shape == (a["n"], a["C"]) and out.shape == x.shape
    v = x
    if a["scale"] is not None:
        v = v * a["scale"]
    if a["shift"] is not None:
        v = v + a["shift"]
    if a["add"] is not None:
        v = v + a["add"]
    out.copy_(_ACT[a["act"]](v))


def cg3d_table_transpose(**a):
    nbr, T = a["nbr"], a["nbrT"]
    assert nbr.shape == (a["K"], a["n_cols"]) and T.shape[0] == a["K"] and T.shape[1] >= a["n_in"]
    rows = a["out_rows"].numpy() if a["out_rows"] is not None else None
    T[:, :a["n_in"]] = torch.from_numpy(Bk.table_transpose(nbr.numpy(), a["n_in"], rows).astype(np.int32))


def cg3d_transpose_weights(**a):
    W, Wt = a["W"], a["Wt"]
    Wt.copy_(W.reshape(a["n_mats"], a["Cin"], a["Cout"]).transpose(1, 2).reshape(Wt.shape))


def cg3d_spconv_wgrad(**a):
    x, dy, dW = a["x"], a["dy"], a["dW"]
    K, Cin, Cout = a["K"], a["Cin"], a["Cout"]
    assert x.shape[1] == Cin and dy.shape[1] == Cout and dW.shape == (K, Cin, Cout)
    assert a["slabs"] is None or a["slabs"].numel() >= K * Cin * Cout
    xin = _ACT[a["in_act"]](x)
    cols = torch.arange(a["col0"], a["col1"])
    rows = a["out_rows"].long()[cols] if a["out_rows"] is not None else cols
    for k in range(K):
        src = a["nbr"][k, cols].long() if a["nbr"] is not None else cols
        hit = src >= 0
        dW[k] = xin[src[hit]].T @ dy[rows[hit]]


def _split_image(v: torch.Tensor) -> torch.Tensor:
    """fp32 [n][C] -> int16 [n][2C]: per 32-channel chunk [hi 32 | lo 32] bf16 bit patterns (cg3d_split_bf16's layout)."""
    n, C = v.shape
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    img = torch.stack([hi.view(n, C // 32, 32), lo.view(n, C // 32, 32)], dim=2)      # [n][C/32][2][32]
    return img.reshape(n, 2 * C).view(torch.int16)


def _unsplit_image(img: torch.Tensor, C: int) -> torch.Tensor:
    n = img.shape[0]
    parts = img[:, :2 * C].contiguous().view(torch.bfloat16).view(n, C // 32, 2, 32).float()
    return (parts[:, :, 0] + parts[:, :, 1]).reshape(n, C)


def cg3d_split_bf16(**a):
    x, n, C = a["in"], a["n"], a["C"]
    assert C % 32 == 0 and a["ld"] == x.stride(0) and a["out"].dtype == torch.int16 and a["out"].shape[1] == 2 * C
    v = x[:n, :C].float()
    if a["relu"]:
        v = torch.relu(v)
    a["out"][:n] = _split_image(v)


def cg3d_spconv_wgrad_tc(**a):
    """the tensor-core weight gradient: the same sums as cg3d_spconv_wgrad over the split-bf16 operands (hi + lo each)"""
    K, Cin, Cout = a["K"], a["Cin"], a["Cout"]
    assert Cin % 64 == 0 and Cout % 64 == 0 and a["dW"].shape == (K, Cin, Cout)
    x, dy = _unsplit_image(a["x_split"], Cin), _unsplit_image(a["dy_split"], Cout)
    cg3d_spconv_wgrad(x=x, dy=dy, dW=a["dW"], K=K, Cin=Cin, Cout=Cout, slabs=a["slabs"], in_act=0, col0=a["col0"], col1=a["col1"],
                      out_rows=a["out_rows"], nbr=a["nbr"])


def cg3d_bn_train_stats(**a):
    x, n = a["x"], a["n"]
    assert x.shape == (n, a["C"])
    mean, var = x.mean(0), x.var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + a["eps"])
    a["mean"].copy_(mean)
    a["rstd"].copy_(rstd)
    g = a["gamma"] if a["gamma"] is not None else torch.ones_like(mean)
    b = a["beta"] if a["beta"] is not None else torch.zeros_like(mean)
    if a["scale"] is not None:
        a["scale"].copy_(g * rstd)
    if a["shift"] is not None:
        a["shift"].copy_(b - mean * g * rstd)
    mom = a["momentum"]
    if a["running_mean"] is not None:
        a["running_mean"].mul_(1 - mom).add_(mom * mean)
    if a["running_var"] is not None:
        a["running_var"].mul_(1 - mom).add_(mom * (x.var(0, unbiased=True) if n > 1 else var))


def cg3d_bn_train_backward(**a):
    x, dy, n = a["x"], a["dy"], a["n"]
    if a["y_mask"] is not None:
        dy = torch.where(a["y_mask"] > 0, dy, torch.zeros_like(dy))
    xh = (x - a["mean"]) * a["rstd"]
    dbeta, dgamma = dy.sum(0), (dy * xh).sum(0)
    g = a["gamma"] if a["gamma"] is not None else torch.ones_like(dbeta)
    a["dx"].copy_(g * a["rstd"] * (dy - dbeta / n - xh * dgamma / n))
    a["dgamma"].copy_(dgamma)
    a["dbeta"].copy_(dbeta)
    if a["dres"] is not None:
        a["dres"].copy_(dy)


def cg3d_interp_trilinear(**a):
    src = _me_of_keys(a["keys"])
    q = a["query"].numpy().astype(np.int64)[:a["nq"]]
    rows, w = Bk.interp_corners(src, q)
    out = a["base"].clone() if a["base"] is not None else torch.zeros((a["nq"], a["C"]), dtype=a["feats"].dtype)
    for c in range(8):
        hit = np.nonzero(rows[:, c] >= 0)[0]
        if len(hit):
            out[torch.from_numpy(hit)] += a["feats"][torch.from_numpy(rows[hit, c])] * torch.from_numpy(w[hit, c]).to(out.dtype)[:, None]
    a["out"].copy_(out)


def cg3d_interp_trilinear_backward(**a):
    qm = _me_of_keys(a["qkeys"])
    assert qm.stride == a["tq"] and a["ts"] % a["tq"] == 0
    srcm = me.CoordMap(a["src_coords"].numpy().astype(np.int64)[:a["n_src"]], a["ts"])
    rows, w = Bk.interp_corners(srcm, qm.coords)
    a["dF"].copy_(Bk.interp_backward(a["dOut"], rows, w, a["n_src"]).to(a["dF"].dtype))


def _window(oc, ic, half):
    d = (ic[None, :, 1:] - oc[:, None, 1:]).abs().max(-1).values
    return ((d <= half) & (ic[None, :, 0] == oc[:, None, 0])).float()


def cg3d_avgpool_window(**a):
    M = _window(a["out_coords"][:a["n_out"]].long(), a["in_coords"][:a["n_in"]].long(), a["half"]).to(a["feats"].dtype)
    a["out"].copy_((M @ a["feats"]) / M.sum(1, keepdim=True).clamp(min=1))


def cg3d_avgpool_window_backward(**a):
    M = _window(a["out_coords"][:a["n_out"]].long(), a["in_coords"][:a["n_in"]].long(), a["half"]).to(a["dOut"].dtype)
    cnt = M.sum(1)
    a["counts"][:a["n_out"]] = cnt
    a["dIn"].copy_(M.T @ (a["dOut"] / cnt[:, None]))


def cg3d_act_backward(**a):
    dy, y = a["dy"], a["y"]
    if a["act"] == 1:
        d = torch.where(y > 0, dy, torch.zeros_like(dy))
    elif a["act"] == 2:
        d = torch.where(y > 0, dy, dy * (y + 1))
    else:
        d = dy
    a["dx"].copy_(d)


def cg3d_segment_mean_backward(**a):
    inv = a["inverse"].long()
    a["dIn"].copy_(a["dOut"][inv] / a["counts"][inv][:, None])


def cg3d_assign(**a):
    from oracle import train_oracle as T
    offs = a["cls_offsets"].tolist()
    pts = [a["locs"][offs[c]:offs[c + 1]] for c in range(a["n_cls"])]
    ct, bt, lb = T.assign(pts, a["gt_boxes"], a["gt_labels"].long(), a["topk"])
    a["centerness"].copy_(torch.nan_to_num(ct))
    a["box_targets"].copy_(bt)
    a["labels"].copy_(lb)
    if a["box_index"] is not None:
        a["box_index"].fill_(-1)


def cg3d_assign_semantic(**a):
    from oracle import train_oracle as T
    sl, il = T.assign_semantic(a["points"], a["gt_boxes"], a["gt_labels"].long())
    a["labels"].copy_(sl)
    a["ins_labels"].copy_(il)


def cg3d_focal_loss(**a):
    from oracle import train_oracle as T
    p = a["pred"].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = T.focal_loss(p, a["labels"], a["avg_factor"], a["gamma"], a["alpha"])
    a["loss"][0] = float(loss)
    if a["grad"] is not None:
        a["grad"].copy_(torch.autograd.grad(loss, p)[0])


def cg3d_bce_loss(**a):
    from oracle import train_oracle as T
    p = a["pred"].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = T.bce_loss(p, a["target"], a["avg_factor"])
    a["loss"][0] = float(loss)
    if a["grad"] is not None:
        a["grad"].copy_(torch.autograd.grad(loss, p)[0])


def cg3d_iou_loss_aa(**a):
    from oracle import train_oracle as T
    p = a["pred"].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = T.axis_aligned_iou_loss(p[:, :6], a["target"][:, :6], a["weight"], a["avg_factor"])
    a["loss"][0] = float(loss)
    if a["grad"] is not None:
        a["grad"].copy_(torch.autograd.grad(loss, p)[0])


def cg3d_smooth_l1_loss(**a):
    from oracle import train_oracle as T
    p = a["pred"].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = T.smooth_l1_sum(p, a["target"], a["weight"], a["beta"])
    a["loss"][0] = float(loss)
    if a["grad"] is not None:
        a["grad"].copy_(torch.autograd.grad(loss, p)[0])


def cg3d_knn(**a):
    assert a["xyz"].shape == (a["b"], a["n"], 3) and a["query"].shape == (a["b"], a["m"], 3) and a["k"] == 1
    d = torch.cdist(a["query"].double(), a["xyz"].double())
    a["idx"].copy_(d.argmin(2, keepdim=True).int())
    a["dist2"].copy_((d.min(2, keepdim=True).values ** 2).float())


def cg3d_vote_targets(**a):
    sp, ins, sem = a["scene_points"][:, :3], a["ins_mask"], a["sem_mask"]
    assert sp.shape[0] == a["n"] and a["ld"] == a["scene_points"].stride(0) and a["workspace"].numel() >= 8 * a["n_inst"]
    centers = torch.zeros((a["n_inst"], 3))
    for i in torch.unique(ins):
        idx = torch.nonzero(ins == i).squeeze(1)
        if sem[idx[0]] < a["n_classes"]:
            c = 0.5 * (sp[idx].min(0)[0] + sp[idx].max(0)[0])
            centers[i] = a["gt_boxes"][torch.argmin(torch.cdist(c.view(1, 3), a["gt_boxes"][:, :3]).view(-1)), :3]
        else:
            centers[i] = -10000.0
    a["centers"].copy_(centers)
    t = centers[ins[a["nearest"].long()]] - a["voxel_points"]
    a["mask"].copy_((t >= -100.0).all(1).float())
    a["targets"].copy_(torch.where(t < -100.0, torch.zeros_like(t), t))


def cg3d_column_sum(**a):
    assert a["x"].shape == (a["n"], a["C"]) and a["ldx"] == a["x"].stride(0)
    a["out"].copy_(a["x"].sum(0))


def cg3d_segment_mean(**a):
    assert a["ref"] is None and a["srcB"] is None, "only the plain (ref == NULL) form is emulated"
    src, inv, U, C = a["srcA"], a["inverse"].long(), a["n_unique"], a["C"]
    assert src.shape == (a["n"], C) and a["ldA"] == src.stride(0)
    assert a["workspace"].dtype == torch.int64 and a["workspace"].numel() >= (3 * (U + 1) + a["n"]) // 2   # counts | offsets | cursors | point list
    cnt = torch.bincount(inv, minlength=U).to(src.dtype)
    a["counts"][:U] = cnt
    a["out"].copy_(torch.zeros((U, C), dtype=src.dtype).index_add_(0, inv, src) / cnt[:, None])


def cg3d_first_rows(**a):
    """first[b] = first row of sample b.  (Always served from here: the kernel groups the lanes of a warp with
    __match_any_sync(__activemask(), ...) inside its row loop, which the CPU harness can only run for whole warps -- it is
    checked natively in tests/test_cuda_on_cpu.py::test_head_coordinate_kernels with a row count that is a multiple of 32.)"""
    b = a["coords"][:a["n"], 0].long()
    for k in range(a["B"]):
        r = torch.nonzero(b == k).flatten()
        a["first"][k] = int(r[0]) if len(r) else 0x7F7F7F7F


def cg3d_sort_pairs(**a):
    n, b0, b1 = a["n"], a["begin_bit"], a["end_bit"]
    k = a["keys"][:n]
    sub = (k >> b0) & ((1 << (b1 - b0)) - 1) if b1 - b0 < 63 else k
    perm = torch.argsort(sub, stable=True)
    a["keys"][:n] = k[perm]
    a["vals"][:n] = a["vals"][:n][perm]


def cg3d_histogram_i32(**a):
    ids = a["ids"][:a["n"]].long()
    a["counts"][:a["nseg"]] = torch.bincount(ids, minlength=a["nseg"])[:a["nseg"]].int()


def cg3d_exclusive_scan_i32(**a):
    params = list(a)
    flags, n = a[params[0]], a[params[1]]
    out, total = a[params[2]], a[params[-1]]
    c = torch.cumsum(flags[:n].long(), 0)
    out[:n] = (c - flags[:n].long()).int()
    total[0] = int(c[-1]) if n else 0


def cg3d_segment_sum_sorted(**a):
    src, order, off = a["src"], a["order"].long(), a["seg_off"].tolist()
    out = torch.zeros((a["n_seg"], a["C"]), dtype=src.dtype)
    for u in range(a["n_seg"]):
        if off[u + 1] > off[u]:
            out[u] = src[order[off[u]:off[u + 1]]].sum(0)
    a["out"].copy_(out)


def cg3d_boxes_pairwise_bev(**a):
    from oracle import iou3d_oracle
    assert a["boxes_a"].shape == (a["na"], 7) and a["boxes_b"].shape == (a["nb"], 7)
    a["out"].copy_(iou3d_oracle.pairwise(a["boxes_a"], a["boxes_b"], {0: "overlap", 1: "iou", 2: "iou_normal"}[a["mode"]]))


class _Setter:
    """monkeypatch stand-in for a process that ends with the test (spawned gloo workers)"""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def install(monkeypatch=None, compiled: bool = False, native_maps: bool = False):
    """route _lib.call to the emulation and the coordinate-map builders to the oracle (for the duration of a test;
    monkeypatch=None: for the rest of the process).

    compiled=True: the entry points of the training kernels are NOT emulated but served by their own CUDA sources compiled
    for the CPU (tests/cuda_on_cpu), called exactly as _lib.call calls the GPU library -- raw data pointers, sizes and
    strides as passed -- so a wrong stride / a non-contiguous view / a wrong dtype handed over by the Python side shows up
    as a wrong result here instead of on the device.

    native_maps=True (with compiled): the coordinate maps are the product's own too -- hash tables built, probed and uniqued
    by csrc/coords.cu on the CPU; nothing of sparse.py is replaced (the inference-side tests, tests/test_inference_wiring_cpu.py)."""
    from cagroup3d_b200 import _lib, sparse as S
    assert compiled or not native_maps
    monkeypatch = monkeypatch or _Setter
    native, protos = None, None
    if compiled:
        import ctypes
        from tests.cuda_on_cpu import build as cpu_build
        native, protos = cpu_build.load(), _lib.parse_header()
    names = _param_names()
    table = {k: v for k, v in globals().items() if k.startswith("cg3d_")}
    calls = []

    def call(name, *args):
        params = names[name]
        assert len(args) == len(params), (name, len(args), params)
        calls.append(name)
        # (entry points that probe a coordinate map's REAL open-addressing table -- parameters `keys` / `capacity`, e.g. the
        # interpolation and its backward -- stay emulated: cpu_map() keeps a sorted key list instead of a hash table)
        if native is not None and hasattr(native, name) and name != "cg3d_first_rows" \
                and (native_maps or not ("capacity" in params or "qcapacity" in params)):
            fn = getattr(native, name)
            fn.argtypes, fn.restype = protos[name], ctypes.c_int
            for a in args:
                assert not isinstance(a, torch.Tensor) or a.device.type == "cpu"
            rc = fn(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args], None)
            assert rc == 0, (name, rc)
            return
        if name not in table:
            raise NotImplementedError(f"{name} is not emulated")
        table[name](**dict(zip(params, args)))

    monkeypatch.setattr(_lib, "call", call)
    if not native_maps:
        monkeypatch.setattr(S, "strided_map", _strided_map)
        monkeypatch.setattr(S, "neighbor_table", _neighbor_table)
        monkeypatch.setattr(S, "transpose_table", _transpose_table)
    return calls
