"""tcgen05 sparse-conv path (bf16x3 split, fp32 accumulate) against the CPU oracle and the exact SIMT path."""
import numpy as np
import pytest
import torch

from oracle import me_cpu as me
from tests.test_gpu_ops import oracle_tensor
from tests.util import to_gpu_sparse

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("cin,cout,k,stride", [(64, 64, 3, 1), (64, 128, 3, 2), (128, 256, 1, 1), (128, 128, 1, 2),
                                               (256, 512, 3, 1), (64, 64, 5, 1), (192, 64, 3, 1)])
def test_spconv_tc_vs_oracle(lib, cin, cout, k, stride):
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(21, cin, n=3000)
    g = torch.Generator().manual_seed(5)
    W = torch.randn((k ** 3, cin, cout), generator=g) / np.sqrt(cin * min(k ** 3, 8))
    scale, shift = torch.rand((cout,), generator=g) + 0.5, torch.randn((cout,), generator=g)
    ref = me.conv(ox.with_F(torch.relu(ox.F)), W[0] if (k == 1 and stride == 1) else W, k, stride)
    res = torch.randn((ref.F.shape[0], cout), generator=g)
    want = torch.nn.functional.elu(ref.F * scale + shift + res)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    Wd = W.to(DEV)
    if k == 1 and stride == 1:
        Wd = Wd[0].contiguous()
    y = S.conv(x, Wd, k, stride, scale=scale.to(DEV), shift=shift.to(DEV), residual=res.to(DEV), act="elu",
               in_act="relu", impl="tc")
    torch.cuda.synchronize()
    err = (y.F.cpu() - want).abs().max().item()
    ref_mag = want.abs().max().item()
    assert err <= 2e-4 * max(1.0, ref_mag), (err, ref_mag)


def test_spconv_tc_grouped_and_slices(lib):
    """grouped tiles (per-class weights) + output into a column slice of a concat buffer."""
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(22, 64, n=2500)
    n = ox.F.shape[0]
    g = torch.Generator().manual_seed(6)
    W = torch.randn((3, 27, 64, 64), generator=g) / 20
    x = to_gpu_sparse(ox.C, ox.F, 1)
    nbr = S.neighbor_table(x.cmap, x.cmap, 3, x.mgr)
    offs = [0, n // 3, n // 3 + 77, n]
    tiles = S.make_tiles(offs, DEV, 128)
    cat = torch.zeros((n, 128), device=DEV)
    S.gemm_rows(x.F, nbr, W.to(DEV), n, 27, tiles=tiles, out=cat[:, 64:], impl="tc")
    ref = torch.cat([me.conv(ox, W[i], 3, 1).F[offs[i]:offs[i + 1]] for i in range(3)])
    assert (cat[:, 64:].cpu() - ref).abs().max().item() <= 2e-4
    assert cat[:, :64].abs().max().item() == 0
    want = S.gemm_rows(x.F, nbr, W.to(DEV), n, 27, tiles=S.make_tiles(offs, DEV, 64), impl="simt")
    assert (cat[:, 64:] - want).abs().max().item() <= 2e-4


@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_tile_ordered_rule_map_gives_identical_rows(lib, impl):
    """Morton tile order: `order` is a permutation, the positional table is the row table permuted, and the conv
    result is the same matrix (same rows in the same places) as with the natural order."""
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(23, 64, n=6000, batch=3)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    n = x.cmap.n
    order, oc = S.tile_order(x.cmap, 4)
    assert order is not None and torch.equal(torch.sort(order.long())[0].cpu(), torch.arange(n))
    assert torch.equal(oc, x.cmap.coords[order.long()])
    b = oc[:, 0].cpu()
    assert (b[1:] >= b[:-1]).all()                                   # samples stay contiguous
    nbr_nat = S.neighbor_table(x.cmap, x.cmap, 3, x.mgr)
    S._TILE_ORDER["mode"], old = "morton", S._TILE_ORDER["mode"]
    try:
        nbr_ord, order2 = S.neighbor_table(x.cmap, x.cmap, 3, x.mgr, ordered=True)
    finally:
        S._TILE_ORDER["mode"] = old
    assert order2 is order and torch.equal(nbr_ord, nbr_nat[:, order.long()])
    g = torch.Generator().manual_seed(9)
    W = (torch.randn((27, 64, 64), generator=g) / 20).to(DEV)
    res = torch.randn((n, 64), generator=g).to(DEV)
    a = S.gemm_rows(x.F, nbr_nat, W, n, 27, residual=res, act="relu", impl=impl)
    o = S.gemm_rows(x.F, nbr_ord, W, n, 27, residual=res, act="relu", impl=impl, out_rows=order)
    assert torch.equal(a, o)                                          # same accumulation order per row -> bit identical


def test_split_rows_layout(lib):
    """cg3d_split_bf16: per 32-channel chunk [hi 32 | lo 32]; hi + lo reproduces x to 2^-16 relative; ReLU variant."""
    from cagroup3d_b200 import sparse as S
    g = torch.Generator().manual_seed(4)
    x = torch.randn((777, 96), generator=g).to(DEV)
    for act in (None, "relu"):
        s = S.split_rows(x, act).view(torch.bfloat16).view(777, 3, 2, 32).float()
        want = torch.relu(x) if act else x
        hi, lo = s[:, :, 0].reshape(777, 96), s[:, :, 1].reshape(777, 96)
        assert torch.equal(hi, want.to(torch.bfloat16).float())
        assert ((hi + lo) - want).abs().max().item() <= 2 ** -16 * want.abs().max().item()


@pytest.mark.parametrize("k,cout,grouped", [(5, 64, False), (5, 128, False), (9, 64, True), (7, 64, False), (3, 64, False)])
def test_spconv_pairs_vs_simt_and_oracle(lib, monkeypatch, k, cout, grouped):
    """pair-compacted tcgen05 kernel (wide kernels over thin maps) == exact fp32 SIMT kernel == CPU oracle, incl. grouped
    per-class weights, a ragged last tile, residual + ELU epilogue, the split-bf16 second output and rows without any pair."""
    from cagroup3d_b200 import sparse as S
    g = torch.Generator().manual_seed(7)
    # a ~30 % occupied volume (as the class maps of the head): many pairs per (tile, tap), the centre tap fills a tile
    c = torch.cat([torch.randint(0, 2, (2600, 1), generator=g), torch.randint(0, 20, (2600, 2), generator=g),
                   torch.randint(0, 9, (2600, 1), generator=g)], 1).float()
    ox = me.from_points(c, torch.randn((2600, 64), generator=g))
    n = ox.F.shape[0]
    G = 3 if grouped else 1
    W = torch.randn((G, k ** 3, 64, cout), generator=g) / np.sqrt(64 * 27)
    scale, shift = torch.rand((G, cout), generator=g) + 0.5, torch.randn((G, cout), generator=g)
    res = torch.randn((n, cout), generator=g)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    nbr = S.neighbor_table(x.cmap, x.cmap, k, x.mgr)
    offs = [0, n // 3, n // 3 + 77, n] if grouped else [0, n]
    Wd = W.to(DEV) if grouped else W[0].to(DEV)
    sc, sh = (scale.to(DEV), shift.to(DEV)) if grouped else (scale[0].to(DEV), shift[0].to(DEV))
    kw = dict(scale=sc, shift=sh, residual=res.to(DEV), act="elu", in_act="relu")
    assert lib.cg3d_spconv_pairs_supported(64, cout, k ** 3) == 1 and k ** 3 >= S.PAIRS_MIN_K
    monkeypatch.setitem(S._PAIRS, "on", True)
    monkeypatch.setattr(S, "PAIRS_MAX_COUT", 128)          # also the two-slice launch (not routed by default)
    assert S.pairs_route(nbr, 64, cout, k ** 3)
    got = S.gemm_rows(x.F, nbr, Wd, n, k ** 3, tiles=S.make_tiles(offs, DEV, 128) if grouped else None, impl="tc",
                      split_out="relu", **kw)
    want = S.gemm_rows(x.F, nbr, Wd, n, k ** 3, tiles=S.make_tiles(offs, DEV, 64) if grouped else None, impl="simt", **kw)
    torch.cuda.synchronize()
    mag = max(1.0, want.abs().max().item())
    assert (got - want).abs().max().item() <= 2e-4 * mag
    # split output = bf16 hi | lo of relu(result)
    sp = got._cg3d_split[(got.data_ptr(), got._version, got.stride(0), 1)].view(torch.bfloat16).view(n, cout // 32, 2, 32).float()
    assert ((sp[:, :, 0] + sp[:, :, 1]).reshape(n, cout) - torch.relu(got)).abs().max().item() <= 2 ** -15 * mag
    # oracle (per group)
    for i in range(G):
        ref = me.conv(ox.with_F(torch.relu(ox.F)), W[i], k, 1).F[offs[i]:offs[i + 1]]
        w = torch.nn.functional.elu(ref * scale[i] + shift[i] + res[offs[i]:offs[i + 1]])
        assert (got[offs[i]:offs[i + 1]].cpu() - w).abs().max().item() <= 2e-4 * mag
    # run-to-run deterministic
    again = S.gemm_rows(x.F, nbr, Wd, n, k ** 3, tiles=S.make_tiles(offs, DEV, 128) if grouped else None, impl="tc", **kw)
    assert torch.equal(again, got)
    # the default route for the same layer (row-stationary kernel) agrees as well
    monkeypatch.setitem(S._PAIRS, "on", False)
    rs = S.gemm_rows(x.F, nbr, Wd, n, k ** 3, tiles=S.make_tiles(offs, DEV, 128) if grouped else None, impl="tc", **kw)
    assert (rs - want).abs().max().item() <= 2e-4 * mag


def test_spconv_pairs_at_query_coordinates(lib, monkeypatch):
    """conv evaluated at foreign query coordinates (RoI grid conv, A12): most queries have few or no neighbours."""
    from cagroup3d_b200 import sparse as S
    monkeypatch.setitem(S._PAIRS, "on", True)
    monkeypatch.setattr(S, "PAIRS_MAX_COUT", 128)
    ox = oracle_tensor(33, 64, n=1500)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    g = torch.Generator().manual_seed(8)
    q = torch.cat([torch.zeros((900, 1), dtype=torch.int32), torch.randint(-6, 40, (900, 3), generator=g, dtype=torch.int32)], 1)
    q = torch.unique(q, dim=0).to(DEV)
    qmap = S.build_map(q.contiguous(), 1, x.mgr)
    nbr = S.neighbor_table(x.cmap, qmap, 5, x.mgr)
    W = (torch.randn((125, 64, 128), generator=g) / 40).to(DEV)
    got = S.gemm_rows(x.F, nbr, W, qmap.n, 125, act="elu", impl="tc")
    want = S.gemm_rows(x.F, nbr, W, qmap.n, 125, act="elu", impl="simt")
    assert (got - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    empty = (nbr < 0).all(0)
    assert empty.any() and got[empty].abs().max().item() == 0


@pytest.mark.parametrize("k,coarse", [(3, False), (5, True)])
def test_mask_ordered_rule_map_gives_identical_rows(lib, monkeypatch, k, coarse):
    """tap-pattern tile order (27-tap maps) and coarse tap-block order (RoI grid conv): `order` is a permutation, the
    positional table is the row table permuted, and the conv writes the same matrix as with the natural order."""
    from cagroup3d_b200 import sparse as S
    monkeypatch.setattr(S, "MASK_MIN_ROWS", 256)
    ox = oracle_tensor(51, 64, n=6000, batch=2)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    n = x.cmap.n
    nbr_nat = S.neighbor_table(x.cmap, x.cmap, k, x.mgr)
    nbr_ord, order = S.neighbor_table(x.cmap, x.cmap, k, x.mgr, ordered=True, spatial=False, coarse_mask=coarse)
    assert order is not None and torch.equal(torch.sort(order.long())[0].cpu(), torch.arange(n))
    assert torch.equal(nbr_ord, nbr_nat[:, order.long()])
    g = torch.Generator().manual_seed(9)
    W = (torch.randn((k ** 3, 64, 64), generator=g) / 20).to(DEV)
    res = torch.randn((n, 64), generator=g).to(DEV)
    a = S.gemm_rows(x.F, nbr_nat, W, n, k ** 3, residual=res, act="relu", impl="tc")
    o = S.gemm_rows(x.F, nbr_ord, W, n, k ** 3, residual=res, act="relu", impl="tc", out_rows=order)
    assert torch.equal(a, o)
