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


@pytest.mark.parametrize("cout", [128, 64])
def test_conv_at_query_coordinates(lib, cout):
    """conv evaluated at foreign query coordinates (RoI grid conv, A12): most queries have few or no neighbours.  Cout = 64
    runs the TMEM-operand kernel (spconv_ts.cu, no stash: K = 125), Cout = 128 the shared-memory one."""
    from cagroup3d_b200 import sparse as S
    ox = oracle_tensor(33, 64, n=1500)
    x = to_gpu_sparse(ox.C, ox.F, 1)
    g = torch.Generator().manual_seed(8)
    q = torch.cat([torch.zeros((900, 1), dtype=torch.int32), torch.randint(-6, 40, (900, 3), generator=g, dtype=torch.int32)], 1)
    q = torch.unique(q, dim=0).to(DEV)
    qmap = S.build_map(q.contiguous(), 1, x.mgr)
    nbr = S.neighbor_table(x.cmap, qmap, 5, x.mgr)
    W = (torch.randn((125, 64, cout), generator=g) / 40).to(DEV)
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
