"""CPU: the oracle's restatement of the MinkowskiEngine ops against DENSE PyTorch ops on the densified grid.

MinkowskiEngine defines a sparse (generalised) convolution as the dense convolution restricted to the active
output sites with absent inputs contributing zero; pooling / transposed conv / interpolation likewise.  These
tests check oracle/me_cpu.py against torch.nn.functional.{conv3d, conv_transpose3d, grid_sample} -- an
independent implementation -- including the tap order (x fastest), the non-centred even kernel (A5/A8),
the stride-2 output lattice (A6), the k=1 stride-2 shortcut (A7), non-zero-count average pooling (A10), trilinear
sampling with absent corners = 0 (A9) and the generative k3s3 transpose (A13)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from oracle import me_cpu as me

G = 24          # dense grid edge (coordinates live in [0, G))


def cloud(seed, n=900, C=5, batch=2, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    c = torch.cat([torch.randint(0, batch, (n, 1), generator=g), torch.randint(2, G - 2, (n, 3), generator=g)], 1)
    c[:, 3] = (c[:, 3] // 3) * 3 // 2 + 4                      # make it surface-like / clustered
    x = me.from_points(c.double(), torch.randn((n, C), generator=g, dtype=dtype))
    return x


def densify(x, batch=2, size=G, unit=1):
    C = x.F.shape[1]
    d = torch.zeros((batch, C, size, size, size), dtype=x.F.dtype)
    c = torch.from_numpy(x.C)
    d[c[:, 0], :, c[:, 3] // unit, c[:, 2] // unit, c[:, 1] // unit] = x.F
    return d


def sample(d, coords, unit=1):
    c = torch.from_numpy(coords)
    return d[c[:, 0], :, c[:, 3] // unit, c[:, 2] // unit, c[:, 1] // unit]


def dense_weight(W, k):
    """(k^3, Cin, Cout) x-fastest taps -> conv3d weight (Cout, Cin, kz, ky, kx)."""
    return W.view(k, k, k, W.shape[1], W.shape[2]).permute(4, 3, 0, 1, 2).contiguous()


@pytest.mark.parametrize("k", [3, 5])
def test_conv_stride1_equals_dense(k):
    x = cloud(1)
    W = torch.randn((k ** 3, 5, 7), dtype=torch.float64)
    y = me.conv(x, W, k, 1)
    want = sample(Fn.conv3d(densify(x), dense_weight(W, k), padding=k // 2), y.C)
    assert (y.C == x.C).all()
    assert (y.F - want).abs().max() < 1e-10


def test_conv_stride2_lattice_and_values():
    x = cloud(2)
    W = torch.randn((27, 5, 4), dtype=torch.float64)
    y = me.conv(x, W, 3, 2)
    assert y.cmap.stride == 2 and (y.C[:, 1:] % 2 == 0).all()
    want_set = {tuple(r) for r in np.concatenate([x.C[:, :1], x.C[:, 1:] // 2 * 2], 1).tolist()}
    assert {tuple(r) for r in y.C.tolist()} == want_set and len(y.C) == len(want_set)
    dense = Fn.conv3d(densify(x), dense_weight(W, 3), padding=1, stride=2)
    assert (y.F - sample(dense, y.C, unit=2)).abs().max() < 1e-10
    # a second strided op from the same input reuses the cached map (A6)
    assert me.conv(x, W, 3, 2).cmap is y.cmap


def test_k1_stride2_shortcut_only_where_the_coordinate_exists():
    x = cloud(3)
    W = torch.randn((1, 5, 6), dtype=torch.float64)
    y = me.conv(x, W, 1, 2)
    present = x.cmap.lookup(y.C)
    want = torch.zeros_like(y.F)
    hit = np.nonzero(present >= 0)[0]
    want[hit] = x.F[present[hit]] @ W[0]
    assert (y.F - want).abs().max() < 1e-12 and 0 < len(hit) < len(y.C)


def test_transpose_k2s2_equals_dense():
    x = cloud(4)
    coarse = me.conv(x, torch.randn((27, 5, 3), dtype=torch.float64), 3, 2)          # stride-2 tensor, map cached
    W = torch.randn((8, 3, 4), dtype=torch.float64)
    up = me.conv_transpose_k2s2(coarse, W)
    assert (up.C == x.C).all()
    wd = W.view(2, 2, 2, 3, 4).permute(3, 4, 0, 1, 2).contiguous()                   # (Cin, Cout, kz, ky, kx)
    dense = Fn.conv_transpose3d(densify(coarse, size=G // 2, unit=2), wd, stride=2)
    assert (up.F - sample(dense, up.C)).abs().max() < 1e-10


def test_generative_transpose_k3s3_equals_dense():
    g = torch.Generator().manual_seed(5)
    fine = cloud(5)
    ec = torch.from_numpy(np.concatenate([fine.C[:, :1], fine.C[:, 1:] // 3 * 3], 1))
    E = me.from_points(ec.double(), torch.randn((len(ec), 5), generator=g, dtype=torch.float64), average=True, stride=3)
    W = torch.randn((27, 5, 4), dtype=torch.float64)
    out = me.generative_transpose_k3s3(E, W, fine.cmap)
    wd = W.view(3, 3, 3, 5, 4).permute(3, 4, 0, 1, 2).contiguous()
    dense = Fn.conv_transpose3d(densify(E, size=G // 3, unit=3), wd, stride=3, padding=1, output_padding=0)
    # dense output index p <-> fine coordinate p (padding 1 centres tap 1 on the coarse voxel)
    c = torch.from_numpy(fine.C)
    ok = (c[:, 1:] < dense.shape[-1]).all(1)
    want = dense[c[ok, 0], :, c[ok, 3], c[ok, 2], c[ok, 1]]
    assert ok.sum() > 100 and (out[ok] - want).abs().max() < 1e-10


@pytest.mark.parametrize("k,s", [(5, 2), (9, 4)])
def test_avg_pool_counts_existing_inputs(k, s):
    x = cloud(6)
    y = me.avg_pool(x, k, s)
    d = densify(x)
    occ = densify(me.SparseTensor(torch.ones((len(x.C), 1), dtype=torch.float64), x.cmap, x.mgr))
    ones = torch.ones((1, 1, k, k, k), dtype=torch.float64)
    num = Fn.conv3d(d.view(-1, 1, G, G, G), ones, padding=k // 2, stride=s).view(2, 5, *[G // s] * 3)
    den = Fn.conv3d(occ, ones, padding=k // 2, stride=s)
    want = sample(num / den.clamp(min=1), y.C, unit=s)
    assert (y.C[:, 1:] % s == 0).all() and (y.F - want).abs().max() < 1e-10


def test_features_at_coordinates_is_trilinear_with_zero_fill():
    x = cloud(7)
    coarse = me.conv(x, torch.randn((27, 5, 6), dtype=torch.float64), 3, 2)
    coarse = me.conv(coarse, torch.randn((27, 6, 6), dtype=torch.float64), 3, 2)     # stride 4
    q = x.C
    got = me.features_at(coarse, q)
    n = G // 4 + 1
    d = torch.zeros((2, 6, n, n, n), dtype=torch.float64)
    c = torch.from_numpy(coarse.C)
    d[c[:, 0], :, c[:, 3] // 4, c[:, 2] // 4, c[:, 1] // 4] = coarse.F
    for b in range(2):
        rows = np.nonzero(q[:, 0] == b)[0]
        idx = torch.from_numpy(q[rows, 1:]).double() / 4                              # (x, y, z) in coarse index units
        grid = (2 * idx / (n - 1) - 1).view(1, -1, 1, 1, 3)
        want = Fn.grid_sample(d[b:b + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        assert (got[rows] - want.view(6, -1).T).abs().max() < 1e-10


def test_unique_first_and_average_quantisation():
    c = np.array([[0, 5, 5, 5], [0, 1, 1, 1], [0, 5, 5, 5], [1, 5, 5, 5], [0, 1, 1, 1]])
    u, inv, first = me.unique_first(c)
    assert u.tolist() == [[0, 5, 5, 5], [0, 1, 1, 1], [1, 5, 5, 5]] and inv.tolist() == [0, 1, 0, 2, 1] and first.tolist() == [0, 1, 3]
    f = torch.arange(10, dtype=torch.float64).view(5, 2)
    x = me.from_points(torch.from_numpy(c).double() + 0.4, f)                          # floor; first point wins
    assert x.F.tolist() == [[0, 1], [2, 3], [6, 7]]
    xa = me.from_points(torch.from_numpy(c).double(), f, average=True)
    assert xa.F.tolist() == [[2, 3], [5, 6], [6, 7]]
    neg = me.from_points(torch.tensor([[0, -0.5, -1.0, 0.99]], dtype=torch.float64), torch.zeros((1, 1)))
    assert neg.C.tolist() == [[0, -1, -1, 0]]                                          # floor toward -inf (A1)
