"""oracle/backward_oracle.py (explicit backward sums of the sparse ops) against torch autograd through the forward
oracle (oracle/me_cpu.py) -- the graph the reference's own training step differentiates in make_train_golden.py."""
import numpy as np
import pytest
import torch

from oracle import backward_oracle as Bk
from oracle import me_cpu as me


def _scene(seed, n=600, extent=14, batch=2):
    rng = np.random.default_rng(seed)
    c = np.concatenate([rng.integers(0, batch, (n, 1)), rng.integers(-extent, extent, (n, 3))], 1).astype(np.int64)
    return me.unique_first(c)[0]


def _tensor(seed, C, stride=1, n=600):
    rng = np.random.default_rng(seed)
    coords = _scene(seed, n) * np.array([1, stride, stride, stride])
    coords = np.unique(coords, axis=0)
    mgr = me.Manager()
    cm = me.CoordMap(coords, stride)
    mgr.by_stride[stride] = cm
    F = torch.from_numpy(rng.standard_normal((len(coords), C))).double().requires_grad_(True)
    return me.SparseTensor(F, cm, mgr), rng


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (2, 2), (5, 1)])
def test_conv_backward_equals_autograd(k, stride):
    x, rng = _tensor(10 * k + stride, 6)
    W = torch.from_numpy(rng.standard_normal((k ** 3, 6, 5))).double().requires_grad_(True)
    y = me.conv(x, W, k, stride)
    dY = torch.from_numpy(rng.standard_normal(tuple(y.F.shape))).double()
    (y.F * dY).sum().backward()
    rules = me.kernel_map(x.cmap, y.cmap.coords, k, x.cmap.stride)
    nbr = Bk.rules_to_table(rules, len(y.cmap))
    assert (nbr >= 0).sum() == sum(len(i) for i, _ in rules) > 0
    dX, dW = Bk.conv_backward(x.F.detach(), W.detach(), nbr, dY)
    assert torch.allclose(dX, x.F.grad, rtol=1e-12, atol=1e-12) and torch.allclose(dW, W.grad, rtol=1e-12, atol=1e-12)
    assert torch.allclose(Bk.conv_backward_by_transpose(x.F.detach(), W.detach(), nbr, dY), x.F.grad, rtol=1e-12, atol=1e-12)
    # positional table (the CUDA path's tile order): column j belongs to output row perm[j]
    perm = rng.permutation(len(y.cmap))
    dX2, dW2 = Bk.conv_backward(x.F.detach(), W.detach(), nbr[:, perm], dY, out_rows=perm)
    assert torch.allclose(dX2, dX, rtol=1e-12, atol=1e-12) and torch.allclose(dW2, dW, rtol=1e-12, atol=1e-12)
    assert np.array_equal(Bk.table_transpose(nbr[:, perm], len(x.cmap), perm), Bk.table_transpose(nbr, len(x.cmap)))


def test_transposed_convs_backward_equal_autograd():
    x, rng = _tensor(5, 4, stride=2)
    fine = _scene(6, 900)
    x.mgr.by_stride[1] = me.CoordMap(fine, 1)
    W = torch.from_numpy(rng.standard_normal((8, 4, 3))).double().requires_grad_(True)
    y = me.conv_transpose_k2s2(x, W)
    dY = torch.from_numpy(rng.standard_normal(tuple(y.F.shape))).double()
    (y.F * dY).sum().backward()
    # the transposed conv's rule map: fine row f <- coarse row floor(f / 2) * 2 at tap (f - parent)
    f = fine
    p = f.copy(); p[:, 1:] = np.floor_divide(f[:, 1:], 2) * 2
    d = f[:, 1:] - p[:, 1:]
    tap = d[:, 0] + 2 * (d[:, 1] + 2 * d[:, 2])
    rows = x.cmap.lookup(p)
    nbr = np.full((8, len(f)), -1, np.int32)
    nbr[tap, np.arange(len(f))] = rows
    assert (nbr >= 0).sum() > 0
    dX, dW = Bk.conv_backward(x.F.detach(), W.detach(), nbr, dY)
    assert torch.allclose(dX, x.F.grad, rtol=1e-12, atol=1e-12) and torch.allclose(dW, W.grad, rtol=1e-12, atol=1e-12)


def test_segment_mean_and_interp_backward_equal_autograd():
    rng = np.random.default_rng(3)
    inv = rng.integers(0, 40, 300)
    inv[:40] = np.arange(40)
    F = torch.from_numpy(rng.standard_normal((300, 5))).double().requires_grad_(True)
    out = torch.zeros((40, 5), dtype=torch.float64).index_add_(0, torch.from_numpy(inv), F) / \
        torch.bincount(torch.from_numpy(inv), minlength=40).double()[:, None]
    dOut = torch.from_numpy(rng.standard_normal((40, 5))).double()
    (out * dOut).sum().backward()
    assert torch.allclose(Bk.segment_mean_backward(dOut, inv, 300), F.grad, rtol=1e-12, atol=1e-12)

    x, rng = _tensor(8, 4, stride=4)
    q = _scene(9, 800, extent=50)
    y = me.features_at(x, q)
    dY = torch.from_numpy(rng.standard_normal(tuple(y.shape))).double()
    (y * dY).sum().backward()
    rows, w = Bk.interp_corners(x.cmap, q)
    assert (rows >= 0).sum() > 0
    assert torch.allclose(Bk.interp_backward(dY, rows, w, len(x.cmap)), x.F.grad, rtol=1e-12, atol=1e-12)


def test_batchnorm_train_backward_equals_autograd():
    rng = np.random.default_rng(4)
    F = torch.from_numpy(rng.standard_normal((200, 7)) * 3 + 1).double().requires_grad_(True)
    p = {"bn.weight": torch.from_numpy(rng.random(7) + 0.5).double().requires_grad_(True),
         "bn.bias": torch.from_numpy(rng.standard_normal(7)).double().requires_grad_(True)}
    out = me.batchnorm(F, p, "bn.", train=True)
    dOut = torch.from_numpy(rng.standard_normal((200, 7))).double()
    (out * dOut).sum().backward()
    dF, dg, db = Bk.batchnorm_train_backward(F.detach(), p["bn.weight"].detach(), dOut)
    assert torch.allclose(dF, F.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dg, p["bn.weight"].grad, rtol=1e-10) and torch.allclose(db, p["bn.bias"].grad, rtol=1e-10)
