"""Time ONE sparse-conv layer shape on realistic coordinates (development aid; used for ncu captures).

    python tools/conv_bench.py --stride 4 --cin 128 --cout 128 --k 3 [--iters 20] [--impl tc]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cagroup3d_b200 import sparse as S, synthetic
from cagroup3d_b200.detector import voxelize


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--voxels", type=int, default=50000)
    ap.add_argument("--stride", type=int, default=4)
    ap.add_argument("--cin", type=int, default=128)
    ap.add_argument("--cout", type=int, default=128)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impl", default="tc")
    a = ap.parse_args()
    data = synthetic.make_batch(a.batch, target_voxels=a.voxels)
    pts = torch.from_numpy(data["points"]).cuda()
    x = voxelize(pts, 0.02)
    cmap = x.cmap
    s = 1
    while s < a.stride:
        cmap = S.strided_map(cmap, x.mgr, 2)
        s *= 2
    n = cmap.n
    F = torch.randn((n, a.cin), device="cuda")
    W = torch.randn((a.k ** 3, a.cin, a.cout), device="cuda") / (a.cin * 8) ** 0.5
    nbr, order = S.neighbor_table(cmap, cmap, a.k, x.mgr, ordered=True)
    P = S.count_rules(nbr)
    res = torch.randn((n, a.cout), device="cuda")
    scale = torch.rand((a.cout,), device="cuda") + 0.5
    for _ in range(3):
        out = S.gemm_rows(F, nbr, W, n, a.k ** 3, scale=scale, shift=scale, residual=res, act="relu", impl=a.impl, out_rows=order)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = S.gemm_rows(F, nbr, W, n, a.k ** 3, scale=scale, shift=scale, residual=res, act="relu", impl=a.impl, out_rows=order)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    byts = 4 * n * (a.cin + 2 * a.cout) + 4 * W.numel() + 8 * P
    print(f"n={n} P={P} pairs/row={P / n:.1f}  {ms:.4f} ms  {byts / ms / 1e6:.1f} GB/s (algorithmic)  "
          f"{2.0 * P * a.cin * a.cout / ms / 1e9:.1f} TFLOP/s useful  "
          f"{2.0 * n * a.k ** 3 * a.cin * a.cout / ms / 1e9:.1f} TFLOP/s dense-equivalent")


if __name__ == "__main__":
    main()
