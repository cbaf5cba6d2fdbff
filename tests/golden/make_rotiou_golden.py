"""Golden vectors of the differentiable rotated 3-D IoU loss (SUN RGB-D training, WITH_YAW), produced by the REFERENCE's own
Python: pcdet/utils/iou3d_loss.py IoU3DLoss(with_yaw=True) -> iou_3d_loss -> pcdet/ops/rotated_iou/oriented_iou_loss.py
cal_iou_3d -> box_intersection_2d.py, run on the CPU.  The one piece of the reference that cannot run here, its CUDA op
sort_vertices (cuda_op/sort_vert_kernel.cu), is replaced by oracle/sort_vertices_oracle.py (pinned to that CUDA op on the GPU
box, tests/test_gpu_ref_ops.py).  Loss AND gradient (autograd through the reference's torch code) are stored.

    python tests/golden/make_rotiou_golden.py      ->  tests/golden/rotiou_loss.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import make_golden as MG          # puts /root/reference first on sys.path, installs the import stubs
from oracle import sort_vertices_oracle as SVO


def main():
    MG.install()
    sv = types.ModuleType("sort_vertices")
    sv.sort_vertices_forward = lambda v, m, nv: torch.from_numpy(
        SVO.sort_vertices(v.detach().numpy(), m.numpy(), nv.numpy())).int()
    sys.modules["sort_vertices"] = sv
    # box2corners_th builds its constants with torch.FloatTensor(...).to(device): fine on the CPU
    from pcdet.utils.iou3d_loss import IoU3DLoss
    g = torch.Generator().manual_seed(11)
    n = 160
    tgt = torch.cat([(torch.rand((n, 3), generator=g) - 0.5) * 4, torch.rand((n, 3), generator=g) * 1.5 + 0.2,
                     (torch.rand((n, 1), generator=g) - 0.5) * 6], 1)
    pred = tgt + torch.randn((n, 7), generator=g) * torch.tensor([0.15, 0.15, 0.15, 0.1, 0.1, 0.1, 0.3])
    pred[:, 3:6] = pred[:, 3:6].abs() + 0.05
    pred[:16, :3] += 5.0                                  # disjoint pairs: zero IoU, zero gradient
    pred[16:24] = tgt[16:24]                              # identical boxes (the num_valid == 8 corner case)
    pred[24:32, 3:6] = tgt[24:32, 3:6] * 0.5              # contained, same centre and yaw
    pred[24:32, [0, 1, 2, 6]] = tgt[24:32][:, [0, 1, 2, 6]]
    pred[32:40, 6] = tgt[32:40, 6] + np.pi / 2            # a quarter turn
    weight = torch.rand((n,), generator=g)
    weight[::7] = 0.0
    out = {"pred": pred.numpy(), "target": tgt.numpy(), "weight": weight.numpy(), "avg_factor": np.float32(37.5)}
    p = pred.clone().requires_grad_(True)
    loss = IoU3DLoss(with_yaw=True, loss_weight=1.0)(p, tgt, weight=weight, avg_factor=37.5)
    loss.backward()
    out["loss"], out["grad"] = loss.detach().numpy(), p.grad.numpy()
    from pcdet.ops.rotated_iou.oriented_iou_loss import cal_iou_3d
    out["iou"] = cal_iou_3d(pred[None], tgt[None])[0].numpy()
    np.savez_compressed(os.path.join(HERE, "rotiou_loss.npz"), **out)
    print("rotiou_loss.npz  loss", float(loss), " iou range", out["iou"].min(), out["iou"].max(), " |grad| max", np.abs(out["grad"]).max())


if __name__ == "__main__":
    main()
