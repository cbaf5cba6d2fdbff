"""One forward of the bench workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off`
(launch list and the --set full capture of the sparse-conv kernel; see profiles/README.md).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cagroup3d_b200 import sparse as S
from stage_times import setup


def main():
    batch = int(os.environ.get("CG3D_BATCH", 8))
    voxels = int(os.environ.get("CG3D_VOXELS", 50000))
    S.set_conv_impl(os.environ.get("CG3D_CONV", "tc"))
    model, pts, n1, n2 = setup(batch, voxels)

    def step():
        return model({"points": pts.clone(), "batch_size": batch, "cur_epoch": 10})

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profiled one step: {n1} voxels, {n2} stride-2 voxels, batch {batch}")


if __name__ == "__main__":
    main()
