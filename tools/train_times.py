"""Per-kernel device times of one whole (two-stage) TRAINING step (development aid for BASELINE config 4: ScanNet training, batch
4 per GPU; not the bench).  Forward + both losses + backward + optimizer through train_step.training_step on synthetic scenes with
per-point masks; prints ms per step and the C-ABI entry points ranked by device time.

    python tools/train_times.py [--batch 4] [--voxels 50000] [--steps 3] [--conv tc|simt]
"""
import argparse
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cagroup3d_b200 import backbone_train as BT, dist as D, model_init, sparse as S, synthetic, train_step as TS
from cagroup3d_b200.detector import voxelize


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--voxels", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--conv", default="tc")
    ap.add_argument("--p_sel", type=float, default=1 / 18)
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--host-profile", action="store_true", help="cProfile over the timed steps: where the HOST time of a step goes")
    a = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    S.set_conv_impl(a.conv)
    ncls, dev = 18, "cuda"
    t0 = time.time()
    scenes = [synthetic.make_scene(1000 * 4 + i, a.voxels, n_classes=ncls, return_masks=True) for i in range(a.batch)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    model = model_init.seeded_model(ncls, False, seed=0).to(dev).train()
    pts = torch.from_numpy(batch["points"]).to(dev)
    # head-occupancy knob of the bench: a declared fraction of the voxels passes the semantic threshold
    p = pts.clone()
    p[:, -3:] /= 255.
    with torch.no_grad():
        out = BT.run_train(model.backbone_3d, voxelize(p, 0.02))
    model_init.calibrate_semantic_bias(model, out.F.detach(), a.p_sel)
    with torch.no_grad():
        model.dense_head.cls_conv.bias.fill_(-2.0)            # stage-1 detections for the RoI stage
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3)
    red = D.GradientAllReducer(params)
    print(f"setup {time.time() - t0:.1f}s  points {tuple(pts.shape)}  parameters {sum(q.numel() for q in params) / 1e6:.1f} M "
          f"({red.nbytes / 1e6:.0f} MB of gradient buckets)")

    def step():
        bd = {"points": pts.clone(), "batch_size": a.batch, "cur_epoch": 10, "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float().to(dev),
              "semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
        return TS.training_step(model, bd, opt, red, grad_norm_clip=10.0)

    for _ in range(2):
        tb = step()
    torch.cuda.synchronize()
    print("tb_dict:", {k: round(v, 4) for k, v in tb.items()})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"training step {ms:.2f} ms -> {a.batch / ms * 1e3:.1f} scenes/s per GPU")
    if a.host_profile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(a.steps):
            step()
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(40)
        st.sort_stats("cumulative").print_stats(60)
    S.Profile.active = []
    S.Profile.stage = "train"
    step()
    torch.cuda.synchronize()
    per = collections.defaultdict(lambda: [0, 0.0])
    for name, _, _, a0, a1 in S.Profile.active:
        per[name][0] += 1
        per[name][1] += a0.elapsed_time(a1)
    S.Profile.active = None
    total = sum(v[1] for v in per.values())
    print(f"C-ABI launches in one step: {sum(v[0] for v in per.values())}, {total:.2f} ms of kernel time (events around each call)")
    for name, (cnt, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"  {name:36s} {cnt:5d} calls {t:9.3f} ms {100 * t / total:5.1f} %")


if __name__ == "__main__":
    main()
