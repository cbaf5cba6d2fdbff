"""Per-stage and per-kernel device times of one CAGroup3D forward (development aid, not the bench).

    python tools/stage_times.py [--batch 8] [--voxels 50000] [--conv simt|tc]
"""
import argparse
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cagroup3d_b200 import model_init, sparse as S, synthetic
from cagroup3d_b200.detector import voxelize


def setup(batch, voxels, ncls=18, yaw=False, p_sel=1 / 18, p_box=0.002, seed=0, config=2, first_scene=0):
    data = synthetic.make_batch(batch, target_voxels=voxels, n_classes=ncls, sunrgbd=yaw, config=config,
                                first_scene=first_scene)
    model = model_init.seeded_model(ncls, yaw, seed=seed).cuda()
    pts = torch.from_numpy(data["points"]).cuda()
    p = pts.clone()
    p[:, -3:] /= 255.
    x = voxelize(p, 0.02)
    out = model.backbone_3d.run(x)
    model_init.calibrate_semantic_bias(model, out.F, p_sel)
    model.dense_head.semantic_threshold = 0.05
    cm = model.dense_head.class_maps(out, batch)
    model_init.calibrate_cls_bias(model, cm["pred"], p_box)
    return model, pts, x.cmap.n, out.cmap.n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--voxels", type=int, default=50000)
    ap.add_argument("--conv", default="simt")
    ap.add_argument("--p_box", type=float, default=0.002)
    ap.add_argument("--top", type=int, default=60)
    a = ap.parse_args()
    S.set_conv_impl(a.conv)
    t0 = time.time()
    model, pts, n1, n2 = setup(a.batch, a.voxels, p_box=a.p_box)
    print(f"setup {time.time() - t0:.1f}s  points {tuple(pts.shape)}  voxels {n1}  stride-2 voxels {n2}")

    def step():
        return model({"points": pts.clone(), "batch_size": a.batch, "cur_epoch": 10})

    for _ in range(2):
        pd, _ = step()
    torch.cuda.synchronize()
    print("detections/sample:", [len(p["pred_boxes"]) for p in pd])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"forward {ms:.2f} ms/batch -> {a.batch / ms * 1e3:.1f} scenes/s")
    # per-stage / per-kernel breakdown
    stages = collections.OrderedDict()
    b = {"points": pts.clone(), "batch_size": a.batch, "cur_epoch": 10}
    S.Profile.active = []
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))
    mark("start")
    b["points"][:, -3:] /= 255.
    model.dense_head.semantic_threshold = 0.05
    S.Profile.stage = "voxelize"; x = voxelize(b["points"], 0.02); mark("voxelize")
    S.Profile.stage = "backbone"; out = model.backbone_3d.run(x); mark("backbone")
    S.Profile.stage = "head.class_maps"; cm = model.dense_head.class_maps(out, a.batch); mark("head.class_maps")
    S.Profile.stage = "head.proposals"; db, ds, dl, off, _ = model.dense_head.proposals(cm, a.batch); mark("head.proposals")
    S.Profile.stage = "roi"; fb, fs, fl, foff, inter = model.roi_head.run(out, db, ds, dl, off, a.batch); mark("roi")
    torch.cuda.synchronize()
    rec, S.Profile.active = S.Profile.active, None
    for (n0, ea), (n1_, eb) in zip(marks[:-1], marks[1:]):
        print(f"  stage {n1_:18s} {ea.elapsed_time(eb):9.3f} ms")
    print(f"  class-map voxels {cm['pred'].shape[0]}  stage-1 rois {off[-1]}  final {foff[-1]}  "
          f"unique grid voxels {inter['uniq'].shape[0]}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, stage, meta, ea, eb in rec:
        k = (stage, name)
        agg[k][0] += 1
        agg[k][1] += ea.elapsed_time(eb)
    for (stage, name), (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"  {stage:18s} {name:28s} x{cnt:4d} {t:9.3f} ms")
    print("  conv layers:")
    for name, stage, meta, ea, eb in rec:
        if meta:
            P = S.count_rules(meta["nbr"]) if meta["nbr"] is not None else meta["n_out"]
            byts = 4 * (meta["n_in"] * meta["Cin"] + meta["n_out"] * meta["Cout"]) + meta["w_bytes"] + \
                (8 * P if meta["nbr"] is not None else 0) + (4 * meta["n_out"] * meta["Cout"] if meta["residual"] else 0)
            t = ea.elapsed_time(eb)
            fl = 2.0 * P * meta["Cin"] * meta["Cout"]
            print(f"    {stage:16s} {name[5:]:12s} K={meta['K']:3d} {meta['Cin']:4d}->{meta['Cout']:4d} n_out={meta['n_out']:7d} P={P:8d} "
                  f"{t:8.3f} ms  {byts / t / 1e6:8.1f} GB/s  {fl / t / 1e9:8.2f} TFLOP/s")


if __name__ == "__main__":
    main()
