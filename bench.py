"""Benchmark of the CAGroup3D inference hot path (BASELINE.json: scenes/s at ~50k voxels/scene).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--conv simt|tc]
                    [--workload scannet|sunrgbd|sweep|train|train_sunrgbd] [--voxel-size 0.02]

--workload (default scannet = BASELINE.json configs[1], the configuration the metric is quoted on):
    sunrgbd   configs[2]: SUN RGB-D-shaped inference, batch 16 per GPU, 10 classes, 100 000 points / scene, yaw boxes and
              rotated-BEV NMS, voxel 0.02 m (tools/cfgs/sunrgbd_models/CAGroup3D.yaml:9; --voxel-size 0.01 = BASELINE's figure)
    sweep     configs[4]: the scannet workload at 10k / 20k / 50k / 100k / 200k active voxels per scene (batch 8 per GPU =
              64 scenes on 8 GPUs); the line's value is the 50k point, config.sweep holds every density
    train     configs[3]: ScanNet-shaped TRAINING step (forward, both stages' losses, backward, bucketed gradient
              all-reduce, AdamW), 4 scenes per GPU, fp32
    train_sunrgbd   the same step for the SUN RGB-D model (WITH_YAW branches of both stages)

One "step" = one forward of the whole detector (voxelise -> BiResNet -> class-aware grouping head ->
RoI-Conv pooling -> NMS) over one batch of 8 synthetic ScanNet-shaped scenes (~50k active voxels each;
BASELINE.json configs[1]).  N > 1 runs one process per GPU under torchrun, every rank on its own
8 scenes (weak scaling, no data-path collective: scenes are independent); the timed region is
bracketed by barrier + synchronize and the max over ranks is reported.

Printed JSON line (rank 0): `value` = scenes/s with inputs resident in HBM, `e2e` = the same through
the pcdet-style `model(batch_dict)` call with pinned-host inputs and host outputs (H2D + D2H inside the
timed region), `roofline` = the sparse-conv launches of the backbone against the measured HBM peak,
`cpu_baseline` = the CPU oracle (restatement of the reference's MinkowskiEngine CPU algorithm) on one
scene of the same workload.  `--impl reference` times that CPU restatement alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "scenes_per_sec_at_50k_voxels_per_scene"
WORKLOAD = "ScanNetV2-shaped CAGroup3D inference, batch 8 per GPU, voxel 0.02 m, ~50k active voxels/scene, 18 classes"
P_SEL, P_BOX = 1.0 / 18, 0.002
# name -> (description, metric, n_classes, with_yaw, batch per GPU, voxel target of the generator, points per scene, seed config)
WORKLOADS = {
    "scannet": (WORKLOAD, METRIC, 18, False, 8, 50000, None, 2),
    "sunrgbd": ("SUN RGB-D-shaped CAGroup3D inference, batch 16 per GPU, 10 classes, 100000 points/scene, yaw boxes + rotated-BEV NMS",
                "scenes_per_sec_sunrgbd_100k_points_per_scene", 10, True, 16, 50000, 100000, 3),
    "sweep": ("density sweep of the ScanNetV2-shaped inference workload, batch 8 per GPU, 10k -> 200k active voxels/scene",
              METRIC, 18, False, 8, 50000, None, 5),
    "train": ("ScanNetV2-shaped CAGroup3D TRAINING step (forward + both stages' losses + backward + gradient all-reduce + AdamW), "
              "4 scenes per GPU, voxel 0.02 m, ~50k active voxels/scene, 18 classes, fp32 master weights",
              "train_scenes_per_sec_at_50k_voxels_per_scene", 18, False, 4, 50000, None, 4),
    "train_sunrgbd": ("SUN RGB-D-shaped CAGroup3D TRAINING step (WITH_YAW: 3 votes per seed, yaw code + rotated IoU loss, RoI stage with "
                      "(cos, sin) heading code and IoU loss), 4 scenes per GPU, voxel 0.02 m, ~50k active voxels/scene, 10 classes, fp32",
                      "train_scenes_per_sec_sunrgbd_at_50k_voxels_per_scene", 10, True, 4, 50000, None, 6),
}
SWEEP_VOXELS = (10000, 20000, 50000, 100000, 200000)


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.enabled = index, [], False, False

    def run(self):
        while not self.stop_flag:
            if self.enabled:
                try:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
                except Exception:
                    pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def cpu_oracle_setup(voxels):
    from cagroup3d_b200 import model_init, synthetic
    from oracle import cagroup3d_oracle as O
    torch.set_num_threads(os.cpu_count())
    batch = synthetic.make_batch(1, target_voxels=voxels, config=2)
    model = model_init.seeded_model(18, False, seed=0)
    pts = torch.from_numpy(batch["points"])
    orc = O.Oracle(model.state_dict(), O.default_cfg(18, False))
    bb = orc.forward(pts, 1, stages="backbone")
    model_init.calibrate_semantic_bias(model, bb["bb_feats"], P_SEL)
    orc = O.Oracle(model.state_dict(), O.default_cfg(18, False))
    return orc, pts


def time_cpu_oracle(voxels, steps, warmup):
    """scenes/s of the CPU restatement: one scene per step, all host threads."""
    orc, pts = cpu_oracle_setup(voxels)
    for _ in range(warmup):
        orc.forward(pts, 1, cur_epoch=10)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward(pts, 1, cur_epoch=10)
    dt = (time.perf_counter() - t0) / steps
    return 1.0 / dt, dt


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, dt = time_cpu_oracle(args.voxels, steps, warmup)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "scenes/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 scene per step (CPU is timed on a bounded sample)"},
        "cpu_baseline": {"value": v, "unit": "scenes/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} x 1 scene of the workload (~{args.voxels} voxels), full forward, "
                                   f"torch CPU fp32 gather->sgemm->scatter per kernel offset (ME CPU algorithm); "
                                   "MinkowskiEngine itself is not installable offline"},
        "e2e": {"value": v, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def conv_bytes(meta, P):
    """4(N_in C_in + N_out C_out) + 4 K C_in C_out + 8 P (+ residual read)."""
    b = 4 * (meta["n_in"] * meta["Cin"] + meta["n_out"] * meta["Cout"]) + meta["w_bytes"]
    if meta["nbr"] is not None:
        b += 8 * P
    if meta["residual"]:
        b += 4 * meta["n_out"] * meta["Cout"]
    return b


def setup_inference(wl, voxels, voxel_size, rank, conv):
    """model + synthetic batch of one rank, head-occupancy knobs calibrated (untimed)."""
    from cagroup3d_b200 import model_init, sparse as S, synthetic
    from cagroup3d_b200.detector import voxelize
    _, _, ncls, yaw, B, _, n_points, cfg = WORKLOADS[wl]
    S.set_conv_impl(conv)
    data = synthetic.make_batch(B, target_voxels=voxels, config=cfg, n_classes=ncls, sunrgbd=yaw, first_scene=rank * B,
                                n_points=n_points)
    host_pts = torch.from_numpy(data["points"]).pin_memory()
    dev_pts = host_pts.cuda()
    model = model_init.seeded_model(ncls, yaw, seed=0)
    if voxel_size != 0.02:
        model.voxel_size = voxel_size
        model.dense_head.voxel_size = voxel_size
        if model.roi_head is not None and hasattr(model.roi_head, "voxel_size"):
            model.roi_head.voxel_size = voxel_size
    model = model.cuda()
    # declared head-occupancy knobs (model_init.py): not timed
    p = dev_pts.clone()
    p[:, -3:] /= 255.
    x = voxelize(p, voxel_size)
    out = model.backbone_3d.run(x)
    n_vox, n_vox2 = x.cmap.n, out.cmap.n
    model_init.calibrate_semantic_bias(model, out.F, 1.0 / ncls)
    model.dense_head.semantic_threshold = 0.05
    cm = model.dense_head.class_maps(out, B)
    model_init.calibrate_cls_bias(model, cm["pred"], P_BOX)
    return model, host_pts, dev_pts, B, n_vox, n_vox2


class Timer:
    """W warm-up steps, then K steps between barrier + synchronize on both sides, CUDA events, max over ranks."""

    def __init__(self, dist, sampler):
        self.dist, self.sampler = dist, sampler

    def __call__(self, fn, steps, warmup, profile_convs=False):
        from cagroup3d_b200 import sparse as S
        dist = self.dist
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        self.sampler.enabled = True
        S.LaunchCounter.n = 0
        if profile_convs:
            S.Profile.active = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        self.sampler.enabled = False
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rec, S.Profile.active = S.Profile.active, None
        return ms.item(), S.LaunchCounter.n, rec


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def bench_inference(args, wl, voxels, rank, world, timed, conv, with_e2e=True):
    """value / roofline / e2e of one inference workload at one density -> dict of the JSON line's pieces."""
    from cagroup3d_b200 import sparse as S
    from cagroup3d_b200 import dist as D
    model, host_pts, dev_pts, B, n_vox, n_vox2 = setup_inference(wl, voxels, args.voxel_size, rank, conv)

    def step_resident():
        return model({"points": dev_pts.clone(), "batch_size": B, "cur_epoch": 10})

    def step_e2e():
        """the user-facing call: pinned host points -> device -> model(batch_dict) -> (N > 1: gather of all ranks'
        detections, the one collective of the path) -> host."""
        pts = host_pts.cuda(non_blocking=True)
        pred, _ = model({"points": pts, "batch_size": B, "cur_epoch": 10})
        if world > 1:
            return D.gather_detections(pred, world * B, to_host=True)
        return D.detections_to_host(pred)

    # ---- instrumented pass: per-launch algorithmic bytes of the sparse-conv kernel (untimed) ----
    S.Profile.active = []
    S.Profile.stage = "all"
    pred, _ = step_resident()
    torch.cuda.synchronize()
    rec, S.Profile.active = S.Profile.active, None
    convs = [(name, meta) for name, _, meta, _, _ in rec if meta is not None and name.startswith("cg3d_spconv")]
    conv_info = []

    def active_tile_taps(meta):
        """(128-row tile, tap) pairs the tensor-core kernel actually runs: a tap is skipped when no row of the tile has
        a neighbour for it.  Counted from the rule map the launch got (positional = tile order), untimed."""
        n_out, nbr = meta["n_out"], meta["nbr"]
        tiles = (n_out + 127) // 128
        if nbr is None:
            return tiles
        pad = tiles * 128 - n_out
        m = nbr[:, :n_out] >= 0
        if pad:
            m = torch.cat([m, torch.zeros((m.shape[0], pad), dtype=torch.bool, device=m.device)], 1)
        return int(m.view(m.shape[0], tiles, 128).any(-1).sum().item())

    n_backbone_convs = 56
    for name, meta in convs:
        P = S.count_rules(meta["nbr"]) if meta["nbr"] is not None else meta["n_out"]
        tt = active_tile_taps(meta) if (name == "cg3d_spconv_tc" and len(conv_info) < n_backbone_convs) else None
        conv_info.append(dict(kernel=name, K=meta["K"], Cin=meta["Cin"], Cout=meta["Cout"], n_in=meta["n_in"],
                              n_out=meta["n_out"], P=P, bytes=conv_bytes(meta, P), flops=2.0 * P * meta["Cin"] * meta["Cout"],
                              tile_taps=tt,
                              mma_flops=(3 * 2.0 * tt * 128 * meta["Cin"] * meta["Cout"]) if tt is not None else None))
    del rec, convs
    n_det = sum(len(d["pred_boxes"]) for d in pred)
    d2h_bytes = sum(d["pred_boxes"].numel() * 4 + d["pred_scores"].numel() * 4 + d["pred_labels"].numel() * 8 for d in pred)

    # value: inputs resident in HBM
    ms_total, launches, _ = timed(step_resident, args.steps, args.warmup)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    # roofline pass: the same K steps with CUDA events around every sparse-conv launch, everything issued
    # on ONE stream so that a launch is timed alone (in the value pass the backbone branches and the rule-map builds overlap)
    from cagroup3d_b200 import backbone as BB
    two, coord = BB._TWO_STREAMS["on"], BB._COORD_STREAM["on"]
    BB._TWO_STREAMS["on"] = BB._COORD_STREAM["on"] = False
    S.Profile.conv_only = True
    ms_serial, _, rec = timed(step_resident, args.steps, 1, profile_convs=True)
    S.Profile.conv_only = False
    BB._TWO_STREAMS["on"], BB._COORD_STREAM["on"] = two, coord
    # the bf16 split pass of a conv's input (cg3d_split_bf16) is charged to the conv launch that follows it
    per_call = [0.0] * len(conv_info)
    split_ms, pending, i = 0.0, 0.0, 0
    for r in rec:
        t = r[3].elapsed_time(r[4])
        if r[0].startswith("cg3d_split"):
            pending += t
            split_ms += t / args.steps
        else:
            per_call[i % len(conv_info)] += (t + pending) / args.steps
            pending = 0.0
            i += 1
    assert i == len(conv_info) * args.steps, (i, len(conv_info))
    for ci, t in zip(conv_info, per_call):
        ci["ms"] = t
        ci["GBps"] = ci["bytes"] / t / 1e6 if t > 0 else None
        ci["TFLOPs"] = ci["flops"] / t / 1e9 if t > 0 else None
    peaks = load_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    bb = conv_info[:n_backbone_convs]
    bb_bytes, bb_ms, bb_flops = sum(c["bytes"] for c in bb), sum(c["ms"] for c in bb), sum(c["flops"] for c in bb)
    all_ms, all_bytes = sum(c["ms"] for c in conv_info), sum(c["bytes"] for c in conv_info)
    # the 64-channel stride-1/2 layers are the HBM-bound part of the backbone (SURVEY 8d "sanity"); reported apart
    hb = [c for c in bb if c["Cin"] <= 64 and c["Cout"] <= 128]
    hbm_layers = {"launches": len(hb), "ms": sum(c["ms"] for c in hb),
                  "achieved": sum(c["bytes"] for c in hb) / max(sum(c["ms"] for c in hb), 1e-9) / 1e6, "unit": "GB/s"}
    hbm_layers["frac"] = hbm_layers["achieved"] / peak
    # every sparse-conv launch of the step (backbone + head incl. the 9^3 class conv + RoI stage), so that the headline
    # fraction cannot hide the most expensive launch
    rest = conv_info[n_backbone_convs:]
    worst = max(conv_info, key=lambda c: c["ms"])
    all_convs = {"launches": len(conv_info), "ms": all_ms, "achieved": all_bytes / max(all_ms, 1e-9) / 1e6, "unit": "GB/s",
                 "frac": all_bytes / max(all_ms, 1e-9) / 1e6 / peak,
                 "head_roi": {"launches": len(rest), "ms": sum(c["ms"] for c in rest),
                              "achieved": sum(c["bytes"] for c in rest) / max(sum(c["ms"] for c in rest), 1e-9) / 1e6},
                 "slowest_launch": {k: worst[k] for k in ("kernel", "K", "Cin", "Cout", "n_out", "P", "ms", "GBps", "TFLOPs")}}
    achieved = bb_bytes / bb_ms / 1e6
    # DRAM traffic per launch of the same 56 launches: from the committed ncu --set full capture (it cannot be measured
    # live: a number taken under a profiler is never a bench value, and the bench never runs under one)
    traffic, traffic_src = None, None
    if wl == "scannet" and voxels == 50000:
        for f in ("r2_spconv_traffic.json", "r1_spconv_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", f)))
                traffic, traffic_src = tj["traffic_bytes_per_launch_avg"], f"profiles/{f} ({tj['launches']} launches, {tj['metric']})"
                break
            except Exception:
                pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "kernel": f"cg3d_spconv_{conv} (56 backbone launches per step, algorithmic bytes = SURVEY 8d formula)",
                "launch_bytes_avg": bb_bytes / len(bb), "launch_ms_avg": bb_ms / len(bb),
                "backbone_ms": bb_ms, "backbone_tflops": bb_flops / bb_ms / 1e9,
                "tensor_peak_tflops": peaks.get("bf16_tflops_sustained"),
                "spconv_share_of_step": all_ms / (ms_serial / args.steps), "split_pass_ms_per_step": split_ms,
                "timed_in": "separate pass of the same steps, single stream, CUDA events around each launch "
                            f"({ms_serial / args.steps:.2f} ms/step)",
                "hbm_bound_layers": hbm_layers, "all_spconv_launches": all_convs}
    # the same launches read against the TENSOR roof: bf16 MMA work the kernel executes (3 products of the bf16x3 split x
    # 128-row tiles x the taps a tile does not skip) / measured sustained cuBLAS bf16 rate
    tcl = [c for c in bb if c.get("mma_flops")]
    if tcl and peaks.get("bf16_tflops_sustained"):
        t_ms = sum(c["ms"] for c in tcl)
        ach = sum(c["mma_flops"] for c in tcl) / t_ms / 1e9
        big = [c for c in tcl if c["Cin"] >= 128 and c["K"] > 1]
        roofline["tensor"] = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": ach / peaks["bf16_tflops_sustained"], "launches": len(tcl),
                              "what": "executed bf16 MMA FLOPs (bf16x3: 3 products; zero rows of a gathered tile included, skipped taps "
                                      "excluded) of the backbone tensor-core launches / sum of their durations",
                              "k27_layers_cin_ge_128": {"launches": len(big),
                                                        "achieved": sum(c["mma_flops"] for c in big) / max(sum(c["ms"] for c in big), 1e-9) / 1e9}}
        roofline["tensor"]["k27_layers_cin_ge_128"]["frac"] = roofline["tensor"]["k27_layers_cin_ge_128"]["achieved"] / peaks["bf16_tflops_sustained"]

    e2e = None
    if with_e2e:
        # e2e: pinned host inputs -> device -> forward -> host outputs
        ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": int(host_pts.numel() * 4), "d2h_bytes_per_step": int(d2h_bytes)}
    info = {"batch_per_gpu": B, "voxels_per_scene": n_vox // B, "stride2_voxels_per_scene": n_vox2 // B,
            "points_per_batch": int(host_pts.shape[0]), "detections_per_batch": n_det,
            "backbone_streams": 2 if two else 1, "coordinate_stream": bool(coord)}
    del model
    torch.cuda.empty_cache()
    return dict(value=value, ms_step=ms_step, launches=launches, roofline=roofline, e2e=e2e, info=info, conv_info=conv_info)


def bench_train(args, rank, world, timed, conv, dist):
    """BASELINE configs[3]: one training step = train_step.training_step (tools/train_utils/train_utils.py:48-72)."""
    from cagroup3d_b200 import backbone_train as BT, dist as D, model_init, sparse as S, synthetic, train_step as TS
    from cagroup3d_b200.detector import voxelize
    _, _, ncls, yaw, B, voxels, _, cfg = WORKLOADS[args.workload]
    B = args.batch or B
    voxels = args.voxels
    S.set_conv_impl(conv)
    dev = "cuda"
    scenes = [synthetic.make_scene(1000 * cfg + rank * B + i, voxels, n_classes=ncls, return_masks=True, sunrgbd=yaw) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    model = model_init.seeded_model(ncls, yaw, seed=0).to(dev).train()
    host_pts = torch.from_numpy(batch["points"]).pin_memory()
    pts = host_pts.to(dev)
    p = pts.clone()
    p[:, -3:] /= 255.
    with torch.no_grad():
        out = BT.run_train(model.backbone_3d, voxelize(p, 0.02))
    n_vox2 = out.cmap.n
    model_init.calibrate_semantic_bias(model, out.F.detach(), P_SEL)
    with torch.no_grad():
        model.dense_head.cls_conv.bias.fill_(-2.0)            # stage-1 detections for the RoI stage (random init has none)
    del out, p
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3)
    red = D.GradientAllReducer(params)
    gt = torch.from_numpy(batch["gt_boxes"]).float()
    # SUN RGB-D items carry no per-point masks (its vote targets come from the boxes)
    masks = {} if yaw else {"semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
    last = {}

    def step_resident():
        bd = {"points": pts.clone(), "batch_size": B, "cur_epoch": 10, "gt_boxes": gt.to(dev), **masks}
        last["tb"] = TS.training_step(model, bd, opt, red, grad_norm_clip=10.0)

    def step_e2e():
        bd = {"points": host_pts.to(dev, non_blocking=True), "batch_size": B, "cur_epoch": 10, "gt_boxes": gt.to(dev, non_blocking=True),
              **masks}
        last["tb"] = TS.training_step(model, bd, opt, red, grad_norm_clip=10.0)      # tb_dict: python floats = D2H of the losses

    ms_total, launches, _ = timed(step_resident, args.steps, args.warmup)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, 1)
    # per-kernel table of one step (CUDA events around every C-ABI call) + the collective's share
    S.Profile.active = []
    S.Profile.stage = "train"
    t_red = {"ms": 0.0}
    orig_reduce = red.reduce

    def timed_reduce():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = orig_reduce()
        b.record()
        t_red["ev"] = (a, b)
        return n
    red.reduce = timed_reduce
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step_resident()
    e1.record()
    torch.cuda.synchronize()
    red.reduce = orig_reduce
    rec, S.Profile.active = S.Profile.active, None
    per = {}
    for name, _, _, a0, a1 in rec:
        c = per.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += a0.elapsed_time(a1)
    kernel_ms = sum(v[1] for v in per.values())
    top = sorted(per.items(), key=lambda kv: -kv[1][1])[:14]
    ar_ms = t_red["ev"][0].elapsed_time(t_red["ev"][1]) if "ev" in t_red else 0.0
    # roofline of the dominant kernel of the step (weight gradient): algorithmic bytes = X rows + dY rows read once per tap
    # that has pairs ... reported as its share; the forward / dX convs reuse the inference kernel (see the scannet line)
    peaks = load_peaks()
    dom = top[0]
    roofline = {"bound": "tensor" if dom[0].startswith("cg3d_spconv") else "hbm", "kernel": dom[0], "calls_per_step": dom[1][0],
                "ms_per_step": dom[1][1], "share_of_kernel_time": dom[1][1] / max(kernel_ms, 1e-9),
                "achieved": None, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": None, "traffic": None,
                "note": "training step: per-kernel device times below; the conv forward / dX launches are the inference kernel "
                        "(roofline in the scannet workload's line)"}
    try:
        # algorithmic bytes (SURVEY 8d formula) of the step's conv forward + dX launches over their own device time
        cb, cms, cn = 0.0, 0.0, 0
        for name, _, meta, a0, a1 in rec:
            if meta is None or not name.startswith("cg3d_spconv"):
                continue
            P = int((meta["nbr"] >= 0).sum()) if meta["nbr"] is not None else 0
            cb += conv_bytes(meta, P)
            cms += a0.elapsed_time(a1)
            cn += 1
        if cn and cms > 0 and roofline["peak"]:
            roofline.update(bound="hbm", kernel="cg3d_spconv_tc / cg3d_spconv_simt (conv forward + dX launches of one training step)",
                            calls_per_step=cn, ms_per_step=cms, share_of_kernel_time=cms / max(kernel_ms, 1e-9),
                            achieved=cb / (cms * 1e-3) / 1e9, frac=cb / (cms * 1e-3) / 1e9 / roofline["peak"],
                            launch_bytes_avg=cb / cn,
                            note="algorithmic bytes = SURVEY 8d formula per launch (dX = the same conv over the transposed rule map); "
                                 "timed inside the instrumented step (events around every C-ABI call); these launches are "
                                 "tensor-bound like the inference ones (scannet line: roofline.tensor), traffic not captured")
    except Exception as e:                                      # the table above stays valid
        roofline["note"] += f" [conv byte accounting failed: {type(e).__name__}: {e}]"
    table = [{"kernel": k, "calls": v[0], "ms": v[1], "share": v[1] / max(kernel_ms, 1e-9)} for k, v in top]
    info = {"batch_per_gpu": B, "stride2_voxels_per_scene": n_vox2 // B, "points_per_batch": int(host_pts.shape[0]),
            "iters_per_sec": 1e3 / ms_step, "parameters_M": sum(q.numel() for q in params) / 1e6,
            "gradient_bucket_MB": red.nbytes / 1e6, "optimizer": "AdamW fp32, grad-norm clip 10",
            "allreduce_ms_per_step": ar_ms, "allreduce_share_of_step": ar_ms / max(e0.elapsed_time(e1), 1e-9),
            "kernel_ms_per_step": kernel_ms, "instrumented_step_ms": e0.elapsed_time(e1), "kernels": table,
            "losses": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in last.get("tb", {}).items()}}
    e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e / args.steps,
           "h2d_bytes_per_step": int(host_pts.numel() * 4 + gt.numel() * 4), "d2h_bytes_per_step": 8 * len(last.get("tb", {}))}
    return dict(value=value, ms_step=ms_step, launches=launches, roofline=roofline, e2e=e2e, info=info, conv_info=[])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="scannet", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="scenes per GPU (default: the workload's)")
    ap.add_argument("--voxels", type=int, default=50000)
    ap.add_argument("--voxel-size", type=float, default=0.02)
    ap.add_argument("--conv", default=os.environ.get("CG3D_CONV", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers-json", default=None, help="write the per-layer roofline table here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    wl = args.workload
    desc, metric = WORKLOADS[wl][0], WORKLOADS[wl][1]
    if args.batch:
        WORKLOADS[wl] = WORKLOADS[wl][:4] + (args.batch,) + WORKLOADS[wl][5:]

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from cagroup3d_b200 import _lib
    _lib.load()
    conv = args.conv
    if conv == "auto":
        conv = "tc" if hasattr(_lib.load(), "cg3d_spconv_tc") else "simt"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed = Timer(dist, sampler)

    extra = {}
    if wl.startswith("train"):
        r = bench_train(args, rank, world, timed, conv, dist)
    elif wl == "sweep":
        pts_ = []
        r = None
        for v in SWEEP_VOXELS:
            rv = bench_inference(args, wl, v, rank, world, timed, conv)
            rl = rv["roofline"]
            pts_.append({"target_voxels": v, "voxels_per_scene": rv["info"]["voxels_per_scene"], "scenes_per_sec": rv["value"],
                         "ms_per_step": rv["ms_step"], "e2e_scenes_per_sec": rv["e2e"]["value"],
                         "backbone_GBps": rl["achieved"], "backbone_frac": rl["frac"],
                         "hbm_bound_layers_GBps": rl["hbm_bound_layers"]["achieved"], "hbm_bound_layers_frac": rl["hbm_bound_layers"]["frac"],
                         "all_spconv_GBps": rl["all_spconv_launches"]["achieved"]})
            if v == 50000:
                r = rv
        extra["sweep"] = pts_
    else:
        r = bench_inference(args, wl, args.voxels, rank, world, timed, conv)
    sampler.stop_flag = True

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu_oracle(args.voxels, 1, 0)
        cpu = {"value": v, "unit": "scenes/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 ScanNet-shaped scene (~{args.voxels} voxels), full inference forward, {dt:.1f} s; torch CPU fp32 "
                         "gather->sgemm->scatter per kernel offset (the ME CPU algorithm; ME is not installable offline)"
                         + ("" if wl in ("scannet", "sweep") else "; the oracle has no timed arm for this workload, the scannet scene is the sample")}

    if rank == 0:
        cfg = {"workload": desc, **r["info"], "conv_impl": conv, "voxel_size_m": args.voxel_size, "p_sel": 1.0 / WORKLOADS[wl][2],
               "p_box": P_BOX, "weights": "seed-0 random init (no checkpoint offline)" + ("" if wl.startswith("train") else ", eval-mode BatchNorm"),
               "l2": "working set > L2: ~0.5 GB of weights (+ images) and the activations are re-streamed every step, no flush needed",
               **extra}
        line = {
            "metric": metric, "value": r["value"], "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if conv == "simt" else "f32 (bf16x3 split on tcgen05, fp32 accumulate)",
            "data": "synthetic", "config": cfg, "e2e": r["e2e"],
            "gpu_launches": r["launches"], "gpu_launches_note": "C-ABI calls in the timed region; each launches >= 1 kernel",
            "roofline": r["roofline"], "cpu_baseline": cpu, "clocks": sampler.summary(),
        }
        print(json.dumps(line))
        if args.layers_json:
            with open(args.layers_json, "w") as f:
                json.dump({"ms_per_step": r["ms_step"], "layers": r["conv_info"]}, f, indent=1)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
