"""Benchmark of the CAGroup3D inference hot path (BASELINE.json: scenes/s at ~50k voxels/scene).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--conv simt|tc]

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--voxels", type=int, default=50000)
    ap.add_argument("--conv", default=os.environ.get("CG3D_CONV", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers-json", default=None, help="write the per-layer roofline table here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from cagroup3d_b200 import _lib, model_init, sparse as S, synthetic
    from cagroup3d_b200.detector import voxelize
    _lib.load()
    conv = args.conv
    if conv == "auto":
        conv = "tc" if hasattr(_lib.load(), "cg3d_spconv_tc") else "simt"
    S.set_conv_impl(conv)

    B = args.batch
    data = synthetic.make_batch(B, target_voxels=args.voxels, config=2, first_scene=rank * B)
    host_pts = torch.from_numpy(data["points"]).pin_memory()
    dev_pts = host_pts.cuda()
    model = model_init.seeded_model(18, False, seed=0).cuda()
    # declared head-occupancy knobs (model_init.py): not timed
    p = dev_pts.clone()
    p[:, -3:] /= 255.
    x = voxelize(p, 0.02)
    out = model.backbone_3d.run(x)
    n_vox, n_vox2 = x.cmap.n, out.cmap.n
    model_init.calibrate_semantic_bias(model, out.F, P_SEL)
    model.dense_head.semantic_threshold = 0.05
    cm = model.dense_head.class_maps(out, B)
    model_init.calibrate_cls_bias(model, cm["pred"], P_BOX)
    del p, x, out, cm

    def step_resident():
        return model({"points": dev_pts.clone(), "batch_size": B, "cur_epoch": 10})

    from cagroup3d_b200 import dist as D

    def step_e2e():
        """the user-facing call: pinned host points -> device -> model(batch_dict) -> (N > 1: gather of all ranks'
        detections, the one collective of the path) -> host."""
        pts = host_pts.cuda(non_blocking=True)
        pred, _ = model({"points": pts, "batch_size": B, "cur_epoch": 10})
        if world > 1:
            pred = D.gather_detections(pred, world * B)
        outs = [(d["pred_boxes"].cpu(), d["pred_scores"].cpu(), d["pred_labels"].cpu()) for d in pred]
        return outs

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

    for name, meta in convs:
        P = S.count_rules(meta["nbr"]) if meta["nbr"] is not None else meta["n_out"]
        tt = active_tile_taps(meta) if (name == "cg3d_spconv_tc" and len(conv_info) < 56) else None
        conv_info.append(dict(kernel=name, K=meta["K"], Cin=meta["Cin"], Cout=meta["Cout"], n_in=meta["n_in"],
                              n_out=meta["n_out"], P=P, bytes=conv_bytes(meta, P), flops=2.0 * P * meta["Cin"] * meta["Cout"],
                              tile_taps=tt,
                              mma_flops=(3 * 2.0 * tt * 128 * meta["Cin"] * meta["Cout"]) if tt is not None else None))
    n_backbone_convs = 56
    del rec, convs
    n_det = sum(len(d["pred_boxes"]) for d in pred)
    d2h_bytes = sum(d["pred_boxes"].numel() * 4 + d["pred_scores"].numel() * 4 + d["pred_labels"].numel() * 8 for d in pred)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed(fn, steps, warmup, profile_convs=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.enabled = True
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
        sampler.enabled = False
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rec, S.Profile.active = S.Profile.active, None
        return ms.item(), S.LaunchCounter.n, rec

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
    try:
        peaks_hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
    except Exception:
        peaks_hbm = 6650.0
    bb = conv_info[:n_backbone_convs]
    bb_bytes, bb_ms, bb_flops = sum(c["bytes"] for c in bb), sum(c["ms"] for c in bb), sum(c["flops"] for c in bb)
    all_ms = sum(c["ms"] for c in conv_info)
    # the 64-channel stride-1/2 layers are the HBM-bound part of the backbone (SURVEY 8d "sanity"); reported apart
    hb = [c for c in bb if c["Cin"] <= 64 and c["Cout"] <= 128]
    hbm_layers = {"launches": len(hb), "ms": sum(c["ms"] for c in hb),
                  "achieved": sum(c["bytes"] for c in hb) / max(sum(c["ms"] for c in hb), 1e-9) / 1e6, "unit": "GB/s"}
    hbm_layers["frac"] = hbm_layers["achieved"] / peaks_hbm
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = bb_bytes / bb_ms / 1e6
    # DRAM traffic per launch of the same 56 launches: from the committed ncu --set full capture (it cannot be measured
    # live: a number taken under a profiler is never a bench value, and the bench never runs under one)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_spconv_traffic.json")))
        traffic, traffic_src = tj["traffic_bytes_per_launch_avg"], f"profiles/r1_spconv_traffic.json ({tj['launches']} launches, {tj['metric']})"
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
                "hbm_bound_layers": hbm_layers}
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

    # e2e: pinned host inputs -> device -> forward -> host outputs
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    sampler.stop_flag = True

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu_oracle(args.voxels, 1, 0)
        cpu = {"value": v, "unit": "scenes/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 scene of the workload (~{args.voxels} voxels), full forward, {dt:.1f} s; torch CPU fp32 "
                         "gather->sgemm->scatter per kernel offset (the ME CPU algorithm; ME is not installable offline)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if conv == "simt" else "f32 (bf16x3 split on tcgen05, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "voxels_per_scene": n_vox // B,
                       "stride2_voxels_per_scene": n_vox2 // B, "points_per_batch": int(host_pts.shape[0]),
                       "conv_impl": conv, "backbone_streams": 2 if two else 1, "coordinate_stream": bool(coord), "p_sel": P_SEL, "p_box": P_BOX, "detections_per_batch": n_det,
                       "weights": "seed-0 random init (no checkpoint offline), eval-mode BatchNorm",
                       "l2": "working set > L2: 506 MB of weights + activations re-streamed every step, no flush needed"},
            "e2e": {"value": e2e_value, "unit": "scenes/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(host_pts.numel() * 4), "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": launches, "gpu_launches_note": "C-ABI calls in the timed region; each launches >= 1 kernel",
            "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
        }
        print(json.dumps(line))
        if args.layers_json:
            with open(args.layers_json, "w") as f:
                json.dump({"ms_per_step": ms_step, "layers": conv_info}, f, indent=1)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
