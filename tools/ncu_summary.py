"""Summarise an .ncu-rep (read here, without a GPU) into the per-launch table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/spconv_tc.ncu-rep profiles/r1_ncu_spconv_tc_backbone [--traffic-json profiles/r1_spconv_traffic.json]

Writes <out>.csv (one row per captured launch) and <out>.md; --traffic-json stores the average DRAM traffic per launch
(dram__bytes_read.sum + dram__bytes_write.sum), which bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel", None),
    ("Grid Size", "grid", None),
    ("gpu__time_duration.sum", "time_us", "us"),
    ("dram__bytes_read.sum", "dram_read_MB", "MB"),
    ("dram__bytes_write.sum", "dram_write_MB", "MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", None),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB", "MB"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct", None),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_wavefront_pct", None),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts", None),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_lsu_wavefronts", None),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", None),
    ("launch__registers_per_thread", "regs", None),
    ("smsp__inst_executed.sum", "warp_insts", None),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    table = []
    for r in data:
        rec = {}
        for name, key, want in COLS:
            if name not in idx:
                continue
            v = r[idx[name]].replace(",", "")
            if want:
                v = float(v) * SCALE.get(units[idx[name]], 1.0)
            else:
                try:
                    v = float(v)
                except ValueError:
                    pass
            rec[key] = v
        if isinstance(rec.get("kernel"), str):
            rec["kernel"] = rec["kernel"].split("(")[0].replace("void <unnamed>::", "")[:48]
        table.append(rec)
    keys = [k for _, k, _ in COLS if any(k in t for t in table)]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for t in table:
            w.writerow({k: (f"{t[k]:.4g}" if isinstance(t.get(k), float) else t.get(k, "")) for k in keys})
    with open(out + ".md", "w") as f:
        f.write(f"ncu --set full --clock-control none, {len(table)} launches from `{rep.split('/')[-1]}` (values per launch)\n\n")
        f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for t in table:
            f.write("| " + " | ".join(f"{t[k]:.4g}" if isinstance(t.get(k), float) else str(t.get(k, "")) for k in keys) + " |\n")
    if tj:
        tr = [(t["dram_read_MB"] + t["dram_write_MB"]) * 1e6 for t in table]
        json.dump({"launches": len(tr), "traffic_bytes_per_launch_avg": sum(tr) / len(tr), "traffic_bytes_total": sum(tr),
                   "time_us_total_under_ncu": sum(t["time_us"] for t in table), "source": rep.split("/")[-1],
                   "metric": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full"}, open(tj, "w"), indent=1)
    print(f"{len(table)} launches -> {out}.csv / .md")


if __name__ == "__main__":
    main()
