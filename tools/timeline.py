"""GPU timeline of one forward from CUDA events around every C-ABI call (development aid; nsys is not installed):
busy time per stream, the union's idle gaps and what follows each gap.

    python tools/timeline.py [--workload scannet] [--gaps 40]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cagroup3d_b200 import sparse as S
from tools.stage_times import setup


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=40)
    ap.add_argument("--min_gap_us", type=float, default=30.0)
    a = ap.parse_args()
    S.set_conv_impl("tc")
    model, pts, n1, n2 = setup(8, 50000)
    for _ in range(2):
        model({"points": pts.clone(), "batch_size": 8, "cur_epoch": 10})
    torch.cuda.synchronize()
    S.Profile.active, S.Profile.streams = [], []
    ref = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    p = pts.clone()
    torch.cuda.synchronize()
    ref.record()
    model({"points": p, "batch_size": 8, "cur_epoch": 10})
    end.record()
    torch.cuda.synchronize()
    rec, streams = S.Profile.active, S.Profile.streams
    S.Profile.active = S.Profile.streams = None
    total = ref.elapsed_time(end)
    iv = sorted((ref.elapsed_time(e0), ref.elapsed_time(e1), name, st) for (name, _, _, e0, e1), st in zip(rec, streams))
    print(f"forward {total:.2f} ms, {len(iv)} calls on {len(set(streams))} streams (event overhead included)")
    ids = {s: i for i, s in enumerate(sorted(set(streams)))}
    for s, i in ids.items():
        print(f"  stream {i}: {sum(b - a for a, b, _, st in iv if st == s):8.2f} ms in {sum(1 for x in iv if x[3] == s)} calls")
    agg = {}
    for a0, b0, name, st in iv:
        c = agg.setdefault((ids[st], name), [0, 0.0])
        c[0] += 1
        c[1] += b0 - a0
    for (st, name), (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"    stream {st} {name:34s} x{cnt:3d} {t:8.3f} ms")
    # union of intervals -> idle gaps
    gaps, cur_end, busy, last = [], 0.0, 0.0, "start"
    for a0, b0, name, st in iv:
        if a0 > cur_end:
            gaps.append((a0 - cur_end, cur_end, last, name, ids[st]))
            busy += b0 - a0
            cur_end = b0
        elif b0 > cur_end:
            busy += b0 - cur_end
            cur_end = b0
        if b0 >= cur_end:
            last = name
    idle = total - busy
    print(f"union busy {busy:.2f} ms, idle {idle:.2f} ms ({100 * idle / total:.1f} %)")
    big = sorted((g for g in gaps if g[0] * 1e3 >= a.min_gap_us), reverse=True)[:a.gaps]
    print(f"{len([g for g in gaps if g[0] * 1e3 >= a.min_gap_us])} gaps >= {a.min_gap_us:.0f} us, sum {sum(g[0] for g in gaps if g[0] * 1e3 >= a.min_gap_us):.2f} ms; largest:")
    for g, t, prev, nxt, st in sorted(big, key=lambda x: x[1]):
        print(f"   t={t:7.2f} ms  idle {g * 1e3:7.0f} us   after {prev:32s} before {nxt} (stream {st})")


main()
