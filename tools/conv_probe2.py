"""Times the 9^3 class conv / a 64-channel 3^3 layer alone (development aid): python tools/conv_probe2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cagroup3d_b200 import sparse as S
from tools.stage_times import setup

def main():
    S.set_conv_impl("tc")
    model, pts, n1, n2 = setup(8, 50000)
    rec = []
    orig = S._call
    def spy(name, *args, meta=None):
        if name == "cg3d_spconv_tc" and (args[2] is not None or os.environ.get("PROBE_ALL")):
            rec.append((name, args))
        return orig(name, *args, meta=meta)
    S._call = spy
    model({"points": pts.clone(), "batch_size": 8, "cur_epoch": 10})
    S._call = orig
    torch.cuda.synchronize()
    for name, args in rec:
        K = args[9]
        if K < int(os.environ.get("PROBE_MIN_K", "27")):
            continue
        n_out = args[6]
        line = f"{name} K={K} n_out={n_out} Cin={args[7]} Cout={args[8]} {'map' if args[2] is not None else 'rows'}"
        for dbg in os.environ.get("PROBE_DEBUGS", "0").split(","):
            os.environ["CG3D_TC_DEBUG"] = dbg                 # timing ablations of the SAME launch (results are garbage for dbg != 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            orig(name, *args)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                orig(name, *args)
            e1.record()
            torch.cuda.synchronize()
            line += f"  dbg{dbg}: {e0.elapsed_time(e1) / 3:.3f} ms"
        os.environ["CG3D_TC_DEBUG"] = "0"
        print(line)

main()
