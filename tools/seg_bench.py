import sys, torch
sys.path.insert(0, '/root/repo')
from cagroup3d_b200 import sparse as S
dev='cuda'
for n_seg, npts, C in ((150000, 175616, 128), (150000, 175616, 64), (2000, 175616, 128)):
    g = torch.Generator().manual_seed(0)
    tgt = torch.randint(0, n_seg, (npts,), generator=g, dtype=torch.int32).to(dev)
    G = torch.randn((npts, C), device=dev)
    keys = tgt.to(torch.int64); order = torch.arange(npts, dtype=torch.int32, device=dev)
    S.sort_pairs(keys, order, npts, end_bit=18)
    counts = torch.zeros((n_seg + 1,), dtype=torch.int32, device=dev)
    S._call("cg3d_histogram_i32", tgt, npts, n_seg, counts)
    seg_off, _ = S.exclusive_scan(counts)
    out = torch.empty((n_seg, C), device=dev)
    for _ in range(2): S._call("cg3d_segment_sum_sorted", G, order, seg_off, n_seg, C, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): S._call("cg3d_segment_sum_sorted", G, order, seg_off, n_seg, C, out)
    e1.record(); torch.cuda.synchronize()
    ref = torch.zeros((n_seg, C), device=dev).index_add_(0, tgt.long(), G)
    print(n_seg, npts, C, "%.3f ms" % (e0.elapsed_time(e1)/5), "max err", float((out-ref).abs().max()))
