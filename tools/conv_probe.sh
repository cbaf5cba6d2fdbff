#!/bin/bash
# development aid: timing experiments on one conv shape (CG3D_TC_DEBUG bits: 1 = 16-byte weight copies, 2 = no feature
# loads, 8 = per-role cycle counters printed to stderr)
for args in "--stride 4 --cin 128 --cout 128" "--stride 2 --cin 64 --cout 64"; do
  echo "== $args order=none debug=8"; CG3D_TILE_ORDER=none CG3D_TC_DEBUG=8 python tools/conv_bench.py $args --iters 2 2>&1 | tail -3
  echo "== $args order=none debug=11"; CG3D_TILE_ORDER=none CG3D_TC_DEBUG=11 python tools/conv_bench.py $args --iters 2 2>&1 | tail -3
done
