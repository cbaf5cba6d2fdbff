#!/bin/bash
# development aid: timing experiments on one conv shape (CG3D_TC_DEBUG bits: 1 = 16-byte weight copies, 2 = no feature
# loads, 4 = no gather copies, 8 = per-role cycle counters printed to stderr, 16 = no MMAs)
for args in "--stride 4 --cin 128 --cout 128" "--stride 2 --cin 64 --cout 64" "--stride 8 --cin 256 --cout 256" "--stride 1 --cin 64 --cout 64"; do
  echo "== $args order=none"; CG3D_TILE_ORDER=none python tools/conv_bench.py $args --iters 20 2>&1 | tail -1
  echo "== $args order=mask"; CG3D_TILE_ORDER=mask python tools/conv_bench.py $args --iters 20 2>&1 | tail -1
done
