#!/bin/bash
# development aid: timing experiments on single conv shapes.  CG3D_TC_DEBUG bits: 1 = 16-byte weight copies,
# 2 = no feature loads, 4 = no gather copies, 8 = per-role cycle counters printed to stderr, 16 = no MMAs,
# 64 = no rule-map loads in the K loop, 128 = no shared-memory stash of the rule-map columns
for args in "--stride 1 --cin 64 --cout 64" "--stride 2 --cin 64 --cout 64" "--stride 4 --cin 128 --cout 128" "--stride 8 --cin 256 --cout 256" "--stride 16 --cin 512 --cout 512" "--stride 32 --cin 512 --cout 512"; do
  echo "== $args"; python tools/conv_bench.py $args --iters 20 2>&1 | tail -1
  echo "== $args nostash"; CG3D_TC_DEBUG=128 python tools/conv_bench.py $args --iters 20 2>&1 | tail -1
done
for args in "--stride 2 --cin 64 --cout 64" "--stride 4 --cin 128 --cout 128"; do
  for dbg in 8 29; do
  echo "== $args debug=$dbg"; CG3D_TC_DEBUG=$dbg python tools/conv_bench.py $args --iters 2 2>&1 | tail -2
  done
done
