#!/bin/bash
for arr in thread warp; do
for args in "--stride 2 --cin 64 --cout 64" "--stride 4 --cin 128 --cout 128"; do
for dbg in 8 24 29 61 125 44 108; do
  echo "== $args arrive=$arr debug=$dbg"; CG3D_TC_ARRIVE=$arr CG3D_TC_DEBUG=$dbg python tools/conv_bench.py $args --iters 2 2>&1 | tail -2
done; done; done
