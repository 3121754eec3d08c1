#!/bin/bash
# A/B of kernel builds on ONE box: tools/ab.sh spp variantA.so variantB.so ...   (files under rtiow_b200/lib/; alternated twice)
spp=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    cp rtiow_b200/lib/$v rtiow_b200/lib/librtiow_cuda.so
    echo -n "$v: "; python tools/profile_render.py $spp | tail -1
  done
done
