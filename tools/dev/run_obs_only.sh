#!/bin/bash
# on the GPU box: the observation kernel alone for every prebuilt variant
mkdir -p gpurun_out
for so in build_variants/*.so; do
  n=$(basename $so .so)
  timeout 200 python tools/dev/obs_bench.py 16384 $PWD/$so 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
