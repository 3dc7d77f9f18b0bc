#!/bin/bash
# on the GPU box: observation kernel and highest-SNR policy for every prebuilt variant
mkdir -p gpurun_out
for so in build_variants/*.so; do
  timeout 200 python tools/dev/obs_bench.py 16384 $PWD/$so 2>&1 | tail -1 | tee -a gpurun_out/variants.log
  timeout 300 python tools/dev/hsnr_bench.py $PWD/$so 4096 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
