#!/bin/bash
# on the GPU box: parity suite on the in-tree library, then the highest-SNR policy of every prebuilt variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_tests.log
tail -15 gpurun_out/ab_tests.log
for so in build_variants/*.so; do
  for n in 4096 16384; do
    timeout 300 python tools/dev/hsnr_bench.py $PWD/$so $n 2>&1 | tail -1 | tee -a gpurun_out/variants.log
  done
  timeout 200 python tools/dev/obs_bench.py 16384 $PWD/$so 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
