#!/bin/bash
# germany50 / 640 slots / load 800 (BASELINE config 4, one GPU's share scaled down) against prebuilt libraries
for so in ${@:-build_variants/*.so}; do
  v=$(QRMSA_LIB=$PWD/$so python bench.py --no-cpu-baseline --no-e2e --topology germany50 --slots 640 --load 800 --envs 65536 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e  %.3f ms' % (d['value'], d['ms_per_step']))")
  echo "g50 $so RING=${QRMSA_RING_SMEM:-1} $v" | tee -a gpurun_out/variants.log
done
