#!/bin/bash
# run the kernel-only bench against every prebuilt library variant in build_variants/ (kernel experiments)
for so in ${@:-build_variants/*.so}; do
  v=$(QRMSA_LIB=$PWD/$so python bench.py --no-cpu-baseline --no-e2e --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e  %.3f ms' % (d['value'], d['ms_per_step']))")
  echo "$so $v" | tee -a gpurun_out/variants.log
done
