#!/bin/bash
# on the GPU box: parity suite on the in-tree library, then the headline and germany50 step kernels of every prebuilt variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_tests.log
tail -15 gpurun_out/ab_tests.log
for so in build_variants/*.so; do
  for cfg in "" "--topology germany50 --slots 640 --load 800 --chunk 128"; do
    v=$(timeout 300 python bench.py --lib "$PWD/$so" --configs none --no-cpu-baseline --no-e2e --steps 8 --warmup 3 $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ps=d.get('parity_sample') or {}; print('%.4e  %.3f ms  parity mism=%s exc=%s bm=%s' % (d['value'], d['ms_per_step'], ps.get('mismatches'), ps.get('excused'), ps.get('bitmap_mismatches')))")
    echo "$so $cfg $v" | tee -a gpurun_out/variants.log
  done
  timeout 200 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum -k regex:k_step_policy -c 6 --csv --log-file gpurun_out/ncu_$(basename $so .so).csv python bench.py --lib $PWD/$so --configs none --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
  grep "inst_executed\|issue_active" gpurun_out/ncu_$(basename $so .so).csv | tail -2 | cut -d, -f13-
done
