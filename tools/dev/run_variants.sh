#!/bin/bash
# on the GPU box: kernel-only bench of every prebuilt variant on nobel-eu/320 and germany50/640, with the oracle parity sample
for so in build_variants/*.so; do
  for cfg in "" "--topology germany50 --slots 640 --load 800 --chunk 128"; do
    v=$(timeout 200 python bench.py --lib "$PWD/$so" --configs none --no-cpu-baseline --no-e2e --steps 8 --warmup 3 $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ps=d.get('parity_sample') or {}; print('%.4e  %.3f ms  parity mism=%s exc=%s bm=%s' % (d['value'], d['ms_per_step'], ps.get('mismatches'), ps.get('excused'), ps.get('bitmap_mismatches')))")
    echo "$so $cfg $v" | tee -a gpurun_out/variants.log
  done
done
