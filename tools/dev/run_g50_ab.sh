#!/bin/bash
# on the GPU box: germany50/640 at 65,536 envs (kernel only, with the oracle parity sample) for every prebuilt variant
mkdir -p gpurun_out
for so in build_variants/*.so; do
  v=$(timeout 300 python bench.py --lib "$PWD/$so" --configs none --no-cpu-baseline --no-e2e --steps 8 --warmup 3 --topology germany50 --slots 640 --load 800 --chunk 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ps=d.get('parity_sample') or {}; print('%.4e  %.3f ms  parity mism=%s bm=%s' % (d['value'], d['ms_per_step'], ps.get('mismatches'), ps.get('bitmap_mismatches')))")
  echo "$so g50/65536 $v" | tee -a gpurun_out/variants.log
done
