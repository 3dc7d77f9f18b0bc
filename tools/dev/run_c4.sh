#!/bin/bash
# on the GPU box: germany50/640 at 65,536 envs (kernel only) and config C4 (524,288 envs per GPU) for the given libraries
for so in "$@"; do
  v=$(timeout 200 python bench.py --lib "$PWD/$so" --configs none --no-cpu-baseline --no-e2e --steps 8 --warmup 3 --topology germany50 --slots 640 --load 800 --chunk 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ps=d.get('parity_sample') or {}; print('%.4e  %.3f ms  parity mism=%s bm=%s' % (d['value'], d['ms_per_step'], ps.get('mismatches'), ps.get('bitmap_mismatches')))")
  echo "$so g50/65536 $v" | tee -a gpurun_out/variants.log
  v=$(timeout 400 python bench.py --lib "$PWD/$so" --configs C4 --no-cpu-baseline --no-e2e --steps 4 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['C4']; ps=c.get('parity_sample') or {}; print('C4 value %.4e e2e %s parity mism=%s bm=%s' % (c['value'], (c.get('e2e') or {}).get('value'), ps.get('mismatches'), ps.get('bitmap_mismatches')))")
  echo "$so $v" | tee -a gpurun_out/variants.log
done
