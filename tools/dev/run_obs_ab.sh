#!/bin/bash
# on the GPU box: parity suite on the in-tree library, then the observation kernel and config C5 of every prebuilt variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_tests.log
tail -5 gpurun_out/ab_tests.log
for so in build_variants/*.so; do
  n=$(basename $so .so)
  timeout 200 python tools/dev/obs_bench.py 16384 $PWD/$so gpurun_out/obs_$n.npz 2>&1 | tail -1 | tee -a gpurun_out/variants.log
  v=$(timeout 400 python bench.py --lib "$PWD/$so" --configs C5 --no-cpu-baseline --no-e2e --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['C5']; print('headline %.4e  C5 value %.4e e2e %s parity %s' % (d['value'], c['value'], (c.get('e2e') or {}).get('value'), c.get('parity_sample')))")
  echo "$so $v" | tee -a gpurun_out/variants.log
done
