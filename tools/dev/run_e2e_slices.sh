#!/bin/bash
# on the GPU box: the headline e2e leg for several env-slice counts
mkdir -p gpurun_out
for s in "$@"; do
  v=$(timeout 200 python bench.py --configs none --no-cpu-baseline --steps 3 --warmup 3 --e2e-slices $s 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4e e2e %.4e (%.3f s) device-requests %.4e' % (d['value'], d['e2e']['value'], d['e2e']['seconds'], d.get('e2e_device_requests',{}).get('value',0)))")
  echo "slices $s: $v" | tee -a gpurun_out/e2e_slices.log
done
