#!/bin/bash
# Kernel experiments: run the kernel-only bench against prebuilt library variants (bench.py --lib).
#   tools/dev/bench_variants.sh [bench.py args, e.g. --topology germany50 --slots 640 --load 800] -- [libs...]
args=(); libs=()
while [ $# -gt 0 ]; do if [ "$1" == "--" ]; then shift; libs=("$@"); break; fi; args+=("$1"); shift; done
[ ${#libs[@]} -eq 0 ] && libs=(build_variants/*.so)
for so in "${libs[@]}"; do
  v=$(python bench.py --lib "$PWD/$so" --configs none --no-cpu-baseline --no-e2e --steps 8 --warmup 3 "${args[@]}" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e  %.3f ms' % (d['value'], d['ms_per_step']))")
  echo "$so ${args[*]} $v" | tee -a gpurun_out/variants.log
done
