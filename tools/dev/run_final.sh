#!/bin/bash
# on the GPU box: [tests] the whole -m gpu suite and smoke(); the full bench line (all configs); the reference arm
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
  tail -4 gpurun_out/final_tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
t0=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.jsonl 2> gpurun_out/bench_n1.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"
tail -c 600 gpurun_out/bench_n1.jsonl
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.jsonl 2>gpurun_out/bench_ref.err; echo "ref rc=$?"
