#!/bin/bash
# on the GPU box: a few ncu counters of the observation kernel for every prebuilt variant
mkdir -p gpurun_out
for so in build_variants/*.so; do
  n=$(basename $so .so)
  timeout 300 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_observation_links -s 3 -c 1 --csv --log-file gpurun_out/m_$n.csv python tools/dev/obs_bench.py 16384 $PWD/$so > /dev/null 2>&1
  echo $n; grep -v "^==" gpurun_out/m_$n.csv | cut -d, -f13- | tail -5
done
