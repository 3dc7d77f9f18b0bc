#!/bin/bash
# on the GPU box: launch list (device time per kernel) of a short config-5 rollout
mkdir -p gpurun_out
timeout 300 python examples/ppo_rollout.py --steps 12 > gpurun_out/c5_plain.log 2>&1 || { tail -5 gpurun_out/c5_plain.log; exit 1; }
tail -1 gpurun_out/c5_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c5_launches.csv python examples/ppo_rollout.py --steps 12 > /dev/null 2>&1
wc -l gpurun_out/c5_launches.csv
