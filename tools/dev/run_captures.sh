#!/bin/bash
# on the GPU box: the round's ncu evidence.  Every capture follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --configs none"
NCU="ncu --set full --clock-control none --import-source on -f"
$B > gpurun_out/cap_plain_step.log 2>&1 || { echo "plain step run failed"; tail -3 gpurun_out/cap_plain_step.log; exit 1; }
timeout 600 $NCU -k regex:k_step_policy -s 2 -c 1 -o gpurun_out/prof_r2_step $B > gpurun_out/cap_step.log 2>&1; echo "step capture rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > /dev/null 2>&1; echo "launch list rc=$?"
G="$B --topology germany50 --slots 640 --load 800 --chunk 128"
$G > gpurun_out/cap_plain_g50.log 2>&1 || { echo "plain g50 run failed"; exit 1; }
timeout 600 $NCU -k regex:k_step_policy -s 2 -c 1 -o gpurun_out/prof_r2_g50 $G > gpurun_out/cap_g50.log 2>&1; echo "g50 capture rc=$?"
python tools/dev/obs_bench.py 16384 > gpurun_out/cap_plain_obs.log 2>&1 || { echo "plain obs run failed"; exit 1; }
timeout 600 $NCU -k regex:k_observation_links -s 3 -c 1 -o gpurun_out/prof_r2_obs python tools/dev/obs_bench.py 16384 > gpurun_out/cap_obs.log 2>&1; echo "obs capture rc=$?"
python tools/dev/hsnr_bench.py > gpurun_out/cap_plain_hsnr.log 2>&1 || { echo "plain hsnr run failed"; exit 1; }
timeout 900 $NCU -k regex:k_step_highest_snr_links -s 1 -c 1 -o gpurun_out/prof_r2_hsnr python tools/dev/hsnr_bench.py > gpurun_out/cap_hsnr.log 2>&1; echo "hsnr capture rc=$?"
tail -1 gpurun_out/cap_plain_obs.log gpurun_out/cap_plain_hsnr.log
ls -la gpurun_out/*.ncu-rep
