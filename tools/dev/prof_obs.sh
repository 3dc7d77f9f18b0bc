#!/bin/bash
# on the GPU box: one ncu --set full capture of the observation kernel (in-tree library) after a plain run has exited 0
mkdir -p gpurun_out
timeout 200 python tools/dev/obs_bench.py 16384 > gpurun_out/obs_plain.log 2>&1 || { tail -5 gpurun_out/obs_plain.log; exit 1; }
tail -1 gpurun_out/obs_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_observation_links -s 3 -c 1 -o gpurun_out/prof_obs -f python tools/dev/obs_bench.py 16384 > gpurun_out/obs_ncu.log 2>&1
tail -2 gpurun_out/obs_ncu.log; ls -la gpurun_out/
