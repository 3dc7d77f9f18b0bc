#!/bin/bash
# on the GPU box: the -m gpu suite (optionally a -k expression)
mkdir -p gpurun_out
timeout ${2:-900} python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/tests_only.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_only.log
tail -25 gpurun_out/tests_only.log
