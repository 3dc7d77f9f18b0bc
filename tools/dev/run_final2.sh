#!/bin/bash
# on the GPU box: tests, smoke, captures, the full bench line
bash tools/dev/run_final.sh tests
bash tools/dev/run_captures.sh
