#!/bin/bash
# B200 box: long-line (1025..2048 cells) x / y sweep variants on a 2048 x 2048 x 128 grid.  usage: tools/gpu_long.sh tag
tag=${1:-l}
mkdir -p gpurun_out
out=gpurun_out/${tag}_long_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt lb=256
run 2048 2048 128 --scalar --full --opt lb=256 --opt tw=1
run 2048 2048 128 --scalar --full --opt tw=1
run 2048 2048 128 --scalar --full --opt kt=4
run 2048 2048 128 --scalar
run 2048 2048 128 --scalar --opt lb=256
run 2048 2048 128 --scalar --full --opt dbg=1
run 2048 2048 128 --scalar --full --opt dbg=1 --opt lb=256
