#!/bin/bash
tag=${1:-z}
mkdir -p gpurun_out
out=gpurun_out/${tag}_zshort_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out | cut -c1-170; }
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt zt=2
run 2048 2048 128 --scalar --full --opt zt=2 --opt zm=32
run 2048 2048 128 --scalar --full --opt zt=2 --opt zm=32 --opt lt=16
run 2048 2048 128 --scalar --full --opt zt=2 --opt zm=32 --opt lt=8
run 2048 2048 128 --scalar --full --opt zt=2 --opt lt=32
run 512 512 128 --opt zt=2 --opt zm=32
run 512 512 128
run 1024 1024 256 --opt zm=16
