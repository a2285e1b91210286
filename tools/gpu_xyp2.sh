#!/bin/bash
tag=${1:-p}
mkdir -p gpurun_out
out=gpurun_out/${tag}_xyp_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt seq=0
run 2048 2048 128 --scalar --full --opt promo=0
run 2048 2048 128 --scalar --full --opt promo=128
run 2048 2048 128 --scalar --full --opt seq=0 --opt promo=0
run 2048 2048 128 --scalar --full --opt xyp=0
run 2048 2048 1024 --scalar --full --steps 3
run 2048 2048 1024 --scalar --full --steps 3 --opt xyp=0
