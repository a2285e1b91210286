#!/bin/bash
# B200 box: explicit-stage launch shapes.  usage: tools/gpu_explicit.sh tag
tag=${1:-e}
mkdir -p gpurun_out
out=gpurun_out/${tag}_explicit_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out | cut -c1-120; }
run 512 512 512
run 512 512 512 --opt ejt=32
run 512 512 512 --opt ejt=64
run 512 512 512 --opt ejt=8
run 512 512 512 --opt eth=64
run 512 512 512 --opt eth=64 --opt ejt=32
run 512 512 512 --opt eorder=1
run 512 512 512 --opt eorder=1 --opt ejt=32
run 2048 2048 128 --scalar --full --opt ejt=32
run 2048 2048 128 --scalar --full --opt ejt=64
