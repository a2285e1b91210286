#!/bin/bash
# Runs on the B200 box: GPU tests, then per-sweep device times of the x / y sweep variants (tools/sweep_probe.py).
# usage: tools/gpu_probe_xy.sh tag
tag=${1:-p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_probe.txt
: > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 512 512 512
run 512 512 512 --opt xy2=0
run 512 512 512 --opt uni=0
run 512 512 512 --opt tw=0
run 512 512 512 --opt m=32
run 512 512 512 --opt kt=16
run 512 512 512 --opt kt=4
run 512 512 512 --scalar
run 512 512 512 --scalar --opt xy2=0
run 512 512 512 --scalar --opt m=32
run 1024 1024 128
run 1024 1024 128 --opt xy2=0
run 2048 2048 64 --scalar --full
run 2048 2048 64 --scalar --full --opt xy2=0
run 2048 2048 64 --scalar --full --opt uni=0
