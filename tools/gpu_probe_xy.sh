#!/bin/bash
# Runs on the B200 box: GPU tests, then per-sweep device times of the x / y sweep variants (tools/sweep_probe.py).
# usage: tools/gpu_probe_xy.sh tag [ncu]
tag=${1:-p}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cart.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_probe.txt
: > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 512 512 512
run 512 512 512 --opt tw=0
run 512 512 512 --opt occ=3
run 512 512 512 --opt occ=3 --opt tw=0
run 512 512 512 --opt occ=4
run 512 512 512 --opt occ=4 --opt tw=0
run 512 512 512 --opt m=32
run 512 512 512 --opt m=32 --opt tw=0
run 512 512 512 --opt wide=1
run 512 512 512 --opt wide=1 --opt tw=0
run 512 512 512 --opt dbg=1 --opt wide=1
run 512 512 512 --scalar --opt occ=3
run 512 512 512 --scalar --opt m=32
run 512 512 512 --scalar --opt wide=1
run 1024 1024 128
run 1024 1024 128 --opt tw=0
run 2048 2048 64 --scalar --full
run 2048 2048 64 --scalar --full --opt tw=0
if [ "$2" = "ncu" ]; then
  for v in "occ=3" "m=32"; do
    n=$(echo $v | tr -d '=')
    ncu --set full --clock-control none --import-source on -k regex:k_sweep_xy -s 6 -c 2 -f -o gpurun_out/${tag}_prof_$n \
        python tools/sweep_probe.py 512 512 512 --steps 2 --opt $v > gpurun_out/${tag}_ncu_$n.log 2>&1
    tail -2 gpurun_out/${tag}_ncu_$n.log
  done
fi
