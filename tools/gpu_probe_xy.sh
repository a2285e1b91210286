#!/bin/bash
# Runs on the B200 box: GPU tests, then per-sweep device times of sweep variants (tools/sweep_probe.py).
# usage: tools/gpu_probe_xy.sh tag [ncu]
tag=${1:-p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_probe.txt
: > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 512 512 512
run 512 512 512 --opt bulk=0
run 512 512 512 --opt zt=0
run 512 512 512 --opt lt=8
run 512 512 512 --scalar
run 512 512 512 --scalar --full
run 1024 1024 128
run 1024 1024 128 --opt zt=0
run 512 512 1024 --scalar --full
run 512 512 1024
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt zt=0
if [ "$2" = "ncu" ]; then
    ncu --set full --clock-control none --import-source on -k regex:"k_sweep_z" -s 4 -c 1 -f -o /tmp/${tag}_prof \
        python tools/sweep_probe.py 512 512 512 --steps 2 > gpurun_out/${tag}_ncu.log 2>&1
    tail -2 gpurun_out/${tag}_ncu.log
    ncu -i /tmp/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw.csv 2>/dev/null
fi
