#!/bin/bash
# B200 box: persistent long-line x / y sweeps (TMA tiles).  usage: tools/gpu_xyp.sh tag
tag=${1:-p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cart.py -m gpu -q -x -k "long_lines or uniform_chunk" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_xyp_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt xyp=0
run 2048 2048 128 --scalar
run 2048 2048 128
run 2048 2048 128 --opt xyp=0
run 1536 1536 256 --scalar
run 1536 1536 256 --scalar --opt xyp=0
