#!/bin/bash
tag=${1:-u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cart.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_xyu_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out | cut -c1-175; }
run 2048 2048 1024 --scalar --full --steps 3
run 2048 2048 1024 --scalar --steps 3
run 2048 2048 512 --scalar --full --steps 3
run 2048 2048 512 --scalar --full --steps 3 --opt ukt=8
run 2048 2048 512 --scalar --full --steps 3 --opt ukt=16
run 2048 2048 256 --scalar --full --steps 3
run 2048 2048 256 --scalar --full --steps 3 --opt ukt=16
run 2048 2048 128 --scalar
run 2048 2048 128 --opt ukt=16
