#!/bin/bash
# B200 box: z sweep variants (hybrid surface chunks).  usage: tools/gpu_z.sh tag
tag=${1:-z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cart.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_z_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 512 512 512
run 512 512 512 --opt hyb=0
run 512 512 512 --scalar
run 512 512 512 --scalar --opt hyb=0
run 512 512 512 --opt lt=16
run 512 512 512 --opt lt=4
run 1024 1024 256
run 1024 1024 256 --opt hyb=0
