#!/bin/bash
# Runs on the B200 box (via gpurun): output-path tests, throughput probe, ncu capture of the text kernels.
# Usage: tools/gpu_output.sh [tag]     outputs -> gpurun_out/<tag>_*
tag=${1:-r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_output.py -x -q > gpurun_out/${tag}_pytest_output.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_output.log
tail -25 gpurun_out/${tag}_pytest_output.log
timeout 600 python tools/output_probe.py --out gpurun_out/${tag}_output_probe.json > gpurun_out/${tag}_output_probe.log 2>&1
echo "probe exit $?"; tail -5 gpurun_out/${tag}_output_probe.log
if [ "${NO_NCU:-0}" != "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_text -s 6 -c 3 -f -o gpurun_out/${tag}_prof_text \
    python tools/output_probe.py --n 256 > gpurun_out/${tag}_ncu_text.log 2>&1
tail -3 gpurun_out/${tag}_ncu_text.log
fi
