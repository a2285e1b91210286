#!/bin/bash
tag=${1:-e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cart.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/sweep_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 512 512 512
run 512 512 512 --opt eorder=1
run 1024 1024 256
run 1024 1024 256 --opt eorder=1
run 2048 2048 128 --scalar --full
run 2048 2048 128 --scalar --full --opt eorder=1
python bench.py --workload c4 --steps 16 --warmup 4 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('c4 steady %.3f birth %.3f' % (r['steady_ms_per_step'], r['birth_ms']), r['kernel_ms'])"
python bench.py --workload c4 --steps 16 --warmup 4 --no-cpu --opt eorder=1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('c4 eorder=1 steady %.3f birth %.3f' % (r['steady_ms_per_step'], r['birth_ms']), r['kernel_ms'])"
