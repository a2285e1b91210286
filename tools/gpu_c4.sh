#!/bin/bash
tag=${1:-c4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cart.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
for o in "" "--opt tiles=0"; do
  timeout 600 python bench.py --workload c4 --steps 16 --warmup 4 $o > gpurun_out/${tag}_c4.json 2> gpurun_out/${tag}_c4.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_c4.json') if l.startswith('{')][-1]); r=d['roofline']
print('$o', 'ms/step %.3f steady %.3f birth %.3f' % (d['ms_per_step'], r['steady_ms_per_step'], r['birth_ms']), 'parity', d['parity'])"
  tail -2 gpurun_out/${tag}_c4.err
done
python tools/sweep_probe.py 512 512 512 | tail -1
