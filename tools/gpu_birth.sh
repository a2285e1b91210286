#!/bin/bash
# B200 box, one GPU: GPU tests, then configs[3] (deposition loop) with the word-form mask kernels on / off.
# usage: tools/gpu_birth.sh tag
tag=${1:-birth}
mkdir -p gpurun_out
quick=${2:-}
if [ -z "$quick" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
fi
for o in "" "--opt maskv=0 --no-cpu"; do
  n=$( [ -z "$o" ] && echo on || echo off )
  timeout 600 python bench.py --workload c4 --steps 16 --warmup 4 $o > gpurun_out/${tag}_c4_${n}.json 2> gpurun_out/${tag}_c4_${n}.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_c4_${n}.json') if l.startswith('{')][-1]); r=d['roofline']
print('maskv ${n}:', 'ms/step %.3f steady %.3f birth %.3f' % (d['ms_per_step'], r['steady_ms_per_step'], r['birth_ms']), 'parity', d['parity'])
print(r.get('birth_ms_each'))"
  tail -2 gpurun_out/${tag}_c4_${n}.err
done
[ -n "$quick" ] && exit 0
B="python bench.py --workload c4 --steps 8 --warmup 4 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_build|k_transpose|k_tile" -c 40 --csv --log-file gpurun_out/${tag}_c4_launches.csv $B > gpurun_out/${tag}_ncu.log 2>&1
tail -12 gpurun_out/${tag}_c4_launches.csv | cut -d, -f5,9,15 | cut -c1-160
