#!/bin/bash
# Runs on the B200 box: GPU tests, then short device-resident bench runs over launch-shape options.
# usage: tools/gpu_tune.sh tag "opts1" "opts2" ...
tag=${1:-t}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.jsonl
if [ $# -eq 0 ]; then set -- ""; fi
for o in "$@"; do
  echo "== $o" | tee -a gpurun_out/${tag}_tune.jsonl
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 $o 2>>gpurun_out/${tag}_tune.err | tee -a gpurun_out/${tag}_tune.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('ms/step %.3f  value %.3e  sweeps %s  step_frac %.3f' % (d['ms_per_step'], d['value'], {k: round(v,3) for k,v in r.get('kernel_ms', r.get('sweep_ms', {})).items()}, r['step_frac']))"
done
