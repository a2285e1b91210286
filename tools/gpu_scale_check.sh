#!/bin/bash
# N-GPU box: NCCL parity (tests/dist_check.py) and bench.py --gpus N exactly as the driver launches it.
# usage: tools/gpu_scale_check.sh tag N
tag=${1:-s}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.txt 2>&1
echo "dist_check exit $?"; grep "dist_check\]" gpurun_out/${tag}_dist_check_${N}gpu.txt | cut -c1-170 | head -12
SECONDS=0
timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${tag}_scale_${N}gpu.jsonl 2> gpurun_out/${tag}_scale_${N}gpu.err
echo "bench exit $? after ${SECONDS}s"
python - <<P
import json
for l in open('gpurun_out/${tag}_scale_${N}gpu.jsonl'):
    if l.startswith('{'):
        d = json.loads(l)
        print('weak ms/step %.4f value %.4g' % (d['ms_per_step'], d['value']), {k: round(v, 4) for k, v in d['roofline']['kernel_ms'].items()})
        print('parity', d.get('parity'))
        c = d.get('c5_strong') or {}
        print('c5_strong', {k: c.get(k) for k in ('ms_per_step', 'value', 'efficiency_vs_c5_n1')})
        print('e2e', d['e2e'].get('value'))
P
tail -3 gpurun_out/${tag}_scale_${N}gpu.err
