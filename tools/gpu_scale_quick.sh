#!/bin/bash
# N-GPU box: NCCL parity, then short weak-scaling runs of the plate over dist options.  usage: tools/gpu_scale_quick.sh tag N
tag=${1:-q}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.txt 2>&1
echo "dist_check exit $?"; grep "dist_check\]" gpurun_out/${tag}_dist_check_${N}gpu.txt | cut -c1-170 | head -12
for o in "" "--opt overlap_halo=0" "--opt spike_thr_log2=-56" "--opt batches=1"; do
  timeout 300 $TR --master-port 29518 bench.py --gpus $N --steps 30 --warmup 8 --no-extras --e2e-steps 2 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$o', 'ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in r['kernel_ms'].items()}, d['exchange']['z_form'])" | tee -a gpurun_out/${tag}_quick_n$N.txt
done
