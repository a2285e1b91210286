#!/bin/bash
# N-GPU box: slab tests on one GPU, NCCL parity, then short runs of the plate (weak) and of configs[4] (strong) over
# dist options.  usage: tools/gpu_scale_quick.sh tag N
tag=${1:-q}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_cart.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.txt 2>&1
echo "dist_check exit $?"; grep "dist_check\]" gpurun_out/${tag}_dist_check_${N}gpu.txt | cut -c1-170 | head -12
for o in "" "--opt spike_thr_log2=-80" "--workload c5 --steps 5" "--workload c5 --steps 5 --opt spike_thr_log2=-80"; do
  timeout 300 $TR --master-port 29518 bench.py --gpus $N --steps 30 --warmup 8 --no-extras --e2e-steps 2 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$o', 'ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in r['kernel_ms'].items()}, d['exchange']['z_form'])" | tee -a gpurun_out/${tag}_quick_n$N.txt
done
