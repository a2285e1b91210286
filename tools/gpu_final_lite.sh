#!/bin/bash
# B200 box, one GPU: GPU tests, smoke, bench.py both arms exactly as the driver launches them, and the ncu launch list of
# a short bench run (no full captures: use gpu_final.sh for those).  usage: tools/gpu_final_lite.sh tag
tag=${1:-f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
SECONDS=0
python bench.py --impl reference --gpus 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
echo "ref exit $? after ${SECONDS}s"; cut -c1-300 gpurun_out/${tag}_bench_ref.json
SECONDS=0
python bench.py --gpus 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $? after ${SECONDS}s"; cut -c1-400 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 --long-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
tail -3 gpurun_out/${tag}_launches.csv | cut -c1-200
