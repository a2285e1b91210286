#!/bin/bash
# Runs on the B200 box: GPU tests, smoke, then bench.py exactly as the driver launches it (both arms).
# usage: tools/gpu_bench.sh tag [N]
tag=${1:-b}; N=${2:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -6 gpurun_out/${tag}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
  echo "ref exit $?"; cut -c1-600 gpurun_out/${tag}_bench_ref.json; grep -E "Elapsed|Maximum resident" gpurun_out/${tag}_bench_ref.err
  python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -12 gpurun_out/${tag}_bench.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/${tag}_bench_ref_n$N.json 2> gpurun_out/${tag}_bench_ref_n$N.err
  echo "ref exit $?"; cut -c1-300 gpurun_out/${tag}_bench_ref_n$N.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
  echo "bench exit $?"; cat gpurun_out/${tag}_bench_n$N.json; tail -15 gpurun_out/${tag}_bench_n$N.err
fi
