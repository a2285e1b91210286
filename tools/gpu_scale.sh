#!/bin/bash
# Runs on an N-GPU B200 box: NCCL parity of the z-slab path (tests/dist_check.py, both sequencings), then bench.py as the
# driver launches it.   usage: tools/gpu_scale.sh tag N [extra bench args]
tag=${1:-s}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.txt 2>&1
echo "dist_check (library sequencing) exit $?" | tee -a gpurun_out/${tag}_dist_check_${N}gpu.txt
grep "dist_check\]" gpurun_out/${tag}_dist_check_${N}gpu.txt | cut -c1-200; tail -5 gpurun_out/${tag}_dist_check_${N}gpu.txt | cut -c1-300
timeout 600 $TR --master-port 29512 tests/dist_check.py --python-seq > gpurun_out/${tag}_dist_check_py_${N}gpu.txt 2>&1
echo "dist_check (python sequencing) exit $?" | tee -a gpurun_out/${tag}_dist_check_py_${N}gpu.txt
grep "dist_check\]" gpurun_out/${tag}_dist_check_py_${N}gpu.txt | cut -c1-200
timeout 900 $TR --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench exit $?"; cat gpurun_out/${tag}_bench_n$N.json; tail -15 gpurun_out/${tag}_bench_n$N.err | cut -c1-400
timeout 300 $TR --master-port 29519 tools/cyl_slab_probe.py > gpurun_out/${tag}_cyl_slab_n$N.json 2> gpurun_out/${tag}_cyl_slab_n$N.err
echo "cyl slab exit $?"; cat gpurun_out/${tag}_cyl_slab_n$N.json; tail -3 gpurun_out/${tag}_cyl_slab_n$N.err | cut -c1-300
