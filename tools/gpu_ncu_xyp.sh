#!/bin/bash
# B200 box: ncu full captures of the long-line x / y sweeps (persistent and plain) and of the z sweep at 512^3.
tag=${1:-n}
mkdir -p gpurun_out
P="python tools/sweep_probe.py"
$P 2048 2048 128 --scalar --full --steps 2 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_xyp -s 6 -c 2 -f -o gpurun_out/${tag}_xyp \
    $P 2048 2048 128 --scalar --full --steps 2 > gpurun_out/${tag}_ncu1.log 2>&1; tail -2 gpurun_out/${tag}_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:k_sweep_xy -s 6 -c 2 -f -o gpurun_out/${tag}_xy \
    $P 2048 2048 128 --scalar --full --steps 2 --opt xyp=0 > gpurun_out/${tag}_ncu2.log 2>&1; tail -2 gpurun_out/${tag}_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:k_sweep_zt -s 3 -c 1 -f -o gpurun_out/${tag}_zt \
    $P 512 512 512 --steps 2 > gpurun_out/${tag}_ncu3.log 2>&1; tail -2 gpurun_out/${tag}_ncu3.log
ncu --set full --clock-control none --import-source on -k regex:k_sweep_zt -s 3 -c 1 -f -o gpurun_out/${tag}_zt_full \
    $P 512 512 512 --full --steps 2 > gpurun_out/${tag}_ncu4.log 2>&1; tail -2 gpurun_out/${tag}_ncu4.log
tools/gpu_cyl_probe.sh ${tag}
