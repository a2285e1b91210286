#!/bin/bash
# B200 box: one ncu full capture of a sweep_probe run; raw + source pages come back as CSV.
# usage: tools/gpu_ncu_one.sh tag name kernel-regex skip count probe-args...
tag=$1; name=$2; rx=$3; skip=$4; cnt=$5; shift 5
mkdir -p gpurun_out /tmp/rep
P="python tools/sweep_probe.py"
$P "$@" > gpurun_out/${tag}_${name}_plain.log 2>&1 || exit 1
tail -1 gpurun_out/${tag}_${name}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/rep/$name $P "$@" > gpurun_out/${tag}_${name}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_${name}_ncu.log
ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/${tag}_${name}_raw.csv 2>/dev/null
ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_${name}_source.csv.gz
rm -f /tmp/rep/$name.ncu-rep
