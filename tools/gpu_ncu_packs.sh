#!/bin/bash
# B200 box: one ncu full capture of the word-form pack builder at 1024^3; raw page as CSV.  usage: tools/gpu_ncu_packs.sh tag
tag=${1:-packs}
mkdir -p gpurun_out /tmp/rep
ncu --set full --clock-control none --import-source on -k regex:k_build_packs_v -s 2 -c 1 -f -o /tmp/rep/packs python tools/packs_probe.py 1024 --one > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ncu -i /tmp/rep/packs.ncu-rep --page raw --csv > gpurun_out/${tag}_packs_raw.csv 2>/dev/null
ncu -i /tmp/rep/packs.ncu-rep --page source --csv > gpurun_out/${tag}_packs_source.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
