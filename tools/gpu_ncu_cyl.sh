#!/bin/bash
tag=${1:-c}
mkdir -p gpurun_out /tmp/rep
C="python tools/cyl_probe.py 256 1024 512 --steps 2"
$C > gpurun_out/${tag}_cyl_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_cyl -s 9 -c 3 -f -o /tmp/rep/cyl $C > gpurun_out/${tag}_cyl_ncu.log 2>&1
tail -1 gpurun_out/${tag}_cyl_ncu.log
ncu -i /tmp/rep/cyl.ncu-rep --page raw --csv > gpurun_out/${tag}_cyl_raw.csv 2>/dev/null
ncu -i /tmp/rep/cyl.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_cyl_source.csv.gz
rm -f /tmp/rep/cyl.ncu-rep; ls -la gpurun_out | tail -4
