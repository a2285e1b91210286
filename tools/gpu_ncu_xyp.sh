#!/bin/bash
# B200 box: ncu full captures of the long-line x / y sweeps (persistent and plain) and of the z sweep at 512^3.
# The reports stay on the box (too large to travel); their raw and source pages come back as gzipped CSV.
tag=${1:-n}
mkdir -p gpurun_out /tmp/rep
P="python tools/sweep_probe.py"
cap() {  # name kernel-regex skip count probe-args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/rep/$name $P "$@" > gpurun_out/${tag}_${name}_ncu.log 2>&1
  tail -1 gpurun_out/${tag}_${name}_ncu.log
  ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/${tag}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_${name}_source.csv.gz
  rm -f /tmp/rep/$name.ncu-rep
}
$P 2048 2048 128 --scalar --full --steps 2 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
cap xyp k_sweep_xyp 6 2 2048 2048 128 --scalar --full --steps 2 --opt xyp=1
cap xy k_sweep_xy 6 2 2048 2048 128 --scalar --full --steps 2 --opt xyp=0
cap zt k_sweep_zt 3 1 512 512 512 --steps 2
cap zt_full k_sweep_zt 3 1 512 512 512 --full --steps 2
ls -la gpurun_out | tail -12
