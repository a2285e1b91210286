#!/bin/bash
tag=${1:-c}
mkdir -p gpurun_out
out=gpurun_out/${tag}_cyl_probe.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python tools/cyl_probe.py "$@" >> $out 2>&1; tail -1 $out; }
run 256 1024 512
run 256 1024 512 --opt kt=16
run 256 1024 512 --opt kt=32
run 256 1024 512 --opt m=32
run 256 1024 512 --opt m=32 --opt kt=16
run 256 1024 512 --opt m=16
run 256 1024 512 --opt m=16 --opt kt=16
run 256 1024 512 --opt kt=4
run 256 1024 512 --masked
