#!/bin/bash
# B200 box, one GPU: the round's evidence run -- GPU tests, smoke, bench.py both arms exactly as the driver launches them,
# ncu launch list of the bench command, ncu full capture of the step's kernels (raw page as CSV; the reports stay on the box).
# usage: tools/gpu_final.sh tag
tag=${1:-f}
mkdir -p gpurun_out /tmp/rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
python bench.py --impl reference --gpus 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
echo "ref exit $?"; cut -c1-300 gpurun_out/${tag}_bench_ref.json
python bench.py --gpus 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cut -c1-400 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
if [ "${NO_NCU:-0}" != "1" ]; then
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 --long-steps 0"
$B > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_explicit" -s 12 -c 4 -f -o /tmp/rep/prof $B > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log
ncu -i /tmp/rep/prof.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_raw.csv 2>/dev/null
C="python tools/cyl_probe.py 256 1024 512 --steps 2"
$C > gpurun_out/${tag}_cyl.log 2>&1
ncu --set full --clock-control none -k regex:k_cyl -s 9 -c 3 -f -o /tmp/rep/prof_cyl $C > gpurun_out/${tag}_ncu3.log 2>&1
ncu -i /tmp/rep/prof_cyl.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_cyl_raw.csv 2>/dev/null
rm -f /tmp/rep/*.ncu-rep
fi
ls -la gpurun_out | tail -15
