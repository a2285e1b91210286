#!/bin/bash
# Runs on the B200 box (via gpurun): GPU tests, bench (both arms), ncu launch list + one full capture.
# Usage: tools/gpu_check.sh [tag]     outputs -> gpurun_out/<tag>_*
tag=${1:-r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_ref.json
if [ "${NO_NCU:-0}" != "1" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_explicit" -s 12 -c 4 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_ncu2.log
python tools/cyl_probe.py 256 1024 512 > gpurun_out/${tag}_cyl.log 2>&1; python tools/cyl_probe.py 256 1024 512 --masked >> gpurun_out/${tag}_cyl.log 2>&1; cat gpurun_out/${tag}_cyl.log
ncu --set full --clock-control none --import-source on -k regex:k_cyl -s 9 -c 3 -f -o gpurun_out/${tag}_prof_cyl \
    python tools/cyl_probe.py 256 1024 512 --steps 2 > gpurun_out/${tag}_ncu3.log 2>&1
tail -3 gpurun_out/${tag}_ncu3.log
fi
if [ "${C4:-0}" = "1" ]; then
timeout 600 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/${tag}_c4.json 2> gpurun_out/${tag}_c4.err; cat gpurun_out/${tag}_c4.json; tail -2 gpurun_out/${tag}_c4.err
fi
