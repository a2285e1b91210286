#!/bin/bash
# Runs on an N-GPU B200 box (gpurun --gpus N): NCCL parity check + scaling runs of bench.py.
# usage: tools/gpu_scale.sh tag "N list weak" "N list c5"
tag=${1:-s}; weak=${2:-"8 4"}; c5=${3:-"8 4 2"}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nmax=$(nvidia-smi -L | wc -l)
timeout 300 $T --nproc-per-node $nmax --master-port 29520 tests/dist_check.py 2>&1 | grep dist_check | tee gpurun_out/${tag}_dist_check.txt
: > gpurun_out/${tag}_scale.jsonl
port=29530
for n in $weak; do
  port=$((port+1))
  if [ "$n" = "1" ]; then timeout 600 python bench.py --steps 10 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/${tag}_scale.jsonl
  else timeout 600 $T --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 10 2>/dev/null | tail -1 >> gpurun_out/${tag}_scale.jsonl; fi
done
for n in $c5; do
  port=$((port+1))
  if [ "$n" = "1" ]; then timeout 900 python bench.py --workload c5 --steps 5 2>/dev/null | tail -1 >> gpurun_out/${tag}_scale.jsonl
  else timeout 900 $T --nproc-per-node $n --master-port $port bench.py --gpus $n --workload c5 --steps 5 2>/dev/null | tail -1 >> gpurun_out/${tag}_scale.jsonl; fi
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_scale.jsonl"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    r = d["roofline"]
    print(d["n_gpus"], d["scaling"], d["config"]["grid"], f'{d["value"]/1e9:.1f} Gcs/s', f'{d["ms_per_step"]:.3f} ms', {k: round(v, 3) for k, v in r["kernel_ms"].items()}, "e2e", d["e2e"]["value"] and round(d["e2e"]["value"]/1e9, 2))
PY
