#!/bin/bash
# Tensor-parallel validation + timing on N GPUs of one box:  gpurun --gpus N -- 'bash tools/r2_tp.sh N [workload]'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
WL=${2:-llama-3-8b}
mkdir -p gpurun_out
L=gpurun_out/r2_tp${N}_${WL}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
{
  nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
  MID=mid-llama; [ "$WL" = "llama-3-70b" ] && MID=mid-llama-70b
  echo "== 1. parity vs the oracle, both decode modes ($MID, tp$N)"
  timeout 400 $TR tools/tp_check.py $MID 6 2>&1 | grep -E "^\[tp|RESULT|rror" | tail -8
  echo "== 2. bench $WL tp$N: persistent, graph"
  for m in persistent graph; do
    extra=""; [ "$m" = "graph" ] && extra="--no-cpu"   # the oracle-backed parity gate runs once (persistent mode)
    [ "$WL" = "llama-3-70b" ] && extra="--no-cpu"      # a 70B oracle step takes minutes of host time: parity is covered by the 2-layer cut above
    timeout 1200 $TR bench.py --gpus $N --workload $WL --no-pp --steps 64 --decode-mode $m $extra > gpurun_out/bench_r2_tp${N}_${WL}_$m.json 2> gpurun_out/bench_r2_tp${N}_${WL}_$m.err
    python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/bench_r2_tp${N}_${WL}_$m.json") if l.startswith("{")][-1])
    print("$m", "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms", round(j["ms_per_step"], 3), "by rank", [round(v, 3) for v in j["ms_per_step_by_rank"]],
          "roofline", round(j["roofline"]["frac"], 3), "parity", j.get("parity"))
except Exception as e:
    print("$m FAILED", e)
PY
    tail -3 gpurun_out/bench_r2_tp${N}_${WL}_$m.err
  done
} 2>&1 | tee $L
