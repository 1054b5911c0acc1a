#!/bin/bash
# Round-2 GPU call: L2 eviction policy on the weight stream (evict_first), with and without L2 look-ahead; KV prefetch moved; parity subset.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run5.log
one() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-pp --no-cpu --steps 64 --decode-mode persistent 2>gpurun_out/tmp.err | grep "^{" | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print('$label', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), j['roofline']['persistent_kernel'], 'upload', j['load']['pipeline'])
except Exception as e:
    print('$label FAILED', e)"
  tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
}
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. parity (persistent + interleave + long context + sampler)"
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampler.py -x -q -k "persistent or interleave or long_context or sampler or mistral" 2>&1 | tail -3
  echo "== 2. bench sweep"
  one "evict_first=1(default)"
  one "evict_first=0" B200_PD_EVICT_FIRST=0
  one "evict_first=1,l2ahead=24" B200_PD_L2_AHEAD=24
  one "evict_first=1,l2ahead=64" B200_PD_L2_AHEAD=64
  one "evict_first=1,l2ahead=64,stages=8" B200_PD_L2_AHEAD=64 B200_PD_STAGES=8
  echo "== 3. timeline (default)"
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2e_persistent.txt 2>&1; tail -34 gpurun_out/decode_timeline_r2e_persistent.txt
  echo "== 4. Qwen3-4B (BASELINE config 4), both modes"
  for m in persistent graph; do
    timeout 400 python bench.py --workload qwen3-4b --no-pp --decode-mode $m > gpurun_out/bench_r2_qwen3-4b_$m.json 2> gpurun_out/tmp.err
    grep "^{" gpurun_out/bench_r2_qwen3-4b_$m.json | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('qwen3-4b $m', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', j.get('parity'), 'cpu', j.get('cpu_baseline'))"
    tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  done
} 2>&1 | tee $L
