#!/bin/bash
# Round-2 GPU call: persistent kernel v3 (16 consumer warps, relaxed polls, descriptors in registers): parity + timing + timeline.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run4.log
one() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-pp --no-cpu --steps 64 --decode-mode persistent 2>gpurun_out/tmp.err | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print('$label', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), j['roofline']['persistent_kernel'])
except Exception as e:
    print('$label FAILED', e)"
  tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
}
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 0. upload pipeline smoke (falls back to B200_UPLOAD_SYNC=1 for the rest of this script if it fails)"
  if ! timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "q8_bit_exact and tiny-llama-graph" 2>&1 | tail -3 | grep -q "1 passed"; then
    echo "UPLOAD PIPELINE FAILED -> B200_UPLOAD_SYNC=1"; export B200_UPLOAD_SYNC=1
  fi
  echo "== 1. parity file (both decode modes)"
  timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
  echo "== 2. bench"
  one "default"
  one "stages=10" B200_PD_STAGES=10
  one "l2ahead=24" B200_PD_L2_AHEAD=24
  timeout 300 python bench.py --no-pp --no-cpu --steps 64 --decode-mode graph 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('graph', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), {k: round(v['GB/s']) for k, v in j['roofline']['other_kernels'].items()})"
  echo "== 3. timeline"
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2d_persistent.txt 2>&1; tail -34 gpurun_out/decode_timeline_r2d_persistent.txt
} 2>&1 | tee $L
