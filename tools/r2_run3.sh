#!/bin/bash
# Round-2 GPU call 4: persistent kernel v2 (in-flight throttle, out-of-line phases, cp.async norm weights, KV L2 prefetch): parity + knob sweep.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run3.log
one() { # label, env...
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
  echo "== 1. parity file (both decode modes)"
  timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
  echo "== 2. knob sweep (persistent, tg64)"
  one "maxfly=6(default)"
  one "maxfly=0" B200_PD_MAXFLY=0
  one "maxfly=2" B200_PD_MAXFLY=2
  one "maxfly=4" B200_PD_MAXFLY=4
  one "maxfly=10" B200_PD_MAXFLY=10
  one "maxfly=4,stages=8" B200_PD_MAXFLY=4 B200_PD_STAGES=8
  one "maxfly=6,stages=12" B200_PD_MAXFLY=6 B200_PD_STAGES=12
  one "maxfly=3,l2ahead=16" B200_PD_MAXFLY=3 B200_PD_L2_AHEAD=16
  echo "== 3. timeline (default knobs), then maxfly=0"
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2c_persistent.txt 2>&1; tail -34 gpurun_out/decode_timeline_r2c_persistent.txt
  B200_PD_MAXFLY=0 timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2c_persistent_fly0.txt 2>&1; tail -34 gpurun_out/decode_timeline_r2c_persistent_fly0.txt | head -24
} 2>&1 | tee $L
