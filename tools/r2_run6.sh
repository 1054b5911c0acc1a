#!/bin/bash
# Round-2 GPU call: attention with quad-parallel K/V loads (one round trip per 128 keys), parity + bench + timeline.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run6.log
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. parity (persistent + interleave + long context + mistral)"
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "persistent or interleave or long_context or mistral" 2>&1 | tail -3
  echo "== 2. bench persistent (tg128, with CPU baseline + parity gate), depth 1024"
  timeout 600 python bench.py --no-pp --decode-mode persistent > gpurun_out/bench_r2f_persistent.json 2> gpurun_out/tmp.err; grep "^{" gpurun_out/bench_r2f_persistent.json | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('tg128', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', j.get('parity'), 'cpu', j.get('cpu_baseline'))"; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  timeout 600 python bench.py --no-pp --no-cpu --steps 64 --depth 1024 --decode-mode persistent 2> gpurun_out/tmp.err | grep "^{" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('tg64 -d 1024', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3))"; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  echo "== 3. timeline"
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2f_persistent.txt 2>&1; tail -32 gpurun_out/decode_timeline_r2f_persistent.txt
} 2>&1 | tee $L
