#!/bin/bash
# Round-2 run 9: FP16 streaming matvec v2 (chain pairs, packed FTZ flush), Phi-3, depth lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run9.log
line() { grep "^{" "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$2', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', (j.get('parity') or {}).get('ids_equal'), (j.get('parity') or {}).get('logits_bit_equal'), 'launches', j.get('gpu_launches'), 'mode', j.get('decode_mode'), {k: round(v['GB/s']) for k, v in j['roofline']['other_kernels'].items()})"; }
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. GPU tests: FP16 rings v2, Phi-3"
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "f16 or phi3" 2>&1 | tail -6
  timeout 600 python -m pytest tests/test_gpu_prefill.py -x -q -k "phi3 or (within and tiny-qwen3-37)" 2>&1 | tail -3
  echo "== 2. FP16 decode: Llama-3.2-1B (BASELINE config 1), Llama-3-8B (config 3 model)"
  timeout 600 python bench.py --workload llama-3.2-1b --quant f16 --no-pp > gpurun_out/bench_r2_1b_f16_v2.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_1b_f16_v2.json 1b-f16-v2; tail -3 gpurun_out/tmp.err | grep -i -E "error|Traceback|PARITY"
  timeout 900 python bench.py --workload llama-3-8b --quant f16 --no-pp --cpu-budget 12 > gpurun_out/bench_r2_8b_f16_v2.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_8b_f16_v2.json 8b-f16-v2; tail -3 gpurun_out/tmp.err | grep -i -E "error|Traceback|PARITY"
  echo "== 3. FP16 timeline (1B)"
  timeout 300 python tools/trace.py llama-3.2-1b 64 f16 > gpurun_out/decode_timeline_r2_1b_f16.txt 2>&1; tail -11 gpurun_out/decode_timeline_r2_1b_f16.txt
  echo "== 4. depth lines (8B Q8_0, graph): tg64 -d 1024, tg32 -d 4096"
  timeout 600 python bench.py --depth 1024 --steps 64 --no-cpu --no-pp > gpurun_out/bench_r2_d1024.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_d1024.json d1024
  timeout 600 python bench.py --depth 4096 --steps 32 --no-cpu --no-pp > gpurun_out/bench_r2_d4096.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_d4096.json d4096
  timeout 300 python tools/trace.py llama-3-8b 4096 > gpurun_out/decode_timeline_r2_d4096.txt 2>&1; tail -11 gpurun_out/decode_timeline_r2_d4096.txt
} 2>&1 | tee $L
