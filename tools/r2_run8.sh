#!/bin/bash
# Round-2 run 8: first GPU execution of the FP16 streaming matvec (stream_matvec_f16.cuh) and of the K-quant load path (kquant.cuh).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run8.log
line() { grep "^{" "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$2', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', (j.get('parity') or {}).get('ids_equal'), (j.get('parity') or {}).get('logits_bit_equal'), 'launches', j.get('gpu_launches'), 'mode', j.get('decode_mode'), {k: round(v['GB/s']) for k, v in j['roofline']['other_kernels'].items()})"; }
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. new GPU tests: K-quants, FP16 streaming kernels"
  timeout 900 python -m pytest tests/test_gpu_kquants.py -x -q 2>&1 | tail -6
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "f16" 2>&1 | tail -6
  timeout 600 python -m pytest tests/test_gpu_prefill.py -x -q -k "tensor_core_prefill_within or mid_llama" 2>&1 | tail -3
  echo "== 2. Llama-3.2-1B FP16 (BASELINE config 1): streaming vs round-1 kernels"
  timeout 600 python bench.py --workload llama-3.2-1b --quant f16 --no-pp > gpurun_out/bench_r2_1b_f16_stream.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_1b_f16_stream.json 1b-f16-stream; tail -3 gpurun_out/tmp.err | grep -i -E "error|Traceback|PARITY"
  B200_F16_STREAM=0 timeout 600 python bench.py --workload llama-3.2-1b --quant f16 --no-pp --no-cpu > gpurun_out/bench_r2_1b_f16_round1.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_1b_f16_round1.json 1b-f16-round1
  echo "== 3. Llama-3-8B FP16 decode (BASELINE config 3 model)"
  timeout 900 python bench.py --workload llama-3-8b --quant f16 --no-pp --cpu-budget 12 > gpurun_out/bench_r2_8b_f16_stream.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_8b_f16_stream.json 8b-f16-stream; tail -3 gpurun_out/tmp.err | grep -i -E "error|Traceback|PARITY"
  echo "== 4. FP16 timeline (1B)"
  timeout 300 python tools/trace.py llama-3.2-1b 64 f16 > gpurun_out/decode_timeline_r2_1b_f16.txt 2>&1; tail -12 gpurun_out/decode_timeline_r2_1b_f16.txt
  echo "== 5. default bench (graph is the default now)"
  timeout 900 python bench.py --no-pp --no-cpu > gpurun_out/bench_r2_default_check.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_default_check.json default
} 2>&1 | tee $L
