#!/bin/bash
# Round-2 final validation A: whole GPU suite on the final build, smoke, both decode modes benched on the same box, prefill variants, load bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_final_a.log
line() { grep "^{" "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
pp = j.get('pp512') or {}
print('$2', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', (j.get('parity') or {}).get('ids_equal'), (j.get('parity') or {}).get('logits_bit_equal'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'pp512', round(pp.get('value', 0)), round((pp.get('roofline') or {}).get('frac', 0), 3))"; }
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. pytest -m gpu (whole suite)"
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== 2. smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== 3. bench, both decode modes, same box"
  timeout 900 python bench.py --decode-mode persistent > gpurun_out/bench_r2_final_persistent.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_persistent.json persistent; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  timeout 900 python bench.py --decode-mode graph --no-pp > gpurun_out/bench_r2_final_graph.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_graph.json graph; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  echo "== 3b. Llama-3.2-1B FP16 (BASELINE config 1 shape; exact lane-order FP16 matvec, graph mode)"
  timeout 600 python bench.py --workload llama-3.2-1b --quant f16 --no-pp > gpurun_out/bench_r2_final_1b_f16.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_1b_f16.json 1b-f16; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback"
  echo "== 4. prefill: persistent residual GEMMs (split-K work items) -- stand-alone check, tests, pp512"
  B200_GEMM_2CTA=1256 B200_GEMM_RESID=2 timeout 120 python tools/gemm_check.py --big 2>&1 | tail -6
  B200_GEMM_PERSIST_RESID=1 timeout 400 python -m pytest tests/test_gpu_prefill.py -x -q 2>&1 | tail -3
  B200_GEMM_PERSIST_RESID=1 timeout 200 python tools/pp_bench.py llama-3-8b 512 5 2>&1 | tail -1
  timeout 200 python tools/pp_bench.py llama-3-8b 512 5 2>&1 | tail -1
  echo "== 5. timelines"
  B200_DECODE=graph timeout 200 python tools/trace.py llama-3-8b 64 > gpurun_out/decode_timeline_r2_final_graph.txt 2>&1; tail -12 gpurun_out/decode_timeline_r2_final_graph.txt
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2_final_persistent.txt 2>&1; tail -32 gpurun_out/decode_timeline_r2_final_persistent.txt | head -24
  echo "== 6. load from a GGUF file on disk"
  timeout 600 python tools/load_bench.py llama-3-8b /tmp 2>&1 | tail -1 | cut -c1-1500
} 2>&1 | tee $L
