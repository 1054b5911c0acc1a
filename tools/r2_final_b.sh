#!/bin/bash
# Round-2 final run: whole GPU suite on the final build (incl. the pipelined attention, deep-context tests, FP16 classifier partials),
# smoke, default bench, depth lines, FP16 line, ncu launch list of the default bench command.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_final_b.log
line() { grep "^{" "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
pp = j.get('pp512') or {}
print('$2', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'parity', (j.get('parity') or {}).get('ids_equal'), (j.get('parity') or {}).get('logits_bit_equal'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'pp512', round(pp.get('value', 0)), 'mode', j.get('decode_mode'))"; }
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. deep-context tests first (new attention code), then the whole suite"
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deep_context" 2>&1 | tail -4
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== 2. smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== 3. default bench (exactly the driver's command)"
  timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final.json default; tail -2 gpurun_out/tmp.err | grep -i -E "error|Traceback|PARITY"
  echo "== 4. depth lines: tg64 -d 1024, tg32 -d 4096 (+ timeline at 4096)"
  timeout 600 python bench.py --depth 1024 --steps 64 --no-cpu --no-pp > gpurun_out/bench_r2_final_d1024.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_d1024.json d1024
  timeout 600 python bench.py --depth 4096 --steps 32 --no-cpu --no-pp > gpurun_out/bench_r2_final_d4096.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_d4096.json d4096
  timeout 300 python tools/trace.py llama-3-8b 4096 > gpurun_out/decode_timeline_r2_final_d4096.txt 2>&1; tail -11 gpurun_out/decode_timeline_r2_final_d4096.txt
  timeout 300 python tools/trace.py llama-3-8b 64 > gpurun_out/decode_timeline_r2_final_graph.txt 2>&1; tail -11 gpurun_out/decode_timeline_r2_final_graph.txt
  echo "== 5. FP16: Llama-3.2-1B (config 1)"
  timeout 600 python bench.py --workload llama-3.2-1b --quant f16 --no-pp --no-cpu > gpurun_out/bench_r2_final_1b_f16.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_1b_f16.json 1b-f16
  echo "== 6. ncu launch list of the bench command (time + DRAM bytes per launch; decode kernels only)"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_stream_matvec|k_rmsnorm|k_attention|k_argmax" -c 1135 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_bench.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_launches_default.csv 681 2>&1 | tee gpurun_out/r2_launches_default.summary.txt
} 2>&1 | tee $L
