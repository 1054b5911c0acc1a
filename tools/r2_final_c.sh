#!/bin/bash
# Round-2 last run: the attention kernel of run 9 (faster than the pipelined variant at every depth measured) against the deep-context tests,
# quick bench lines, ncu launch list of the default bench command.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_final_c.log
line() { grep "^{" "$1" | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$2', 'tok/s', round(j['value'], 1), 'ms', round(j['ms_per_step'], 3), 'frac', round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value'], 1), 'mode', j.get('decode_mode'))"; }
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. deep-context + attention-touching parity tests"
  timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "deep_context or decode_q8_bit_exact or phi3_bit_exact or long_context or mid_geometries_q8" 2>&1 | tail -4
  echo "== 2. smoke + default bench (no CPU leg)"
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 300 python bench.py --no-cpu --no-pp > gpurun_out/bench_r2_final_c.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_c.json default
  echo "== 3. depth lines"
  timeout 300 python bench.py --depth 1024 --steps 64 --no-cpu --no-pp > gpurun_out/bench_r2_final_c_d1024.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_c_d1024.json d1024
  timeout 300 python bench.py --depth 4096 --steps 32 --no-cpu --no-pp > gpurun_out/bench_r2_final_c_d4096.json 2> gpurun_out/tmp.err; line gpurun_out/bench_r2_final_c_d4096.json d4096
  echo "== 4. ncu launch list of the bench command"
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_stream_matvec|k_rmsnorm|k_attention|k_argmax" -c 1135 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_bench.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_launches_default.csv 681 2>&1 | tee gpurun_out/r2_launches_default.summary.txt
} 2>&1 | tee $L
