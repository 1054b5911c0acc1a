#!/bin/bash
# Round-2 GPU call 3: optimisation pass on the persistent kernel (term walk, register-resident norm, batched loads): parity + timings.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run2.log
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. parity file (both decode modes)"
  timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -6
  echo "== 2. bench persistent / graph"
  timeout 400 python bench.py --no-pp --no-cpu --decode-mode persistent > gpurun_out/bench_r2b_persistent.json 2> gpurun_out/bench_r2b_persistent.err; python -c "
import json; j=json.load(open('gpurun_out/bench_r2b_persistent.json')); print('persistent', j['value'], j['ms_per_step'], j['roofline']['frac'], j['e2e']['value'])"; tail -2 gpurun_out/bench_r2b_persistent.err
  timeout 300 python bench.py --no-pp --no-cpu --decode-mode graph > gpurun_out/bench_r2b_graph.json 2> gpurun_out/bench_r2b_graph.err; python -c "
import json; j=json.load(open('gpurun_out/bench_r2b_graph.json')); print('graph', j['value'], j['ms_per_step'], j['roofline']['frac'], j['e2e']['value'], {k: round(v['GB/s']) for k, v in j['roofline']['other_kernels'].items()})"; tail -2 gpurun_out/bench_r2b_graph.err
  echo "== 3. timelines"
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2b_persistent.txt 2>&1; tail -40 gpurun_out/decode_timeline_r2b_persistent.txt
  B200_DECODE=graph timeout 200 python tools/trace.py llama-3-8b 64 > gpurun_out/decode_timeline_r2b_graph.txt 2>&1; tail -12 gpurun_out/decode_timeline_r2b_graph.txt
  echo "== 4. sampler smoke (device-side temperature/top-p vs oracle)"
  timeout 300 python -m pytest tests/test_gpu_sampler.py -x -q 2>&1 | tail -5
} 2>&1 | tee $L
