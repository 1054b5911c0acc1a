#!/bin/bash
# Round-2 GPU call 1: seqsum2 + norm v2 on the graph path, the persistent decode kernel under the bit-exact suite, first timings.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_run1.log
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== 1. exact accumulators (v1, v2@1024, v2@256)"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "sequential_sum" 2>&1 | tail -4
  echo "== 2. everything that is not the persistent kernel (graph path with norm v2, FP16, long context on graph...)"
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "not persistent and not sequential_sum and not interleave and not long_context" 2>&1 | tail -12
  echo "== 3. persistent kernel: one process per test, short timeouts"
  for t in "test_decode_q8_bit_exact[tiny-llama-persistent]" "test_decode_q8_bit_exact[tiny-llama-tied-persistent]" "test_decode_q8_bit_exact[tiny-qwen3-persistent]" \
           "test_decode_small_llama_q8[persistent]" "test_decode_mid_geometries_q8[mid-llama-persistent]" "test_decode_mid_geometries_q8[mid-qwen3-4b-persistent]" \
           "test_decode_mid_geometries_q8[mid-llama-1b-persistent]" "test_decode_mid_geometries_q8[mid-llama-70b-persistent]" "test_prefill_graph_then_decode[persistent]" \
           "test_batch_prefill_matches_oracle[persistent]" "test_decode_sequence_device_loop[persistent]" "test_generation_loops_match_oracle[persistent]" \
           "test_kv_reset_and_determinism[persistent]" "test_modes_interleave" "test_long_context_score_row_in_global_memory"; do
    echo "-- $t"
    timeout 150 python -m pytest "tests/test_gpu_parity.py::$t" -x -q 2>&1 | tail -6
  done
  echo "== 5. bench: graph, persistent"
  timeout 300 python bench.py --no-pp --no-cpu --decode-mode graph > gpurun_out/bench_r2_graph.json 2> gpurun_out/bench_r2_graph.err; tail -c 1500 gpurun_out/bench_r2_graph.json; tail -3 gpurun_out/bench_r2_graph.err
  timeout 400 python bench.py --no-pp --decode-mode persistent > gpurun_out/bench_r2_persistent.json 2> gpurun_out/bench_r2_persistent.err; tail -c 2500 gpurun_out/bench_r2_persistent.json; tail -3 gpurun_out/bench_r2_persistent.err
  for a in 8 32; do
    echo "-- persistent, L2 look-ahead $a tiles"
    B200_PD_L2_AHEAD=$a timeout 200 python bench.py --no-pp --no-cpu --decode-mode persistent 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['frac'])"
  done
  echo "== 6. timelines"
  B200_DECODE=graph timeout 200 python tools/trace.py llama-3-8b 64 > gpurun_out/decode_timeline_r2_graph.txt 2>&1; tail -14 gpurun_out/decode_timeline_r2_graph.txt
  timeout 200 python tools/trace_persistent.py llama-3-8b 64 > gpurun_out/decode_timeline_r2_persistent.txt 2>&1; tail -30 gpurun_out/decode_timeline_r2_persistent.txt
  echo "== 7. persistent CTA-pair GEMM (prefill): stand-alone check, prefill tests, pp512"
  B200_GEMM_2CTA=1256 timeout 90 python tools/gemm_check.py --big 2>&1 | tail -7
  B200_GEMM_2CTA=1256 B200_GEMM_RESID=1 timeout 60 python tools/gemm_check.py 2>&1 | tail -2
  B200_GEMM_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_prefill.py -x -q 2>&1 | tail -3
  B200_GEMM_PERSIST=1 timeout 200 python tools/pp_bench.py llama-3-8b 512 5 2>&1 | tail -1
  timeout 200 python tools/pp_bench.py llama-3-8b 512 5 2>&1 | tail -1
} 2>&1 | tee $L
