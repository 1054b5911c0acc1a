#!/bin/bash
# First GPU call of round 2 (one `gpurun --timeout 900 -- 'bash tools/r2_first_run.sh'`): validates and times the
# round-2 RMSNorm accumulator (tools/seqsum2/README.md) and records the new decode timeline.  Writes to gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== 1. stand-alone harness: v2 vs the literal loop, v1 vs v2 timing"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I gpullama3.java_b200/csrc -o /tmp/seqsum2_harness tools/seqsum2/harness.cu \
    && for n in 4096 2048 8192 2560; do timeout 120 /tmp/seqsum2_harness $n 1500; done
  echo "== 2. library built with the v2 accumulator: bit-exact decode tests + adversarial sums"
  B200_NVCC_DEFINES=B200_SEQSUM_V2 python -c "import __graft_entry__ as g; g.build()" \
    && B200_NVCC_DEFINES=B200_SEQSUM_V2 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
  echo "== 3. decode timeline with v2 (compare with profiles/decode_timeline_r1.txt: rmsnorm 17.6 us x 65)"
  B200_NVCC_DEFINES=B200_SEQSUM_V2 timeout 200 python tools/trace.py llama-3-8b 64 2>&1 | tail -12
  echo "== 4. persistent decode kernel (experimental/decode_persistent.cuh): smallest bit-exact test first, under a short timeout"
  export B200_NVCC_DEFINES="B200_SEQSUM_V2 B200_PERSISTENT_DECODE"
  python -c "import __graft_entry__ as g; g.build()" \
    && timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "decode_q8_bit_exact" 2>&1 | tail -5 \
    && timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 \
    && timeout 200 python bench.py --no-cpu --no-pp --steps 64 2>&1 | tail -1 | cut -c1-400
  unset B200_NVCC_DEFINES
  echo "== 5. persistent CTA-pair GEMM (prefill_gemm.cuh, k_gemm_f16_2cta_persist): stand-alone check, then the prefill tests and pp512 with it"
  python -c "import __graft_entry__ as g; g.build()"
  B200_GEMM_2CTA=1256 timeout 90 python tools/gemm_check.py --big 2>&1 | tail -7
  B200_GEMM_2CTA=1256 B200_GEMM_RESID=1 timeout 60 python tools/gemm_check.py 2>&1 | tail -2
  B200_GEMM_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_prefill.py -x -q 2>&1 | tail -3
  B200_GEMM_PERSIST=1 timeout 200 python tools/pp_bench.py llama-3-8b 512 5 2>&1 | tail -1
  echo "== 6. back to the default build"
  python -c "import __graft_entry__ as g; g.build()"
} 2>&1 | tee gpurun_out/r2_first_run.log
