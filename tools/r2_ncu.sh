#!/bin/bash
# ncu --set full of ONE launch of the persistent decode kernel (source-level stall sampling) + the launch list of a short bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_persistent -s 10 -c 1 -o gpurun_out/r2_pd_full -f \
      python bench.py --no-pp --no-cpu --steps 8 --warmup 3 --decode-mode persistent 2>&1 | tail -5
  ls -la gpurun_out/r2_pd_full.ncu-rep
} 2>&1 | tee gpurun_out/r2_ncu.log
