#!/bin/bash
# FSFB_GEMM_DEBUG sweep of the SS gather-GEMM (1 no A loads, 2 no W copies, 4 no MMAs, 8 no A stores, 16 no proxy fence)
mkdir -p gpurun_out
for d in 0 1 2 4 3 5 6 7; do
  echo "=== FSFB_GEMM_DEBUG=$d"
  FSFB_GEMM_DEBUG=$d timeout 120 python tools/gemm_role_timers.py 2>&1 | grep -A13 "conv sorted" | grep -v "^--"
done > gpurun_out/ss_debug_sweep.txt 2>&1
cat gpurun_out/ss_debug_sweep.txt
