#!/bin/bash
# Cholesky restructure: parity tests, then timings
timeout 900 python -m pytest tests/test_gpu_cov.py tests/test_gpu_gemm.py tests/test_gpu_gemm_tma.py -x -q 2>&1 | tail -5
for r in 8 0 16; do
  echo "reserve $r"; GMB_CHOL_RESERVE_SMS=$r timeout 300 python tools/prof_chol.py 5000 2>&1 | tail -1
  GMB_CHOL_RESERVE_SMS=$r timeout 300 python tools/prof_chol.py 10000 2>&1 | tail -1
done
timeout 300 python tools/prof_gemm.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "C5 or C3" 2>&1 | tail -3
