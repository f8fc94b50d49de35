#!/bin/bash
# round-2 GPU job B: new kernels (Cholesky / TRSM look-ahead, family codes, Wendland) + parity suites
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_cov.py tests/test_gpu_families.py tests/test_rcpp_adapters.py tests/test_gpu_entry_points.py -m gpu -x -q > gpurun_out/b_pytest1.log 2>&1 ) 2>> gpurun_out/b_times.txt
echo "pytest1 rc=$?" >> gpurun_out/b_times.txt
( time timeout 600 python tools/bench_configs.py C3 C5 > gpurun_out/b_configs.jsonl 2> gpurun_out/b_configs.err ) 2>> gpurun_out/b_times.txt
echo "configs rc=$?" >> gpurun_out/b_times.txt
( time timeout 1500 python -m pytest tests/test_gpu_fit_parity.py -m gpu -x -q -s > gpurun_out/b_pytest2.log 2>&1 ) 2>> gpurun_out/b_times.txt
echo "pytest2 rc=$?" >> gpurun_out/b_times.txt
tail -5 gpurun_out/b_pytest1.log; tail -3 gpurun_out/b_configs.err; tail -8 gpurun_out/b_pytest2.log; cat gpurun_out/b_times.txt
