#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 env GMB_GEMM_TMA=2 python -m pytest tests/test_gpu_cov.py tests/test_gpu_estep.py -m gpu -x -q > gpurun_out/c_pytest_tma.log 2>&1 ) 2>> gpurun_out/c_times.txt
echo "tma forced rc=$?" >> gpurun_out/c_times.txt
( time timeout 900 python -m pytest tests/test_gpu_gemm_tma.py tests/test_gpu_cov.py -m gpu -x -q > gpurun_out/c_pytest.log 2>&1 ) 2>> gpurun_out/c_times.txt
echo "pytest rc=$?" >> gpurun_out/c_times.txt
python tools/prof_chol.py 5000 10000 > gpurun_out/c_chol5000.txt 2>&1
python tools/prof_chol.py 10000 250 > gpurun_out/c_chol10000.txt 2>&1
GMB_GEMM_TMA=0 python tools/prof_chol.py 5000 10000 > gpurun_out/c_chol5000_notma.txt 2>&1
( time timeout 600 python tools/bench_configs.py C3 C5 > gpurun_out/c_configs.jsonl 2> gpurun_out/c_configs.err ) 2>> gpurun_out/c_times.txt
tail -5 gpurun_out/c_pytest_tma.log; tail -5 gpurun_out/c_pytest.log; cat gpurun_out/c_chol*.txt; cat gpurun_out/c_times.txt
