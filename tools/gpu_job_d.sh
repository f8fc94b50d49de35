#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_cov.py tests/test_gpu_estep.py tests/test_golden.py tests/test_gpu_families.py -m gpu -x -q > gpurun_out/e_pytest.log 2>&1 ) 2>> gpurun_out/e_times.txt
echo "pytest rc=$?" >> gpurun_out/e_times.txt
( time timeout 900 env GMB_GEMM_TMA=2 python -m pytest tests/test_gpu_cov.py tests/test_gpu_estep.py tests/test_gpu_hmc.py -m gpu -x -q > gpurun_out/e_pytest_tma.log 2>&1 ) 2>> gpurun_out/e_times.txt
echo "pytest tma rc=$?" >> gpurun_out/e_times.txt
python tools/prof_chol.py 5000 10000 > gpurun_out/e_chol5000.txt 2>&1
python tools/prof_chol.py 10000 250 > gpurun_out/e_chol10000.txt 2>&1
( time timeout 600 python tools/bench_configs.py C3 C4 C5 > gpurun_out/e_configs.jsonl 2> gpurun_out/e_configs.err ) 2>> gpurun_out/e_times.txt
python tools/estep_bw.py > gpurun_out/e_estep_bw.txt 2>&1
tail -5 gpurun_out/e_pytest.log; tail -5 gpurun_out/e_pytest_tma.log; cat gpurun_out/e_chol*.txt; cat gpurun_out/e_estep_bw.txt; cat gpurun_out/e_times.txt; tail -3 gpurun_out/e_configs.err
