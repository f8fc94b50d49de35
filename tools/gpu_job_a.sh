#!/bin/bash
# round-2 GPU job A: large configs at full size, the GPU test suite, the bench line and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
( time timeout 900 python tools/bench_configs.py C3 C4 C5 > gpurun_out/a_configs.jsonl 2> gpurun_out/a_configs.err ) 2>> gpurun_out/a_times.txt
echo "configs rc=$?" >> gpurun_out/a_times.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1 ) 2>> gpurun_out/a_times.txt
echo "pytest rc=$?" >> gpurun_out/a_times.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err ) 2>> gpurun_out/a_times.txt
echo "bench rc=$?" >> gpurun_out/a_times.txt
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err ) 2>> gpurun_out/a_times.txt
echo "ref rc=$?" >> gpurun_out/a_times.txt
tail -5 gpurun_out/a_pytest.log; tail -3 gpurun_out/a_configs.err; cat gpurun_out/a_times.txt
