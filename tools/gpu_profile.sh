#!/bin/bash
# Run on the GPU box through gpurun: plain runs first, then the ncu passes (B200_PROFILING.md recipe).
set -u
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/profile_kernels.py all > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'loglik_kernel|mcnr_pass1|hmc_fused' -c 6 \
    -o gpurun_out/prof_r01 -f python tools/profile_kernels.py all > gpurun_out/ncu_prof.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/plain_prof.log gpurun_out/ncu_prof.log
