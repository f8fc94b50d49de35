#!/bin/bash
set -u
mkdir -p gpurun_out
W=${1:-all}
python tools/profile_kernels.py $W > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'loglik_kernel|mcnr_pass1|hmc_fused' -c 6 \
    -o gpurun_out/prof_r01b -f python tools/profile_kernels.py $W > gpurun_out/ncu_prof.log 2>&1
echo "full rc=$?"
tail -n 3 gpurun_out/plain_prof.log
