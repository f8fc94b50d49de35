#!/bin/bash
# Run on the GPU box through gpurun: plain runs first, then the ncu passes (B200_PROFILING.md recipe).
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r01'
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/profile_hmc.py > gpurun_out/plain_hmc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"hmc_fused|hmc_sparse" -c 2 -o gpurun_out/prof_hmc_${TAG} -f \
    python tools/profile_hmc.py > gpurun_out/ncu_hmc.log 2>&1
echo "hmc full rc=$?"
python tools/profile_kernels.py estep > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'loglik|mcnr_pass1' -c 6 -o gpurun_out/prof_estep_${TAG} -f \
    python tools/profile_kernels.py estep > gpurun_out/ncu_prof.log 2>&1
echo "estep full rc=$?"
tail -n 2 gpurun_out/plain_hmc.log gpurun_out/plain_prof.log
# summaries are extracted here, on the box (the reports exceed gpurun's 64 MiB return limit): profiles/ -> gpurun_out/profiles_out/
bash tools/make_profiles.sh ${TAG} > gpurun_out/make_profiles.log 2>&1
mkdir -p gpurun_out/profiles_out && cp profiles/${TAG}_ncu_raw_extract_final.txt profiles/${TAG}_ncu_source_*_final.txt profiles/${TAG}_launches_bench.csv gpurun_out/profiles_out/
rm -f gpurun_out/prof_*_${TAG}.ncu-rep gpurun_out/launches_bench_${TAG}.csv
