#!/bin/bash
# Round-2 validation on one B200: full GPU test suite, smoke, the bench (both arms), supporting measurements -> gpurun_out/profiles_out/
set -u
P=gpurun_out/profiles_out; mkdir -p $P
( time timeout 1500 python -m pytest tests/ -m gpu -q 2>&1 | tail -6 ) > gpurun_out/j_pytest.log 2>&1
tail -8 gpurun_out/j_pytest.log
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $P/r02_bench_final.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/j_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $P/r02_bench_reference_arm.json 2> gpurun_out/j_ref.err; echo "ref rc=$?"
python tools/tf32_error_sweep.py > $P/r02_tf32_error_sweep.txt 2>&1; tail -4 $P/r02_tf32_error_sweep.txt
{ echo "# python tools/prof_chol.py N M: factorisation of one N x N fexp block; mvn_ll with M device-resident sample columns (CUDA events)";
  for a in "3000 0" "5000 10000" "10000 10000"; do echo "## N M = $a"; python tools/prof_chol.py $a 2>&1 | tail -3; done; } > $P/r02_chol_timings.txt
cat $P/r02_chol_timings.txt
python tools/bench_mcml_full.py > $P/r02_mcml_full_fits.jsonl 2> gpurun_out/j_fits.err; tail -3 $P/r02_mcml_full_fits.jsonl | cut -c1-400
