#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): plain run first, then the launch list of the bench command and one `--set full` capture per kernel family.
# Reports are summarised on the box (they exceed gpurun's return limit): gpurun_out/profiles_out/*.txt -> copied to profiles/ afterwards.
set -u
mkdir -p gpurun_out/profiles_out
P=gpurun_out/profiles_out
if [ -z "${SKIP_LAUNCH_LIST:-}" ]; then
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/p_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/p_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/p_ncu_bench.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/p_launches_bench.csv', errors='replace')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum': continue
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
    k = r[idx['Kernel Name']][:110]; t[k] += v; n[k] += 1
tot = sum(t.values())
with open('gpurun_out/profiles_out/r02_launches_bench.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs` (whole process: 4 steps, the e2e calls and the roofline probes); per kernel: launches, total us, share\n')
    f.write('kernel,launches,total_us,avg_us,share\n')
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        f.write('"%s",%d,%.1f,%.2f,%.4f\n' % (k, n[k], v, v / n[k], v / tot))
PY
fi
SPECS=${SPECS:-"hmc:hmc_lane:2:hmc_hmc_lane estep:loglik_logit_factor_kernel|mcnr_tma_kernel|mcnr_tail:12:estep_loglik_logit estep:loglik_logit_agg_kernel|mcnr_agg_kernel|agg_finish:6:estep_loglik_logit_agg gemm:dgemm_tma_kernel|sgemm3_tf32_kernel|split_tf32:6:gemm_dgemm_tma_ke chol:potrf_diag_kernel:1:chol_potrf_diag_k chol:dgemm_tma_kernel:4:chol_dgemm_tma_ke"}
for spec in $SPECS; do
    IFS=: read what rx cnt tag <<< "$spec"
    python tools/prof_r02.py $what > gpurun_out/p_plain_$what.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/p_$tag -f python tools/prof_r02.py $what > gpurun_out/p_ncu_$tag.log 2>&1
    echo "$tag full rc=$?"
    python tools/ncu_extract.py gpurun_out/p_$tag.ncu-rep "python tools/prof_r02.py $what — ncu --set full --clock-control none -k regex:$rx -c $cnt" > $P/r02_ncu_$tag.txt 2>&1
    ncu -i gpurun_out/p_$tag.ncu-rep --page source --csv 2>/dev/null | head -400 > $P/r02_ncu_source_$tag.csv
    rm -f gpurun_out/p_$tag.ncu-rep
done
rm -f gpurun_out/p_launches_bench.csv
ls -la $P
