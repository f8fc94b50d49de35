#!/bin/bash
# Turns the ncu captures of tools/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/.
#   bash tools/make_profiles.sh r01
set -u
TAG=${1:-r01}
OUT=profiles
python - "$TAG" <<'PY'
import csv, subprocess, sys
tag = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio']
def block(rep, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ['## ' + title]
    for r in rows[2:]:
        out.append('### ' + r[idx['Kernel Name']][:110])
        for w in WANT:
            if w in idx: out.append('  %-85s %s %s' % (w, r[idx[w]], units[idx[w]]))
        out.append('')
    return '\n'.join(out)
txt = ['# %s (final kernels of the round) — extracts of `ncu --set full --clock-control none --import-source on` captures (tools/gpu_profile.sh, tools/make_profiles.sh).' % tag,
       '# Units as printed by `ncu --page raw --csv`; one block per captured launch.', '']
txt.append(block('gpurun_out/prof_hmc_%s.ncu-rep' % tag, 'sampler, C2 model, 62 proposals (tools/profile_hmc.py): launch 1 = the structure-aware kernel as the bench runs it (1000 chains, 50 distinct rows, '
                 '150 non-zeros of Z L, one warp per chain); launch 2 = the dense on-chip kernel, forced, 1184 chains without row aggregation (500 rows, DMMA)'))
txt.append(block('gpurun_out/prof_estep_%s.ncu-rep' % tag, 'E-step kernels on 1 GB (C2 model, m = 250000; tools/profile_kernels.py estep): log-likelihood and MCNR kernels'))
open('profiles/%s_ncu_raw_extract_final.txt' % tag, 'w').write('\n'.join(txt) + '\n')
PY
for spec in "hmc:hmc_sparse:hmc_sparse_final" "hmc:hmc_fused:hmc_fused_final" "estep:loglik:loglik_final" "estep:mcnr_pass1:mcnr_final"; do
    IFS=: read rep rx name <<< "$spec"
    ncu -i gpurun_out/prof_${rep}_${TAG}.ncu-rep --page source --csv --kernel-name regex:$rx 2>/dev/null > /tmp/src_$name.csv
    python tools/ncu_source_summary.py /tmp/src_$name.csv 40 > $OUT/${TAG}_ncu_source_$name.txt 2>&1
done
cp gpurun_out/launches_bench_${TAG}.csv /tmp/launches_full.csv
python - "$TAG" <<'PY'
import csv, collections, sys
tag = sys.argv[1]
rows = list(csv.reader(open('/tmp/launches_full.csv', errors='replace')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum': continue
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}.get(u, 1.0)
    k = r[idx['Kernel Name']][:100]; t[k] += v; n[k] += 1
tot = sum(t.values())
with open('profiles/%s_launches_bench.csv' % tag, 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (whole process: 4 steps, the e2e calls and the roofline probes); per kernel: launches, total us, share\n')
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        f.write('"%s",%d,%.1f,%.2f,%.4f\n' % (k, n[k], v, v / n[k], v / tot))
PY
head -12 profiles/${TAG}_launches_bench.csv
