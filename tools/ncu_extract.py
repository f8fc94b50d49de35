"""Selected metrics of every launch in an .ncu-rep (ncu --page raw --csv), as a text block for profiles/."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print('## ' + title)
for r in rows[2:]:
    print('### ' + r[idx['Kernel Name']][:120])
    for w in WANT:
        if w in idx and r[idx[w]] != '':
            print('  %-90s %s %s' % (w, r[idx[w]], units[idx[w]]))
    print()
