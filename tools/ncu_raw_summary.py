"""Key metrics per kernel from `ncu --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_xu.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_active' if False else 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:60])
    for w in want:
        if w in d: print(f"    {w:75s} {d[w]}")
    top = sorted(((float(d[s] or 0), s) for s in stalls), reverse=True)[:6]
    for v, s in top: print(f"    stall {s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.2f}")
