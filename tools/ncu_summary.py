"""Summarise an ncu report exported as CSV (raw page + source page): key metrics, stall mix,
and stall samples per code region.  Usage: ncu_summary.py raw.csv src.csv"""
import csv, sys, collections
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'lts__t_bytes.sum.per_second',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__block_size', 'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_lsu.sum']
for k in keys:
    if k in m: print(f'{k:85s} {m[k]:>18s} {units[hdr.index(k)]}')
print('stall mix (warps per issue-active):')
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
        v = float(m[h])
        if v > 0.05: print(f'   {h[34:-23]:25s} {v:6.2f}')
rows = list(csv.reader(open(src)))
h2 = rows[1]; col = {h: i for i, h in enumerate(h2)}; data = rows[2:]
tot = sum(int(r[col['# Samples']] or 0) for r in data)
stall_cols = [h for h in h2 if h.startswith('stall_') and 'Not' not in h]
print('instructions', len(data), 'samples', tot)
W = 256
for i in range(0, len(data), W):
    blk = data[i:i + W]
    s = sum(int(r[col['# Samples']] or 0) for r in blk)
    if s < 0.004 * tot: continue
    ex = sum(int(r[col['Instructions Executed']] or 0) for r in blk)
    nd = sum(1 for r in blk if 'DMMA' in r[col['Source']])
    st = {h: sum(int(r[col[h]] or 0) for r in blk) for h in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f'{i:6d} {100 * s / tot:5.1f}% exec {ex / 1e6:8.1f}M dmma {nd:3d} ', ' '.join(f'{k[6:]}:{100 * v / tot:.1f}' for k, v in top))
