"""Summarise an .ncu-rep: key raw metrics of the first kernel + top stall lines of the source page.
usage: python tools/ncu_summary.py REPORT [N_LINES]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in d: print(f"| {w} | {d[w][0]} | {d[w][1]} |")
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
        v = float(d[h][1])
        if v >= 0.3: print(f"| {h} | {d[h][0]} | {v:.3f} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if any('Source' == c for c in r))
h = rows[hi]
def col(name):
    for i, c in enumerate(h):
        if c == name: return i
    return None
ci = {n: col(n) for n in ['#', 'Source', 'Warp Stall Sampling (All Samples)', 'Instructions Executed', '# Samples']}
sc = ci['Warp Stall Sampling (All Samples)'] or ci['# Samples']
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(float(r[sc] or 0) for r in body)
body.sort(key=lambda r: -float(r[sc] or 0))
print(f"\nTop {nlines} source lines by warp stall samples (total {tot:.0f}):")
for r in body[:nlines]:
    print(f"{float(r[sc]) / tot * 100:6.2f}%  L{r[ci['#']]:>4}  {r[ci['Source']].strip()[:150]}")
