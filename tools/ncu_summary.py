#!/usr/bin/env python
"""Key counters per kernel launch of an ncu report (raw page): usage ncu_summary.py report.ncu-rep [out.csv]"""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall = [i for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    table = []
    for r in data:
        d = {'kernel': r[hdr.index('Kernel Name')][:70]}
        for w in WANT:
            if w in hdr: d[w] = r[hdr.index(w)] + ' ' + units[hdr.index(w)]
        top = sorted(((float(r[i].replace(',', '') or 0), hdr[i][len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for i in stall), reverse=True)[:6]
        d['stalls'] = ' '.join(f"{n}={v:.2f}" for v, n in top)
        table.append(d)
    for d in table:
        print('---')
        for k, v in d.items(): print(f"  {k}: {v}")
    if len(sys.argv) > 2:
        keys = list(table[0].keys())
        with open(sys.argv[2], 'w', newline='') as f:
            w = csv.writer(f); w.writerow(keys)
            for d in table: w.writerow([d.get(k, '') for k in keys])
if __name__ == "__main__":
    main()
