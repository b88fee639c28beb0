#!/usr/bin/env python
"""tools/ncu_summary.py <raw.csv>... -- one-screen summary of `ncu --page raw --csv` exports (the counters DESIGN.md cites)."""
import csv
import json
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.per_cycle_active', 'launch__occupancy_limit_blocks', 'launch__occupancy_limit_warps']


def summarise(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {'kernel': vals[hdr.index('Kernel Name')]}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out[w] = f'{vals[i]} {units[i]}'.strip()
    return out


if __name__ == '__main__':
    res = {p: summarise(p) for p in sys.argv[1:]}
    print(json.dumps(res, indent=1))
