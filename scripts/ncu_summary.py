#!/usr/bin/env python
"""Condense an .ncu-rep capture (ncu --set full) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "free-text header" > profiles/rNN_x.ncu.txt

Needs the `ncu` CLI (no GPU): reads the raw page as CSV and prints duration, DRAM bytes, pipe and
issue utilisation, the stall breakdown (warps stalled per issued instruction) and launch shape.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory pipe % of peak'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared bank conflicts'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared wavefronts'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
]


def main():
    rep = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('# %s' % header)
    print('# source: %s (ncu --set full --clock-control none), condensed by scripts/ncu_summary.py' % rep)
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print('\nkernel: %s' % d.get('Kernel Name', ('?', ''))[0])
        for k, label in KEYS:
            if k in d:
                print('  %-34s %s %s' % (label, d[k][0], d[k][1]))
        stalls = sorted(((float(v[0]), h) for h, v in d.items()
                         if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')
                         and v[0] not in ('', 'n/a')), reverse=True)
        print('  stalls (warps stalled per issued instruction):')
        for val, h in stalls[:8]:
            print('    %-22s %.2f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], val))


if __name__ == '__main__':
    main()
