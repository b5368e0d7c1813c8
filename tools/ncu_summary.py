#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics the
roofline discussion uses.  Usage: python tools/ncu_summary.py <report.ncu-rep> [> profiles/x.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'sm__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_sleeping_per_warp_active.pct',
        'smsp__warp_issue_stalled_membar_per_warp_active.pct']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')], 'grid', r[hdr.index('Grid Size')],
              'block', r[hdr.index('Block Size')])
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith('.' + k):
                    print('  {:85s} {} {}'.format(h, r[i], units[i]))
                    break
        print()


if __name__ == '__main__':
    main()
