"""Print the headline metrics of an ncu report (raw page). usage: python tools/ncu_raw.py rep.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit',
        'launch__waves_per_multiprocessor', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'launch__grid_size', 'launch__block_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warp', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__pcsamp_warps_issue_stalled']


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('---', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(w) for w in WANT) and 'per_second' not in h and 'pct_of_peak_sustained_elapsed' not in h.replace('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '').replace('sm__throughput.avg.pct_of_peak_sustained_elapsed', ''):
                print('  %-75s %-12s %s' % (h, u, v))


if __name__ == "__main__":
    main()
