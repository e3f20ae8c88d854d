"""Summarise an .ncu-rep (first kernel) into the handful of numbers the design decisions rest on."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct']


def main(path, px=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, u, v = r[0], r[1], r[-1]
    d = dict(zip(h, v))
    units = dict(zip(h, u))
    print(f"# {path}: {d.get('Kernel Name', '')[:100]}")
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {d[k]:>18s} {units[k]}")
    stalls = sorted(((float(d[k]), k) for k in h if 'pcsamp_warps_issue_stalled' in k and not k.endswith('_not_issued') and d[k]),
                    reverse=True)
    tot = sum(s for s, _ in stalls) or 1
    print("warp-state samples: " + ", ".join(f"{k.split('stalled_')[1]} {100 * s / tot:.1f}%" for s, k in stalls[:9]))
    if px:
        print(f"warp-instructions per 32-px row: {float(d['smsp__inst_executed.sum']) / (px / 32):.1f}")


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
