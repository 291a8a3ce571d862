"""Small helper used during development: print selected metrics of an `ncu --page raw --csv` dump."""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct']


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('----')
        for w in list(WANT) + list(extra):
            if w in idx:
                print(f"{w:80s} {r[idx[w]]} {units[idx[w]]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
