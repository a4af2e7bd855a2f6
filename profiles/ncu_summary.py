"""Summarise an `ncu --page raw --csv` dump: one block of key metrics per captured kernel launch."""
import csv
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_registers', 'occ limit regs'), ('launch__occupancy_limit_shared_mem', 'occ limit smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'DMMA pipe % (active)'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % (elapsed)'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU pipe %'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / warp inst'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long scoreboard'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short scoreboard'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math pipe throttle'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no instruction'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall branch resolving'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio throttle'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg throttle'),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('## ' + r[idx['Kernel Name']])
        for key, label in KEYS:
            if key in idx:
                print(f'  {label:32s} {r[idx[key]]} {units[idx[key]]}')
        print()


if __name__ == '__main__':
    main(sys.argv[1])
