#!/usr/bin/env python
"""Prints the metrics we track from an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'memory_l1_wavefronts_shared_ideal',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('==', name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k} = {r[i]} {units[i]}')
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                try:
                    st.append((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
