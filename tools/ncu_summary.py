"""Key metrics of one kernel from an `ncu --set full` report, as text for profiles/.
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv, sys
KEEP = ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'sm__inst_executed_pipe_uniform.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__cycles_elapsed.avg',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_reads.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_writes.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second')
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
hdr, units = rows[0], rows[1]
for rec in rows[2:]:
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            print('%s [%s] = %s' % (name, units[i], rec[i]))
    print()
