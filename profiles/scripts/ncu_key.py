"""ncu -i X.ncu-rep --page raw --csv | python profiles/scripts/ncu_key.py  -> the metrics the profiles/ summaries quote"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_requests_srcunit_tex_op_read.sum', 'lts__t_requests_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__sectors_read.sum', 'dram__sectors_write.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-80s %s %s" % (w, r[i], units[i]))
    print()
