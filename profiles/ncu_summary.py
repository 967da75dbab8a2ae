#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'lts__t_bytes.sum','l1tex__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.sum',
 'sm__inst_issued.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct',
 'sm__inst_executed_pipe_fp64.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum',
 'sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_tex.sum',
 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct',
 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct',
 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct',
 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_imc_miss_per_warp_active.pct',
 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_sleeping_per_warp_active.pct',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active', 'local_load_sectors', 'smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('--- launch', r[hdr.index('ID')], r[hdr.index('Kernel Name')][:40])
    for w in WANT:
        if w in hdr:
            print('   %-78s %16s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
if len(sys.argv) > 2:
    for h in hdr:
        if sys.argv[2] in h: print(h)
