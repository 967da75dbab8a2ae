#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box):  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source]

Per launch: the metrics the roofline discussion in DESIGN.md quotes -- duration, occupancy, issue-slot use, lanes per
instruction, DRAM / L2 / L1 traffic and hit rates, the L1 data-pipe split (LSU vs texture wavefronts), local-memory
(stack) behaviour, and the full warp-stall breakdown as stall cycles per issued instruction.  --source adds the
per-SASS-instruction stall hot spots of the first launch (top 25) so the numbers can be traced to code.
"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
    'smsp__warps_eligible.avg.per_cycle_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    # L1 data pipes: half of every node record goes through the texture pipe, the other half, triangles and path records through the LSU
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
    'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
    'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'l1tex__t_sector_pipe_tex_hit_rate.pct',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
    'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_tex.sum', 'sm__inst_executed_pipe_fp64.sum',
]
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'lg_throttle', 'tex_throttle', 'mio_throttle',
          'no_instruction', 'branch_resolving', 'not_selected', 'selected', 'barrier', 'dispatch_stall', 'imc_miss',
          'drain', 'membar', 'sleeping', 'misc']


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print('--- launch', r[col['ID']], r[col['Kernel Name']][:60])
        for w in WANT:
            if w in col:
                print('   %-78s %16s %s' % (w, r[col[w]], units[col[w]]))
        # stall cycles per issued instruction (ncu "Warp State Statistics")
        tot = 0.0
        parts = []
        for s in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if k in col and num(r[col[k]]) is not None:
                parts.append((num(r[col[k]]), s))
                tot += num(r[col[k]])
        if parts:
            print('   warp stall cycles per issued instruction: total %.2f' % tot)
            for v, s in sorted(parts, reverse=True):
                if v >= 0.005 * tot:
                    print('      %-22s %6.2f  (%4.1f %%)' % (s, v, 100 * v / tot))
        sec, req = 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'smsp__inst_executed_op_local_ld.sum'
        if sec in col and req in col and num(r[col[req]]):
            # a stack pop reads 4 bytes per active lane; a 32-byte sector fetched for it carries bytes_used/32 useful data
            print('   local (stack) loads: %.2f sectors per instruction' % (num(r[col[sec]]) / num(r[col[req]])))
    if '--source' in sys.argv:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', ':::1'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h = rows[1]
        ix = {n: i for i, n in enumerate(h)}
        data = [r for r in rows[2:] if r and r[0].startswith('0x') and len(r) >= len(h) - 2]
        data = data[:len(data) // 2] if len(data) > 1 and data[0][1] == data[len(data) // 2][1] else data
        tot = sum(int(r[ix['# Samples']]) for r in data) or 1
        stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
        print('--- source hot spots of the first launch (share of warp-stall samples, SASS, top stall reasons)')
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:25]:
            top = sorted(((int(r[ix[x]]), x) for x in stalls), reverse=True)[:2]
            print('   %5.2f %%  lanes %-4s %-64s %s' % (100.0 * int(r[ix['# Samples']]) / tot, r[ix['Avg. Threads Executed']],
                                                        r[ix['Source']].strip()[:64], ' '.join('%s=%d' % (b[6:], a) for a, b in top if a)))


if __name__ == '__main__':
    main()
