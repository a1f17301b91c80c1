"""Summarise an ncu report: headline metrics + per-region instruction/stall shares from the source page (developer tool)."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', 'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active_mem_lg.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print('%-90s %s %s' % (w, data[i], units[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
data = [r for r in rows[2:] if len(r) > max(ia, isrc, isamp) and r[ia].isdigit()]
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print('total warp instr', tot, 'samples', tots)
i = 0
while i < len(data):
    j = i; c = int(data[i][ia])
    while j < len(data) and abs(int(data[j][ia]) - c) <= 0.03 * max(c, 1): j += 1
    n = sum(int(r[ia]) for r in data[i:j]); s = sum(int(r[isamp]) for r in data[i:j])
    ops = {}
    for r in data[i:j]:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    if n > 0.01 * tot or s > 0.02 * tots:
        print('instr %4d-%4d (%3d) exec %9d  instr-share %.3f  stall-sample-share %.3f ' % (i, j, j - i, c, n / tot, s / tots), dict(sorted(ops.items(), key=lambda kv: -kv[1])[:7]))
    i = j
