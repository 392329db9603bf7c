"""Brief per-kernel metrics + top stall hot spots from an .ncu-rep (reads via `ncu -i`)."""
import csv, subprocess, sys
rep = sys.argv[1]; which = sys.argv[2] if len(sys.argv) > 2 else None; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 18
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
stall = [c for c in h if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    name = r[h.index('Kernel Name')].split('(')[0]
    if which and which not in name: continue
    print('---', name, 'id', r[h.index('ID')])
    for w in want:
        if w in h: print(f'   {w} [{units[h.index(w)]}] = {r[h.index(w)]}')
    st = sorted(((float(r[h.index(c)] or 0), c.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for c in stall), reverse=True)[:6]
    print('   stalls/issue:', ', '.join(f'{n}={v:.2f}' for v, n in st))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
seen = set()
for k, i0 in enumerate(hi):
    name = rows[i0 - 1][1].split('(')[0] if rows[i0 - 1] and rows[i0 - 1][0] == 'Kernel Name' else '?'
    if (which and which not in name) or name in seen: continue
    seen.add(name)
    hh = rows[i0]; end = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    body = [r for r in rows[i0 + 1:end] if len(r) == len(hh)]
    S = hh.index('# Samples'); SRC = hh.index('Source'); IE = hh.index('Instructions Executed')
    sc = [i for i, c in enumerate(hh) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[S]) for r in body) or 1
    print(f'== {name}: {tot} samples, {len(body)} SASS instrs, {sum(int(r[IE]) for r in body)} warp-instr executed')
    for r in sorted(body, key=lambda r: -int(r[S]))[:topn]:
        st = sorted(((int(r[i]), hh[i]) for i in sc), reverse=True)[:2]
        print(f'{int(r[S]):6d} {100*int(r[S])/tot:5.1f}% exec={r[IE]:>9} {r[SRC].strip()[:58]:58s} {st}')
