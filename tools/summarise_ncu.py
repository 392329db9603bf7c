"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

  python tools/summarise_ncu.py <tag> <rep.ncu-rep> [launches.csv] [kernel-regex-for-traffic]
"""
import csv, json, os, re, subprocess, sys, collections
tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != '-' else None
traffic_kernel = sys.argv[4] if len(sys.argv) > 4 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, 'profiles')
os.makedirs(out, exist_ok=True)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
KEEP = re.compile(r'^(dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__time_duration\.sum|'
                  r'lts__t_sectors_srcunit_tex_op_red\.sum|lts__t_sectors_srcunit_tex_op_red\.avg\.pct_of_peak_sustained_elapsed|'
                  r'lts__t_sectors_srcunit_tex_op_atom\.sum|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'dram__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|'
                  r'launch__shared_mem_per_block_static|launch__shared_mem_per_block_dynamic|smsp__inst_executed\.sum|'
                  r'sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                  r'smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|barrier|mio_throttle|lg_throttle|wait|not_selected|math_pipe_throttle)_per_issue_active\.ratio|'
                  r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__inst_executed_op_shared_atom\.sum|'
                  r'l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active)$')
summ = []
for r in rows[2:]:
    d = {'kernel': r[h.index('Kernel Name')].split('(')[0], 'id': r[h.index('ID')]}
    for i, c in enumerate(h):
        if KEEP.match(c):
            d[f'{c} [{units[i]}]'] = r[i]
    summ.append(d)
json.dump(summ, open(os.path.join(out, f'{tag}_ncu_full_summary.json'), 'w'), indent=1)
# traffic of the dominant kernel (per launch, average)
def mb(d, key):
    for k, v in d.items():
        if k.startswith(key):
            f = float(v.replace(',', ''))
            unit = k[k.index('[') + 1:-1]
            return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return 0.0
sel = [d for d in summ if (traffic_kernel is not None and traffic_kernel in d['kernel'])]
tr = [mb(d, 'dram__bytes_read.sum') + mb(d, 'dram__bytes_write.sum') for d in sel]
if sel:
  json.dump({'kernel': sel[0]['kernel'], 'dram_bytes_per_launch': sum(tr) / len(tr), 'launches_captured': len(tr),
           'source': f'profiles/{tag}_ncu_full_summary.json'}, open(os.path.join(out, 'ncu_traffic.json'), 'w'), indent=1)   # only when the traffic kernel was captured
# stall hot spots
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
open('/tmp/_src.csv', 'w').write(src)
top = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_top.py'), '/tmp/_src.csv', '25'], capture_output=True, text=True).stdout
open(os.path.join(out, f'{tag}_ncu_stall_hotspots.txt'), 'w').write(top)
if launches:
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hh = rows[hi]; kn = hh.index('Kernel Name'); mv = hh.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            agg.setdefault(r[kn].split('(')[0], []).append(float(r[mv].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out, f'{tag}_ncu_launches.csv'), 'w') as f:
        f.write('kernel,launches,total_ns,avg_ns,share_of_all\n')
        for k, v in agg.items():
            f.write(f'"{k}",{len(v)},{sum(v):.0f},{sum(v)/len(v):.0f},{sum(v)/tot:.4f}\n')
print(open(os.path.join(out, 'ncu_traffic.json')).read())
