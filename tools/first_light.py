"""Scratch measurement script for gpurun: scatter probes + integrate timing."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from occgrid_b200 import _native, dual_bot_mapper as M, simulation_tools as st

lib = _native.lib()
dev = torch.device('cuda:0')
def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(it+1)]
    ev[0].record()
    for i in range(it):
        fn(); ev[i+1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i+1]) for i in range(it)]
    return float(np.median(ts)), float(min(ts))

out = {}
stream = torch.cuda.current_stream().cuda_stream
nops = 1 << 26
for name, kind, cells, dtype in [('global_atomicmax_u32_64MiB', 0, 4096*4096, torch.int32),
                                 ('global_atomicmax_u32_2GiB', 0, 8192*65536, torch.int32),
                                 ('global_atomicmax_u32_1.2MiB', 0, 307200, torch.int32),
                                 ('global_store_u8_16MiB', 1, 4096*4096, torch.uint8),
                                 ('smem_atomicmax_16KiB', 2, 4096, torch.int32),
                                 ('smem_atomicmax_64KiB', 2, 16384, torch.int32),
                                 ('smem_store_64KiB', 3, 16384, torch.int32),
                                 ('hot256_red_add', 4, 256, torch.int32), ('hot256_atom_add', 5, 256, torch.int32),
                                 ('hot4096_red_add', 4, 4096, torch.int32), ('hot16_red_add', 4, 16, torch.int32)]:
    plane = torch.zeros(max(cells * 64, 1<<20), dtype=dtype, device=dev)
    def f():
        rc = lib.occgrid_scatter_probe(kind, plane.data_ptr(), cells, nops, 123, stream)
        assert rc == 0, _native.last_error()
    med, best = timeit(f, 2, 5)
    out[name] = {'ms': med, 'Gops_per_s': nops / med / 1e6}
    print(name, out[name], flush=True)
    del plane

npk = int(os.environ.get('NPK', 2_500_000))
s = st.generate_session(n_agents=64, n_packets=npk, seed=42)
for strategy in ('global_atomic', 'tiled'):
    try:
        g = M.OccupancyGrid(strategy=strategy, max_batch=npk, **s['grid'])
    except M.OccGridError as e:
        print(strategy, 'unavailable', e); continue
    pk, _ = g.stage_packets(s['packets'])
    off = torch.from_numpy(s['agent_offsets']).to(dev)
    def f():
        rc = lib.occgrid_integrate_packets(g._geom, pk.data_ptr(), pk.shape[0], 42, 42, None, None, off.data_ptr(), 64,
                                           g.grid_tensor.data_ptr(), g._ws.data_ptr(), g._ws.numel(),
                                           g._counters.data_ptr(), g._strategy, stream)
        assert rc == 0, _native.last_error()
    g._counters.zero_()
    f(); torch.cuda.synchronize()
    c = g.counters(reset=True)
    med, best = timeit(f, 3, 10)
    out['integrate_' + strategy] = {'ms': med, 'best_ms': best, 'updates': c['updates'],
                                    'Gupd_per_s': c['updates'] / med / 1e6, 'slowpath': c['slowpath'], 'records': c['records']}
    print(strategy, out['integrate_' + strategy], flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/first_light.json', 'w'), indent=1)
