import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from occgrid_b200 import dual_bot_mapper as M, simulation_tools as st
dev = torch.device('cuda:0')
s = st.generate_session(n_agents=64, n_packets=2_500_000, seed=42)
host = torch.from_numpy(s['packets']).pin_memory()
d = torch.empty_like(host, device=dev)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
print('H2D 105MB pinned: %.3f ms' % t(lambda: d.copy_(host, non_blocking=True)))
for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 22):
    g = M.OccupancyGrid(max_batch=min(chunk, 2_500_000), **s['grid'])
    g.h2d_chunk = chunk
    def f():
        g.update_packets(host, agent_offsets=s['agent_offsets'])
        g.counters(reset=True)
    print('chunk', chunk, 'e2e %.3f ms' % t(f))
