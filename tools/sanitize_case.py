"""Small end-to-end case for compute-sanitizer: both integrate strategies, explicit rays, routing
(send-buffer and peer-store kernels on local buffers), pose integration and a 3-grid merge."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from occgrid_b200 import dual_bot_mapper as M, simulation_tools as st, map_merger as MM
from occgrid_b200.distributed import BandLayout, CudaBandOps
from oracle import c_oracle
s = st.generate_session(n_agents=8, n_packets=6000, grid_size=512, origin=(-12.8, -12.8), seed=3)
want = np.full((512, 512), -1, np.int8)
c_oracle.integrate_packets(s['packets'], want, -12.8, -12.8, 0.05, agent_offsets=s['agent_offsets'])
for strat in ('global_atomic', 'tiled'):
    g = M.OccupancyGrid(strategy=strat, max_batch=6000, **s['grid'])
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'])
    assert np.array_equal(g.grid, want), strat
g = M.OccupancyGrid(size=64)
g.update_rays(np.random.default_rng(0).uniform(-1, 1, (500, 4)), np.ones(500, np.uint8))
layout = BandLayout(512, 2)
from occgrid_b200.distributed import BandBuffers, BandStep
steps = [BandStep(layout, r, 512, 0.05, -12.8, -12.8, 'cuda', 3000) for r in range(2)]
BandBuffers.link([b.buf for b in steps])
for b in steps:
    b.finish_init()
tab = torch.from_numpy(s['agent_offsets']).cuda()
ops = [CudaBandOps(layout, r, 512, 0.05, -12.8, -12.8, 'cuda', 'auto', 6000) for r in range(2)]
for r in range(2):
    sl = slice(r * 3000, (r + 1) * 3000)
    pk = ops[r].stage(s['packets'][sl])
    ops[r].route(pk, None, None, tab)                     # NCCL-variant routing kernels
    steps[r].step(pk, None, None, tab, wait=False)        # fused raycast + route (first step: route only)
for b in steps:
    b.step(None, None, None, None)
    b.check_status()
assert np.array_equal(np.concatenate([b.grid.grid for b in steps]), want)
import __graft_entry__ as ge
ge.merge_smoke()
print('sanitize_case ok')
