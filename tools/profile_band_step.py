"""Single-GPU driver of the fused band step (world = 1, plain receive buffers): the same
k_home_raycast<false, true> kernel the multi-GPU path runs, for ncu.  Prints ms per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occgrid_b200 import simulation_tools as st
from occgrid_b200.distributed import TiledSwarmMap
npk = int(os.environ.get('NPK', '2500000'))
agents = int(os.environ.get('AGENTS', '64'))
size = int(os.environ.get('GRID', '4096'))
origin = (-size * 0.05 / 2,) * 2
dev = torch.device('cuda', 0)
tmap = TiledSwarmMap(size, 0.05, origin[0], origin[1], device=dev, max_batch=npk, exchange='p2p')
sess = [st.generate_session(n_agents=agents, n_packets=npk, grid_size=size, origin=origin, seed=42 + i) for i in range(3)]
d = [(tmap.local.stage_packets(s['packets'])[0], torch.from_numpy(s['agent_idx']).to(dev), torch.from_numpy(s['agent_offsets']).to(dev)) for s in sess]
for i in range(6):
    tmap.update_packets(d[i % 3][0], agent_offsets=d[i % 3][2], agent_idx=d[i % 3][1])
tmap.flush(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = int(os.environ.get('STEPS', '20'))
e0.record()
for i in range(K):
    tmap.update_packets(d[i % 3][0], agent_offsets=d[i % 3][2], agent_idx=d[i % 3][1])
tmap.flush()
e1.record(); torch.cuda.synchronize()
print('ms per step', e0.elapsed_time(e1) / K, tmap.counters())
from occgrid_b200 import _native
_native.profile_begin()
for i in range(5):
    tmap.update_packets(d[i % 3][0], agent_offsets=d[i % 3][2], agent_idx=d[i % 3][1])
tmap.flush(); torch.cuda.synchronize()
print({k: round(v[0] / 5, 4) for k, v in _native.profile_end().items()})
