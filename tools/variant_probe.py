import os, sys
sys.path.insert(0, '/root/repo')
import torch
from occgrid_b200 import simulation_tools as st, _native
from occgrid_b200.distributed import TiledSwarmMap
npk=2500000; size=4096; origin=(-size*0.05/2,)*2
dev=torch.device('cuda',0)
tmap=TiledSwarmMap(size,0.05,origin[0],origin[1],device=dev,max_batch=npk,exchange='p2p')
s=st.generate_session(n_agents=64,n_packets=npk,grid_size=size,origin=origin,seed=42)
pk=tmap.local.stage_packets(s['packets'])[0]; idx=torch.from_numpy(s['agent_idx']).to(dev); off=torch.from_numpy(s['agent_offsets']).to(dev)
# alternate: route a batch, then flush (raycast only) -> the flush launch is the raycast-only kernel
for rep in range(3):
    tmap.update_packets(pk, agent_offsets=off, agent_idx=idx); tmap.flush()
torch.cuda.synchronize()
_native.profile_begin()
for rep in range(5):
    tmap.update_packets(pk, agent_offsets=off, agent_idx=idx); tmap.flush()
torch.cuda.synchronize()
p=_native.profile_end()
print(os.environ.get('OCC_ROUTE_VARIANT_ALWAYS'), {k:round(v[0]/5,4) for k,v in p.items()})
