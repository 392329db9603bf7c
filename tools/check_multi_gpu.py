"""Multi-GPU parity check (run under torchrun on N GPUs): every rank integrates its share of a
seeded session through TiledSwarmMap (fused raycast + route over peer memory, and the NCCL all-to-all variant), the bands are all-gathered and rank 0 compares the assembled map with the C oracle
run over the canonical stream (batch by batch, rank 0's share first).  Prints PASS/FAIL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
from occgrid_b200 import simulation_tools as st
from occgrid_b200.distributed import TiledSwarmMap
size = int(os.environ.get('CHECK_GRID_PER_GPU', '2048')) * world
agents_per_gpu = int(os.environ.get('CHECK_AGENTS_PER_GPU', '32'))
origin = (-size * 0.05 / 2,) * 2
n_batches, per_rank = 3, 60_000
sess = st.generate_session(n_agents=agents_per_gpu * world, n_packets=n_batches * per_rank * world, grid_size=size, origin=origin, seed=77)
pk, idx, offs = sess['packets'], sess['agent_idx'], sess['agent_offsets']
ok = True
modes = ['p2p'] if os.environ.get('CHECK_P2P_ONLY') else ['nccl', 'p2p']
for exchange in modes:
    tmap = TiledSwarmMap(size, 0.05, origin[0], origin[1], device=dev, max_batch=per_rank, exchange=exchange)
    order = []
    for b in range(n_batches):
        base = b * per_rank * world
        sl = slice(base + rank * per_rank, base + (rank + 1) * per_rank)
        tmap.update_packets(pk[sl], agent_offsets=offs, agent_idx=idx[sl])
        order.append(np.arange(base, base + per_rank * world))      # canonical: rank 0's share, rank 1's, ...
    got = tmap.gather_grid()
    if rank == 0:
        from oracle import c_oracle
        o = np.concatenate(order)
        want = np.full((size, size), -1, np.int8)
        c = c_oracle.integrate_packets(pk[o], want, origin[0], origin[1], 0.05, agent_offsets=offs, agent_idx=idx[o])
        same = bool(np.array_equal(got, want))
        ok &= same
        print(f'world={world} exchange={tmap.exchange}: map {"bit-exact" if same else "DIFFERS"} vs oracle '
              f'({c["beams"]} beams, {int((want != -1).sum())} known cells)', flush=True)
    dist.barrier()
# map fusion: agent blocks per rank, batched extraction sharded, ordered voxel chain replicated (mode 'exact');
# and the SURVEY §8e partitioning (per-rank chains, bounds all-reduce, max fuse of partial rasters; not a parity mode)
import math
from occgrid_b200.distributed import ShardedMapMerger, agent_block
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from merge_util import synth_agent_grid
from oracle import merge_oracle as MO
A, S = 3 * world + 1, 256
rgen = np.random.default_rng(3)
grids = [synth_agent_grid(S, 500 + a) for a in range(A)]
tf = [MO.se2_matrix(*rgen.uniform(-5, 5, 2), rgen.uniform(-math.pi, math.pi)) for _ in range(A)]
lo, hi = agent_block(A, world, rank)
got, origin = ShardedMapMerger(device=dev).merge(grids[lo:hi], [(-6.4, -6.4)] * (hi - lo), 0.05, tf[lo:hi], A)
fused, forigin = ShardedMapMerger(device=dev, mode='raster_fuse').merge(grids[lo:hi], [(-6.4, -6.4)] * (hi - lo), 0.05, tf[lo:hi], A)
if rank == 0:
    o = MO.OracleMerger()
    for a in range(A):
        want = o.map_callback(grids[a].ravel(), S, S, 0.05, -6.4, -6.4, tf[a])
    same = bool(np.array_equal(got, want[0]) and origin == want[1])
    ok &= same
    print(f'world={world} sharded merge of {A} grids (exact): {"bit-exact" if same else "DIFFERS"} vs oracle', flush=True)
    if fused.shape == want[0].shape:
        print(f'world={world} raster_fuse: {int((fused != want[0]).sum())} of {fused.size} cells differ from the reference map '
              f'(occupied: {int((fused == 100).sum())} vs {int((want[0] == 100).sum())})', flush=True)
    else:
        print(f'world={world} raster_fuse: map {fused.shape} vs reference {want[0].shape}', flush=True)
dist.barrier()
if rank == 0:
    print('PASS' if ok else 'FAIL', flush=True)
dist.destroy_process_group()
