"""Per-phase wall-clock of TiledSwarmMap.update_packets and raw all-to-all bandwidth (torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
def sync(): torch.cuda.synchronize()
# raw all-to-all
for mb in (8, 105):
    n = mb * 1000 * 1000 // world * world
    a = torch.empty(n, dtype=torch.uint8, device=dev); b = torch.empty_like(a)
    for _ in range(3): dist.all_to_all_single(b, a)
    sync(); t = time.perf_counter()
    for _ in range(10): dist.all_to_all_single(b, a)
    sync(); dt = (time.perf_counter() - t) / 10
    if rank == 0: print(f'all_to_all {mb} MB: {dt*1e3:.3f} ms  ({n/dt/1e9:.1f} GB/s per rank)', flush=True)
from occgrid_b200.distributed import make_rank_sessions
tmap, sessions, step = make_rank_sessions(world, rank, dev, 2_500_000, 2, 'auto')
for i in range(3): step(i)
sync(); dist.barrier()
import occgrid_b200.distributed as D
ops = tmap.ops
s = sessions[0]
tab = tmap._agent_table(0.0, s['agent_offsets'])
for rep in range(3):
    T = {}
    sync(); t0 = time.perf_counter()
    send, s_idx, s_dr, counts = ops.route(s['packets'], s['agent_idx'], None, tab); sync(); T['route'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    c_in = torch.tensor(counts, dtype=torch.int64, device=dev); c_out = torch.empty_like(c_in)
    dist.all_to_all_single(c_out, c_in); tmap._recv_counts = c_out.cpu().tolist(); sync(); T['counts'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    recv = tmap._exchange(send, counts, send.shape[1], torch.uint8); sync(); T['a2a_packets'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    r_idx = tmap._exchange(s_idx, counts, 0, torch.int32); sync(); T['a2a_idx'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ops.integrate(recv, r_idx, None, tab); sync(); T['integrate'] = time.perf_counter() - t0
    if rank == 0: print({k: round(v * 1e3, 3) for k, v in T.items()}, 'counts', counts, 'recv', tmap._recv_counts, flush=True)
dist.destroy_process_group()
