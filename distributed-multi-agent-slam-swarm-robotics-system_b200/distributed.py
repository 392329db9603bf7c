"""Multi-GPU layout of the hot path: one process per GPU, ``torch.distributed``.

Integration — the global ``size`` x ``size`` grid is cut into ``world`` row bands, one per
rank.  Cells are independent given a global beam order and a ray is at most
MAX_DIST_M / res cells long (server_nodes/dual_bot_mapper.py:57), so the only exchange step is
routing records to the band(s) their rays can reach:

    local share --occgrid_route_packets--> decoded 48-byte records grouped by band
                --all_to_all_single (NCCL over NVLink)--> records for my band, canonical order
                --occgrid_integrate_poses(window = my band)--> my rows of the map

There is no collective on the grid itself: bands are disjoint.  The canonical stream order is
"rank 0's share, then rank 1's, ..."; routing is stable and all-to-all concatenates by source
rank, so last-writer-wins on every band equals the single-GPU result (tests run the same
exchange on one GPU and, with the oracle as the device, on 2 gloo ranks on the CPU).

Merge — the reference's fuse is a sequential chain (each callback voxel-filters the whole
accumulated cloud, map_merger.py:58-60), so only the HBM-bound part shards: every rank extracts
and transforms its agents' grids, the point lists are all-gathered, and the (cheap, ordered)
voxel chain is replicated on every rank — see ``ShardedMapMerger``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _native
from ._native import Geom, OccGridError


class BandLayout:
    """Row-band partition of a ``size`` x ``size`` grid (pure host logic)."""

    def __init__(self, size, n_bands):
        if n_bands < 1 or n_bands > 32 or n_bands > size:
            raise ValueError('n_bands must be in 1..min(32, size)')
        self.size = int(size)
        self.n_bands = int(n_bands)
        self.band_y0 = [(self.size * b) // self.n_bands for b in range(self.n_bands + 1)]

    def window(self, b):
        """(x0, y0, w, h) of band b."""
        return (0, self.band_y0[b], self.size, self.band_y0[b + 1] - self.band_y0[b])

    def bands_of_row_interval(self, lo, hi):
        return [b for b in range(self.n_bands) if hi >= self.band_y0[b] and lo < self.band_y0[b + 1]]


class CudaBandOps:
    """Device side of one rank: routing kernel + windowed OccupancyGrid."""

    def __init__(self, layout, rank, size, resolution, origin_x, origin_y, device, strategy, max_batch):
        from .dual_bot_mapper import OccupancyGrid
        self.layout = layout
        self.device = torch.device(device)
        self.grid = OccupancyGrid(size, resolution, origin_x, origin_y, device=device,
                                  window=layout.window(rank), strategy=strategy, max_batch=max_batch)
        self._lib = _native.lib()
        self._geom = Geom(float(origin_x), float(origin_y), float(resolution), int(size), int(size), 0, 0, int(size), int(size))
        self._band_y0 = np.asarray(layout.band_y0, np.int32)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._counts = torch.zeros(layout.n_bands, dtype=torch.int64, device=self.device)
        self._ws = None
        self._send = None

    def stage(self, packets):
        return self.grid.stage_packets(packets)[0]

    def route(self, packets, agent_idx, drift, agent_table):
        """-> (rows uint8 [m, 48] of decoded records grouped by destination band, counts list)"""
        n, stride = packets.shape
        nb = self.layout.n_bands
        cap = 2 * n + 1024
        need = self._lib.occgrid_route_workspace_bytes(n, nb)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if self._send is None or self._send.shape[0] < cap:
            self._send = torch.empty((cap, 48), dtype=torch.uint8, device=self.device)
        rc = self._lib.occgrid_route_packets(
            self._geom, nb, self._band_y0.ctypes.data, packets.data_ptr(), n, stride, 42 if stride >= 42 else 41,
            agent_idx.data_ptr() if agent_idx is not None else None, drift.data_ptr() if drift is not None else None,
            agent_table.data_ptr(), agent_table.shape[0] - 1, self._send.data_ptr(), cap,
            self._counts.data_ptr(), self._status.data_ptr(), None,
            self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        _native.check(rc, 'occgrid_route_packets')
        host = torch.cat([self._counts, self._status.to(torch.int64)]).cpu().tolist()   # the only host sync of a step
        counts, status = host[:nb], host[nb]
        if status & 1:
            self._status.zero_()
            raise OccGridError('route: send buffer overflow')
        return self._send[:int(sum(counts))], counts

    def route_p2p(self, packets, agent_idx, drift, agent_table, ordinal_base, recv_ptrs, count_ptrs, capacity):
        """Fused route + exchange: records go straight into the owners' receive buffers."""
        n, stride = packets.shape
        rc = self._lib.occgrid_route_packets_p2p(
            self._geom, self.layout.n_bands, self._band_y0.ctypes.data, packets.data_ptr(), n, stride,
            42 if stride >= 42 else 41,
            agent_idx.data_ptr() if agent_idx is not None else None, drift.data_ptr() if drift is not None else None,
            agent_table.data_ptr(), agent_table.shape[0] - 1, int(ordinal_base), recv_ptrs.data_ptr(), count_ptrs.data_ptr(),
            int(capacity), self._status.data_ptr(), None, torch.cuda.current_stream(self.device).cuda_stream)
        _native.check(rc, 'occgrid_route_packets_p2p')

    def empty(self, rows, stride, dtype):
        shape = (rows, stride) if stride else (rows,)
        return torch.empty(shape, dtype=dtype, device=self.device)

    def integrate(self, rows):
        self.grid.update_poses(rows)

    def band_tensor(self):
        return self.grid.grid_tensor


class PeerExchange:
    """Receive buffers in symmetric (peer-mapped) memory for the fused route+exchange kernel
    ``occgrid_route_packets_p2p``: three slots per rank so that one cross-GPU barrier per step is
    enough (passing barrier j implies every rank finished integrating batch j-2, whose slot is the
    one batch j+1 will be written into)."""
    SLOTS = 3

    def __init__(self, device, group, capacity):
        import torch.distributed._symmetric_memory as symm
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.capacity = int(capacity)
        pg = group if group is not None else dist.group.WORLD
        self.recv = symm.empty((self.SLOTS, self.capacity, 48), dtype=torch.uint8, device=device)
        self.count = symm.empty((self.SLOTS, 64), dtype=torch.int32, device=device)      # one counter per 256 B
        self.count.zero_()
        self.h_recv = symm.rendezvous(self.recv, pg)
        self.h_count = symm.rendezvous(self.count, pg)
        self.recv_ptrs, self.count_ptrs = [], []
        for s in range(self.SLOTS):
            self.recv_ptrs.append(torch.tensor([int(p) + s * self.capacity * 48 for p in self.h_recv.buffer_ptrs],
                                               dtype=torch.int64, device=device))
            self.count_ptrs.append(torch.tensor([int(p) + s * 256 for p in self.h_count.buffer_ptrs],
                                                dtype=torch.int64, device=device))
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        self.group = group

    def barrier(self):
        """Cross-GPU barrier ordered on the current stream."""
        try:
            self.h_recv.barrier(channel=0)
        except Exception:
            dist.barrier(group=self.group)


class TiledSwarmMap:
    """A global occupancy grid spatially tiled over the ranks of a process group.  Mirrors
    ``OccupancyGrid``'s batched entry; each rank passes ITS share of the packet stream."""

    def __init__(self, size, resolution=0.05, origin_x=-5.0, origin_y=-5.0, *, group=None, device=None,
                 strategy='auto', max_batch=1 << 16, ops=None, pipeline=False, exchange='nccl'):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size, self.res, self.ox, self.oy = int(size), float(resolution), float(origin_x), float(origin_y)
        self.layout = BandLayout(size, self.world)
        if ops is None:
            if not torch.cuda.is_available():
                raise OccGridError('no CUDA device: TiledSwarmMap has no CPU fallback')
            device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
            ops = CudaBandOps(self.layout, self.rank, size, resolution, origin_x, origin_y, device, strategy, max_batch)
        self.ops = ops
        self.local = getattr(ops, 'grid', None)
        self._recv_bufs = {}
        self.pipeline = bool(pipeline) and hasattr(ops, 'grid')
        self._pending = None
        self._step = 0
        self._slot_free = [None, None]
        self._side = torch.cuda.Stream(device=ops.device, priority=-1) if self.pipeline else None   # router first
        # exchange = 'p2p': routed records are stored straight into the owner's receive buffer over
        # NVLink by the routing kernel (symmetric memory); 'nccl': send buffer + all_to_all_single
        self.peer = None
        self.exchange = 'nccl'
        if exchange in ('p2p', 'auto') and self.world > 1 and hasattr(ops, 'grid'):
            try:
                self.peer = PeerExchange(ops.device, group, int(max_batch * 1.6) + 4096)
                self.exchange = 'p2p'
            except Exception as e:                      # no symmetric memory on this system: NCCL path
                if exchange == 'p2p':
                    raise
                self.peer = None
        self._done = {}

    def _agent_table(self, separation, agent_offsets):
        if isinstance(agent_offsets, torch.Tensor):
            t = agent_offsets.to(dtype=torch.float64).reshape(-1, 2).contiguous()
        elif agent_offsets is None:
            t = torch.tensor([[0.0, 0.0], [0.0, 0.0], [float(separation), 0.0]], dtype=torch.float64)
        else:
            t = torch.from_numpy(np.ascontiguousarray(agent_offsets, np.float64).reshape(-1, 2))
        return t.to(self.ops.device)

    def _exchange(self, send, counts, stride, dtype, slot=0):
        """all_to_all_single of row-segments; returns the rows received (concatenated by source
        rank, i.e. in canonical stream order)."""
        if self.world == 1 and not self.pipeline:
            return send
        recv_counts = self._recv_counts if self.world > 1 else [int(send.shape[0])]
        rows = int(sum(recv_counts))
        key = (stride, dtype, slot)
        buf = self._recv_bufs.get(key)
        if buf is None or buf.shape[0] < rows:          # grown geometrically, then reused every step
            buf = self.ops.empty(int(rows * 1.25) + 1024, stride, dtype)
            self._recv_bufs[key] = buf
        out = buf[:rows]
        if self.world == 1:              # pipelined single rank: the send buffer is reused by the next route
            out.copy_(send)
            return out
        dist.all_to_all_single(out, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=counts,
                               group=self.group)
        return out

    def _route_and_exchange(self, packets, separation, drift, agent_offsets, agent_idx, slot):
        tab = self._agent_table(separation, agent_offsets)
        pk = self.ops.stage(packets)
        dev = pk.device
        idx = torch.as_tensor(agent_idx, dtype=torch.int32).to(dev).contiguous() if agent_idx is not None else None
        dr = torch.as_tensor(drift, dtype=torch.float64).reshape(-1, 2).to(dev).contiguous() if drift is not None else None
        send, counts = self.ops.route(pk, idx, dr, tab)
        if self.world > 1:
            c_in = torch.tensor(counts, dtype=torch.int64, device=dev)
            c_out = torch.empty_like(c_in)
            dist.all_to_all_single(c_out, c_in, group=self.group)
            self._recv_counts = c_out.cpu().tolist()
        return self._exchange(send, counts, send.shape[1], torch.uint8, slot)

    def update_packets(self, packets, separation=0.0, drift=None, agent_offsets=None, agent_idx=None):
        """Integrate this rank's share of the stream into the tiled map (all ranks must call).

        With ``pipeline=True`` the call routes and exchanges THIS batch on a side stream while
        the PREVIOUS batch is being integrated (NVLink traffic hidden behind SM work); the map is
        complete after ``flush()`` (``gather_grid``/``counters`` flush first).  Batch order, and
        with it last-writer-wins, is unchanged."""
        if self.peer is not None:
            return self._update_packets_p2p(packets, separation, drift, agent_offsets, agent_idx)
        if not self.pipeline:
            recv = self._route_and_exchange(packets, separation, drift, agent_offsets, agent_idx, 0)
            self.ops.integrate(recv)
            return int(recv.shape[0])
        main = torch.cuda.current_stream(self.ops.device)
        slot = self._step & 1
        self._step += 1
        with torch.cuda.stream(self._side):
            if self._slot_free[slot] is not None:
                self._side.wait_event(self._slot_free[slot])      # integrate that last read this buffer
            recv = self._route_and_exchange(packets, separation, drift, agent_offsets, agent_idx, slot)
            ready = torch.cuda.Event()
            ready.record(self._side)
        prev, self._pending = self._pending, (recv, ready, slot)
        if prev is not None:
            self._integrate_pending(prev, main)
        return int(recv.shape[0])

    def _update_packets_p2p(self, packets, separation, drift, agent_offsets, agent_idx):
        """Fused route+exchange over peer memory.  Side stream (or the current one when not
        pipelined): wait for my integrate of batch j-2, route batch j into the owners' slot j%3,
        cross-GPU barrier, read my fill counter.  Main stream: integrate my slot with the
        ordinals carried by the records, zero its counter."""
        px = self.peer
        dev = self.ops.device
        main = torch.cuda.current_stream(dev)
        j = self._step
        self._step += 1
        slot = j % px.SLOTS
        # pipeline: put the PREVIOUS batch's integration on the main stream first, so that it runs
        # while this batch is routed over NVLink on the side stream (the host then waits only for
        # the router's fill counter)
        if self.pipeline and self._pending is not None:
            prev, self._pending = self._pending, None
            self._integrate_p2p(prev, main)
        stream = self._side if self.pipeline else main
        with torch.cuda.stream(stream):
            if self.pipeline and (j - 2) in self._done:
                stream.wait_event(self._done.pop(j - 2))
            tab = self._agent_table(separation, agent_offsets)
            pk = self.ops.stage(packets)
            idx = torch.as_tensor(agent_idx, dtype=torch.int32).to(dev).contiguous() if agent_idx is not None else None
            dr = torch.as_tensor(drift, dtype=torch.float64).reshape(-1, 2).to(dev).contiguous() if drift is not None else None
            self.ops.route_p2p(pk, idx, dr, tab, self.rank * ((1 << 29) // (self.world + 1)), px.recv_ptrs[slot],
                               px.count_ptrs[slot], px.capacity)
            px.barrier()
            host = torch.cat([px.count[slot, :1], self.ops._status]).cpu().tolist()      # the one host sync of a step
            n_recv, status = int(host[0]), int(host[1])
            if status & 1:
                self.ops._status.zero_()
                raise OccGridError('route_p2p: a receive buffer overflowed')
            ready = torch.cuda.Event()
            ready.record(stream)
        pending = (slot, n_recv, ready, j)
        if not self.pipeline:
            self._integrate_p2p(pending, main)
        else:
            self._pending = pending
        return n_recv

    def _integrate_p2p(self, pending, main):
        slot, n_recv, ready, j = pending
        px = self.peer
        main.wait_event(ready)
        self.ops.grid.update_poses(px.recv[slot, :n_recv], ordinals_in_records=True)
        px.count[slot].zero_()
        done = torch.cuda.Event()
        done.record(main)
        self._done[j] = done

    def _integrate_pending(self, pending, main):
        recv, ready, slot = pending
        main.wait_event(ready)
        self.ops.integrate(recv)
        done = torch.cuda.Event()
        done.record(main)
        self._slot_free[slot] = done

    def flush(self):
        """Integrate the batch still in flight (pipeline mode)."""
        if self.pipeline and self._pending is not None:
            pending, self._pending = self._pending, None
            main = torch.cuda.current_stream(self.ops.device)
            if self.peer is not None:
                self._integrate_p2p(pending, main)
            else:
                self._integrate_pending(pending, main)

    def gather_grid(self):
        """Assemble the global map on every rank (all_gather of the disjoint bands)."""
        self.flush()
        band = self.ops.band_tensor()
        if self.world == 1:
            return band.cpu().numpy()
        bands = [self.ops.empty(self.layout.window(b)[3], self.size, torch.int8) for b in range(self.world)]
        dist.all_gather(bands, band.contiguous(), group=self.group)
        return torch.cat(bands, dim=0).cpu().numpy()


# ----------------------------------------------------------------------------------------------
#  Merge: shard the extraction, replicate the ordered voxel chain
# ----------------------------------------------------------------------------------------------

class ShardedMapMerger:
    """Agents are dealt round-robin to ranks; every rank extracts + transforms the occupied cells
    of ITS agents (the HBM-bound scan, map_merger.py:71-77 + :58), the per-agent point lists are
    all-gathered, and the order-dependent voxel chain (:59-60) is replayed identically on every
    rank.  Result == MapMerger.merge on one GPU."""

    def __init__(self, group=None, device=None):
        from .map_merger import MapMerger
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merger = MapMerger(device=device if device is not None else 'cuda')

    def merge(self, local_grids, local_origins, res, local_transforms, n_agents_total):
        """``local_*`` hold the agents a with a % world == rank, in increasing a."""
        from .map_merger import MapMerger, make_grid_msg, se2_matrix
        dev = self.merger.device
        mine = list(range(self.rank, n_agents_total, self.world))
        # the first NON-EMPTY grid is adopted untransformed (:40-43): find it globally first
        nonempty = torch.zeros(n_agents_total, dtype=torch.int64, device=dev)
        for j, a in enumerate(mine):
            g = local_grids[j]
            g = g if isinstance(g, torch.Tensor) else torch.from_numpy(np.asarray(g))
            nonempty[a] = int(bool((g > 50).any()))
        if self.world > 1:
            dist.all_reduce(nonempty, group=self.group)
        ne = nonempty.cpu().tolist()
        first_agent = next((a for a in range(n_agents_total) if ne[a]), -1)
        pts = []
        for j, a in enumerate(mine):
            tmp = MapMerger(device=dev)
            tmp._ensure_capacity(1 << 16)
            T = local_transforms[j] if (local_transforms is not None and a != first_agent) else None
            if T is not None and np.asarray(T).size == 3:
                T = se2_matrix(*np.asarray(T, np.float64).tolist())
            h, w = local_grids[j].shape
            k = tmp._extract(make_grid_msg(local_grids[j], w, h, res, local_origins[j][0], local_origins[j][1]), T)
            pts.append(torch.stack([tmp._cloud.x[:k], tmp._cloud.y[:k]], dim=1))
        # all-gather variable-length point lists (pad to the global max)
        lens = torch.zeros(n_agents_total, dtype=torch.int64, device=dev)
        for j, a in enumerate(mine):
            lens[a] = pts[j].shape[0]
        if self.world > 1:
            dist.all_reduce(lens, group=self.group)
        lens_h = lens.cpu().tolist()
        per_rank = -(-n_agents_total // self.world)
        mx = max(max(lens_h), 1)
        buf = torch.zeros((per_rank, mx, 2), dtype=torch.float64, device=dev)
        for j in range(len(mine)):
            buf[j, :pts[j].shape[0]] = pts[j]
        if self.world > 1:
            allbuf = [torch.empty_like(buf) for _ in range(self.world)]
            dist.all_gather(allbuf, buf, group=self.group)
        else:
            allbuf = [buf]
        # every rank now holds every slice: stage them in agent order and replay the ordered voxel
        # chain (:59-60) with the incremental chain kernels, identically on every rank
        from .map_merger import _Cloud
        from . import _native
        m = self.merger
        order = [a for a in range(n_agents_total) if lens_h[a] > 0]
        total = sum(lens_h)
        if total:
            with torch.cuda.device(dev):
                stage = _Cloud(total + 16, dev)
                offs_h = np.zeros(n_agents_total + 1, np.int64)
                offs_h[1:] = np.cumsum(lens_h)
                for a in order:
                    p = allbuf[a % self.world][a // self.world, :lens_h[a]]
                    stage.x[offs_h[a]:offs_h[a + 1]].copy_(p[:, 0])
                    stage.y[offs_h[a]:offs_h[a + 1]].copy_(p[:, 1])
                offs = torch.from_numpy(offs_h).to(dev)
                lo = torch.stack([stage.x[:total].min(), stage.y[:total].min()])
                hi = torch.stack([stage.x[:total].max(), stage.y[:total].max()])
                if m._n_global:
                    n = m._n_global
                    lo = torch.minimum(lo, torch.stack([m._cloud.x[:n].min(), m._cloud.y[:n].min()]))
                    hi = torch.maximum(hi, torch.stack([m._cloud.x[:n].max(), m._cloud.y[:n].max()]))
                bb = torch.cat([lo, hi]).cpu().tolist()
                m._ensure_capacity(m._n_global + total + 1024)
                if m._n_global == 0:                    # adopted as is (:40-43)
                    a0 = order.pop(0)
                    rc = m._lib.mapmerge_append_slice(stage.x.data_ptr(), stage.y.data_ptr(), offs.data_ptr(), a0,
                                                      m._cloud.x.data_ptr(), m._cloud.y.data_ptr(), m._cloud.capacity,
                                                      m._cloud.count.data_ptr(), m._status.data_ptr(), None, m._stream())
                    _native.check(rc, 'mapmerge_append_slice')
                    m.map_resolution = float(res)
                v = m.map_resolution
                if order:
                    # means stay inside the hull of the points, so these bounds cover every cloud of the chain
                    lat_w, lat_h = int((bb[2] - bb[0]) / v) + 4, int((bb[3] - bb[1]) / v) + 4
                    m.chain_stats = m._run_chain(stage, offs, n_agents_total, order, lat_w, lat_h, max(lens_h))
                m._n_global = int(m._cloud.count.item())
                m._check_status()
        out = m.publish_global_map()
        return (out.data, (out.info.origin.position.x, out.info.origin.position.y)) if out is not None else (None, None)


# ----------------------------------------------------------------------------------------------
#  bench.py helper: weak-scaling sessions
# ----------------------------------------------------------------------------------------------

def make_rank_sessions(world, rank, device, packets_per_rank, pool, strategy, pipeline=True,
                       grid_per_gpu=4096, agents_per_gpu=64, exchange='auto', ingest='uniform'):
    """Weak scaling of BASELINE configs[1]: (4096*world)^2 map, 64*world agents, each rank ingests
    its own `packets_per_rank` share.  ``ingest='uniform'`` (default, the worst case): every
    rank's share covers ALL agents, so (world-1)/world of the records are routed to another GPU;
    ``'affine'``: a rank receives the agents that drive inside its own band (only rays crossing a
    band edge are routed).  Returns (TiledSwarmMap, sessions, step_fn)."""
    from . import simulation_tools as st
    side = grid_per_gpu * world
    origin = (-side * 0.05 / 2.0,) * 2
    tmap = TiledSwarmMap(side, 0.05, origin[0], origin[1], device=device, strategy=strategy,
                         max_batch=int(packets_per_rank * 1.25), pipeline=pipeline, exchange=exchange)
    sessions = []
    for i in range(pool):
        # this rank's share of the stream: all 64*world agents, `packets_per_rank` records
        if ingest == 'affine':
            # this rank's own agents, their rooms on a lattice inside the band it owns
            y0 = tmap.layout.window(rank)[1]
            sh = st.generate_session(n_agents=agents_per_gpu, n_packets=packets_per_rank, grid_size=grid_per_gpu,
                                     origin=(origin[0], origin[1] + y0 * 0.05),
                                     seed=1000 + 97 * i + rank)
        else:
            sh = st.generate_session(n_agents=agents_per_gpu * world, n_packets=packets_per_rank, grid_size=side,
                                     origin=origin, seed=1000 + 97 * i + rank)
        sessions.append({'packets': tmap.ops.stage(sh['packets']),
                         'agent_idx': torch.from_numpy(sh['agent_idx']).to(device),
                         'agent_offsets': torch.from_numpy(sh['agent_offsets']).to(device), 'grid': sh['grid']})

    def step(i):
        s = sessions[i % pool]
        tmap.update_packets(s['packets'], agent_offsets=s['agent_offsets'], agent_idx=s['agent_idx'])

    return tmap, sessions, step
