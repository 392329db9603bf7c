"""Multi-GPU layout of the hot path: one process per GPU, ``torch.distributed``.

Integration — the global ``size`` x ``size`` grid is cut into ``world`` row bands, one per
rank.  Cells are independent given a global beam order and a ray is at most
MAX_DIST_M / res cells long (server_nodes/dual_bot_mapper.py:57), so the only exchange step is
routing records to the band(s) their rays can reach.  Two exchange paths, same map:

``exchange='p2p'`` (default on GPUs) — the fused band step of ``csrc/occgrid_band.cu``:

    call j:  occgrid_band_prepare(slot j-1)         bin what arrived during call j-1
             occgrid_band_raycast_route(...)        ONE persistent kernel: raycast batch j-1 and,
                                                    between its work items, decode batch j and
                                                    store the 48-byte records straight into the
                                                    band owners' per-source segments over NVLink
             occgrid_band_publish(epoch j+1)        fill counts to the owners + cross-GPU barrier

  everything stream-ordered on the caller's stream, no host read-back, two receive slots in
  ``torch.distributed._symmetric_memory`` (plumbing: it hands out the peer pointers).

``exchange='nccl'`` — ``occgrid_route_packets`` (records grouped by band in a send buffer) +
``all_to_all_single`` + ``occgrid_integrate_poses``; also the path the CPU test-suite drives on
``gloo`` with the oracle standing in for the device.

There is no collective on the grid itself: bands are disjoint.  The canonical stream order is
"rank 0's share, then rank 1's, ..."; every record carries its ordinal, so last-writer-wins on
every band equals the single-GPU result.

Merge — ``ShardedMapMerger``: agents are sharded, every rank rasterises its agents' clouds into a
partial global grid over the all-reduced bounds and the partial grids are fused with a max
reduction (occupied wins, map_merger.py:103-111 is idempotent and order-free).
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np
import torch
import torch.distributed as dist

from . import _native
from ._native import Geom, OccGridError

log = logging.getLogger('occgrid_b200.distributed')
SEG_CHUNK = 2048            # kSegChunk: segment capacities are multiples of this


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
    """Device side of one rank: windowed OccupancyGrid + the NCCL-variant routing kernel."""

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
        host = torch.cat([self._counts, self._status.to(torch.int64)]).cpu().tolist()   # the NCCL variant needs the split sizes
        counts, status = host[:nb], host[nb]
        if status & 1:
            self._status.zero_()
            raise OccGridError('route: send buffer overflow')
        return self._send[:int(sum(counts))], counts

    def empty(self, rows, stride, dtype):
        shape = (rows, stride) if stride else (rows,)
        return torch.empty(shape, dtype=dtype, device=self.device)

    def integrate(self, rows):
        self.grid.update_poses(rows)

    def band_tensor(self):
        return self.grid.grid_tensor


class BandBuffers:
    """Receive side of the fused band step for ONE rank: two slots of ``world`` per-source segments
    of ``seg_cap`` 48-byte records, the per-slot fill counts and the barrier flags.  ``symmetric``:
    allocate in ``torch.distributed._symmetric_memory`` and rendezvous, so that every rank can
    address every other rank's buffers over NVLink; otherwise plain device tensors (world == 1, or
    several logical ranks on one GPU in the tests)."""
    SLOTS = 2
    CTL_WORDS = 64 * (SLOTS + 1)          # seg_counts[slot][64] ..., flags[64]

    def __init__(self, device, world, seg_cap, symmetric=False, group=None):
        self.device, self.world, self.seg_cap = device, int(world), int(seg_cap)
        rec_shape = (self.SLOTS, self.world, self.seg_cap, 48)
        if symmetric:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            self.recv = symm.empty(rec_shape, dtype=torch.uint8, device=device)
            self.tiles = symm.empty(rec_shape[:3], dtype=torch.int32, device=device)
            self.ctl = symm.empty((self.CTL_WORDS,), dtype=torch.int32, device=device)
            self.ctl.zero_()
            torch.cuda.synchronize(device)
            h_recv, h_ctl, h_tiles = symm.rendezvous(self.recv, pg), symm.rendezvous(self.ctl, pg), symm.rendezvous(self.tiles, pg)
            self._handles = (h_recv, h_ctl, h_tiles)
            self.recv_bases = [int(p) for p in h_recv.buffer_ptrs]
            self.tile_bases = [int(p) for p in h_tiles.buffer_ptrs]
            self.ctl_bases = [int(p) for p in h_ctl.buffer_ptrs]
        else:
            self.recv = torch.empty(rec_shape, dtype=torch.uint8, device=device)
            self.tiles = torch.empty(rec_shape[:3], dtype=torch.int32, device=device)
            self.ctl = torch.zeros((self.CTL_WORDS,), dtype=torch.int32, device=device)
            self.recv_bases = self.ctl_bases = self.tile_bases = None         # filled by link()

    def slot_ptr(self, slot):
        return self.recv.data_ptr() + slot * self.world * self.seg_cap * 48

    def tiles_ptr(self, slot):
        return self.tiles.data_ptr() + slot * self.world * self.seg_cap * 4

    def seg_counts_ptr(self, slot):
        return self.ctl.data_ptr() + slot * 256

    def flags_ptr(self):
        return self.ctl.data_ptr() + self.SLOTS * 256

    @staticmethod
    def link(buffers):
        """Plain (non-symmetric) buffers of all logical ranks on one device: exchange base pointers."""
        for b in buffers:
            b.recv_bases = [o.recv.data_ptr() for o in buffers]
            b.ctl_bases = [o.ctl.data_ptr() for o in buffers]
            b.tile_bases = [o.tiles.data_ptr() for o in buffers]

    def pointer_tables(self):
        """Device arrays of peer pointers: per slot the owners' slot bases and seg_counts, plus flags."""
        dev = self.device
        slot_bytes = self.world * self.seg_cap * 48
        t = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        self.peer_recs = [t([p + s * slot_bytes for p in self.recv_bases]) for s in range(self.SLOTS)]
        self.peer_tiles = [t([p + s * (slot_bytes // 12) for p in self.tile_bases]) for s in range(self.SLOTS)]
        self.peer_seg_counts = [t([p + s * 256 for p in self.ctl_bases]) for s in range(self.SLOTS)]
        self.peer_flags = t([p + self.SLOTS * 256 for p in self.ctl_bases])


class BandStep:
    """The fused band step of one rank (C ABI ``occgrid_band_*``): owns the windowed grid, the tiled
    workspace sized for ``world`` segments, the local reservation counters and the step state."""

    def __init__(self, layout, rank, size, resolution, origin_x, origin_y, device, max_batch, buffers=None,
                 symmetric=False, group=None):
        from .dual_bot_mapper import OccupancyGrid
        self.layout, self.rank, self.world = layout, int(rank), layout.n_bands
        self.device = torch.device(device)
        self._lib = _native.lib()
        self.size, self.res, self.ox, self.oy = int(size), float(resolution), float(origin_x), float(origin_y)
        self.seg_cap = -(-max(int(max_batch), 1) // SEG_CHUNK) * SEG_CHUNK
        self.ordinal_stride = (1 << 29) // (self.world + 1)
        if self.seg_cap > self.ordinal_stride:
            raise OccGridError(f'max_batch {max_batch} exceeds the per-rank ordinal slice {self.ordinal_stride} '
                               f'(2^29 order stamps shared by {self.world} ranks)')
        # the band's own grid: strategy tiled (the fused kernel IS the tiled raycast); its generic workspace stays minimal
        self.grid = OccupancyGrid(size, resolution, origin_x, origin_y, device=device, window=layout.window(rank),
                                  strategy='tiled', max_batch=1, lazy_workspace=True)
        self._geom = self.grid._geom
        with torch.cuda.device(self.device):
            nbytes = self._lib.occgrid_band_workspace_bytes(self._geom, self.world, self.seg_cap)
            if nbytes == 0:
                raise OccGridError('occgrid_band_workspace_bytes: ' + _native.last_error())
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._resv = torch.zeros(64, dtype=torch.int32, device=self.device)
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.buf = buffers if buffers is not None else BandBuffers(self.device, self.world, self.seg_cap, symmetric, group)
        self._job = _native.RouteJob()
        j = self._job
        j.ox, j.oy, j.res, j.size_x = self.ox, self.oy, self.res, self.size
        j.n_bands, j.src_rank = self.world, self.rank
        for b in range(33):
            j.band_y0[b] = layout.band_y0[min(b, self.world)]
        j.seg_capacity = self.seg_cap
        j.d_resv, j.d_status = self._resv.data_ptr(), self._status.data_ptr()
        j.d_counters = self.grid._counters.data_ptr()
        with torch.cuda.device(self.device):
            self._side = torch.cuda.Stream(device=self.device)
        self._step = 0
        self._pending = False       # a routed batch sits in slot (_step - 1) % 2, not yet integrated
        self._keep = None           # tensors the in-flight job points at

    def finish_init(self):
        """After every rank's buffers exist (and, for plain buffers, BandBuffers.link ran)."""
        with torch.cuda.device(self.device):
            self.buf.pointer_tables()
        b, c = self.buf, _native.BandCtx()
        c.band_geom = self._geom
        c.n_bands, c.rank, c.seg_capacity = self.world, self.rank, self.seg_cap
        for s in range(b.SLOTS):
            c.d_recv[s], c.d_recv_tiles[s], c.d_seg_counts[s] = b.slot_ptr(s), b.tiles_ptr(s), b.seg_counts_ptr(s)
            c.d_peer_recs[s], c.d_peer_tiles[s] = b.peer_recs[s].data_ptr(), b.peer_tiles[s].data_ptr()
            c.d_peer_seg_counts[s] = b.peer_seg_counts[s].data_ptr()
        c.d_peer_flags, c.d_my_flags = b.peer_flags.data_ptr(), b.flags_ptr()
        c.d_resv, c.d_status = self._resv.data_ptr(), self._status.data_ptr()
        c.d_grid = self.grid.grid_tensor.data_ptr()
        c.d_workspace, c.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        c.d_counters = self.grid._counters.data_ptr()
        self._ctx = c
        self._ctx_ref = ctypes.byref(c)
        self._job_ref = ctypes.byref(self._job)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def step(self, pk, agent_idx, drift, agent_table, wait=True):
        """Integrate the pending batch and route `pk` (uint8 cuda [n, stride]; None = nothing more to
        route) in ONE fused kernel, then publish + barrier — one host call (occgrid_band_step)."""
        job = None
        if pk is not None:
            n, stride = pk.shape
            if n > self.seg_cap:
                raise OccGridError(f'batch of {n} records exceeds max_batch (segment capacity {self.seg_cap})')
            j = self._job
            j.n, j.stride, j.rec_len = n, stride, 42 if stride >= 42 else 41
            if n:
                j.d_packets = pk.data_ptr()
                j.d_agent_idx = agent_idx.data_ptr() if agent_idx is not None else None
                j.d_drift = drift.data_ptr() if drift is not None else None
                j.d_agent_off, j.n_agents = agent_table.data_ptr(), agent_table.shape[0] - 1
                j.ordinal_base = self.rank * self.ordinal_stride
                self._keep = (pk, agent_idx, drift, agent_table)
            job = self._job_ref
        # the real barrier (wait) runs on a side stream, overlapping the resolve pass of this step
        side = self._side.cuda_stream if wait else None
        rc = self._lib.occgrid_band_step(self._ctx_ref, self._step, 1 if self._pending else 0, job, 1 if wait else 0, self._stream(), side)
        if rc:
            _native.check(rc, 'occgrid_band_step')
        if self._pending:
            self.grid._host_cache = None
        self._pending = pk is not None      # every rank publishes every step, even an empty share, so the barrier lines up
        if pk is not None:
            self._step += 1

    def check_status(self):
        _native.check(self._lib.occgrid_band_join(self._ctx_ref, self._stream()), 'occgrid_band_join')
        st = int(self._status.item())
        if st:
            self._status.zero_()
            raise OccGridError('band step: ' + ', '.join(n for bit, n in ((2, 'a receive segment overflowed (records dropped)'),
                                                                          (4, 'the cross-GPU barrier timed out')) if st & bit))


class TiledSwarmMap:
    """A global occupancy grid spatially tiled over the ranks of a process group.  Mirrors
    ``OccupancyGrid``'s batched entry; each rank passes ITS share of the packet stream.

    exchange   'p2p' (default): the fused raycast + route kernel over NVLink peer memory; the map is
               complete after ``flush()`` (``gather_grid`` / ``counters`` flush first) because a
               batch is integrated while the NEXT one is being routed.  'nccl': route kernel +
               ``all_to_all_single`` + integrate, synchronous.  No silent switching: when symmetric
               memory is unavailable 'p2p' raises.
    ops        test seam of the 'nccl' path (tests/test_distributed_cpu.py drives the host logic on
               gloo with the oracle as the device)."""

    def __init__(self, size, resolution=0.05, origin_x=-5.0, origin_y=-5.0, *, group=None, device=None,
                 strategy='auto', max_batch=1 << 16, ops=None, exchange=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size, self.res, self.ox, self.oy = int(size), float(resolution), float(origin_x), float(origin_y)
        self.layout = BandLayout(size, self.world)
        if exchange is None:
            exchange = 'nccl' if ops is not None else 'p2p'
        if exchange not in ('p2p', 'nccl'):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange
        self.band = None
        self._recv_bufs = {}
        if exchange == 'p2p':
            if ops is not None:
                raise ValueError("the 'ops' test seam belongs to exchange='nccl'")
            if not torch.cuda.is_available():
                raise OccGridError('no CUDA device: TiledSwarmMap has no CPU fallback')
            device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
            self.band = BandStep(self.layout, self.rank, size, resolution, origin_x, origin_y, device, max_batch,
                                 symmetric=self.world > 1, group=group)
            if self.world == 1:
                BandBuffers.link([self.band.buf])
            self.band.finish_init()
            self.local = self.band.grid
            self.device = device
            self._tab_cache = None
            if self.world > 1:
                torch.cuda.synchronize(device)
                dist.barrier(group=group)
            log.info('TiledSwarmMap rank %d/%d: exchange=p2p (fused raycast+route over peer memory), %d-record segments',
                     self.rank, self.world, self.band.seg_cap)
            return
        if ops is None:
            if not torch.cuda.is_available():
                raise OccGridError('no CUDA device: TiledSwarmMap has no CPU fallback')
            device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
            ops = CudaBandOps(self.layout, self.rank, size, resolution, origin_x, origin_y, device, strategy, max_batch)
        self.ops = ops
        self.device = ops.device
        self.local = getattr(ops, 'grid', None)
        log.info('TiledSwarmMap rank %d/%d: exchange=nccl (send buffer + all_to_all_single)', self.rank, self.world)

    def _agent_table(self, separation, agent_offsets):
        if isinstance(agent_offsets, torch.Tensor):
            t = agent_offsets.to(dtype=torch.float64).reshape(-1, 2).contiguous()
        elif agent_offsets is None:
            t = torch.tensor([[0.0, 0.0], [0.0, 0.0], [float(separation), 0.0]], dtype=torch.float64)
        else:
            t = torch.from_numpy(np.ascontiguousarray(agent_offsets, np.float64).reshape(-1, 2))
        return t.to(self.device)

    # ---- exchange = 'nccl' -----------------------------------------------------------------
    def _exchange(self, send, counts, stride, dtype):
        """all_to_all_single of row-segments; returns the rows received (concatenated by source
        rank, i.e. in canonical stream order)."""
        if self.world == 1:
            return send
        rows = int(sum(self._recv_counts))
        key = (stride, dtype)
        buf = self._recv_bufs.get(key)
        if buf is None or buf.shape[0] < rows:          # grown geometrically, then reused every step
            buf = self.ops.empty(int(rows * 1.25) + 1024, stride, dtype)
            self._recv_bufs[key] = buf
        out = buf[:rows]
        dist.all_to_all_single(out, send.contiguous(), output_split_sizes=self._recv_counts, input_split_sizes=counts,
                               group=self.group)
        return out

    def _update_packets_nccl(self, packets, separation, drift, agent_offsets, agent_idx):
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
        recv = self._exchange(send, counts, send.shape[1], torch.uint8)
        self.ops.integrate(recv)
        return int(recv.shape[0])

    # ---- public entry ----------------------------------------------------------------------
    def update_packets(self, packets, separation=0.0, drift=None, agent_offsets=None, agent_idx=None):
        """Integrate this rank's share of the stream into the tiled map (ALL ranks must call, with
        an empty share if they have nothing).  exchange='p2p': asynchronous — this batch is routed
        while the previous one is integrated; ``flush()`` completes the map."""
        if self.band is None:
            return self._update_packets_nccl(packets, separation, drift, agent_offsets, agent_idx)
        dev = self.device
        T = torch.Tensor
        if type(packets) is T and packets.is_cuda and packets.dtype == torch.uint8 and packets.dim() == 2 and packets.is_contiguous() \
                and type(agent_offsets) is T and agent_offsets.is_cuda and agent_offsets.dtype == torch.float64 and agent_offsets.is_contiguous() \
                and (agent_idx is None or (type(agent_idx) is T and agent_idx.is_cuda and agent_idx.dtype == torch.int32 and agent_idx.is_contiguous())) \
                and (drift is None or (type(drift) is T and drift.is_cuda and drift.dtype == torch.float64 and drift.is_contiguous())):
            # everything already lives on the device in the wire layout: no conversions, one C call
            self.band.step(packets, agent_idx, drift, agent_offsets.view(-1, 2))
            return int(packets.shape[0])
        with torch.cuda.device(dev):
            tab = self._agent_table(separation, agent_offsets)
            pk = self.local.stage_packets(packets)[0]
            idx = torch.as_tensor(agent_idx, dtype=torch.int32).to(dev).contiguous() if agent_idx is not None else None
            dr = torch.as_tensor(drift, dtype=torch.float64).reshape(-1, 2).to(dev).contiguous() if drift is not None else None
            self.band.step(pk, idx, dr, tab)
        return int(pk.shape[0])

    def flush(self):
        """exchange='p2p': integrate the batch still in flight and surface overflow / barrier errors."""
        if self.band is not None:
            if self.band._pending:
                self.band.step(None, None, None, None)
            self.band.check_status()

    def counters(self, reset=False):
        self.flush()
        return self.local.counters(reset=reset)

    def gather_grid(self):
        """Assemble the global map on every rank (all_gather of the disjoint bands)."""
        self.flush()
        band = self.local.grid_tensor if self.band is not None else self.ops.band_tensor()
        if self.world == 1:
            return band.cpu().numpy()
        bands = [torch.empty((self.layout.window(b)[3], self.size), dtype=torch.int8, device=band.device) for b in range(self.world)]
        dist.all_gather(bands, band.contiguous(), group=self.group)
        return torch.cat(bands, dim=0).cpu().numpy()


# ----------------------------------------------------------------------------------------------
#  Merge: what shards, and what the reference's semantics keep sequential
# ----------------------------------------------------------------------------------------------

def agent_block(n_agents_total, world, rank):
    """Agents are dealt to ranks in contiguous blocks (rank order == agent order)."""
    return (n_agents_total * rank) // world, (n_agents_total * (rank + 1)) // world


class ShardedMapMerger:
    """Multi-GPU ``MapMerger.merge``: rank r holds the grids of the agents of ITS block
    (``agent_block``).  Two modes:

    ``mode='exact'`` (default) — results identical to the reference.  The reference's fuse is a
    sequential chain: callback a voxel-filters the WHOLE accumulated cloud, and a voxel's point is
    the unweighted mean of [the cloud's point, the slice's points] (map_merger.py:58-60), which is
    order-dependent and not associative, so the chain itself cannot be split across agents without
    changing the map.  What shards is the HBM-bound part: every rank scans, extracts and transforms
    its agents' grids in ONE batched launch pair (bulk-copy staged scan), the per-agent point lists
    are all-gathered (NCCL), and the ordered chain is replayed identically on every rank
    ("replicas only" for the chain).  == ``MapMerger.merge`` on one GPU, bit for bit.

    ``mode='raster_fuse'`` — the partitioning SURVEY §8e sketches: every rank runs the chain over
    its OWN agents only, the cloud bounds are all-reduced (min / max, map_merger.py:95-98), every
    rank rasterises its cloud into a partial int8 grid over the GLOBAL bounds (:100-111) and the
    partial grids are fused with ReduceScatter(max) + AllGather on int8 (values {-1, 100}: max ==
    "occupied wins", idempotent and order-free).  Scales with the ranks, but voxels that mix points
    of agents living on different ranks are averaged per rank instead of jointly, so a small
    fraction of cells differs from the reference's map — NOT a parity mode; callers get the count
    of differing cells from bench.py / tools/check_multi_gpu.py, never a silent substitution."""

    def __init__(self, group=None, device=None, mode='exact'):
        from .map_merger import MapMerger
        if mode not in ('exact', 'raster_fuse'):
            raise ValueError("mode must be 'exact' or 'raster_fuse'")
        self.mode = mode
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merger = MapMerger(device=device if device is not None else 'cuda')
        self.stats = {}

    # ---- collectives (plumbing) --------------------------------------------------------------
    def _all_gather_vec(self, local, total_len, lo):
        """Variable-length all-gather of a per-agent vector: everybody contributes its block at [lo, ...)."""
        full = torch.zeros(total_len, dtype=local.dtype, device=local.device)
        full[lo:lo + local.numel()] = local
        if self.world > 1:
            dist.all_reduce(full, group=self.group)
        return full

    def merge(self, local_grids, local_origins, res, local_transforms, n_agents_total, fitness=None, to_host=True):
        """``local_*``: the agents [lo, hi) = agent_block(n_agents_total, world, rank), in order.
        ``fitness``: None or the GLOBAL float array [n_agents_total] (gate at map_merger.py:54-56).
        Returns (int8 grid [H', W'], (origin_x, origin_y)) on every rank."""
        if self.mode == 'raster_fuse':
            return self._merge_raster_fuse(local_grids, local_origins, res, local_transforms, n_agents_total, fitness, to_host)
        from .map_merger import _Cloud
        m = self.merger
        dev = m.device
        A = int(n_agents_total)
        lo, hi = agent_block(A, self.world, self.rank)
        if len(local_grids) != hi - lo:
            raise ValueError(f'rank {self.rank} must hold agents [{lo}, {hi}) of {A}')
        with torch.cuda.device(dev):
            dgr = m._as_device_grids(local_grids)
            shape = dgr[0].shape if dgr else (0, 0)
            shp = torch.tensor([shape[0], shape[1], -shape[0], -shape[1]], dtype=torch.int64, device=dev)
            if self.world > 1:
                dist.all_reduce(shp, op=dist.ReduceOp.MAX, group=self.group)
            H, W = int(shp[0].item()), int(shp[1].item())
            if not all(d.shape == (H, W) for d in dgr) or (int(shp[2].item()), int(shp[3].item())) != (-H, -W):
                raise OccGridError('ShardedMapMerger: all agent grids must have the same shape')
            # pass 1 (sharded): occupied cells of my agents; everybody learns every count
            if dgr:
                counts, ptrs, bws = m._count_occupied(dgr, True)
            else:
                counts, ptrs, bws = torch.zeros(0, dtype=torch.int64, device=dev), None, None
            n_occ = self._all_gather_vec(counts, A, lo).cpu().numpy()
            meta = np.zeros((A, 18))                 # every rank plans with the GLOBAL callback table (tiny, host-side)
            if hi > lo:
                meta[lo:hi, :2] = np.asarray(local_origins, np.float64).reshape(hi - lo, 2)
                meta[lo:hi, 2:] = m._as_matrices(local_transforms, hi - lo).reshape(hi - lo, 16)
            if self.world > 1:
                md = torch.from_numpy(meta).to(dev)
                dist.all_reduce(md, group=self.group)       # disjoint rows: the sum is the concatenation, bit for bit
                meta = md.cpu().numpy()
            org, Tm = np.ascontiguousarray(meta[:, :2]), np.ascontiguousarray(meta[:, 2:]).reshape(A, 4, 4)
            fit = np.ones(A) if fitness is None else np.asarray(fitness, np.float64).reshape(A)
            hw = np.tile(np.array([[H, W]], np.float64), (A, 1))
            plan = m._plan_merge(n_occ, fit, Tm, org, hw, res)
            if plan is None:
                return None, None
            use, order, first, bb, n_new = plan
            # pass 2 (sharded): my used agents' points, transformed, in one launch
            n_mine = int(n_occ[lo:hi][use[lo:hi]].sum())
            if dgr:
                lstage, loffs = m._write_slices(dgr, ptrs, bws, counts, use[lo:hi], Tm[lo:hi], org[lo:hi], res, n_mine)
            # all-gather the slices: agent order == rank order, so the gathered stage is simply the concatenation
            stage = _Cloud(n_new + 16, dev)
            offs_h = np.zeros(A + 1, np.int64)
            offs_h[1:] = np.cumsum(np.where(use, n_occ, 0))
            per_rank = [int(offs_h[agent_block(A, self.world, r)[1]] - offs_h[agent_block(A, self.world, r)[0]]) for r in range(self.world)]
            if self.world > 1:
                mx = max(max(per_rank), 1)
                send = torch.zeros((2, mx), dtype=torch.float64, device=dev)
                if n_mine:
                    send[0, :n_mine] = lstage.x[:n_mine]
                    send[1, :n_mine] = lstage.y[:n_mine]
                recv = torch.empty((self.world, 2, mx), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(recv, send, group=self.group)
                for r in range(self.world):
                    b0 = int(offs_h[agent_block(A, self.world, r)[0]])
                    stage.x[b0:b0 + per_rank[r]] = recv[r, 0, :per_rank[r]]
                    stage.y[b0:b0 + per_rank[r]] = recv[r, 1, :per_rank[r]]
            elif n_mine:
                stage.x[:n_mine] = lstage.x[:n_mine]
                stage.y[:n_mine] = lstage.y[:n_mine]
            offs = torch.from_numpy(offs_h).to(dev)
            # the ordered voxel chain (:59-60), replayed identically on every rank
            m._ensure_capacity(m._n_global + n_new + 1024)
            v = float(res) if first else m.map_resolution
            lat_w, lat_h = int((bb[2] - bb[0]) / v) + 4, int((bb[3] - bb[1]) / v) + 4
            if first:                                   # adopted as is (:40-43)
                a0 = order.pop(0)
                rc = m._lib.mapmerge_append_slice(stage.x.data_ptr(), stage.y.data_ptr(), offs.data_ptr(), a0,
                                                  m._cloud.x.data_ptr(), m._cloud.y.data_ptr(), m._cloud.capacity,
                                                  m._cloud.count.data_ptr(), m._status.data_ptr(), None, m._stream())
                _native.check(rc, 'mapmerge_append_slice')
                m.map_resolution = float(res)
                m.map_origin = [float(org[a0, 0]), float(org[a0, 1])]
            m.chain_stats = m._run_chain(stage, offs, A, order, lat_w, lat_h, int(n_occ[use].max())) if order else None
            m._n_global = int(m._cloud.count.item())
            m._check_status()
        out = m.publish_global_map(to_host=to_host)
        return (out.data, (out.info.origin.position.x, out.info.origin.position.y)) if out is not None else (None, None)

    def _merge_raster_fuse(self, local_grids, local_origins, res, local_transforms, n_agents_total, fitness, to_host):
        """per-rank chain -> AllReduce(min/max) of the bounds -> partial raster over the
        global bounds -> ReduceScatter(max, int8) + AllGather."""
        from .map_merger import make_grid_msg
        m = self.merger
        dev = m.device
        A = int(n_agents_total)
        lo, hi = agent_block(A, self.world, self.rank)
        fit = None if fitness is None else np.asarray(fitness, np.float64).reshape(A)[lo:hi]
        with torch.cuda.device(dev):
            # my agents' own chain; only the rank that holds agent 0's block adopts its first cloud untransformed
            if hi > lo:
                m.merge(local_grids, local_origins, res, local_transforms, fitness=fit, publish=False, adopt_first=self.rank == 0)
            big = 1.0e300
            b = torch.tensor([big, big, big, big], dtype=torch.float64, device=dev)
            if m._n_global:
                m._bounds_of(m._cloud)
                b = torch.stack([m._bounds[0], m._bounds[1], -m._bounds[2], -m._bounds[3]])
            if self.world > 1:
                dist.all_reduce(b, op=dist.ReduceOp.MIN, group=self.group)          # min of (min_x, min_y, -max_x, -max_y)
            bh = b.cpu().numpy()
            if bh[0] >= big:
                return None, None
            min_x, min_y, max_x, max_y = float(bh[0]), float(bh[1]), float(-bh[2]), float(-bh[3])
            v = m.map_resolution if m._n_global else float(res)
            width = int(np.ceil((max_x - min_x) / v)) + 1                            # :100
            height = int(np.ceil((max_y - min_y) / v)) + 1                           # :101
            cells = width * height
            padded = -(-cells // (16 * self.world)) * (16 * self.world)
            part = torch.full((padded,), -1, dtype=torch.int8, device=dev)
            if m._n_global:
                gb = torch.tensor([min_x, min_y, max_x, max_y], dtype=torch.float64, device=dev)
                rc = m._lib.mapmerge_rasterise(m._cloud.x.data_ptr(), m._cloud.y.data_ptr(), m._cloud.count.data_ptr(), v,
                                               gb.data_ptr(), width, height, part.data_ptr(), m._stream())
                _native.check(rc, 'mapmerge_rasterise')
            if self.world > 1:
                shard = torch.empty(padded // self.world, dtype=torch.int8, device=dev)
                dist.reduce_scatter_tensor(shard, part, op=dist.ReduceOp.MAX, group=self.group)   # occupied wins
                dist.all_gather_into_tensor(part, shard, group=self.group)
            grid = part[:cells].view(height, width)
            data = grid.cpu().numpy() if to_host else grid
        m.published = make_grid_msg(data, width, height, v, min_x, min_y, frame_id='map_global')
        return data, (min_x, min_y)


# ----------------------------------------------------------------------------------------------
#  bench.py helper: weak-scaling sessions
# ----------------------------------------------------------------------------------------------

def make_rank_sessions(world, rank, device, packets_per_rank, pool, strategy='auto',
                       grid_per_gpu=4096, agents_per_gpu=64, exchange='p2p', ingest='uniform', layout='bands'):
    """Weak scaling of BASELINE configs[1] / configs[3]: (grid_per_gpu*world)^2 map cut into `world`
    row bands, agents_per_gpu*world agents, each rank ingests its own `packets_per_rank` share.
    ``ingest='uniform'`` (default, the worst case): every rank's share covers ALL agents, so
    (world-1)/world of the records travel to another GPU; ``'affine'``: a rank receives the agents
    that drive inside its own band (only rays crossing a band edge travel).
    Returns (TiledSwarmMap, sessions, step_fn)."""
    from . import simulation_tools as st
    side = grid_per_gpu * world
    origin = (-side * 0.05 / 2.0,) * 2
    tmap = TiledSwarmMap(side, 0.05, origin[0], origin[1], device=device, strategy=strategy,
                         max_batch=int(packets_per_rank), exchange=exchange)
    sessions = []
    for i in range(pool):
        if ingest == 'affine':
            # this rank's own agents, their rooms on a lattice inside the band it owns
            y0 = tmap.layout.window(rank)[1]
            sh = st.generate_session(n_agents=agents_per_gpu, n_packets=packets_per_rank, grid_size=grid_per_gpu,
                                     origin=(origin[0], origin[1] + y0 * 0.05),
                                     seed=1000 + 97 * i + rank)
        else:
            # layout='bands': the same number of rooms in every band (weak scaling: same work per GPU);
            # 'lattice': one square lattice over the whole map, whose rows fall unevenly into the bands
            sh = st.generate_session(n_agents=agents_per_gpu * world, n_packets=packets_per_rank, grid_size=side,
                                     origin=origin, seed=1000 + 97 * i + rank, bands=world if layout == 'bands' else 1)
        sessions.append({'packets': tmap.local.stage_packets(sh['packets'])[0],
                         'agent_idx': torch.from_numpy(sh['agent_idx']).to(device),
                         'agent_offsets': torch.from_numpy(sh['agent_offsets']).to(device), 'grid': sh['grid'],
                         'host': sh})

    def step(i):
        s = sessions[i % pool]
        tmap.update_packets(s['packets'], agent_offsets=s['agent_offsets'], agent_idx=s['agent_idx'])

    return tmap, sessions, step
