"""world_size-2 `gloo` test of the multi-GPU plumbing on the CPU (no GPU in this container).

The product's device side (CudaBandOps: routing kernel + windowed OccupancyGrid) is swapped for
a test double built on the ORACLE, so what is exercised here is the host logic that has to be
right for N > 1: the band layout, the split sizes of the all-to-all, the canonical ordering
argument (rank-blocked stream + stable routing => last-writer-wins equals one grid), and the
all-gather that assembles the map."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleBandOps:
    """CPU stand-in for CudaBandOps (same interface)."""

    def __init__(self, layout, rank, size, res, ox, oy):
        self.layout, self.rank = layout, rank
        self.size, self.res, self.ox, self.oy = size, res, ox, oy
        x0, y0, w, h = layout.window(rank)
        self.win = (x0, y0)
        self.band = np.full((h, w), -1, np.int8)
        self.device = torch.device('cpu')

    def stage(self, packets):
        return torch.from_numpy(np.ascontiguousarray(packets, np.uint8))

    def route(self, packets, agent_idx, drift, agent_table):
        pk = packets.numpy()
        n = pk.shape[0]
        tab = agent_table.numpy()
        ids = agent_idx.numpy() if agent_idx is not None else pk[:, 4].astype(np.int64)
        ok = (pk[:, 0] == ord('Q')) & (pk[:, 1] == ord('S')) & (pk[:, 2] == ord('R')) & (pk[:, 3] == ord('L'))
        ok &= (ids >= 1) & (ids <= tab.shape[0] - 1)
        y = np.ascontiguousarray(pk[:, 9:13]).view('<f4')[:, 0].astype(np.float64)
        x = np.ascontiguousarray(pk[:, 5:9]).view('<f4')[:, 0].astype(np.float64)
        yaw = np.ascontiguousarray(pk[:, 13:17]).view('<f4')[:, 0].astype(np.float64)
        safe = np.where(ok, ids, 1)
        x = x + tab[safe, 0]
        y = y + tab[safe, 1]
        if drift is not None:
            x = x + drift.numpy()[:, 0]
            y = y + drift.numpy()[:, 1]
        ok &= np.isfinite(x) & np.isfinite(y) & np.isfinite(yaw)
        gy = np.trunc((np.where(ok, y, 0.0) - self.oy) / self.res).astype(np.int64)
        reach = int(math.ceil(1.2 / self.res)) + 2
        segs, counts = [], []
        for b in range(self.layout.n_bands):
            sel = ok & (gy + reach >= self.layout.band_y0[b]) & (gy - reach < self.layout.band_y0[b + 1])
            segs.append(np.nonzero(sel)[0])
            counts.append(int(sel.sum()))
        order = np.concatenate(segs) if segs else np.zeros(0, np.int64)
        # opaque 64-byte rows: packet (42) | agent index (4) | drift (16) | pad
        rows = np.zeros((order.shape[0], 64), np.uint8)
        rows[:, :42] = pk[order][:, :42]
        rows[:, 42:46] = ids[order].astype('<i4').view(np.uint8).reshape(-1, 4)
        d = drift.numpy()[order] if drift is not None else np.zeros((order.shape[0], 2))
        rows[:, 46:62] = np.ascontiguousarray(d, '<f8').view(np.uint8).reshape(-1, 16)
        self._tab = tab
        return torch.from_numpy(rows), counts

    def empty(self, rows, stride, dtype):
        return torch.empty((rows, stride) if stride else (rows,), dtype=dtype)

    def integrate(self, rows):
        from oracle import c_oracle
        r = rows.numpy()
        if r.shape[0] == 0:
            return
        ids = np.ascontiguousarray(r[:, 42:46]).view('<i4')[:, 0]
        d = np.ascontiguousarray(r[:, 46:62]).view('<f8').reshape(-1, 2)
        c_oracle.integrate_packets(np.ascontiguousarray(r[:, :42]), self.band, self.ox, self.oy, self.res,
                                   agent_offsets=self._tab, agent_idx=ids, drift=d,
                                   window=self.win, size_x=self.size, size_y=self.size)

    def band_tensor(self):
        return torch.from_numpy(self.band)


def _worker(rank, world, port, tmpdir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from occgrid_b200 import simulation_tools as st
    from occgrid_b200.distributed import BandLayout, TiledSwarmMap
    size = 512
    origin = (-12.8, -12.8)
    sess = st.generate_session(n_agents=4, n_packets=8000, grid_size=size, origin=origin, seed=3)
    # put the rooms near the band boundary so that both bands see traffic and some packets go to both
    offs = sess['agent_offsets'].copy()
    offs[1:3] = (-3.0, -1.5)
    offs[3:5] = (-2.0, 1.0)
    n = sess['packets'].shape[0]
    share = slice(rank * n // world, (rank + 1) * n // world)
    rng = np.random.default_rng(5)
    drift = rng.normal(0, 0.02, (n, 2))
    layout = BandLayout(size, world)
    ops = OracleBandOps(layout, rank, size, 0.05, origin[0], origin[1])
    tmap = TiledSwarmMap(size, 0.05, origin[0], origin[1], ops=ops)
    got_rows = 0
    for half in range(2):          # two consecutive batches: later batch wins
        lo = share.start + half * (share.stop - share.start) // 2
        hi = share.start + (half + 1) * (share.stop - share.start) // 2
        got_rows += tmap.update_packets(sess['packets'][lo:hi], agent_offsets=offs,
                                        agent_idx=sess['agent_idx'][lo:hi], drift=drift[lo:hi])
    full = tmap.gather_grid()
    np.save(os.path.join(tmpdir, f'grid_{rank}.npy'), full)
    np.save(os.path.join(tmpdir, f'rows_{rank}.npy'), np.array([got_rows]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_tiled_map_two_ranks_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from occgrid_b200 import simulation_tools as st
    from oracle import c_oracle
    size, origin = 512, (-12.8, -12.8)
    sess = st.generate_session(n_agents=4, n_packets=8000, grid_size=size, origin=origin, seed=3)
    offs = sess['agent_offsets'].copy()
    offs[1:3] = (-3.0, -1.5)
    offs[3:5] = (-2.0, 1.0)
    n = sess['packets'].shape[0]
    drift = np.random.default_rng(5).normal(0, 0.02, (n, 2))
    # canonical stream: batch 0 = [rank0 first half, rank1 first half], batch 1 = second halves
    idx = []
    for half in range(2):
        for r in range(world):
            s0, s1 = r * n // world, (r + 1) * n // world
            idx.append(np.arange(s0 + half * (s1 - s0) // 2, s0 + (half + 1) * (s1 - s0) // 2))
    idx = np.concatenate(idx)
    want = np.full((size, size), -1, np.int8)
    c_oracle.integrate_packets(sess['packets'][idx], want, origin[0], origin[1], 0.05, agent_offsets=offs,
                               agent_idx=sess['agent_idx'][idx], drift=drift[idx])
    g0 = np.load(tmp_path / 'grid_0.npy')
    g1 = np.load(tmp_path / 'grid_1.npy')
    assert np.array_equal(g0, g1)
    assert np.array_equal(g0, want)
    rows = int(np.load(tmp_path / 'rows_0.npy')[0] + np.load(tmp_path / 'rows_1.npy')[0])
    assert n <= rows <= 2 * n and rows > n          # some packets straddle the boundary and go to both bands
    assert (want[:256] != -1).any() and (want[256:] != -1).any()


def test_band_layout():
    from occgrid_b200.distributed import BandLayout
    L = BandLayout(4096, 8)
    assert L.band_y0 == [512 * b for b in range(9)]
    assert L.window(3) == (0, 1536, 4096, 512)
    L = BandLayout(200, 3)
    assert L.band_y0 == [0, 66, 133, 200]
    assert sum(L.window(b)[3] for b in range(3)) == 200
    assert L.bands_of_row_interval(60, 70) == [0, 1]
    assert L.bands_of_row_interval(-30, -1) == []
    assert L.bands_of_row_interval(199, 260) == [2]
    with pytest.raises(ValueError):
        BandLayout(4, 8)
