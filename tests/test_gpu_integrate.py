"""GPU parity tests of the integration path: the CUDA kernels (through the C ABI, via the
reference-shaped ``OccupancyGrid``) against the oracle and the golden fixtures produced by
the unmodified reference.  Integer/byte work: the bar is BIT-EXACT grids."""
import hashlib
import os
import struct

import numpy as np
import pytest

from conftest import GOLD, load_packet_stream, normalise_datagrams, session_packets

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

STRATEGIES = ['global_atomic', 'tiled']


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope='module')
def M():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import dual_bot_mapper
    return dual_bot_mapper


def supported(M, strategy, **kw):
    try:
        return M.OccupancyGrid(strategy=strategy, **kw)
    except M.OccGridError as e:
        pytest.skip(f'{strategy}: {e}')


@pytest.mark.parametrize('strategy', STRATEGIES)
@pytest.mark.parametrize('order', ['file', 'time'])
@pytest.mark.parametrize('slam', [True, False])
def test_golden_session(M, golden, strategy, order, slam):
    """BASELINE config 1: the seeded 2-bot session replayed into the default 200^2 grid must
    reproduce the reference's census and sha1."""
    pk, _ = session_packets(time_sorted=(order == 'time'))
    want = golden['session'][f'{order}_order_slam_{"on" if slam else "off"}']
    g = supported(M, strategy)
    if slam:
        drift, s = M.slam_drift_table(pk, separation=0.0)
        assert len(s.closures) == want['closures']
        assert np.array_equal(drift, np.load(os.path.join(GOLD, f'session_drift_{order}.npy')))
    else:
        drift = None
    g.update_packets(pk, separation=0.0, drift=drift)
    grid = g.grid
    assert grid.dtype == np.int8 and grid.shape == (200, 200)
    assert (int((grid == 0).sum()), int((grid == 100).sum()), int((grid == -1).sum())) == \
        (want['free'], want['occ'], want['unk'])
    assert sha1(grid) == want['sha1']
    c = g.counters()
    assert c['packets'] == 687 and c['beams'] == 2748
    if not slam:
        assert c['updates'] == 54049 and c['hits'] == 1041


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_golden_session_separation_and_replay_helper(M, golden, strategy):
    pk, _ = session_packets(time_sorted=True)
    g = supported(M, strategy)
    g.update_packets(pk, separation=-2.5)
    assert sha1(g.grid) == golden['session']['time_order_slam_off_sep_-2.5']['sha1']
    g2, slam = M.replay_session(pk, separation=0.0, use_slam=True, strategy=strategy)
    assert sha1(g2.grid) == golden['session']['time_order_slam_on']['sha1']
    assert len(slam.closures) == 10


@pytest.mark.parametrize('strategy', STRATEGIES)
@pytest.mark.parametrize('name', ['mixed_a', 'mixed_b_sep', 'mixed_c_4096', 'mixed_d_edge'])
def test_adversarial_packet_streams(M, golden, strategy, name):
    """Bad magic, foreign agents, v1/junk datagram sizes, NaN/inf/sentinel ranges, float32 edge
    values of the trust filter, quantised yaws on cell boundaries — vs the reference's result."""
    want = golden['packet_streams'][name]
    pk, ref_drift = load_packet_stream(name)
    g = supported(M, strategy, **want['grid_kwargs'])
    drift = ref_drift if want['slam'] else None
    if want['slam']:
        d2, s = M.slam_drift_table(pk, separation=want['separation'])
        assert np.array_equal(d2, ref_drift)
        assert len(s.closures) == want['closures']
    g.update_packets(pk, separation=want['separation'], drift=drift)
    assert sha1(g.grid) == want['sha1']


@pytest.mark.parametrize('name', ['default200', 'g512_r0.1', 'g64_r0.02'])
def test_update_rays_fixture(M, golden, name):
    z = np.load(os.path.join(GOLD, f'rays_{name}.npz'))
    want = golden['rays'][name]
    g = M.OccupancyGrid(**want['grid_kwargs'])
    rays = np.stack([z['x0'], z['y0'], z['x1'], z['y1']], axis=1)
    g.update_rays(rays, z['hit'])
    assert sha1(g.grid) == want['sha1']


def test_update_ray_single_calls(M, golden):
    """The reference's per-ray API, one call per ray (first 200 rays of a fixture)."""
    from oracle import occgrid_oracle as O
    z = np.load(os.path.join(GOLD, 'rays_default200.npz'))
    g = M.OccupancyGrid()
    o = O.OracleGrid()
    for i in range(200):
        a = (float(z['x0'][i]), float(z['y0'][i]), float(z['x1'][i]), float(z['y1'][i]), bool(z['hit'][i]))
        g.update_ray(*a)
        o.update_ray(*a)
    assert np.array_equal(g.grid, o.grid)
    assert g.world_to_grid(-5.01, -5.01) == (0, 0) and g.world_to_grid(0.0, 0.0) == (100, 100)
    assert g._bresenham(0, 0, 3, -7) == O.OracleGrid.bresenham(0, 0, 3, -7)
    assert g.in_bounds(199, 0) and not g.in_bounds(200, 0)


def test_bresenham_exhaustive_on_device(M):
    """All (dx,dy) in [-32,32]^2 line shapes, each drawn in its own 80x80 block of one big grid,
    against the reference's `_bresenham` cell lists (golden KAT)."""
    z = np.load(os.path.join(GOLD, 'bresenham_kat.npz'))
    xs, ys, offs = z['x'].astype(np.int64), z['y'].astype(np.int64), z['offsets']
    B, n = 80, 65
    size = B * n
    res, ox, oy = 0.05, -3.0, 7.0
    want = np.full((size, size), -1, np.int8)
    rays = np.empty((n * n, 4))
    hit = np.zeros(n * n, np.uint8)
    i = 0
    for j, dy in enumerate(range(-32, 33)):
        for k, dx in enumerate(range(-32, 33)):
            cx, cy = k * B + 40, j * B + 40
            cells_x = xs[offs[i]:offs[i + 1]] + cx
            cells_y = ys[offs[i]:offs[i + 1]] + cy
            h = (i % 3) != 0
            want[cells_y[:-1], cells_x[:-1]] = 0
            if h:
                want[cells_y[-1], cells_x[-1]] = 100
            hit[i] = h
            rays[i] = (ox + (cx + 0.5) * res, oy + (cy + 0.5) * res,
                       ox + (cx + dx + 0.5) * res, oy + (cy + dy + 0.5) * res)
            i += 1
    g = M.OccupancyGrid(size=size, resolution=res, origin_x=ox, origin_y=oy)
    g.update_rays(rays, hit)
    assert np.array_equal(g.grid, want)


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_all_line_shapes_through_packets(M, strategy):
    """Every integer line shape a 1.2 m beam can take at 5 cm (|d| <= 24 cells, all octants and
    all tie cases) is produced by a packet (front beam aimed from a cell centre at another cell
    centre; the other three beams come along) and drawn through the packet path of each
    strategy; the grid must equal the oracle's."""
    import math
    from oracle import c_oracle
    res, size = 0.05, 4096
    ox = oy = -102.4
    pk = []
    slot = 0
    for dy in range(-24, 25):
        for dx in range(-24, 25):
            r = math.hypot(dx, dy) * res
            if r > 1.19 or (dx == 0 and dy == 0):
                continue
            cx, cy = 40 + 60 * (slot % 66), 40 + 60 * (slot // 66)
            slot += 1
            x, y = ox + (cx + 0.5) * res, oy + (cy + 0.5) * res
            yaw = math.atan2(dy, dx)
            for hit in (True, False):
                d = r if hit else 3.0
                pk.append(struct.pack(M.PACKET_FMT, b'QSRL', 1, x, y, yaw, 0, 0, d, 0.3 + 0.01 * (slot % 50), 0.0, 1.0, 0))
    arr, _ = normalise_datagrams(pk)
    g = supported(M, strategy, size=size, resolution=res, origin_x=ox, origin_y=oy)
    g.update_packets(arr)
    want = np.full((size, size), -1, np.int8)
    c = c_oracle.integrate_packets(arr, want, ox, oy, res)
    assert np.array_equal(g.grid, want)
    assert g.counters()['updates'] == c['updates'] and c['packets'] > 3000


def synthetic(n_packets, seed=3, n_agents=64):
    from occgrid_b200 import simulation_tools as st
    return st.generate_session(n_agents=n_agents, n_packets=n_packets, seed=seed)


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_synthetic_64_agents_vs_c_oracle(M, strategy):
    """BASELINE config 2 geometry (64 agents, 4096^2 @ 5 cm) at 2e5 packets: grid and counters
    bit-exact vs the C oracle; agent offset table exercised."""
    from oracle import c_oracle
    s = synthetic(200_000)
    g = supported(M, strategy, max_batch=200_000, **s['grid'])
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'])
    want = np.full((4096, 4096), -1, np.int8)
    c = c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'])
    got = g.grid
    assert np.array_equal(got, want)
    k = g.counters()
    for name in ('packets', 'accepted', 'dropped', 'bad_pose', 'beams', 'hits', 'updates'):
        assert k[name] == c[name], name
    assert k['owned_updates'] == k['updates']


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_agent_idx_override_and_drift(M, strategy):
    from oracle import c_oracle
    s = synthetic(50_000, seed=5)
    rng = np.random.default_rng(0)
    drift = rng.normal(0, 0.05, (50_000, 2))
    idx = s['agent_idx'].copy()
    idx[::97] = 0          # dropped
    idx[::89] = 9999       # dropped
    g = supported(M, strategy, max_batch=50_000, **s['grid'])
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'], agent_idx=idx, drift=drift)
    want = np.full((4096, 4096), -1, np.int8)
    c = c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'],
                                   agent_idx=idx, drift=drift)
    assert np.array_equal(g.grid, want)
    assert g.counters()['dropped'] == c['dropped'] > 0


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_batches_later_wins(M, strategy):
    """Streaming: three consecutive update_packets calls == one oracle pass over the
    concatenation (a later batch always overwrites an earlier one)."""
    from oracle import c_oracle
    s = synthetic(90_000, seed=8)
    g = supported(M, strategy, max_batch=30_000, **s['grid'])
    for i in range(3):
        g.update_packets(s['packets'][i * 30_000:(i + 1) * 30_000], agent_offsets=s['agent_offsets'])
    want = np.full((4096, 4096), -1, np.int8)
    c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'])
    assert np.array_equal(g.grid, want)


@pytest.mark.parametrize('pinned', [False, True])
def test_streamed_host_batch(M, pinned):
    """Large host buffers are copied chunk by chunk on a side stream while earlier chunks
    integrate; result and order semantics must not change (drift and agent_idx are sliced too)."""
    from oracle import c_oracle
    s = synthetic(70_000, seed=13)
    drift = np.random.default_rng(2).normal(0, 0.03, (70_000, 2))
    g = supported(M, 'auto', max_batch=8_000, **s['grid'])
    g.h2d_chunk = 8_000                      # 9 chunks through 2 device slots
    pk = torch.from_numpy(s['packets'])
    if pinned:
        pk = pk.pin_memory()
    g.update_packets(pk, agent_offsets=s['agent_offsets'], agent_idx=s['agent_idx'], drift=drift)
    g.update_packets(s['packets'][:20_000], agent_offsets=s['agent_offsets'], agent_idx=s['agent_idx'][:20_000])
    want = np.full((4096, 4096), -1, np.int8)
    c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'],
                               agent_idx=s['agent_idx'], drift=drift)
    c_oracle.integrate_packets(s['packets'][:20_000], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'],
                               agent_idx=s['agent_idx'][:20_000])
    assert np.array_equal(g.grid, want)


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_window_tiles_equal_crop(M, strategy):
    """Spatial tiles (multi-GPU layout run on one GPU): integrating the same stream into
    row-band windows reproduces the crops of the single-grid result."""
    from oracle import c_oracle
    pk, _ = load_packet_stream('mixed_a')
    arr, _ = normalise_datagrams(pk)
    full = np.full((200, 200), -1, np.int8)
    c_oracle.integrate_packets(arr, full, -5.0, -5.0, 0.05)
    owned = 0
    for (x0, y0, w, h) in [(0, 0, 200, 64), (0, 64, 200, 64), (0, 128, 200, 72)]:
        g = supported(M, strategy, window=(x0, y0, w, h))
        g.update_packets(arr)
        assert np.array_equal(g.grid, full[y0:y0 + h, x0:x0 + w])
        owned += g.counters()['owned_updates']
    g = supported(M, strategy, window=(37, 51, 64, 80))
    g.update_packets(arr)
    assert np.array_equal(g.grid, full[51:131, 37:101])


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_v1_records_and_stride(M, strategy):
    """41-byte v1 records packed back to back (stride 41) and 42-byte records in 48-byte slots."""
    from oracle import c_oracle
    s = synthetic(20_000, seed=11)
    want = np.full((4096, 4096), -1, np.int8)
    c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'])
    v1 = np.ascontiguousarray(s['packets'][:, :41])
    g = supported(M, strategy, max_batch=20_000, **s['grid'])
    g.update_packets(v1, agent_offsets=s['agent_offsets'], rec_len=41)
    assert np.array_equal(g.grid, want)
    padded = np.zeros((20_000, 48), np.uint8)
    padded[:, :42] = s['packets']
    g2 = supported(M, strategy, max_batch=20_000, **s['grid'])
    g2.update_packets(padded, agent_offsets=s['agent_offsets'])
    assert np.array_equal(g2.grid, want)


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_nan_pose_and_empty(M, strategy):
    good = struct.pack(M.PACKET_FMT, b'QSRL', 1, 0.0, 0.0, 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    bad = struct.pack(M.PACKET_FMT, b'QSRL', 1, float('nan'), 0.0, 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    inf = struct.pack(M.PACKET_FMT, b'QSRL', 2, 0.0, float('inf'), 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    far = struct.pack(M.PACKET_FMT, b'QSRL', 1, 3.0e30, -1.0e25, 1.0e20, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    from oracle import occgrid_oracle as O
    o, _ = O.replay([good])
    g = supported(M, strategy)
    g.update_packets([good, bad, inf, far, b'', b'xx'])
    assert np.array_equal(g.grid, o.grid)
    c = g.counters()
    assert c['bad_pose'] == 2 and c['packets'] == 4
    g.update_packets([])
    g.update_packets(np.zeros((0, 42), np.uint8))
    assert np.array_equal(g.grid, o.grid)


def test_errors_are_loud(M):
    g = M.OccupancyGrid()
    with pytest.raises((M.OccGridError, ValueError, TypeError)):
        g.update_packets(np.zeros(43, np.uint8))
    with pytest.raises(M.OccGridError):
        g.update_packets(np.zeros((4, 42), np.uint8), rec_len=40)
    with pytest.raises(M.OccGridError):
        M.OccupancyGrid(window=(0, 0, 300, 10))
    with pytest.raises(M.OccGridError):
        M.OccupancyGrid(device='cpu')


@pytest.mark.parametrize('strategy', STRATEGIES)
def test_full_size_config2_batch(M, strategy):
    """BASELINE config 2 at full size: 2.5e6 packets = 1e7 beams into 4096^2 — bit-exact vs the
    C oracle, plus size-independent properties (idempotence of a repeated batch; a checksum of
    tile checksums equal to the whole-grid checksum of the oracle)."""
    from oracle import c_oracle
    s = synthetic(2_500_000, seed=42)
    g = supported(M, strategy, max_batch=2_500_000, **s['grid'])
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'])
    want = np.full((4096, 4096), -1, np.int8)
    c = c_oracle.integrate_packets(s['packets'], want, -102.4, -102.4, 0.05, agent_offsets=s['agent_offsets'])
    got = g.grid.copy()
    assert np.array_equal(got, want)
    k = g.counters(reset=True)
    assert k['updates'] == c['updates'] and k['beams'] == 10_000_000
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'])     # idempotent
    assert np.array_equal(g.grid, want)
    tiles = [sha1(got[y:y + 512, x:x + 512]) for y in range(0, 4096, 512) for x in range(0, 4096, 512)]
    tiles_want = [sha1(want[y:y + 512, x:x + 512]) for y in range(0, 4096, 512) for x in range(0, 4096, 512)]
    assert hashlib.sha1(''.join(tiles).encode()).hexdigest() == hashlib.sha1(''.join(tiles_want).encode()).hexdigest()


def test_fine_resolution_falls_back_to_global_atomics(M):
    """At 2 cm cells a beam reaches 62 cells and the (64+2R)^2 shared-memory window no longer fits
    the budget: 'tiled' is refused loudly, 'auto' picks the global-atomic kernels; result = oracle."""
    from oracle import c_oracle
    rng = np.random.default_rng(4)
    n = 4000
    pk = np.zeros((n, 42), np.uint8)
    rec = pk.view(np.dtype([('magic', 'S4'), ('agent', 'u1'), ('x', '<f4'), ('y', '<f4'), ('yaw', '<f4'), ('enc', '<i4'),
                            ('v2v', '<u4'), ('d', '<f4', 4), ('lm', 'u1')]))[:, 0]
    rec['magic'] = b'QSRL'
    rec['agent'] = rng.integers(1, 3, n)
    rec['x'] = rng.uniform(-1, 9, n)
    rec['y'] = rng.uniform(-1, 9, n)
    rec['yaw'] = rng.uniform(-3.2, 3.2, n)
    rec['d'] = rng.uniform(0, 1.5, (n, 4))
    with pytest.raises(M.OccGridError):
        M.OccupancyGrid(size=400, resolution=0.02, origin_x=0.0, origin_y=0.0, strategy='tiled')
    g = M.OccupancyGrid(size=400, resolution=0.02, origin_x=0.0, origin_y=0.0, strategy='auto', max_batch=n)
    g.update_packets(pk, separation=0.25)
    want = np.full((400, 400), -1, np.int8)
    c = c_oracle.integrate_packets(pk, want, 0.0, 0.0, 0.02, separation=0.25)
    assert np.array_equal(g.grid, want)
    assert g.counters()['updates'] == c['updates']


@pytest.mark.parametrize('frame', [1, 20, 137])
def test_ingest_batcher_frames_equal_one_batch(M, golden, frame):
    """The server loop's shape (:797-919): datagrams pushed one at a time, one flush per frame
    (the reference caps a frame at 20 datagrams, :816).  SLAM state and drift carry over between
    frames, later frames overwrite earlier ones: the final grid is the golden one, with SLAM on."""
    pk, _ = session_packets(time_sorted=True)
    want = golden['session']['time_order_slam_on']
    g = supported(M, 'auto')
    b = M.IngestBatcher(g, separation=0.0, use_slam=True, capacity=256)
    junk = 0
    for k, d in enumerate(pk):
        b.push(d)
        if k % 50 == 7:
            b.push(b'QSRL' + bytes(13))            # heartbeat-sized datagram: dropped by size (:836-838)
            junk += 1
        if (k + 1) % frame == 0:
            b.flush()
    b.flush()
    assert b.total == len(pk) and b.dropped == junk and b.flush() == 0
    assert sha1(g.grid) == want['sha1']
    assert len(b.slam.closures) == want['closures']


def test_ingest_batcher_v1_datagrams(M):
    """41-byte v1 datagrams carry no landmark byte: LM_NONE (:832-835)."""
    from oracle import occgrid_oracle as O
    pk, _ = session_packets(time_sorted=True)
    v1 = [d[:41] if i % 3 else d for i, d in enumerate(pk)]
    og, os_ = O.OracleGrid(), O.OracleSLAM()
    O.replay(v1, grid=og, separation=0.3, slam=os_)
    g = supported(M, 'auto')
    b = M.IngestBatcher(g, separation=0.3, use_slam=True, capacity=100)
    for d in v1:
        b.push(d)
    b.flush()
    assert np.array_equal(g.grid, og.grid)
    assert b.slam.closures == [tuple(c) for c in os_.closures]


def test_pageable_batches_back_to_back_do_not_share_the_staging_buffer():
    """update_packets on pageable host input goes through ONE reusable pinned staging buffer and an
    asynchronous H2D copy; the next call must not overwrite the buffer while that copy is still
    queued behind earlier kernels (it would integrate the later batch twice)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import dual_bot_mapper as M, simulation_tools as st
    from oracle import c_oracle
    s = st.generate_session(n_agents=8, n_packets=300_000, grid_size=1024, origin=(-25.6, -25.6), seed=11)
    g = M.OccupancyGrid(max_batch=300_000, **s['grid'])
    big = g.stage_packets(s['packets'])[0]
    want = np.full((1024, 1024), -1, np.int8)
    a, b = np.ascontiguousarray(s['packets'][:1500]), np.ascontiguousarray(s['packets'][150_000:151_500])
    for _ in range(6):                                   # keep the stream busy so the small copies queue up behind kernels
        g._integrate_device(big, None, 0.0, None, s['agent_offsets'], None, 42)
        c_oracle.integrate_packets(s['packets'], want, -25.6, -25.6, 0.05, agent_offsets=s['agent_offsets'])
    g.update_packets(a, agent_offsets=s['agent_offsets'])            # pageable numpy -> staging buffer
    g.update_packets(b, agent_offsets=s['agent_offsets'])            # reuses the same staging buffer right away
    c_oracle.integrate_packets(a, want, -25.6, -25.6, 0.05, agent_offsets=s['agent_offsets'])
    c_oracle.integrate_packets(b, want, -25.6, -25.6, 0.05, agent_offsets=s['agent_offsets'])
    assert np.array_equal(g.grid, want)
