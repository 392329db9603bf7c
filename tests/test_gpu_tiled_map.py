"""Multi-GPU integration path emulated on ONE GPU (SURVEY §4 'Multi-GPU without a cluster'):
N logical ranks each route their share with the CUDA routing kernel, the all-to-all is done by
concatenating segments in source-rank order, every band integrates what it receives, and the
assembled map must equal the untiled oracle result bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.mark.parametrize('world', [2, 4])
@pytest.mark.parametrize('strategy', ['global_atomic', 'tiled'])
def test_emulated_ranks_equal_single_grid(world, strategy):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import simulation_tools as st
    from occgrid_b200.distributed import BandLayout, CudaBandOps
    from occgrid_b200.dual_bot_mapper import OccGridError
    from oracle import c_oracle
    size, origin = 1024, (-25.6, -25.6)
    sess = st.generate_session(n_agents=16, n_packets=60_000, grid_size=size, origin=origin, seed=9)
    n = sess['packets'].shape[0]
    offs = sess['agent_offsets'].copy()
    # park two rooms right on band boundaries so that packets go to two bands
    offs[1:3] = (-2.0, origin[1] + size * 0.05 / world - 1.0)
    offs[3:5] = (3.0, origin[1] + size * 0.05 / 2 + 0.2)
    drift = np.random.default_rng(1).normal(0, 0.02, (n, 2))
    layout = BandLayout(size, world)
    try:
        ops = [CudaBandOps(layout, r, size, 0.05, origin[0], origin[1], 'cuda', strategy, n) for r in range(world)]
    except OccGridError as e:
        pytest.skip(str(e))
    tab = torch.from_numpy(offs).cuda()
    routed = []
    for r in range(world):
        sl = slice(r * n // world, (r + 1) * n // world)
        pk = ops[r].stage(sess['packets'][sl])
        idx = torch.from_numpy(sess['agent_idx'][sl].copy()).cuda()
        dr = torch.from_numpy(drift[sl].copy()).cuda()
        send, counts = ops[r].route(pk, idx, dr, tab)
        offs_c = np.concatenate([[0], np.cumsum(counts)])
        routed.append([send[offs_c[b]:offs_c[b + 1]].clone() for b in range(world)])
    total_rows = 0
    bands = []
    for b in range(world):
        recv = torch.cat([routed[r][b] for r in range(world)])
        total_rows += recv.shape[0]
        ops[b].integrate(recv)
        bands.append(ops[b].band_tensor().cpu().numpy())
    got = np.concatenate(bands, axis=0)
    want = np.full((size, size), -1, np.int8)
    c = c_oracle.integrate_packets(sess['packets'], want, origin[0], origin[1], 0.05, agent_offsets=offs,
                                   agent_idx=sess['agent_idx'], drift=drift)
    assert np.array_equal(got, want)
    assert n < total_rows < 1.5 * n                      # boundary rooms are duplicated, the rest is not
    owned = sum(ops[b].grid.counters()['owned_updates'] for b in range(world))
    assert owned == c['updates']                         # every beam counted exactly once across bands


def test_pipelined_swarm_map_single_rank():
    """pipeline=True (route/exchange of batch i+1 on a side stream while batch i integrates):
    same map as the oracle after flush(); exercised here with world = 1."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import simulation_tools as st
    from occgrid_b200.distributed import TiledSwarmMap
    from oracle import c_oracle
    size, origin = 1024, (-25.6, -25.6)
    sess = st.generate_session(n_agents=16, n_packets=90_000, grid_size=size, origin=origin, seed=21)
    tmap = TiledSwarmMap(size, 0.05, origin[0], origin[1], max_batch=30_000, pipeline=True)
    assert tmap.pipeline
    for i in range(3):
        sl = slice(i * 30_000, (i + 1) * 30_000)
        tmap.update_packets(sess['packets'][sl], agent_offsets=sess['agent_offsets'], agent_idx=sess['agent_idx'][sl])
    got = tmap.gather_grid()            # flushes
    want = np.full((size, size), -1, np.int8)
    c = c_oracle.integrate_packets(sess['packets'], want, origin[0], origin[1], 0.05, agent_offsets=sess['agent_offsets'],
                                   agent_idx=sess['agent_idx'])
    assert np.array_equal(got, want)
    assert tmap.local.counters()['owned_updates'] == c['updates']


@pytest.mark.parametrize('world', [2, 4])
def test_p2p_route_kernel_emulated_ranks(world):
    """occgrid_route_packets_p2p on one GPU: every logical rank stores its records straight into
    the band owners' receive buffers (plain device buffers here; peer-mapped over NVLink in
    production) in arbitrary arrival order; owners integrate with the ordinals carried by the
    records.  Assembled map == untiled oracle."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import simulation_tools as st
    from occgrid_b200.distributed import BandLayout, CudaBandOps
    from oracle import c_oracle
    size, origin = 1024, (-25.6, -25.6)
    sess = st.generate_session(n_agents=16, n_packets=60_000, grid_size=size, origin=origin, seed=9)
    n = sess['packets'].shape[0]
    offs = sess['agent_offsets'].copy()
    offs[1:3] = (-2.0, origin[1] + size * 0.05 / world - 1.0)
    offs[3:5] = (3.0, origin[1] + size * 0.05 / 2 + 0.2)
    drift = np.random.default_rng(1).normal(0, 0.02, (n, 2))
    layout = BandLayout(size, world)
    ops = [CudaBandOps(layout, r, size, 0.05, origin[0], origin[1], 'cuda', 'auto', n) for r in range(world)]
    tab = torch.from_numpy(offs).cuda()
    cap = n + 1024
    recv = [torch.zeros((cap, 48), dtype=torch.uint8, device='cuda') for _ in range(world)]
    cnt = [torch.zeros(64, dtype=torch.int32, device='cuda') for _ in range(world)]
    recv_ptrs = torch.tensor([t.data_ptr() for t in recv], dtype=torch.int64, device='cuda')
    cnt_ptrs = torch.tensor([t.data_ptr() for t in cnt], dtype=torch.int64, device='cuda')
    stride = (1 << 29) // (world + 1)
    for r in reversed(range(world)):                     # issue order must not matter
        sl = slice(r * n // world, (r + 1) * n // world)
        ops[r].route_p2p(ops[r].stage(sess['packets'][sl]), torch.from_numpy(sess['agent_idx'][sl].copy()).cuda(),
                         torch.from_numpy(drift[sl].copy()).cuda(), tab, r * stride, recv_ptrs, cnt_ptrs, cap)
    torch.cuda.synchronize()
    bands, total = [], 0
    for b in range(world):
        m = int(cnt[b][0].item())
        total += m
        ops[b].grid.update_poses(recv[b][:m], ordinals_in_records=True)
        bands.append(ops[b].band_tensor().cpu().numpy())
    want = np.full((size, size), -1, np.int8)
    c = c_oracle.integrate_packets(sess['packets'], want, origin[0], origin[1], 0.05, agent_offsets=offs,
                                   agent_idx=sess['agent_idx'], drift=drift)
    assert np.array_equal(np.concatenate(bands, axis=0), want)
    assert n < total < 1.5 * n
    assert sum(ops[b].grid.counters()['owned_updates'] for b in range(world)) == c['updates']
