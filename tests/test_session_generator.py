"""SURVEY §8d (cfg 2): the vectorised many-agent generator (simulation_tools.generate_session)
against the reference's SCALAR generator on 2 agents — the golden fake_dual_session in
tests/golden/ is the unmodified generate_fake_dual_session.py output (md5 in golden.json).
Same room, same waypoint loops, same noise model: the statistics the benchmark depends on must
agree (cells per beam, hit rate, landmark rate, 15-degree yaw quantisation, wire rounding)."""
import numpy as np

from conftest import session_packets
from occgrid_b200 import simulation_tools as st
from oracle import occgrid_oracle as O


def _stats(packets, grid, offsets=None):
    _, s = O.replay([bytes(p) for p in packets], grid=grid, agent_offsets=offsets)
    arr = np.frombuffer(b''.join(bytes(p) for p in packets), np.uint8).reshape(-1, 42)
    yaw = np.degrees(np.ascontiguousarray(arr[:, 13:17]).view('<f4')[:, 0].astype(np.float64))
    rng_m = np.ascontiguousarray(arr[:, 25:41]).view('<f4').astype(np.float64)
    return {'cells_per_beam': s['updates'] / s['beams'], 'hit_rate': s['hits'] / s['beams'],
            'landmark_rate': float((arr[:, 41] != 0).mean()), 'yaw_mod15': np.abs(((yaw + 7.5) % 15.0) - 7.5).max(),
            'range_mm_residual': np.abs(rng_m * 1000.0 - np.round(rng_m * 1000.0)).max(), 'beams': s['beams']}


def test_vectorised_generator_matches_scalar_reference_generator_statistics():
    gold_pk, _ = session_packets(True)
    gold = _stats(gold_pk, O.OracleGrid())
    sess = st.generate_session(n_agents=2, n_packets=20_000, grid_size=400, origin=(-10.0, -10.0), seed=42)
    offs = {a: (float(sess['agent_offsets'][a, 0]), float(sess['agent_offsets'][a, 1])) for a in (1, 2)}
    vec = _stats(sess['packets'], O.OracleGrid(size=400, resolution=0.05, origin_x=-10.0, origin_y=-10.0), offs)
    assert gold['beams'] == 2748 and abs(gold['cells_per_beam'] - 19.67) < 0.01
    assert abs(vec['cells_per_beam'] / gold['cells_per_beam'] - 1.0) < 0.10, (vec, gold)
    assert abs(vec['hit_rate'] - gold['hit_rate']) < 0.10, (vec, gold)
    assert abs(vec['landmark_rate'] - gold['landmark_rate']) < 0.08, (vec, gold)
    assert vec['yaw_mod15'] < 1e-3 and gold['yaw_mod15'] < 1e-3                          # 15-degree quantised headings (:468)
    assert vec['range_mm_residual'] < 2e-2 and gold['range_mm_residual'] < 2e-2          # ranges written to 1 mm (:476-479)
    assert set(np.unique(sess['agent_idx'])) == {1, 2}


def test_band_layout_puts_the_same_number_of_rooms_in_every_band():
    for bands, rooms, size in ((8, 512, 65536), (4, 256, 32768), (2, 128, 16384)):
        origin = (-size * 0.05 / 2,) * 2
        r = st.room_lattice(rooms, size, 0.05, origin, bands=bands)
        gy = ((r[:, 1] - origin[1]) / 0.05).astype(int)
        assert np.bincount(gy // (size // bands), minlength=bands).tolist() == [rooms // bands] * bands
        assert np.unique(r, axis=0).shape[0] == rooms
    # bands = 1 is the original square lattice (the single-GPU workloads are unchanged)
    assert np.array_equal(st.room_lattice(32, 4096, 0.05, (-102.4, -102.4)), st.room_lattice(32, 4096, 0.05, (-102.4, -102.4), bands=1))


def test_dispersed_variant_keeps_packets_but_spreads_poses():
    s = st.generate_session(n_agents=8, n_packets=4000, grid_size=1024, origin=(-25.6, -25.6), seed=3)
    d = st.disperse_poses(s, seed=1)
    a, b = s['packets'].view(st.PACKET_DTYPE).reshape(-1), d['packets'].view(st.PACKET_DTYPE).reshape(-1)
    for f in ('magic', 'agent', 'yaw', 'front', 'left', 'back', 'right', 'lm'):
        assert np.array_equal(a[f], b[f])
    wx = b['x'].astype(np.float64) + d['agent_offsets'][d['agent_idx'], 0]
    assert wx.min() > -25.6 and wx.max() < 25.6 and np.std(wx) > 10.0
