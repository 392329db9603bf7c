"""Authoring-container only (skipped where /root/reference is absent): the oracle restatement
against the LIVE, unmodified reference objects on fresh random inputs — not just the frozen
fixtures."""
import math

import numpy as np
import pytest

from oracle import ref_loader
from oracle import occgrid_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason='/root/reference not present')


@pytest.fixture(scope='module')
def ref():
    return ref_loader.load_dual_bot_mapper()


def test_constants_and_formats(ref):
    assert (ref.PACKET_FMT, ref.PACKET_FMT_V1) == (O.PACKET_FMT, O.PACKET_FMT_V1)
    assert (ref.MAX_DIST_M, ref.MIN_DIST_M) == (O.MAX_DIST_M, O.MIN_DIST_M)
    assert tuple(ref.SENSOR_ANGLES_RAD.values()) == O.SENSOR_ANGLES_RAD
    assert (ref.CLOSURE_RADIUS, ref.MIN_POSES_BETWEEN, ref.CLOSURE_CORRECTION) == \
        (O.CLOSURE_RADIUS, O.MIN_POSES_BETWEEN, O.CLOSURE_CORRECTION)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_update_ray_random(ref, seed):
    rng = np.random.default_rng(seed)
    kw = dict(size=int(rng.integers(20, 300)), resolution=float(rng.choice([0.02, 0.05, 0.1, 0.25])),
              origin_x=float(rng.uniform(-20, 5)), origin_y=float(rng.uniform(-20, 5)))
    a, b = ref.OccupancyGrid(**kw), O.OracleGrid(**kw)
    span = kw['size'] * kw['resolution']
    for _ in range(3000):
        x0, y0 = rng.uniform(kw['origin_x'] - 1, kw['origin_x'] + span + 1, 2)
        ang, d = rng.uniform(-math.pi, math.pi), rng.uniform(0, 1.3)
        hv = bool(rng.random() < 0.5)
        args = (float(x0), float(y0), float(x0 + d * math.cos(ang)), float(y0 + d * math.sin(ang)), hv)
        a.update_ray(*args)
        b.update_ray(*args)
    assert np.array_equal(a.grid, b.grid)
    for _ in range(200):
        p = [int(v) for v in rng.integers(-50, 50, 4)]
        assert a._bresenham(*p) == b.bresenham(*p)
        w = rng.uniform(-30, 30, 2)
        assert a.world_to_grid(*w) == b.world_to_grid(*w) and a.grid_to_world(3, 4) == b.grid_to_world(3, 4)


def test_slam_chain_random(ref):
    rng = np.random.default_rng(5)
    a, b = ref_loader.QuietSLAM(ref), O.OracleSLAM()
    for k in range(1500):
        args = (float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)), float(rng.uniform(-3, 3)), int(rng.integers(1, 3)),
                int(rng.integers(0, 6)) if rng.random() < 0.4 else 0, float(k))
        assert a.add_pose(*args) == b.add_pose(*args)
    assert [tuple(c) for c in a.closures] == b.closures and len(b.closures) > 3


@pytest.mark.parametrize('seed', [11, 12])
def test_map_merger_live_reference_vs_restatement(seed):
    """The unmodified server_nodes/map_merger.py (under rclpy / nav_msgs / open3d stubs) against
    oracle/merge_oracle.OracleMerger on fresh random sequences: same early returns, clouds,
    published grids and origins.  The Open3D arithmetic is the restatement on both sides; the
    sequencing (:35-62) and the NumPy lines (:64-85, :87-127) are the reference's own."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from merge_util import synth_agent_grid
    from oracle import merge_oracle as MO
    mod, script = ref_loader.load_map_merger()
    script.mode = 'script'
    script.queue.clear()
    rng = np.random.default_rng(seed)
    node, o = mod.MapMerger(), MO.OracleMerger()
    for a in range(6):
        n = int(rng.integers(40, 120))
        g = synth_agent_grid(n, seed * 50 + a) if rng.random() > 0.15 else np.full((n, n), 0, np.int8)
        res, ox, oy = 0.05, float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))
        T = MO.se2_matrix(*rng.uniform(-3, 3, 2), rng.uniform(-math.pi, math.pi))
        fit = float(rng.choice([0.3, 0.6, 0.95]))
        if not node.global_pcd.is_empty() and (g > 50).any():
            script.queue.append((T, fit))
        n_pub = len(node.published)
        node.map_callback(ref_loader.make_ref_grid_msg(mod, g.ravel(), n, n, res, ox, oy), a + 1)
        want = o.map_callback(g.ravel(), n, n, res, ox, oy, T, accept=fit >= 0.6)
        assert (len(node.published) > n_pub) == (want is not None)
        pts = np.asarray(node.global_pcd.points).reshape(-1, 3)
        assert np.array_equal(pts[:, 0], o.gx) and np.array_equal(pts[:, 1], o.gy)
        if want is not None:
            m = node.published[-1]
            assert np.array_equal(np.array(m.data, np.int8).reshape(m.info.height, m.info.width), want[0])
            assert (m.info.origin.position.x, m.info.origin.position.y) == want[1]
    assert not script.queue
