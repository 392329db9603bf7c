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
