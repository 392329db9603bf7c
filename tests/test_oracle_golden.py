"""The oracle (Python restatement + C restatement) against fixtures produced by the
unmodified reference (oracle/make_golden.py).  CPU only."""
import hashlib
import math
import os
import struct

import numpy as np
import pytest

from conftest import GOLD, load_packet_stream, normalise_datagrams, session_packets
from oracle import c_oracle
from oracle import occgrid_oracle as O


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_constants_match_reference(golden):
    c = golden['reference_constants']
    assert O.PACKET_SIZE == c['PACKET_SIZE'] == 42
    assert O.PACKET_SIZE_V1 == c['PACKET_SIZE_V1'] == 41
    assert (O.MAX_DIST_M, O.MIN_DIST_M) == (c['MAX_DIST_M'], c['MIN_DIST_M'])
    assert (O.GRID_RESOLUTION, O.GRID_SIZE) == (c['GRID_RESOLUTION'], c['GRID_SIZE'])
    assert (O.CELL_UNKNOWN, O.CELL_FREE, O.CELL_OCCUPIED) == (-1, 0, 100)
    assert list(O.SENSOR_ANGLES_RAD) == c['SENSOR_ANGLES_RAD']
    assert c['sensor_order'] == ['front', 'left', 'back', 'right']


def test_telemetry_fixture_md5(golden):
    data = open(os.path.join(GOLD, 'fake_dual_session', 'telemetry.csv'), 'rb').read()
    assert hashlib.md5(data).hexdigest() == golden['telemetry_md5'] == 'f7a9c7184da98c83af126973c0611e5a'


def test_bresenham_exhaustive_kat():
    z = np.load(os.path.join(GOLD, 'bresenham_kat.npz'))
    xs, ys, offs = z['x'], z['y'], z['offsets']
    i = 0
    for dy in range(-32, 33):
        for dx in range(-32, 33):
            want = list(zip(xs[offs[i]:offs[i + 1]].tolist(), ys[offs[i]:offs[i + 1]].tolist()))
            assert O.OracleGrid.bresenham(0, 0, dx, dy) == want
            assert c_oracle.bresenham(0, 0, dx, dy) == want
            assert len(want) == max(abs(dx), abs(dy)) + 1
            i += 1
    assert i == 65 * 65


def test_bresenham_long_lines(golden):
    ends = np.load(os.path.join(GOLD, 'bresenham_long_endpoints.npy'))
    h1, h2 = hashlib.sha1(), hashlib.sha1()
    for x0, y0, x1, y1 in ends.tolist():
        h1.update(np.asarray(O.OracleGrid.bresenham(x0, y0, x1, y1), np.int32).tobytes())
        h2.update(np.asarray(c_oracle.bresenham(x0, y0, x1, y1), np.int32).tobytes())
    assert h1.hexdigest() == h2.hexdigest() == golden['bresenham_long_sha1']


def test_world_to_grid_truncation(golden):
    g = O.OracleGrid()
    for e in golden['world_to_grid']:
        assert list(g.world_to_grid(*e['w'])) == e['g']
    g2 = O.OracleGrid(size=8, resolution=0.05, origin_x=-102.4, origin_y=-102.4)
    for e in golden['world_to_grid_4096']:
        assert list(g2.world_to_grid(*e['w'])) == e['g']
    # truncation, not floor
    assert g.world_to_grid(-5.01, -5.01) == (0, 0)
    assert g.world_to_grid(-5.06, -5.06) == (-1, -1)
    assert g.world_to_grid(0.0, 0.0) == (100, 100)


def test_hit_classification(golden):
    for e in golden['hit_classification']:
        dist = struct.unpack('<f', struct.pack('<I', e['f32_bits']))[0]
        beams = O.expand_beams(0.0, 0.0, 0.0, (dist, dist, dist, dist))
        assert beams[0][4] == e['hit']
        want_range = e['range']
        assert beams[0][2] == 0.0 + want_range * math.cos(0.0)
    # float32(1.2) > 1.2 -> miss; float32(0.05) > 0.05 -> hit
    f32 = lambda v: struct.unpack('<f', struct.pack('<f', v))[0]
    assert not (O.MIN_DIST_M < f32(1.2) <= O.MAX_DIST_M)
    assert O.MIN_DIST_M < f32(0.05) <= O.MAX_DIST_M


@pytest.mark.parametrize('name', ['default200', 'g512_r0.1', 'g64_r0.02'])
def test_update_ray_streams(golden, name):
    z = np.load(os.path.join(GOLD, f'rays_{name}.npz'))
    want = golden['rays'][name]
    kw = want['grid_kwargs']
    g = O.OracleGrid(**kw)
    for i in range(len(z['x0'])):
        g.update_ray(float(z['x0'][i]), float(z['y0'][i]), float(z['x1'][i]), float(z['y1'][i]),
                     bool(z['hit'][i]))
    assert sha1(g.grid) == want['sha1']
    gc = np.full_like(g.grid, -1)
    rays = np.stack([z['x0'], z['y0'], z['x1'], z['y1']], axis=1)
    cnt = c_oracle.update_rays(rays, z['hit'], gc, g.ox, g.oy, g.res)
    assert sha1(gc) == want['sha1']
    assert cnt['updates'] == g.updates


@pytest.mark.parametrize('order', ['file', 'time'])
@pytest.mark.parametrize('slam', [True, False])
def test_golden_session_python_oracle(golden, order, slam):
    pk, rows = session_packets(time_sorted=(order == 'time'))
    want = golden['session'][f'{order}_order_slam_{"on" if slam else "off"}']
    drift = []
    s = O.OracleSLAM() if slam else None
    g, st = O.replay(pk, separation=0.0, slam=s, drift_out=drift)
    c = O.grid_census(g.grid)
    assert (c['free'], c['occ'], c['unk'], c['sha1']) == (want['free'], want['occ'], want['unk'], want['sha1'])
    assert st['packets'] == 687 and st['beams'] == 2748
    if slam:
        assert len(s.closures) == want['closures'] == 10
        ref_drift = np.load(os.path.join(GOLD, f'session_drift_{order}.npy'))
        assert np.array_equal(np.asarray(drift), ref_drift)
    if order == 'file' and not slam:
        assert st['updates'] == 54049 and st['hits'] == 1041


def test_golden_session_order_sensitivity(golden):
    s = golden['session']
    assert s['file_order_slam_off']['sha1'] != s['time_order_slam_off']['sha1']


@pytest.mark.parametrize('order', ['file', 'time'])
@pytest.mark.parametrize('slam', [True, False])
def test_golden_session_c_oracle(golden, order, slam):
    pk, _ = session_packets(time_sorted=(order == 'time'))
    want = golden['session'][f'{order}_order_slam_{"on" if slam else "off"}']
    drift = np.load(os.path.join(GOLD, f'session_drift_{order}.npy')) if slam else None
    arr, d = normalise_datagrams(pk, drift)
    grid = np.full((200, 200), -1, np.int8)
    cnt = c_oracle.integrate_packets(arr, grid, -5.0, -5.0, 0.05, drift=d)
    assert sha1(grid) == want['sha1']
    assert cnt['beams'] == 2748


def test_golden_session_separation(golden):
    pk, _ = session_packets(time_sorted=True)
    want = golden['session']['time_order_slam_off_sep_-2.5']
    g, _ = O.replay(pk, separation=-2.5)
    assert sha1(g.grid) == want['sha1']
    arr, _ = normalise_datagrams(pk)
    grid = np.full((200, 200), -1, np.int8)
    c_oracle.integrate_packets(arr, grid, -5.0, -5.0, 0.05, separation=-2.5)
    assert sha1(grid) == want['sha1']


@pytest.mark.parametrize('name', ['mixed_a', 'mixed_b_sep', 'mixed_c_4096', 'mixed_d_edge'])
def test_adversarial_packet_streams(golden, name):
    want = golden['packet_streams'][name]
    pk, ref_drift = load_packet_stream(name)
    kw = want['grid_kwargs']
    g = O.OracleGrid(**kw)
    drift = []
    s = O.OracleSLAM() if want['slam'] else None
    O.replay(pk, grid=g, separation=want['separation'], slam=s, drift_out=drift)
    assert sha1(g.grid) == want['sha1']
    assert np.array_equal(np.asarray(drift), ref_drift)
    if s is not None:
        assert len(s.closures) == want['closures']
    # C oracle, fed the reference's own drift table
    arr, d = normalise_datagrams(pk, ref_drift)
    grid = np.full_like(g.grid, -1)
    cnt = c_oracle.integrate_packets(arr, grid, g.ox, g.oy, g.res, separation=want['separation'], drift=d)
    assert sha1(grid) == want['sha1']
    assert cnt['updates'] == g.updates


def test_c_oracle_window_equals_crop():
    """Tile semantics used by the multi-GPU path: integrating into a window of the global
    grid == cropping the globally integrated grid (per-cell clipping)."""
    pk, _ = load_packet_stream('mixed_a')
    arr, _ = normalise_datagrams(pk)
    full = np.full((200, 200), -1, np.int8)
    c_oracle.integrate_packets(arr, full, -5.0, -5.0, 0.05)
    for (x0, y0, w, h) in [(0, 0, 200, 100), (0, 100, 200, 100), (37, 51, 64, 80)]:
        tile = np.full((h, w), -1, np.int8)
        c_oracle.integrate_packets(arr, tile, -5.0, -5.0, 0.05, window=(x0, y0), size_x=200, size_y=200)
        assert np.array_equal(tile, full[y0:y0 + h, x0:x0 + w])


def test_nan_pose_is_skipped_not_fatal():
    good = struct.pack(O.PACKET_FMT, b'QSRL', 1, 0.0, 0.0, 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    bad = struct.pack(O.PACKET_FMT, b'QSRL', 1, float('nan'), 0.0, 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    inf = struct.pack(O.PACKET_FMT, b'QSRL', 2, 0.0, float('inf'), 0.0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0)
    g1, st = O.replay([good, bad, inf])
    g2, _ = O.replay([good])
    assert np.array_equal(g1.grid, g2.grid) and st['bad_pose'] == 2
    arr, _ = normalise_datagrams([good, bad, inf])
    grid = np.full((200, 200), -1, np.int8)
    cnt = c_oracle.integrate_packets(arr, grid, -5.0, -5.0, 0.05)
    assert np.array_equal(grid, g2.grid) and cnt['bad_pose'] == 2
