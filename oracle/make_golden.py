"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Generates the golden fixtures under ``tests/golden/`` by EXECUTING THE UNMODIFIED
REFERENCE in the authoring container (``/root/reference`` must exist; it does not on
the GPU box, which is why the outputs are committed).  Run:

    python oracle/make_golden.py

Recipe = SURVEY.md Appendix B.  The per-packet loop of ``main()`` is inline in the
reference (server_nodes/dual_bot_mapper.py:826-919) and cannot be imported, so the
~25 lines below restate it around the reference's own ``OccupancyGrid`` and
``PoseGraphSLAM`` objects; everything numeric is done by reference code.
"""
import hashlib
import json
import math
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def census(g):
    return {'free': int((g == 0).sum()), 'occ': int((g == 100).sum()),
            'unk': int((g == -1).sum()), 'sha1': sha1(g)}


def run_reference_generator():
    """simulation_tools/generate_fake_dual_session.py writes beside itself (:322-325) and the
    mount is read-only, so run a byte-identical copy from a temp dir."""
    src = os.path.join(ref_loader.REFERENCE_ROOT, 'simulation_tools',
                       'generate_fake_dual_session.py')
    tmp = tempfile.mkdtemp(prefix='refgen_')
    shutil.copy(src, tmp)
    subprocess.run([sys.executable, 'generate_fake_dual_session.py'], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    out = os.path.join(tmp, 'server_nodes', 'logs', 'dual_session_20260611_062145')
    dst = os.path.join(GOLD, 'fake_dual_session')
    os.makedirs(dst, exist_ok=True)
    shutil.copy(os.path.join(out, 'telemetry.csv'), os.path.join(dst, 'telemetry.csv'))
    md5 = hashlib.md5(open(os.path.join(dst, 'telemetry.csv'), 'rb').read()).hexdigest()
    shutil.rmtree(tmp)
    return md5


def reference_replay(m, packets, separation, use_slam, grid_kwargs=None):
    """dual_bot_mapper.py:826-919 around reference objects.  Returns (grid, drift table,
    closures).  ``packets`` are raw datagrams."""
    occ = m.OccupancyGrid(**(grid_kwargs or {}))
    slam = ref_loader.QuietSLAM(m) if use_slam else None
    drift_correction = {1: (0.0, 0.0), 2: (0.0, 0.0)}
    drift_table = []
    for k, data in enumerate(packets):
        landmark_type = m.LM_NONE
        if len(data) == m.PACKET_SIZE:
            (magic, agent_id, rx, ry, ryaw, enc, v2v, d_front, d_left, d_back, d_right,
             landmark_type) = struct.unpack(m.PACKET_FMT, data)
        elif len(data) == m.PACKET_SIZE_V1:
            (magic, agent_id, rx, ry, ryaw, enc, v2v, d_front, d_left, d_back,
             d_right) = struct.unpack(m.PACKET_FMT_V1, data)
        else:
            drift_table.append((0.0, 0.0))
            continue
        if magic != b'QSRL' or agent_id not in [1, 2]:
            drift_table.append((0.0, 0.0))
            continue
        if agent_id == 2:
            rx += separation
        cdx, cdy = drift_correction[agent_id]
        drift_table.append((cdx, cdy))
        rx += cdx
        ry += cdy
        sensors = {'front': d_front, 'left': d_left, 'back': d_back, 'right': d_right}
        for name, dist in sensors.items():
            ray_angle = ryaw + m.SENSOR_ANGLES_RAD[name]
            hit_valid = m.MIN_DIST_M < dist <= m.MAX_DIST_M
            if hit_valid:
                wx = rx + dist * math.cos(ray_angle)
                wy = ry + dist * math.sin(ray_angle)
                occ.update_ray(rx, ry, wx, wy, True)
            else:
                max_range = min(dist, m.MAX_DIST_M) if dist > m.MIN_DIST_M else m.MAX_DIST_M
                end_x = rx + max_range * math.cos(ray_angle)
                end_y = ry + max_range * math.sin(ray_angle)
                occ.update_ray(rx, ry, end_x, end_y, False)
        if slam is not None:
            closure, cdx_new, cdy_new = slam.add_pose(rx, ry, ryaw, agent_id, landmark_type,
                                                      float(k))
            if closure:
                drift_correction[agent_id] = (drift_correction[agent_id][0] + cdx_new,
                                              drift_correction[agent_id][1] + cdy_new)
    return occ.grid, drift_table, (slam.closures if slam else [])


def random_packets(rng, n, span=4.0):
    """Adversarial packet stream: both agents, quantised and continuous yaws, range
    sentinels from the firmware (0, 4.0 timeout: AgentFirmware_Bot1.ino:239,386), NaN/inf
    ranges, float32 edge values of the trust filter, bad magic, foreign agent ids, v1 size,
    junk sizes.  Poses stay finite (the reference crashes on NaN poses)."""
    out = []
    special = [0.0, -1.0, 4.0, float('nan'), float('inf'), 0.05, 1.2, 1.2000001, 0.050000004,
               0.01, 2.5, 1.1999999]
    for i in range(n):
        agent = int(rng.integers(1, 3))
        x = float(rng.uniform(-span, span))
        y = float(rng.uniform(-span, span))
        r = rng.random()
        if r < 0.4:
            yaw = math.radians(15.0 * int(rng.integers(-12, 13)))
        elif r < 0.5:
            yaw = float(rng.uniform(-50.0, 50.0))
        else:
            yaw = float(rng.uniform(-math.pi, math.pi))
        if rng.random() < 0.2:
            x = round(x / 0.05) * 0.05
            y = round(y / 0.05) * 0.05
        d = []
        for _ in range(4):
            if rng.random() < 0.25:
                d.append(special[int(rng.integers(0, len(special)))])
            else:
                d.append(round(float(rng.uniform(0.0, 1.6)), 3))
        lm = int(rng.integers(0, 6)) if rng.random() < 0.3 else 0
        magic = b'QSRL'
        kind = rng.random()
        if kind < 0.03:
            magic = b'QSRX'
        elif kind < 0.06:
            agent = int(rng.integers(3, 256)) if rng.random() < 0.7 else 0
        pkt = struct.pack('<4sBfffiIffffB', magic, agent, x, y, yaw,
                          int(rng.integers(-1000, 100000)), int(rng.integers(0, 1000)),
                          d[0], d[1], d[2], d[3], lm)
        if 0.06 <= kind < 0.10:
            pkt = pkt[:41]                      # v1 datagram
        elif 0.10 <= kind < 0.12:
            pkt = pkt[:int(rng.integers(0, 41))]  # junk size -> dropped
        elif 0.12 <= kind < 0.13:
            pkt = pkt + b'\x00'                   # 43 bytes -> dropped
        out.append(pkt)
    return out


def main():
    assert ref_loader.reference_available(), 'needs /root/reference'
    m = ref_loader.load_dual_bot_mapper()
    os.makedirs(GOLD, exist_ok=True)
    gold = {'reference_constants': {
        'PACKET_SIZE': m.PACKET_SIZE, 'PACKET_SIZE_V1': m.PACKET_SIZE_V1,
        'MAX_DIST_M': m.MAX_DIST_M, 'MIN_DIST_M': m.MIN_DIST_M,
        'GRID_RESOLUTION': m.GRID_RESOLUTION, 'GRID_SIZE': m.GRID_SIZE,
        'GRID_ORIGIN_X': m.GRID_ORIGIN_X, 'GRID_ORIGIN_Y': m.GRID_ORIGIN_Y,
        'CELL_UNKNOWN': m.CELL_UNKNOWN, 'CELL_FREE': m.CELL_FREE,
        'CELL_OCCUPIED': m.CELL_OCCUPIED,
        'SENSOR_ANGLES_RAD': [m.SENSOR_ANGLES_RAD[k] for k in ('front', 'left', 'back', 'right')],
        'sensor_order': list(m.SENSOR_ANGLES_RAD.keys()),
    }}

    # 1. golden 2-bot session (BASELINE config 1) ---------------------------------------
    gold['telemetry_md5'] = run_reference_generator()
    from oracle.occgrid_oracle import load_session_rows, rows_to_packets
    csv_path = os.path.join(GOLD, 'fake_dual_session', 'telemetry.csv')
    sess = {}
    for order in ('file', 'time'):
        rows = load_session_rows(csv_path, time_sorted=(order == 'time'))
        pk = rows_to_packets(rows)
        for use_slam in (True, False):
            g, drift, closures = reference_replay(m, pk, 0.0, use_slam)
            c = census(g)
            c['closures'] = len(closures)
            c['packets'] = len(pk)
            sess[f'{order}_order_slam_{"on" if use_slam else "off"}'] = c
            if use_slam:
                np.save(os.path.join(GOLD, f'session_drift_{order}.npy'),
                        np.asarray(drift, dtype=np.float64))
    # separation != 0 variant exercises the agent-2 shift (:851-852)
    rows = load_session_rows(csv_path, time_sorted=True)
    g, _, _ = reference_replay(m, rows_to_packets(rows), -2.5, False)
    sess['time_order_slam_off_sep_-2.5'] = census(g)
    gold['session'] = sess

    # 2. exhaustive Bresenham shape KAT over [-32,32]^2 --------------------------------
    og = m.OccupancyGrid()
    xs, ys, offs = [], [], [0]
    for dy in range(-32, 33):
        for dx in range(-32, 33):
            cells = og._bresenham(0, 0, dx, dy)
            xs.extend(c[0] for c in cells)
            ys.extend(c[1] for c in cells)
            offs.append(len(xs))
    np.savez_compressed(os.path.join(GOLD, 'bresenham_kat.npz'),
                        x=np.asarray(xs, np.int8), y=np.asarray(ys, np.int8),
                        offsets=np.asarray(offs, np.int32))
    # translated / long lines: a few hundred random absolute endpoints
    rng = np.random.default_rng(7)
    ends = rng.integers(-300, 300, size=(300, 4))
    h = hashlib.sha1()
    for x0, y0, x1, y1 in ends.tolist():
        h.update(np.asarray(og._bresenham(x0, y0, x1, y1), np.int32).tobytes())
    np.save(os.path.join(GOLD, 'bresenham_long_endpoints.npy'), ends.astype(np.int32))
    gold['bresenham_long_sha1'] = h.hexdigest()

    # 3. world_to_grid truncation KATs --------------------------------------------------
    pts = [(-5.01, -5.01), (-5.06, -5.06), (0.0, 0.0), (4.999999, 4.95), (5.0, 5.0),
           (-4.95, -4.9), (0.15, 0.3), (1e-9, -1e-9), (-5.0, -5.05), (-5.049999, 2.0),
           (123.456, -77.7)]
    gold['world_to_grid'] = [{'w': list(p), 'g': list(og.world_to_grid(*p))} for p in pts]
    og2 = m.OccupancyGrid(size=4096, resolution=0.05, origin_x=-102.4, origin_y=-102.4)
    gold['world_to_grid_4096'] = [{'w': list(p), 'g': list(og2.world_to_grid(*p))} for p in pts]

    # 4. random update_ray streams on three geometries ----------------------------------
    rays = {}
    for name, seed, kw, span in (
            ('default200', 21, {}, 6.0),
            ('g512_r0.1', 22, dict(size=512, resolution=0.1, origin_x=-20.0, origin_y=-31.3), 30.0),
            ('g64_r0.02', 23, dict(size=64, resolution=0.02, origin_x=0.0, origin_y=-0.64), 1.5)):
        rng = np.random.default_rng(seed)
        n = 3000
        x0 = rng.uniform(-span, span, n)
        y0 = rng.uniform(-span, span, n)
        ang = rng.uniform(-math.pi, math.pi, n)
        d = rng.uniform(0.0, 1.2, n)
        x1 = x0 + d * np.cos(ang)
        y1 = y0 + d * np.sin(ang)
        hv = rng.random(n) < 0.5
        g = m.OccupancyGrid(**kw)
        for i in range(n):
            g.update_ray(float(x0[i]), float(y0[i]), float(x1[i]), float(y1[i]), bool(hv[i]))
        np.savez_compressed(os.path.join(GOLD, f'rays_{name}.npz'),
                            x0=x0, y0=y0, x1=x1, y1=y1, hit=hv)
        rays[name] = dict(census(g.grid), grid_kwargs=kw)
    gold['rays'] = rays

    # 5. adversarial packet streams through the reference loop --------------------------
    streams = {}
    for name, seed, n, sep, use_slam, kw, span in (
            ('mixed_a', 11, 4000, 0.0, False, {}, 4.5),
            ('mixed_b_sep', 12, 4000, 0.75, True, {}, 4.5),
            ('mixed_c_4096', 13, 6000, 0.0, False,
             dict(size=4096, resolution=0.05, origin_x=-102.4, origin_y=-102.4), 104.0),
            ('mixed_d_edge', 14, 5000, 0.5, True,
             dict(size=96, resolution=0.05, origin_x=-2.4, origin_y=-2.4), 3.2)):
        rng = np.random.default_rng(seed)
        pk = random_packets(rng, n, span)
        g, drift, closures = reference_replay(m, pk, sep, use_slam, kw)
        lens = np.asarray([len(p) for p in pk], np.int32)
        blob = np.frombuffer(b''.join(pk), np.uint8)
        np.savez_compressed(os.path.join(GOLD, f'packets_{name}.npz'),
                            blob=blob, lens=lens, drift=np.asarray(drift, np.float64))
        streams[name] = dict(census(g), separation=sep, slam=use_slam, grid_kwargs=kw,
                             closures=len(closures), n=n)
    gold['packet_streams'] = streams

    # 6. hit/miss classification KATs (expression at :888 and :900 with reference constants)
    cls = []
    for v in [0.0, -1.0, 4.0, 0.05, 1.2, 1.25, 0.050000004, 1.2000001, 0.5, 1.1999999,
              float('nan'), float('inf')]:
        dist = struct.unpack('<f', struct.pack('<f', v))[0]
        hit_valid = m.MIN_DIST_M < dist <= m.MAX_DIST_M
        rng_used = dist if hit_valid else (min(dist, m.MAX_DIST_M) if dist > m.MIN_DIST_M
                                           else m.MAX_DIST_M)
        cls.append({'f32_bits': struct.unpack('<I', struct.pack('<f', v))[0],
                    'hit': bool(hit_valid), 'range': rng_used})
    gold['hit_classification'] = cls

    with open(os.path.join(GOLD, 'golden.json'), 'w') as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print(json.dumps(gold['session'], indent=1))
    print('telemetry md5', gold['telemetry_md5'])


if __name__ == '__main__':
    main()
