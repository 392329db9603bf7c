"""Synthetic multi-agent sessions and the session replay format.

Mirrors the two tools of the reference that feed the hot path:

* ``simulation_tools/generate_fake_dual_session.py`` — a seeded 2-bot simulation of one
  6 m x 4 m room (walls :44-54, waypoints :137-214, sensor model :93-110).  Here the same
  room/sensor model is vectorised and scaled to many rooms ("generate_fake_dual_session scaled
  to many agents", BASELINE.json configs 2, 4, 5): every room hosts a BOT1/BOT2 pair that laps
  its waypoint loop; packets are emitted in the 42-byte QuasarPacket v2 wire format
  (server_nodes/dual_bot_mapper.py:41-42).  Statistics preserved from the reference generator:
  range noise N(0, 0.035 m) (:100), 6 % outliers U(0.02, 2.5) (:104-105), floor 0.01 (:107),
  yaw quantised to 15 deg (:468), pose printed to 1e-4 m and ranges to 1 mm (:476-479) before the
  fp32 cast, 5 % duplicated packets (:471), bot-2 time jitter +-0.08 s (:505).
  The scalar generator's exact random stream is NOT reproduced (it interleaves Python's
  Mersenne Twister calls per step); the golden 2-bot session itself is replayed from the
  committed CSV fixture instead.

* ``simulation_tools/playback_dual_session.py:58-105`` ``load_session`` — CSV -> rows, sorted
  by time; ``rows_to_packets`` re-packs rows into wire records.

Host-side NumPy only: this is workload generation, not the hot path.
"""
from __future__ import annotations

import csv
import math
import os
import struct

import numpy as np

PACKET_DTYPE = np.dtype([('magic', 'S4'), ('agent', 'u1'), ('x', '<f4'), ('y', '<f4'), ('yaw', '<f4'),
                         ('enc', '<i4'), ('v2v', '<u4'), ('front', '<f4'), ('left', '<f4'),
                         ('back', '<f4'), ('right', '<f4'), ('lm', 'u1')], align=False)
assert PACKET_DTYPE.itemsize == 42

# Room of the reference generator (:44-54): x in [-0.5, 5.5], y in [-2, 2] in room coordinates.
ROOM_X = (-0.5, 5.5)
ROOM_Y = (-2.0, 2.0)
MAX_SENSOR_RANGE = 1.20
SENSOR_ANGLES = (0.0, math.pi / 2, math.pi, -math.pi / 2)

# Waypoint loops (x, y, heading_deg) of the two bots (reference :137-214): BOT1 maps the left
# half and BOT2 the right half of the room; both return to their start pose.
_BOT1 = [(0.0, 0.0, 90), (0.0, 1.3, 90), (0.0, 1.3, 180), (-0.2, 1.3, 180), (-0.2, 1.3, 90), (-0.2, 1.7, 90),
         (-0.2, 1.7, 0), (2.45, 1.7, 0), (2.45, 1.7, -90), (2.45, -1.7, -90), (2.45, -1.7, 180),
         (0.0, -1.7, 180), (0.0, -1.7, 90), (0.0, 0.0, 90)]
_BOT2 = [(5.0, 0.0, 90), (5.0, 1.3, 90), (5.0, 1.3, 0), (5.2, 1.3, 0), (5.2, 1.3, 90), (5.2, 1.7, 90),
         (5.2, 1.7, 180), (3.0, 1.7, 180), (3.0, 1.7, -90), (3.0, -1.7, -90), (3.0, -1.7, 0),
         (5.0, -1.7, 0), (5.0, -1.7, 90), (5.0, 0.0, 90)]


def _lap(waypoints, steps_per_meter=25, turn_steps=4):
    """Dense (x, y, yaw) poses of one lap: 25 steps per metre on straights, 4 steps per turn
    (reference :225-311, without its stochastic wall-following wiggle, which is added per
    packet in ``generate_session``)."""
    px, py, pyaw = [], [], []
    for (x1, y1, a1), (x2, y2, a2) in zip(waypoints[:-1], waypoints[1:]):
        dist = math.hypot(x2 - x1, y2 - y1)
        if dist < 0.05:
            d = math.radians(a2 - a1)
            d = (d + math.pi) % (2 * math.pi) - math.pi
            for j in range(turn_steps):
                px.append(x1), py.append(y1), pyaw.append(math.radians(a1) + d * j / turn_steps)
        else:
            n = max(5, int(dist * steps_per_meter))
            heading = math.atan2(y2 - y1, x2 - x1)
            for j in range(n):
                t = j / n
                px.append(x1 + t * (x2 - x1)), py.append(y1 + t * (y2 - y1)), pyaw.append(heading)
    return np.array(px), np.array(py), np.array(pyaw)


def room_lattice(n_rooms, grid_size, resolution, origin, bands=1):
    """Room origins on a near-square lattice inside the grid, each 6 m x 4 m room fully inside.
    ``bands`` > 1 (multi-GPU row bands): every band of grid_size / bands rows gets the same number of
    rooms, on its own lattice inside the band and clear of the band edges — the weak-scaling premise
    (same work per GPU) made true for the synthetic swarm."""
    span = grid_size * resolution
    if bands > 1:
        if n_rooms % bands:
            raise ValueError('rooms must divide evenly into the bands')
        per = n_rooms // bands
        band_h = span / bands
        rows = max(1, int(round(math.sqrt(per * band_h / span / 1.5))))
        cols = int(math.ceil(per / rows))
        pitch_x, pitch_y = span / cols, band_h / rows
        if pitch_x < 8.0 or pitch_y < 6.0:
            raise ValueError('grid too small for that many rooms')
        out = []
        for b in range(bands):
            for r in range(per):
                i, j = r % cols, r // cols
                out.append((origin[0] + (i + 0.5) * pitch_x - 2.5, origin[1] + b * band_h + (j + 0.5) * pitch_y))
        return np.asarray(out, np.float64)
    cols = int(math.ceil(math.sqrt(n_rooms * 1.5)))
    rows = int(math.ceil(n_rooms / cols))
    pitch_x, pitch_y = span / cols, span / rows
    if pitch_x < 8.0 or pitch_y < 6.0:
        raise ValueError('grid too small for that many rooms')
    out = []
    for r in range(n_rooms):
        i, j = r % cols, r // cols
        out.append((origin[0] + (i + 0.5) * pitch_x - 2.5, origin[1] + (j + 0.5) * pitch_y))
    return np.asarray(out, np.float64)


def generate_session(n_agents=64, n_packets=2_500_000, grid_size=4096, resolution=0.05,
                     origin=(-102.4, -102.4), seed=42, time_sorted=True, bands=1):
    """Synthetic many-agent session.

    Returns a dict:
      packets        uint8 [n_packets, 42] wire records in stream (time) order
      agent_idx      int32 [n_packets]     1-based agent index (== the wire byte while <= 255)
      agent_offsets  float64 [n_agents+1, 2] start offset of each agent's odometry frame
                     (the many-agent generalisation of ``--separation``,
                     server_nodes/dual_bot_mapper.py:851-852); row 0 unused
      grid           dict(size, resolution, origin_x, origin_y)
    Agents 2r+1 / 2r+2 are BOT1 / BOT2 of room r; packet poses are in room coordinates.
    """
    if n_agents % 2:
        raise ValueError('agents come in BOT1/BOT2 pairs')
    rng = np.random.default_rng(seed)
    n_rooms = n_agents // 2
    rooms = room_lattice(n_rooms, grid_size, resolution, origin, bands)
    offsets = np.zeros((n_agents + 1, 2), np.float64)
    offsets[1::2] = rooms
    offsets[2::2] = rooms
    laps = (_lap(_BOT1), _lap(_BOT2))

    per_agent = -(-n_packets // n_agents)
    agent = np.repeat(np.arange(1, n_agents + 1, dtype=np.int32), per_agent)
    step = np.tile(np.arange(per_agent, dtype=np.int64), n_agents)
    bot = (agent - 1) % 2
    m = agent.shape[0]

    # true pose: lap position + wall-following wiggle (lateral +-0.15 m hysteresis in the
    # reference; a bounded random lateral/heading perturbation here)
    x = np.empty(m), np.empty(m), np.empty(m)
    tx, ty, tyaw = x
    for b in (0, 1):
        sel = bot == b
        lx, ly, lyaw = laps[b]
        idx = step[sel] % lx.shape[0]
        tx[sel], ty[sel], tyaw[sel] = lx[idx], ly[idx], lyaw[idx]
    lat = rng.normal(0.0, 0.05, m).clip(-0.2, 0.2)
    tx = tx - lat * np.sin(tyaw) + rng.normal(0.0, 0.004, m)
    ty = ty + lat * np.cos(tyaw) + rng.normal(0.0, 0.004, m)
    tyaw = tyaw + rng.normal(0.0, 0.08, m)
    tx = tx.clip(ROOM_X[0] + 0.05, ROOM_X[1] - 0.05)
    ty = ty.clip(ROOM_Y[0] + 0.05, ROOM_Y[1] - 0.05)

    # 4 ranges from the TRUE pose (:93-110): nearest wall of the rectangular room
    dists = np.empty((m, 4))
    for s, rel in enumerate(SENSOR_ANGLES):
        c, sn = np.cos(tyaw + rel), np.sin(tyaw + rel)
        with np.errstate(divide='ignore', invalid='ignore'):
            t_x = np.where(c > 0, (ROOM_X[1] - tx) / c, np.where(c < 0, (ROOM_X[0] - tx) / c, np.inf))
            t_y = np.where(sn > 0, (ROOM_Y[1] - ty) / sn, np.where(sn < 0, (ROOM_Y[0] - ty) / sn, np.inf))
        d = np.minimum(t_x, t_y) + rng.normal(0.0, 0.035, m)
        outlier = rng.random(m) < 0.06
        d = np.where(outlier, rng.uniform(0.02, 2.5, m), d)
        dists[:, s] = np.maximum(0.01, d)

    # odometry estimate: true pose + slowly growing drift; yaw quantised to 15 deg (:468)
    drift_scale = 0.0004 * np.sqrt(step + 1.0)
    ex = tx + rng.normal(0.0, 1.0, m) * drift_scale
    ey = ty + rng.normal(0.0, 1.0, m) * drift_scale
    eyaw_deg = np.round(np.degrees(tyaw + rng.normal(0.0, 0.01, m)) / 15.0) * 15.0
    eyaw_deg = (eyaw_deg + 180.0) % 360.0 - 180.0

    # landmark signature (:113-129)
    f, l, r_ = dists[:, 0], dists[:, 1], dists[:, 3]
    close = 0.30
    lm = np.zeros(m, np.uint8)
    lm[(f > MAX_SENSOR_RANGE) & (l > MAX_SENSOR_RANGE) & (r_ > MAX_SENSOR_RANGE)] = 5
    lm[(f < close) & (l < close) & (r_ < close)] = 4
    lm[(l < close) & (r_ < close) & (f > close)] = 3
    lm[(f < close) & (r_ < close) & (l > close)] = 2
    lm[(f < close) & (l < close) & (r_ > close)] = 1

    # time stamps: ~0.55 s per step (:389), bot-2 jitter (:505); 5 % duplicates (:471)
    t = step * 0.55 + rng.uniform(-0.05, 0.05, m) + np.where(bot == 1, rng.uniform(-0.08, 0.08, m), 0.0)
    dup = rng.random(m) < 0.05
    order_src = np.concatenate([np.arange(m), np.nonzero(dup)[0]])
    t_all = np.concatenate([t, t[dup] + rng.uniform(-0.01, 0.01, int(dup.sum()))])
    if time_sorted:
        perm = np.argsort(t_all, kind='stable')
    else:
        perm = np.arange(order_src.shape[0])
    src = order_src[perm][:n_packets]

    rec = np.zeros(src.shape[0], PACKET_DTYPE)
    rec['magic'] = b'QSRL'
    rec['agent'] = np.minimum(agent[src], 255).astype(np.uint8)
    rec['x'] = np.round(ex[src], 4).astype(np.float32)          # f"{x:.4f}" then fp32 wire field
    rec['y'] = np.round(ey[src], 4).astype(np.float32)
    rec['yaw'] = np.radians(eyaw_deg[src]).astype(np.float32)
    rec['enc'] = (step[src] * 4).astype(np.int32)
    rec['v2v'] = 0
    for s, name in enumerate(('front', 'left', 'back', 'right')):
        rec[name] = (np.round(dists[src, s] * 100.0, 1) / 100.0).astype(np.float32)   # f"{cm:.1f}" / 100
    rec['lm'] = lm[src]
    packets = rec.view(np.uint8).reshape(-1, 42)
    return {
        'packets': packets,
        'agent_idx': agent[src].astype(np.int32),
        'agent_offsets': offsets,
        'grid': dict(size=grid_size, resolution=resolution, origin_x=origin[0], origin_y=origin[1]),
        'n_rooms': n_rooms,
    }


def disperse_poses(sess, seed=0):
    """Workload-sensitivity variant of a session: the same packets (ranges, yaws, agents), but every
    pose is moved to a uniformly random place of the WHOLE map, so that consecutive packets share
    neither tiles nor cells (few packets per tile, no cell written twice) — the opposite of robots
    re-scanning their rooms.  Returns a new session dict."""
    rng = np.random.default_rng(seed)
    g = sess['grid']
    span = g['size'] * g['resolution']
    rec = sess['packets'].copy().view(PACKET_DTYPE).reshape(-1)
    n = rec.shape[0]
    wx = rng.uniform(g['origin_x'] + 2.0, g['origin_x'] + span - 2.0, n)
    wy = rng.uniform(g['origin_y'] + 2.0, g['origin_y'] + span - 2.0, n)
    off = sess['agent_offsets'][sess['agent_idx']]
    rec['x'] = np.round(wx - off[:, 0], 4).astype(np.float32)
    rec['y'] = np.round(wy - off[:, 1], 4).astype(np.float32)
    out = dict(sess)
    out['packets'] = rec.view(np.uint8).reshape(-1, 42)
    return out


def beam_cell_updates(sess, packets=None):
    """Closed-form count of beam-cell updates of a session (SURVEY §8d unit) is produced by the
    device counters; this helper only reports the packet/beam totals."""
    pk = sess['packets'] if packets is None else packets
    return {'packets': int(pk.shape[0]), 'beams': int(pk.shape[0]) * 4}


# ------------------------------------------------------------------------------------------
#  Session replay format (reference simulation_tools/playback_dual_session.py:58-105)
# ------------------------------------------------------------------------------------------

def load_session(folder, time_sorted=True):
    """telemetry.csv -> list of row dicts; ranges stay in cm (:80-83), yaw in radians (:77);
    stable sort by time (:102).  (The point-cloud CSV is a by-product of the grid path and is
    not needed to replay it.)"""
    path = os.path.join(folder, 'telemetry.csv')
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    telemetry = []
    with open(path, 'r') as f:
        for row in csv.DictReader(f):
            telemetry.append({
                'time': float(row['time']), 'agent': int(row['agent']),
                'x': float(row['x']), 'y': float(row['y']),
                'yaw': math.radians(float(row['yaw_deg'])),
                'enc': int(row['encoder']), 'v2v': int(row['v2v']),
                'front': float(row['front_cm']), 'left': float(row['left_cm']),
                'back': float(row['back_cm']), 'right': float(row['right_cm']),
                'lm': int(row['landmark']),
            })
    if time_sorted:
        telemetry.sort(key=lambda r: r['time'])
    return telemetry


def rows_to_packets(rows):
    """Row dicts -> list of 42-byte v2 datagrams (cm -> m; every field rounds to its wire type)."""
    fmt = '<4sBfffiIffffB'
    return [struct.pack(fmt, b'QSRL', r['agent'], r['x'], r['y'], r['yaw'], r['enc'], r['v2v'],
                        r['front'] / 100.0, r['left'] / 100.0, r['back'] / 100.0, r['right'] / 100.0, r['lm'])
            for r in rows]
