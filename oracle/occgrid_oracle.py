"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (pure Python + NumPy storage) of the reference's occupancy-grid
integration path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file; the product
package (``distributed-multi-agent-slam-swarm-robotics-system_b200/``) never does.

Parity status: PINNED BY EXECUTION.  The reference ships no tests or golden
vectors, but its integration classes import under a pygame stub
(`oracle/ref_loader.py`).  `oracle/make_golden.py` runs the *unmodified* reference
`OccupancyGrid` / `PoseGraphSLAM` in the authoring container and freezes the results
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures (and, where ``/root/reference`` is present, against the
live reference objects).

Every function cites the reference lines it follows; paths are relative to
``/root/reference/``.
"""
from __future__ import annotations

import math
import struct

import numpy as np

# -- protocol: server_nodes/dual_bot_mapper.py:41-46 ---------------------------
PACKET_FMT = '<4sBfffiIffffB'
PACKET_SIZE = struct.calcsize(PACKET_FMT)          # 42
PACKET_FMT_V1 = '<4sBfffiIffff'
PACKET_SIZE_V1 = struct.calcsize(PACKET_FMT_V1)    # 41

# -- trust filter: dual_bot_mapper.py:57-58 ------------------------------------
MAX_DIST_M = 1.20
MIN_DIST_M = 0.05

# -- sensor angles, dict order front,left,back,right: dual_bot_mapper.py:61-66 -
SENSOR_ANGLES_RAD = (0.0, math.pi / 2, math.pi, -math.pi / 2)

# -- grid constants: dual_bot_mapper.py:87-94 ----------------------------------
GRID_RESOLUTION = 0.05
GRID_SIZE = 200
GRID_ORIGIN_X = -5.0
GRID_ORIGIN_Y = -5.0
CELL_UNKNOWN = -1
CELL_FREE = 0
CELL_OCCUPIED = 100

# -- SLAM constants: dual_bot_mapper.py:97-99 ----------------------------------
CLOSURE_RADIUS = 0.60
MIN_POSES_BETWEEN = 30
CLOSURE_CORRECTION = 0.5
LM_NONE = 0


class OracleGrid:
    """Restates ``OccupancyGrid`` — dual_bot_mapper.py:110-179."""

    def __init__(self, size=GRID_SIZE, resolution=GRID_RESOLUTION,
                 origin_x=GRID_ORIGIN_X, origin_y=GRID_ORIGIN_Y):
        # :113-119
        self.size = size
        self.res = resolution
        self.ox = origin_x
        self.oy = origin_y
        self.grid = np.full((size, size), CELL_UNKNOWN, dtype=np.int8)
        self.updates = 0   # beam-cell updates (SURVEY §8d unit); not in the reference

    def world_to_grid(self, wx, wy):
        # :121-125 — true division by res, int() truncates toward zero
        gx = int((wx - self.ox) / self.res)
        gy = int((wy - self.oy) / self.res)
        return gx, gy

    def grid_to_world(self, gx, gy):
        # :127-131 — cell centre
        return self.ox + (gx + 0.5) * self.res, self.oy + (gy + 0.5) * self.res

    def in_bounds(self, gx, gy):
        # :133-134
        return 0 <= gx < self.size and 0 <= gy < self.size

    @staticmethod
    def bresenham(x0, y0, x1, y1):
        # :158-179
        cells = []
        dx = abs(x1 - x0)
        dy = abs(y1 - y0)
        sx = 1 if x0 < x1 else -1
        sy = 1 if y0 < y1 else -1
        err = dx - dy
        while True:
            cells.append((x0, y0))
            if x0 == x1 and y0 == y1:
                break
            e2 = 2 * err
            if e2 > -dy:
                err -= dy
                x0 += sx
            if e2 < dx:
                err += dx
                y0 += sy
        return cells

    _bresenham = bresenham

    def update_ray(self, robot_x, robot_y, hit_x, hit_y, hit_valid):
        # :136-156
        x0, y0 = self.world_to_grid(robot_x, robot_y)
        x1, y1 = self.world_to_grid(hit_x, hit_y)
        cells = self.bresenham(x0, y0, x1, y1)
        self.updates += len(cells)
        g = self.grid
        n = self.size
        for gx, gy in cells[:-1]:
            if 0 <= gx < n and 0 <= gy < n:
                g[gy, gx] = CELL_FREE
        if cells and hit_valid:
            gx, gy = cells[-1]
            if 0 <= gx < n and 0 <= gy < n:
                g[gy, gx] = CELL_OCCUPIED


FRONTIER_MIN_CLUSTER = 3      # dual_bot_mapper.py:102


def get_frontiers(grid_array):
    """OccupancyGrid.get_frontiers — dual_bot_mapper.py:181-197 (loop restated verbatim)."""
    g = grid_array
    size = g.shape[0]
    frontiers = []
    for y in range(1, size - 1):
        for x in range(1, size - 1):
            if g[y, x] != CELL_FREE:
                continue
            for dx, dy in [(-1, 0), (1, 0), (0, -1), (0, 1)]:
                if g[y + dy, x + dx] == CELL_UNKNOWN:
                    frontiers.append((x, y))
                    break
    return frontiers


def get_frontiers_fast(grid_array):
    """Vectorised equivalent of get_frontiers for large grids (checked against it in the tests)."""
    g = np.asarray(grid_array)
    free = g == CELL_FREE
    unk = g == CELL_UNKNOWN
    nb = np.zeros_like(free)
    nb[1:-1, 1:-1] = unk[1:-1, :-2] | unk[1:-1, 2:] | unk[:-2, 1:-1] | unk[2:, 1:-1]
    inner = np.zeros_like(free)
    inner[1:-1, 1:-1] = True
    ys, xs = np.nonzero(free & nb & inner)
    return list(zip(xs.tolist(), ys.tolist()))


def cluster_frontiers(frontier_cells):
    """OccupancyGrid.cluster_frontiers — dual_bot_mapper.py:199-231 (BFS flood fill)."""
    if not frontier_cells:
        return []
    cell_set = set(frontier_cells)
    visited = set()
    clusters = []
    for cell in frontier_cells:
        if cell in visited:
            continue
        cluster = []
        queue = [cell]
        while queue:
            c = queue.pop(0)
            if c in visited:
                continue
            visited.add(c)
            cluster.append(c)
            cx, cy = c
            for dx, dy in [(-1, 0), (1, 0), (0, -1), (0, 1)]:
                nbr = (cx + dx, cy + dy)
                if nbr in cell_set and nbr not in visited:
                    queue.append(nbr)
        if len(cluster) >= FRONTIER_MIN_CLUSTER:
            clusters.append(cluster)
    return clusters


def cluster_centroid_world(cluster, ox, oy, res):
    """dual_bot_mapper.py:233-237 + grid_to_world :127-131."""
    avg_x = sum(c[0] for c in cluster) / len(cluster)
    avg_y = sum(c[1] for c in cluster) / len(cluster)
    return ox + (avg_x + 0.5) * res, oy + (avg_y + 0.5) * res


class OracleSLAM:
    """Restates ``PoseGraphSLAM`` — dual_bot_mapper.py:244-338 (prints dropped).

    Sequential host logic; it only matters to the hot path as the producer of the
    per-packet drift input (dual_bot_mapper.py:855-857, 908-914).
    """

    def __init__(self):
        self.nodes = []        # (x, y, yaw, agent_id, landmark_type, timestamp)
        self.landmarks = []    # (x, y, landmark_type, node_index)
        self.closures = []     # (lm_idx, node_idx, dx, dy)
        self.last_closure_idx = {1: -MIN_POSES_BETWEEN, 2: -MIN_POSES_BETWEEN}

    def add_pose(self, x, y, yaw, agent_id, landmark_type, timestamp):
        # :261-280
        idx = len(self.nodes)
        self.nodes.append((x, y, yaw, agent_id, landmark_type, timestamp))
        closure, cdx, cdy = False, 0.0, 0.0
        if landmark_type != LM_NONE:
            closure, cdx, cdy = self._check_closure(x, y, agent_id, landmark_type, idx)
            self.landmarks.append((x, y, landmark_type, idx))
        return closure, cdx, cdy

    def _check_closure(self, nx, ny, agent_id, landmark_type, index):
        # :282-322
        for lm_x, lm_y, lm_type, lm_idx in self.landmarks:
            if lm_type != landmark_type:
                continue
            if index - lm_idx < MIN_POSES_BETWEEN:
                continue
            if index - self.last_closure_idx.get(agent_id, -999) < MIN_POSES_BETWEEN:
                continue
            dist = math.sqrt((nx - lm_x) ** 2 + (ny - lm_y) ** 2)
            if dist < CLOSURE_RADIUS:
                error_x = lm_x - nx
                error_y = lm_y - ny
                cdx = error_x * CLOSURE_CORRECTION
                cdy = error_y * CLOSURE_CORRECTION
                self.closures.append((lm_idx, index, cdx, cdy))
                self.last_closure_idx[agent_id] = index
                return True, cdx, cdy
        return False, 0.0, 0.0


def decode_packet(data):
    """dual_bot_mapper.py:826-843.  Returns the 12 fields or None if dropped."""
    if len(data) == PACKET_SIZE:
        u = struct.unpack(PACKET_FMT, data)
        magic, agent_id, rx, ry, ryaw, enc, v2v, df, dl, db, dr, lm = u
    elif len(data) == PACKET_SIZE_V1:
        u = struct.unpack(PACKET_FMT_V1, data)
        magic, agent_id, rx, ry, ryaw, enc, v2v, df, dl, db, dr = u
        lm = LM_NONE
    else:
        return None
    if magic != b'QSRL':
        return None
    return magic, agent_id, rx, ry, ryaw, enc, v2v, df, dl, db, dr, lm


def expand_beams(rx, ry, ryaw, dists):
    """dual_bot_mapper.py:881-903 — 4 beams (x0,y0,x1,y1,hit_valid) in sensor order."""
    out = []
    for ang, dist in zip(SENSOR_ANGLES_RAD, dists):
        ray_angle = ryaw + ang                                   # :887
        hit_valid = MIN_DIST_M < dist <= MAX_DIST_M              # :888
        if hit_valid:
            wx = rx + dist * math.cos(ray_angle)                 # :890
            wy = ry + dist * math.sin(ray_angle)                 # :891
            out.append((rx, ry, wx, wy, True))
        else:
            max_range = min(dist, MAX_DIST_M) if dist > MIN_DIST_M else MAX_DIST_M   # :900
            ex = rx + max_range * math.cos(ray_angle)            # :901
            ey = ry + max_range * math.sin(ray_angle)            # :902
            out.append((rx, ry, ex, ey, False))
    return out


def pose_is_integrable(rx, ry, ryaw):
    """The reference crashes (``int(nan)`` ValueError, :123) on non-finite poses.  The
    build skips such packets instead (documented divergence, SURVEY §8b 'Errors')."""
    return math.isfinite(rx) and math.isfinite(ry) and math.isfinite(ryaw)


def replay(packets, grid=None, separation=0.0, slam=None, agent_offsets=None,
           drift_out=None, timestamps=None):
    """Restates the per-packet loop of ``main()`` — dual_bot_mapper.py:826-919.

    packets       iterable of bytes objects (one datagram each)
    separation    --separation CLI flag (:716, :851-852)
    slam          OracleSLAM (or the reference's PoseGraphSLAM) or None for drift-free replay
    agent_offsets EXTENSION (no reference counterpart): {agent_id: (ox, oy)} accepted ids
                  and their start offsets.  Default = the reference's rule: ids {1,2},
                  only agent 2 shifted by ``separation`` along x.
    drift_out     optional list; receives the (cdx, cdy) that was applied to each
                  *accepted or dropped* packet (dropped packets get (0,0)) so that the
                  device path can be fed the same per-packet drift table.
    """
    if grid is None:
        grid = OracleGrid()
    if agent_offsets is None:
        agent_offsets = {1: (0.0, 0.0), 2: (separation, 0.0)}
    drift = {a: (0.0, 0.0) for a in agent_offsets}               # :782
    stats = {'packets': 0, 'accepted': 0, 'dropped': 0, 'bad_pose': 0, 'beams': 0, 'hits': 0}
    for k, data in enumerate(packets):
        stats['packets'] += 1
        f = decode_packet(data)
        if f is None or f[1] not in agent_offsets:               # :838-843
            stats['dropped'] += 1
            if drift_out is not None:
                drift_out.append((0.0, 0.0))
            continue
        _, agent_id, rx, ry, ryaw, enc, v2v, df, dl, db, dr, lm = f
        offx, offy = agent_offsets[agent_id]
        rx += offx                                               # :851-852
        ry += offy                                               # extension; +0.0 in reference mode
        cdx, cdy = drift[agent_id]                               # :855
        if drift_out is not None:
            drift_out.append((cdx, cdy))
        rx += cdx                                                # :856
        ry += cdy                                                # :857
        if not pose_is_integrable(rx, ry, ryaw):
            stats['bad_pose'] += 1
            continue
        stats['accepted'] += 1
        for (x0, y0, x1, y1, hv) in expand_beams(rx, ry, ryaw, (df, dl, db, dr)):
            if not (math.isfinite(x1) and math.isfinite(y1)):
                continue
            grid.update_ray(x0, y0, x1, y1, hv)                  # :897 / :903
            stats['beams'] += 1
            stats['hits'] += int(hv)
        if slam is not None:                                     # :908-914
            now = timestamps[k] if timestamps is not None else float(k)
            closure, cdx_new, cdy_new = slam.add_pose(rx, ry, ryaw, agent_id, lm, now)
            if closure:
                drift[agent_id] = (drift[agent_id][0] + cdx_new,
                                   drift[agent_id][1] + cdy_new)
    stats['updates'] = grid.updates
    return grid, stats


# ------------------------------------------------------------------------------
#  Session CSV -> packets (replay format)
# ------------------------------------------------------------------------------

def load_session_rows(telem_csv_path, time_sorted=True):
    """Restates ``load_session`` — simulation_tools/playback_dual_session.py:58-105
    (telemetry half).  Ranges stay in cm as in the reference (:80-83); yaw becomes
    radians (:77).  ``time_sorted`` applies the stable sort at :102."""
    import csv
    rows = []
    with open(telem_csv_path, 'r') as f:
        for row in csv.DictReader(f):
            rows.append({
                'time': float(row['time']), 'agent': int(row['agent']),
                'x': float(row['x']), 'y': float(row['y']),
                'yaw': math.radians(float(row['yaw_deg'])),
                'enc': int(row['encoder']), 'v2v': int(row['v2v']),
                'front': float(row['front_cm']), 'left': float(row['left_cm']),
                'back': float(row['back_cm']), 'right': float(row['right_cm']),
                'lm': int(row['landmark']),
            })
    if time_sorted:
        rows.sort(key=lambda r: r['time'])
    return rows


def rows_to_packets(rows):
    """pack each CSV row into the v2 wire format
    (dual_bot_mapper.py:41-42; firmware struct AgentFirmware_Bot1.ino:172-185)."""
    return [struct.pack(PACKET_FMT, b'QSRL', r['agent'], r['x'], r['y'], r['yaw'],
                        r['enc'], r['v2v'], r['front'] / 100.0, r['left'] / 100.0,
                        r['back'] / 100.0, r['right'] / 100.0, r['lm'])
            for r in rows]


def grid_census(grid_array):
    import hashlib
    g = np.ascontiguousarray(grid_array, dtype=np.int8)
    return {
        'free': int((g == CELL_FREE).sum()),
        'occ': int((g == CELL_OCCUPIED).sum()),
        'unk': int((g == CELL_UNKNOWN).sum()),
        'sha1': hashlib.sha1(g.tobytes()).hexdigest(),
    }
