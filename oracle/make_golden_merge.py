"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Freezes merge fixtures by EXECUTING the unmodified reference
`server_nodes/map_merger.py` in the authoring container under rclpy / nav_msgs / open3d stubs
(oracle/ref_loader.load_map_merger).  What the reference itself computes here:

  * `MapMerger.grid_to_pcd`        (:64-85)   — np.array(msg.data).reshape, argwhere(> 50), corners
  * `MapMerger.map_callback`       (:35-62)   — empty return, first-map adoption, fitness gate,
                                                transform -> += -> voxel_down_sample sequencing
  * `MapMerger.publish_global_map` (:87-127)  — bounds, ceil+1 extents, astype(int), clip, scatter,
                                                frame id, flatten().tolist()

What stays restated (the library is absent and un-pinned): the arithmetic inside
`PointCloud.transform`, `+=`, `voxel_down_sample` (oracle/merge_oracle.py) and
`registration_icp` (oracle/icp_oracle.py, used by the 'icp' sequence; the other sequences
inject the transform the way BASELINE configs[2] supplies it).

Writes tests/golden/merge_ref.npz.   Run:  python oracle/make_golden_merge.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_loader, merge_oracle as MO  # noqa: E402
from merge_util import synth_agent_grid  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def partial_view(world, seed, n):
    """An agent's view of a common world: a window of it, moved by a small rigid motion that
    ICP has to find (the agent's map frame is off by (dx, dy, dth))."""
    r = np.random.default_rng(seed)
    g = np.full((n, n), -1, np.int8)
    x0, y0 = (int(v) for v in r.integers(0, n // 6, 2))
    w = n - n // 6
    g[y0:y0 + w, x0:x0 + w] = world[y0:y0 + w, x0:x0 + w]
    return g


def run_sequence(mod, script, name, steps, out):
    """steps: list of dicts {grid, res, ox, oy, T (4x4) | None, fitness}.  Drives the reference
    node's own map_callback and records what it published and the cloud it holds."""
    node = mod.MapMerger()
    script.queue.clear()
    script.calls.clear()
    for k, s in enumerate(steps):
        g = s['grid']
        h, w = g.shape
        n_pub = len(node.published)
        have_cloud = not node.global_pcd.is_empty()
        msg = ref_loader.make_ref_grid_msg(mod, g.ravel(), w, h, s['res'], s['ox'], s['oy'])
        if script.mode == 'script' and have_cloud and (g > 50).any():
            script.queue.append((s['T'], s['fitness']))
        # (:64-85) on its own, as the reference computes it
        pts = np.asarray(mod.MapMerger.grid_to_pcd(node, msg).points, np.float64).reshape(-1, 3)
        node.map_callback(msg, k + 1)
        published = len(node.published) > n_pub
        out[f'{name}/{k}/grid'] = g
        out[f'{name}/{k}/geom'] = np.array([s['res'], s['ox'], s['oy']])
        out[f'{name}/{k}/T'] = np.asarray(s['T'] if s['T'] is not None else np.eye(4), np.float64)
        out[f'{name}/{k}/fitness'] = np.float64(s['fitness'])
        out[f'{name}/{k}/pcd'] = pts
        out[f'{name}/{k}/published'] = np.bool_(published)
        if published:
            m = node.published[-1]
            assert m.header.frame_id == 'map_global'
            out[f'{name}/{k}/out'] = np.array(m.data, np.int8).reshape(m.info.height, m.info.width)
            out[f'{name}/{k}/out_origin'] = np.array([m.info.origin.position.x, m.info.origin.position.y])
            out[f'{name}/{k}/out_res'] = np.float64(m.info.resolution)
        out[f'{name}/{k}/cloud'] = np.asarray(node.global_pcd.points, np.float64).reshape(-1, 3)[:, :2].copy()
        out[f'{name}/{k}/state'] = np.array([node.map_resolution, node.map_origin[0], node.map_origin[1]])
    if script.mode == 'icp':
        out[f'{name}/icp_calls'] = np.array(script.calls, np.float64).reshape(-1, 4)
        return
    assert not script.queue, 'a scripted registration result was not consumed'
    out[f'{name}/icp_calls'] = np.array(script.calls, np.float64).reshape(-1, 4)


def main():
    mod, script = ref_loader.load_map_merger()
    out = {}
    r = np.random.default_rng(2026)

    # A: six callbacks, random SE(2), one rejected (fitness 0.59 < 0.6, :54-56), one empty grid (:37-38)
    steps = []
    for a in range(7):
        g = synth_agent_grid(96, 500 + a) if a != 4 else np.full((96, 96), -1, np.int8)
        T = MO.se2_matrix(*r.uniform(-3, 3, 2), r.uniform(-math.pi, math.pi))
        steps.append(dict(grid=g, res=0.05, ox=-2.4, oy=-2.4, T=T, fitness=0.59 if a == 2 else 1.0))
    script.mode = 'script'
    run_sequence(mod, script, 'A', steps, out)

    # B: other resolution and a far-away origin; rectangular grids; first message is empty
    steps = [dict(grid=np.full((20, 30), 0, np.int8), res=0.1, ox=12.3, oy=-40.7, T=None, fitness=1.0)]
    for a in range(4):
        g = synth_agent_grid(120, 700 + a)[:90, :]
        T = MO.se2_matrix(*r.uniform(-6, 6, 2), r.uniform(-math.pi, math.pi))
        steps.append(dict(grid=g, res=0.1, ox=12.3, oy=-40.7, T=T, fitness=[0.5, 0.6, 0.75, 0.9][a]))
    run_sequence(mod, script, 'B', steps, out)

    # C: heavy overlap (identity and tiny motions: many points per voxel, means move)
    base = synth_agent_grid(80, 900)
    steps = []
    for a in range(5):
        T = MO.se2_matrix(*(r.uniform(-0.04, 0.04, 2) if a else (0.0, 0.0)), r.uniform(-0.01, 0.01) if a else 0.0)
        steps.append(dict(grid=base, res=0.05, ox=-2.0, oy=-2.0, T=T, fitness=0.9))
    run_sequence(mod, script, 'C', steps, out)

    # D: the reference's own callback drives the (restated) registration on partial views
    world = synth_agent_grid(160, 1100, occ_segments=40)
    steps = []
    for a in range(4):
        g = partial_view(world, 1200 + a, 160)
        d = r.uniform(-0.12, 0.12, 2)
        steps.append(dict(grid=g, res=0.05, ox=-4.0 + (d[0] if a else 0.0), oy=-4.0 + (d[1] if a else 0.0),
                          T=None, fitness=1.0))
    script.mode = 'icp'
    run_sequence(mod, script, 'D', steps, out)
    script.mode = 'script'

    # publish_global_map (:87-127) alone on hand-made clouds: clip and ceil+1 behaviour
    for j, pts in enumerate([np.array([[0.0, 0.0, 0.0], [0.05, 0.1, 0.0], [0.149999, 0.25, 0.0]]),
                             np.array([[1.0, -1.0, 0.0]]),
                             np.c_[r.uniform(-3, 3, (400, 2)), np.zeros(400)]]):
        node = mod.MapMerger()
        node.global_pcd = mod.o3d.geometry.PointCloud()
        node.global_pcd.points = pts
        node.map_resolution = [0.05, 0.05, 0.1][j]
        node.publish_global_map('map')
        m = node.published[-1]
        out[f'P/{j}/points'] = pts
        out[f'P/{j}/res'] = np.float64(node.map_resolution)
        out[f'P/{j}/out'] = np.array(m.data, np.int8).reshape(m.info.height, m.info.width)
        out[f'P/{j}/out_origin'] = np.array([m.info.origin.position.x, m.info.origin.position.y])

    np.savez_compressed(os.path.join(GOLD, 'merge_ref.npz'), **out)
    n_pub = sum(1 for k in out if k.endswith('/published') and bool(out[k]))
    print(f'merge_ref.npz: {len(out)} arrays, {n_pub} published maps, '
          f'{os.path.getsize(os.path.join(GOLD, "merge_ref.npz"))} bytes')
    for s in 'ABCD':
        print(s, 'icp calls', out[f'{s}/icp_calls'].tolist())


if __name__ == '__main__':
    main()
