"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Golden hit/miss count planes for the count-mode EXTENSION (the reference itself keeps
no counts).  The cells are enumerated by the UNMODIFIED reference: a subclass of its
`OccupancyGrid` (server_nodes/dual_bot_mapper.py:110-179) overrides only the two store
statements of `update_ray` (:150, :156) with `+= 1` on a miss / hit plane, and calls the
reference's own `world_to_grid`, `_bresenham` and `in_bounds`; the per-packet loop is the
reference's constants driven by oracle/occgrid_oracle.replay (pinned by make_golden.py).
Writes tests/golden/session_counts.npz (time order, SLAM on and off).

    python oracle/make_golden_counts.py          # authoring container only (/root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_loader, occgrid_oracle as O  # noqa: E402
from conftest import session_packets  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def main():
    m = ref_loader.load_dual_bot_mapper()

    class RefCountGrid(m.OccupancyGrid):
        def __init__(self):
            super().__init__()
            self.hit = np.zeros((self.size, self.size), np.int32)
            self.miss = np.zeros((self.size, self.size), np.int32)
            self.updates = 0

        def update_ray(self, robot_x, robot_y, hit_x, hit_y, hit_valid):     # :136-156 with the stores replaced
            gx0, gy0 = self.world_to_grid(robot_x, robot_y)
            gx1, gy1 = self.world_to_grid(hit_x, hit_y)
            cells = self._bresenham(gx0, gy0, gx1, gy1)
            self.updates += len(cells)
            for (cx, cy) in cells[:-1]:
                if self.in_bounds(cx, cy):
                    self.miss[cy, cx] += 1
            if hit_valid and cells:
                ex, ey = cells[-1]
                if self.in_bounds(ex, ey):
                    self.hit[ey, ex] += 1

    pk, _ = session_packets(True)
    out = {}
    for name, slam in (('slam_off', None), ('slam_on', m.PoseGraphSLAM())):
        g = RefCountGrid()
        O.replay(pk, grid=g, separation=0.0, slam=slam)
        out[f'hit_{name}'] = g.hit
        out[f'miss_{name}'] = g.miss
        print(name, int(g.hit.sum()), int(g.miss.sum()), int(g.hit.max()), int(g.miss.max()))
    np.savez_compressed(os.path.join(GOLD, 'session_counts.npz'), **out)


if __name__ == '__main__':
    main()
