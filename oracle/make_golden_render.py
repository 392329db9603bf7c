"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Freezes overlay-render fixtures (second half) by EXECUTING the unmodified
`MapRenderer._draw_occupancy` / `world_to_screen` of server_nodes/dual_bot_mapper.py (:404-408,
:492-527) in the authoring container.  PyGame is absent, so the two primitives the method calls
are supplied by a recording surface: `set_at` (no effect outside the surface) and
`pygame.draw.rect` (fill clipped to the surface) — their documented behaviour; everything else
(cell size, visible ranges, cell centres, truncations, which values are skipped) is the
reference's own code.  Writes tests/golden/render_ref.npz.   Run: python oracle/make_golden_render.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_loader, occgrid_oracle as O  # noqa: E402
from conftest import session_packets  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


class RecordingSurface:
    def __init__(self, width, height, bg):
        self.img = np.empty((height, width, 3), np.uint8)
        self.img[:] = np.asarray(bg, np.uint8)

    def set_at(self, pos, color):
        x, y = pos
        h, w, _ = self.img.shape
        if 0 <= x < w and 0 <= y < h:
            self.img[y, x] = color

    def fill_rect(self, color, rect):
        x, y, rw, rh = rect
        h, w, _ = self.img.shape
        xa, xb, ya, yb = max(0, x), min(w, x + rw), max(0, y), min(h, y + rh)
        if xa < xb and ya < yb:
            self.img[ya:yb, xa:xb] = color


def render_with_reference(m, occ, width, height, scale, offset_x, offset_y):
    R = types.SimpleNamespace(width=width, height=height, scale=scale, offset_x=offset_x, offset_y=offset_y,
                              screen=RecordingSurface(width, height, m.BG_COLOR))
    R.world_to_screen = lambda wx, wy: m.MapRenderer.world_to_screen(R, wx, wy)
    m.pygame.draw = types.SimpleNamespace(rect=lambda surface, color, rect: surface.fill_rect(color, rect))
    m.MapRenderer._draw_occupancy(R, occ)
    return R.screen.img


def main():
    m = ref_loader.load_dual_bot_mapper()
    pk, _ = session_packets(True)
    g, _ = O.replay(pk)
    occ = m.OccupancyGrid()
    occ.grid[:] = g.grid
    out = {}
    cases = [  # (name, grid object, width, height, scale, offset_x, offset_y)
        ('session_default', occ, 1000, 800, 100.0, 500.0, 400.0),             # the renderer's start view (:395-397)
        ('session_zoom_out', occ, 640, 480, 45.0, 200.5, 300.25),             # cell_px == 2: single pixels (:523-524)
        ('session_too_small', occ, 320, 240, 30.0, 160.0, 120.0),             # cell_px < 2: nothing drawn (:495-496)
        ('session_zoom_in_pan', occ, 400, 300, 260.0, -350.0, 820.0),         # large rects, view mostly off the map
    ]
    rng = np.random.default_rng(8)
    occ2 = m.OccupancyGrid(size=96, resolution=0.1, origin_x=-3.3, origin_y=7.7)
    occ2.grid[:] = rng.choice(np.array([-1, 0, 100], np.int8), size=(96, 96), p=[0.4, 0.5, 0.1])
    cases.append(('random96_res01', occ2, 512, 384, 37.5, 180.0, 600.0))
    for name, o, w, h, sc, ofx, ofy in cases:
        img = render_with_reference(m, o, w, h, sc, ofx, ofy)
        out[f'{name}/img'] = img
        out[f'{name}/view'] = np.array([w, h, sc, ofx, ofy], np.float64)
        out[f'{name}/geom'] = np.array([o.size, o.res, o.ox, o.oy], np.float64)
        out[f'{name}/grid'] = o.grid.copy()
        print(name, img.shape, int((img != np.asarray(m.BG_COLOR, np.uint8)).any(axis=2).sum()), 'painted pixels')
    out['colors'] = np.array([m.BG_COLOR, m.CELL_COLOR_FREE], np.uint8)
    np.savez_compressed(os.path.join(GOLD, 'render_ref.npz'), **out)
    print(os.path.getsize(os.path.join(GOLD, 'render_ref.npz')), 'bytes')


if __name__ == '__main__':
    main()
