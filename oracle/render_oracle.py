"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement of the occupancy overlay of the reference's renderer,
server_nodes/dual_bot_mapper.py:492-527 (`MapRenderer._draw_occupancy`) with
`world_to_screen` (:404-408) and `grid_to_world` (:127-131), drawing into an RGB array instead
of a PyGame surface.  PyGame's two primitives on this path are restated from their documented
behaviour: `Surface.set_at` outside the surface has no effect; `draw.rect(surface, color, rect)`
fills the rectangle clipped to the surface.  Pinned by tests/golden/render_ref.npz, which
oracle/make_golden_render.py produced by EXECUTING the unmodified method against a recording
surface.
"""
import numpy as np

CELL_UNKNOWN, CELL_FREE, CELL_OCCUPIED = -1, 0, 100
CELL_COLOR_FREE = (30, 45, 70)          # :373
BG_COLOR = (22, 33, 62)                 # :346


def draw_occupancy(grid, ox, oy, res, scale=100.0, offset_x=None, offset_y=None, width=1000, height=800,
                   background=BG_COLOR, color=CELL_COLOR_FREE):
    """-> uint8 [height, width, 3]: `background` with the FREE cells painted (:492-527)."""
    size = grid.shape[0]
    offset_x = width / 2 if offset_x is None else offset_x               # :396-397
    offset_y = height / 2 if offset_y is None else offset_y
    img = np.empty((height, width, 3), np.uint8)
    img[:] = np.asarray(background, np.uint8)
    cell_px = max(1, int(res * scale))                                    # :494
    if cell_px < 2:                                                       # :495-496
        return img
    world_left = -offset_x / scale                                        # :500-503
    world_right = (width - offset_x) / scale
    world_top = offset_y / scale
    world_bottom = -(height - offset_y) / scale
    gx_min = max(0, int((world_left - ox) / res) - 1)                     # :505-508
    gx_max = min(size, int((world_right - ox) / res) + 1)
    gy_min = max(0, int((world_bottom - oy) / res) - 1)
    gy_max = min(size, int((world_top - oy) / res) + 1)
    for gy in range(gy_min, gy_max):                                      # :510-511
        for gx in range(gx_min, gx_max):
            val = grid[gy, gx]
            if val == CELL_UNKNOWN or val == CELL_OCCUPIED:               # :513-520
                continue
            wx = ox + (gx + 0.5) * res                                    # :129-130
            wy = oy + (gy + 0.5) * res
            sx = int(offset_x + wx * scale)                               # :406-407
            sy = int(offset_y - wy * scale)
            if cell_px <= 2:                                              # :523-524
                if 0 <= sx < width and 0 <= sy < height:
                    img[sy, sx] = color
            else:                                                         # :525-527
                x0, y0 = sx - cell_px // 2, sy - cell_px // 2
                xa, xb = max(0, x0), min(width, x0 + cell_px)
                ya, yb = max(0, y0), min(height, y0 + cell_px)
                if xa < xb and ya < yb:
                    img[ya:yb, xa:xb] = color
    return img
