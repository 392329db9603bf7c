"""Overlay render (server_nodes/dual_bot_mapper.py:492-527).  The fixtures in
tests/golden/render_ref.npz were produced by EXECUTING the unmodified MapRenderer._draw_occupancy
against a recording surface (oracle/make_golden_render.py).  CPU: the restatement equals them;
GPU: occgrid_render_overlay equals them byte for byte."""
import os

import numpy as np
import pytest

from conftest import GOLD

CASES = ['session_default', 'session_zoom_out', 'session_too_small', 'session_zoom_in_pan', 'random96_res01']


def _load():
    return np.load(os.path.join(GOLD, 'render_ref.npz'))


@pytest.mark.parametrize('name', CASES)
def test_restatement_equals_reference_executed_render(name):
    from oracle import render_oracle as RO
    z = _load()
    w, h, sc, ofx, ofy = z[f'{name}/view'].tolist()
    size, res, ox, oy = z[f'{name}/geom'].tolist()
    assert tuple(z['colors'][0]) == RO.BG_COLOR and tuple(z['colors'][1]) == RO.CELL_COLOR_FREE
    img = RO.draw_occupancy(z[f'{name}/grid'], ox, oy, res, sc, ofx, ofy, int(w), int(h))
    assert np.array_equal(img, z[f'{name}/img'])
    if name == 'session_too_small':
        assert (img == np.asarray(RO.BG_COLOR, np.uint8)).all()          # cell_px < 2: nothing is drawn (:495-496)
    if name == 'session_zoom_out':                                       # cell_px == 2: one pixel per FREE cell (:523-524)
        assert (img != np.asarray(RO.BG_COLOR, np.uint8)).any(axis=2).sum() == (z[f'{name}/grid'] == 0).sum()


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_device_render_equals_reference_executed_render(name):
    torch = pytest.importorskip('torch')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import dual_bot_mapper as M
    z = _load()
    w, h, sc, ofx, ofy = z[f'{name}/view'].tolist()
    size, res, ox, oy = z[f'{name}/geom'].tolist()
    g = M.OccupancyGrid(int(size), res, ox, oy)
    g.grid_tensor.copy_(torch.from_numpy(z[f'{name}/grid']))
    img = g.render_overlay(int(w), int(h), sc, ofx, ofy)
    assert img.dtype == np.uint8 and np.array_equal(img, z[f'{name}/img'])
    assert (M.BG_COLOR, M.CELL_COLOR_FREE) == (tuple(z['colors'][0]), tuple(z['colors'][1]))


@pytest.mark.gpu
def test_device_render_large_map_against_restatement():
    """4096^2 map after a packet batch, default view and a panned zoom: CUDA == restatement."""
    torch = pytest.importorskip('torch')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from occgrid_b200 import dual_bot_mapper as M, simulation_tools as st
    from oracle import render_oracle as RO
    s = st.generate_session(n_agents=8, n_packets=40_000, grid_size=1024, origin=(-25.6, -25.6), seed=5)
    g = M.OccupancyGrid(max_batch=40_000, **s['grid'])
    g.update_packets(s['packets'], agent_offsets=s['agent_offsets'])
    for view in [(1000, 800, 100.0, None, None), (700, 500, 61.0, 900.0, -300.0), (256, 256, 499.9, 2000.5, 1000.25)]:
        w, h, sc, ofx, ofy = view
        want = RO.draw_occupancy(g.grid, g.ox, g.oy, g.res, sc, ofx, ofy, w, h)
        assert np.array_equal(g.render_overlay(w, h, sc, ofx, ofy), want), view
