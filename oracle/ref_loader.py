"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference modules from ``/root/reference`` (authoring container
only; the GPU box has no ``/root/reference``).  ``dual_bot_mapper`` and
``playback_dual_session`` hard-exit when pygame is missing
(server_nodes/dual_bot_mapper.py:29-34, simulation_tools/playback_dual_session.py:27-32),
so an empty ``pygame`` / ``pygame.gfxdraw`` stub is injected first.  Nothing on the
integration path touches pygame.
"""
import contextlib
import importlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('OCCGRID_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'server_nodes', 'dual_bot_mapper.py'))


def _stub_pygame():
    if 'pygame' not in sys.modules:
        pg = types.ModuleType('pygame')
        pg.gfxdraw = types.ModuleType('pygame.gfxdraw')
        sys.modules['pygame'] = pg
        sys.modules['pygame.gfxdraw'] = pg.gfxdraw


def load_dual_bot_mapper():
    """Returns the reference ``dual_bot_mapper`` module (OccupancyGrid, PoseGraphSLAM, …)."""
    _stub_pygame()
    p = os.path.join(REFERENCE_ROOT, 'server_nodes')
    if p not in sys.path:
        sys.path.insert(0, p)
    return importlib.import_module('dual_bot_mapper')


def load_playback():
    _stub_pygame()
    p = os.path.join(REFERENCE_ROOT, 'simulation_tools')
    if p not in sys.path:
        sys.path.insert(0, p)
    return importlib.import_module('playback_dual_session')


class QuietSLAM:
    """Wraps the reference PoseGraphSLAM, swallowing its per-closure print (:316-318)."""

    def __init__(self, ref_module):
        self._s = ref_module.PoseGraphSLAM()

    def add_pose(self, *a):
        with contextlib.redirect_stdout(io.StringIO()):
            return self._s.add_pose(*a)

    @property
    def closures(self):
        return self._s.closures
