"""The C-ABI library loads without a GPU and exports every function include/occgrid_b200.h
declares; host-only entry points behave (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'occgrid_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b((?:occgrid|mapmerge)_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    from occgrid_b200 import _native
    lib = _native.lib()
    names = declared_functions()
    assert len(names) >= 18, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_only_entry_points():
    from occgrid_b200 import _native
    lib = _native.lib()
    assert lib.occgrid_abi_version() == 2
    g = _native.Geom(-5.0, -5.0, 0.05, 200, 200, 0, 0, 200, 200)
    ga = lib.occgrid_workspace_bytes(g, 1000, _native.STRATEGY['global_atomic'])
    assert ga >= 200 * 200 * 4
    ti = lib.occgrid_workspace_bytes(g, 1000, _native.STRATEGY['tiled'])
    assert ti > ga
    assert lib.occgrid_workspace_bytes(g, 1000, _native.STRATEGY['auto']) == ti
    bad = _native.Geom(-5.0, -5.0, 0.05, 200, 200, 0, 0, 300, 10)
    assert lib.occgrid_workspace_bytes(bad, 1000, 0) == 0
    assert b'window' in lib.occgrid_last_error()
    fine = _native.Geom(0.0, 0.0, 0.01, 64, 64, 0, 0, 64, 64)        # 1 cm cells: smem window too large
    assert lib.occgrid_workspace_bytes(fine, 10, _native.STRATEGY['tiled']) == 0
    assert lib.occgrid_workspace_bytes(fine, 10, _native.STRATEGY['auto']) == 64 * 64 * 4
    assert lib.mapmerge_extract_workspace_bytes(2048 * 2048) > 0
    assert lib.mapmerge_voxel_workspace_bytes(1000, 1000) > 0
    assert lib.occgrid_route_workspace_bytes(1000, 8) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'distributed-multi-agent-slam-swarm-robotics-system_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'from oracle' not in src and 'import oracle' not in src, f


def test_compute_requires_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from occgrid_b200 import dual_bot_mapper as M, map_merger as MM
    with pytest.raises(M.OccGridError):
        M.OccupancyGrid()
    with pytest.raises(M.OccGridError):
        MM.MapMerger()
