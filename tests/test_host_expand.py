"""CPU checks of the product's host+device arithmetic headers (csrc/sincos_dd.cuh,
csrc/beam_expand.cuh), compiled for the host by g++:

* sincos_dd is correctly rounded (vs mpmath) — it is what the CUDA path falls back to whenever
  a beam endpoint sits within the screening tolerance of a cell boundary;
* the packet expansion (one library sincos per packet + 1/res multiply, screened, with exact
  re-evaluation) returns the reference's cells even when the library sincos is off by 2 ulp.
"""
import ctypes as C
import math
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_packet_stream, normalise_datagrams, session_packets
from oracle import occgrid_oracle as O

SRC = os.path.join(ROOT, 'tests', 'host_harness.cpp')
LIB = os.path.join(ROOT, 'tests', '_build', 'libhost_harness.so')


@pytest.fixture(scope='module')
def hh():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, 'distributed-multi-agent-slam-swarm-robotics-system_b200', 'csrc', f)
                    for f in ('beam_expand.cuh', 'sincos_dd.cuh')]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fno-fast-math', '-x', 'c++', '-shared', '-fPIC',
                        '-o', LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.hh_sincos_dd.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hh_expand_packets.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    L.hh_bresenham.restype = C.c_int64
    L.hh_bresenham.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int64]
    return L


def dd(L, x):
    s, c = C.c_double(), C.c_double()
    ok = L.hh_sincos_dd(x, C.byref(s), C.byref(c))
    return ok, s.value, c.value


def test_sincos_dd_correctly_rounded(hh):
    import mpmath as mp
    mp.mp.prec = 300
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(-7, 7, 4000)) + list(rng.uniform(-1e3, 1e3, 1500)) + list(rng.uniform(-1e6, 1e6, 1500))
    xs += [float(np.float32(k * math.pi / 12)) + a for k in range(-24, 25) for a in (0.0, math.pi / 2, math.pi, -math.pi / 2)]
    xs += [0.0, 1e-300, -1e-10, 1e-5, 0.5, math.pi / 4, math.pi / 2, math.pi, 3 * math.pi / 2, 2 * math.pi, 1e6, -1e6]
    glibc_diff = 0
    for x in xs:
        ok, s, c = dd(hh, float(x))
        assert ok
        ws, wc = float(mp.sin(mp.mpf(float(x)))), float(mp.cos(mp.mpf(float(x))))
        assert s == ws, (x, s, ws)
        assert c == wc, (x, c, wc)
        glibc_diff += (math.sin(x) != ws) + (math.cos(x) != wc)
    # glibc is correctly rounded in the vast majority of cases (bound 0.55 ulp)
    assert glibc_diff <= 0.02 * 2 * len(xs)
    assert dd(hh, 1.0e6 + 1)[0] == 0 and dd(hh, float('nan'))[0] == 0 and dd(hh, float('inf'))[0] == 0


def test_bresenham_walk_matches_reference_kat(hh):
    z = np.load(os.path.join(GOLD, 'bresenham_kat.npz'))
    xs, ys, offs = z['x'], z['y'], z['offsets']
    bx, by = np.empty(80, np.int32), np.empty(80, np.int32)
    i = 0
    for dy in range(-32, 33):
        for dx in range(-32, 33):
            n = hh.hh_bresenham(5, -7, 5 + dx, -7 + dy, bx.ctypes.data, by.ctypes.data, 80)
            assert n == offs[i + 1] - offs[i]
            assert np.array_equal(bx[:n] - 5, xs[offs[i]:offs[i + 1]]) and np.array_equal(by[:n] + 7, ys[offs[i]:offs[i + 1]])
            i += 1


def expand(hh, arr, ox, oy, res, mode, separation=0.0, drift=None, agent_offsets=None):
    n = arr.shape[0]
    if agent_offsets is None:
        agent_offsets = np.array([[0.0, 0.0], [0.0, 0.0], [separation, 0.0]])
    agent_offsets = np.ascontiguousarray(agent_offsets, np.float64)
    out = np.zeros((n, 4, 7), np.int32)
    st = np.zeros(n, np.int32)
    d = None if drift is None else np.ascontiguousarray(drift, np.float64)
    hh.hh_expand_packets(arr.ctypes.data, n, arr.shape[1], None, d.ctypes.data if d is not None else None,
                         agent_offsets.ctypes.data, agent_offsets.shape[0] - 1, ox, oy, res, mode,
                         out.ctypes.data, st.ctypes.data)
    return out, st


def oracle_cells(arr, ox, oy, res, separation=0.0, drift=None, agent_offsets=None):
    g = O.OracleGrid(size=8, resolution=res, origin_x=ox, origin_y=oy)
    offs = agent_offsets if agent_offsets is not None else {1: (0.0, 0.0), 2: (separation, 0.0)}
    want = {}
    for k in range(arr.shape[0]):
        f = O.decode_packet(arr[k].tobytes())
        if f is None or f[1] not in offs:
            continue
        _, a, rx, ry, ryaw, _, _, df, dl, db, dr, _ = f
        rx += offs[a][0]
        ry += offs[a][1]
        if drift is not None:
            rx += drift[k][0]
            ry += drift[k][1]
        if not O.pose_is_integrable(rx, ry, ryaw):
            continue
        cells = []
        for (x0, y0, x1, y1, hv) in O.expand_beams(rx, ry, ryaw, (df, dl, db, dr)):
            cells.append(g.world_to_grid(x0, y0) + g.world_to_grid(x1, y1) + (int(hv),))
        want[k] = cells
    return want


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
def test_expansion_survives_sloppy_sincos(hh, mode):
    """Golden session + adversarial streams (quantised yaws, poses on cell boundaries): the
    cells must equal the oracle's for an exact and for three 2-ulp-perturbed library sincos."""
    cases = []
    pk, _ = session_packets(time_sorted=True)
    cases.append((normalise_datagrams(pk)[0], -5.0, -5.0, 0.05, 0.0, None))
    for name, kw in (('mixed_a', (-5.0, -5.0, 0.05, 0.0)), ('mixed_b_sep', (-5.0, -5.0, 0.05, 0.75)),
                     ('mixed_c_4096', (-102.4, -102.4, 0.05, 0.0)), ('mixed_d_edge', (-2.4, -2.4, 0.05, 0.5))):
        pk, drift = load_packet_stream(name)
        arr, d = normalise_datagrams(pk, drift)
        cases.append((arr, *kw, d))
    slow_total = 0
    for arr, ox, oy, res, sep, drift in cases:
        got, st = expand(hh, arr, ox, oy, res, mode, separation=sep, drift=drift)
        want = oracle_cells(arr, ox, oy, res, separation=sep, drift=drift)
        for k in range(arr.shape[0]):
            if k not in want:
                assert st[k] != 0
                continue
            assert st[k] == 0
            for s in range(4):
                assert got[k, s, 5] == 1
                assert tuple(got[k, s, :5]) == want[k][s], (k, s, got[k, s], want[k][s])
        slow_total += int(got[:, :, 6].sum())
    assert slow_total > 0          # the exact path was exercised


def test_expansion_on_synthetic_swarm_and_boundary_poses(hh):
    from occgrid_b200 import simulation_tools as st
    s = st.generate_session(n_agents=64, n_packets=30_000, seed=12)
    offs = {a: tuple(s['agent_offsets'][a]) for a in range(1, 65)}
    want = oracle_cells(s['packets'], -102.4, -102.4, 0.05, agent_offsets=offs)
    for mode in (1, 3):
        got, stt = expand(hh, s['packets'], -102.4, -102.4, 0.05, mode, agent_offsets=s['agent_offsets'])
        for k, cells in want.items():
            for b in range(4):
                assert tuple(got[k, b, :5]) == cells[b]
    # poses exactly on cell corners with axis-aligned and diagonal yaws, ranges that land on corners
    pk = []
    for i in range(4000):
        x, y = (i % 41 - 20) * 0.05, (i // 41 - 20) * 0.25
        yaw = math.radians(45.0 * (i % 8))
        d = [0.05 * ((i + j) % 25) for j in range(4)]
        pk.append(struct.pack('<4sBfffiIffffB', b'QSRL', 1, x, y, yaw, 0, 0, *d, 0))
    arr = normalise_datagrams(pk)[0]
    want = oracle_cells(arr, -5.0, -5.0, 0.05)
    for mode in (0, 1, 2, 3):
        got, _ = expand(hh, arr, -5.0, -5.0, 0.05, mode)
        for k, cells in want.items():
            for b in range(4):
                assert tuple(got[k, b, :5]) == cells[b], (mode, k, b)
    assert got[:, :, 6].sum() > 100

