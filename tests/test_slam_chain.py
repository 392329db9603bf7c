"""The C library's drift-table producer (occgrid_slam_*, csrc/slam_chain.cpp) against the
reference-generated golden drift tables, the Python mirror of PoseGraphSLAM and the oracle's
replay (dual_bot_mapper.py:261-338, :826-857, :908-914).  Host code: runs without a GPU.
Bit-exact: drift tables and closure lists must be identical."""
import math
import os
import struct

import numpy as np
import pytest

from conftest import GOLD, load_packet_stream, session_packets

torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def M():
    from occgrid_b200 import dual_bot_mapper
    return dual_bot_mapper


@pytest.mark.parametrize('order', ['file', 'time'])
def test_golden_session_drift(M, golden, order):
    pk, _ = session_packets(time_sorted=(order == 'time'))
    drift, slam = M.slam_drift_table(pk, separation=0.0)
    assert isinstance(slam, M.NativePoseGraphSLAM)
    assert np.array_equal(drift, np.load(os.path.join(GOLD, f'session_drift_{order}.npy')))
    assert len(slam.closures) == golden['session'][f'{order}_order_slam_on']['closures']
    # the Python mirror of the reference class gives the same closures, value for value
    d2, py = M.slam_drift_table(pk, separation=0.0, slam=M.PoseGraphSLAM())
    assert np.array_equal(d2, drift)
    assert py.closures == slam.closures
    for agent in (1, 2):
        assert py.get_correction_for_agent(agent) == slam.get_correction_for_agent(agent)
    assert slam.n_nodes == len(py.nodes) and slam.n_landmarks == len(py.landmarks)


@pytest.mark.parametrize('name', ['mixed_a', 'mixed_b_sep', 'mixed_c_4096', 'mixed_d_edge'])
def test_adversarial_streams_drift(M, golden, name):
    want = golden['packet_streams'][name]
    if not want['slam']:
        pytest.skip('stream recorded without SLAM')
    pk, ref_drift = load_packet_stream(name)
    drift, slam = M.slam_drift_table(pk, separation=want['separation'])
    assert np.array_equal(drift, ref_drift)
    assert len(slam.closures) == want['closures']


def _walk_packets(n, seed, v1_every=0):
    """Two agents wandering in a small room with frequent landmarks: many revisits, so closures
    happen all the time; sprinkled with dropped datagrams."""
    r = np.random.default_rng(seed)
    pos = {1: np.array([0.0, 0.0]), 2: np.array([1.0, 0.5])}
    out = []
    for k in range(n):
        a = 1 + int(r.integers(0, 2))
        pos[a] = np.clip(pos[a] + r.normal(0, 0.08, 2), -2.0, 2.0)
        lm = int(r.integers(1, 6)) if r.random() < 0.35 else 0
        magic = b'QSRL'
        agent = a
        u = r.random()
        if u < 0.01:
            magic = b'QSRX'
        elif u < 0.02:
            agent = 3
        x, y = pos[a]
        if 0.02 <= u < 0.025:
            x = float('nan')
        d = struct.pack('<4sBfffiIffffB', magic, agent, x, y, float(r.uniform(-3, 3)), k, 0, 0.5, 0.6, 0.7, 0.8, lm)
        if v1_every and k % v1_every == 0:
            d = d[:41]
        if 0.025 <= u < 0.03:
            d = d[:17]
        out.append(d)
    return out


@pytest.mark.parametrize('seed,v1_every', [(1, 0), (2, 7), (3, 3)])
def test_random_walks_match_python_mirror_and_oracle(M, seed, v1_every):
    from oracle import occgrid_oracle as O
    pk = _walk_packets(6000, seed, v1_every)
    sep = 0.25 * seed
    drift, slam = M.slam_drift_table(pk, separation=sep)
    d2, py = M.slam_drift_table(pk, separation=sep, slam=M.PoseGraphSLAM())
    assert np.array_equal(drift, d2)
    assert slam.closures == py.closures and len(py.closures) > 20
    os_, dl = O.OracleSLAM(), []
    O.replay(pk, grid=O.OracleGrid(), separation=sep, slam=os_, drift_out=dl)
    assert np.array_equal(drift, np.asarray(dl, np.float64).reshape(-1, 2))
    assert [tuple(c) for c in os_.closures] == slam.closures


def test_state_carries_over_batches(M):
    pk = _walk_packets(4000, 11)
    whole, s1 = M.slam_drift_table(pk, separation=0.1)
    s2 = M.NativePoseGraphSLAM()
    parts = [M.slam_drift_table(pk[i:i + 700], separation=0.1, slam=s2)[0] for i in range(0, len(pk), 700)]
    assert np.array_equal(np.concatenate(parts), whole)
    assert s1.closures == s2.closures


def test_array_input_and_errors(M):
    pk = _walk_packets(500, 5)
    arr = np.zeros((len(pk), 42), np.uint8)
    sizes = np.array([len(d) for d in pk], np.int32)
    for k, d in enumerate(pk):
        arr[k, :len(d)] = np.frombuffer(d, np.uint8)
    a = M.NativePoseGraphSLAM().drift_table(arr, 0.0, sizes)
    b, _ = M.slam_drift_table(pk, 0.0)
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        M.NativePoseGraphSLAM().drift_table(arr.ravel(), 0.0)
    with pytest.raises(M.OccGridError):
        M.NativePoseGraphSLAM().drift_table(arr[:, :30], 0.0)
