"""CPU checks of the merge oracle (restatement of server_nodes/map_merger.py:35-127).

Pinned by execution for the reference's own code: tests/golden/merge_ref.npz was produced by
running the UNMODIFIED map_merger.py (grid_to_pcd :64-85, map_callback :35-62,
publish_global_map :87-127) under rclpy / nav_msgs / open3d stubs (oracle/make_golden_merge.py).
The arithmetic inside Open3D's transform / += / voxel_down_sample / registration_icp stays a
restatement of the published algorithm (the library is absent and un-pinned)."""
import math

import numpy as np

from merge_util import load_merge_ref, synth_agent_grid
from oracle import merge_oracle as MO


def test_grid_to_points_is_corner_and_row_major():
    g = np.full((3, 4), -1, np.int8)
    g[0, 1] = 100
    g[2, 3] = 100
    g[1, 0] = 51
    g[1, 1] = 50
    x, y = MO.grid_to_points(g.ravel(), 4, 3, 0.05, -1.0, 2.0)
    assert x.tolist() == [1 * 0.05 + -1.0, 0 * 0.05 + -1.0, 3 * 0.05 + -1.0]
    assert y.tolist() == [0 * 0.05 + 2.0, 1 * 0.05 + 2.0, 2 * 0.05 + 2.0]


def test_first_callback_adopts_without_downsample_and_identity_property():
    """SURVEY §4 merge KAT: one grid -> output occupied set == input occupied set shifted to
    its bbox, up to the reference's own int((p-min)/res) rounding collisions."""
    g = synth_agent_grid(128, 1)
    m = MO.OracleMerger()
    out, origin = m.map_callback(g.ravel(), 128, 128, 0.05, -3.2, -3.2)
    ys, xs = np.nonzero(g > 50)
    assert origin == (xs.min() * 0.05 - 3.2, ys.min() * 0.05 - 3.2)
    assert m.map_resolution == 0.05 and m.map_origin == [-3.2, -3.2]
    assert set(np.unique(out)) <= {-1, 100}
    assert 0.9 * len(xs) <= (out == 100).sum() <= len(xs)
    # with res = 0.5 (exact in binary) there are no rounding collisions
    m2 = MO.OracleMerger()
    out2, origin2 = m2.map_callback(g.ravel(), 128, 128, 0.5, -32.0, -32.0)
    want = np.full_like(out2, -1)
    want[ys - ys.min(), xs - xs.min()] = 100
    assert np.array_equal(out2, want)


def test_empty_and_rejected_callbacks_return_none():
    m = MO.OracleMerger()
    assert m.map_callback(np.full(16, -1, np.int8), 4, 4, 0.05, 0, 0) is None
    g = synth_agent_grid(64, 2)
    assert m.map_callback(g.ravel(), 64, 64, 0.05, 0, 0) is not None
    n = m.gx.size
    assert m.map_callback(g.ravel(), 64, 64, 0.05, 0, 0, MO.se2_matrix(1, 1, 0.3), accept=False) is None
    assert m.gx.size == n


def test_voxel_down_sample_means_and_order():
    px = np.array([0.0, 0.01, 0.2, 0.21, 0.22, 1.0])
    py = np.array([0.0, 0.01, 0.0, 0.0, 0.01, 1.0])
    x, y = MO.voxel_down_sample(px, py, 0.05)
    # voxel origin = min - 0.025: points 0,1 share a voxel; 2,3,4 share one; 5 alone
    assert x.shape == (3,)
    assert x[0] == (0.0 + 0.01) / 2 and y[0] == (0.0 + 0.01) / 2
    assert x[1] == ((0.2 + 0.21) + 0.22) / 3
    assert (x[2], y[2]) == (1.0, 1.0)
    # canonical order: first appearance (voxels ordered by their smallest point index)
    px = np.array([1.0, 0.0, 0.5, 1.01]); py = np.array([0.0, 1.0, 0.0, 0.01])
    x, y = MO.voxel_down_sample(px, py, 0.05)
    assert x.tolist() == [(1.0 + 1.01) / 2, 0.0, 0.5] and y.tolist() == [(0.0 + 0.01) / 2, 1.0, 0.0]


def test_sequence_is_deterministic_and_only_unknown_or_occupied():
    rng = np.random.default_rng(3)
    outs = []
    for rep in range(2):
        r = np.random.default_rng(3)
        m = MO.OracleMerger()
        for a in range(5):
            g = synth_agent_grid(200, 10 + a)
            T = MO.se2_matrix(*r.uniform(-4, 4, 2), r.uniform(-math.pi, math.pi))
            out = m.map_callback(g.ravel(), 200, 200, 0.05, -5.0, -5.0, T)
        outs.append(out)
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert set(np.unique(outs[0][0])) == {-1, 100}


def test_host_transform_normalisation_matches_oracle_matrices():
    """MapMerger.merge accepts [A,4,4], [A,3] (tx, ty, theta), None or a mixed list; the matrices it
    hands to the device must be the oracle's se2_matrix bit for bit (cos/sin from the same libm)."""
    pytest = __import__('pytest')
    pytest.importorskip('torch')
    from occgrid_b200.map_merger import MapMerger, se2_matrix
    from oracle import merge_oracle as MO
    import numpy as np
    r = np.random.default_rng(3)
    p = r.uniform(-5, 5, (6, 3))
    want = np.stack([MO.se2_matrix(*row) for row in p])
    assert np.array_equal(MapMerger._as_matrices(p, 6), want)
    assert np.array_equal(MapMerger._as_matrices(want, 6), want)
    assert np.array_equal(MapMerger._as_matrices(None, 3), np.tile(np.eye(4), (3, 1, 1)))
    mixed = [tuple(p[0]), want[1], None, p[3], want[4].ravel(), tuple(p[5])]
    got = MapMerger._as_matrices(mixed, 6)
    want[2] = np.eye(4)
    assert np.array_equal(got, want) and got.flags['C_CONTIGUOUS'] and got.dtype == np.float64
    assert np.array_equal(se2_matrix(*p[2]), MO.se2_matrix(*p[2]))


# ---- fixtures produced by executing the unmodified reference -------------------------------------
def _replay_restatement(steps, icp=False):
    from oracle import icp_oracle as IO
    m = MO.OracleMerger()
    for k, s in enumerate(steps):
        g = s['grid']
        h, w = g.shape
        res, ox, oy = s['geom'].tolist()
        px, py = MO.grid_to_points(g.ravel(), w, h, res, ox, oy)            # a10
        assert np.array_equal(px, s['pcd'][:, 0]) and np.array_equal(py, s['pcd'][:, 1]), k
        assert not s['pcd'][:, 2].any()
        T, fit = s['T'], float(s['fitness'])
        if icp and m.gx.size and px.size:
            T, fit, _, _ = IO.registration_icp(px, py, m.gx, m.gy, 1.0, 30)
        out = m.map_callback(g.ravel(), w, h, res, ox, oy, T, accept=fit >= 0.6)
        assert (out is not None) == bool(s['published']), k
        assert np.array_equal(m.gx, s['cloud'][:, 0]) and np.array_equal(m.gy, s['cloud'][:, 1]), k
        assert [m.map_resolution] + list(m.map_origin) == s['state'].tolist(), k
        if out is not None:
            assert out[0].dtype == np.int8 and np.array_equal(out[0], s['out']), k
            assert list(out[1]) == s['out_origin'].tolist() and m.map_resolution == float(s['out_res']), k
    return m


def test_restatement_equals_reference_executed_sequences():
    """a10 / a11 sequencing / a12: every callback of sequences A (reject + empty), B (other
    geometry, rectangular, empty first, fitness exactly 0.6) and C (heavy overlap) gives the
    point list, the cloud, the node state and the published grid the reference produced."""
    ref = load_merge_ref()
    assert _replay_restatement(ref['A']).gx.size == ref['A'][-1]['cloud'].shape[0] > 1000
    _replay_restatement(ref['B'])
    _replay_restatement(ref['C'])
    published = [bool(s['published']) for s in ref['A']]
    assert published == [True, True, False, True, False, True, True]         # :54-56 and :37-38
    assert ref['A_meta']['icp_calls'][:, 3].tolist() == [1.0, 0.59, 1.0, 1.0, 1.0]
    assert not bool(ref['B'][0]['published']) and ref['B_meta']['icp_calls'][0, 3] == 0.6


def test_restatement_equals_reference_driven_registration():
    """Sequence D: the reference's own map_callback called registration_icp (restated) on the
    clouds IT built; replaying with the restated ICP must reproduce every cloud and grid."""
    ref = load_merge_ref()
    _replay_restatement(ref['D'], icp=True)
    calls = ref['D_meta']['icp_calls']
    assert calls.shape == (3, 4) and (calls[:, 2] == 1.0).all() and (calls[:, 3] >= 0.6).all()


def test_rasterise_equals_reference_publish_on_handmade_clouds():
    ref = load_merge_ref()
    for case in ref['P']:
        g, origin = MO.rasterise(case['points'][:, 0], case['points'][:, 1], float(case['res']))
        assert np.array_equal(g, case['out']) and list(origin) == case['out_origin'].tolist()
    assert ref['P'][1]['out'].shape == (1, 1)
