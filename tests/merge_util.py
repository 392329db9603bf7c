import numpy as np


def synth_agent_grid(n, seed, occ_segments=20, free_frac=0.4):
    """Wall-like synthetic agent map (SURVEY §8d cfg 3): ~40 % free, thin occupied segments,
    rest unknown."""
    r = np.random.default_rng(seed)
    g = np.full((n, n), -1, np.int8)
    g[r.random((n, n)) < free_frac] = 0
    for _ in range(occ_segments):
        x0, y0 = (int(v) for v in r.integers(0, n, 2))
        L = int(r.integers(n // 25 + 2, n // 3 + 3))
        if r.random() < 0.5:
            g[y0, x0:x0 + L] = 100
        else:
            g[y0:y0 + L, x0] = 100
    g[r.random((n, n)) < 0.0005] = 60      # a few >50 non-100 values (threshold test, map_merger.py:72)
    g[r.random((n, n)) < 0.0005] = 50      # exactly 50 is NOT occupied
    return g


def load_merge_ref():
    """tests/golden/merge_ref.npz — produced by oracle/make_golden_merge.py, which EXECUTES the
    unmodified reference server_nodes/map_merger.py under rclpy / nav_msgs / open3d stubs.
    -> {sequence: [step dicts]} plus 'P' (publish_global_map on hand-made clouds)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'merge_ref.npz'))
    seqs = {}
    for key in z.files:
        parts = key.split('/')
        if len(parts) == 3:
            seqs.setdefault(parts[0], {}).setdefault(int(parts[1]), {})[parts[2]] = z[key]
        else:
            seqs.setdefault(parts[0] + '_meta', {})[parts[1]] = z[key]
    return {k: ([v[i] for i in sorted(v)] if not k.endswith('_meta') else v) for k, v in seqs.items()}
