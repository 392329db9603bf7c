"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Adds the frontier fixtures to tests/golden/golden.json by running the
UNMODIFIED reference `OccupancyGrid.get_frontiers / cluster_frontiers / cluster_centroid_world`
(server_nodes/dual_bot_mapper.py:181-237) in the authoring container:
  * on the golden 2-bot session grid (time order, SLAM off),
  * on a seeded random 96x96 grid with other geometry (saved as frontier_random_grid.npy).
Run after oracle/make_golden.py:   python oracle/make_golden_frontiers.py
"""
import hashlib
import json
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


def describe(fr, cl, ce):
    return {'n_frontiers': len(fr), 'frontiers_sha1': hashlib.sha1(np.asarray(fr, np.int32).tobytes()).hexdigest(),
            'n_clusters': len(cl), 'cluster_sizes': [len(c) for c in cl], 'cluster_first': [list(c[0]) for c in cl],
            'centroids': [list(c) for c in ce],
            'cluster_sets_sha1': hashlib.sha1(json.dumps([sorted(map(list, c)) for c in cl]).encode()).hexdigest()}


def main():
    m = ref_loader.load_dual_bot_mapper()
    pk, _ = session_packets(True)
    g, _ = O.replay(pk)
    occ = m.OccupancyGrid()
    occ.grid[:] = g.grid
    fr = occ.get_frontiers()
    cl = occ.cluster_frontiers(fr)
    ce = [occ.cluster_centroid_world(c) for c in cl]
    gold = json.load(open(os.path.join(GOLD, 'golden.json')))
    gold['frontiers_session_time_off'] = describe(fr, cl, ce)
    rng = np.random.default_rng(31)
    rg = rng.choice(np.array([-1, 0, 100], np.int8), size=(96, 96), p=[0.45, 0.45, 0.10])
    occ2 = m.OccupancyGrid(size=96, resolution=0.1, origin_x=-3.3, origin_y=7.7)
    occ2.grid[:] = rg
    fr2 = occ2.get_frontiers()
    cl2 = occ2.cluster_frontiers(fr2)
    ce2 = [occ2.cluster_centroid_world(c) for c in cl2]
    np.save(os.path.join(GOLD, 'frontier_random_grid.npy'), rg)
    gold['frontiers_random96'] = describe(fr2, cl2, ce2)
    json.dump(gold, open(os.path.join(GOLD, 'golden.json'), 'w'), indent=1, sort_keys=True)
    print(len(fr), len(cl), len(fr2), len(cl2))


if __name__ == '__main__':
    main()
