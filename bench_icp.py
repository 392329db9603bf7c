#!/usr/bin/env python
"""bench_icp.py — registration of one agent grid against the fused global cloud (SURVEY §8 rows
a13 / f3): `registration_icp(local, global, 1.0, I, point-to-point, max_iteration=30)`,
server_nodes/map_merger.py:45-52.

Workload: global cloud = fusion of `--agents` grids of `--size`^2 (BASELINE configs[2] maps and
transforms); local = one more view of agent 0's map displaced by (0.2 m, -0.15 m, 2 deg).

    python bench_icp.py [--agents 32] [--size 2048] [--repeat 5]

Prints ONE JSON line: registrations/s on one B200 (device time of the whole call: cell-list build
over the global cloud + up to 31 association passes), and the CPU baseline (SciPy cKDTree
restatement, oracle/icp_oracle.py, one core).
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=32)
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--repeat', type=int, default=5)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    import torch
    from bench_merge import device_grids
    from occgrid_b200 import _native
    from occgrid_b200.map_merger import MapMerger, make_grid_msg, se2_matrix
    assert torch.cuda.is_available(), 'bench_icp.py needs a CUDA device (no CPU fallback)'
    dev = torch.device('cuda', 0)
    A, S, res = args.agents, args.size, 0.05
    grids = device_grids(torch, dev, A, S)
    rng = np.random.default_rng(0)
    origins = np.tile(np.array([[-S * res / 2, -S * res / 2]]), (A, 1))
    tf = [se2_matrix(*rng.uniform(-50, 50, 2), rng.uniform(-math.pi, math.pi)) for _ in range(A)]
    m = MapMerger(device=dev, registration='icp')
    m.merge(grids, origins, res, tf, to_host=False)
    # agent 0 was adopted untransformed: a displaced second view of its map is the local cloud
    D = se2_matrix(0.2, -0.15, math.radians(2.0))
    Dinv = np.linalg.inv(D)
    ox, oy = origins[0]
    # moving the grid's frame: cell (c, r) sits at D^-1 * (ox + c res, oy + r res); a pure frame shift
    # keeps it a grid only for translations, so displace by the translation part and let ICP find the rest
    msg = make_grid_msg(grids[0], S, S, res, ox + Dinv[0, 3], oy + Dinv[1, 3])
    reg = m.register(msg)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.repeat):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reg = m.register(msg)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    _native.profile_begin()
    m.register(msg)
    torch.cuda.synchronize()
    prof = _native.profile_end()
    n_local = int((grids[0] > 50).sum().item())
    result = {
        'metric': 'icp_registrations_per_sec', 'value': 1e3 / ms, 'unit': 'registrations/s', 'n_gpus': 1, 'ms_per_registration': ms,
        'higher_is_better': True, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'registration_icp of one {S}^2 agent grid ({n_local} points) against the fused cloud of {A} grids '
                               f'({m._n_global} points), threshold 1.0 m, max_iteration 30',
                   'iterations_run': reg.iterations, 'fitness': reg.fitness, 'inlier_rmse': reg.inlier_rmse},
        'kernels': {k: {'ms_total': v[0], 'kernels': v[1]} for k, v in prof.items()},
    }
    if not args.no_cpu:
        from oracle import icp_oracle as IO
        from oracle import merge_oracle as MO
        pc = m.global_pcd
        lx, ly = MO.grid_to_points(grids[0].cpu().numpy().ravel(), S, S, res, ox + Dinv[0, 3], oy + Dinv[1, 3])
        t0 = time.perf_counter()
        Tw, fw, rw, iw = IO.registration_icp(lx, ly, pc[:, 0], pc[:, 1])
        dt = time.perf_counter() - t0
        result['cpu_baseline'] = {'value': 1.0 / dt, 'unit': 'registrations/s', 'cores': 1, 'kind': 'port',
                                  'sample': 'the same registration, SciPy cKDTree + NumPy restatement (oracle/icp_oracle.py)',
                                  'fitness': fw, 'iterations_run': iw,
                                  'max_abs_T_diff': float(np.abs(Tw - reg.transformation).max())}
    print(json.dumps(result))


if __name__ == '__main__':
    main()
