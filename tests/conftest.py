import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs /root/reference (authoring container only)')


@pytest.fixture(scope='session')
def golden():
    with open(os.path.join(GOLD, 'golden.json')) as f:
        return json.load(f)


def load_packet_stream(name):
    """-> (list of datagrams, drift table [n,2]) from tests/golden/packets_<name>.npz"""
    z = np.load(os.path.join(GOLD, f'packets_{name}.npz'))
    blob = z['blob'].tobytes()
    lens = z['lens']
    offs = np.concatenate([[0], np.cumsum(lens)])
    pk = [blob[offs[i]:offs[i + 1]] for i in range(len(lens))]
    return pk, z['drift']


def normalise_datagrams(datagrams, drift=None):
    """Host-side batcher rule (dual_bot_mapper.py:828-838): 42-byte datagrams are v2, 41-byte
    ones are v1 (landmark := 0 -> pad with one zero byte), every other size is dropped.
    Returns uint8 [m,42] and the matching rows of the drift table."""
    keep = [i for i, p in enumerate(datagrams) if len(p) in (41, 42)]
    arr = np.zeros((len(keep), 42), np.uint8)
    for j, i in enumerate(keep):
        p = datagrams[i]
        arr[j, :len(p)] = np.frombuffer(p, np.uint8)
    d = None if drift is None else np.ascontiguousarray(np.asarray(drift)[keep], np.float64)
    return arr, d


def session_packets(time_sorted=True):
    from oracle.occgrid_oracle import load_session_rows, rows_to_packets
    rows = load_session_rows(os.path.join(GOLD, 'fake_dual_session', 'telemetry.csv'),
                             time_sorted=time_sorted)
    return rows_to_packets(rows), rows
