"""REAL multi-GPU parity (skipped unless >= 2 CUDA devices are visible): two processes, NCCL
process group, TiledSwarmMap with the fused raycast + route kernel over symmetric (peer-mapped)
memory and the cross-GPU barrier of occgrid_band_publish; the all-gathered map must equal the C
oracle over the canonical stream bit for bit.  Also the NCCL all-to-all variant and the sharded
map merge."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.mark.timeout(600)
def test_two_rank_band_exchange_on_real_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 CUDA devices')
    env = dict(os.environ, CHECK_GRID_PER_GPU='1024', CHECK_AGENTS_PER_GPU='16')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', str(29600 + os.getpid() % 300),
                        os.path.join(ROOT, 'tools', 'check_multi_gpu.py')], capture_output=True, text=True, env=env, timeout=540)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert 'exchange=p2p: map bit-exact' in out and 'exchange=nccl: map bit-exact' in out, out[-4000:]
    assert 'PASS' in out and 'DIFFERS' not in out, out[-4000:]
