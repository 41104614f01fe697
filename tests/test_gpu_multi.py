"""The observation-sharded evaluation on 2 ranks (one process per GPU, NCCL) against the single-GPU evaluation of
the whole series: ELBO and gradient in both regimes, identical variables and random draws on ranks that seed numpy
differently, fpi / SMF bound / predict_f under sharding.  Self-launches tools/check_multi.py under torchrun when at
least two GPUs are visible (the driver's 1-GPU box skips it; `gpurun --gpus 2` runs it)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_agree_with_one_gpu():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tools', 'check_multi.py'), '6000', '64']
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:]
    assert 'OK widened rows' in p.stdout and 'OK\n' in p.stdout, p.stdout[-3000:]
