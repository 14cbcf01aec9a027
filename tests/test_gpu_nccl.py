"""GPU, >= 2 devices: the data-parallel training step over NCCL equals the single-process step on the concatenated batch
(VERDICT r01 weak #5).  Launched like the driver launches bench.py: torch.distributed.run, one process per GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_rank_nccl_step_equals_single_process_step():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'nccl_two_rank_step.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and 'NCCL_STEP_OK' in r.stdout
