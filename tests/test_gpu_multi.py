"""Several GPUs of one box, one process per GPU (torchrun): the histogram exchange step over NVLink peer memory."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_hist_allreduce_over_peer_memory():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    import socket
    with socket.socket() as sk:          # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HERE, "dist_hist_reduce.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HIST_REDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
