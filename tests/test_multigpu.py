"""Multi-GPU parity: owner-sharded count pass + distributed lookup pass under torchrun, 2 ranks."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _n_gpus():
    import torch
    return torch.cuda.device_count()


# transports of the round exchange: peer-mapped receive buffers written by the copy engines (default) or by k_push_copy,
# and ncclSend / ncclRecv of the same parts when the buffers cannot be mapped (KMN_P2P=0); with and without the
# shared-memory phase 2
# lookup pass: request / response rounds (default) or the owners' tables probed over peer memory (KMN_PEER_LOOKUP=1)
@pytest.mark.parametrize("world,env", [(2, {}), (2, {"KMN_P2P": "0"}), (2, {"KMN_PUSH": "kernel"}), (2, {"KMN_SMEM_COUNT": "0"}), (2, {"KMN_PEER_LOOKUP": "1"}),
                                       (4, {}), (4, {"KMN_P2P": "0"}), (8, {})])
def test_owner_sharded_count_and_lookup(world, env):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420, env=dict(os.environ, **env))
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-6000:]
    assert p.stdout.count("mgpu ok") == 4
