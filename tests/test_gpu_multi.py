"""GPU, >= 2 devices: the hash-sharded count + replica statistics over REAL NCCL and NVLink peer memory (CUDA IPC), one
process per GPU launched like the bench (torch.distributed.run), bit-exact against the CPU oracle on the concatenated
reads.  Skipped on a single-GPU box; `gpurun --gpus 2|8 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it
(the logs of those runs are committed under profiles/)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, exchange, fold, coarse, nreads=6000):
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mp_sharded_worker.py"), exchange, str(fold), str(coarse), str(nreads)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP_SHARDED_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-6000:])


def test_sharded_world1_collectives_and_routed_lookups():
    """one rank (runs on a one-GPU box too): the whole sharded code path -- log exchange through NCCL collectives, replica,
    and the ROUTED lookups (keys out, counts back, statistics from counts) -- against the oracle"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    _run(1, "collective", 0, 0, nreads=4000)


@pytest.mark.parametrize("exchange,fold,coarse", [("peer", 0, 0), ("peer", 1, 2), ("collective", 0, 0), ("collective", 1, 1)])
def test_sharded_world2(exchange, fold, coarse):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, exchange, fold, coarse)


@pytest.mark.parametrize("exchange,fold,coarse", [("peer", 1, 2), ("peer", 0, 0), ("collective", 1, 0)])
def test_sharded_world8(exchange, fold, coarse):
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, exchange, fold, coarse, nreads=12000)
