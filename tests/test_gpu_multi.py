"""Multi-GPU parity on hardware (needs >= 2 GPUs; skipped on a single-GPU box): the NCCL data-parallel step against the
concatenated batch on one rank, and the sharded optimiser against all-reduce + Adam.  Workers: tests/multi_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs on the box")]
HERE = os.path.dirname(os.path.abspath(__file__))


def run(mode, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multi_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_two_nccl_ranks_match_the_concatenated_batch():
    assert "MULTI grads" in run("grads")


def test_partial_table_exchange_matches_the_concatenated_batch():
    assert "MULTI partial" in run("partial")


def test_peer_memory_exchange_matches_the_concatenated_batch():
    assert "MULTI peer" in run("peer")


def test_peer_memory_exchange_with_p2p_store_kernels():
    assert "MULTI peerk" in run("peerk")


def test_sharded_adam_matches_allreduce_plus_adam():
    assert "MULTI sharded" in run("sharded")
