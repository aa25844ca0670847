"""GPU, two devices: the sharded replica-exchange driver and the device-resident ladder over NCCL.

Skipped on a single-GPU box.  Each test launches ``torchrun --nproc-per-node 2`` on a helper script; the helper
compares the two-rank run with a single-process run of the same ladder and exits non-zero on any difference."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_gpus():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


def _torchrun(script, *args):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.skipif(not _two_gpus(), reason='needs two GPUs')
def test_sharded_replica_exchange_over_nccl_equals_single_process():
    """ShardedReplicaExchange over unmodified GrandCanonicalEnsemble replicas (FixAtoms on two atoms), one
    B200Calculator per GPU, records all_gathered and cross-rank states shipped over NCCL: same swap decisions,
    particle numbers and positions as the single-process run, energies to 1e-10."""
    proc = _torchrun('run_sharded_re_nccl.py')
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert 'identical_to_single_process=True' in proc.stdout


@pytest.mark.skipif(not _two_gpus(), reason='needs two GPUs')
def test_resident_replica_exchange_over_nccl_equals_single_process_host_driver():
    """ResidentReplicaExchange on two ranks (every replica's trials on its GPU, a CTA each; local swaps on the device,
    the boundary swap shipped over NCCL) against the host-driven single-process run over B200Calculator: same swap
    decisions, particle numbers and positions, energies to 1e-9."""
    proc = _torchrun('run_sharded_re_nccl.py', 'resident')
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert 'identical_to_single_process=True' in proc.stdout


@pytest.mark.skipif(not _two_gpus(), reason='needs two GPUs')
def test_resident_replica_exchange_swaps_device_to_device_across_ranks():
    """Same run with nothing attached to the Atoms that the configuration block does not hold: every accepted boundary
    swap sends the block GPU to GPU over NCCL (``ResidentGCMC.remote_swap``), no host round trip; same final state."""
    proc = _torchrun('run_sharded_re_nccl.py', 'resident', 'device-swaps')
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert 'identical_to_single_process=True' in proc.stdout


@pytest.mark.skipif(not _two_gpus(), reason='needs two GPUs')
def test_resident_ladder_shard_over_nccl_equals_single_process():
    """ResidentLadderShard: configurations resident in HBM, ONE all_gather per exchange attempt carrying records and
    boundary configurations; after the exchange rounds every slot holds the configuration, box and energy the
    single-process run puts there, and trial dE scored against the swapped configurations agree bit for bit."""
    proc = _torchrun('run_resident_shard_nccl.py')
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert 'identical_to_single_process=True' in proc.stdout
