"""The sharded replica-exchange driver on real GPUs over NCCL (launched by tests/test_gpu_nccl.py, or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tests/run_sharded_re_nccl.py

Every rank drives its block of replicas with a B200Calculator on its own GPU; the record all_gather and
the state transport of accepted cross-rank swaps run over NCCL (device tensors).  Rank 0 also repeats the
whole ladder in one process (world_size = 1 semantics, its own GPU) and checks that the swap decisions,
particle numbers and positions of every replica are identical and the energies (and the swap exponents
computed from them) agree to 1e-10: a replica's energy is E(base) + dE(base -> now), and which
configuration is the resident base differs between the sharded and the single-process run (a
configuration that arrives from another rank re-bases, one that arrives from another local slot does
not), so the last bits may differ while every decision stays the same."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from baseline import refpath
refpath.enable()

import torch
import torch.distributed as dist

from helpers import make_factory
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.ensembles import ResidentReplicaExchange, ShardedReplicaExchange


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_rep, steps = 8, 24
    temps = list(np.linspace(1.5, 3.0, n_rep))
    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0, device=local)
    resident = 'resident' in sys.argv[1:]
    origin = 'device-swaps' not in sys.argv[1:]      # Atoms.info travels with the whole Atoms only: forces the host path
    if resident:
        # the replicas' trials on the device (a CTA each), swaps of local replicas there too; the boundary replica of an
        # accepted cross-rank swap comes back to the host and travels as in the sharded driver (no FixAtoms: out of the
        # resident chains' scope)
        steps = 40
        re = ResidentReplicaExchange(make_factory(calc, n_moves=3, origin=origin), calc, temperatures=temps, gcmc_steps=steps,
                                     exchange_interval=3, seed=31, rank=rank, world_size=world, device=f'cuda:{local}')
    else:
        re = ShardedReplicaExchange(make_factory(calc, fix=(0, 3)), calc, temperatures=temps, gcmc_steps=steps,
                                    exchange_interval=3, seed=31, rank=rank, world_size=world, device=f'cuda:{local}')
    re.run()
    mine = {re.lo + k: (r.E_old, r.n_atoms, r.atoms.positions.copy()) for k, r in enumerate(re.replicas)}
    gathered = [None] * world
    dist.all_gather_object(gathered, (re.swap_log, mine))
    dev_swaps = [None] * world
    dist.all_gather_object(dev_swaps, getattr(re, 'device_swaps', 0))
    ok = True
    if rank == 0:
        calc1 = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0, device=local)
        ref = ShardedReplicaExchange(make_factory(calc1, n_moves=3, origin=origin) if resident else make_factory(calc1, fix=(0, 3)), calc1,
                                     temperatures=temps, gcmc_steps=steps, exchange_interval=3, seed=31, rank=0, world_size=1)
        ref.run()
        merged = {}
        for log, st in gathered:
            ok = ok and len(log) == len(ref.swap_log)
            for (s1, i1, j1, d1, a1), (s2, i2, j2, d2, a2) in zip(log, ref.swap_log):
                ok = ok and (s1, i1, j1, a1) == (s2, i2, j2, a2) and \
                    (d1 == d2 or abs(d1 - d2) <= 1e-9 * max(1.0, abs(d2)))
            merged.update(st)
        accepted = sum(1 for *_, a in ref.swap_log if a)
        for i, r in enumerate(ref.replicas):
            e, n, p = merged[i]
            ok = ok and n == r.n_atoms and abs(e - r.E_old) <= (1e-9 if resident else 1e-10) * max(1.0, abs(r.E_old)) \
                and np.array_equal(p, r.atoms.positions)
        cross = sum(1 for _, i, j, _, a in ref.swap_log if a and i == n_rep // world - 1)
        if resident:
            ok = ok and cross > 0 and re.blocks < steps
            # every accepted boundary swap went device to device on both ranks, or none did (Atoms.info present)
            ok = ok and dev_swaps == ([0] * world if origin else [cross] * world)
        print(f'{"resident" if resident else "sharded"} RE over NCCL: world={world} cross_rank_swaps={cross} device_to_device={dev_swaps}  replicas={n_rep} swaps attempted={len(ref.swap_log)} '
              f'accepted={accepted} identical_to_single_process={ok}')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
