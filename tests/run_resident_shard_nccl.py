"""ResidentLadderShard over NCCL on two GPUs against a single-process run (launched by tests/test_gpu_nccl.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from mcpy_b200 import Engine
from mcpy_b200.ensembles import ResidentLadderShard
from mcpy_b200.utils import synthetic
from mcpy_b200.utils.chunking import shard_range


def ladder(n_rep):
    cfgs = []
    for g in range(n_rep):
        c = synthetic.s2_lj(400 + 13 * (g % 3), 2.5, 0.5, seed=40 + g)          # ragged sizes, boxes that differ
        cfgs.append((c['positions'], c['numbers'], c['cell'], c['pbc']))
    return cfgs, np.full(n_rep, 0.5), np.linspace(-3.0, -2.0, n_rep)


def run(rank, world, local, n_rep, rounds):
    cfgs, betas, mus = ladder(n_rep)
    lo, hi = shard_range(n_rep, rank, world)
    eng = Engine(local)
    eng.set_lj(1.0, 1.0, 2.5)
    sh = ResidentLadderShard(eng, cfgs[lo:hi], betas=betas, mus=mus, seed=31, rank=rank, world_size=world,
                             device=f'cuda:{local}')
    for _ in range(rounds):
        sh.attempt_exchanges()
    # score one deletion per local replica against whatever configuration now sits in each slot
    R = hi - lo
    dE = eng.trial_delta_e(np.arange(R + 1, dtype=np.int32), np.full(R, 7, np.int32), np.zeros(R + 1, np.int32),
                           np.zeros((0, 3)), rep=np.arange(R, dtype=np.int32))
    state = {lo + k: (p, z, float(sh.E[k]), sh.cells[k].copy(), float(dE[k])) for k, (p, z) in enumerate(sh.positions_host())}
    return sh.swap_log, state


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_rep, rounds = 7, 9
    log, state = run(rank, world, local, n_rep, rounds)
    gathered = [None] * world
    dist.all_gather_object(gathered, (log, state))
    ok = True
    if rank == 0:
        log1, state1 = run(0, 1, local, n_rep, rounds)
        merged = {}
        for lg, st in gathered:
            ok = ok and lg == log1
            merged.update(st)
        ok = ok and sorted(merged) == sorted(state1) and any(a for *_, a in log1) and any(a and i == 3 for _, i, _, _, a in log1)
        for i in state1:
            a, b = state1[i], merged[i]
            ok = ok and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] \
                and np.array_equal(a[3], b[3]) and a[4] == b[4]
        print(f'resident ladder over NCCL: world={world} replicas={n_rep} attempts={len(log1)} '
              f'accepted={sum(1 for *_, a in log1 if a)} identical_to_single_process={ok}')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
