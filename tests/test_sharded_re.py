"""CPU: the sharded replica-exchange driver -- swap rule vs the reference's analytic cases,
world_size-2 gloo run == single-process run (the only collective is the record all_gather)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swap_exponent_reduces_to_the_textbook_forms():
    from mcpy_b200.ensembles import swap_exponent
    # temperature ladder, shared mu: delta = (beta_j - beta_i)(Phi_j - Phi_i)
    #   (tests/test_batched_re_acceptance.py of the reference checks the same identity)
    mu, S = 0.3, 1
    ri = np.array([-10.0, 1 / 1.5, 12, mu, 1.0])
    rj = np.array([-14.0, 1 / 2.5, 17, mu, 1.0])
    phi_i, phi_j = -10.0 - mu * 12, -14.0 - mu * 17
    assert swap_exponent(ri, rj, S) == pytest.approx((rj[1] - ri[1]) * (phi_j - phi_i), rel=1e-13)
    # mu ladder, shared beta: delta = beta (mu_i - mu_j)(N_Y - N_X); energies cancel
    b = 0.5
    ri = np.array([-3.0, b, 10, -2.0, 1.0])
    rj = np.array([-9.0, b, 15, -1.5, 1.0])
    assert swap_exponent(ri, rj, S) == pytest.approx(b * (-2.0 - -1.5) * (15 - 10), rel=1e-13)
    # unequal temperatures and particle numbers pick up 3 dN ln(Lambda_i / Lambda_j)
    ri = np.array([-3.0, 0.6, 10, 0.0, 1.7])
    rj = np.array([-3.0, 0.4, 13, 0.0, 1.2])
    assert swap_exponent(ri, rj, S) == pytest.approx(3.0 * (10 - 13) * np.log(1.7 / 1.2), rel=1e-13)


def _run(rank, world, port, n_rep, steps, out, kind='T'):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from baseline import refpath
    refpath.enable()
    import torch.distributed as dist
    from helpers import OracleLJCalculator, make_factory
    from mcpy_b200.ensembles import ShardedReplicaExchange
    if world > 1:
        dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    calc = OracleLJCalculator()
    if kind == 'T':
        ladder = dict(temperatures=list(np.linspace(1.5, 3.0, n_rep)))
        factory = make_factory(calc, fix=(0, 3))
    else:
        # a mu ladder whose replicas do not all carry the same chemical-potential keys
        ladder = dict(mus=[{'H': -3.0 + 0.3 * k} for k in range(n_rep)])
        factory = make_factory(calc, fix=(1,), extra_mu={1: {'H': -2.7, 'He': -1.0}})
    re = ShardedReplicaExchange(factory, calc, gcmc_steps=steps, exchange_interval=3, seed=31,
                                rank=rank, world_size=world, **ladder)
    re.run()
    state = {re.lo + k: (r.E_old, r.n_atoms, r.atoms.positions.copy(), r.cells[0].get_volume(),
                         [list(map(int, c.get_indices())) for c in r.atoms.constraints],
                         dict(r.atoms.info), r._step)
             for k, r in enumerate(re.replicas)}
    out.put((rank, re.swap_log, state, (re.global_attempts, re.global_successes)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _launch(world, n_rep=5, steps=16, kind='T'):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, n_rep, steps, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(results, key=lambda t: t[0])


@pytest.mark.parametrize('kind', ['T', 'mu'])
def test_two_rank_gloo_run_equals_single_process_run(kind):
    """World size 2 over gloo == one process: same swap decisions, and after cross-rank swaps the same
    energies, positions, free volumes, FixAtoms constraints and ``atoms.info`` in every slot."""
    single = _launch(1, kind=kind)
    double = _launch(2, kind=kind)
    log1 = single[0][1]
    assert len(log1) > 0 and any(a for *_, a in log1), 'no swap was attempted / accepted'
    assert any(a and i == 2 for _, i, _, _, a in log1), 'no cross-rank swap (slots 2 <-> 3) was accepted'
    # every rank made the same decisions as the single-process run
    for _, log, _, tallies in double:
        assert log == log1
        assert tallies == single[0][3]
    state1 = single[0][2]
    state2 = {}
    for _, _, st, _ in double:
        state2.update(st)
    assert sorted(state2) == sorted(state1)
    origins = set()
    for i in state1:
        assert state2[i][0] == state1[i][0] and state2[i][1] == state1[i][1]
        assert np.array_equal(state2[i][2], state1[i][2])
        assert state2[i][3:] == state1[i][3:]          # volume, constraints, info, step counter
        origins.add(state1[i][5]['origin'] == i)
    assert False in origins, 'no configuration changed slots'


def test_single_rank_run_equals_reference_batched_replica_exchange():
    """Same factory, same seeds: the reference's BatchedReplicaExchange and this driver end in
    identical states after identical swap decisions (reference rows a5-a7 of SURVEY.md 8a)."""
    from helpers import OracleLJCalculator, make_factory
    from mcpy.ensembles import BatchedReplicaExchange
    from mcpy_b200.ensembles import ShardedReplicaExchange

    calc = OracleLJCalculator()
    factory = make_factory(calc, n0=20, base_seed=200, n_moves=2)
    for ladder in (dict(temperatures=[1.6, 2.0, 2.5, 3.1]),
                   dict(mus=[{'H': -3.0}, {'H': -2.6}, {'H': -2.2}])):
        ref = BatchedReplicaExchange(factory, calc, gcmc_steps=14, exchange_interval=3, seed=31,
                                     outfile=None, global_minimum_file=None, **ladder)
        ref.run()
        mine = ShardedReplicaExchange(factory, calc, gcmc_steps=14, exchange_interval=3, seed=31,
                                      rank=0, world_size=1, **ladder)
        mine.run()
        assert mine.exchange_attempts == ref.exchange_attempts
        assert mine.exchange_successes == ref.exchange_successes
        assert sum(mine.exchange_successes) > 0
        for a, b in zip(ref.replicas, mine.replicas):
            assert a.E_old == b.E_old and a.n_atoms == b.n_atoms
            assert np.array_equal(a.atoms.positions, b.atoms.positions)
            assert a.cells[0].get_volume() == b.cells[0].get_volume()


# ------------------------------------------------------------------ the device-resident form of the ladder
class _HostEngine:
    """Stands in for ``Engine`` on the CPU: ``upload_dev`` records what it is handed, energies from the oracle."""
    device = 0

    def __init__(self):
        self.uploads = 0

    def upload_dev(self, off, d_pos, d_num, cells, pbcs):
        self.off, self.cells, self.pbcs = np.array(off), np.array(cells), np.array(pbcs)
        self.uploads += 1

    def copy_segments_dev(self, segments):
        """Engine.copy_segments_dev on host memory: the same moves, through the same addresses."""
        import ctypes
        ctype = {0: (ctypes.c_double, ctypes.c_double), 1: (ctypes.c_int32, ctypes.c_int32),
                 2: (ctypes.c_double, ctypes.c_int32), 3: (ctypes.c_int32, ctypes.c_double)}
        for src, dst, n, kind in segments:
            if n:
                ts, td = ctype[kind]
                a = np.ctypeslib.as_array((ts * n).from_address(src))
                np.ctypeslib.as_array((td * n).from_address(dst))[:] = a

    def bind(self, shard):
        self.shard = shard

    def energies(self):
        from oracle import cfast
        sh, off = self.shard, self.off
        pos = sh.d_pos.numpy()
        return np.array([cfast.lj_energy(pos[off[k]:off[k + 1]], self.cells[k].reshape(3, 3), self.pbcs[k], 1.0, 1.0, 2.5)
                         for k in range(len(off) - 1)])


def _resident_configs(n_rep):
    out = []
    for g in range(n_rep):
        rng = np.random.default_rng(700 + g)
        L = 6.0 + 0.1 * g                                   # boxes differ: they must travel with the configuration
        n = 20 + (g % 3)                                    # ragged sizes
        out.append((rng.uniform(0.3, L - 0.3, size=(n, 3)), np.full(n, 1 + g % 2, np.int32), np.eye(3) * L, np.ones(3, bool)))
    return out


def _run_resident(rank, world, port, n_rep, rounds, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mcpy_b200.ensembles.resident_shard import ResidentLadderShard
    if world > 1:
        dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    from mcpy_b200.utils.chunking import shard_range
    lo, hi = shard_range(n_rep, rank, world)
    cfgs = _resident_configs(n_rep)
    eng = _HostEngine()

    class _Shard(ResidentLadderShard):
        def upload(self, energies=False):
            eng.bind(self)
            super().upload(energies)
    sh = _Shard(eng, cfgs[lo:hi], betas=np.full(n_rep, 0.5), mus=np.linspace(-3.0, -1.0, n_rep), seed=31,
                rank=rank, world_size=world, device='cpu')
    e0 = sh.E.copy()
    for _ in range(rounds):
        sh.attempt_exchanges()
    state = {lo + k: (p, z, float(sh.E[k]), sh.cells[k].copy()) for k, (p, z) in enumerate(sh.positions_host())}
    out.put((rank, sh.swap_log, state, e0, eng.uploads))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_resident_ladder_shard_two_ranks_equal_one():
    """ResidentLadderShard: after a few exchange rounds every slot of the world-size-2 run holds the same
    configuration, box and energy as the single-process run, decisions are identical on all ranks, and every
    configuration still carries the energy it was uploaded with."""
    import torch.multiprocessing as mp
    n_rep, rounds = 7, 9
    res = {}
    for world in (1, 2):
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_run_resident, args=(r, world, port, n_rep, rounds, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        res[world] = got
    log1 = res[1][0][1]
    assert any(ok for *_, ok in log1) and any(not ok for *_, ok in log1)
    assert any(ok and i == 3 for _, i, _, _, ok in log1), 'no cross-rank swap (slots 3 <-> 4)'
    for _, log, *_ in res[2]:
        assert log == log1
    s1 = res[1][0][2]
    s2 = {}
    for _, _, st, *_ in res[2]:
        s2.update(st)
    assert sorted(s1) == sorted(s2) == list(range(n_rep))
    e_initial = res[1][0][3]
    cfgs = _resident_configs(n_rep)
    moved = 0
    for i in range(n_rep):
        assert np.array_equal(s1[i][0], s2[i][0]) and np.array_equal(s1[i][1], s2[i][1])
        assert s1[i][2] == s2[i][2] and np.array_equal(s1[i][3], s2[i][3])
        src = [g for g in range(n_rep) if cfgs[g][0].shape == s1[i][0].shape and np.array_equal(cfgs[g][0], s1[i][0])]
        assert len(src) == 1                                  # a permutation of the initial configurations ...
        assert s1[i][2] == e_initial[src[0]]                  # ... each still with its own energy
        assert np.array_equal(s1[i][3], cfgs[src[0]][2].reshape(9))
        moved += src[0] != i
    assert moved >= 2
