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


def _run(rank, world, port, n_rep, steps, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import importlib.util
    if importlib.util.find_spec('ase') is None:
        sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ase_shim'))
    import torch.distributed as dist
    from helpers import OracleLJCalculator, make_factory
    from mcpy_b200.ensembles import ShardedReplicaExchange
    if world > 1:
        dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    temps = list(np.linspace(1.5, 3.0, n_rep))
    re = ShardedReplicaExchange(make_factory(), OracleLJCalculator(), temperatures=temps,
                                gcmc_steps=steps, exchange_interval=3, seed=31,
                                rank=rank, world_size=world)
    re.run()
    state = {re.lo + k: (r.E_old, r.n_atoms, r.atoms.positions.copy(), r.volume_refreshes)
             for k, r in enumerate(re.replicas)}
    out.put((rank, re.swap_log, state))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _launch(world, n_rep=5, steps=16):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, n_rep, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(results, key=lambda t: t[0])


def test_two_rank_gloo_run_equals_single_process_run():
    single = _launch(1)
    double = _launch(2)
    log1 = single[0][1]
    assert len(log1) > 0 and any(a for *_, a in log1), 'no swap was attempted / accepted'
    # every rank made the same decisions as the single-process run
    for _, log, _ in double:
        assert log == log1
    state1 = single[0][2]
    state2 = {}
    for _, _, st in double:
        state2.update(st)
    assert sorted(state2) == sorted(state1)
    for i in state1:
        assert state2[i][0] == state1[i][0] and state2[i][1] == state1[i][1]
        assert np.array_equal(state2[i][2], state1[i][2])
        assert state2[i][3] == state1[i][3]          # free-volume refreshes after swaps too


@pytest.mark.skipif(not os.path.isdir('/root/reference/mcpy'), reason='reference checkout not present')
def test_single_rank_run_equals_reference_batched_replica_exchange():
    """Same factory, same seeds: the reference's BatchedReplicaExchange and this driver end in
    identical states after identical swap decisions (reference rows a5-a7 of SURVEY.md 8a)."""
    sys.path.insert(0, '/root/reference')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from ase import Atoms
    from helpers import OracleLJCalculator
    from mcpy.cell import Cell
    from mcpy.ensembles import BatchedReplicaExchange, GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    from mcpy_b200.ensembles import ShardedReplicaExchange

    L = 7.0

    def factory(T=None, mu=None, rank=0):
        rng = np.random.default_rng(200 + rank)
        atoms = Atoms('H20', positions=rng.uniform(0.5, L - 0.5, size=(20, 3)), cell=[L, L, L], pbc=True)
        cell = Cell(atoms, seed=300 + rank)
        moves = [InsertionMove(cell, species=['H'], seed=11 + rank), DeletionMove(cell, species=['H'], seed=12 + rank),
                 DisplacementMove(species=['H'], seed=13 + rank, max_displacement=0.3)]
        ms = MoveSelector([1, 1, 1], moves, seed=14 + rank, n_moves=2)
        return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=calc,
                                      mu=mu if mu is not None else {'H': -2.5}, species=['H'],
                                      temperature=T if T is not None else 2.0, move_selector=ms,
                                      random_seed=40 + rank, traj_file=None, outfile=None)

    calc = OracleLJCalculator()
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
