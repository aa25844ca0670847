"""GPU: device-resident GCMC chains (row f2, ``csrc/gcmc_chain.cuh``) against the UNMODIFIED reference loop.

Two copies of the same reference ``GrandCanonicalEnsemble`` (``baseline/_ref``), same seeds: one advanced by the
reference's own ``do_gcmc_step`` (``grand_canonical_ensemble.py:244-308``) over the CPU oracle calculator, one handed
to ``ResidentGCMC``.  Asserted: the same move, decision and particle number at EVERY trial, dE within
``1e-10 max(1, |dE|) + 1e-13 |E|``, and afterwards the same configuration to the bit, the same five Mersenne-Twister
states and PCG64 state (so the next host trial draws the same numbers), the same counters and the same running minimum.
"""
import numpy as np
import pytest

from helpers import OracleLJCalculator, de_tol
from mcpy_b200.utils import synthetic

pytestmark = pytest.mark.gpu


def box_gcmc(cfg, symbol, mu, seed=0, n_moves=3, weights=(1, 1, 1), max_atoms=None, min_atoms=None, temperature=None):
    from ase import Atoms
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    atoms = Atoms(numbers=cfg['numbers'], positions=cfg['positions'].copy(), cell=cfg['cell'], pbc=cfg['pbc'])
    calc = OracleLJCalculator(rc=cfg['rc'], sigma=cfg['sigma'], epsilon=cfg['epsilon'])
    cell = Cell(atoms, seed=21 + seed)
    moves = [InsertionMove(cell, [symbol], 31 + seed, max_atoms=max_atoms),
             DeletionMove(cell, [symbol], 32 + seed, min_atoms=min_atoms),
             DisplacementMove([symbol], 33 + seed, max_displacement=cfg['max_displacement'])]
    ms = MoveSelector(list(weights), moves, seed=34 + seed, n_moves=n_moves)
    return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type=cfg['units'], calculator=calc, mu={symbol: mu},
                                  species=[symbol], temperature=temperature or cfg['temperature'], move_selector=ms,
                                  random_seed=35 + seed, traj_file=None, outfile=None)


def small_box(n=40, L=7.0, seed=5, rc=2.5):
    pos = synthetic.jittered_lattice(n, L, seed)
    return dict(positions=pos, numbers=np.ones(n, np.int32), cell=np.eye(3) * L, pbc=np.ones(3, bool), rc=rc, sigma=1.0,
                epsilon=1.0, units='LJ', temperature=2.0, max_displacement=0.3)


def host_trace(g):
    """(move, dE, accepted, N after) of every scored trial + the failed proposals, in order, of the reference loop."""
    trace = []
    ms = g.move_selector
    inner_acc = g._acceptance_condition
    inner_trial = ms.do_trial_move

    def trial(atoms):
        res = inner_trial(atoms)
        a = res[0] if isinstance(res, tuple) else res
        if a is False or a is None:
            trace.append([ms.to_use, float('nan'), False, True, len(g.atoms), float(g.E_old)])
        return res

    def acc(dE, dN, volume, species, n):
        ok = inner_acc(dE, dN, volume, species, n)
        trace.append([ms.to_use, float(dE), bool(ok), False, n + (dN if ok else 0), float(g.E_old)])
        return ok
    ms.do_trial_move = trial
    g._acceptance_condition = acc
    return trace


def gen_states(g):
    ms = g.move_selector
    gens = [ms.rng.random] + [m.rng.random for m in ms.move_list] + [g.rng_acceptance.random]
    cells = [m.cell for m in ms.move_list if hasattr(m.cell, '_rng')]      # the cells the moves draw their points from
    return [r.getstate() for r in gens], [c._rng.bit_generator.state for c in cells]


def assert_same_chain(g_ref, g_dev, exact_positions=True):
    assert len(g_ref.atoms) == len(g_dev.atoms) == g_ref.n_atoms == g_dev.n_atoms
    if exact_positions:
        assert np.array_equal(g_ref.atoms.positions, g_dev.atoms.positions)
    else:
        assert np.allclose(g_ref.atoms.positions, g_dev.atoms.positions, rtol=0, atol=1e-9)
    assert np.array_equal(g_ref.atoms.numbers, g_dev.atoms.numbers)
    assert gen_states(g_ref) == gen_states(g_dev)
    a, b = g_ref.move_selector, g_dev.move_selector
    for name in ('move_counter', 'move_acceptance', 'move_failed_counter', 'move_counter_total',
                 'move_acceptance_total', 'move_failed_counter_total'):
        assert getattr(a, name) == getattr(b, name), name
    assert a.to_use == b.to_use
    assert abs(g_ref.E_old - g_dev.E_old) <= 1e-9 * max(1.0, abs(g_ref.E_old))


def assert_same_trace(ref, dev):
    assert len(ref) == len(dev), (len(ref), len(dev))
    for k, (r, d) in enumerate(zip(ref, dev)):
        failed = bool(d['flags'] & 2)
        assert (r[0], r[2], r[3], r[4]) == (int(d['move']), bool(d['flags'] & 1), failed, int(d['n_atoms'])), \
            f'trial {k}: reference {r} vs device {d}'
        if not failed:
            assert abs(r[1] - d['dE']) <= de_tol(r[1], r[5]), f'trial {k}: dE {r[1]!r} vs {d["dE"]!r}'


def run_pair(make, steps, **kw):
    from mcpy_b200.ensembles import ResidentGCMC
    g_ref, g_dev = make(), make()
    trace = host_trace(g_ref)
    for _ in range(steps):
        g_ref.do_gcmc_step()
    n_trials = steps * g_dev.move_selector.n_moves
    cfg = kw.pop('cfg')
    rg = ResidentGCMC(g_dev, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc']), trace=n_trials, **kw)
    rg.run_trials(n_trials)
    assert_same_trace(trace, rg.traces[0])
    assert_same_chain(g_ref, g_dev)
    return g_ref, g_dev, rg, trace


def test_small_box_every_trial_and_the_state_handed_back():
    cfg = small_box()
    g_ref, g_dev, rg, trace = run_pair(lambda: box_gcmc(cfg, 'H', -2.5), 400, cfg=cfg)
    accepted = sum(t[2] for t in trace)
    kinds = {t[0] for t in trace if t[2]}
    assert accepted > 100 and kinds == {0, 1, 2}, (accepted, kinds)          # all three moves were accepted
    assert rg.info()['in_shared_memory']
    # the running minimum
    assert g_dev._best_score == pytest.approx(g_ref._best_score, rel=1e-10, abs=1e-10)
    assert len(g_dev._best_atoms) == len(g_ref._best_atoms)
    assert np.array_equal(g_dev._best_atoms.positions, g_ref._best_atoms.positions)


def test_c1_argon():
    cfg = synthetic.s1_argon()
    run_pair(lambda: box_gcmc(cfg, 'Ar', -0.085), 50, cfg=cfg)


def test_c2_lj_2000_attractive_and_repulsive():
    for rc in (3.0, 2.0 ** (1.0 / 6.0)):
        cfg = synthetic.s2_lj(rc=rc)
        g_ref, g_dev, rg, trace = run_pair(lambda: box_gcmc(cfg, 'H', -1.0, n_moves=4), 40, cfg=cfg)
        assert sum(t[2] for t in trace) > 10


def test_device_and_host_blocks_alternate_on_one_chain():
    """device block, the reference's own loop, device block == the reference's loop alone."""
    from mcpy_b200.ensembles import ResidentGCMC
    cfg = small_box(seed=9)
    g_ref, g_dev = box_gcmc(cfg, 'H', -2.0, seed=3), box_gcmc(cfg, 'H', -2.0, seed=3)
    for _ in range(90):
        g_ref.do_gcmc_step()
    rg = ResidentGCMC(g_dev, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc']))
    rg.run_trials(30 * 3)
    for _ in range(30):
        g_dev.do_gcmc_step()
    rg.run_trials(30 * 3)
    assert_same_chain(g_ref, g_dev)


def test_many_chains_one_launch():
    from mcpy_b200.ensembles import ResidentGCMC
    cfg = small_box(seed=11)
    mus = [-3.0, -2.5, -2.0, -1.5, -1.0, -0.5]
    refs = [box_gcmc(cfg, 'H', mu, seed=10 * k) for k, mu in enumerate(mus)]
    devs = [box_gcmc(cfg, 'H', mu, seed=10 * k) for k, mu in enumerate(mus)]
    for g in refs:
        for _ in range(60):
            g.do_gcmc_step()
    rg = ResidentGCMC(devs, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc']))
    rg.run_trials(180)
    for a, b in zip(refs, devs):
        assert_same_chain(a, b)
    assert len({len(g.atoms) for g in devs}) > 1


def test_growth_limits_and_an_emptied_box():
    from mcpy_b200.ensembles import ResidentGCMC
    # high mu: the box fills up, the chain stops at its capacities, is rebuilt larger and carries on
    cfg = small_box(n=20, seed=13)
    mk = lambda: box_gcmc(cfg, 'H', 4.0, seed=1, weights=(3, 1, 1), n_moves=5, max_atoms=150)  # noqa: E731
    g_ref, g_dev = mk(), mk()
    for _ in range(120):
        g_ref.do_gcmc_step()
    rg = ResidentGCMC(g_dev, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc']), atom_capacity=32, cell_capacity=4)
    rg.run_trials(600)
    assert_same_chain(g_ref, g_dev)
    assert len(g_dev.atoms) > 60
    assert g_dev.move_selector.move_failed_counter_total[0] > 0 or len(g_dev.atoms) < 150
    # very low mu: the box empties; deletions of an empty box fail, min_atoms holds; no displacement move in the list
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, InsertionMove, MoveSelector
    from ase import Atoms

    def emptied(min_atoms):
        c = small_box(n=6, seed=17)
        atoms = Atoms(numbers=c['numbers'], positions=c['positions'].copy(), cell=c['cell'], pbc=c['pbc'])
        cell = Cell(atoms, seed=2)
        ms = MoveSelector([1, 3], [InsertionMove(cell, ['H'], 3), DeletionMove(cell, ['H'], 4, min_atoms=min_atoms)],
                          seed=5, n_moves=2)
        return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=OracleLJCalculator(rc=2.5),
                                      mu={'H': -12.0}, species=['H'], temperature=1.0, move_selector=ms, random_seed=6,
                                      traj_file=None, outfile=None)
    for min_atoms in (None, 2):
        g_ref, g_dev = emptied(min_atoms), emptied(min_atoms)
        for _ in range(60):
            g_ref.do_gcmc_step()
        ResidentGCMC(g_dev, lj=(1.0, 1.0, 2.5)).run_trials(120)
        assert_same_chain(g_ref, g_dev)
        assert len(g_dev.atoms) == (0 if min_atoms is None else 2)
        assert g_dev.move_selector.move_failed_counter_total[1] > 0


def test_run_writes_the_same_files_as_the_reference(tmp_path):
    """``ResidentGCMC.run(steps)`` against ``gcmc.run(steps)``: outfile rows and trajectory frames."""
    from mcpy_b200.ensembles import ResidentGCMC
    cfg = small_box(seed=19)

    def make(tag):
        g = box_gcmc(cfg, 'H', -2.0, seed=7)
        g._outfile, g._traj_file = str(tmp_path / f'{tag}.out'), str(tmp_path / f'{tag}.xyz')
        g._outfile_write_interval, g._trajectory_write_interval = 5, 10
        return g
    g_ref, g_dev = make('ref'), make('dev')
    g_ref.run(40)
    ResidentGCMC(g_dev, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc'])).run(40)
    assert_same_chain(g_ref, g_dev)
    rows_ref = [ln.split() for ln in open(tmp_path / 'ref.out') if ln[:1].isdigit()]
    rows_dev = [ln.split() for ln in open(tmp_path / 'dev.out') if ln[:1].isdigit()]
    assert len(rows_ref) == len(rows_dev) >= 9
    for a, b in zip(rows_ref, rows_dev):
        assert a[:2] == b[:2] and a[3:] == b[3:], (a, b)                     # step, N, acceptance ratios of the interval
        assert abs(float(a[2]) - float(b[2])) <= 2e-6 * max(1.0, abs(float(a[2])))
    xyz_ref = [ln for ln in open(tmp_path / 'ref.xyz') if not ln.startswith('energy=')]
    xyz_dev = [ln for ln in open(tmp_path / 'dev.xyz') if not ln.startswith('energy=')]
    assert xyz_ref == xyz_dev                                                # positions print to the last digit


def test_out_of_scope_chains_are_refused():
    from mcpy_b200.ensembles import ResidentGCMC
    from mcpy_b200.ensembles.resident_gcmc import supports
    cfg = small_box()
    g = box_gcmc(cfg, 'H', -2.0)
    assert supports(g, lj=(1.0, 1.0, 2.5)) is None
    assert 'exceed 2 rc' in supports(g, lj=(1.0, 1.0, 3.6))
    g.atoms.pbc = [True, True, False]
    assert 'periodic' in supports(g, lj=(1.0, 1.0, 2.5))
    g = box_gcmc(cfg, 'H', -2.0)
    g.move_selector.move_list[0].min_insert = 0.5
    with pytest.raises(ValueError, match='min_insert'):
        ResidentGCMC(g, lj=(1.0, 1.0, 2.5))
    g = box_gcmc(cfg, 'H', -2.0)
    with pytest.raises(ValueError, match='Lennard-Jones parameters'):
        ResidentGCMC(g)


@pytest.mark.parametrize('kind', ['T', 'mu'])
def test_replica_exchange_with_resident_replicas_equals_the_reference_driver(kind, tmp_path):
    """The reference's ``BatchedReplicaExchange`` over the CPU oracle against ``ResidentReplicaExchange`` (all replicas
    advance in one launch per block, swaps of local replicas happen on the device): same swap decisions, same replicas."""
    from helpers import make_factory
    from mcpy.ensembles import BatchedReplicaExchange
    from mcpy_b200.ensembles import ResidentReplicaExchange
    calc = OracleLJCalculator()
    factory = make_factory(calc, n0=24, base_seed=400, n_moves=3)
    ladder = dict(temperatures=[1.5, 1.8, 2.2, 2.7, 3.3, 4.0]) if kind == 'T' else \
        dict(mus=[{'H': m} for m in (-3.2, -2.9, -2.6, -2.3, -2.0, -1.7)])
    ref = BatchedReplicaExchange(factory, calc, gcmc_steps=45, exchange_interval=4, seed=31, write_out_interval=10,
                                 outfile=str(tmp_path / 'ref.out'), global_minimum_file=str(tmp_path / 'ref_min.xyz'),
                                 **ladder)
    ref.run()
    mine = ResidentReplicaExchange(factory, calc, gcmc_steps=45, exchange_interval=4, seed=31, write_out_interval=10,
                                   outfile=str(tmp_path / 'dev.out'), global_minimum_file=str(tmp_path / 'dev_min.xyz'),
                                   rank=0, world_size=1, lj=(1.0, 1.0, 2.5), **ladder)
    mine.run()
    assert mine.exchange_attempts == ref.exchange_attempts
    assert mine.exchange_successes == ref.exchange_successes
    assert sum(mine.exchange_successes) > 2
    assert mine.blocks < 45                                                  # several steps per launch
    for a, b in zip(ref.replicas, mine.replicas):
        assert_same_chain(a, b)
        assert a._step == b._step
    # the replica-exchange outfile: replica, step and atom counts of every row
    rows = [[ln.split()[:4] for ln in open(tmp_path / f'{t}.out') if ln[:1].isdigit()] for t in ('ref', 'dev')]
    assert rows[0] == rows[1] and len(rows[0]) >= 6 * 5
    lines = [[ln for ln in open(tmp_path / f'{t}_min.xyz')][2:] for t in ('ref', 'dev')]
    assert lines[0] == lines[1] and len(lines[0]) > 0                        # the global minimum, atom for atom


def test_reference_lammps_parity_benchmark_stage1_series():
    """The reference's own LJ GCMC benchmark (``benchmark/lammps_gcmc_parity.py:205-221,331-345``: L = 9, rc = 3, T* = 2,
    20 randomly placed atoms, insertion / deletion(min_atoms=1) / displacement 1:1:1, one trial per step, the Cell built
    from a separate empty Atoms): its per-step series N_i and E_i, from the device chain's per-trial trace."""
    from ase import Atoms
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    from mcpy_b200.ensembles import ResidentGCMC
    L, RC, T, steps = 9.0, 3.0, 2.0, 3000

    def make(mu):
        pos = np.random.default_rng(25).uniform(0.9, L - 0.9, size=(20, 3))
        atoms = Atoms('H20', positions=pos, cell=[L, L, L], pbc=True)
        cell = Cell(Atoms(cell=[L, L, L], pbc=True), seed=7)       # (the benchmark leaves it unseeded)
        moves = [InsertionMove(cell, species=['H'], seed=21), DeletionMove(cell, species=['H'], seed=22, min_atoms=1),
                 DisplacementMove(species=['H'], seed=23, max_displacement=0.3)]
        ms = MoveSelector([1, 1, 1], moves, seed=25, n_moves=1)
        return GrandCanonicalEnsemble(atoms=atoms, cells=[Cell(atoms)], units_type='LJ',
                                      calculator=OracleLJCalculator(rc=RC), mu={'H': mu}, species=['H'], temperature=T,
                                      move_selector=ms, random_seed=24, traj_file=None, outfile=None)
    for mu in (-3.0, -1.0):
        g_ref, g_dev = make(mu), make(mu)
        n_series, e_series = np.empty(steps), np.empty(steps)
        for i in range(steps):                                       # run_mcpy of the benchmark
            g_ref.do_gcmc_step()
            n_series[i], e_series[i] = len(g_ref.atoms), g_ref.E_old
        E0 = g_dev.E_old
        rg = ResidentGCMC(g_dev, lj=(1.0, 1.0, RC), trace=steps)
        rg.run_trials(steps)
        tr = rg.traces[0]
        assert np.array_equal(tr['n_atoms'], n_series)
        e_dev = E0 + np.cumsum(np.where(tr['flags'] & 1, tr['dE'], 0.0))
        scale = np.maximum.accumulate(np.maximum(np.abs(e_series), abs(E0)))      # the running sum rounds at the largest |E| so far
        assert np.all(np.abs(e_dev - e_series) <= 1e-10 * np.maximum(1.0, np.abs(e_series)) + 1e-13 * scale)
        assert_same_chain(g_ref, g_dev)
        assert n_series.max() > 25 and n_series.min() >= 1


def test_state_in_global_memory_when_it_does_not_fit_shared_memory():
    """A chain whose storage (here: room for 20 000 atoms) exceeds the 227 KB of shared memory runs from global memory
    with the same code: same trials, same state."""
    cfg = small_box(seed=23)
    g_ref, g_dev, rg, trace = run_pair(lambda: box_gcmc(cfg, 'H', -2.2, seed=4), 150, cfg=cfg, atom_capacity=20000)
    assert not rg.info()['in_shared_memory'] and rg.info()['state_bytes'] > 227 * 1024
    assert sum(t[2] for t in trace) > 50


@pytest.mark.parametrize('box', [(7.0, 9.0, 12.5), (5.5, 6.4, 5.2), (16.0, 5.3, 8.1)])
def test_anisotropic_boxes_two_insertion_cells_fractional_weights_unwrapped_start(box):
    """Boxes whose axes fall on different sides of the five-cell stencil (fewer than five cells along an axis: every
    cell of that axis is a neighbour), two insertion moves drawing from two cells' PCG64 streams, non-integer move
    weights, deletion and insertion limits, and a start with atoms one or two boxes outside the cell."""
    from ase import Atoms
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    from mcpy_b200.ensembles import ResidentGCMC
    rc = 2.5

    def make():
        rng = np.random.default_rng(77)
        n = 30
        pos = rng.uniform(0.0, 1.0, (n, 3)) * np.array(box)
        pos[::5] += np.array(box) * rng.integers(-2, 3, (len(pos[::5]), 3))        # unwrapped images
        atoms = Atoms(numbers=np.ones(n, int), positions=pos, cell=np.diag(box), pbc=True)
        cell_a, cell_b = Cell(atoms, seed=5), Cell(atoms, seed=6)
        moves = [InsertionMove(cell_a, ['H'], 41, max_atoms=60), InsertionMove(cell_b, ['H'], 42),
                 DeletionMove(cell_a, ['H'], 43, min_atoms=8), DisplacementMove(['H'], 44, max_displacement=1.7)]
        ms = MoveSelector([0.7, 0.45, 1.3, 2.05], moves, seed=45, n_moves=4)
        return GrandCanonicalEnsemble(atoms=atoms, cells=[cell_a, cell_b], units_type='LJ',
                                      calculator=OracleLJCalculator(rc=rc), mu={'H': -1.2}, species=['H'], temperature=1.7,
                                      move_selector=ms, random_seed=46, traj_file=None, outfile=None)
    g_ref, g_dev = make(), make()
    trace = host_trace(g_ref)
    steps = 150
    for _ in range(steps):
        g_ref.do_gcmc_step()
    rg = ResidentGCMC(g_dev, lj=(1.0, 1.0, rc), trace=4 * steps)
    rg.run_trials(4 * steps)
    assert_same_trace(trace, rg.traces[0])
    assert_same_chain(g_ref, g_dev)
    assert {t[0] for t in trace if t[2]} == {0, 1, 2, 3}
    cells = rg.info()['cells']
    assert min(cells) < 5 or max(cells) >= 5


def test_install_routes_an_unmodified_run_call_through_the_device(tmp_path):
    """``resident_gcmc.install()``: ``gcmc.run(steps)`` of an unmodified script with a ``B200Calculator('lj')`` runs on
    the device (one launch per block between file writes instead of one calculator call per trial), an out-of-scope
    ensemble keeps the reference's loop, and both leave the state and files the reference's loop over the oracle leaves."""
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy_b200.calculators import B200Calculator
    from mcpy_b200.ensembles import resident_gcmc
    cfg = small_box(seed=29)
    calc = B200Calculator('lj', epsilon=cfg['epsilon'], sigma=cfg['sigma'], rc=cfg['rc'])

    def make(tag, calculator=None, min_insert=None):
        g = box_gcmc(cfg, 'H', -2.1, seed=8)
        if calculator is not None:
            g._calculator = calculator
            g.E_old = g.compute_energy(g.atoms)
        g.move_selector.move_list[0].min_insert = min_insert
        g._outfile, g._traj_file = str(tmp_path / f'{tag}.out'), str(tmp_path / f'{tag}.xyz')
        g._outfile_write_interval, g._trajectory_write_interval = 10, 20
        return g
    g_ref = make('ref')
    g_ref.run(60)                                                   # the reference's loop over the CPU oracle
    original = resident_gcmc.install()
    try:
        assert GrandCanonicalEnsemble.run is not original
        g_dev = make('dev', calc)
        calls0 = calc.n_energy_calls
        g_dev.run(60)
        assert calc.n_energy_calls == calls0                        # not one calculator call: the chain ran on the device
        assert_same_chain(g_ref, g_dev)
        xyz = [[ln for ln in open(tmp_path / f'{t}.xyz') if not ln.startswith('energy=')] for t in ('ref', 'dev')]
        assert xyz[0] == xyz[1] and len(xyz[0]) > 100
        g_out = make('out', calc, min_insert=0.3)                   # out of scope: the reference's own loop, per-trial calls
        g_out.run(5)
        assert calc.n_energy_calls >= calls0 + 5
    finally:
        resident_gcmc.uninstall(original)
    assert 'run' not in GrandCanonicalEnsemble.__dict__ or GrandCanonicalEnsemble.run is original


def test_run_with_the_reference_defaults_a_row_and_a_frame_every_step(tmp_path):
    """``outfile_write_interval = trajectory_write_interval = 1`` (the reference's defaults): one launch per step, the
    positions of every step come back for the frame while generators and the rest stay on the device; same files."""
    from mcpy_b200.ensembles import ResidentGCMC
    cfg = small_box(seed=31)

    def make(tag):
        g = box_gcmc(cfg, 'H', -1.8, seed=9)
        g._outfile, g._traj_file = str(tmp_path / f'{tag}.out'), str(tmp_path / f'{tag}.xyz')
        g._outfile_write_interval, g._trajectory_write_interval = 1, 1
        return g
    g_ref, g_dev = make('ref'), make('dev')
    g_ref.run(35)
    rg = ResidentGCMC(g_dev, lj=(cfg['epsilon'], cfg['sigma'], cfg['rc']))
    rg.run(35)
    assert_same_chain(g_ref, g_dev)
    frames = [[ln for ln in open(tmp_path / f'{t}.xyz') if not ln.startswith('energy=')] for t in ('ref', 'dev')]
    assert frames[0] == frames[1] and sum(1 for ln in frames[0] if ln.strip().isdigit()) == 37      # step 0 .. 35 + final
    rows = [[ln.split()[:2] + ln.split()[3:] for ln in open(tmp_path / f'{t}.out') if ln[:1].isdigit()] for t in ('ref', 'dev')]
    assert rows[0] == rows[1] and len(rows[0]) >= 36


def test_long_run_in_one_launch_equals_many_launches_with_full_hand_backs():
    """120 000 trials of a 500-atom argon chain in ONE launch against the same trials in 120 launches with the whole state
    handed back to the ensemble and uploaded again in between (generator states through ``setstate`` / ``getstate``,
    configuration, cell lists rebuilt): the same chain to the bit, and its energy is the oracle's for the final
    configuration (the running sum is re-summed on the device when its rounding bound says so)."""
    from oracle import cfast
    from mcpy_b200.ensembles import ResidentGCMC
    cfg = synthetic.s1_argon()
    lj = (cfg['epsilon'], cfg['sigma'], cfg['rc'])
    a, b = box_gcmc(cfg, 'Ar', -0.085, seed=12), box_gcmc(cfg, 'Ar', -0.085, seed=12)
    ra, rb = ResidentGCMC(a, lj=lj), ResidentGCMC(b, lj=lj)
    ra.run_trials(120_000)
    for _ in range(120):
        rb.run_trials(1000)
    assert len(a.atoms) == len(b.atoms) and np.array_equal(a.atoms.positions, b.atoms.positions)
    assert gen_states(a) == gen_states(b)
    assert a.move_selector.move_acceptance_total == b.move_selector.move_acceptance_total
    assert sum(a.move_selector.move_acceptance_total) > 5000
    e = cfast.lj_energy(a.atoms.positions, np.asarray(a.atoms.cell), a.atoms.pbc, cfg['sigma'], cfg['epsilon'], cfg['rc'])
    for g in (a, b):
        assert abs(g.E_old - e) <= 1e-10 * max(1.0, abs(e)), (g.E_old, e)
