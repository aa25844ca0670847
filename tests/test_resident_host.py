"""CPU: the host side of the device-resident chains -- what is in scope, and the block planning / file writes / swaps of
``ResidentReplicaExchange`` with the replicas advanced by a host stand-in (the reference's own ``do_gcmc_step``) instead
of the CUDA chains.  The chains themselves are tested on the GPU (``tests/test_gpu_resident_gcmc.py``)."""
import numpy as np
import pytest

from helpers import OracleLJCalculator, make_factory


class HostChains:
    """Same interface as ``ResidentGCMC``; every block is the reference's loop on the host."""

    def __init__(self, ensembles, engine=None, lj=None, atom_capacity=None, cell_capacity=0):
        self.ensembles = list(ensembles)
        self.blocks = []
        self.frames = []
        self.swaps = 0

    def advance(self, n_trials, sync='full', counters=True):
        n_trials = np.broadcast_to(np.asarray(n_trials), (len(self.ensembles),))
        self.blocks.append((n_trials.tolist(), sync))
        for g, t in zip(self.ensembles, n_trials):
            assert t % g.move_selector.n_moves == 0
            for _ in range(int(t) // g.move_selector.n_moves):
                g.do_gcmc_step()
        return 0.0

    def swap_configurations(self, i, j):
        self.swaps += 1
        return False

    def sync_host(self, chains=None):
        pass

    def sync_positions(self, chains=None):
        self.frames.append((len(self.blocks) - 1, list(chains)))

    def close(self):
        pass


def test_what_is_in_scope():
    from mcpy_b200.ensembles.resident_gcmc import supports
    f = make_factory(OracleLJCalculator())
    lj = (1.0, 1.0, 2.5)
    assert supports(f(rank=0), lj) is None
    assert 'Lennard-Jones parameters' in supports(f(rank=0))
    assert 'exceed 2 rc' in supports(f(rank=0), (1.0, 1.0, 3.6))
    g = f(rank=0)
    g.atoms.pbc = [True, False, True]
    assert 'periodic' in supports(g, lj)
    g = f(rank=0)
    g.atoms.set_cell([[7, 0, 0], [1, 7, 0], [0, 0, 7]])
    assert 'orthorhombic' in supports(g, lj)
    assert 'constraints' in supports(make_factory(OracleLJCalculator(), fix=(0, 1))(rank=0), lj)
    g = f(rank=0)
    g.atoms.numbers[3] = 2
    assert 'every atom' in supports(g, lj)
    g = f(rank=0)
    g.move_selector.move_list[0].min_insert = 0.4
    assert 'min_insert' in supports(g, lj)
    g = f(rank=0)
    g.move_selector.move_list[2].n_steps = 2
    assert 'n_steps' in supports(g, lj)
    g = f(rank=0)
    g._minima_file = 'minima.xyz'
    assert 'minima_file' in supports(g, lj)
    g = f(rank=0)
    g.atoms.set_array('molecule_id', np.full(len(g.atoms), -1))
    assert 'molecule_id' in supports(g, lj)


@pytest.mark.parametrize('kind', ['T', 'mu'])
def test_block_planning_files_and_swaps_equal_the_reference_driver(kind, tmp_path):
    """Blocks end exactly where the host looks at the replicas: every exchange attempt, every row of the replica-exchange
    outfile, every replica's own outfile row / trajectory frame, the end of the run -- and the run equals
    ``BatchedReplicaExchange`` step for step, files included."""
    from mcpy.ensembles import BatchedReplicaExchange
    from mcpy_b200.ensembles import ResidentReplicaExchange
    calc = OracleLJCalculator()
    base = make_factory(calc, n0=20, base_seed=600, n_moves=2)

    def factory_for(tag):
        def factory(rank=0, **kw):
            g = base(rank=rank, **kw)
            if rank == 1:                                   # one replica keeps its own outfile and trajectory
                g._outfile, g._traj_file = str(tmp_path / f'{tag}_r1.out'), str(tmp_path / f'{tag}_r1.xyz')
                g._outfile_write_interval, g._trajectory_write_interval = 7, 11
            return g
        return factory
    ladder = dict(temperatures=[1.6, 2.0, 2.5, 3.1]) if kind == 'T' else dict(mus=[{'H': m} for m in (-3.0, -2.6, -2.2, -1.8)])
    ref = BatchedReplicaExchange(factory_for('ref'), calc, gcmc_steps=38, exchange_interval=5, seed=31, write_out_interval=12,
                                 outfile=str(tmp_path / 'ref.out'), global_minimum_file=None, **ladder)
    ref.run()

    class Mine(ResidentReplicaExchange):
        chains_class = HostChains
    mine = Mine(factory_for('dev'), calc, gcmc_steps=38, exchange_interval=5, seed=31, write_out_interval=12,
                outfile=str(tmp_path / 'dev.out'), global_minimum_file=None, rank=0, world_size=1, **ladder)
    mine.run()
    assert mine.exchange_attempts == ref.exchange_attempts and mine.exchange_successes == ref.exchange_successes
    assert sum(mine.exchange_successes) > 0 and mine.resident.swaps == sum(mine.exchange_successes) // 2
    for a, b in zip(ref.replicas, mine.replicas):
        assert a.E_old == b.E_old and a.n_atoms == b.n_atoms and a._step == b._step == 38
        assert np.array_equal(a.atoms.positions, b.atoms.positions)
        assert a.move_selector.move_counter_total == b.move_selector.move_counter_total
    for name in ('.out', '_r1.out', '_r1.xyz'):
        assert open(tmp_path / f'ref{name}').read() == open(tmp_path / f'dev{name}').read(), name
    # the blocks: steps 0..37; the host looks after step s when s % 5 == 0 (s > 0), s % 12 == 0, replica 1's step count
    # (s + 1) is a multiple of 7 or 11, and at the end
    looks = sorted({s for s in range(38) if (s > 0 and s % 5 == 0) or s % 12 == 0 or (s + 1) % 7 == 0 or (s + 1) % 11 == 0} | {37})
    sizes = np.diff([-1] + looks).tolist()
    assert [b[0][0] // 2 for b in mine.resident.blocks] == sizes
    # everything comes back at the end; where replica 1 writes a frame only its positions do; scalars suffice elsewhere
    assert [looks[k] for k, b in enumerate(mine.resident.blocks) if b[1] == 'full'] == [37]
    assert [(looks[k], chains) for k, chains in mine.resident.frames] == [(s, [1]) for s in looks if (s + 1) % 11 == 0 and s != 37]
