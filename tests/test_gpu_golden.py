"""GPU: replay the reference's own GCMC runs (tests/golden) through the CUDA calculator and cell.

Given the identical proposals and the identical uniform draws, the accept / reject sequence
must be the reference's, every energy within 1e-10 relative, every free volume identical."""
import os

import numpy as np
import pytest

from mcpy_b200.calculators import B200Calculator
from mcpy_b200.cell import B200SphericalCell
from mcpy_b200.ensembles.acceptance import Units
from test_golden_cpu import GOLD, replay_accepts

pytestmark = pytest.mark.gpu


class _Cfg:
    def __init__(self, pos, num, cell, pbc):
        self.positions, self.numbers, self.cell, self.pbc = pos, num, cell, pbc


def _energies(calc, d):
    off = d['cfg_off']
    out = np.empty(len(off) - 1)
    for t in range(len(out)):
        out[t] = calc.get_potential_energy(_Cfg(d['cfg_pos'][off[t]:off[t + 1]], d['cfg_num'][off[t]:off[t + 1]],
                                                d['cell'], d['pbc']))
    return out


def test_lj_trace_replay():
    d = np.load(os.path.join(GOLD, 'lj_gcmc_trace.npz'))
    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=float(d['rc']))
    E = _energies(calc, d)
    scale = np.maximum(1.0, np.abs(d['cfg_energy']))
    assert np.max(np.abs(E - d['cfg_energy']) / scale) <= 1e-10
    units = Units('LJ', float(d['temperature']), {'H': 1.0})
    flags, _ = replay_accepts(d, E, units, 'H', lambda dn: 'H' if dn else 'X')
    assert np.array_equal(flags, d['accept'])
    # the batched contract gives the same numbers (ragged sizes in one launch)
    off = d['cfg_off']
    batch = [_Cfg(d['cfg_pos'][off[t]:off[t + 1]], d['cfg_num'][off[t]:off[t + 1]], d['cell'], d['pbc'])
             for t in range(0, 64)]
    Eb = calc.get_potential_energies(batch)
    # the per-trial path answers base + delta, the batched path a fresh total: equal to rounding
    assert np.max(np.abs(Eb - E[:64]) / scale[:64]) <= 1e-10
    assert calc.path_counts[1] > 150           # the replay ran on the incremental path
    assert np.array_equal(calc.get_potential_energies(batch, chunk_size=7), Eb)
    assert calc.get_potential_energies([]).shape == (0,)


def test_emt_trace_replay_with_free_volume(fv_engine):
    d = np.load(os.path.join(GOLD, 'emt_sphere_trace.npz'))
    calc = B200Calculator('emt')
    E = _energies(calc, d)
    scale = np.maximum(1.0, np.abs(d['cfg_energy']))
    assert np.max(np.abs(E - d['cfg_energy']) / scale) <= 1e-10
    units = Units('metal', float(d['temperature']), {'O': 15.999})
    flags, _ = replay_accepts(d, E, units, 'O', lambda dn: 'O' if dn else 'X')
    assert np.array_equal(flags, d['accept'])
    # every free-volume refresh of the run, host- and device-generated points
    from ase import Atoms
    voff = d['vol_off']
    for mode in ('host', 'device'):
        for k in range(len(d['vol_value'])):
            pos, num = d['vol_pos'][voff[k]:voff[k + 1]], d['vol_num'][voff[k]:voff[k + 1]]
            atoms = Atoms(numbers=num, positions=pos)
            cell = B200SphericalCell.__new__(B200SphericalCell)
            cell.species_radii = {'Ag': 2.947, 'O': 0.0}
            cell.mc_sample_points = int(d['n_points'])
            cell.radius = float(d['sphere_radius'])
            cell.center = np.zeros(3)
            cell.sphere_volume = (4.0 / 3.0) * np.pi * cell.radius ** 3
            cell.engine, cell.points = fv_engine, mode
            cell._rng = np.random.default_rng(0)
            st = cell._rng.bit_generator.state
            w = [int(x) for x in d['vol_state'][k]]
            st['state'] = {'state': (w[0] << 64) | w[1], 'inc': (w[2] << 64) | w[3]}
            cell._rng.bit_generator.state = st
            cell.calculate_volume(atoms)
            assert cell.get_volume() == d['vol_value'][k], (mode, k)
