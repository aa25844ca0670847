"""CPU: the oracle and the host-side acceptance mirror against fixtures produced by the
UNMODIFIED reference drivers (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from mcpy_b200.ensembles.acceptance import Units, accept_trial
from oracle import cfast, emt, free_volume as ofv, lj

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _draws(d):
    it = iter(d['draw'])
    return it


def replay_accepts(d, energies, units, mu_key, species_of):
    """Re-run the reference's accept/reject chain from energies; returns the accept flags."""
    E_old = energies[0]
    flags = []
    for t in range(len(d['dE'])):
        E_new = energies[t + 1]
        used = {'n': 0}

        def draw(t=t):
            used['n'] += 1
            assert not np.isnan(d['draw'][t]), 'the reference consumed no uniform on this trial'
            return float(d['draw'][t])

        ok = accept_trial(units, {mu_key: float(d['mu'])}, E_new - E_old, int(d['dN'][t]),
                          float(d['volume'][t]), species_of(int(d['dN'][t])), int(d['n'][t]), draw)
        # a uniform is consumed exactly when the reference consumed one
        assert (used['n'] == 1) == (not np.isnan(d['draw'][t]))
        flags.append(ok)
        if ok:
            E_old = E_new
    return np.array(flags), E_old


def test_lj_trace_oracle_and_acceptance_mirror():
    d = np.load(os.path.join(GOLD, 'lj_gcmc_trace.npz'))
    off = d['cfg_off']
    # numpy oracle on a sample, C oracle on all: both equal the recorded energies
    for t in range(len(off) - 1):
        pos = d['cfg_pos'][off[t]:off[t + 1]]
        e = cfast.lj_energy(pos, d['cell'], d['pbc'], 1.0, 1.0, float(d['rc']))
        assert e == d['cfg_energy'][t]
        if t % 40 == 0:
            assert lj.lj_energy(pos, d['cell'], d['pbc'], 1.0, 1.0, float(d['rc'])) == pytest.approx(e, rel=1e-12)
    units = Units('LJ', float(d['temperature']), {'H': 1.0})
    flags, E_final = replay_accepts(d, d['cfg_energy'], units, 'H', lambda dn: 'H' if dn else 'X')
    assert np.array_equal(flags, d['accept'])
    assert E_final == float(d['final_energy'])


def test_emt_trace_oracle_acceptance_and_free_volume():
    d = np.load(os.path.join(GOLD, 'emt_sphere_trace.npz'))
    off = d['cfg_off']
    for t in range(0, len(off) - 1, 9):
        pos, num = d['cfg_pos'][off[t]:off[t + 1]], d['cfg_num'][off[t]:off[t + 1]]
        assert emt.emt_energy(pos, num, d['cell'], d['pbc']) == pytest.approx(d['cfg_energy'][t], rel=1e-12)
    # metal units: beta and Lambda as the reference computed them
    units = Units('metal', float(d['temperature']), {'O': 15.999})
    assert units.beta == float(d['beta']) and units.lambda_dbs['O'] == float(d['lambda_O'])
    flags, E_final = replay_accepts(d, d['cfg_energy'], units, 'O', lambda dn: 'O' if dn else 'X')
    assert np.array_equal(flags, d['accept'])
    assert E_final == float(d['final_energy'])
    # free-volume refreshes: same PCG64 state -> same volume
    voff = d['vol_off']
    for k in range(0, len(d['vol_value']), 7):
        rng = np.random.default_rng(0)
        st = rng.bit_generator.state
        w = [int(x) for x in d['vol_state'][k]]
        st['state'] = {'state': (w[0] << 64) | w[1], 'inc': (w[2] << 64) | w[3]}
        rng.bit_generator.state = st
        pos, num = d['vol_pos'][voff[k]:voff[k + 1]], d['vol_num'][voff[k]:voff[k + 1]]
        radii = np.where(num == 47, 2.947, 0.0)
        vol, _ = ofv.spherical_free_volume(rng, pos, radii, float(d['sphere_radius']), int(d['n_points']))
        assert vol == d['vol_value'][k]
