"""GPU parity: EMT energies, density sums and trial delta-E vs the oracle (1e-10 relative)."""
import numpy as np
import pytest

from helpers import de_tol
from mcpy_b200.utils import synthetic
from oracle import cfast
from oracle import emt as oemt
from oracle.neighbors import neighbor_pairs

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _energy(engine, pos, num, cell, pbc):
    return engine.potential_energies([pos], [np.asarray(num, np.int32)], np.asarray(cell, float).reshape(1, 9),
                                     np.asarray(pbc, bool).reshape(1, 3))[0]


def test_derived_constants_match_oracle(emt_engine):
    rc, rc_list, acut = oemt.cutoffs()
    assert emt_engine.cutoff == rc_list
    for sym, Z in oemt.ATOMIC_NUMBERS.items():
        p = oemt.species_parameters(Z)
        ref = np.array([p['E0'], p['s0'], p['V0'], p['eta2'], p['kappa'], p['lambda'], p['n0'],
                        p['gamma1'], p['gamma2']])
        np.testing.assert_allclose(emt_engine.emt_species(Z), ref, rtol=1e-14, atol=0)


def test_known_answers(emt_engine):
    z = np.zeros((3, 3))
    # docs/make_figures.py:334: E_EMT(O2)/2 = +0.46 eV (g2 O2, d = 1.245956)
    e = _energy(emt_engine, np.array([[0, 0, 0], [0, 0, 1.245956]], float), [8, 8], z, [0, 0, 0])
    assert abs(e / 2 - 0.46) < 0.005
    assert abs(e - 0.9226813251586565) <= RTOL
    # ASE N2/Cu tutorial: E(N2, d = 1.10) = 0.440343572
    e = _energy(emt_engine, np.array([[0, 0, 0], [0, 0, 1.10]], float), [7, 7], z, [0, 0, 0])
    assert abs(e - 0.440343572) < 5e-9
    # isolated atom: log(0) branch, E = -E0
    e = _energy(emt_engine, np.array([[0.3, 0.1, 0.2]]), [29], z, [0, 0, 0])
    assert e == 3.51


def test_bulk_and_small_cells(emt_engine):
    a = 3.61
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    rng = np.random.default_rng(2)
    for reps in (1, 2, 3):        # 1 and 2: box thinner than 2 rc_list -> several images per pair
        pos = np.concatenate([base + np.array([i, j, k]) * a for i in range(reps)
                              for j in range(reps) for k in range(reps)])
        pos = pos + rng.normal(scale=0.05, size=pos.shape)
        num = np.full(len(pos), 29)
        num[::3] = 46
        cell = np.eye(3) * a * reps
        e = _energy(emt_engine, pos, num, cell, [1, 1, 1])
        ref = oemt.emt_energy(pos, num, cell, [1, 1, 1])
        assert abs(e - ref) <= RTOL * max(1.0, abs(ref)), (reps, e, ref)


def test_s3_energy_sigma_pairs(emt_engine):
    cfg = synthetic.s3_ag_o()
    emt_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    E, pairs = emt_engine.energies(return_pairs=True)
    ref, sigma_ref, npairs = oemt.emt_energy(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc'],
                                             return_details=True)
    assert int(pairs[0]) == npairs
    assert abs(E[0] - ref) <= RTOL * abs(ref)
    e_atom, sigma1 = emt_engine.atom_energies(want_sigma1=True)
    np.testing.assert_allclose(sigma1, sigma_ref, rtol=1e-12, atol=1e-300)
    assert abs(e_atom.sum() - ref) <= RTOL * abs(ref)
    rc_list = emt_engine.cutoff
    gi, gj, gS, gr2 = emt_engine.neighbor_pairs(0, rc_list, mode=2)
    oi, oj, oS, or2 = neighbor_pairs(cfg['positions'], cfg['cell'], cfg['pbc'], rc_list, use_sqrt=True)
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gr2, or2)


def test_unsupported_species_raises(emt_engine):
    with pytest.raises(Exception):
        _energy(emt_engine, np.zeros((1, 3)), [26], np.zeros((3, 3)), [0, 0, 0])


def _check_trials(engine, cfg, tb, tol_scale=None):
    """dE of every trial against the difference of two oracle totals: 1e-10 of |dE| (helpers.de_tol)."""
    E0 = cfast.emt_energy(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc'])
    dE = engine.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
    for t in range(len(dE)):
        p1, n1 = synthetic.apply_trial(cfg, tb, t)
        ref = cfast.emt_energy(p1, n1, cfg['cell'], cfg['pbc']) - E0
        assert abs(dE[t] - ref) <= de_tol(ref, E0), (t, dE[t], ref)
    return dE


def test_trial_delta_e_cluster(emt_engine):
    cfg = synthetic.s3_ag_o()
    emt_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    tb = synthetic.trial_batch(cfg, 30, seed=4, insert_Z=8)
    cfg2 = dict(cfg); cfg2['max_displacement'] = 0.4
    E0 = abs(oemt.emt_energy(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']))
    _check_trials(emt_engine, cfg, tb, max(1.0, E0))


def test_trial_delta_e_periodic_and_molecule(emt_engine):
    # 4x4x4 fcc Cu/Pd (a*4 = 14.44 A > 2 rc_list + spread) with CO moves
    a = 3.61
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    rng = np.random.default_rng(6)
    pos = np.concatenate([base + np.array([i, j, k]) * a for i in range(4) for j in range(4) for k in range(4)])
    pos = pos + rng.normal(scale=0.04, size=pos.shape)
    num = np.full(len(pos), 29, np.int32); num[::5] = 46
    # carve a void and put a CO in it so molecule deletion / displacement have something to act on
    keep = np.linalg.norm(pos - np.array([7.2, 7.2, 7.2]), axis=1) > 3.0
    pos, num = pos[keep], num[keep]
    co = np.array([[7.2, 7.2, 6.7], [7.2, 7.2, 7.83]])
    pos = np.concatenate([pos, co]); num = np.concatenate([num, [6, 8]]).astype(np.int32)
    n = len(pos)
    cell = np.eye(3) * 4 * a
    cfg = dict(positions=pos, numbers=num, cell=cell, pbc=np.ones(3, bool))
    emt_engine.upload([pos], [num], cell.reshape(1, 9), np.ones((1, 3), bool))
    old_off, new_off, old_idx, new_pos, new_Z = [0], [0], [], [], []
    def add(old, newp, newz):
        old_idx.extend(old); new_pos.extend(newp); new_Z.extend(newz)
        old_off.append(len(old_idx)); new_off.append(len(new_pos))
    add([n - 2, n - 1], co + np.array([0.3, -0.2, 0.25]), [6, 8])          # rigid translation
    add([n - 2, n - 1], [], [])                                              # molecule deletion
    add([], np.array([[7.0, 7.5, 6.9], [7.6, 7.1, 7.85]]), [6, 8])          # second molecule inserted
    add([5], [pos[5] + np.array([0.1, 0.2, -0.15])], [int(num[5])])         # metal atom displaced
    add([0], [pos[0] + np.array([-0.3, -0.3, -0.3])], [int(num[0])])        # across the periodic face
    add([17], [], [])                                                        # atom deletion
    add([], [np.array([0.05, 14.3, 7.0])], [8])                              # insertion near a face
    tb = dict(old_off=np.array(old_off, np.int32), new_off=np.array(new_off, np.int32),
              old_idx=np.array(old_idx, np.int32), new_pos=np.array(new_pos, float).reshape(-1, 3),
              new_Z=np.array(new_Z, np.int32))
    E0 = abs(oemt.emt_energy(pos, num, cell, [1, 1, 1]))
    _check_trials(emt_engine, cfg, tb, max(1.0, E0))


def test_thin_box_reports_nan(emt_engine):
    # 2x2x2 Cu: the box is thinner than 2 rc_list, so the incremental path must refuse
    a = 3.61
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    pos = np.concatenate([base + np.array([i, j, k]) * a for i in range(2) for j in range(2) for k in range(2)])
    emt_engine.upload([pos], [np.full(32, 29, np.int32)], (np.eye(3) * 2 * a).reshape(1, 9), np.ones((1, 3), bool))
    dE = emt_engine.trial_delta_e([0, 1], [3], [0, 1], [pos[3] + 0.1], [29])
    assert np.isnan(dE[0])


def test_s5_ten_thousand_atoms_cluster_build(emt_engine):
    """~10.8k-atom CuPd + CO (BASELINE config 4): the cell list is built by an 8-CTA cluster."""
    from oracle import cfast
    cfg = synthetic.s5_cupd_co()
    emt_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    E, pairs = emt_engine.energies(return_pairs=True)
    ref, sigma_ref, npairs = cfast.emt_energy(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc'],
                                              return_details=True)
    assert int(pairs[0]) == npairs
    assert abs(E[0] - ref) <= RTOL * max(1.0, abs(ref))
    _, sigma1 = emt_engine.atom_energies(want_sigma1=True)
    np.testing.assert_allclose(sigma1, sigma_ref, rtol=1e-11, atol=1e-300)
    # a few CO moves against it (rigid translation, deletion, insertion)
    n = len(cfg['positions'])
    co = cfg['positions'][n - 2:]
    tb = dict(old_off=np.array([0, 2, 4, 4], np.int32), old_idx=np.array([n - 2, n - 1, n - 4, n - 3], np.int32),
              new_off=np.array([0, 2, 2, 4], np.int32),
              new_pos=np.concatenate([co + [0.2, -0.1, 0.15], co + [3.0, 2.0, 1.0]]),
              new_Z=np.array([6, 8, 6, 8], np.int32))
    dE = emt_engine.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
    for t in range(3):
        p1, n1 = synthetic.apply_trial(cfg, tb, t)
        r = cfast.emt_energy(p1, n1, cfg['cell'], cfg['pbc']) - ref
        assert abs(dE[t] - r) <= de_tol(r, ref), (t, dE[t], r)
