"""GPU parity: Lennard-Jones energies, neighbour pairs and trial delta-E vs the oracle.

Bars (BASELINE.json north_star): pair sets bit-exact; energies and delta-E within
1e-10 relative in fp64.
"""
import numpy as np
import pytest

from helpers import de_tol
from mcpy_b200.utils import synthetic
from oracle import cfast
from oracle import lj as olj
from oracle.neighbors import neighbor_pairs

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _rel(a, b, scale=None):
    scale = max(1.0, abs(b)) if scale is None else scale
    return abs(a - b) / scale


def _energy(engine, pos, cell, pbc, numbers=None):
    numbers = np.ones(len(pos), np.int32) if numbers is None else numbers
    return engine.potential_energies([pos], [numbers], np.asarray(cell, float).reshape(1, 9),
                                     np.asarray(pbc, bool).reshape(1, 3))[0]


def test_two_body_probe_matches_lammps_convention(lj_engine):
    # benchmark/lammps_gcmc_parity.py:285-324: E(r) = 4(r^-12 - r^-6) - 4(3^-12 - 3^-6)
    lj_engine.set_lj(1.0, 1.0, 3.0)
    for r in [0.9, 1.0, 1.5, 2.5, 2.99]:
        pos = np.array([[5, 5, 5], [5, 5, 5 + r]], float)
        e = _energy(lj_engine, pos, np.diag([20.0] * 3), [1, 1, 1])
        exact = 4 * (r ** -12 - r ** -6) - 4 * (3.0 ** -12 - 3.0 ** -6)
        assert abs(e - exact) <= 1e-8 * max(1.0, abs(exact))
    # beyond the cutoff and exactly at it
    assert _energy(lj_engine, np.array([[5, 5, 5], [5, 5, 8.5]], float), np.diag([20.0] * 3), [1, 1, 1]) == 0.0


def test_ljcalc_configs_match(lj_engine):
    # benchmark/lammps_gcmc_parity.py:190-203: 5 x 40 atoms, default_rng(99), L = 9, rc = 3
    lj_engine.set_lj(1.0, 1.0, 3.0)
    rng = np.random.default_rng(99)
    L = 9.0
    for _ in range(5):
        pos = rng.uniform(0, L, size=(40, 3))
        e = _energy(lj_engine, pos, np.diag([L] * 3), [1, 1, 1])
        ref = olj.lj_energy(pos, np.diag([L] * 3), [1, 1, 1], 1.0, 1.0, 3.0)
        fast = olj.LJCalcRestated(L, 3.0).energy(pos)
        assert _rel(e, ref) <= RTOL
        assert _rel(e, fast) <= 1e-9


@pytest.mark.parametrize('builder,kw', [(synthetic.s1_argon, {}), (synthetic.s2_lj, {}),
                                        (synthetic.s2_lj, {'rc': 2 ** (1 / 6)})])
def test_synthetic_energy_and_pairs(lj_engine, builder, kw):
    cfg = builder(**kw)
    lj_engine.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
    lj_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    E, pairs = lj_engine.energies(return_pairs=True)
    ref, ref_pairs = cfast.lj_energy(cfg['positions'], cfg['cell'], cfg['pbc'], cfg['sigma'],
                                     cfg['epsilon'], cfg['rc'], return_pairs=True)
    assert int(pairs[0]) == ref_pairs
    assert _rel(E[0], ref) <= RTOL
    # pair set bit-exact, including the fp64 r2 of every image
    gi, gj, gS, gr2 = lj_engine.neighbor_pairs(0, cfg['rc'], mode=1)
    oi, oj, oS, or2 = cfast.neighbor_pairs(cfg['positions'], cfg['cell'], cfg['pbc'], cfg['rc'], mode=1)
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gS, oS)
    assert np.array_equal(gr2, or2)


def _random_cfg(rng, n, cell, pbc, spread=1.0):
    cell = np.asarray(cell, float)
    frac = rng.random((n, 3)) * spread + (1 - spread) / 2
    return frac @ cell if np.any(cell) else rng.normal(size=(n, 3)) * 3


@pytest.mark.parametrize('case', ['triclinic', 'slab', 'cluster', 'thin', 'unwrapped', 'tiny'])
def test_geometries(lj_engine, case):
    rng = np.random.default_rng(11)
    lj_engine.set_lj(0.7, 1.1, 2.75)
    if case == 'triclinic':
        cell = np.array([[9.0, 0, 0], [2.5, 8.5, 0], [1.0, -2.0, 10.0]]); pbc = [1, 1, 1]
        pos = _random_cfg(rng, 150, cell, pbc)
    elif case == 'slab':
        cell = np.diag([8.0, 9.0, 30.0]); pbc = [1, 1, 0]
        pos = _random_cfg(rng, 120, cell, pbc); pos[:, 2] = rng.uniform(-4, 9, 120)
    elif case == 'cluster':
        cell = np.zeros((3, 3)); pbc = [0, 0, 0]
        pos = rng.normal(size=(90, 3)) * 2.5
    elif case == 'thin':       # box thinner than the cutoff along y and z: several images
        cell = np.diag([7.0, 2.1, 1.3]); pbc = [1, 1, 1]
        pos = _random_cfg(rng, 12, cell, pbc)
    elif case == 'unwrapped':  # atoms several boxes away from the primary cell
        cell = np.diag([7.5, 8.0, 6.5]); pbc = [1, 1, 1]
        pos = _random_cfg(rng, 100, cell, pbc) + rng.integers(-3, 4, size=(100, 3)) * np.diag(cell)
    else:
        cell = np.diag([5.0, 5.0, 5.0]); pbc = [1, 1, 1]
        pos = np.array([[1.0, 2.0, 3.0]])
    # keep atoms from overlapping absurdly
    e = _energy(lj_engine, pos, cell, pbc)
    ref = olj.lj_energy(pos, cell, pbc, 1.1, 0.7, 2.75)
    assert _rel(e, ref, max(1.0, abs(ref))) <= RTOL
    gi, gj, gS, gr2 = lj_engine.neighbor_pairs(0, 2.75, mode=1)
    oi, oj, oS, or2 = neighbor_pairs(pos, cell, pbc, 2.75, inclusive=True)
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gS, oS)
    assert np.array_equal(gr2, or2)


def test_cutoff_boundary_is_inclusive_and_bit_exact(lj_engine):
    # pairs placed exactly at r == rc and one ulp either side (ASE keeps r2 <= rc^2)
    rc = 2.5
    lj_engine.set_lj(1.0, 1.0, rc)
    base = np.array([3.0, 3.0, 3.0])
    pos = [base]
    for k, d in enumerate([rc, np.nextafter(rc, 0), np.nextafter(rc, 10)]):
        pos.append(base + np.array([0.0, 0.0, 0.0]) + np.eye(3)[k] * d)
    pos = np.array(pos)
    cell = np.diag([40.0] * 3)
    lj_engine.upload([pos], [np.ones(4, np.int32)], cell.reshape(1, 9), np.ones((1, 3), bool))
    for mode, kw in [(0, {}), (1, {'inclusive': True}), (2, {'use_sqrt': True})]:
        gi, gj, gS, gr2 = lj_engine.neighbor_pairs(0, rc, mode=mode)
        oi, oj, oS, or2 = neighbor_pairs(pos, cell, [1, 1, 1], rc, **kw)
        assert np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gr2, or2)


def test_empty_and_ragged_batch(lj_engine):
    lj_engine.set_lj(1.0, 1.0, 3.0)
    rng = np.random.default_rng(3)
    L = 9.0
    sizes = [0, 1, 37, 2, 120, 0, 64]
    poss = [rng.uniform(0, L, size=(n, 3)) for n in sizes]
    nums = [np.ones(n, np.int32) for n in sizes]
    cells = np.tile(np.diag([L] * 3).reshape(1, 9), (len(sizes), 1))
    cells[3] = np.diag([12.0, 9.0, 10.0]).reshape(9)
    pbcs = np.ones((len(sizes), 3), bool)
    pbcs[4] = [True, False, True]
    E = lj_engine.potential_energies(poss, nums, cells, pbcs)
    for r, n in enumerate(sizes):
        ref = olj.lj_energy(poss[r], cells[r].reshape(3, 3), pbcs[r], 1.0, 1.0, 3.0)
        assert _rel(E[r], ref) <= RTOL
    assert lj_engine.potential_energies([], [], np.zeros((0, 9)), np.zeros((0, 3), bool)).shape == (0,)


@pytest.mark.parametrize('builder,kw,ntr', [(synthetic.s1_argon, {}, 48), (synthetic.s2_lj, {}, 24)])
def test_trial_delta_e_matches_two_full_energies(lj_engine, builder, kw, ntr):
    cfg = builder(**kw)
    lj_engine.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
    lj_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    tb = synthetic.trial_batch(cfg, ntr, seed=5)
    for group in ('small', 'large'):
        if group == 'large':      # the same trials replicated into a much larger batch
            reps = 600 // ntr + 1
            big = {k: v for k, v in tb.items()}
            big['old_off'] = np.concatenate([[0], np.cumsum(np.tile(np.diff(tb['old_off']), reps))]).astype(np.int32)
            big['new_off'] = np.concatenate([[0], np.cumsum(np.tile(np.diff(tb['new_off']), reps))]).astype(np.int32)
            big['old_idx'] = np.tile(tb['old_idx'], reps)
            big['new_pos'] = np.tile(tb['new_pos'], (reps, 1))
            big['new_Z'] = np.tile(tb['new_Z'], reps)
            use = big
        else:
            use = tb
        dE, pairs = lj_engine.trial_delta_e(use['old_off'], use['old_idx'], use['new_off'],
                                            use['new_pos'], use['new_Z'], return_pairs=True)
        if group == 'small':
            E0 = cfast.lj_energy(cfg['positions'], cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc'])
            first = dE.copy()
            for t in range(ntr):
                p1, _ = synthetic.apply_trial(cfg, tb, t)
                ref = cfast.lj_energy(p1, cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc']) - E0
                assert abs(dE[t] - ref) <= de_tol(ref, E0), (t, tb['kind'][t], dE[t], ref)
        else:
            assert np.array_equal(dE[:ntr], first)          # a trial's bits do not depend on its batch
            assert np.array_equal(dE[ntr:2 * ntr], first)


def test_rigid_molecule_trials_and_periodic_self_images(lj_engine):
    # k = 2 and k = 3 moved sets, including a box small enough that moved atoms see
    # each other's periodic images
    rng = np.random.default_rng(21)
    lj_engine.set_lj(1.0, 1.0, 2.5)
    for L, n in [(9.0, 80), (5.5, 30)]:
        cell = np.diag([L] * 3)
        pos = rng.uniform(0, L, size=(n, 3))
        cfg = dict(positions=pos, numbers=np.ones(n, np.int32), cell=cell, pbc=np.ones(3, bool))
        lj_engine.upload([pos], [cfg['numbers']], cell.reshape(1, 9), np.ones((1, 3), bool))
        old_off, new_off, old_idx, new_pos = [0], [0], [], []
        for k in (2, 3, 2, 3):
            idx = rng.choice(n, size=k, replace=False)
            old_idx += list(idx)
            new_pos += list(pos[idx] + rng.uniform(-0.8, 0.8, size=(1, 3)) + rng.uniform(-0.2, 0.2, size=(k, 3)))
            old_off.append(len(old_idx)); new_off.append(len(new_pos))
        # molecule insertion (k = 2) and deletion (k = 3)
        new_pos += list(rng.uniform(0, L, size=(1, 3)) + np.array([[0, 0, 0], [0, 0, 1.1]]))
        old_off.append(len(old_idx)); new_off.append(len(new_pos))
        old_idx += list(rng.choice(n, size=3, replace=False))
        old_off.append(len(old_idx)); new_off.append(len(new_pos))
        tb = dict(old_off=np.array(old_off, np.int32), new_off=np.array(new_off, np.int32),
                  old_idx=np.array(old_idx, np.int32), new_pos=np.array(new_pos),
                  new_Z=np.ones(len(new_pos), np.int32))
        dE = lj_engine.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
        E0 = cfast.lj_energy(pos, cell, [1, 1, 1], 1.0, 1.0, 2.5)
        for t in range(len(dE)):
            p1, _ = synthetic.apply_trial(cfg, tb, t)
            ref = cfast.lj_energy(p1, cell, [1, 1, 1], 1.0, 1.0, 2.5) - E0
            assert abs(dE[t] - ref) <= de_tol(ref, E0), (L, t, dE[t], ref)


def test_trials_on_a_replica_batch(lj_engine):
    rng = np.random.default_rng(8)
    lj_engine.set_lj(1.0, 1.0, 3.0)
    L = 10.0
    sizes = [50, 0, 75, 31]
    poss = [rng.uniform(0, L, size=(n, 3)) for n in sizes]
    nums = [np.ones(n, np.int32) for n in sizes]
    cells = np.tile(np.diag([L] * 3).reshape(1, 9), (4, 1))
    pbcs = np.ones((4, 3), bool)
    lj_engine.upload(poss, nums, cells, pbcs)
    rep = np.array([0, 2, 3, 1, 2, 0], np.int32)
    old_off = np.array([0, 1, 1, 2, 2, 3, 3], np.int32)
    old_idx = np.array([7, 30, 74], np.int32)
    new_off = np.array([0, 1, 2, 2, 3, 3, 4], np.int32)
    new_pos = rng.uniform(0, L, size=(4, 3))
    dE = lj_engine.trial_delta_e(old_off, old_idx, new_off, new_pos, np.ones(4, np.int32), rep=rep)
    for t in range(6):
        r = rep[t]
        cfg = dict(positions=poss[r], numbers=nums[r])
        tb = dict(old_off=old_off, old_idx=old_idx, new_off=new_off, new_pos=new_pos, new_Z=np.ones(4, np.int32))
        p1, _ = synthetic.apply_trial(cfg, tb, t)
        c = cells[r].reshape(3, 3)
        e_old = cfast.lj_energy(poss[r], c, [1, 1, 1])
        ref = cfast.lj_energy(p1, c, [1, 1, 1]) - e_old
        assert abs(dE[t] - ref) <= de_tol(ref, e_old)


def test_errors(lj_engine):
    lj_engine.set_lj(1.0, 1.0, 3.0)
    pos = np.zeros((2, 3)); pos[1, 0] = 1.2
    with pytest.raises(Exception):          # periodic axis with zero lattice vector
        _energy(lj_engine, pos, np.zeros((3, 3)), [1, 1, 1])
    lj_engine.upload([pos], [np.ones(2, np.int32)], np.diag([9.0] * 3).reshape(1, 9), np.ones((1, 3), bool))
    with pytest.raises(ValueError):         # atom index out of range
        lj_engine.trial_delta_e([0, 1], [5], [0, 0], np.zeros((0, 3)), [])
    with pytest.raises(ValueError):
        lj_engine.trial_delta_e([0, 0], [], [0, 1], np.zeros((1, 3)), [1], rep=[3])


def test_large_periodic_box_cluster_build(lj_engine):
    """N >= 4096 switches the cell-list build to an 8-CTA thread-block cluster."""
    cfg = synthetic.s2_lj(n=6000, rc=2.5, rho=0.7, seed=9)
    lj_engine.set_lj(1.0, 1.0, 2.5)
    lj_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    E, pairs = lj_engine.energies(return_pairs=True)
    ref, ref_pairs = cfast.lj_energy(cfg['positions'], cfg['cell'], cfg['pbc'], 1.0, 1.0, 2.5, return_pairs=True)
    assert int(pairs[0]) == ref_pairs
    assert _rel(E[0], ref) <= RTOL
    gi, gj, gS, gr2 = lj_engine.neighbor_pairs(0, 2.5, mode=1)
    oi, oj, oS, or2 = cfast.neighbor_pairs(cfg['positions'], cfg['cell'], cfg['pbc'], 2.5, mode=1)
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gS, oS) and np.array_equal(gr2, or2)
    # two big replicas and a small one in one batch
    small = synthetic.s2_lj(n=300, rc=2.5, rho=0.5, seed=3)
    Eb = lj_engine.potential_energies([cfg['positions'], small['positions'], cfg['positions'][:5000]],
                                      [cfg['numbers'], small['numbers'], cfg['numbers'][:5000]],
                                      np.stack([cfg['cell'].reshape(9), small['cell'].reshape(9), cfg['cell'].reshape(9)]),
                                      np.ones((3, 3), bool))
    assert Eb[0] == E[0]
    assert _rel(Eb[1], cfast.lj_energy(small['positions'], small['cell'], small['pbc'], 1.0, 1.0, 2.5)) <= RTOL
    assert _rel(Eb[2], cfast.lj_energy(cfg['positions'][:5000], cfg['cell'], cfg['pbc'], 1.0, 1.0, 2.5)) <= RTOL


def test_stencil_lists_pinned_arrays_and_repeated_batches_give_the_same_bits(lj_engine):
    """The first batch against a freshly uploaded configuration stages every stencil from the cell list, the
    following ones read the per-configuration stencil lists (rows_block.cuh, expand_stencils_kernel); page-locked
    caller arrays take the direct DMA path of mcpyb_trial_delta_e_host.  None of that may change a bit, and a
    sample of the batch is checked against two full oracle energies."""
    from mcpy_b200 import pinned_copy, pinned_empty
    cfg = synthetic.s2_lj()
    lj_engine.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
    lj_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    T = 20_000                                    # > 64 KB of arguments: the large-batch path
    tb = synthetic.trial_batch(cfg, T, seed=11)
    args = [tb[k] for k in ('old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')]
    first = lj_engine.trial_delta_e(*args)                        # no lists yet
    second = lj_engine.trial_delta_e(*args)                       # lists built by this call
    third = lj_engine.trial_delta_e(*args)
    assert np.array_equal(first, second) and np.array_equal(first, third)
    out = pinned_empty(T, np.float64)
    out[:] = np.nan
    res = lj_engine.trial_delta_e(*[pinned_copy(a) for a in args], out=out)
    assert res is out and np.array_equal(out, first)
    # a fresh upload drops the lists; the answer does not move
    lj_engine.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
    assert np.array_equal(lj_engine.trial_delta_e(*args), first)
    E0 = cfast.lj_energy(cfg['positions'], cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc'])
    for t in range(0, T, T // 16):
        p1, _ = synthetic.apply_trial(cfg, tb, t)
        ref = cfast.lj_energy(p1, cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc']) - E0
        assert abs(first[t] - ref) <= de_tol(ref, E0), (t, tb['kind'][t], first[t], ref)
    with pytest.raises(ValueError):
        lj_engine.trial_delta_e(*args, out=np.empty(T - 1))


def test_full_size_batch_properties(lj_engine):
    """BASELINE's headline size (65 536 proposals against the 2 000-atom S2 box), where two oracle energies per
    trial are out of reach: size-independent properties instead.
      * order: permuting the batch permutes the answers, bit for bit;
      * identity: displacing an atom onto its own position costs exactly 0.0;
      * round trip: dE(insert p | C) = -dE(delete it | C + p), to 1e-10;
      * composition: dE(move i -> b | C) = dE(delete i | C) + dE(insert b | C - i), to 1e-10."""
    cfg = synthetic.s2_lj()
    pos, num, cell, pbc = cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']
    lj_engine.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
    up = lambda p: lj_engine.upload([p], [np.ones(len(p), np.int32)], cell.reshape(1, 9), pbc.reshape(1, 3))
    up(pos)
    T = 65536
    tb = synthetic.trial_batch(cfg, T, seed=7)
    keys = ('old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')
    dE = lj_engine.trial_delta_e(*[tb[k] for k in keys])
    assert np.all(np.isfinite(dE))
    # ---- order
    rng = np.random.default_rng(3)
    perm = rng.permutation(T)
    no, nn = np.diff(tb['old_off']), np.diff(tb['new_off'])
    old_rows = np.concatenate([np.arange(tb['old_off'][t], tb['old_off'][t + 1]) for t in perm]).astype(int)
    new_rows = np.concatenate([np.arange(tb['new_off'][t], tb['new_off'][t + 1]) for t in perm]).astype(int)
    pb = dict(old_off=np.concatenate([[0], np.cumsum(no[perm])]).astype(np.int32), old_idx=tb['old_idx'][old_rows],
              new_off=np.concatenate([[0], np.cumsum(nn[perm])]).astype(np.int32), new_pos=tb['new_pos'][new_rows],
              new_Z=tb['new_Z'][new_rows])
    assert np.array_equal(lj_engine.trial_delta_e(*[pb[k] for k in keys]), dE[perm])
    # ---- identity
    n = len(pos)
    idx = np.arange(n, dtype=np.int32)
    off = np.arange(n + 1, dtype=np.int32)
    same = lj_engine.trial_delta_e(off, idx, off, pos[idx], np.ones(n, np.int32))
    assert np.all(same == 0.0)
    # ---- round trip and composition on a sample of the batch
    ins = [t for t in range(T) if no[t] == 0 and nn[t] == 1][:24]
    mov = [t for t in range(T) if no[t] == 1 and nn[t] == 1][:24]
    one = np.array([0, 1], np.int32)
    zero = np.array([0, 0], np.int32)
    for t in ins:
        p = tb['new_pos'][tb['new_off'][t]]
        up(np.vstack([pos, p]))
        back = lj_engine.trial_delta_e(one, np.array([n], np.int32), zero, np.zeros((0, 3)), np.zeros(0, np.int32))[0]
        assert abs(dE[t] + back) <= RTOL * max(1.0, abs(dE[t])), (t, dE[t], back)
    up(pos)
    for t in mov:
        i = int(tb['old_idx'][tb['old_off'][t]])
        b = tb['new_pos'][tb['new_off'][t]]
        d_del = lj_engine.trial_delta_e(one, np.array([i], np.int32), zero, np.zeros((0, 3)), np.zeros(0, np.int32))[0]
        up(np.delete(pos, i, axis=0))
        d_ins = lj_engine.trial_delta_e(zero, np.zeros(0, np.int32), one, b.reshape(1, 3), np.ones(1, np.int32))[0]
        up(pos)
        assert abs(dE[t] - (d_del + d_ins)) <= RTOL * max(1.0, abs(d_del), abs(d_ins)), (t, dE[t], d_del, d_ins)


@pytest.mark.parametrize('name,n_trials', [('s1_argon', 3000), ('s2_lj', 6000)])
def test_fp32_pair_math_mode_within_1e5(name, n_trials):
    """north_star's second precision: fp32 pair math (r2, reciprocal, c6 on the FP32 pipe, fp64 accumulation) in the
    batched row sums.  dE within 1e-5 relative of the oracle (tolerance: 1e-5 max(1, |dE|)), the in-cutoff pair
    counts -- membership is still decided bit-exactly -- identical to the fp64 mode's."""
    from mcpy_b200 import Engine
    cfg = getattr(synthetic, name)()
    tb = synthetic.trial_batch(cfg, n_trials, seed=21)
    args = (tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
    out = {}
    for prec in ('fp64', 'fp32'):
        eng = Engine(0)
        eng.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
        eng.set_precision(prec)
        eng.upload([cfg['positions']], [cfg['numbers']], cfg['cell'].reshape(1, 9), cfg['pbc'].reshape(1, 3))
        out[prec] = eng.trial_delta_e(*args, return_pairs=True)
        eng.close()
    d64, p64 = out['fp64']
    d32, p32 = out['fp32']
    assert np.array_equal(p64, p32)
    assert not np.array_equal(d64, d32)                                   # the mode really ran in another arithmetic
    E0 = cfast.lj_energy(cfg['positions'], cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc'])
    worst = 0.0
    for t in range(0, n_trials, n_trials // 60):
        p1, _ = synthetic.apply_trial(cfg, tb, t)
        ref = cfast.lj_energy(p1, cfg['cell'], cfg['pbc'], cfg['sigma'], cfg['epsilon'], cfg['rc']) - E0
        worst = max(worst, abs(d32[t] - ref) / max(1.0, abs(ref)))
        assert abs(d32[t] - ref) <= 1e-5 * max(1.0, abs(ref)) + 1e-13 * abs(E0), (t, d32[t], ref)
    rel = np.abs(d32 - d64) / np.maximum(1.0, np.abs(d64))
    assert rel.max() <= 1e-5, rel.max()
    print(f'fp32 pair math on {name}: max |dE32 - dE64| / max(1, |dE|) = {rel.max():.2e}, vs oracle {worst:.2e}')


def test_direct_placement_overflow_and_scan_pipelines_give_the_same_bits(monkeypatch):
    """A dense batch takes the direct placement (site_kernels.cuh: the rank a site draws from its bin's histogram is
    its slot, no scan, no scatter).  The same batch through (a) 8 slots per bin, which sends most sites through the
    overflow kernel, (b) the scan + scatter pipeline and (c) a replica batch with ragged boxes must not move a bit,
    and a sample is checked against the oracle."""
    from mcpy_b200 import Engine
    cfgs = [synthetic.s2_lj(2000, 3.0, rho, seed=40 + k) for k, rho in enumerate((0.6, 0.8442, 0.7))]
    rng = np.random.default_rng(12)
    parts = [synthetic.trial_batch(c, 12_000, seed=50 + k) for k, c in enumerate(cfgs)]
    rep = np.concatenate([np.full(12_000, k, np.int32) for k in range(3)])
    old_off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(p['old_off']) for p in parts]))]).astype(np.int32)
    new_off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(p['new_off']) for p in parts]))]).astype(np.int32)
    old_idx = np.concatenate([p['old_idx'] for p in parts])
    new_pos = np.concatenate([p['new_pos'] for p in parts])
    new_Z = np.concatenate([p['new_Z'] for p in parts])
    perm = rng.permutation(len(rep))              # interleave the replicas' trials
    def take(off, rows, order):
        seg = [np.arange(off[t], off[t + 1]) for t in order]
        idx = np.concatenate(seg).astype(int)
        return np.concatenate([[0], np.cumsum([len(s) for s in seg])]).astype(np.int32), rows[idx]
    oo, oi = take(old_off, old_idx, perm)
    no, npos = take(new_off, new_pos, perm)
    _, nz = take(new_off, new_Z, perm)
    res = {}
    for label, env in (('direct', {}), ('overflow', {'MCPYB_SLOT_CAP': '8'}), ('scan', {'MCPYB_DIRECT': '0'})):
        for k in ('MCPYB_SLOT_CAP', 'MCPYB_DIRECT'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = Engine(0)
        eng.set_lj(1.0, 1.0, 3.0)
        eng.upload([c['positions'] for c in cfgs], [c['numbers'] for c in cfgs],
                   np.stack([c['cell'].reshape(9) for c in cfgs]), np.stack([c['pbc'] for c in cfgs]))
        res[label] = [eng.trial_delta_e(oo, oi, no, npos, nz, rep=rep[perm]).copy() for _ in range(2)]
        eng.close()
    base = res['direct'][0]
    assert np.all(np.isfinite(base))
    for label, runs in res.items():
        for r in runs:
            assert np.array_equal(r, base), label
    for t in range(0, len(perm), len(perm) // 12):
        k, tt = int(rep[perm[t]]), int(perm[t]) % 12_000
        c = cfgs[k]
        p1, _ = synthetic.apply_trial(c, parts[k], tt)
        E0 = cfast.lj_energy(c['positions'], c['cell'], c['pbc'], 1.0, 1.0, 3.0)
        ref = cfast.lj_energy(p1, c['cell'], c['pbc'], 1.0, 1.0, 3.0) - E0
        assert abs(base[t] - ref) <= de_tol(ref, E0), (t, base[t], ref)
