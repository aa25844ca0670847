"""CPU: the EMT oracle against the only EMT numbers the reference documents, plus invariants."""
import numpy as np
import pytest

from oracle import cfast, emt

Z0 = np.zeros((3, 3))


def test_known_answers():
    # docs/make_figures.py:334 -- E_EMT(O2)/2 = +0.46 eV (g2 O2, d = 1.245956 A)
    e = emt.emt_energy([[0, 0, 0], [0, 0, 1.245956]], [8, 8], Z0, [0, 0, 0])
    assert round(e / 2, 2) == 0.46
    # ASE's N2 / Cu(111) tutorial prints 0.440343572 eV for N2 at d = 1.10 A
    e = emt.emt_energy([[0, 0, 0], [0, 0, 1.10]], [7, 7], Z0, [0, 0, 0])
    assert abs(e - 0.440343572) < 5e-9
    # isolated atoms: log(0) branch -> -E0 each
    assert emt.emt_energy([[0, 0, 0]], [29], Z0, [0, 0, 0]) == 3.51
    assert emt.emt_energy([[0, 0, 0], [0, 0, 50.0]], [47, 8], Z0, [0, 0, 0]) == pytest.approx(2.96 + 4.60)


def test_cutoffs():
    rc, rc_list, acut = emt.cutoffs()
    assert rc == pytest.approx(5.376798324905922, rel=1e-15)
    assert rc_list == rc + 0.5
    assert acut == pytest.approx(23.85845476790713, rel=1e-14)


def test_bulk_copper_near_zero_at_lattice_constant():
    a = 3.61
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    pos = np.concatenate([base + np.array([i, j, k]) * a for i in range(2) for j in range(2) for k in range(2)])
    e = emt.emt_energy(pos, [29] * 32, np.eye(3) * 2 * a, [1, 1, 1]) / 32
    assert abs(e) < 0.01          # EMT's own energy zero is the bulk metal
    assert cfast.emt_energy(pos, [29] * 32, np.eye(3) * 2 * a, [1, 1, 1]) / 32 == pytest.approx(e, abs=1e-12)


def test_invariances_and_c_port():
    rng = np.random.default_rng(0)
    pos = rng.normal(size=(25, 3)) * 2.2
    num = rng.choice([29, 46, 8, 6, 47], size=25)
    e = emt.emt_energy(pos, num, Z0, [0, 0, 0])
    # translation, rotation, permutation
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    perm = rng.permutation(25)
    assert emt.emt_energy(pos + [3.0, -1.0, 2.0], num, Z0, [0, 0, 0]) == pytest.approx(e, rel=1e-12)
    assert emt.emt_energy(pos @ q.T, num, Z0, [0, 0, 0]) == pytest.approx(e, rel=1e-12)
    assert emt.emt_energy(pos[perm], num[perm], Z0, [0, 0, 0]) == pytest.approx(e, rel=1e-12)
    ec, sig, npairs = cfast.emt_energy(pos, num, Z0, [0, 0, 0], return_details=True)
    en, sign, npn = emt.emt_energy(pos, num, Z0, [0, 0, 0], return_details=True)
    assert ec == pytest.approx(en, rel=1e-12) and npairs == npn
    np.testing.assert_allclose(sig, sign, rtol=1e-12)
    # periodic: a lattice translation of one atom changes nothing
    cell = np.diag([12.0, 13.0, 14.0])
    p2 = rng.random((30, 3)) @ cell
    n2 = rng.choice([29, 46], size=30)
    e2 = emt.emt_energy(p2, n2, cell, [1, 1, 1])
    p3 = p2.copy(); p3[4] += [12.0, -13.0, 28.0]
    assert emt.emt_energy(p3, n2, cell, [1, 1, 1]) == pytest.approx(e2, rel=1e-12)
