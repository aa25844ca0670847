"""CPU: the LJ oracle against the reference's golden values (runs without a GPU)."""
import numpy as np
import pytest

from oracle import cfast, lj
from oracle.neighbors import neighbor_pairs


def test_two_body_probe_values():
    # benchmark/lammps_gcmc_parity.py:285-324 -- ASE LennardJones == LAMMPS lj/cut 3.0 + shift yes
    # to 1e-8 at r in {0.9, 1.0, 1.5, 2.5, 2.99}: E = 4(r^-12 - r^-6) - 4(3^-12 - 3^-6)
    cell = np.diag([20.0, 20.0, 20.0])
    for r in [0.9, 1.0, 1.5, 2.5, 2.99]:
        pos = np.array([[5, 5, 5], [5, 5, 5 + r]], float)
        exact = 4 * (r ** -12 - r ** -6) - 4 * (3.0 ** -12 - 3.0 ** -6)
        for fn in (lj.lj_energy, cfast.lj_energy):
            e = fn(pos, cell, [1, 1, 1], 1.0, 1.0, 3.0)
            assert abs(e - exact) <= 1e-8 * max(1.0, abs(exact))


def test_matches_reference_ljcalc_on_its_own_check_configs():
    # benchmark/lammps_gcmc_parity.py:190-203: 5 x 40 atoms, default_rng(99), L = 9; the
    # reference asserts |fast - ASE| <= 1e-9 max(1, |E|)
    rng = np.random.default_rng(99)
    L = 9.0
    cell = np.diag([L] * 3)
    for _ in range(5):
        pos = rng.uniform(0, L, size=(40, 3))
        fast = lj.LJCalcRestated(L, 3.0).energy(pos)
        full = lj.lj_energy(pos, cell, [1, 1, 1], 1.0, 1.0, 3.0)
        port = lj.AseStyleLJ(1.0, 1.0, 3.0).energy(pos, cell, [1, 1, 1])
        c = cfast.lj_energy(pos, cell, [1, 1, 1], 1.0, 1.0, 3.0)
        for e in (full, port, c):
            assert abs(e - fast) <= 1e-9 * max(1.0, abs(fast))


def test_ase_defaults_and_empty():
    assert lj.lj_energy(np.zeros((0, 3)), np.eye(3) * 5, [1, 1, 1]) == 0.0
    assert lj.lj_energy([[0, 0, 0]], np.zeros((3, 3)), [0, 0, 0]) == 0.0
    # rc defaults to 3 sigma
    pos = np.array([[0, 0, 0], [0, 0, 5.9]], float)
    assert lj.lj_energy(pos, np.zeros((3, 3)), [0, 0, 0], sigma=2.0) != 0.0
    assert lj.lj_energy(pos, np.zeros((3, 3)), [0, 0, 0], sigma=1.9) == 0.0


@pytest.mark.parametrize('pbc', [[1, 1, 1], [1, 1, 0], [0, 0, 0]])
def test_c_oracle_pairs_are_bit_identical_to_numpy(pbc):
    rng = np.random.default_rng(4)
    cell = np.array([[9.0, 0, 0], [2.5, 8.5, 0], [1.0, -2.0, 10.0]])
    pos = rng.random((70, 3)) @ cell + rng.integers(-2, 3, (70, 3)) @ cell
    for mode, kw in [(0, {}), (1, {'inclusive': True}), (2, {'use_sqrt': True})]:
        a = neighbor_pairs(pos, cell, pbc, 2.75, **kw)
        b = cfast.neighbor_pairs(pos, cell, pbc, 2.75, mode)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_thin_box_needs_several_images():
    # L < rc: an atom interacts with its own periodic images (ASE's list contains them)
    cell = np.diag([2.0, 30.0, 30.0])
    e = lj.lj_energy([[0.3, 1, 1]], cell, [1, 0, 0], 1.0, 1.0, 3.0)
    expect = 0.5 * 2 * lj.lj_pair_energy(np.array([4.0]), 1.0, 1.0, 3.0)[0]
    assert e == pytest.approx(expect, rel=1e-14)
    assert cfast.lj_energy([[0.3, 1, 1]], cell, [1, 0, 0], 1.0, 1.0, 3.0) == pytest.approx(expect, rel=1e-14)


def test_delta_e_is_difference_of_totals():
    rng = np.random.default_rng(1)
    L = 8.0
    cell = np.diag([L] * 3)
    pos = rng.uniform(0, L, (30, 3))
    new = pos.copy()
    new[7] += [0.2, -0.1, 0.15]
    d = lj.lj_delta_e(pos, new, cell, [1, 1, 1])
    assert d == lj.lj_energy(new, cell, [1, 1, 1]) - lj.lj_energy(pos, cell, [1, 1, 1])


def test_c_oracle_bin_pruning_changes_no_bit():
    """The plain-C oracle skips atoms whose bin cannot hold an image within the cutoff; the visited
    (i, j, S) sequence must be the plain O(N^2 * images) loop's, so every result is bit-identical."""
    from oracle import cfast
    rng = np.random.default_rng(11)
    tri = np.array([[11.0, 0.0, 0.0], [3.0, 9.5, 0.0], [-2.0, 1.5, 10.0]])
    cases = [
        (rng.random((400, 3)) @ (np.eye(3) * 12.0) * 1.3 - 2.0, np.eye(3) * 12.0, [1, 1, 1], 3.0),      # unwrapped atoms
        (rng.random((300, 3)) @ tri, tri, [1, 1, 1], 2.5),                                             # triclinic
        (rng.random((300, 3)) * [9.0, 9.0, 30.0], np.diag([9.0, 9.0, 40.0]), [1, 1, 0], 3.0),           # slab
        (rng.normal(size=(250, 3)) * 6.0, np.zeros((3, 3)), [0, 0, 0], 3.0),                            # cluster
        (rng.random((60, 3)) * 4.0, np.eye(3) * 4.0, [1, 1, 1], 3.0),                                   # box thinner than 2 rc
    ]
    for pos, cell, pbc, rc in cases:
        num = rng.choice([29, 46, 8, 6], size=len(pos))
        cfast.set_bruteforce(False)
        fast = (cfast.lj_energy(pos, cell, pbc, 1.0, 1.0, rc, return_pairs=True),
                cfast.emt_energy(pos, num, cell, pbc, return_details=True),
                cfast.neighbor_pairs(pos, cell, pbc, rc, mode=1))
        cfast.set_bruteforce(True)
        try:
            slow = (cfast.lj_energy(pos, cell, pbc, 1.0, 1.0, rc, return_pairs=True),
                    cfast.emt_energy(pos, num, cell, pbc, return_details=True),
                    cfast.neighbor_pairs(pos, cell, pbc, rc, mode=1))
        finally:
            cfast.set_bruteforce(False)
        assert fast[0] == slow[0] and fast[0][1] > 0
        assert fast[1][0] == slow[1][0] and np.array_equal(fast[1][1], slow[1][1]) and fast[1][2] == slow[1][2]
        for a, b in zip(fast[2], slow[2]):
            assert np.array_equal(a, b)
