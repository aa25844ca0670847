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
