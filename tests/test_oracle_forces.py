"""CPU: the oracle's analytic forces against central finite differences of the oracle's energies.

The reference's tests assert no force value, so the analytic restatements of ASE's
``LennardJones`` / ``EMT`` forces are pinned by the definition F = -dE/dx of the already pinned
energies.  Tolerance: central differences with h = 1e-5 A are good to ~1e-9 relative here."""
import numpy as np
import pytest

from oracle import emt as oemt, lj as olj


def _fd(energy, pos, h=1e-5):
    F = np.zeros_like(pos)
    for a in range(len(pos)):
        for c in range(3):
            p = pos.copy(); p[a, c] += h
            m = pos.copy(); m[a, c] -= h
            F[a, c] = -(energy(p) - energy(m)) / (2 * h)
    return F


@pytest.mark.parametrize('pbc', [(True, True, True), (False, False, False), (True, False, True)])
def test_lj_forces_match_finite_differences(pbc):
    rng = np.random.default_rng(4)
    L = 7.5
    g = np.stack(np.meshgrid(*[np.arange(3)] * 3, indexing='ij'), -1).reshape(-1, 3) * (L / 3)
    pos = g + rng.uniform(-0.25, 0.25, g.shape)
    cell = np.diag([L] * 3)
    # keep every pair away from the (unsmoothed) cutoff so the energy is differentiable there
    F = olj.lj_forces(pos, cell, pbc, sigma=1.0, epsilon=0.7, rc=2.5)
    Ffd = _fd(lambda p: olj.lj_energy(p, cell, pbc, sigma=1.0, epsilon=0.7, rc=2.5), pos)
    assert np.max(np.abs(F - Ffd)) <= 2e-6 * max(1.0, np.max(np.abs(F)))
    assert np.allclose(F.sum(0), 0.0, atol=1e-10 * max(1.0, np.max(np.abs(F))))     # Newton's third law


@pytest.mark.parametrize('pbc', [(True, True, True), (False, False, False)])
def test_emt_forces_match_finite_differences(pbc):
    rng = np.random.default_rng(5)
    a0 = 3.61
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a0
    pos = np.concatenate([base + np.array([i, j, k]) * a0 for i in range(2) for j in range(2) for k in range(2)])
    pos = pos + rng.uniform(-0.15, 0.15, pos.shape)
    num = np.full(len(pos), 29)
    num[[3, 11, 17]] = 46                       # Cu / Pd mixture
    pos = np.concatenate([pos, [[3.3, 3.1, 8.4], [3.3, 3.1, 9.55]]])
    num = np.concatenate([num, [6, 8]])          # a CO on top
    cell = np.diag([2 * a0, 2 * a0, 14.0])
    F = oemt.emt_forces(pos, num, cell, pbc)
    Ffd = _fd(lambda p: oemt.emt_energy(p, num, cell, pbc), pos, h=2e-5)
    assert np.max(np.abs(F - Ffd)) <= 5e-7 * max(1.0, np.max(np.abs(F)))
    assert np.allclose(F.sum(0), 0.0, atol=1e-9 * max(1.0, np.max(np.abs(F))))


def test_isolated_atom_has_zero_force():
    F = oemt.emt_forces(np.array([[0.0, 0, 0], [30.0, 0, 0]]), [29, 29], np.zeros((3, 3)), (False,) * 3)
    assert np.array_equal(F, np.zeros((2, 3)))
