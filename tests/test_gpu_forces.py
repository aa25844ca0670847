"""GPU: LJ and EMT forces (mcpyb_forces_host / mcpyb_energy_forces_host) against the oracle's analytic
forces on the same seeded inputs, and a relaxation driven by them.

Tolerance: |F - F_oracle| <= 1e-10 * max(1, max|F|) per component (fp64)."""
import numpy as np
import pytest

from mcpy_b200 import Engine
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.utils import synthetic
from oracle import emt as oemt, lj as olj

pytestmark = pytest.mark.gpu


def _close(F, Fo):
    scale = max(1.0, float(np.max(np.abs(Fo)))) if Fo.size else 1.0
    return float(np.max(np.abs(F - Fo))) <= 1e-10 * scale if Fo.size else True


@pytest.mark.parametrize('pbc', [(True, True, True), (False, False, False), (True, True, False)])
def test_lj_forces_match_oracle(pbc):
    cfg = synthetic.s1_argon()
    pos = cfg['positions'][:300]
    eng = Engine(0)
    eng.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
    e, F = eng.energy_forces(pos, cfg['numbers'][:300], cfg['cell'], np.array(pbc))
    Fo = olj.lj_forces(pos, cfg['cell'], pbc, cfg['sigma'], cfg['epsilon'], cfg['rc'])
    assert _close(F, Fo)
    eo = olj.lj_energy(pos, cfg['cell'], pbc, cfg['sigma'], cfg['epsilon'], cfg['rc'])
    assert abs(e - eo) <= 1e-10 * max(1.0, abs(eo))
    eng.close()


def test_lj_forces_thin_box_and_batch():
    """A box thinner than the cutoff (self images) and a ragged batch through upload + forces."""
    rng = np.random.default_rng(8)
    eng = Engine(0)
    eng.set_lj(1.0, 1.0, 3.0)
    cell = np.diag([2.2, 9.0, 9.0])
    pos = rng.uniform(0, 1, (20, 3)) @ cell
    pos = pos[np.argsort(pos[:, 1])]
    pos[:, 1] = np.linspace(0.3, 8.5, 20)               # keep atoms apart
    _, F = eng.energy_forces(pos, np.ones(20, np.int32), cell, np.ones(3, bool))
    assert _close(F, olj.lj_forces(pos, cell, (True,) * 3, 1.0, 1.0, 3.0))
    a, b = synthetic.s2_lj(400, 3.0, 0.6, seed=3), synthetic.s2_lj(150, 3.0, 0.5, seed=4)
    eng.upload([a['positions'], b['positions']], [a['numbers'], b['numbers']],
               np.stack([a['cell'].reshape(9), b['cell'].reshape(9)]), np.ones((2, 3), bool))
    F = eng.forces()
    assert _close(F[:400], olj.lj_forces(a['positions'], a['cell'], a['pbc'], 1.0, 1.0, 3.0))
    assert _close(F[400:], olj.lj_forces(b['positions'], b['cell'], b['pbc'], 1.0, 1.0, 3.0))
    eng.close()


def test_emt_forces_match_oracle():
    cfg = synthetic.s3_ag_o()
    eng = Engine(0)
    eng.set_emt()
    e, F = eng.energy_forces(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc'])
    Fo = oemt.emt_forces(cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc'])
    assert _close(F, Fo)
    assert np.max(np.abs(F.sum(0))) <= 1e-9 * max(1.0, np.max(np.abs(F)))
    # periodic CuPd slab + CO, and an isolated atom (deds = 0, zero force)
    rng = np.random.default_rng(5)
    a0 = 3.61
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a0
    pos = np.concatenate([base + np.array([i, j, k]) * a0 for i in range(3) for j in range(3) for k in range(2)])
    pos = pos + rng.uniform(-0.1, 0.1, pos.shape)
    num = np.full(len(pos), 29, np.int32)
    num[rng.choice(len(pos), 6, replace=False)] = 46
    pos = np.concatenate([pos, [[3.3, 3.1, 8.4], [3.3, 3.1, 9.55], [5.0, 5.0, 20.0]]])
    num = np.concatenate([num, [6, 8, 1]]).astype(np.int32)
    cell = np.diag([3 * a0, 3 * a0, 30.0])
    pbc = np.array([True, True, False])
    _, F = eng.energy_forces(pos, num, cell, pbc)
    Fo = oemt.emt_forces(pos, num, cell, pbc)
    assert _close(F, Fo)
    assert np.array_equal(F[-1], np.zeros(3))
    eng.close()


def test_steepest_descent_on_gpu_forces_lowers_the_energy():
    """The calculator serves the ASE protocol with forces: a relaxation loop written against
    ``calculate(atoms, ['energy', 'forces'])`` converges (what CanonicalEnsemble.relax needs)."""
    cfg = synthetic.s3_ag_o()

    from ase import Atoms            # real ASE when installed, else the test-only shim (tests/conftest.py)
    atoms = Atoms(numbers=cfg['numbers'], positions=cfg['positions'].copy(), cell=cfg['cell'], pbc=cfg['pbc'])
    calc = B200Calculator('emt')
    assert 'forces' in calc.implemented_properties
    calc.calculate(atoms, ['energy', 'forces'])
    e0, f0 = calc.results['energy'], calc.results['forces']
    fmax0 = np.abs(f0).max()
    for _ in range(60):
        calc.calculate(atoms, ['energy', 'forces'])
        atoms.positions = atoms.positions + 0.02 * np.clip(calc.results['forces'], -2.0, 2.0)
        assert atoms.positions.shape == (len(cfg['numbers']), 3)
    calc.calculate(atoms, ['energy', 'forces'])
    assert calc.results['energy'] < e0 - 1e-3
    assert np.abs(calc.results['forces']).max() < fmax0
