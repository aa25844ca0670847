"""GPU: relax-then-energy (mcpyb_relax_lbfgs_host, B200RelaxCalculator) against the LBFGS restatement of
oracle/lbfgs.py driven by the oracle's forces: the same number of steps, the same relaxed geometry and energy.

Tolerances: the two loops run the same iteration on forces that agree to 1e-10.  Steps taken: identical.  Energy:
|dE| <= 1e-8 * max(1, |E|).  Positions: |dx| <= 1e-7 A over the first few steps; over a full relaxation of a free
cluster the rounding differences drift along its soft modes (rigid motions cost no energy, so nothing restores
them), hence |dx| <= 1e-4 A there -- still far below the optimiser's own resolution fmax / curvature."""
import numpy as np
import pytest

from mcpy_b200 import Engine
from mcpy_b200.calculators import B200RelaxCalculator
from mcpy_b200.utils import synthetic
from oracle import emt as oemt, lbfgs, lj as olj

pytestmark = pytest.mark.gpu


def _lj_cluster(seed=4, n=20):
    rng = np.random.default_rng(seed)
    g = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], float)[:n] * 1.12
    return g + rng.uniform(-0.08, 0.08, size=g.shape) + 6.0


@pytest.mark.parametrize('pbc', [False, True])
def test_lj_relaxation_follows_the_oracle_iteration(pbc):
    pos = _lj_cluster()
    cell, pb = np.diag([14.0] * 3), np.array([pbc] * 3)
    num = np.ones(len(pos), np.int32)
    eng = Engine(0)
    eng.set_lj(1.0, 1.0, 2.5)
    ef = lambda p: (olj.lj_energy(p, cell, pb, 1.0, 1.0, 2.5), olj.lj_forces(p, cell, pb, 1.0, 1.0, 2.5))
    for steps, fmax, fixed in [(0, 0.05, None), (1, 0.05, None), (7, 1e-9, None), (400, 1e-4, None), (400, 1e-4, [0, 5, 9])]:
        eo, po, so, fo = lbfgs.lbfgs_relax(ef, pos, fmax=fmax, steps=steps, fixed=fixed)
        before = pos.copy()
        e, p, s, fm = eng.relax_lbfgs(pos, num, cell, pb, fmax=fmax, steps=steps, fixed=fixed)
        assert np.array_equal(pos, before)                         # the engine call works on a copy
        assert s == so, (steps, fmax, s, so)
        assert np.max(np.abs(p - po)) <= (1e-7 if steps <= 7 else 1e-4) and abs(e - eo) <= 1e-8 * max(1.0, abs(eo))
        if steps == 400:
            assert fm < fmax and e < ef(pos)[0]
        if fixed:
            assert np.array_equal(p[fixed], pos[fixed])
    eng.close()


def test_emt_relaxation_and_the_base_calculator_mirror():
    """O atoms on the Ag octahedron (S3): B200RelaxCalculator relaxes atoms IN PLACE like
    BaseCalculator.get_potential_energy (base_calculator.py:14-22) and returns the relaxed energy."""
    from ase import Atoms
    from ase.constraints import FixAtoms
    cfg = synthetic.s3_ag_o(n_o=4)
    syms = ['Ag' if z == 47 else 'O' for z in cfg['numbers']]
    cell = np.diag([40.0] * 3)
    atoms = Atoms(syms, positions=cfg['positions'] + 20.0, cell=cell, pbc=False)
    fixed = list(range(0, 140, 3))
    atoms.set_constraint(FixAtoms(indices=fixed))
    start = atoms.positions.copy()
    pb = np.zeros(3, bool)
    ef = lambda p: (oemt.emt_energy(p, cfg['numbers'], cell, pb), oemt.emt_forces(p, cfg['numbers'], cell, pb))
    eo, po, so, fo = lbfgs.lbfgs_relax(ef, start, fmax=0.05, steps=12, fixed=fixed)
    calc = B200RelaxCalculator('emt', steps=12, fmax=0.05)
    e = calc.get_potential_energy(atoms)
    assert calc.last_steps == so and so > 0
    assert np.max(np.abs(atoms.positions - po)) <= 1e-7 and abs(e - eo) <= 1e-8 * max(1.0, abs(eo))
    assert np.array_equal(atoms.positions[fixed], start[fixed]) and e < ef(start)[0]
    # steps = 0: single point, geometry untouched (the reference's documented way to switch relaxation off)
    single = B200RelaxCalculator('emt', steps=0, fmax=0.05, calculator=calc.calculator)
    frozen = atoms.positions.copy()
    e1 = single.get_potential_energy(atoms)
    assert np.array_equal(atoms.positions, frozen) and abs(e1 - e) <= 1e-10 * max(1.0, abs(e))
    with pytest.raises(ValueError):
        B200RelaxCalculator('emt', steps=-1)
