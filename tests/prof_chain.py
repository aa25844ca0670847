"""Profiling helper (not a test): the sequential one-call-per-trial chain (B200Calculator.get_potential_energy on a
mutated copy per call) on the 2 000-atom S2 box, timed in blocks inside one process."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.utils import synthetic


class _A:
    pass


cfg = synthetic.s2_lj()
tb = synthetic.trial_batch(cfg, 2000, seed=7)
atoms = []
for t in range(2000):
    a = _A()
    a.positions, a.numbers = synthetic.apply_trial(cfg, tb, t)
    a.cell, a.pbc = cfg['cell'], cfg['pbc']
    atoms.append(a)
calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0)
for a in atoms[:50]:
    calc.get_potential_energy(a)
for block in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    t0 = time.perf_counter()
    for a in atoms:
        calc.get_potential_energy(a)
    print('block', block, 'us per call', 1e6 * (time.perf_counter() - t0) / len(atoms))
