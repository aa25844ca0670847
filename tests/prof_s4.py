"""Profiling helper (not a test): get_potential_energies on 64 x 2000-atom LJ replicas through resident bases,
timed repeatedly in one process; `concat` forces the concatenating entry point."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.engine import Engine
from mcpy_b200.utils import synthetic


class _A:
    pass


base, calls = [], []
for k in range(64):
    c = synthetic.s2_lj(2000, 3.0, 0.6, seed=100 + k)
    tbk = synthetic.trial_batch(c, 8, seed=300 + k)
    base.append(c)
    calls.append([synthetic.apply_trial(c, tbk, t) for t in range(8)])


def batch_of(t):
    reps = []
    for k in range(64):
        a = _A()
        a.positions, a.numbers = calls[k][t]
        a.cell, a.pbc = base[k]['cell'], base[k]['pbc']
        reps.append(a)
    return reps


batches = [batch_of(t) for t in range(8)]
calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0, incremental=True)
orig = Engine._row_pointers
for mode in sys.argv[1:] or ['rows', 'concat', 'rows', 'concat']:
    Engine._row_pointers = staticmethod((lambda p, z: None) if mode == 'concat' else orig)
    for bt in batches[:2]:
        calc.get_potential_energies(bt)
    t0 = time.perf_counter()
    for _ in range(5):
        for bt in batches:
            calc.get_potential_energies(bt)
    print(mode, 'ms per call', 1e3 * (time.perf_counter() - t0) / 40)
