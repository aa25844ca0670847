import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.utils import synthetic
cfg = synthetic.s2_lj(2000, 3.0, 0.6)
tb = synthetic.trial_batch(cfg, 65536, seed=7)
calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0, device=0)
class A: pass
a = A(); a.positions, a.numbers, a.cell, a.pbc = cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']
th = {k: np.ascontiguousarray(tb[k]) for k in ('old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')}
for i in range(6):
    t0 = time.perf_counter(); calc.set_configuration(a); t1 = time.perf_counter()
    calc.engine.trial_delta_e(th['old_off'], th['old_idx'], th['new_off'], th['new_pos'], th['new_Z']); t2 = time.perf_counter()
    print('set_config %.1f us, trials %.1f us' % ((t1 - t0) * 1e6, (t2 - t1) * 1e6), file=sys.stderr)
