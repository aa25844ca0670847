"""Profiling helper (not a test): a short drop-in chain + one trial batch, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200 import Engine
from mcpy_b200.utils import synthetic

which = sys.argv[1] if len(sys.argv) > 1 else 'lj'
eng = Engine(0)
if which == 'lj':
    cfg = synthetic.s2_lj()
    eng.set_lj(1.0, 1.0, 3.0)
    T = 65536
else:
    cfg = synthetic.s5_cupd_co()
    eng.set_emt()
    T = 8192
cell, pbc = cfg['cell'], cfg['pbc']
for _ in range(5):
    eng.potential_energy(cfg['positions'], cfg['numbers'], cell, pbc)
eng.upload([cfg['positions']], [cfg['numbers']], cell.reshape(1, 9), pbc.reshape(1, 3))
tb = synthetic.trial_batch(cfg, T, seed=7, insert_Z=(None if which == 'lj' else 8))
for _ in range(3):
    dE = eng.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
print('ok', float(dE[:4].sum()), eng.launch_count)
