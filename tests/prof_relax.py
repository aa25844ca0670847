"""Profiling helper (not a test): a short LBFGS relaxation through mcpyb_relax_lbfgs_host, for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
from mcpy_b200 import Engine
from mcpy_b200.utils import synthetic

which = sys.argv[1] if len(sys.argv) > 1 else 'lj'
eng = Engine(0)
if which == 'lj':
    cfg = synthetic.s1_argon()
    eng.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
else:
    cfg = synthetic.s3_ag_o()
    eng.set_emt()
pos, num, cell, pbc = cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']
eng.relax_lbfgs(pos, num, cell, pbc, fmax=0.0, steps=3)
t0 = time.perf_counter()
e, p, n, fm = eng.relax_lbfgs(pos, num, cell, pbc, fmax=0.0, steps=10)
print('ok', e, n, fm, (time.perf_counter() - t0) / (n + 1) * 1e6, 'us/step')
