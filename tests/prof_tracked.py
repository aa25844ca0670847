"""Profiling helper (not a test): the incremental drop-in chain (lj on S2, emt on S3 / S5)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200 import Engine
from mcpy_b200.utils import synthetic
which = sys.argv[1] if len(sys.argv) > 1 else 'lj'
eng = Engine(0)
if which == 'lj':
    cfg = synthetic.s2_lj(); eng.set_lj(1.0, 1.0, 3.0); insert_Z = None
else:
    cfg = synthetic.s3_ag_o() if which == 'emt3' else synthetic.s5_cupd_co(); eng.set_emt(); insert_Z = 8
tb = synthetic.trial_batch(cfg, 64, seed=7, insert_Z=insert_Z)
cfgs = [synthetic.apply_trial(cfg, tb, t) for t in range(64)]
paths = [0, 0, 0]
for rep in range(3):
    t0 = time.perf_counter()
    for p, z in cfgs:
        e, path = eng.tracked_energy(p, z, cfg['cell'], cfg['pbc']); paths[path] += 1
    dt = (time.perf_counter() - t0) / len(cfgs)
print(which, 'us/call', dt * 1e6, paths)
