"""Profiling helper (not a test): the bench's sweep -- 8 resident replicas of S2, 8 192 proposals each -- for ncu.

    python tests/prof_shard.py [proposals per replica] [replicas]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from mcpy_b200 import Engine

per = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
bench.TRIALS_PER_REPLICA = per
bench.REPLICAS_PER_GPU = R
cfgs = bench.local_configs(0)[:R]
eng = Engine(0)
eng.set_lj(1.0, 1.0, 3.0)
eng.upload([c['positions'] for c in cfgs], [c['numbers'] for c in cfgs],
           np.stack([c['cell'].reshape(9) for c in cfgs]), np.stack([c['pbc'] for c in cfgs]))
for b in range(5):
    tb = bench.proposal_batch(cfgs, seed=b)
    dE = eng.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'], rep=tb['rep'])
print('ok', float(dE[:4].sum()), eng.launch_count)
