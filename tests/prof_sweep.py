"""Timing helper (not a test): device time per stage of the batched trial pipeline for R replicas of S2 with
`per` proposals each.   python tests/prof_sweep.py <per> <R>"""
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
tbs = [bench.proposal_batch(cfgs, seed=b) for b in range(4)]
for tb in tbs:
    dE, pairs = eng.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'], rep=tb['rep'],
                                  return_pairs=True)
eng.stage_timing(True)
for k in range(12):
    tb = tbs[k % 4]
    eng.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'], rep=tb['rep'])
    st, calls = eng.stage_timing(True)
eng.stage_timing(False)
sites = int(tbs[0]['old_off'][-1] + tbs[0]['new_off'][-1])
flop = sites * 27 * 0.6 * 27 * 17 + int(pairs.sum()) * 8
peak = eng.fp64_peak_tflops()
print(f'per={per} R={R} trials={per * R} sites={sites}', {k: round(1e3 * v, 1) for k, v in st.items()},
      f'rows frac {flop / (st["row_sums"] * 1e-3) / 1e12 / peak:.3f}  pipeline {1e3 * sum(st.values()):.1f} us'
      f'  {per * R / (sum(st.values()) * 1e-3) / 1e6:.0f} M dE/s')
