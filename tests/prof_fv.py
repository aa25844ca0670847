"""Profiling helper (not a test): device-PCG64 free volume on S3 and S5."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcpy_b200 import Engine
from mcpy_b200.utils import synthetic
eng = Engine(0)
for cfg in (synthetic.s3_ag_o(), synthetic.s5_cupd_co()):
    P = int(cfg['mc_sample_points'])
    rng = np.random.default_rng(3)
    for _ in range(3):
        t0 = time.perf_counter()
        r = eng.sphere_free_volume_pcg64(rng, P, cfg['sphere_radius'], np.zeros(3), cfg['positions'], cfg['radii'])
        dt = time.perf_counter() - t0
    print(cfg['name'], r, f'{dt*1e6:.0f} us')
