"""``gcmc.run(steps)`` with a row and a frame written EVERY step (the reference's default intervals), 2 000-atom LJ box:
the reference's loop over B200Calculator against ResidentGCMC.run (one launch per step, positions back for the frame).

    python tests/prof_gcmc_files.py [steps]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: E402,F401

from prof_gcmc import make  # noqa: E402
from mcpy_b200.calculators import B200Calculator  # noqa: E402
from mcpy_b200.ensembles import ResidentGCMC  # noqa: E402
from mcpy_b200.utils import synthetic  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    cfg = synthetic.s2_lj(rc=3.0)
    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag in ('reference_loop', 'resident'):
            for interval in (1, 100):
                g = make(cfg, 'H', -1.0, calc, n_moves=3)
                g._outfile, g._traj_file = os.path.join(tmp, f'{tag}.out'), os.path.join(tmp, f'{tag}.xyz')
                g._outfile_write_interval = g._trajectory_write_interval = interval
                t0 = time.perf_counter()
                if tag == 'resident':
                    ResidentGCMC(g).run(steps)
                else:
                    g.run(steps)
                dt = time.perf_counter() - t0
                out[f'{tag}_write_every_{interval}'] = {'steps_per_s': steps / dt, 'trials_per_s': 3 * steps / dt, 'ms_per_step': 1e3 * dt / steps}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
