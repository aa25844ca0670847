"""Host-side profile of ResidentReplicaExchange: 8 replicas of the 2 000-atom LJ box, an exchange attempt every 10 trials.

    python tests/prof_resident_re.py [steps]      # cProfile of the step loop, top entries by cumulative time
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: E402,F401
import numpy as np  # noqa: E402

from prof_gcmc import make  # noqa: E402
from mcpy_b200.calculators import B200Calculator  # noqa: E402
from mcpy_b200.ensembles import ResidentReplicaExchange  # noqa: E402
from mcpy_b200.utils import synthetic  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    cfg = synthetic.s2_lj(rc=3.0)
    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0)
    mus = np.linspace(-3.0, -1.8, 8)

    def factory(T=None, mu=None, rank=0):
        g = make(cfg, 'H', mu['H'], calc, seed=11 * rank, n_moves=1)
        return g
    re = ResidentReplicaExchange(factory, calc, mus=[{'H': float(m)} for m in mus], gcmc_steps=10 ** 9, exchange_interval=10,
                                 write_out_interval=10 ** 9, seed=31, rank=0, world_size=1, outfile=None,
                                 global_minimum_file=None)
    for r in re.replicas:
        r.initialize_run()
    re._rebatch_initial_energies()

    def loop(n):
        for _ in range(n):
            re._batched_gcmc_step()
            if re._re_step > 0 and re._re_step % 10 == 0:
                re._attempt_exchanges()
            re._re_step += 1
    loop(50)
    t0 = time.perf_counter()
    loop(steps)
    dt = time.perf_counter() - t0
    print(f'{steps} steps x 8 replicas: {dt:.3f} s, {8 * steps / dt:.0f} replica-trials/s, device {re.device_ms:.1f} ms, blocks {re.blocks}')
    pr = cProfile.Profile()
    pr.enable()
    loop(steps)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(22)


if __name__ == '__main__':
    main()
