"""One device-resident 2 000-atom LJ chain (or a few): python tests/prof_gcmc_one.py [chains] [trials] -- the ncu / cycle-counter target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa
from prof_gcmc import make
from mcpy_b200.utils import synthetic
from mcpy_b200.calculators import B200Calculator
from mcpy_b200.ensembles import ResidentGCMC
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
cfg = synthetic.s2_lj(rc=3.0)
lj = (1.0, 1.0, 3.0)
calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=3.0)
gs = [make(cfg, 'H', -1.0, calc, seed=7 * k) for k in range(chains)]
rg = ResidentGCMC(gs, engine=calc.engine, lj=lj)
ms = rg.run_trials(trials)
print('ms', ms, rg.cycles(0))
