"""Timing probe of the device-resident GCMC chains (row f2): trials per second of whole reference chains.

    python tests/prof_gcmc.py [trials]

Prints one JSON line: per configuration the device time per trial of ONE chain, of 64 chains in one launch (C4's
replica count) and of 148 / 296 chains (one and two per SM), next to the reference's own loop over ``B200Calculator``
(one C-ABI call per trial) and over the CPU oracle calculator on the same chain.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: E402,F401  (puts the reference and the oracle on sys.path)
import numpy as np  # noqa: E402

from mcpy_b200.utils import synthetic  # noqa: E402


def make(cfg, symbol, mu, calc, seed=0, n_moves=3):
    from ase import Atoms
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    atoms = Atoms(numbers=cfg['numbers'], positions=cfg['positions'].copy(), cell=cfg['cell'], pbc=cfg['pbc'])
    cell = Cell(atoms, seed=21 + seed)
    moves = [InsertionMove(cell, [symbol], 31 + seed), DeletionMove(cell, [symbol], 32 + seed),
             DisplacementMove([symbol], 33 + seed, max_displacement=cfg['max_displacement'])]
    ms = MoveSelector([1, 1, 1], moves, seed=34 + seed, n_moves=n_moves)
    return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type=cfg['units'], calculator=calc, mu={symbol: mu},
                                  species=[symbol], temperature=cfg['temperature'], move_selector=ms,
                                  random_seed=35 + seed, traj_file=None, outfile=None)


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    from helpers import OracleLJCalculator
    from mcpy_b200.calculators import B200Calculator
    from mcpy_b200.ensembles import ResidentGCMC
    out = {}
    for name, cfg, symbol, mu in (('S1_argon_500', synthetic.s1_argon(), 'Ar', -0.085),
                                  ('S2_lj_2000_rc3', synthetic.s2_lj(rc=3.0), 'H', -1.0)):
        lj = (cfg['epsilon'], cfg['sigma'], cfg['rc'])
        calc = B200Calculator('lj', epsilon=lj[0], sigma=lj[1], rc=lj[2])
        row = {}
        for chains in (1, 64, 148, 296):
            gs = [make(cfg, symbol, mu, calc, seed=7 * k) for k in range(chains)]
            rg = ResidentGCMC(gs, engine=calc.engine, lj=lj)
            rg.run_trials(200)                                        # warm-up
            t0 = time.perf_counter()
            ms = rg.run_trials(trials)
            wall = time.perf_counter() - t0
            acc = sum(sum(g.move_selector.move_acceptance_total) for g in gs)
            row[f'chains_{chains}'] = {'device_us_per_trial_per_chain': 1e3 * ms / trials,
                                       'trials_per_s': chains * trials / (1e-3 * ms),
                                       'wall_trials_per_s': chains * trials / wall,
                                       'accept_ratio': acc / (chains * (trials + 200)),
                                       'mean_atoms': float(np.mean([len(g.atoms) for g in gs])), **rg.info(),
                                       'cycles_last_launch': rg.cycles(0)}
            rg.close()
        # the reference's loop, one calculator call per trial
        g = make(cfg, symbol, mu, calc)
        for _ in range(20):
            g.do_gcmc_step()
        n = 300
        t0 = time.perf_counter()
        for _ in range(n):
            g.do_gcmc_step()
        row['reference_loop_over_B200Calculator_trials_per_s'] = 3 * n / (time.perf_counter() - t0)
        g = make(cfg, symbol, mu, OracleLJCalculator(rc=lj[2], sigma=lj[1], epsilon=lj[0]))
        n = 40
        t0 = time.perf_counter()
        for _ in range(n):
            g.do_gcmc_step()
        row['reference_loop_over_cpu_oracle_trials_per_s'] = 3 * n / (time.perf_counter() - t0)
        out[name] = row
    print(json.dumps(out))


if __name__ == '__main__':
    main()
