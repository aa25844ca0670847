"""TEST INFRASTRUCTURE -- ASE's LBFGS iteration, restated.

``mcpy.calculators.BaseCalculator.get_potential_energy`` (``mcpy/calculators/base_calculator.py:14-22``) and
``CanonicalEnsemble.relax`` (``mcpy/ensembles/canonical_ensemble.py:115-123``) relax every trial configuration with
``ase.optimize.LBFGS`` before they read its energy.  ASE (pinned 3.23.0, ``uv.lock:9-20``) is not vendored and not
installable here, so the iteration below follows ``ase/optimize/lbfgs.py`` and ``ase/optimize/optimize.py`` AS
RECALLED: ``LBFGS.step`` (two-loop recursion over the stored ``(s, y, rho)`` with ``H0 = 1/alpha``),
``LBFGS.update`` (pair appended from the previous positions / forces, oldest pair dropped beyond ``memory``),
``LBFGS.determine_step`` (whole step scaled so that the longest atomic displacement is ``maxstep``),
``Optimizer.run`` / ``converged`` (``max_i |f_i| < fmax`` checked before every step, at most ``steps`` steps), no line
search, defaults ``maxstep=0.2, memory=100, alpha=70, damping=1``.

PARITY UNPINNED: the reference's tests hold no optimiser trajectory or relaxed energy for LJ / EMT, and ASE itself
cannot run here.  What the tests can and do check: the end state is a stationary point to ``fmax`` whose energy is
not above the start, the engine's loop (``mcpyb_relax_lbfgs_host``) follows this restatement step for step, and
``steps=0`` is the single-point energy.
"""
import numpy as np


def lbfgs_relax(energy_forces, positions, fmax=0.05, steps=100_000_000, fixed=None, maxstep=0.2, memory=100,
                alpha=70.0, damping=1.0):
    """``energy_forces(pos) -> (E, F (n,3))``.  Returns ``(E, relaxed positions, steps taken, max |f|)``."""
    pos = np.array(positions, dtype=float, copy=True)
    n = len(pos)
    fixed = np.zeros(0, int) if fixed is None else np.asarray(fixed, int)

    def evaluate(p):
        e, f = energy_forces(p)
        f = np.array(f, dtype=float, copy=True)
        f[fixed] = 0.0                              # FixAtoms.adjust_forces
        return e, f

    e, f = evaluate(pos)
    s, y, rho = [], [], []
    r0 = f0 = None
    iteration = nsteps = 0
    H0 = 1.0 / alpha
    while n and nsteps < steps and not (np.linalg.norm(f, axis=1).max() < fmax):
        if iteration > 0:                           # LBFGS.update
            s0 = pos.reshape(-1) - r0.reshape(-1)
            y0 = f0.reshape(-1) - f.reshape(-1)
            s.append(s0); y.append(y0); rho.append(1.0 / np.dot(y0, s0))
        if iteration > memory:
            s.pop(0); y.pop(0); rho.pop(0)
        loopmax = min(memory, iteration)
        a = np.empty(loopmax)
        q = -f.reshape(-1)
        for i in range(loopmax - 1, -1, -1):
            a[i] = rho[i] * np.dot(s[i], q)
            q = q - a[i] * y[i]
        z = H0 * q
        for i in range(loopmax):
            b = rho[i] * np.dot(y[i], z)
            z = z + s[i] * (a[i] - b)
        p = -z.reshape(-1, 3)
        longest = np.sqrt((p ** 2).sum(1)).max()    # LBFGS.determine_step
        if longest >= maxstep:
            p = p * (maxstep / longest)
        r0, f0 = pos.copy(), f.copy()
        pos = pos + p * damping
        iteration += 1
        e, f = evaluate(pos)
        nsteps += 1
    return e, pos, nsteps, float(np.linalg.norm(f, axis=1).max()) if n else 0.0
