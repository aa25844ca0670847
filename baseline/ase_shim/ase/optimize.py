"""``ase.optimize`` stand-in.  ``LBFGS`` runs ASE's iteration as restated in ``oracle/lbfgs.py`` (no line
search, ASE's defaults) on ``atoms.get_forces()``; the other names only support ``run(steps=0)``."""
import numpy as np


class _Optimizer:
    def __init__(self, atoms, logfile=None, trajectory=None, **kwargs):
        self.atoms = atoms
        self.nsteps = 0

    def run(self, fmax=0.05, steps=100000000):
        if steps == 0:
            return True
        raise NotImplementedError(f'{type(self).__name__} is outside the shim (only LBFGS relaxes)')

    def get_number_of_steps(self):
        return self.nsteps


class LBFGS(_Optimizer):
    def __init__(self, atoms, logfile=None, trajectory=None, maxstep=0.2, memory=100, damping=1.0, alpha=70.0,
                 **kwargs):
        super().__init__(atoms, logfile=logfile)
        self.maxstep, self.memory, self.damping, self.alpha = maxstep, memory, damping, alpha

    def run(self, fmax=0.05, steps=100000000):
        from oracle.lbfgs import lbfgs_relax
        atoms = self.atoms
        fixed = [int(i) for c in atoms.constraints for i in c.get_indices()]

        def energy_forces(p):
            atoms.positions[...] = p
            f = atoms.calc.get_forces(atoms)
            return atoms.calc.get_potential_energy(atoms), f

        _, pos, self.nsteps, fm = lbfgs_relax(energy_forces, atoms.positions, fmax=fmax, steps=steps,
                                              fixed=fixed or None, maxstep=self.maxstep, memory=self.memory,
                                              alpha=self.alpha, damping=self.damping)
        atoms.positions[...] = pos
        return bool(fm < fmax)


class FIRE(_Optimizer):
    pass


class BFGS(_Optimizer):
    pass
