"""Optimizer names; ``run(steps=0)`` is a no-op, anything else needs forces."""


class _Optimizer:
    def __init__(self, atoms, logfile=None, **kwargs):
        self.atoms = atoms
        self.nsteps = 0

    def run(self, fmax=0.05, steps=100000000):
        if steps == 0:
            return True
        raise NotImplementedError('relaxation needs forces: outside the hot path (shim)')


class LBFGS(_Optimizer):
    pass


class FIRE(_Optimizer):
    pass


class BFGS(_Optimizer):
    pass
