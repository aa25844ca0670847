"""``LennardJones`` stand-in whose energy is the CPU oracle (oracle/lj.py)."""
import numpy as np

from .calculator import Calculator, all_changes


class LennardJones(Calculator):
    implemented_properties = ['energy', 'free_energy']
    default_parameters = {'epsilon': 1.0, 'sigma': 1.0, 'rc': None, 'ro': None, 'smooth': False}

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.parameters.rc is None:
            self.parameters.rc = 3 * self.parameters.sigma
        if self.parameters.smooth:
            raise NotImplementedError('smooth=True is outside the hot path (shim)')

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        from oracle.lj import lj_energy
        super().calculate(atoms, properties, system_changes)
        p = self.parameters
        e = lj_energy(self.atoms.positions, np.asarray(self.atoms.cell), self.atoms.pbc,
                      p.sigma, p.epsilon, p.rc)
        self.results['energy'] = e
        self.results['free_energy'] = e
