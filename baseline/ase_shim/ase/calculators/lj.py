"""``LennardJones`` stand-in whose energy and forces are the CPU oracle (oracle/lj.py, oracle/cfast.py)."""
import numpy as np

from .calculator import Calculator, all_changes


class LennardJones(Calculator):
    implemented_properties = ['energy', 'free_energy', 'forces']
    default_parameters = {'epsilon': 1.0, 'sigma': 1.0, 'rc': None, 'ro': None, 'smooth': False}

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.parameters.rc is None:
            self.parameters.rc = 3 * self.parameters.sigma
        if self.parameters.smooth:
            raise NotImplementedError('smooth=True is outside the hot path (shim)')

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        from oracle import cfast
        from oracle.lj import lj_forces
        super().calculate(atoms, properties, system_changes)
        p = self.parameters
        a = self.atoms
        e = cfast.lj_energy(a.positions, np.asarray(a.cell), a.pbc, p.sigma, p.epsilon, p.rc)
        self.results['energy'] = e
        self.results['free_energy'] = e
        if properties is not None and 'forces' in properties:
            self.results['forces'] = lj_forces(a.positions, np.asarray(a.cell), a.pbc, p.sigma, p.epsilon, p.rc)
