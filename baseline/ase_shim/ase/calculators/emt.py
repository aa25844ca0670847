"""``EMT`` stand-in whose energy and forces are the CPU oracle (oracle/emt.py, oracle/cfast.py)."""
import numpy as np

from .calculator import Calculator, all_changes


class EMT(Calculator):
    implemented_properties = ['energy', 'free_energy', 'forces']
    default_parameters = {'asap_cutoff': False}

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        from oracle import cfast
        from oracle.emt import emt_forces
        super().calculate(atoms, properties, system_changes)
        a = self.atoms
        e = cfast.emt_energy(a.positions, a.numbers, np.asarray(a.cell), a.pbc)
        self.results['energy'] = e
        self.results['free_energy'] = e
        if properties is not None and 'forces' in properties:
            self.results['forces'] = emt_forces(a.positions, a.numbers, np.asarray(a.cell), a.pbc)
