"""``ase.calculators.calculator`` stand-in: the caching ``Calculator`` protocol."""
import numpy as np

all_changes = ['positions', 'numbers', 'cell', 'pbc', 'initial_charges', 'initial_magmoms']


class PropertyNotImplementedError(NotImplementedError):
    pass


class CalculatorError(RuntimeError):
    pass


def compare_atoms(atoms1, atoms2, tol=1e-15):
    if atoms1 is None:
        return all_changes[:]
    changes = []
    if len(atoms1) != len(atoms2) or not np.array_equal(atoms1.numbers, atoms2.numbers):
        changes.append('numbers')
    if len(atoms1) != len(atoms2) or not np.allclose(atoms1.positions, atoms2.positions,
                                                      rtol=0, atol=tol):
        changes.append('positions')
    if not np.allclose(np.asarray(atoms1.cell), np.asarray(atoms2.cell), rtol=0, atol=tol):
        changes.append('cell')
    if not np.array_equal(atoms1.pbc, atoms2.pbc):
        changes.append('pbc')
    return changes


class _Parameters(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value


class Calculator:
    implemented_properties = []
    default_parameters = {}
    nolabel = True

    def __init__(self, restart=None, label=None, atoms=None, **kwargs):
        self.atoms = None
        self.results = {}
        self.parameters = _Parameters(self.default_parameters)
        self.parameters.update(kwargs)
        if atoms is not None:
            atoms.calc = self

    def reset(self):
        self.atoms = None
        self.results = {}

    def check_state(self, atoms, tol=1e-15):
        return compare_atoms(self.atoms, atoms, tol=tol)

    def calculate(self, atoms=None, properties=('energy',), system_changes=all_changes):
        if atoms is not None:
            self.atoms = atoms.copy()

    def get_property(self, name, atoms=None, allow_calculation=True):
        if name not in self.implemented_properties:
            raise PropertyNotImplementedError(f'{name} property not implemented')
        if atoms is None:
            atoms = self.atoms
            system_changes = []
        else:
            system_changes = self.check_state(atoms)
            if system_changes:
                self.reset()
        if name not in self.results:
            if not allow_calculation:
                return None
            self.calculate(atoms, [name], system_changes)
        if name not in self.results:
            raise PropertyNotImplementedError(f'{name} not present in this calculation')
        result = self.results[name]
        if isinstance(result, np.ndarray):
            result = result.copy()
        return result

    def get_potential_energy(self, atoms=None, force_consistent=False):
        return self.get_property('energy', atoms)

    def get_forces(self, atoms=None):
        return self.get_property('forces', atoms)
