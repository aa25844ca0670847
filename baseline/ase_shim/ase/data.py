"""Symbols, numbers and standard atomic weights (``ase.data`` stand-in)."""
import numpy as np

chemical_symbols = [
    'X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al',
    'Si', 'P', 'S', 'Cl', 'Ar', 'K', 'Ca', 'Sc', 'Ti', 'V', 'Cr', 'Mn', 'Fe',
    'Co', 'Ni', 'Cu', 'Zn', 'Ga', 'Ge', 'As', 'Se', 'Br', 'Kr', 'Rb', 'Sr', 'Y',
    'Zr', 'Nb', 'Mo', 'Tc', 'Ru', 'Rh', 'Pd', 'Ag', 'Cd', 'In', 'Sn', 'Sb', 'Te',
    'I', 'Xe', 'Cs', 'Ba', 'La', 'Ce', 'Pr', 'Nd', 'Pm', 'Sm', 'Eu', 'Gd', 'Tb',
    'Dy', 'Ho', 'Er', 'Tm', 'Yb', 'Lu', 'Hf', 'Ta', 'W', 'Re', 'Os', 'Ir', 'Pt',
    'Au', 'Hg', 'Tl', 'Pb', 'Bi', 'Po', 'At', 'Rn']
atomic_numbers = {s: z for z, s in enumerate(chemical_symbols)}

_masses = {
    'X': 1.0, 'H': 1.008, 'He': 4.002602, 'Li': 6.94, 'Be': 9.0121831, 'B': 10.81,
    'C': 12.011, 'N': 14.007, 'O': 15.999, 'F': 18.998403163, 'Ne': 20.1797,
    'Na': 22.98976928, 'Mg': 24.305, 'Al': 26.9815385, 'Si': 28.085,
    'P': 30.973761998, 'S': 32.06, 'Cl': 35.45, 'Ar': 39.948, 'K': 39.0983,
    'Ca': 40.078, 'Ti': 47.867, 'Fe': 55.845, 'Co': 58.933194, 'Ni': 58.6934,
    'Cu': 63.546, 'Zn': 65.38, 'Rh': 102.9055, 'Pd': 106.42, 'Ag': 107.8682,
    'Pt': 195.084, 'Au': 196.966569, 'Kr': 83.798, 'Xe': 131.293,
}
atomic_masses = np.array([_masses.get(s, float(2 * z)) for z, s in
                          enumerate(chemical_symbols)], dtype=float)
