"""CODATA-2014 values as used by ASE 3.23 (``ase.units``), the few mcpy needs."""
kB = 8.6173303e-05            # eV / K
Bohr = 0.52917721067          # Angstrom
Hartree = 27.211386024367243  # eV
_e = 1.6021766208e-19
_amu = 1.66053904e-27
second = 1e10 * (_e / _amu) ** 0.5
fs = 1e-15 * second
eV = 1.0
Angstrom = 1.0
