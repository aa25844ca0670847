"""extxyz trajectory frames.

``write_xyz`` has the signature and the output -- byte for byte -- of the reference's
``mcpy.ensembles.base_ensemble.write_xyz`` (``base_ensemble.py:267-305``), which every ensemble calls once per step
at the default ``trajectory_write_interval = 1`` (``base_ensemble.py:27,125-131``) and once per new minimum
(``:142-160``).  The reference builds the frame with one f-string per atom (1.3 ms for 2 000 atoms); here the text
comes from ``_marshal.xyz_frame`` (csrc/marshal.cpp: shortest round-trip digits laid out by ``repr(float)``'s rule).

``install()`` rebinds the name in the reference's modules so an unmodified run writes through it.
"""
from __future__ import annotations

import numpy as np

try:
    from .. import _marshal
except ImportError:                                   # helper not built: python formatting below
    _marshal = None


def frame_text(atoms, energy, extra: str = '') -> str:
    """The frame ``write_xyz`` writes, as a string."""
    cell = np.ascontiguousarray(np.asarray(atoms.get_cell()), dtype=np.float64)
    positions = atoms.positions
    mol_ids = atoms.arrays.get('molecule_id') if hasattr(atoms, 'arrays') else None
    if _marshal is not None:
        try:
            txt = _marshal.xyz_frame(atoms.numbers, positions, cell, float(energy), extra or '', mol_ids)
        except (TypeError, ValueError, OverflowError):        # an energy that is not a real number, odd `extra`
            txt = None
        if txt is not None:
            return txt
    # the reference's own construction (arrays the helper cannot read in place: views, other dtypes)
    symbols = atoms.get_chemical_symbols()
    cell_str = " ".join(f"{v:.8f}" for v in cell.ravel())
    comment = f'energy={energy:.6f}'
    if extra:
        comment += f' {extra}'
    comment += f' Lattice="{cell_str}"'
    if mol_ids is not None:
        comment += ' Properties=species:S:1:pos:R:3:molecule_id:I:1'
    parts = [f"{len(atoms)}\n", comment + '\n']
    for i, (s, p) in enumerate(zip(symbols, positions)):
        line = f"{s} {p[0]} {p[1]} {p[2]}"
        if mol_ids is not None:
            line += f" {mol_ids[i]}"
        parts.append(line + "\n")
    return "".join(parts)


def write_xyz(atoms, energy, file_or_path, extra: str = ''):
    """Append one frame to an open handle (flushed, as the reference does) or to a path."""
    text = frame_text(atoms, energy, extra)
    if isinstance(file_or_path, str):
        with open(file_or_path, 'a') as fh:
            fh.write(text)
    else:
        file_or_path.write(text)
        file_or_path.flush()


def install():
    """Route the reference's trajectory / minima / per-replica writers through ``write_xyz``.  Returns the modules
    that were rebound."""
    import importlib
    done = []
    for name in ('mcpy.ensembles.base_ensemble', 'mcpy.ensembles.batched_replica_exchange',
                 'mcpy.ensembles.replica_exchange'):
        try:
            mod = importlib.import_module(name)
        except ImportError:
            continue
        if hasattr(mod, 'write_xyz'):
            mod.write_xyz = write_xyz
            done.append(name)
    return done
