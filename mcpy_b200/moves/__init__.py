"""Proposal-side routing (boundary of ``mcpy/moves``): the ``min_insert`` distance loop on the GPU,
rigid-molecule deletion / displacement with cached candidate bookkeeping."""
from .insertion import B200InsertionMove, B200MoleculeInsertionMove  # noqa: F401
from .molecules import B200MoleculeDeletionMove, B200MoleculeDisplacementMove, MoleculeIndex  # noqa: F401

__all__ = ['B200InsertionMove', 'B200MoleculeInsertionMove', 'B200MoleculeDeletionMove',
           'B200MoleculeDisplacementMove', 'MoleculeIndex']
