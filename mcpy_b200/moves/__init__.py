"""Proposal-side routing (boundary of ``mcpy/moves``): the ``min_insert`` distance loop on the GPU."""
from .insertion import B200InsertionMove, B200MoleculeInsertionMove  # noqa: F401

__all__ = ['B200InsertionMove', 'B200MoleculeInsertionMove']
