"""Names only: the hot path never reaches these (``mcpy/moves/go_moves.py``)."""


def neighbor_list(*args, **kwargs):
    raise NotImplementedError('ase.neighborlist.neighbor_list is outside the shim')


def natural_cutoffs(atoms, mult=1, **kwargs):
    raise NotImplementedError('ase.neighborlist.natural_cutoffs is outside the shim')


class NeighborList:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('ase.neighborlist.NeighborList is outside the shim')
