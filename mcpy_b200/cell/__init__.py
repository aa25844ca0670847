"""Sampling-region cells with device free volume (boundary of ``mcpy/cell``)."""
from .b200_cells import B200SphericalCell, B200CustomCell, B200DomeCell  # noqa: F401

__all__ = ['B200SphericalCell', 'B200CustomCell', 'B200DomeCell']
