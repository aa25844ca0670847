"""Calculator plug-ins (the boundary of ``mcpy/calculators``).

A maintainer of the reference registers the class in
``mcpy/calculators/__init__.py:5-19`` with the same optional-import block the
MACE / Alchemi back-ends use -- see INTEGRATION.md.
"""
from .b200_calculator import B200Calculator, B200LBFGS, B200RelaxCalculator  # noqa: F401

__all__ = ['B200Calculator', 'B200RelaxCalculator', 'B200LBFGS']
