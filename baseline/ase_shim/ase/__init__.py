"""TEST-ONLY stand-in for the third-party ``ase`` package (NOT PRODUCT CODE).

ASE 3.23.0 is the reference's only hard dependency (``pyproject.toml:7-10``)
and cannot be installed in the build container (no network, no wheel).  This
package re-implements, from ASE's documented behaviour, the small slice of the
API that ``/root/reference/mcpy`` touches on the GCMC hot path -- ``Atoms``
(arrays, ``+=``, ``del``, ``wrap``, ``copy``, constraints), ``find_mic``,
``wrap_positions``, ``atomic_masses`` -- so that the UNMODIFIED reference
drivers (``GrandCanonicalEnsemble``, cells, moves, ``BatchedReplicaExchange``)
can be imported in this container to generate golden fixtures
(``tests/golden/make_golden.py``) and so that the product's Python classes can
be exercised against an ASE-shaped ``Atoms`` in the tests.  Energies inside
``ase.calculators.{lj,emt}`` here come from the CPU oracle.

It is put on ``sys.path`` only by ``tests/conftest.py`` / the golden
generator, and only when a real ``ase`` is not importable.  Nothing under
``mcpy_b200/`` depends on it.
"""
from .atoms import Atoms, Atom  # noqa: F401
from . import units  # noqa: F401

__version__ = '3.23.0+shim'
