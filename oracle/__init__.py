"""CPU oracle for the mcpy hot path (TEST INFRASTRUCTURE -- NOT PRODUCT CODE).

This package restates, on the CPU in numpy / plain C, the arithmetic that the
reference (farrisric/mcpy v1.4.0) performs on the trial-move energy path.  It
exists only so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` can check and time
the reference algorithm.  Nothing under ``mcpy_b200/`` may import it: the
product path must fail loudly when the CUDA library is missing.

Where the arithmetic lives
--------------------------
mcpy delegates the energy itself to a third-party dependency that is NOT in
``/root/reference``: ``ase`` pinned ``3.23.0`` (``uv.lock:9-20``,
``pyproject.toml:8``) -- ``ase/calculators/lj.py::LennardJones``,
``ase/calculators/emt.py::EMT``, ``ase/neighborlist.py``.  The restatements in
:mod:`oracle.lj`, :mod:`oracle.emt` and :mod:`oracle.neighbors` follow the
published algorithm of those classes and are pinned by

* the reference's own numpy restatement ``LJCalc``
  (``benchmark/lammps_gcmc_parity.py:168-187``), which the reference asserts
  equal to ASE within 1e-9 (``:190-203``);
* the LAMMPS two-body probe values (``benchmark/lammps_gcmc_parity.py:285-324``);
* the EMT known answers ``E_EMT(O2)/2 = +0.46 eV`` (``docs/make_figures.py:334``)
  and the energies / optimiser fmax values ASE's documentation prints for its EMT
  tutorials (N2 on Cu(111), H2O, Au on Al(100), bulk Cu, the Ag equation of state),
  reproduced to the last printed digit -- see ``tests/test_oracle_emt.py``;
* for the free volume, ``scipy.spatial.cKDTree`` -- the very library the
  reference calls (``mcpy/cell/spherical_cell.py:130-138``) -- is installed, so
  :mod:`oracle.free_volume` runs the reference's algorithm verbatim in spirit
  and is additionally pinned by ``tests/test_audit_regressions.py:363-372``.

Parity status: LJ and free volume are pinned by reference golden values; EMT
energies and forces by ASE's documented tutorial outputs (7 species, 6-9 digits)
plus the algorithm's identities for the remaining metals; the neighbour list has
no reference test at all -- it is pinned through the energies above (a missed or
doubled pair changes them) and its brute-force definition.  See DESIGN.md section 3.
"""
