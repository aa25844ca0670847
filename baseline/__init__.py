"""Reference-arm material (NOT PRODUCT CODE).

``_ref/``      the UNMODIFIED reference (farrisric/mcpy), installed by ``__graft_entry__.build()`` with
               ``pip install --no-deps --target baseline/_ref`` from a copy of ``/root/reference``; git-ignored,
               travels to the GPU box with the snapshot.
``ase_shim/``  a stand-in for the reference's only hard dependency, ``ase`` (3.23.0, not installable offline):
               ``Atoms``, ``geometry``, ``data``, ``units``, ``build``, the ``Calculator`` protocol -- what the
               unmodified reference needs to import and run.  Put on ``sys.path`` only when no real ``ase`` is
               importable.  Its ``LennardJones`` / ``EMT`` / ``LBFGS`` delegate to ``oracle/`` (they are the CPU
               side of the parity tests and of ``bench.py --impl reference``).
``refpath``    puts both on ``sys.path``.
"""
