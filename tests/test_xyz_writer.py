"""The extxyz frame writer against the reference's own write_xyz (base_ensemble.py:267-305): identical bytes for
periodic / cluster frames, molecule ids, an `extra` field, awkward floats, empty frames; and a whole unmodified
GrandCanonicalEnsemble run writes the same trajectory through the installed writer.  No GPU needed."""
import io
import struct

import numpy as np
import pytest

from baseline import refpath

refpath.enable()
from ase import Atoms                                              # noqa: E402
from mcpy.ensembles import base_ensemble as ref_be                 # noqa: E402

from mcpy_b200.utils import xyz                                    # noqa: E402

REF_WRITE = ref_be.write_xyz           # bound before any install()


def _both(atoms, energy, extra=''):
    a, b = io.StringIO(), io.StringIO()
    REF_WRITE(atoms, energy, a, extra=extra)
    xyz.write_xyz(atoms, energy, b, extra=extra)
    return a.getvalue(), b.getvalue()


def test_helper_is_built():
    assert xyz._marshal is not None and hasattr(xyz._marshal, 'xyz_frame')


@pytest.mark.parametrize('n', [0, 1, 57, 2000])
def test_frames_are_byte_identical(n):
    rng = np.random.default_rng(n)
    atoms = Atoms(numbers=rng.choice([8, 18, 29, 46, 47], n), positions=rng.normal(0, 9, (n, 3)),
                  cell=rng.random((3, 3)) * 20, pbc=[True, False, True])
    want, got = _both(atoms, -1234.56789012)
    assert got == want
    want, got = _both(atoms, 7.0, extra='step=42 score=-1.234000')
    assert got == want


def test_molecule_ids_and_awkward_floats():
    rnd = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, 1e-4, 1e-5, 9.999e-5, 1e15, 1e16, 9999999999999998.0, 1.2345678901234568e17, 5e-324,
            1.7976931348623157e308, float('nan'), float('inf'), -float('inf'), 100000.0, 1e22, 1e23, 2 / 3, 0.1 + 0.2]
    vals += [struct.unpack('d', struct.pack('Q', int(x)))[0] for x in rnd.integers(0, 2 ** 63, 1000)]
    vals += list(10.0 ** rnd.uniform(-320, 308, 1000) * rnd.choice([-1, 1], 1000))
    pos = np.array(vals[:len(vals) // 3 * 3]).reshape(-1, 3)
    atoms = Atoms(numbers=np.full(len(pos), 6), positions=pos, cell=np.eye(3) * 1e-9, pbc=False)
    atoms.arrays['molecule_id'] = rnd.integers(-1, 500, len(pos))
    for energy in (0.0, -0.0000004, float('nan'), float('inf'), 1e300):
        want, got = _both(atoms, energy)
        assert got == want


def test_views_take_the_python_path_with_the_same_bytes():
    rng = np.random.default_rng(9)
    atoms = Atoms(numbers=[29] * 6, positions=rng.random((6, 3)), cell=np.eye(3) * 5, pbc=True)
    atoms.arrays['positions'] = np.asfortranarray(rng.random((6, 3)))     # not C-contiguous: helper declines
    want, got = _both(atoms, 1.0)
    assert got == want


def test_an_unmodified_run_writes_the_same_trajectory(tmp_path):
    from helpers import OracleLJCalculator, make_factory
    outs = []
    for k, patch in enumerate((False, True)):
        if patch:
            assert 'mcpy.ensembles.base_ensemble' in xyz.install()
        try:
            ens = make_factory(OracleLJCalculator(rc=2.5, sigma=1.0, epsilon=1.0), L=7.0, n0=30, base_seed=77)(T=2.0, rank=0)
            ens._traj_file = str(tmp_path / f'traj{k}.xyz')
            ens.run(25)
            outs.append(open(ens._traj_file).read())
        finally:
            ref_be.write_xyz = REF_WRITE
    assert outs[0] == outs[1] and outs[0].count('energy=') >= 25
