"""The CPython marshalling helper (csrc/marshal.cpp) against the numpy marshalling it replaces (Engine._row_pointers):
same pointer / size tables, cells and pbcs for ase.Atoms-like and plain attribute objects, -1 for anything that
cannot be read in place.  No GPU needed."""
import numpy as np
import pytest

from mcpy_b200 import build as mb

mb.build_marshal()
from mcpy_b200 import _marshal                      # noqa: E402
from mcpy_b200.engine import Engine                 # noqa: E402


def _tables(R):
    return (np.zeros(R, np.uintp), np.zeros(R, np.uintp), np.zeros(R, np.int64), np.zeros((R, 9)),
            np.zeros((R, 3), np.uint8))


class _Plain:
    pass


def _plain(n, seed, zdt=np.int64):
    rng = np.random.default_rng(seed)
    a = _Plain()
    a.positions = rng.random((n, 3))
    a.numbers = rng.integers(1, 80, n).astype(zdt)
    a.cell = rng.random((3, 3))
    a.pbc = np.array([True, seed % 2 == 0, False])
    return a


def test_rows_match_numpy_marshalling_for_atoms():
    from ase import Atoms
    rng = np.random.default_rng(0)
    batch = [Atoms('Ar%d' % (5 + k), positions=rng.random((5 + k, 3)) * 9, cell=[9, 9 + k, 9], pbc=[True, k % 2 == 0, True])
             for k in range(7)]
    pp, zp, n, cells, pbcs = _tables(7)
    is64 = _marshal.rows(batch, pp, zp, n, cells, pbcs)
    a, b, want64, n2 = Engine._row_pointers([x.positions for x in batch], [x.numbers for x in batch])
    assert is64 == want64
    assert np.array_equal(pp, a) and np.array_equal(zp, b) and np.array_equal(n, n2)
    assert np.array_equal(cells, np.stack([np.asarray(x.cell).reshape(9) for x in batch]))
    assert np.array_equal(pbcs, np.stack([x.pbc for x in batch]).astype(np.uint8))


@pytest.mark.parametrize('zdt', [np.int32, np.int64])
def test_rows_plain_objects_and_empty_configuration(zdt):
    batch = [_plain(4, 1, zdt), _plain(0, 2, zdt), _plain(9, 3, zdt)]
    pp, zp, n, cells, pbcs = _tables(3)
    assert _marshal.rows(batch, pp, zp, n, cells, pbcs) == (1 if zdt is np.int64 else 0)
    assert list(n) == [4, 0, 9] and pp[1] == 0 and zp[1] == 0
    assert pp[0] == batch[0].positions.ctypes.data and zp[2] == batch[2].numbers.ctypes.data
    assert np.array_equal(cells[2], batch[2].cell.reshape(9)) and list(pbcs[1]) == [1, 1, 0]
    assert _marshal.rows([], pp, zp, n, cells, pbcs) == 0


def test_rows_refuses_what_it_cannot_read_in_place():
    pp, zp, n, cells, pbcs = _tables(2)
    ok = _plain(4, 1)
    for spoil in ('f32', 'strided', 'mixed_numbers', 'pbc_scalar', 'no_cell', 'short_numbers'):
        bad = _plain(4, 2)
        if spoil == 'f32':
            bad.positions = bad.positions.astype(np.float32)
        elif spoil == 'strided':
            bad.positions = np.zeros((8, 3))[::2]
        elif spoil == 'mixed_numbers':
            bad.numbers = bad.numbers.astype(np.int32)
        elif spoil == 'pbc_scalar':
            bad.pbc = True
        elif spoil == 'no_cell':
            del bad.cell
        else:
            bad.numbers = bad.numbers[:3]
        assert _marshal.rows([ok, bad], pp, zp, n, cells, pbcs) == -1, spoil
    with pytest.raises(ValueError):
        _marshal.rows([ok, ok], pp[:1], zp, n, cells, pbcs)
    with pytest.raises(ValueError):
        _marshal.rows([ok], pp, zp, n.astype(np.int32), cells, pbcs)
