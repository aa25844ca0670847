"""CPU: the PCG64 restatement in csrc/pcg64.cuh against numpy's own generator, and the
floating-point association numpy uses in the sphere acceptance test."""
import numpy as np

from mcpy_b200 import _lib
from mcpy_b200.engine import Engine


def _doubles(rng, skip, n):
    lib = _lib.load()
    state, inc = Engine._pcg_words(rng)
    out = np.empty(n)
    assert lib.mcpyb_pcg64_host_doubles(state.ctypes.data, inc.ctypes.data, skip, n, out.ctypes.data) == 0
    return out


def test_stream_matches_numpy_including_jump_ahead():
    for seed in (0, 3, 7, 12345):
        ref = np.random.default_rng(seed).random(5000)
        rng = np.random.default_rng(seed)
        assert np.array_equal(_doubles(rng, 0, 5000), ref)
        for skip in (1, 2, 63, 64, 1000, 4097):
            assert np.array_equal(_doubles(rng, skip, 100), ref[skip:skip + 100])
    # jump far ahead: numpy's own advance is the reference
    a = np.random.default_rng(5)
    b = np.random.default_rng(5)
    b.bit_generator.advance(3 * 1_500_001 + 17)
    assert np.array_equal(_doubles(a, 3 * 1_500_001 + 17, 64), b.random(64))


def test_sphere_acceptance_association_is_xx_plus_zz_plus_yy():
    # mcpy/cell/spherical_cell.py:100: d2 = np.einsum('ij,ij->i', pts, pts).  The device kernel
    # restates numpy's evaluation order (x*x + z*z) + y*y, unfused.
    rng = np.random.default_rng(0)
    p = (rng.random((200_000, 3)) - 0.5) * 2.0 * 11.43
    d2 = np.einsum('ij,ij->i', p, p)
    assert np.array_equal(d2, (p[:, 0] * p[:, 0] + p[:, 2] * p[:, 2]) + p[:, 1] * p[:, 1])
