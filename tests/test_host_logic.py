"""CPU: host-side mirrors of the reference interface (chunking, argument errors, synthetic inputs)."""
import numpy as np
import pytest

from mcpy_b200.utils import synthetic
from mcpy_b200.utils.chunking import chunk_ranges, shard_range


def test_chunk_ranges_matches_reference_contract():
    # tests/test_chunking.py of the reference, same expectations
    assert list(chunk_ranges(5, None)) == [(0, 5)]
    assert list(chunk_ranges(3, 10)) == [(0, 3)]
    assert list(chunk_ranges(6, 2)) == [(0, 2), (2, 4), (4, 6)]
    assert list(chunk_ranges(7, 3)) == [(0, 3), (3, 6), (6, 7)]
    assert list(chunk_ranges(3, 1)) == [(0, 1), (1, 2), (2, 3)]
    assert list(chunk_ranges(0, 2)) == []
    for bad in (0, -1):
        assert list(chunk_ranges(4, bad)) == [(0, 4)]


def test_shard_range_partitions_replicas():
    for n, w in [(64, 8), (10, 4), (3, 8), (0, 2)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(64, 3, 8) == (24, 32)
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_synthetic_configs_are_deterministic_and_sized():
    a, b = synthetic.s2_lj(), synthetic.s2_lj()
    assert np.array_equal(a['positions'], b['positions']) and len(a['positions']) == 2000
    assert abs(len(a['positions']) / np.linalg.det(a['cell']) - 0.6) < 1e-12
    s1 = synthetic.s1_argon()
    assert len(s1['positions']) == 500 and abs(500 / np.linalg.det(s1['cell']) - 0.0225) < 1e-12
    s3 = synthetic.s3_ag_o()
    assert (s3['numbers'] == 47).sum() == 140 and (s3['numbers'] == 8).sum() == 20
    tb = synthetic.trial_batch(a, 100, seed=3)
    assert tb['old_off'][-1] == len(tb['old_idx']) and tb['new_off'][-1] == len(tb['new_pos'])
    kinds = tb['kind']
    ko, kn = np.diff(tb['old_off']), np.diff(tb['new_off'])
    assert np.all(ko[kinds == 0] == 0) and np.all(kn[kinds == 0] == 1)
    assert np.all(ko[kinds == 1] == 1) and np.all(kn[kinds == 1] == 0)
    assert np.all(ko[kinds == 2] == 1) and np.all(kn[kinds == 2] == 1)
    p1, z1 = synthetic.apply_trial(a, tb, int(np.argmax(kinds == 0)))
    assert len(p1) == 2001 and len(z1) == 2001
