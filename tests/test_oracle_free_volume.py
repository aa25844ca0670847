"""CPU: free-volume oracle -- cKDTree (the reference's library) == brute force, reference values."""
import math

import numpy as np
import pytest

from oracle import cfast, free_volume as ofv


def test_strict_boundary_kdtree_equals_bruteforce():
    rng = np.random.default_rng(12)
    atoms = rng.uniform(-5, 5, size=(30, 3))
    radii = rng.choice([1.3, 2.0, 0.0], size=30)
    pts = []
    for a, r in zip(atoms, radii):
        if r <= 0:
            continue
        u = rng.standard_normal((100, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        p = a + u * r
        pts += [p, np.nextafter(p, a), np.nextafter(p, p + (p - a))]
    pts = np.concatenate(pts)
    kd = ofv.covered_mask_kdtree(pts, atoms, radii)
    assert np.array_equal(kd, ofv.covered_mask_bruteforce(pts, atoms, radii))
    assert np.array_equal(kd, cfast.fv_covered(pts, atoms, radii))


def test_reference_periodic_image_value():
    # tests/test_audit_regressions.py:363-372: |V - 966.49| < 3 with 200 000 points, seed 7
    rng = np.random.default_rng(7)
    dims = np.diag([10.0, 10.0, 10.0])
    vol, ncov = ofv.custom_free_volume(rng, [[0.0, 5.0, 5.0]], [2.0], dims, np.zeros(3), dims,
                                       [True, True, True], 200_000)
    assert vol == pytest.approx(1000.0 - (4.0 / 3.0) * math.pi * 8.0, abs=3.0)


def test_floor_and_empty():
    # tests/test_zero_free_volume.py:22-63: a fully covered region floors at V / P
    rng = np.random.default_rng(1)
    vol, nfree = ofv.spherical_free_volume(rng, [[0, 0, 0]], [100.0], 5.0, 1000)
    assert nfree == 0 and vol == (4.0 / 3.0) * np.pi * 125.0 / 1000
    vol, nfree = ofv.spherical_free_volume(rng, np.zeros((0, 3)), [], 5.0, 1000)
    assert vol == (4.0 / 3.0) * np.pi * 125.0


def test_sphere_sampler_consumes_the_stream_like_the_reference():
    # spherical_cell.py:93-106: oversample int(1.5 n) + 1 candidates per pass, 3 doubles each
    rng = np.random.default_rng(5)
    pts, draws = ofv.sample_sphere_points(rng, 1000, 7.5)
    assert draws % (3 * 1501) == 0 and np.all(np.einsum('ij,ij->i', pts, pts) <= 7.5 ** 2)
    rng2 = np.random.default_rng(5)
    rng2.random(draws)
    assert rng.random() == rng2.random()
