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


def _supported_particle():
    # the fixture of the reference's tests/test_dome_cell.py:17-40: Al support layer, off-centre Ag cluster
    from ase import Atoms
    ag = np.array([[14.0, 10.0, 0.5], [14.0, 10.0, 5.0], [16.0, 10.0, 2.5], [12.0, 10.0, 2.5],
                   [14.0, 12.0, 2.5], [14.0, 8.0, 2.5]])
    xs = np.linspace(1.0, 19.0, 7)
    support = [(x, y, 0.0) for x in xs for y in xs]
    return Atoms(['Al'] * len(support) + ['Ag'] * len(ag), positions=support + [tuple(p) for p in ag],
                 cell=[20, 20, 20], pbc=True), ag


def test_dome_oracle_equals_the_reference_class():
    """oracle.dome_free_volume == the unmodified mcpy.cell.DomeCell (same seed, same stream)."""
    import os, sys
    if not os.path.isdir('/root/reference/mcpy'):
        pytest.skip('reference checkout not present (GPU box)')
    sys.path.insert(0, '/root/reference')
    try:
        from mcpy.cell import DomeCell
    finally:
        sys.path.remove('/root/reference')
    atoms, ag = _supported_particle()
    radii_map = {'Al': 1.0, 'Ag': 1.0, 'O': 0.0}
    ref = DomeCell(atoms, particle_species='Ag', bottom_z=0.0, vacuum=1.0, species_radii=radii_map,
                   mc_sample_points=20_000, seed=4)
    rng = np.random.default_rng(4)
    radii = np.array([radii_map[s] for s in atoms.get_chemical_symbols()])
    for _ in range(2):
        ref.calculate_volume(atoms)
        vol, n_free, n_above = ofv.dome_free_volume(rng, atoms.positions, radii, ref.radius, ref.center,
                                                    0.0, 20_000)
        assert vol == ref.get_volume() and 0 < n_free < n_above < 20_000
    assert rng.bit_generator.state == ref._rng.bit_generator.state
    # no atoms: the dome fraction of the ball (dome_cell.py:118-121)
    from ase import Atoms
    ref.calculate_volume(Atoms())
    vol, n_free, n_above = ofv.dome_free_volume(rng, np.zeros((0, 3)), [], ref.radius, ref.center, 0.0, 20_000)
    assert vol == ref.get_volume() and n_free == n_above
