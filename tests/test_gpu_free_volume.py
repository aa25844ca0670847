"""GPU parity: free-volume counts are bit-exact integers vs scipy cKDTree (the
reference's own library, mcpy/cell/spherical_cell.py:130-138)."""
import math

import numpy as np
import pytest

from mcpy_b200.utils import synthetic
from oracle import free_volume as ofv

pytestmark = pytest.mark.gpu


def test_sphere_counts_match_kdtree_s3(fv_engine):
    cfg = synthetic.s3_ag_o()
    rng = np.random.default_rng(3)
    pts, _ = ofv.sample_sphere_points(rng, 100_000, cfg['sphere_radius'])
    covered = ofv.covered_mask_kdtree(pts, cfg['positions'], cfg['radii'])
    n_free, n_cov, mask = fv_engine.free_volume_count(pts, cfg['positions'], cfg['radii'], return_mask=True)
    assert n_cov == int(covered.sum()) and n_free == int((~covered).sum())
    assert np.array_equal(mask, covered)


def test_boundary_points_are_strict(fv_engine):
    # points within a few ulps of sphere surfaces: cKDTree's d < r is strict
    rng = np.random.default_rng(12)
    atoms = rng.uniform(-5, 5, size=(40, 3))
    radii = rng.choice([1.3, 2.0, 0.0], size=40)
    pts = []
    for a, r in zip(atoms, radii):
        if r <= 0:
            continue
        u = rng.standard_normal((200, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        p = a + u * r
        for k in range(-3, 4):
            q = p.copy()
            for _ in range(abs(k)):
                q = np.nextafter(q, a if k < 0 else q + (q - a))
            pts.append(q)
    pts = np.concatenate(pts)
    covered = ofv.covered_mask_kdtree(pts, atoms, radii)
    brute = ofv.covered_mask_bruteforce(pts, atoms, radii)
    assert np.array_equal(covered, brute)
    n_free, n_cov, mask = fv_engine.free_volume_count(pts, atoms, radii, return_mask=True)
    assert np.array_equal(mask, covered)
    assert n_cov == int(covered.sum())


def test_custom_cell_periodic_images_reference_value(fv_engine):
    # tests/test_audit_regressions.py:363-372: one r = 2 sphere on the x face of a 10 A cube,
    # 200 000 points, seed 7 -> |V - 966.49| < 3
    from mcpy_b200.cell import B200CustomCell
    from ase import Atoms
    atoms = Atoms('H', positions=[[0.0, 5.0, 5.0]], cell=[10, 10, 10], pbc=True)
    cell = B200CustomCell(atoms, custom_height=10.0, bottom_z=0.0, species_radii={'H': 2.0},
                          mc_sample_points=200_000, seed=7, engine=fv_engine)
    cell.calculate_volume(atoms)
    expected = 1000.0 - (4.0 / 3.0) * math.pi * 8.0
    assert cell.get_volume() == pytest.approx(expected, abs=3.0)
    rng = np.random.default_rng(7)
    vol, n_cov = ofv.custom_free_volume(rng, atoms.positions, [2.0], cell.dimensions, cell.offset,
                                        cell.original_dimensions, atoms.pbc, 200_000)
    assert cell.last_counts[1] == n_cov
    assert cell.get_volume() == vol


def test_custom_cell_slab_images_and_offset(fv_engine):
    rng = np.random.default_rng(5)
    dims = np.array([[11.0, 0, 0], [5.5, 9.5, 0], [0, 0, 30.0]])
    pos = rng.random((60, 3)) @ dims
    radii = rng.choice([1.5, 2.2, 0.0], size=60)
    region = dims.copy(); region[2, 2] = 12.0
    offset = np.array([0, 0, 6.0])
    pbc = [True, True, False]
    prng = np.random.default_rng(9)
    pts = ofv.sample_box_points(prng, 50_000, region)
    images = ofv.periodic_images(pos, offset, dims, pbc)
    rad = np.tile(radii, len(images) // len(pos))
    covered = ofv.covered_mask_kdtree(pts, images, rad)
    shifts = np.array([np.array([i, j, 0], float) @ dims for i in (-1, 0, 1) for j in (-1, 0, 1)])
    n_free, n_cov, mask = fv_engine.free_volume_count(pts, pos, radii, offset=offset, shifts=shifts,
                                                      return_mask=True)
    assert np.array_equal(mask, covered)


def test_s5_million_points(fv_engine):
    cfg = synthetic.s5_cupd_co()
    rng = np.random.default_rng(5)
    pts, _ = ofv.sample_sphere_points(rng, 1_000_000, cfg['sphere_radius'])
    covered = ofv.covered_mask_kdtree(pts, cfg['positions'], cfg['radii'])
    n_free, n_cov = fv_engine.free_volume_count(pts, cfg['positions'], cfg['radii'])
    assert (n_free, n_cov) == (int((~covered).sum()), int(covered.sum()))


def test_spherical_cell_mirror_matches_oracle_and_floor(fv_engine):
    from mcpy_b200.cell import B200SphericalCell
    from ase import Atoms
    cfg = synthetic.s3_ag_o()
    syms = ['Ag' if z == 47 else 'O' for z in cfg['numbers']]
    atoms = Atoms(syms, positions=cfg['positions'])
    cell = B200SphericalCell(atoms, vacuum=3.0, species_radii={'Ag': 2.947, 'O': 0.0},
                             mc_sample_points=20_000, seed=3, engine=fv_engine)
    cell.calculate_volume(atoms)
    rng = np.random.default_rng(3)
    vol, n_free = ofv.spherical_free_volume(rng, atoms.positions, cfg['radii'], cell.radius, 20_000)
    assert cell.last_counts[0] == n_free
    assert cell.get_volume() == vol
    # zero free volume -> floored at one point's worth (tests/test_zero_free_volume.py:22-51)
    full = B200SphericalCell(atoms, vacuum=0.0, species_radii={'Ag': 50.0, 'O': 50.0},
                             mc_sample_points=1000, seed=1, engine=fv_engine)
    full.calculate_volume(atoms)
    assert full.get_volume() == full.sphere_volume / 1000
    # empty atoms -> whole sphere
    empty = Atoms()
    full.calculate_volume(empty)
    assert full.get_volume() == full.sphere_volume


# ---- K8: sample points regenerated on the GPU from numpy's PCG64 state -------------------

@pytest.mark.parametrize('n_points', [1, 1000, 100_000])
def test_device_sphere_points_match_numpy_stream(fv_engine, n_points):
    cfg = synthetic.s3_ag_o()
    rng_dev = np.random.default_rng(3)
    rng_ref = np.random.default_rng(3)
    for _ in range(2):                       # two consecutive refreshes share one stream
        n_free, n_cov = fv_engine.sphere_free_volume_pcg64(rng_dev, n_points, cfg['sphere_radius'],
                                                           np.zeros(3), cfg['positions'], cfg['radii'])
        pts, draws = ofv.sample_sphere_points(rng_ref, n_points, cfg['sphere_radius'])
        covered = ofv.covered_mask_kdtree(pts, cfg['positions'], cfg['radii'])
        assert (n_free, n_cov) == (int((~covered).sum()), int(covered.sum()))
        # the host generator was advanced by exactly what the reference sampler consumes
        assert rng_dev.bit_generator.state == rng_ref.bit_generator.state


def test_device_box_points_match_numpy_stream(fv_engine):
    rng_pos = np.random.default_rng(5)
    for dims in (np.diag([11.0, 9.5, 12.0]),
                 np.array([[11.0, 0, 0], [5.5, 9.5, 0], [0, 0, 12.0]])):     # cubic and hexagonal
        original = dims.copy(); original[2, 2] = 30.0
        pos = rng_pos.random((60, 3)) @ original
        radii = rng_pos.choice([1.5, 2.2, 0.0], size=60)
        offset = np.array([0, 0, 6.0])
        pbc = [True, True, False]
        shifts = np.array([np.array([i, j, 0], float) @ original for i in (-1, 0, 1) for j in (-1, 0, 1)])
        rng_dev = np.random.default_rng(9)
        rng_ref = np.random.default_rng(9)
        n_free, n_cov = fv_engine.box_free_volume_pcg64(rng_dev, 50_000, dims, pos, radii, offset, shifts)
        pts = ofv.sample_box_points(rng_ref, 50_000, dims)
        images = ofv.periodic_images(pos, offset, original, pbc)
        covered = ofv.covered_mask_kdtree(pts, images, np.tile(radii, 9))
        assert (n_free, n_cov) == (int((~covered).sum()), int(covered.sum()))
        assert rng_dev.bit_generator.state == rng_ref.bit_generator.state


def test_cells_in_device_point_mode_track_the_reference_stream(fv_engine):
    from mcpy_b200.cell import B200SphericalCell
    from ase import Atoms
    cfg = synthetic.s3_ag_o()
    syms = ['Ag' if z == 47 else 'O' for z in cfg['numbers']]
    a1 = Atoms(syms, positions=cfg['positions'])
    a2 = Atoms(syms, positions=cfg['positions'])
    kw = dict(vacuum=3.0, species_radii={'Ag': 2.947, 'O': 0.0}, mc_sample_points=30_000, seed=11,
              engine=fv_engine)
    host = B200SphericalCell(a1, points='host', **kw)
    dev = B200SphericalCell(a2, points='device', **kw)
    for _ in range(3):
        host.calculate_volume(a1)
        dev.calculate_volume(a2)
        assert host.last_counts == dev.last_counts and host.get_volume() == dev.get_volume()
        # insertion points drawn afterwards come from the same stream position
        assert np.array_equal(host.get_random_point(), dev.get_random_point())


def test_dome_cell_matches_oracle_in_both_point_modes(fv_engine):
    """B200DomeCell (dome_cell.py:8-138): same integer counts, same volume, same stream position as
    the oracle restatement (itself pinned to the unmodified DomeCell in test_oracle_free_volume.py)."""
    from mcpy_b200.cell import B200DomeCell
    from ase import Atoms
    ag = np.array([[14.0, 10.0, 0.5], [14.0, 10.0, 5.0], [16.0, 10.0, 2.5], [12.0, 10.0, 2.5],
                   [14.0, 12.0, 2.5], [14.0, 8.0, 2.5]])
    xs = np.linspace(1.0, 19.0, 7)
    support = [(x, y, 0.0) for x in xs for y in xs]
    atoms = Atoms(['Al'] * len(support) + ['Ag'] * len(ag), positions=support + [tuple(p) for p in ag],
                  cell=[20, 20, 20], pbc=True)
    radii_map = {'Al': 1.0, 'Ag': 1.0, 'O': 0.0}
    radii = np.array([radii_map[s] for s in atoms.get_chemical_symbols()])
    before = atoms.positions.copy()
    for mode in ('host', 'device'):
        cell = B200DomeCell(atoms, particle_species='Ag', bottom_z=0.0, vacuum=1.0, species_radii=radii_map,
                            mc_sample_points=50_000, seed=4, engine=fv_engine, points=mode)
        assert np.array_equal(atoms.positions, before)            # the dome does not translate the atoms
        np.testing.assert_allclose(cell.center, ag.mean(axis=0), atol=1e-12)
        rng = np.random.default_rng(4)
        for _ in range(2):
            cell.calculate_volume(atoms)
            vol, n_free, n_above = ofv.dome_free_volume(rng, atoms.positions, radii, cell.radius, cell.center,
                                                        0.0, 50_000)
            assert cell.last_counts[0] == n_free and sum(cell.last_counts) == n_above
            assert cell.get_volume() == vol
        cell.calculate_volume(Atoms())                            # empty: dome fraction of the ball
        vol, n_free, n_above = ofv.dome_free_volume(rng, np.zeros((0, 3)), [], cell.radius, cell.center, 0.0, 50_000)
        assert cell.get_volume() == vol and cell.last_counts == (n_above, 0)
        assert rng.bit_generator.state == cell._rng.bit_generator.state
        p = cell.get_random_point()
        assert p[2] >= 0.0 and cell.is_point_inside(p)
        below = cell.center.copy(); below[2] = -0.5
        assert not cell.is_point_inside(below)
        trial = atoms.copy()
        trial += Atoms('O', positions=[p])
        trial += Atoms('O', positions=[below])
        assert list(cell.get_atoms_specie_inside_cell(trial, ['O'])) == [len(atoms)]


def test_region_mask_kernel_equals_the_reference_cells_bit_for_bit():
    """K10: ``species AND inside`` of the unmodified ``SphericalCell`` / ``DomeCell.get_atoms_specie_inside_cell`` against
    the device mask, atoms placed on the sphere's surface to the ulp, below / on / above the dome's floor, three species."""
    from ase import Atoms
    from mcpy.cell import DomeCell, SphericalCell
    from mcpy_b200.cell import B200DomeCell, B200SphericalCell
    rng = np.random.default_rng(11)
    n = 12000
    core = Atoms('Ag40', positions=rng.uniform(8.0, 14.0, (40, 3)), cell=[40.0, 40.0, 40.0], pbc=False)
    radii = {'Ag': 1.4, 'O': 0.7, 'C': 0.8}
    ref = SphericalCell(core.copy(), 3.0, radii, mc_sample_points=10, seed=1)      # (construction translates its atoms in place)
    mine = B200SphericalCell(core.copy(), 3.0, radii, mc_sample_points=10, seed=1)
    assert ref.radius == mine.radius and np.array_equal(ref.center, mine.center)
    c, R = np.asarray(ref.center, dtype=float), float(ref.radius)
    pos = c + rng.uniform(-1.3 * R, 1.3 * R, (n, 3))
    u = rng.standard_normal((3000, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    pos[:3000] = c + u * R                                            # on the surface, to rounding
    for k in range(1, 4):                                             # ... and a few ulps to either side
        pos[3000 * 0 + 500 * k:3000 * 0 + 500 * k + 250] = c + u[500 * k:500 * k + 250] * np.nextafter(R, np.inf)
        pos[500 * k + 250:500 * k + 500] = c + u[500 * k + 250:500 * k + 500] * np.nextafter(R, 0.0)
    numbers = rng.choice([47, 8, 6, 1], size=n)
    atoms = Atoms(numbers=numbers, positions=pos, cell=[40.0, 40.0, 40.0], pbc=False)
    assert len(atoms) >= mine.mask_on_device_from
    for specie in ('O', ['O', 'C'], ['Ag', 'O', 'C'], 'Xx', []):
        a, b = ref.get_atoms_specie_inside_cell(atoms, specie), mine.get_atoms_specie_inside_cell(atoms, specie)
        assert a.dtype.kind == b.dtype.kind == 'i' and np.array_equal(a, b), specie
    assert 2000 < len(ref.get_atoms_specie_inside_cell(atoms, ['Ag', 'O', 'C', 'H'])) < n
    # the numpy path of the same class (small systems) gives the same indices as the device path
    small = mine.mask_on_device_from
    try:
        type(mine).mask_on_device_from = 10 ** 9
        assert np.array_equal(mine.get_atoms_specie_inside_cell(atoms, ['O', 'C']), ref.get_atoms_specie_inside_cell(atoms, ['O', 'C']))
    finally:
        type(mine).mask_on_device_from = small
    dref = DomeCell(core.copy(), 'Ag', 10.5, 3.0, radii, mc_sample_points=10, seed=1)
    dmine = B200DomeCell(core.copy(), 'Ag', 10.5, 3.0, radii, mc_sample_points=10, seed=1)
    assert np.array_equal(dref.center, dmine.center) and dref.radius == dmine.radius
    cd, Rd = np.asarray(dref.center, dtype=float), float(dref.radius)
    pos2 = cd + rng.uniform(-1.2 * Rd, 1.2 * Rd, (n, 3))
    pos2[:2000] = cd + u[:2000] * Rd                                  # on the ball's surface
    pos2[4000:4400, 2] = dref.bottom_z                                # on the floor
    pos2[4400:4800, 2] = np.nextafter(dref.bottom_z, -np.inf)         # one ulp below it
    atoms2 = Atoms(numbers=numbers, positions=pos2, cell=[40.0, 40.0, 40.0], pbc=False)
    assert 500 < len(dref.get_atoms_specie_inside_cell(atoms2, ['Ag', 'O', 'C', 'H'])) < n
    for specie in ('O', ['Ag', 'O', 'C']):
        assert np.array_equal(dref.get_atoms_specie_inside_cell(atoms2, specie), dmine.get_atoms_specie_inside_cell(atoms2, specie))
