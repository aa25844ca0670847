"""SphericalCell / CustomCell with the Monte-Carlo free volume counted on a B200.

Same duck-typed surface as the reference cells (``mcpy/cell/base_cell.py:19-76``,
``cell.py:27-97``): ``calculate_volume(atoms)``, ``get_volume()``,
``get_random_point()``, ``get_atoms_specie_inside_cell(atoms, species)``,
``get_species()``, ``is_point_inside``, ``is_point_exchangeable``, attributes
``volume`` / ``mc_sample_points``.  The host parts (region predicates, insertion
points, the numpy PCG64 stream ``np.random.default_rng(seed)`` of ``cell.py:25``)
consume random numbers exactly as the reference does, so a run seeded like the
reference proposes the same points.  Only the covered test -- the cKDTree loops
of ``spherical_cell.py:127-140`` and ``custom_cell.py:80-89`` -- moves to the GPU,
and its result is the same integer.

``points='host'`` draws the sample points with numpy and uploads them;
``points='device'`` regenerates the identical PCG64 stream on the GPU
(see csrc/pcg64.cuh) and advances the host generator by the same amount.
"""
from __future__ import annotations

import logging

import numpy as np

from ..engine import Engine

logger = logging.getLogger(__name__)


def _symbols(atoms):
    return atoms.get_chemical_symbols()


class _B200CellBase:
    def __init__(self, atoms, species_radii, seed, engine, device, points):
        if points not in ('host', 'device'):
            raise ValueError("points must be 'host' or 'device'")
        self.original_dimensions = np.array(atoms.cell, dtype=float).reshape(3, 3)
        self.dimensions = self.original_dimensions
        self.species_radii = species_radii if species_radii else {}
        self._box_volume = float(abs(np.linalg.det(self.dimensions)))
        self.volume = self._box_volume
        self._rng = np.random.default_rng(seed)
        self.engine = engine if engine is not None else Engine(device)
        self.points = points
        self.last_counts = None           # (free, covered) of the last device count

    # -- shared with mcpy.cell.Cell -------------------------------------------
    def get_volume(self):
        return self.volume

    def get_species(self):
        return list(self.species_radii.keys())

    def _radius_table(self):
        """radius by atomic number (NaN = no radius given), rebuilt when ``species_radii`` changes."""
        key = tuple(sorted(self.species_radii.items()))
        if getattr(self, '_radius_key', None) != key:
            from ase.data import atomic_numbers
            table = np.full(max(len(atomic_numbers), 120), np.nan)
            for sym, r in self.species_radii.items():
                table[atomic_numbers[sym]] = float(r)
            self._radius_key, self._radius_by_Z = key, table
        return self._radius_by_Z

    def _radii_for(self, atoms):
        """Per-atom exclusion radii (``cell.py:69-89``), ValueError naming any species without a
        radius.  The reference walks the list of chemical symbols in Python (milliseconds per call
        for 10 000 atoms, several times the GPU work it feeds); here one table lookup by atomic number."""
        radii = self._radius_table()[np.asarray(atoms.numbers)]
        if np.isnan(radii).any():
            missing = sorted(set(_symbols(atoms)) - set(self.species_radii))
            raise ValueError(
                f'{type(self).__name__}.species_radii has no radius for {missing}. '
                f'Every species the system can contain needs one, including species '
                f'inserted during the run; got {sorted(self.species_radii)}.')
        return radii

    def _species_mask(self, atoms, specie):
        """``np.isin(atoms.get_chemical_symbols(), specie)`` by atomic number."""
        from ase.data import atomic_numbers
        zs = [atomic_numbers[s] for s in self._species_list(specie) if s in atomic_numbers]
        return np.isin(np.asarray(atoms.numbers), zs)

    def _clamp_free_volume(self, free_volume, region_volume):
        """Floor at one sample point's worth of the region (``base_cell.py:46-60``)."""
        floor = region_volume / self.mc_sample_points
        if free_volume < floor:
            logger.warning(
                'Estimated free volume %.3g A^3 is below the sampler resolution '
                '(%d points over %.3g A^3); flooring to %.3g A^3. The region is '
                'effectively full -- consider enlarging it.',
                free_volume, self.mc_sample_points, region_volume, floor)
            return floor
        return free_volume

    def is_point_exchangeable(self, point):
        return self.is_point_inside(point)

    @staticmethod
    def _species_list(specie):
        return [specie] if isinstance(specie, str) else list(specie)


class B200SphericalCell(_B200CellBase):
    """Drop-in for ``mcpy.cell.SphericalCell`` (``spherical_cell.py:7-142``).

    Like the reference, construction TRANSLATES ``atoms`` so that their centre
    of mass sits at the origin.
    """

    def __init__(self, atoms, vacuum, species_radii, mc_sample_points=100_000, seed=None,
                 engine=None, device=0, points='host'):
        super().__init__(atoms, species_radii, seed, engine, device, points)
        self.center = np.zeros(3)
        atoms.translate(self.center - atoms.get_center_of_mass())
        self.radius = float(np.linalg.norm(atoms.positions - self.center, axis=1).max() + vacuum)
        self.species_radii = species_radii
        self.mc_sample_points = int(mc_sample_points)
        self.sphere_volume = (4.0 / 3.0) * np.pi * (self.radius ** 3)
        self.volume = self.sphere_volume

    def is_point_inside(self, point):
        d = point - self.center
        return float(np.dot(d, d)) <= self.radius ** 2

    def get_random_point(self):
        """Cube-root method, 3 normals + 1 uniform from the cell stream
        (``spherical_cell.py:55-63``)."""
        direction = self._rng.standard_normal(3)
        direction /= np.linalg.norm(direction)
        r = self.radius * (self._rng.random() ** (1.0 / 3.0))
        return self.center + r * direction

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        mask = self._species_mask(atoms, specie)
        d = atoms.positions - self.center
        mask &= np.einsum('ij,ij->i', d, d) <= self.radius ** 2
        return np.where(mask)[0]

    def _sample_sphere_points(self, n):
        """Host restatement of the cube-rejection sampler (``spherical_cell.py:88-106``):
        first ``n`` accepted candidates of the PCG64 stream, in stream order."""
        out = np.empty((n, 3), dtype=float)
        filled = 0
        radius_sq = self.radius ** 2
        oversample = int(1.5 * n) + 1
        while filled < n:
            pts = (self._rng.random((oversample, 3)) - 0.5) * 2.0 * self.radius
            keep = pts[np.einsum('ij,ij->i', pts, pts) <= radius_sq]
            take = min(len(keep), n - filled)
            out[filled:filled + take] = keep[:take]
            filled += take
        return out + self.center

    def calculate_volume(self, atoms):
        """Free volume = (#uncovered points / P) * sphere volume, floored
        (``spherical_cell.py:108-142``)."""
        if len(atoms) == 0:
            self.volume = self.sphere_volume
            return
        radii = self._radii_for(atoms)
        if self.points == 'device':
            n_free, n_cov = self.engine.sphere_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.radius, self.center,
                atoms.positions, radii)
        else:
            pts = self._sample_sphere_points(self.mc_sample_points)
            n_free, n_cov = self.engine.free_volume_count(pts, atoms.positions, radii)
        self.last_counts = (n_free, n_cov)
        free_fraction = float(n_free) / self.mc_sample_points
        self.volume = self._clamp_free_volume(free_fraction * self.sphere_volume,
                                              self.sphere_volume)


class B200DomeCell(B200SphericalCell):
    """Drop-in for ``mcpy.cell.DomeCell`` (``dome_cell.py:8-138``): the ball around the centroid of
    the particle species, clipped at the support surface ``z >= bottom_z``.  Unlike the spherical
    cell it does NOT translate the atoms."""

    def __init__(self, atoms, particle_species, bottom_z, vacuum, species_radii,
                 mc_sample_points: int = 100_000, seed=None, engine=None, device=0, points='host'):
        _B200CellBase.__init__(self, atoms, species_radii, seed, engine, device, points)
        from ase.data import atomic_numbers
        zs = [atomic_numbers[s] for s in self._species_list(particle_species) if s in atomic_numbers]
        particle_mask = np.isin(np.asarray(atoms.numbers), zs)
        if not particle_mask.any():
            raise ValueError(
                f'No atoms of particle_species={self._species_list(particle_species)} found in the cell.')
        particle_pos = atoms.positions[particle_mask]
        self.center = particle_pos.mean(axis=0)
        self.bottom_z = float(bottom_z)
        self.radius = float(np.linalg.norm(particle_pos - self.center, axis=1).max() + vacuum)
        self.species_radii = species_radii
        self.mc_sample_points = int(mc_sample_points)
        self.sphere_volume = (4.0 / 3.0) * np.pi * (self.radius ** 3)
        self.volume = self.sphere_volume

    def is_point_inside(self, point):
        if point[2] < self.bottom_z:
            return False
        d = point - self.center
        return float(np.dot(d, d)) <= self.radius ** 2

    def get_random_point(self):
        """Full-ball cube-root sampler with rejection below the surface (``dome_cell.py:65-77``)."""
        while True:
            direction = self._rng.standard_normal(3)
            direction /= np.linalg.norm(direction)
            r = self.radius * (self._rng.random() ** (1.0 / 3.0))
            point = self.center + r * direction
            if point[2] >= self.bottom_z:
                return point

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        mask = self._species_mask(atoms, specie)
        d = atoms.positions - self.center
        mask &= np.einsum('ij,ij->i', d, d) <= self.radius ** 2
        mask &= atoms.positions[:, 2] >= self.bottom_z
        return np.where(mask)[0]

    def calculate_volume(self, atoms):
        """Free dome volume = (#uncovered points above the surface / P of the FULL ball) * sphere
        volume, floored (``dome_cell.py:106-138``).  Without atoms: the dome fraction of the ball."""
        n_atoms = len(atoms)
        radii = self._radii_for(atoms) if n_atoms else np.empty(0)
        if self.points == 'device':
            n_free, n_cov = self.engine.dome_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.radius, self.center, self.bottom_z,
                atoms.positions.reshape(-1, 3), radii)
        else:
            pts = self._sample_sphere_points(self.mc_sample_points)
            pts = pts[pts[:, 2] >= self.bottom_z]
            if n_atoms:
                n_free, n_cov = self.engine.free_volume_count(pts, atoms.positions, radii)
            else:
                n_free, n_cov = len(pts), 0
        self.last_counts = (n_free, n_cov)
        if n_atoms == 0:
            self.volume = (float(n_free) / self.mc_sample_points) * self.sphere_volume
            return
        free_fraction = float(n_free) / self.mc_sample_points
        self.volume = self._clamp_free_volume(free_fraction * self.sphere_volume, self.sphere_volume)


class B200CustomCell(_B200CellBase):
    """Drop-in for ``mcpy.cell.CustomCell`` (``custom_cell.py:7-162``)."""

    def __init__(self, atoms, custom_height=None, bottom_z=None, species_radii=None,
                 mc_sample_points: int = 100_000, seed=None, engine=None, device=0,
                 points='host'):
        super().__init__(atoms, species_radii, seed, engine, device, points)
        self.dimensions, self.offset = self.get_custom_height_cell(custom_height, bottom_z)
        self.cell_volume = float(abs(np.linalg.det(self.dimensions)))
        self.mc_sample_points = int(mc_sample_points)
        self._dim_inv = np.linalg.inv(self.dimensions)
        self.volume = self.cell_volume

    def get_custom_height_cell(self, custom_height, bottom_z):
        """Region = original box with its c-height replaced (``custom_cell.py:29-52``)."""
        if custom_height is None or bottom_z is None:
            raise ValueError(
                'CustomCell requires explicit custom_height and bottom_z '
                f'(got custom_height={custom_height}, bottom_z={bottom_z})')
        top = self.original_dimensions[2][2]
        if custom_height > top:
            raise ValueError('Custom height cannot be greater than the original cell height.')
        if bottom_z < 0 or bottom_z + custom_height > top:
            raise ValueError('Custom cell exceeds the bounds of the original cell.')
        dims = np.copy(self.original_dimensions)
        dims[2][2] = custom_height
        offset = np.zeros(3)
        offset[2] = bottom_z
        return dims, offset

    def get_random_point(self):
        return self._rng.random(3) @ self.dimensions + self.offset

    def _frac(self, p):
        return (p - self.offset) @ self._dim_inv

    def is_point_inside(self, point):
        f = self._frac(point)
        return bool(np.all((f >= 0.0) & (f < 1.0)))

    def is_point_exchangeable(self, point):
        """xy inside the footprint, z at or above the floor; upper z bound dropped
        (``custom_cell.py:142-156``)."""
        f = self._frac(point)
        return bool(np.all(f[:2] >= 0.0) and np.all(f[:2] < 1.0) and f[2] >= 0.0)

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        mask = self._species_mask(atoms, specie)
        f = self._frac(atoms.positions)
        mask &= np.all((f[:, :2] >= 0.0) & (f[:, :2] < 1.0), axis=1) & (f[:, 2] >= 0.0)
        return np.where(mask)[0]

    def _image_shifts(self, atoms):
        """Lattice shifts of the -1/0/+1 images along the periodic axes of the
        ORIGINAL box, in the reference's loop order (``custom_cell.py:99-106``)."""
        pbc = np.asarray(atoms.pbc, dtype=bool)
        if not pbc.any():
            return np.zeros((1, 3))
        rng3 = [(-1, 0, 1) if p else (0,) for p in pbc]
        return np.array([np.array([i, j, k], dtype=float) @ self.original_dimensions
                         for i in rng3[0] for j in rng3[1] for k in rng3[2]])

    def calculate_volume(self, atoms):
        """Free volume = V_cell * (1 - covered / P), floored (``custom_cell.py:61-91``)."""
        if len(atoms) == 0:
            self.volume = self.cell_volume
            return
        radii = self._radii_for(atoms)
        shifts = self._image_shifts(atoms)
        if self.points == 'device':
            n_free, n_cov = self.engine.box_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.dimensions, atoms.positions, radii,
                self.offset, shifts)
        else:
            cart = self._rng.random((self.mc_sample_points, 3)) @ self.dimensions
            n_free, n_cov = self.engine.free_volume_count(cart, atoms.positions, radii,
                                                          offset=self.offset, shifts=shifts)
        self.last_counts = (n_free, n_cov)
        occupied_fraction = float(n_cov) / self.mc_sample_points
        self.volume = self._clamp_free_volume(self.cell_volume * (1.0 - occupied_fraction),
                                              self.cell_volume)
