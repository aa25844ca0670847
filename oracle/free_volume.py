"""Free-volume / sample-point oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``SphericalCell.calculate_volume`` / ``_sample_sphere_points``
(``mcpy/cell/spherical_cell.py:88-142``), ``CustomCell.calculate_volume`` /
``_periodic_images`` (``mcpy/cell/custom_cell.py:61-107``) and
``BaseCell._clamp_free_volume`` (``mcpy/cell/base_cell.py:46-60``) with the
same libraries the reference uses (numpy PCG64 ``default_rng``,
``scipy.spatial.cKDTree``), plus a brute-force restatement of the covered
test in the association order the CUDA kernel must reproduce bit for bit:
``(dx*dx + dy*dy) + dz*dz < r*r`` in unfused fp64 (strict).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def sample_sphere_points(rng, n, radius, center=(0.0, 0.0, 0.0)):
    """``SphericalCell._sample_sphere_points`` (``spherical_cell.py:88-106``):
    cube rejection from the PCG64 stream, first ``n`` accepted in stream
    order.  Returns ``(points, draws)`` where ``draws`` is the number of
    doubles consumed from ``rng`` (3 per candidate)."""
    out = np.empty((n, 3), dtype=float)
    filled = 0
    radius_sq = radius ** 2
    oversample = int(1.5 * n) + 1
    draws = 0
    while filled < n:
        pts = (rng.random((oversample, 3)) - 0.5) * 2.0 * radius
        draws += 3 * oversample
        d2 = np.einsum('ij,ij->i', pts, pts)
        keep = pts[d2 <= radius_sq]
        take = min(len(keep), n - filled)
        out[filled:filled + take] = keep[:take]
        filled += take
    return out + np.asarray(center, dtype=float), draws


def sample_box_points(rng, n, dimensions):
    """``CustomCell.calculate_volume`` points (``custom_cell.py:74-75``):
    ``rng.random((n, 3)) @ dimensions`` in the cell frame."""
    return rng.random((n, 3)) @ np.asarray(dimensions, dtype=float)


def periodic_images(positions, offset, original_dimensions, pbc):
    """``CustomCell._periodic_images`` (``custom_cell.py:93-107``)."""
    positions = np.asarray(positions, dtype=float) - np.asarray(offset, dtype=float)
    pbc = np.asarray(pbc, dtype=bool)
    if not pbc.any():
        return positions
    shift_ranges = [(-1, 0, 1) if p else (0,) for p in pbc]
    images = [
        positions + np.array([i, j, k], dtype=float) @ np.asarray(original_dimensions)
        for i in shift_ranges[0]
        for j in shift_ranges[1]
        for k in shift_ranges[2]
    ]
    return np.concatenate(images)


def covered_mask_kdtree(points, positions, radii):
    """The reference's covered test (``spherical_cell.py:127-138``): one kd-tree
    per unique positive radius, ``query(k=1, distance_upper_bound=r)``."""
    points = np.asarray(points, dtype=float)
    positions = np.asarray(positions, dtype=float)
    radii = np.asarray(radii, dtype=float)
    covered = np.zeros(len(points), dtype=bool)
    for r in np.unique(radii):
        if r <= 0.0:
            continue
        mask = radii == r
        tree = cKDTree(positions[mask])
        dists, _ = tree.query(points, k=1, distance_upper_bound=float(r))
        covered |= np.isfinite(dists)
    return covered


def covered_mask_bruteforce(points, positions, radii, chunk=4096):
    """Brute-force covered test in the kernel's association order."""
    points = np.asarray(points, dtype=float)
    positions = np.asarray(positions, dtype=float)
    radii = np.asarray(radii, dtype=float)
    sel = radii > 0.0
    pos, rad = positions[sel], radii[sel]
    covered = np.zeros(len(points), dtype=bool)
    if len(pos) == 0:
        return covered
    rr = rad * rad
    for s in range(0, len(points), chunk):
        p = points[s:s + chunk]
        dx = p[:, None, 0] - pos[None, :, 0]
        dy = p[:, None, 1] - pos[None, :, 1]
        dz = p[:, None, 2] - pos[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        covered[s:s + chunk] = (d2 < rr[None, :]).any(axis=1)
    return covered


def clamp_free_volume(free_volume, region_volume, n_points):
    """``BaseCell._clamp_free_volume`` (``base_cell.py:46-60``) without the log."""
    floor = region_volume / n_points
    return floor if free_volume < floor else free_volume


def spherical_free_volume(rng, positions, radii, radius, n_points,
                          center=(0.0, 0.0, 0.0), brute=False):
    """``SphericalCell.calculate_volume`` (``spherical_cell.py:108-142``).
    Returns ``(volume, n_free)``; advances ``rng`` exactly as the reference."""
    sphere_volume = (4.0 / 3.0) * np.pi * (radius ** 3)
    if len(positions) == 0:
        return sphere_volume, n_points
    pts, _ = sample_sphere_points(rng, n_points, radius, center)
    fn = covered_mask_bruteforce if brute else covered_mask_kdtree
    covered = fn(pts, positions, radii)
    n_free = int(np.count_nonzero(~covered))
    free_fraction = float(n_free) / n_points
    return clamp_free_volume(free_fraction * sphere_volume, sphere_volume, n_points), n_free


def custom_free_volume(rng, positions, radii, dimensions, offset,
                       original_dimensions, pbc, n_points, brute=False):
    """``CustomCell.calculate_volume`` (``custom_cell.py:61-91``).
    Returns ``(volume, n_covered)``."""
    cell_volume = float(abs(np.linalg.det(np.asarray(dimensions, dtype=float))))
    n_atoms = len(positions)
    if n_atoms == 0:
        return cell_volume, 0
    cart = sample_box_points(rng, n_points, dimensions)
    images = periodic_images(positions, offset, original_dimensions, pbc)
    rad = np.tile(np.asarray(radii, dtype=float), len(images) // n_atoms)
    fn = covered_mask_bruteforce if brute else covered_mask_kdtree
    covered = fn(cart, images, rad)
    n_cov = int(np.count_nonzero(covered))
    occupied_fraction = float(n_cov) / n_points
    return clamp_free_volume(cell_volume * (1.0 - occupied_fraction),
                             cell_volume, n_points), n_cov


def dome_free_volume(rng, positions, radii, radius, center, bottom_z, n_points, brute=False):
    """``DomeCell.calculate_volume`` (``dome_cell.py:106-138``): full-ball sample, points below
    ``bottom_z`` dropped, free fraction normalised by the FULL-ball point count (``:136``).
    Returns ``(volume, n_free_above, n_above)``; advances ``rng`` exactly as the reference."""
    sphere_volume = (4.0 / 3.0) * np.pi * (radius ** 3)
    pts, _ = sample_sphere_points(rng, n_points, radius, center)
    pts = pts[pts[:, 2] >= bottom_z]
    n_above = len(pts)
    if len(positions) == 0:
        return (float(n_above) / n_points) * sphere_volume, n_above, n_above
    fn = covered_mask_bruteforce if brute else covered_mask_kdtree
    covered = fn(pts, positions, radii)
    n_free = int(np.count_nonzero(~covered))
    free_fraction = float(n_free) / n_points
    return clamp_free_volume(free_fraction * sphere_volume, sphere_volume, n_points), n_free, n_above
