"""Seeded synthetic configurations S1..S5 of SURVEY.md section 8d.

Plain numpy (no ASE): each builder returns a dict with ``positions`` (N,3) f64,
``numbers`` (N,) int32, ``cell`` (3,3), ``pbc`` (3,) bool and the potential
parameters, so bench.py, the tests and the CPU oracle all see identical inputs.
"""
from __future__ import annotations

import numpy as np


def jittered_lattice(n, L, seed, jitter=0.1):
    """First ``n`` sites of the smallest simple-cubic lattice that holds them,
    each displaced uniformly by +-``jitter`` lattice spacings."""
    m = int(np.ceil(n ** (1.0 / 3.0) - 1e-9))
    a = L / m
    rng = np.random.default_rng(seed)
    idx = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing='ij'),
                   axis=-1).reshape(-1, 3)[:n]
    pos = (idx + 0.5) * a + rng.uniform(-jitter * a, jitter * a, size=(n, 3))
    return np.ascontiguousarray(pos)


def s1_argon(n=500, seed=1):
    """Argon LJ, periodic cubic box, rho = 0.0225 A^-3 (BASELINE config 0)."""
    L = (n / 0.0225) ** (1.0 / 3.0)
    return dict(name='S1-argon', positions=jittered_lattice(n, L, seed),
                numbers=np.full(n, 18, np.int32), cell=np.eye(3) * L, pbc=np.ones(3, bool),
                potential='lj', epsilon=0.0104, sigma=3.40, rc=10.2, temperature=87.0,
                units='metal', max_displacement=0.3 * 3.40)


def s2_lj(n=2000, rc=3.0, rho=0.6, seed=2):
    """Reduced-unit LJ fluid, N = 2000, rho* = 0.6 (BASELINE config 1).  ``rc=3.0`` is
    the headline attractive variant; ``rc=2**(1/6)`` the repulsive (WCA) one."""
    L = (n / rho) ** (1.0 / 3.0)
    return dict(name=f'S2-lj-rc{rc:.3f}', positions=jittered_lattice(n, L, seed),
                numbers=np.ones(n, np.int32), cell=np.eye(3) * L, pbc=np.ones(3, bool),
                potential='lj', epsilon=1.0, sigma=1.0, rc=float(rc), temperature=2.0,
                units='LJ', max_displacement=0.3)


def fcc_octahedron(length, cutoff, a):
    """fcc octahedron with ``length`` atoms on an edge, ``cutoff`` layers cut at
    each vertex (the geometry of ``ase.cluster.Octahedron``)."""
    m = length - 1
    g = np.arange(-m, m + 1)
    i, j, k = [x.ravel() for x in np.meshgrid(g, g, g, indexing='ij')]
    keep = ((np.abs(i) + np.abs(j) + np.abs(k)) <= m) & (((i + j + k) - m) % 2 == 0)
    lim = m - cutoff
    keep &= (np.abs(i) <= lim) & (np.abs(j) <= lim) & (np.abs(k) <= lim)
    return np.stack([i[keep], j[keep], k[keep]], axis=1) * (a / 2.0)


def _random_points_in_sphere(rng, n, radius):
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return d * (radius * rng.random(n) ** (1.0 / 3.0))[:, None]


def s3_ag_o(n_o=20, seed=3, min_insert=0.5, vacuum=3.0):
    """Ag Octahedron(6,1) (140 atoms) + ``n_o`` O, EMT, SphericalCell (BASELINE config 2)."""
    ag = fcc_octahedron(6, 1, 4.09)
    ag = ag - ag.mean(axis=0)
    radius = float(np.linalg.norm(ag, axis=1).max() + vacuum)
    rng = np.random.default_rng(seed)
    pos = [p for p in ag]
    while len(pos) < len(ag) + n_o:
        p = _random_points_in_sphere(rng, 1, radius)[0]
        if np.min(np.linalg.norm(np.array(pos) - p, axis=1)) >= min_insert:
            pos.append(p)
    numbers = np.array([47] * len(ag) + [8] * n_o, np.int32)
    return dict(name='S3-Ag140-O', positions=np.ascontiguousarray(pos), numbers=numbers,
                cell=np.zeros((3, 3)), pbc=np.zeros(3, bool), potential='emt',
                sphere_radius=radius, species_radii={'Ag': 2.947, 'O': 0.0},
                radii=np.where(numbers == 47, 2.947, 0.0), mc_sample_points=100_000,
                temperature=300.0, units='metal')


def s5_cupd_co(length=25, n_co=200, seed=5, min_insert=1.3, vacuum=3.5, d_co=1.128):
    """~10k-atom CuPd octahedron + CO, EMT, SphericalCell, 1e6 points (BASELINE config 4)."""
    cu = fcc_octahedron(length, 0, 3.61)
    cu = cu - cu.mean(axis=0)
    rng = np.random.default_rng(seed)
    numbers = np.full(len(cu), 29, np.int32)
    numbers[rng.random(len(cu)) < 0.05] = 46
    radius = float(np.linalg.norm(cu, axis=1).max() + vacuum)
    from scipy.spatial import cKDTree
    tree = cKDTree(cu)
    mol = []
    while len(mol) < n_co:
        c = _random_points_in_sphere(rng, 1, radius)[0]
        u = rng.standard_normal(3)
        u /= np.linalg.norm(u)
        pc, po = c - u * d_co * 0.5714, c + u * d_co * 0.4286     # C and O about the COM
        ok = tree.query(pc)[0] >= min_insert and tree.query(po)[0] >= min_insert
        for q in mol:
            ok = ok and min(np.linalg.norm(q[0] - pc), np.linalg.norm(q[1] - po),
                            np.linalg.norm(q[0] - po), np.linalg.norm(q[1] - pc)) >= min_insert
        if ok:
            mol.append((pc, po))
    pos = np.concatenate([cu] + [np.array(m) for m in mol])
    numbers = np.concatenate([numbers, np.tile(np.array([6, 8], np.int32), n_co)])
    radii = np.select([numbers == 29, numbers == 46], [2.4, 2.5], 0.0)
    return dict(name='S5-CuPd-CO', positions=np.ascontiguousarray(pos), numbers=numbers,
                cell=np.zeros((3, 3)), pbc=np.zeros(3, bool), potential='emt',
                sphere_radius=radius, radii=radii, mc_sample_points=1_000_000,
                species_radii={'Cu': 2.4, 'Pd': 2.5, 'C': 0.0, 'O': 0.0},
                temperature=300.0, units='metal')


def trial_batch(cfg, n_trials, seed=7, mix=(1, 1, 1), insert_Z=None):
    """``n_trials`` pre-generated atomic proposals against ``cfg``: insertions
    (uniform point in the box / sphere), deletions (uniform atom), displacements
    (uniform in a ball of ``max_displacement``), in the ratio ``mix``.

    Returns the ragged arrays of the C ABI: ``old_off, old_idx, new_off, new_pos, new_Z``
    plus ``kind`` (0 insert, 1 delete, 2 displace)."""
    rng = np.random.default_rng(seed)
    n = len(cfg['positions'])
    kinds = rng.choice(3, size=n_trials, p=np.array(mix, float) / np.sum(mix))
    if insert_Z is None:
        insert_Z = int(cfg['numbers'][-1]) if n else 1
    periodic = bool(np.all(cfg['pbc']))
    old_off = np.zeros(n_trials + 1, np.int32)
    new_off = np.zeros(n_trials + 1, np.int32)
    old_idx, new_pos, new_Z = [], [], []
    maxd = cfg.get('max_displacement', 0.3)
    for t, k in enumerate(kinds):
        if k == 0:
            if periodic:
                p = rng.random(3) @ cfg['cell']
            else:
                p = _random_points_in_sphere(rng, 1, cfg['sphere_radius'])[0]
            new_pos.append(p); new_Z.append(insert_Z)
        elif k == 1:
            old_idx.append(int(rng.integers(n)))
        else:
            i = int(rng.integers(n))
            while True:
                d = 2.0 * rng.random(3) - 1.0
                if d @ d <= 1.0:
                    break
            old_idx.append(i)
            new_pos.append(cfg['positions'][i] + d * maxd)
            new_Z.append(int(cfg['numbers'][i]))
        old_off[t + 1] = len(old_idx)
        new_off[t + 1] = len(new_pos)
    return dict(kind=kinds.astype(np.int32), old_off=old_off,
                old_idx=np.array(old_idx, np.int32), new_off=new_off,
                new_pos=np.array(new_pos, np.float64).reshape(-1, 3),
                new_Z=np.array(new_Z, np.int32))


def apply_trial(cfg, tb, t):
    """Positions / numbers of the configuration after trial ``t`` (for the oracle)."""
    pos, num = cfg['positions'], cfg['numbers']
    old = tb['old_idx'][tb['old_off'][t]:tb['old_off'][t + 1]]
    ns, ne = tb['new_off'][t], tb['new_off'][t + 1]
    keep = np.ones(len(pos), bool)
    keep[old] = False
    return (np.concatenate([pos[keep], tb['new_pos'][ns:ne]]),
            np.concatenate([num[keep], tb['new_Z'][ns:ne]]))


def trial_batch_periodic(cfg, n_trials, seed=7, mix=(1, 1, 1)):
    """Vectorised :func:`trial_batch` for a periodic box and atomic moves (no Python loop: the bench draws
    64 x 65 536 proposals per rank).  Same layout and the same three kinds; a different random stream."""
    rng = np.random.default_rng(seed)
    n = len(cfg['positions'])
    kinds = rng.choice(3, size=n_trials, p=np.array(mix, float) / np.sum(mix)).astype(np.int32)
    has_old = kinds != 0
    has_new = kinds != 1
    old_off = np.zeros(n_trials + 1, np.int32)
    new_off = np.zeros(n_trials + 1, np.int32)
    np.cumsum(has_old, out=old_off[1:])
    np.cumsum(has_new, out=new_off[1:])
    atom = rng.integers(n, size=n_trials)
    old_idx = atom[has_old].astype(np.int32)
    d = 2.0 * rng.random((4 * n_trials, 3)) - 1.0                   # uniform in the unit ball by rejection
    d = d[np.einsum('ij,ij->i', d, d) <= 1.0][:n_trials]
    ins = rng.random((n_trials, 3)) @ cfg['cell']
    disp = cfg['positions'][atom] + d * cfg.get('max_displacement', 0.3)
    new_pos = np.where((kinds == 0)[:, None], ins, disp)[has_new]
    insert_Z = int(cfg['numbers'][-1]) if n else 1
    new_Z = np.where(kinds == 0, insert_Z, cfg['numbers'][atom])[has_new].astype(np.int32)
    return dict(kind=kinds, old_off=old_off, old_idx=old_idx, new_off=new_off,
                new_pos=np.ascontiguousarray(new_pos, dtype=np.float64), new_Z=new_Z)


def sphere_points(rng, n, radius):
    """``n`` uniform points in the ball by cube rejection (the sampling scheme of ``SphericalCell``; bench input)."""
    out = np.empty((n, 3))
    filled = 0
    while filled < n:
        pts = (rng.random((int(1.5 * n) + 1, 3)) - 0.5) * 2.0 * radius
        keep = pts[np.einsum('ij,ij->i', pts, pts) <= radius * radius]
        take = min(len(keep), n - filled)
        out[filled:filled + take] = keep[:take]
        filled += take
    return out
