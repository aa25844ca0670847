"""GPU: the incremental drop-in path (mcpyb_tracked_energy_host) against the ORACLE and against the engine's own
full path on a chain of GCMC-like configurations: displacements, insertions, deletions, accepted moves that
accumulate, ULP-level perturbation of every coordinate (atoms.wrap()), box changes."""
import numpy as np
import pytest

from mcpy_b200 import Engine
from mcpy_b200.utils import synthetic
from oracle import cfast


def _oracle_energy(potential, cfg, pos, num, cell, pbc):
    if potential == 'lj':
        return cfast.lj_energy(pos, cell, pbc, cfg.get('sigma', 1.0), cfg.get('epsilon', 1.0), cfg.get('rc', 3.0))
    return cfast.emt_energy(pos, num, cell, pbc)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('potential', ['lj', 'emt'])
def test_tracked_chain_equals_full_evaluations(potential):
    rng = np.random.default_rng(17)
    if potential == 'lj':
        cfg = synthetic.s1_argon()
        setup = lambda e: e.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
        new_Z, step = 18, 0.8
    else:
        cfg = synthetic.s3_ag_o()
        setup = lambda e: e.set_emt()
        new_Z, step = 8, 0.3
    trk, full = Engine(0), Engine(0)
    setup(trk); setup(full)
    cell, pbc = cfg['cell'], cfg['pbc']
    pos, num = cfg['positions'].copy(), cfg['numbers'].copy()
    paths = [0, 0, 0]
    for it in range(160):
        tp, tn = pos.copy(), num.copy()
        kind = rng.integers(4)
        if kind == 0:                                  # displacement
            i = rng.integers(len(tp)); tp[i] += rng.uniform(-step, step, 3)
        elif kind == 1:                                # insertion (appended, like atoms += Atoms(...))
            p = tp[rng.integers(len(tp))] + rng.uniform(-3, 3, 3)
            tp = np.concatenate([tp, [p]]); tn = np.concatenate([tn, [new_Z]]).astype(np.int32)
        elif kind == 2 and len(tp) > 5:                # deletion (row removed, like del atoms[i])
            i = rng.integers(len(tp)); tp = np.delete(tp, i, axis=0); tn = np.delete(tn, i)
        else:                                          # two atoms at once (rigid pair move)
            i = rng.choice(len(tp), 2, replace=False); tp[i] += rng.uniform(-step, step, 3)
        e_t, path = trk.tracked_energy(tp, tn, cell, pbc)
        paths[path] += 1
        e_f = full.potential_energy(tp, tn, cell, pbc)
        assert abs(e_t - e_f) <= 1e-10 * max(1.0, abs(e_f)), (it, kind, path, e_t, e_f)
        e_o = _oracle_energy(potential, cfg, tp, tn, cell, pbc)            # the third leg: the CPU oracle
        assert abs(e_t - e_o) <= 1e-10 * max(1.0, abs(e_o)), (it, kind, path, e_t, e_o)
        if rng.random() < 0.5:                         # accepted: the trial becomes the state ...
            pos, num = tp, tn
            if potential == 'lj':                      # ... and atoms.wrap() perturbs every coordinate
                L = cell[0, 0]
                pos = (np.linalg.solve(cell.T, pos.T).T % 1.0) @ cell
        # the same configuration asked twice in a row costs nothing
        if it % 37 == 0:
            e2, p2 = trk.tracked_energy(tp, tn, cell, pbc)
            assert e2 == e_t and p2 in (2, 1, 0)
    assert paths[1] > 100 and paths[0] >= 2          # mostly incremental, re-based a few times
    # a different box invalidates the base
    cell2 = cell.copy()
    if potential == 'lj':
        cell2[0, 0] *= 1.01
        e_t, path = trk.tracked_energy(pos, num, cell2, pbc)
        assert path == 0 and abs(e_t - full.potential_energy(pos, num, cell2, pbc)) <= 1e-10 * max(1.0, abs(e_t))
    trk.close(); full.close()


@pytest.mark.parametrize('potential,zdt', [('lj', np.int32), ('lj', np.int64), ('emt', np.int64)])
def test_tracked_batch_equals_full_evaluations(potential, zdt):
    """get_potential_energies through resident bases: per-replica moves that accumulate, replicas
    swapping slots (replica exchange), ragged sizes, a subset of the replicas, a re-base.  The configurations
    are read where they are (mcpyb_tracked_energies_rows_host; int32 or ASE's int64 atomic numbers); every
    seventh call hands over a non-contiguous array, which takes the concatenating entry point."""
    rng = np.random.default_rng(23)
    R = 6
    if potential == 'lj':
        cfgs = [synthetic.s2_lj(300 + 7 * r, 3.0, 0.5, seed=50 + r) for r in range(R)]
        setup = lambda e: e.set_lj(1.0, 1.0, 3.0)
        new_Z, step = 1, 0.4
    else:
        cfgs = [synthetic.s3_ag_o() for _ in range(R)]
        for r, c in enumerate(cfgs):
            c['positions'] = c['positions'] + rng.normal(0, 0.02, c['positions'].shape)
        setup = lambda e: e.set_emt()
        new_Z, step = 8, 0.3
    trk, full = Engine(0), Engine(0)
    setup(trk); setup(full)
    pos = [c['positions'].copy() for c in cfgs]
    num = [np.asarray(c['numbers'], zdt).copy() for c in cfgs]
    cells = np.stack([c['cell'].reshape(9) for c in cfgs])
    pbcs = np.stack([c['pbc'] for c in cfgs])
    paths = np.zeros(3, int)
    for it in range(60):
        tp, tn = [p.copy() for p in pos], [z.copy() for z in num]
        for r in range(R):                                  # one trial move per replica
            kind = rng.integers(3)
            if kind == 0:
                i = rng.integers(len(tp[r])); tp[r][i] += rng.uniform(-step, step, 3)
            elif kind == 1:
                p = tp[r][rng.integers(len(tp[r]))] + rng.uniform(-2.5, 2.5, 3)
                tp[r] = np.concatenate([tp[r], [p]]); tn[r] = np.concatenate([tn[r], [new_Z]]).astype(zdt)
            elif len(tp[r]) > 5:
                i = rng.integers(len(tp[r])); tp[r] = np.delete(tp[r], i, axis=0); tn[r] = np.delete(tn[r], i)
        sel = list(range(R)) if it % 11 else [0, 2, 3]      # sometimes only a subset is evaluated
        given = [tp[r] for r in sel]
        if it % 7 == 6:                                     # a strided view: not readable in place
            wide = np.zeros((len(given[0]), 6)); wide[:, ::2] = given[0]; given[0] = wide[:, ::2]
            assert Engine._row_pointers(given, [tn[r] for r in sel]) is None
        else:
            assert Engine._row_pointers(given, [tn[r] for r in sel]) is not None
        e_t, pth = trk.tracked_energies(given, [tn[r] for r in sel], cells[sel], pbcs[sel])
        e_f = full.potential_energies([tp[r] for r in sel], [tn[r] for r in sel], cells[sel], pbcs[sel])
        assert np.all(np.abs(e_t - e_f) <= 1e-10 * np.maximum(1.0, np.abs(e_f))), (it, e_t - e_f)
        e_o = np.array([_oracle_energy(potential, {'rc': 3.0}, tp[r], tn[r], cells[k].reshape(3, 3), pbcs[k])
                        for r, k in zip(sel, sel)])
        assert np.all(np.abs(e_t - e_o) <= 1e-10 * np.maximum(1.0, np.abs(e_o))), (it, e_t - e_o)
        for x in pth:
            paths[x] += 1
        for r in sel:                                       # accept about half of the moves
            if rng.random() < 0.5:
                pos[r], num[r] = tp[r], tn[r]
        if it % 9 == 8:                                     # replica exchange: two slots swap configurations
            i, j = rng.choice(R, 2, replace=False)
            pos[i], pos[j] = pos[j], pos[i]
            num[i], num[j] = num[j], num[i]
            if potential == 'lj':                           # boxes differ per replica: they travel with the configuration
                cells[[i, j]] = cells[[j, i]]
    assert paths[1] > 150 and paths[0] >= R                 # mostly incremental, re-based at least once
    trk.close(); full.close()


@pytest.mark.parametrize('potential', ['lj', 'emt'])
def test_resident_chain_answers_like_the_launch_path(potential):
    """The resident chain CTA (mailbox in pinned memory, no launch per call) and the one-launch-per-call path give
    the same bits for the same chain of configurations; the CTA survives accepted moves that accumulate, leaves
    when the chain pauses and comes back, and is replaced when the base is rebuilt."""
    import time
    cfg = synthetic.s2_lj() if potential == 'lj' else synthetic.s3_ag_o()
    new_Z = 1 if potential == 'lj' else 8
    cell, pbc = cfg['cell'], cfg['pbc']
    rng = np.random.default_rng(5)
    steps = []
    pos, num = cfg['positions'].copy(), cfg['numbers'].copy()
    for it in range(120):
        tp, tn = pos.copy(), num.copy()
        kind = it % 3
        if kind == 0:
            i = rng.integers(len(tp)); tp[i] += rng.uniform(-0.3, 0.3, 3)
        elif kind == 1:
            p_new = rng.random(3) @ cell if potential == 'lj' else tp[rng.integers(len(tp))] + rng.uniform(-2.5, 2.5, 3)
            tp = np.concatenate([tp, [p_new]]); tn = np.concatenate([tn, [new_Z]]).astype(np.int32)
        else:
            i = rng.integers(len(tp)); tp = np.delete(tp, i, axis=0); tn = np.delete(tn, i)
        steps.append((tp, tn))
        if rng.random() < 0.3:
            pos, num = tp, tn
    out = {}
    for mode in (True, False):
        eng = Engine(0)
        if potential == 'lj':
            eng.set_lj(1.0, 1.0, 3.0)
        else:
            eng.set_emt()
        eng.resident_chain(mode)
        res = []
        for k, (tp, tn) in enumerate(steps):
            if k in (40, 41):
                time.sleep(0.01)                      # longer than the grid's idle timeout: it leaves and is relaunched
            res.append(eng.tracked_energy(tp, tn, cell, pbc))
        out[mode] = (np.array([r[0] for r in res]), [r[1] for r in res], eng.resident_chain())
        eng.close()
    e_on, p_on, (launches, trials) = out[True]
    e_off, p_off, (l_off, t_off) = out[False]
    assert p_on == p_off and p_on.count(1) > 90
    assert np.array_equal(e_on, e_off)
    # resident: most incremental calls went through the mailbox (not the first one after a re-base -- the stencil
    # lists appear with the second -- nor differences of more than 12 sites), with far fewer launches than trials
    assert trials >= 0.6 * p_on.count(1) and 3 <= launches < 20, (trials, launches, p_on.count(1))
    assert (l_off, t_off) == (0, 0)
    for k in (0, 17, 63, 119):
        tp, tn = steps[k]
        ref = _oracle_energy(potential, {'rc': 3.0}, tp, tn, cell, pbc)
        assert abs(e_on[k] - ref) <= 1e-10 * max(1.0, abs(ref))
