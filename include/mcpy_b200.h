/* mcpy_b200.h -- C ABI of the B200-native energy-and-move engine for mcpy.
 *
 * This header is the drop-in boundary.  Every entry point
 * takes plain pointers and sizes; there are no torch / numpy / ASE types.  The
 * reference (farrisric/mcpy v1.4.0) is pure Python, so the binding a maintainer
 * adds on the reference side is a ctypes stub -- shown in INTEGRATION.md and
 * shipped as mcpy_b200/_lib.py.
 *
 * Conventions
 *   - All functions return 0 on success or a negative mcpyb_status; the message is
 *     available from mcpyb_last_error(engine).  Nothing falls back to the CPU.
 *   - "_host" entry points take host buffers and include the host<->device
 *     copies (this is the end-to-end path).  "_dev" entry points take device
 *     pointers, enqueue on the engine's stream and do not synchronise.
 *   - positions are fp64, row-major [n][3], Angstrom (ase.Atoms.positions);
 *     numbers are int32 atomic numbers; cell is the row-major 3x3 lattice
 *     (np.asarray(atoms.cell)); pbc is 3 bytes.
 *   - A "batch" is R configurations (replicas) concatenated; atom_off[R+1]
 *     delimits them.  R = 1 is the single-GCMC case.
 */
#ifndef MCPY_B200_H
#define MCPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcpyb_engine mcpyb_engine;

typedef enum {
    MCPYB_OK = 0,
    MCPYB_ERR_CUDA = -1,        /* a CUDA runtime call failed                      */
    MCPYB_ERR_ARG = -2,         /* invalid argument (ValueError on the Python side) */
    MCPYB_ERR_STATE = -3,       /* call sequence error (no potential / no batch)   */
    MCPYB_ERR_SPECIES = -4,     /* atomic number without parameters                */
    MCPYB_ERR_GEOMETRY = -5,    /* atom more than 511 boxes away / singular cell   */
    MCPYB_ERR_CAPACITY = -6     /* internal capacity exceeded                      */
} mcpyb_status;

/* ---- lifetime ---------------------------------------------------------- */
/* Create an engine on CUDA device `device` (fails if there is no such GPU).  */
int mcpyb_create(int device, mcpyb_engine** out);
void mcpyb_destroy(mcpyb_engine* e);
const char* mcpyb_last_error(const mcpyb_engine* e);
const char* mcpyb_version(void);
/* Use an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream).
 * NULL restores the engine's own stream.                                      */
int mcpyb_set_stream(mcpyb_engine* e, void* cuda_stream);
int mcpyb_synchronize(mcpyb_engine* e);

/* ---- potentials --------------------------------------------------------- */
/* ase.calculators.lj.LennardJones(sigma, epsilon, rc, smooth=False); rc <= 0
 * selects ASE's default 3*sigma.  (Reference call sites:
 * examples/gcmc_molecule.py:27-31, benchmark/lammps_gcmc_parity.py:197.)      */
int mcpyb_set_lj(mcpyb_engine* e, double epsilon, double sigma, double rc);
/* ase.calculators.emt.EMT() with the default parameter table, asap_cutoff=False.
 * (Reference call sites: tests/test_canonical_nvt.py:26,
 * docs/examples/gcmc_simulations.rst:31.)                                      */
int mcpyb_set_emt(mcpyb_engine* e);
/* Arithmetic of the batched LJ trial-move row sums (mcpyb_trial_delta_e_*, and the energies of large batches that
 * run through the same pipeline).  0 (default): fp64 throughout -- energies and dE within 1e-10 relative of
 * ase.calculators.lj.LennardJones.  1: fp32 pair math (r^2, reciprocal, c6 on the FP32 pipe from an fp64 distance
 * vector; sums accumulated in fp64) -- within 1e-5 relative; membership r^2 <= rc^2 stays bit-exact (candidates
 * within 1e-6 of the cutoff are decided and evaluated in fp64).  The one-call-per-trial paths stay fp64.        */
int mcpyb_set_precision(mcpyb_engine* e, int fp32_pair_math);
/* Neighbour cutoff in use: LJ rc, EMT rc_list.                               */
double mcpyb_cutoff(const mcpyb_engine* e);
/* Derived EMT constants for atomic number Z: out9 = E0,s0,V0,eta2,kappa,lambda,
 * n0,gamma1,gamma2 (tests compare them with the oracle's).                    */
int mcpyb_emt_species(const mcpyb_engine* e, int Z, double* out9);

/* ---- total energies: replaces calculator.get_potential_energy(atoms)
 *      (mcpy/ensembles/base_ensemble.py:190-191) and
 *      calculator.get_potential_energies(atoms_list)
 *      (mcpy/ensembles/batched_replica_exchange.py:202,278) ------------------ */
/* Upload R configurations from host memory, build their cell lists (resident
 * in HBM afterwards).                                                         */
int mcpyb_upload_host(mcpyb_engine* e, int R, const int64_t* atom_off,
                      const double* positions, const int32_t* numbers,
                      const double* cells /* [R][9] */, const uint8_t* pbc /* [R][3] */);
/* Same from device memory (positions/numbers device pointers; the small
 * atom_off / cells / pbc arrays stay on the host).                            */
int mcpyb_upload_dev(mcpyb_engine* e, int R, const int64_t* atom_off,
                     const double* d_positions, const int32_t* d_numbers,
                     const double* cells, const uint8_t* pbc);
/* nseg small copies between device-accessible buffers in ONE launch on the engine's stream (48 segments per launch):
 * count[k] elements from src[k] to dst[k], kind[k] = 0: float64 -> float64, 1: int32 -> int32, 2: float64 -> int32,
 * 3: int32 -> float64.  A ladder shard that keeps its replicas in HBM uses it twice per exchange round
 * (BatchedReplicaExchange._attempt_exchanges / _swap_states, batched_replica_exchange.py:310-410): to pack its
 * records (page-locked host memory, read in place) and boundary configurations into the all_gather buffer, and to
 * move the replica blocks into their new slots afterwards -- a dozen framework copies otherwise, ~10 us of host
 * time each.  Sources and destinations must not overlap.                                                        */
int mcpyb_copy_segments_dev(mcpyb_engine* e, int nseg, const uint64_t* src, const uint64_t* dst,
                            const int64_t* count, const uint8_t* kind);
/* Energies of the resident batch -> host array [R] (synchronises).
 * pairs_out (may be NULL) receives the number of unordered in-cutoff pairs.   */
int mcpyb_energies_host(mcpyb_engine* e, double* energies_out, int64_t* pairs_out);
/* Energies of the resident batch -> device array [R], no synchronisation.    */
int mcpyb_energies_dev(mcpyb_engine* e, double* d_energies_out);
/* upload + energies in one call: the whole plug-in call with host buffers.   */
int mcpyb_potential_energies_host(mcpyb_engine* e, int R, const int64_t* atom_off,
                                  const double* positions, const int32_t* numbers,
                                  const double* cells, const uint8_t* pbc,
                                  double* energies_out);
/* Incremental form of the single-configuration call, for the one-call-per-trial pattern of
 * GrandCanonicalEnsemble.do_gcmc_step (grand_canonical_ensemble.py:244-308).  The engine keeps a
 * fully evaluated base configuration resident, aligns the incoming one with it on the host and
 * returns E_base + dE(base -> incoming); it re-bases (full evaluation) when more than a few atoms
 * differ or the box changes.  Same result as mcpyb_potential_energies_host to rounding (one full
 * evaluation plus one delta, no drift).  *path_out (may be NULL): 0 full, 1 incremental,
 * 2 identical to the base.  Any mcpyb_upload_* call resets the tracker.
 * Alignment bound: a row of the incoming configuration whose three coordinates are within 1e-10 A of a
 * base row of the same species counts as that row unchanged, and the answer is the energy with the BASE's
 * coordinates for it: an error of at most |F_i| x 1.7e-10 A per such row (~1e-10 eV for forces of order
 * 1 eV/A; ASE's own wrap() moves atoms by ulps, far below this).  A displacement smaller than 1e-10 A is
 * therefore scored as no move; the trial-move kernels (mcpyb_trial_delta_e_*) have no such threshold.   */
int mcpyb_tracked_energy_host(mcpyb_engine* e, const double* positions, const int32_t* numbers,
                              int n, const double* cell9, const uint8_t* pbc3,
                              double* energy_out, int* path_out);
/* The same call with the atomic numbers as int64 (what ase.Atoms.numbers holds): the caller converts nothing. */
int mcpyb_tracked_energy_z64_host(mcpyb_engine* e, const double* positions, const int64_t* numbers,
                                  int n, const double* cell9, const uint8_t* pbc3,
                                  double* energy_out, int* path_out);

/* Forces of the resident batch, host array [sum N][3] (eV/A): ASE's results['forces'] of
 * LennardJones / EMT.  Replaces the forces the ASE calculators hand to the optimisers behind
 * BaseCalculator.get_potential_energy (mcpy/calculators/base_calculator.py:14-22) and
 * CanonicalEnsemble.relax (mcpy/ensembles/canonical_ensemble.py:115-123).                     */
int mcpyb_forces_host(mcpyb_engine* e, double* forces_out);
/* Upload + energies + forces in one call (Calculator.calculate(atoms, ['energy', 'forces'])). */
int mcpyb_energy_forces_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                             const int32_t* numbers, const double* cells, const uint8_t* pbc,
                             double* energies_out, double* forces_out);

/* Relax-then-energy: BaseCalculator.get_potential_energy (mcpy/calculators/base_calculator.py:14-22) and
 * CanonicalEnsemble.relax (mcpy/ensembles/canonical_ensemble.py:115-123) run ase.optimize.LBFGS on every trial
 * configuration and then read its energy.  This call does both: ASE's LBFGS iteration without line search
 * (H0 = 1/alpha, `memory` curvature pairs, longest atomic step clipped to maxstep, damping; ASE defaults
 * maxstep 0.2, memory 100, alpha 70, damping 1), converged when max_i |f_i| < fmax, at most max_steps steps
 * (0 = single point).  positions [n][3] are relaxed IN PLACE, as the optimiser does to atoms.positions;
 * fixed (may be NULL): n bytes, nonzero = the atom is held (ase.constraints.FixAtoms).
 * energy_out: energy of the final positions; steps_out / fmax_out (may be NULL): steps taken, final max |f|. */
int mcpyb_relax_lbfgs_host(mcpyb_engine* e, double* positions, const int32_t* numbers, int n,
                           const double* cell9, const uint8_t* pbc3, const uint8_t* fixed,
                           double fmax, int max_steps, double maxstep, int memory, double alpha,
                           double damping, double* energy_out, int* steps_out, double* fmax_out);

/* Batched incremental form for get_potential_energies(atoms_list) (BatchedReplicaExchange
 * ._batched_single_move, batched_replica_exchange.py:235-302): R resident bases, one batched delta
 * launch per call; configurations are matched with the base in their own slot first, then with any
 * other (replica exchange swaps configurations between slots); a configuration more than a few atoms
 * away from every base re-bases the whole batch with a full evaluation.  paths_out[r] (may be NULL):
 * 0 full, 1 incremental, 2 identical to its base.  Any mcpyb_upload_* call resets the bases.       */
int mcpyb_tracked_energies_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                                const int32_t* numbers, const double* cells, const uint8_t* pbc,
                                double* energies_out, int* paths_out);

/* The same with the caller's OWN per-configuration arrays, read where they are (no concatenated copy: for
 * 64 x 2 000 atoms building one costs more host time than the rest of the call).  positions_of[r] -> n_atoms[r]
 * rows of 3 doubles (atoms.positions), numbers_of[r] -> n_atoms[r] atomic numbers, int32 or int64
 * (numbers_are_int64 != 0: what ase.Atoms.numbers holds); numbers_of may be NULL for LJ.                       */
int mcpyb_tracked_energies_rows_host(mcpyb_engine* e, int R, const double* const* positions_of,
                                     const void* const* numbers_of, int numbers_are_int64,
                                     const int64_t* n_atoms, const double* cells, const uint8_t* pbc,
                                     double* energies_out, int* paths_out);

/* Per-atom energies of the resident batch (ASE's results['energies']) and, for
 * EMT, the density sums sigma1; either pointer may be NULL. Host arrays [sum N]. */
int mcpyb_atom_energies_host(mcpyb_engine* e, double* e_atom_out, double* sigma1_out);

/* ---- trial moves: delta E = E(trial) - E(resident)
 *      (mcpy/ensembles/grand_canonical_ensemble.py:283-284) ------------------
 * Trial t acts on replica rep[t] (rep may be NULL: replica 0).  It removes the
 * atoms old_idx[old_off[t]..old_off[t+1]) (indices inside the replica) and adds
 * atoms at new_pos[new_off[t]..new_off[t+1]) with atomic numbers new_Z (may be
 * NULL for LJ).  Displacement / rigid-molecule move: both lists, same length;
 * insertion: new only; deletion: old only.  At most 32 atoms per list.
 * pairs_out (may be NULL): in-cutoff (moved atom, neighbour) pairs evaluated.  */
int mcpyb_trial_delta_e_host(mcpyb_engine* e, int T, const int32_t* rep,
                             const int32_t* old_off, const int32_t* old_idx,
                             const int32_t* new_off, const double* new_pos,
                             const int32_t* new_Z,
                             double* dE_out, int32_t* pairs_out);
int mcpyb_trial_delta_e_dev(mcpyb_engine* e, int T, const int32_t* d_rep,
                            const int32_t* d_old_off, const int32_t* d_old_idx,
                            const int32_t* d_new_off, const double* d_new_pos,
                            const int32_t* d_new_Z, int n_old_total, int n_new_total,
                            double* d_dE_out, int32_t* d_pairs_out);

/* ---- page-locked host buffers ------------------------------------------------------------
 * The "_host" entry points accept any host memory.  Arrays that are page-locked (allocated here, or by
 * cudaHostAlloc / cudaHostRegister) are read and written by the copy engine where they are; pageable
 * arrays pass through a pinned staging block inside the engine (one extra host copy each way).
 * mcpyb_trial_delta_e_host takes the direct path when every array of the call is page-locked.
 * Returns NULL when the allocation fails.                                                  */
void* mcpyb_alloc_pinned(size_t bytes);
void mcpyb_free_pinned(void* p);

/* ---- neighbour list (bit-exactness check; oracle/neighbors.py) -------------
 * Lists every image (i, j, S) of replica `rep` of the resident batch with
 * mode 0: r2 < cutoff^2, 1: r2 <= cutoff^2, 2: sqrt(r2) < cutoff.  cutoff must
 * not exceed mcpyb_cutoff().  Returns the number of pairs found in *npairs
 * (may exceed max_pairs: then only max_pairs were written).                   */
int mcpyb_neighbor_pairs_host(mcpyb_engine* e, int rep, double cutoff, int mode,
                              int64_t max_pairs, int32_t* i_out, int32_t* j_out,
                              int32_t* S_out /* [max_pairs][3] */, double* r2_out,
                              int64_t* npairs);

/* ---- free volume: replaces the cKDTree loops of
 *      SphericalCell.calculate_volume (mcpy/cell/spherical_cell.py:127-140) and
 *      CustomCell.calculate_volume (mcpy/cell/custom_cell.py:80-89) -----------
 * Spheres are (positions[a] - offset) + shifts[s] for every atom a with
 * radii[a] > 0 and every image shift s < nshift (custom_cell.py:93-107; pass
 * nshift = 1, shift = 0, offset = 0 for the spherical cell).  A point is covered
 * iff (dx*dx+dy*dy)+dz*dz < r*r for some sphere.  counts_out[0] = free points,
 * counts_out[1] = covered points.  covered_out (may be NULL): per-point mask.  */
int mcpyb_free_volume_count_host(mcpyb_engine* e, const double* points, int64_t P,
                                 const double* positions, const double* radii, int64_t M,
                                 const double* offset3, const double* shifts, int nshift,
                                 int64_t* counts_out, uint8_t* covered_out);
int mcpyb_free_volume_count_dev(mcpyb_engine* e, const double* d_points, int64_t P,
                                const double* d_positions, const double* d_radii, int64_t M,
                                const double* offset3, const double* shifts, int nshift,
                                uint64_t* d_counts2);

/* ---- free volume with the sample points regenerated on the GPU ---------------
 * The reference draws its points from numpy's PCG64 (np.random.default_rng(seed),
 * mcpy/cell/cell.py:25).  These two calls take that generator's 128-bit state and
 * increment (as {high, low} 64-bit words, from Generator.bit_generator.state) and
 * reproduce the same stream on the device, so no points are uploaded.
 * Sphere: SphericalCell._sample_sphere_points (spherical_cell.py:88-106): cube rejection,
 * first P accepted in stream order, + center.  *draws_out = doubles consumed; the host
 * generator must be advanced by that many (bit_generator.advance).
 * Box: CustomCell.calculate_volume (custom_cell.py:74-75): rng.random((P,3)) @ dims,
 * 3 P doubles consumed.                                                        */
int mcpyb_sphere_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo,
                                   const uint64_t* inc_hi_lo, int64_t P, double radius,
                                   const double* center3, const double* positions,
                                   const double* radii, int64_t M,
                                   int64_t* counts_out, int64_t* draws_out);
/* DomeCell.calculate_volume (mcpy/cell/dome_cell.py:106-138): the same full-ball stream as the
 * sphere call, but a point with z < bottom_z is drawn and then ignored.  counts_out[0] = free points
 * at or above bottom_z, counts_out[1] = covered points at or above it; the reference normalises by
 * the full-ball P (dome_cell.py:136).                                               */
int mcpyb_dome_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo,
                                 const uint64_t* inc_hi_lo, int64_t P, double radius,
                                 const double* center3, double bottom_z, const double* positions,
                                 const double* radii, int64_t M,
                                 int64_t* counts_out, int64_t* draws_out);
int mcpyb_box_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo,
                                const uint64_t* inc_hi_lo, int64_t P, const double* dims9,
                                const double* positions, const double* radii, int64_t M,
                                const double* offset3, const double* shifts, int nshift,
                                int64_t* counts_out);

/* K10: `species AND inside` of SphericalCell.get_atoms_specie_inside_cell (mcpy/cell/spherical_cell.py:65-80) and of
 * DomeCell's (mcpy/cell/dome_cell.py: the same with `z >= bottom_z`): mask_out[i] = numbers[i] in species_Z and
 * |positions[i] - center|^2 <= radius_squared (numpy's arithmetic of einsum('ij,ij->i'): (dx^2 + dz^2) + dy^2, unfused)
 * and, with clip_z, positions[i][2] >= bottom_z.  Host arrays in, one byte per atom out; the caller's np.where keeps
 * the reference's index order.                                                                                   */
int mcpyb_region_mask_host(mcpyb_engine* e, const double* positions, const int32_t* numbers, int n, const double* center3,
                           double radius_squared, int clip_z, double bottom_z, const int32_t* species_Z, int n_species,
                           uint8_t* mask_out);
/* ---- min_insert test: replaces the distance loop of InsertionMove.do_trial_move
 *      (mcpy/moves/insertion_move.py:54-62) and MoleculeInsertionMove
 *      (molecule_insertion_move.py:61-72).  For each of C candidate points: the smallest
 *      squared image distance to the atoms of replica `rep` with mask[j] != 0 (all atoms if
 *      mask is NULL) that lie within mcpyb_cutoff(); +inf if none does.                */
int mcpyb_min_dist2_host(mcpyb_engine* e, int rep, const double* points, int64_t C,
                         const uint8_t* mask, double* out_d2);

/* Host-side evaluation of the same PCG64 restatement (no GPU needed): n doubles of
 * Generator.random() after skipping `skip` draws.  Used by the CPU tests to pin the
 * restatement to numpy.                                                          */
int mcpyb_pcg64_host_doubles(const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo,
                             uint64_t skip, int64_t n, double* out);

/* ---- instrumentation ---------------------------------------------------- */
/* Measured FP64 FMA throughput of this GPU (TFLOP/s, 2 flop per DFMA): the
 * roofline denominator for the pair kernels.                                 */
int mcpyb_fp64_peak_tflops(mcpyb_engine* e, double* tflops_out);
/* Per-stage device time of the batched LJ trial pipeline (CUDA events between its launches):
 * trial_sites | trial_scan | trial_scatter | row sums | combine.  enable != 0 turns the
 * recording on; every call folds the most recent trial launch into running means (ms).          */
int mcpyb_stage_timing(mcpyb_engine* e, int enable, double* mean_ms_out5, int64_t* calls_out);
/* Number of kernels this engine has launched since creation.                 */
int64_t mcpyb_launch_count(const mcpyb_engine* e);
/* The resident chain behind mcpyb_tracked_energy_host (LJ): instead of one launch per call -- the pattern of
 * GrandCanonicalEnsemble.do_gcmc_step, mcpy/ensembles/grand_canonical_ensemble.py:244-308 -- a small grid stays
 * resident while calls keep coming and takes each trial from a mailbox in page-locked host memory; it leaves after
 * 1.5 ms without a request and before anything it reads is rebuilt.  enable = 0 / 1 switches it, -1 leaves it as it
 * is.  launches_out / trials_out (may be NULL): grids launched and trials answered through the mailbox so far.    */
int mcpyb_resident_chain(mcpyb_engine* e, int enable, int64_t* launches_out, int64_t* trials_out);

/* ---- device-resident GCMC chains ---------------------------------------------
 * Whole Markov chains of GrandCanonicalEnsemble.do_gcmc_step (mcpy/ensembles/grand_canonical_ensemble.py:244-308)
 * on the device, ONE CTA per chain, the chain's state (positions, atom order, cell lists, random-number states) held
 * in the CTA's shared memory for the length of a launch: MoveSelector.do_trial_move (moves/move_selector.py:74-96),
 * InsertionMove / DeletionMove / DisplacementMove.do_trial_move (insertion_move.py:30-74, deletion_move.py:19-47,
 * displacement_move.py:51-84) over a periodic box Cell (cell/cell.py:34-41), the LJ trial dE, _acceptance_condition
 * (:197-242), atoms.wrap() and _record_minimum (base_ensemble.py:142-150) on acceptance.  The five MT19937 streams
 * (random.Random: selector, one per move, acceptance -- utils/random_number_generator.py:4-47) and the cells' PCG64
 * streams are taken from and handed back to the caller, so a chain can alternate between the device and the
 * reference's own host loop and the draw sequence, the accept / reject sequence and the configuration stay those of
 * the host loop.  Scope: one atomic species, Lennard-Jones, orthorhombic box periodic along all three axes with
 * rc < L/2, no constraints, no min_insert; anything else goes through the calculator path.                      */
#define MCPYB_GCMC_MAX_MOVES 8
#define MCPYB_GCMC_MT_WORDS 625      /* 624 state words + the position, as random.Random.getstate()[1] */
typedef struct mcpyb_gcmc mcpyb_gcmc;
enum { MCPYB_GCMC_INSERT = 0, MCPYB_GCMC_DELETE = 1, MCPYB_GCMC_DISPLACE = 2 };
enum {                                   /* per-chain status after mcpyb_gcmc_run                               */
    MCPYB_GCMC_DONE = 0,                 /* all requested trials done                                           */
    MCPYB_GCMC_CELL_FULL = 1,            /* stopped between two trials: a cell list reached its capacity         */
    MCPYB_GCMC_ATOMS_FULL = 2,           /* stopped between two trials: the atom storage is full                 */
    MCPYB_GCMC_TRACE_FULL = 3,           /* stopped between two trials: the trace buffer is full                 */
    MCPYB_GCMC_EMPTY_DISPLACE = 4        /* DisplacementMove on an empty configuration (the reference raises)    */
};
typedef struct mcpyb_gcmc_desc {
    double box[3];                                   /* diagonal of atoms.cell                                   */
    double beta, mu, lambda3;                        /* SetUnits.beta, mu[species], lambda_dbs[species] ** 3      */
    double lj_epsilon, lj_sigma, lj_rc;
    double move_cum_weight[MCPYB_GCMC_MAX_MOVES];    /* MoveSelector.rho                                         */
    double move_volume[MCPYB_GCMC_MAX_MOVES];        /* move.get_volume() (insertion / deletion)                 */
    double move_max_displacement[MCPYB_GCMC_MAX_MOVES];
    int32_t move_kind[MCPYB_GCMC_MAX_MOVES];         /* MCPYB_GCMC_INSERT / _DELETE / _DISPLACE                  */
    int32_t move_limit[MCPYB_GCMC_MAX_MOVES];        /* max_atoms (insertion) / min_atoms (deletion); -1 = none  */
    int32_t move_pcg[MCPYB_GCMC_MAX_MOVES];          /* insertion: which PCG64 stream (cell) it draws from        */
    int32_t n_move_types;
    int32_t atom_capacity;                           /* storage, in atoms (<= 65535)                             */
    int32_t cell_capacity;                           /* slots per cell of the chain's cell list; 0 = choose       */
} mcpyb_gcmc_desc;
typedef struct mcpyb_gcmc_scalars {                  /* read / written by get_state / set_state                  */
    double energy;                                   /* E_old                                                    */
    double best_score, best_energy;                  /* _best_score / _best_energy                               */
    int64_t attempts[MCPYB_GCMC_MAX_MOVES];          /* since set_state: move_counter                            */
    int64_t accepted[MCPYB_GCMC_MAX_MOVES];          /*                  move_acceptance                         */
    int64_t failed[MCPYB_GCMC_MAX_MOVES];            /*                  move_failed_counter                     */
    int64_t trials_done;                             /* since set_state                                          */
    uint64_t pcg[MCPYB_GCMC_MAX_MOVES][4];           /* per stream: state hi, lo, inc hi, lo                     */
    int32_t n_atoms;
    int32_t best_n;                                  /* atoms of the best configuration; -1 = not improved        */
    int32_t last_move;                               /* MoveSelector.to_use                                      */
    int32_t status;
} mcpyb_gcmc_scalars;
typedef struct mcpyb_gcmc_trial {                    /* one record per trial when a trace is requested           */
    double dE;                                       /* NaN for a failed proposal                                */
    int32_t move;                                    /* index into the move list                                 */
    int32_t flags;                                   /* bit 0 accepted, bit 1 failed proposal, bit 2 a draw was consumed */
    int32_t n_atoms;                                 /* after the trial                                          */
    int32_t index;                                   /* deleted / displaced atom index, insertion: N              */
} mcpyb_gcmc_trial;
/* n_chains chains, desc[c] each; trace_capacity > 0 keeps the last launch's trial records per chain.  */
int mcpyb_gcmc_create(mcpyb_engine* e, int n_chains, const mcpyb_gcmc_desc* desc, int64_t trace_capacity,
                      mcpyb_gcmc** out);
void mcpyb_gcmc_destroy(mcpyb_gcmc* g);
/* positions[n][3] in atom order; mt[(n_move_types + 2)][625]: selector, moves in list order, acceptance.  */
int mcpyb_gcmc_set_state(mcpyb_gcmc* g, int chain, const double* positions, const mcpyb_gcmc_scalars* s,
                         const uint32_t* mt);
/* positions_out[atom_capacity][3], best_positions_out[atom_capacity][3] (may be NULL). */
int mcpyb_gcmc_get_state(mcpyb_gcmc* g, int chain, double* positions_out, mcpyb_gcmc_scalars* s_out,
                         uint32_t* mt_out, double* best_positions_out);
/* Every chain advances by up to n_trials[c] trials in ONE launch (grid = chains) and the call returns when the
 * launch has finished; a chain that stops early says why in its status and how far it got in trials_done.   */
int mcpyb_gcmc_run(mcpyb_gcmc* g, const int64_t* n_trials, int32_t* status_out);
/* scalars of a chain as of the end of the last mcpyb_gcmc_run / set_state / swap (host copy, no device traffic) */
int mcpyb_gcmc_get_scalars(mcpyb_gcmc* g, int chain, mcpyb_gcmc_scalars* s_out);
/* the same for all chains at once, only what a replica driver reads between two blocks: E_old, the particle number, trials
 * done since set_state (arrays of n_chains entries) and {best score, best energy} (2 n_chains) */
int mcpyb_gcmc_get_block_scalars(mcpyb_gcmc* g, double* energy, int32_t* n_atoms, int64_t* trials_done, double* best2);
/* BatchedReplicaExchange._swap_states (batched_replica_exchange.py:401-410) between two chains of one set, on the
 * device: configuration, energy and particle number change slots; generators, counters, the running minimum,
 * temperature and chemical potential stay.  MCPYB_ERR_CAPACITY when the chains' boxes or storage layouts differ. */
int mcpyb_gcmc_swap_configurations(mcpyb_gcmc* g, int i, int j);
/* A swap across two processes without the host (ReplicaExchange's sendrecv of atoms, replica_exchange.py:188-192):
 * the device address and size of a chain's configuration block (positions by slot, atom order, cell lists, wrap
 * bookkeeping -- position-independent, so a peer with the same storage layout can take it byte for byte over NCCL) and
 * the five scalars that go with it {energy, atoms, unsettled atoms, max |E| since the last full sum, accepted moves since
 * then}; set_config_scalars installs the scalars that came with a received block.  Valid between two launches.  */
int mcpyb_gcmc_config_ptr(mcpyb_gcmc* g, int chain, void** dev_ptr, int64_t* bytes, double* scalars5);
int mcpyb_gcmc_set_config_scalars(mcpyb_gcmc* g, int chain, const double* scalars5);
int mcpyb_gcmc_get_trace(mcpyb_gcmc* g, int chain, mcpyb_gcmc_trial* out, int64_t capacity, int64_t* n_out);
/* cell list geometry chosen for a chain: cells per axis, slots per cell, bytes of state, 1 if it lives in shared memory */
int mcpyb_gcmc_info(mcpyb_gcmc* g, int chain, int32_t* out6);
/* out8[0..2]: SM cycles of a chain's last launch -- thread 0's serial section (acceptance, commit, the part of the
 * next proposal that was not drawn ahead), the whole trial loop, loop trips: what the serial Markov chain costs next to
 * the CTA-wide row sums.  out8[4]: cycles thread 0 spent drawing the next proposal ahead, behind the row sums.
 * out8[3]: how often the chain's running energy was replaced by a full sum over its configuration (done at the end of a
 * launch when accepted moves x 2^-53 x max |E| since the last full sum exceeds 1e-12 max(1, |E|)).
 * out8[5]: cycles of the row sums (first evaluator thread); out8[6]: thread 0 from the end of its look-ahead to its next
 * serial section (waiting for the row sums + the barrier); out8[7]: from the end of a serial section to the look-ahead
 * (the barrier that releases the evaluator).  */
int mcpyb_gcmc_cycles(mcpyb_gcmc* g, int chain, int64_t* out8);
/* device time of the last mcpyb_gcmc_run launch (CUDA events on the engine's stream) */
int mcpyb_gcmc_last_run_ms(mcpyb_gcmc* g, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* MCPY_B200_H */
