// K2 / K3 / K6 -- Lennard-Jones total energy and trial-move delta-E row sums.
//
// Arithmetic follows ase.calculators.lj.LennardJones (ASE 3.23.0, smooth=False;
// SURVEY.md 8a row a8; in-repo restatement benchmark/lammps_gcmc_parity.py:168-187):
//   c6 = (sigma^2 / r2)^3, pairs dropped when r2 > rc^2, c12 = c6^2,
//   e_pair = 4 eps (c12 - c6) - e0,  e0 = 4 eps ((sigma/rc)^12 - (sigma/rc)^6),
//   energies[i] = 0.5 * sum_j e_pair,  E = sum_i energies[i].
// The reference obtains a trial move's delta-E by recomputing that total
// (mcpy/ensembles/grand_canonical_ensemble.py:283-284); K3 evaluates only the
// rows of the moved atoms: sum_j u(new_a, j) - sum_j u(old_b, j) over atoms j
// that are not part of the move, plus the pairs inside the moved set.
#pragma once
#include "common.cuh"
#include "stencil.cuh"

struct LJParams {
    double eps4;      // 4 * epsilon
    double sigma2;    // sigma^2
    double rc2;       // rc^2
    double e0;        // shift at rc
    double rc;        // rc
};

// 1/x to within an ulp: hardware seed (rcp.approx.ftz.f64, rel. error e0 = 2^-23) + ONE third-order step
// y (1 + e + e^2), e = 1 - x y: the error after it is e^3 = 2^-69, below the rounding of the last fma.  Three DFMA
// instead of the four of two Newton steps (the row-sum loop is bound by the FP64 pipe: 14 instead of 15 per
// candidate); a full IEEE division costs ~3x more.  The pair energy only needs 1e-16 relative.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

__device__ __forceinline__ double lj_pair(const LJParams& p, double r2) {
    const double t = p.sigma2 * fast_rcp(r2);
    const double c6 = t * t * t;
    return fma(p.eps4, fma(c6, c6, -c6), -p.e0);      // 4 eps (c12 - c6) - e0
}

// ---------------------------------------------------------------------------
// K2: per-atom energies, one warp per atom.  e_atom[off+i] = 0.5 * sum_j e_pair.
// grid = ceil(total_atoms / warps_per_block); atom -> replica via binary search
// over atom_off (R is small).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int find_replica(const int* __restrict__ atom_off, int R, int a) {
    int lo = 0, hi = R;            // atom_off[lo] <= a < atom_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (atom_off[mid] <= a) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
lj_atom_energy_kernel(BatchView b, const int* __restrict__ atom_off, int total_atoms,
                      LJParams p, double* __restrict__ e_atom, int* __restrict__ pair_count) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int a = warp; a < total_atoms; a += nwarps) {
        const int r = find_replica(atom_off, b.R, a);
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int i = a - off;
        const double4 me = b.upos[a];
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        double acc = 0.0;
        int cnt = 0;
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                if (c.r2 <= p.rc2 && !(c.same_image && tag_index(c.tag) == i)) {
                    acc += lj_pair(p, c.r2);
                    ++cnt;
                }
            });
        acc = warp_sum(acc);
        cnt = warp_sum_int(cnt);
        if (lane == 0) {
            e_atom[a] = 0.5 * acc;
            if (pair_count) pair_count[a] = cnt;
        }
    }
}

// ---------------------------------------------------------------------------
// LJ forces (ase.calculators.lj.LennardJones, smooth=False):
//   F_i = sum_j  -24 eps (2 c12 - c6) / r2 * d_ij,   d_ij = pos_j + S @ cell - pos_i,  r2 <= rc^2,
// the input of the relax-then-energy contract (mcpy/calculators/base_calculator.py:14-22,
// CanonicalEnsemble.relax canonical_ensemble.py:115-123).  One warp per atom; forces[3a .. 3a+2].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lj_forces_kernel(BatchView b, const int* __restrict__ atom_off, int total_atoms,
                 LJParams p, double* __restrict__ forces) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int a = warp; a < total_atoms; a += nwarps) {
        const int r = find_replica(atom_off, b.R, a);
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int i = a - off;
        const double4 me = b.upos[a];
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                if (c.r2 <= p.rc2 && !(c.same_image && tag_index(c.tag) == i)) {
                    const double inv = fast_rcp(c.r2);
                    const double t = p.sigma2 * inv;
                    const double c6 = t * t * t;
                    const double w = -6.0 * p.eps4 * fma(2.0 * c6, c6, -c6) * inv;     // -24 eps (2 c12 - c6) / r2
                    fx = fma(w, c.dx, fx); fy = fma(w, c.dy, fy); fz = fma(w, c.dz, fz);
                }
            });
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
        if (lane == 0) { forces[3 * (size_t)a] = fx; forces[3 * (size_t)a + 1] = fy; forces[3 * (size_t)a + 2] = fz; }
    }
}

// Energy AND forces in one pass for the relaxation loop (mcpyb_relax_lbfgs_host): the pair loop that gives the
// force also gives the pair energy, so the separate energy kernel (31 of ~100 us of GPU time per optimiser step at
// 500 atoms) disappears.  A relaxation works on ONE small configuration, where a warp per atom leaves the GPU
// empty (500 warps on 148 SMs): here a CTA of EF_WARPS warps shares one atom, warp w walking the stencil rows
// w, w + EF_WARPS, ...; the warps' partial sums are folded in warp order, so the result does not depend on timing.
// e_atom[a] = 0.5 * sum of pair energies (ASE's results['energies']), pair_count[a] = in-cutoff pairs.
#define EF_WARPS 4
__global__ void __launch_bounds__(32 * EF_WARPS)
lj_energy_forces_kernel(BatchView b, const int* __restrict__ atom_off, int total_atoms, LJParams p,
                        double* __restrict__ forces, double* __restrict__ e_atom, int* __restrict__ pair_count) {
    __shared__ double s_part[EF_WARPS][4];
    __shared__ int s_cnt[EF_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = blockIdx.x; a < total_atoms; a += gridDim.x) {
        const int r = find_replica(atom_off, b.R, a);
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int i = a - off;
        const double4 me = b.upos[a];
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        double fx = 0.0, fy = 0.0, fz = 0.0, en = 0.0;
        int cnt = 0;
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                if (c.r2 <= p.rc2 && !(c.same_image && tag_index(c.tag) == i)) {
                    const double inv = fast_rcp(c.r2);
                    const double t = p.sigma2 * inv;
                    const double c6 = t * t * t;
                    const double w = -6.0 * p.eps4 * fma(2.0 * c6, c6, -c6) * inv;     // -24 eps (2 c12 - c6) / r2
                    fx = fma(w, c.dx, fx); fy = fma(w, c.dy, fy); fz = fma(w, c.dz, fz);
                    en += fma(p.eps4, fma(c6, c6, -c6), -p.e0);                          // 4 eps (c12 - c6) - e0
                    ++cnt;
                }
            }, wid, EF_WARPS);
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz); en = warp_sum(en); cnt = warp_sum_int(cnt);
        if (lane == 0) { s_part[wid][0] = fx; s_part[wid][1] = fy; s_part[wid][2] = fz; s_part[wid][3] = en; s_cnt[wid] = cnt; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double v = 0.0;
#pragma unroll
            for (int ww = 0; ww < EF_WARPS; ++ww) v += s_part[ww][threadIdx.x];
            if (threadIdx.x < 3) forces[3 * (size_t)a + threadIdx.x] = v;
            else e_atom[a] = 0.5 * v;
        }
        if (threadIdx.x == 4) {
            int c = 0;
#pragma unroll
            for (int ww = 0; ww < EF_WARPS; ++ww) c += s_cnt[ww];
            pair_count[a] = c;
        }
        __syncthreads();
    }
}

// Fixed-order per-replica reduction: E[r] = sum_i e_atom[off+i]; one CTA per replica.
// With e_part != nullptr it first folds the stencil-plane partials of the lane-per-atom LJ kernel:
// e_atom[a] = 0.5 * sum_planes e_part[a][plane], pair_count[a] = sum_planes cnt_part[a][plane].
__global__ void __launch_bounds__(1024)
reduce_replica_kernel(double* __restrict__ e_atom, const int* __restrict__ atom_off,
                      double* __restrict__ E, int* __restrict__ pair_count,
                      long long* __restrict__ pairs, const GridDesc* __restrict__ grid,
                      int* __restrict__ status, const double* __restrict__ e_part,
                      const int* __restrict__ cnt_part, int nplanes) {
    __shared__ double s[32];
    __shared__ long long sc[32];
    const int r = blockIdx.x;
    const int beg = atom_off[r], end = atom_off[r + 1];
    double acc = 0.0;
    long long cnt = 0;
    for (int a = beg + threadIdx.x; a < end; a += blockDim.x) {
        if (e_part) {
            double sp = 0.0;
            int c = 0;
            for (int pl = 0; pl < nplanes; ++pl) {
                sp += e_part[(size_t)a * nplanes + pl];
                c += cnt_part[(size_t)a * nplanes + pl];
            }
            e_atom[a] = 0.5 * sp;
            pair_count[a] = c;
        }
        acc += e_atom[a];
        if (pair_count) cnt += pair_count[a];
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5] = acc; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        long long c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s[w]; c += sc[w]; }
        E[r] = t;
        if (pairs) pairs[r] = c / 2;          // both-ways list -> unordered pairs
        if (status) status[r] = grid[r].error;
    }
}

// ---------------------------------------------------------------------------
// K3 / K6: batched trial moves.  Trial t acts on replica rep[t]; it removes the
// atoms old_idx[old_off[t] .. old_off[t+1]) and adds atoms at
// new_pos[new_off[t] .. new_off[t+1]).  displacement = both, same count;
// insertion = new only; deletion = old only.  G threads cooperate on one trial.
// ---------------------------------------------------------------------------
struct TrialBatch {
    int T;
    const int* rep;        // [T] or nullptr (all trials on replica 0)
    const int* old_off;    // [T+1]
    const int* old_idx;    // [old_off[T]]
    const int* new_off;    // [T+1]
    const double* new_pos; // [new_off[T]][3]
    const int* new_Z;      // [new_off[T]] atomic numbers (EMT) or nullptr
};

template <int G>
__device__ __forceinline__ double group_sum(double v, double* s_part, int gl) {
    v = warp_sum(v);
    if (G == 32) return v;
    // G > 32: one group per block
    __syncthreads();
    if ((gl & 31) == 0) s_part[gl >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) t += s_part[w];
    return t;
}
template <int G>
__device__ __forceinline__ int group_sum_int(int v, int* s_part, int gl) {
    v = warp_sum_int(v);
    if (G == 32) return v;
    __syncthreads();
    if ((gl & 31) == 0) s_part[gl >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) t += s_part[w];
    return t;
}

// Pairs inside a moved set of k atoms at positions P (k x 3, generic pointer):
// every unordered pair (a < a', all images) and every self image (a == a', S > 0 lexicographically).
template <class PosFn>
__device__ __forceinline__ void intra_set_lj(const GridDesc& g, const LJParams& p, int k, PosFn pos,
                                             int gl, int G, double& acc, int& cnt) {
    if (k == 0) return;
    int nim[3];
    bool mic_only = true;          // every periodic height >= 2 rc: at most the minimum image is in range
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        nim[d] = g.pbc[d] ? (int)ceil(p.rc * (1.0 + 1e-12) / g.height[d]) : 0;
        if (g.pbc[d] && !(g.height[d] >= 2.0 * p.rc * (1.0 + 1e-12))) mic_only = false;
    }
    if (mic_only) {
        // pairs a < a2, minimum image only; an atom cannot see its own images
        const int npair = k * (k - 1) / 2;
        for (int pr = gl; pr < npair; pr += G) {
            int a = 0, rem = pr;
            while (rem >= k - 1 - a) { rem -= (k - 1 - a); ++a; }
            const int a2 = a + 1 + rem;
            double xa, ya, za, xb, yb, zb;
            pos(a, xa, ya, za);
            pos(a2, xb, yb, zb);
            double f[3];
            to_frac(g, xb - xa, yb - ya, zb - za, f);
            const int s0 = g.pbc[0] ? -(int)rint(f[0]) : 0, s1 = g.pbc[1] ? -(int)rint(f[1]) : 0,
                      s2 = g.pbc[2] ? -(int)rint(f[2]) : 0;
            double sx, sy, sz, dx, dy, dz;
            shift_vec(g, s0, s1, s2, sx, sy, sz);
            const double r2 = image_r2(xb, yb, zb, sx, sy, sz, xa, ya, za, dx, dy, dz);
            if (r2 <= p.rc2) { acc += lj_pair(p, r2); ++cnt; }
        }
        return;
    }
    const int w0 = 2 * nim[0] + 1, w1 = 2 * nim[1] + 1, w2 = 2 * nim[2] + 1;
    const int nimg = w0 * w1 * w2;
    const int npair = k * (k + 1) / 2;
    for (int item = gl; item < npair * nimg; item += G) {
        const int pr = item / nimg, im = item - pr * nimg;
        // decode pair index -> (a <= a2)
        int a = 0, rem = pr;
        while (rem >= k - a) { rem -= (k - a); ++a; }
        const int a2 = a + rem;
        int s0 = im % w0 - nim[0], s1 = (im / w0) % w1 - nim[1], s2 = im / (w0 * w1) - nim[2];
        double xa, ya, za, xb, yb, zb;
        pos(a, xa, ya, za);
        pos(a2, xb, yb, zb);
        if (a == a2) {
            const bool lexpos = (s0 > 0) || (s0 == 0 && (s1 > 0 || (s1 == 0 && s2 > 0)));
            if (!lexpos) continue;
        } else {
            // centre the image window on the minimum image of (b - a)
            double f[3];
            to_frac(g, xb - xa, yb - ya, zb - za, f);
            if (g.pbc[0]) s0 -= (int)rint(f[0]);
            if (g.pbc[1]) s1 -= (int)rint(f[1]);
            if (g.pbc[2]) s2 -= (int)rint(f[2]);
        }
        double sx, sy, sz, dx, dy, dz;
        shift_vec(g, s0, s1, s2, sx, sy, sz);
        const double r2 = image_r2(xb, yb, zb, sx, sy, sz, xa, ya, za, dx, dy, dz);
        if (r2 <= p.rc2) { acc += lj_pair(p, r2); ++cnt; }
    }
}

template <int G>
__global__ void __launch_bounds__(256)
lj_trial_delta_kernel(BatchView b, TrialBatch tb, LJParams p,
                      double* __restrict__ dE, int* __restrict__ npairs) {
    __shared__ double s_part[8];
    __shared__ int s_parti[8];
    const int gl = threadIdx.x % G;                       // lane inside the group
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ngroups = (gridDim.x * blockDim.x) / G;
    for (int t = group; t < tb.T; t += ngroups) {
        const int r = tb.rep ? tb.rep[t] : 0;
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const double4* spos = b.spos + off;
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
        const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
        const int kold = oe - ob;
        double acc = 0.0;
        int cnt = 0;
        auto excluded = [&](int j) {
            bool ex = false;
            for (int e = ob; e < oe; ++e) ex |= (tb.old_idx[e] == j);
            return ex;
        };
        // rows of the new sites
        for (int s = nb; s < ne; ++s) {
            const double px = tb.new_pos[3 * s], py = tb.new_pos[3 * s + 1], pz = tb.new_pos[3 * s + 2];
            for_each_candidate(g, spos, cstart, px, py, pz, gl, G,
                [&](const Cand& c) {
                    if (c.r2 <= p.rc2 && !(kold && excluded(tag_index(c.tag)))) {
                        acc += lj_pair(p, c.r2);
                        ++cnt;
                    }
                });
        }
        // rows of the old sites (subtract)
        double accm = 0.0;
        int cntm = 0;
        for (int e = ob; e < oe; ++e) {
            const double4 me = b.upos[off + tb.old_idx[e]];
            for_each_candidate(g, spos, cstart, me.x, me.y, me.z, gl, G,
                [&](const Cand& c) {
                    if (c.r2 <= p.rc2 && !excluded(tag_index(c.tag))) {
                        accm += lj_pair(p, c.r2);
                        ++cntm;
                    }
                });
        }
        // pairs inside the moved sets
        intra_set_lj(g, p, ne - nb,
            [&](int a, double& x, double& y, double& z) {
                x = tb.new_pos[3 * (nb + a)]; y = tb.new_pos[3 * (nb + a) + 1]; z = tb.new_pos[3 * (nb + a) + 2];
            }, gl, G, acc, cnt);
        intra_set_lj(g, p, kold,
            [&](int a, double& x, double& y, double& z) {
                const double4 q = b.upos[off + tb.old_idx[ob + a]];
                x = q.x; y = q.y; z = q.z;
            }, gl, G, accm, cntm);
        const double plus = group_sum<G>(acc, s_part, gl);
        const double minus = group_sum<G>(accm, s_part, gl);
        const int c = group_sum_int<G>(cnt + cntm, s_parti, gl);
        if (gl == 0) {
            dE[t] = plus - minus;
            if (npairs) npairs[t] = c;
        }
    }
}

// ---------------------------------------------------------------------------
// Neighbour-pair listing (bit-exactness check of the cell list against
// oracle/neighbors.py).  mode 0: r2 < rc2; 1: r2 <= rc2; 2: sqrt(r2) < rc.
// One warp per atom; pairs appended through a global cursor (order is fixed up
// by the host-side sort in the tests).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
neighbor_pairs_kernel(BatchView b, int r, double rc, int mode, long long max_pairs,
                      int* __restrict__ out_i, int* __restrict__ out_j, int* __restrict__ out_S,
                      double* __restrict__ out_r2, unsigned long long* __restrict__ cursor) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const GridDesc& g = b.grid[r];
    const int off = g.atom_off;
    const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
    const double rc2 = rc * rc;
    for (int i = warp; i < g.n; i += nwarps) {
        const double4 me = b.upos[off + i];
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                const double r2 = c.r2;
                bool in = (mode == 0) ? (r2 < rc2) : (mode == 1) ? (r2 <= rc2) : (sqrt(r2) < rc);
                const int j = tag_index(c.tag);
                if (in && !(c.same_image && j == i)) {
                    const unsigned long long slot = atomicAdd(cursor, 1ULL);
                    if ((long long)slot < max_pairs) {
                        out_i[slot] = i;
                        out_j[slot] = j;
                        // S from its shift vector: S = round((S @ cell) @ inv)
                        double f[3];
                        to_frac(g, c.ax, c.ay, c.az, f);
                        out_S[3 * slot + 0] = (int)rint(f[0]);
                        out_S[3 * slot + 1] = (int)rint(f[1]);
                        out_S[3 * slot + 2] = (int)rint(f[2]);
                        out_r2[slot] = r2;
                    }
                }
            });
    }
}

// ---------------------------------------------------------------------------
// K9: min_insert test of InsertionMove / MoleculeInsertionMove
// (mcpy/moves/insertion_move.py:54-62, molecule_insertion_move.py:61-72): for each candidate
// point, the smallest squared image distance to the atoms selected by `mask` (all atoms when
// mask == nullptr) that lie within the engine cutoff; +inf when there is none.  One warp per
// candidate point; the cell list limits the search to the stencil.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
min_dist2_kernel(BatchView b, int r, const double* __restrict__ points, int C,
                 const unsigned char* __restrict__ mask, double rcut, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const GridDesc& g = b.grid[r];
    const int off = g.atom_off;
    const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
    const double rc2 = rcut * rcut;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int c = warp; c < C; c += nwarps) {
        double best = inf;
        for_each_candidate(g, b.spos + off, cstart, points[3 * c], points[3 * c + 1], points[3 * c + 2], lane, 32,
            [&](const Cand& cd) {
                if (cd.r2 <= rc2 && (!mask || mask[tag_index(cd.tag)])) best = fmin(best, cd.r2);
            });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) out[c] = best;
    }
}
