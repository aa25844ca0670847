// K4 / K5 -- effective-medium-theory pair pass, embedding pass and trial delta-E.
//
// Arithmetic follows ase.calculators.emt.EMT (ASE 3.23.0, asap_cutoff=False;
// SURVEY.md 8a row a9).  ASE walks a half list and updates both atoms of a pair;
// here every atom sums its own row (both-ways), which is the same total:
//   theta(r)  = 1 / (1 + exp(acut (r - rc)))                      for r < rc_list
//   sigma1[i] = sum_j exp(-eta2_j (r - beta s0_j)) * ksi_ij * theta / gamma1_i
//   e_pair[i] = -sum_j 0.5 V0_i exp(-kappa_j (r/beta - s0_j)) * ksi_ij / gamma2_i * theta
//   ksi_ij    = n0_j / n0_i
//   ds = -log(sigma1/12) / (beta eta2);  x = lambda ds
//   e_embed[i] = E0 ((1+x) exp(-x) - 1) + 6 V0 exp(-kappa ds)     (-E0 when sigma1 == 0)
//   E = sum_i (e_pair[i] + e_embed[i])
// (ASE books a pair's -(y1+y2) half to each atom; here y1 goes to the atom it
// belongs to.  The total is identical; per-atom values differ for unlike pairs.)
#pragma once
#include "common.cuh"
#include "stencil.cuh"
#include "lj_kernels.cuh"      // TrialBatch, group_sum, find_replica

struct EmtTable {
    double rc, rc_list, acut, beta, inv_beta;
    int nspecies;
    EmtSpecies sp[MCPYB_MAX_SPECIES];
    double ksi[MCPYB_MAX_SPECIES][MCPYB_MAX_SPECIES];   // ksi[a][b] = n0_b / n0_a
};

// Contribution of a neighbour of species b at distance r to an atom of species a.
__device__ __forceinline__ void emt_row_terms(const EmtTable& t, int a, int b, double r,
                                              double& dsigma, double& depair) {
    const EmtSpecies& pa = t.sp[a];
    const EmtSpecies& pb = t.sp[b];
    const double ksi = t.ksi[a][b];
    const double x = exp(t.acut * (r - t.rc));
    const double theta = fast_rcp(1.0 + x);
    // the three divisions of ASE's expression (r / beta, / gamma2, / gamma1) become multiplications
    // by reciprocals that are within an ulp: 1e-16 relative, far inside the 1e-10 bar
    depair = -(0.5 * pa.V0 * exp(-pb.kappa * (r * t.inv_beta - pb.s0)) * ksi * pa.inv_gamma2 * theta);
    dsigma = exp(-pb.eta2 * (r - t.beta * pb.s0)) * ksi * theta * pa.inv_gamma1;
}

__device__ __forceinline__ double emt_embed(const EmtTable& t, int a, double sigma1) {
    const EmtSpecies& p = t.sp[a];
    if (!(sigma1 > 0.0)) return -p.E0;               // ASE: log(0) raises -> energy -= E0
    const double ds = -log(sigma1 / 12.0) / (t.beta * p.eta2);
    const double x = p.lambda * ds;
    const double y = exp(-x);
    const double z = 6.0 * p.V0 * exp(-p.kappa * ds);
    return p.E0 * ((1.0 + x) * y - 1.0) + z;
}

// K4: one warp per atom; writes sigma1[a] and e_atom[a] = e_pair (embedding added by K5).
__global__ void __launch_bounds__(256)
emt_pair_kernel(BatchView b, const int* __restrict__ atom_off, int total_atoms,
                const EmtTable* __restrict__ tp, double* __restrict__ e_atom,
                double* __restrict__ sigma1, int* __restrict__ pair_count) {
    const EmtTable& t = *tp;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const double rcl = t.rc_list;
    for (int a = warp; a < total_atoms; a += nwarps) {
        const int r = find_replica(atom_off, b.R, a);
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int i = a - off;
        const double4 me = b.upos[a];
        const int si = tag_species(tag_of(me));
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        double sig = 0.0, ep = 0.0;
        int cnt = 0;
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                const double rr = sqrt(c.r2);
                if (rr < rcl && !(c.same_image && tag_index(c.tag) == i)) {
                    double ds, de;
                    emt_row_terms(t, si, tag_species(c.tag), rr, ds, de);
                    sig += ds;
                    ep += de;
                    ++cnt;
                }
            });
        sig = warp_sum(sig);
        ep = warp_sum(ep);
        cnt = warp_sum_int(cnt);
        if (lane == 0) {
            sigma1[a] = sig;
            e_atom[a] = ep;
            if (pair_count) pair_count[a] = cnt;
        }
    }
}

// K5: e_emb[a] = embedding(sigma1[a]), e_atom[a] += e_emb[a]; one thread per atom.
__global__ void __launch_bounds__(256)
emt_embed_kernel(BatchView b, int total_atoms, const EmtTable* __restrict__ tp,
                 const double* __restrict__ sigma1, double* __restrict__ e_atom, double* __restrict__ e_emb) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total_atoms) return;
    const int sp = tag_species(tag_of(b.upos[a]));
    const double v = emt_embed(*tp, sp, sigma1[a]);
    e_emb[a] = v;
    e_atom[a] += v;
}

// ---------------------------------------------------------------------------
// EMT forces (ase.calculators.emt.EMT.interact1 / interact2, ASE 3.23.0), per-atom rows over the
// both-ways neighbourhood.  With d = pos_j + S @ cell - pos_i, r = |d|, x = exp(acut (r - rc)),
// theta = 1 / (1 + x), a / b the species of i / j:
//   y1 = 0.5 V0_a exp(-kappa_b (r/beta - s0_b)) ksi_ab / gamma2_a theta,  y2 = the same with a <-> b
//   pair term      g  = (y1 kappa_b + y2 kappa_a) / beta + (y1 + y2) acut theta x
//   z1 = exp(-eta2_b (r - beta s0_b)) ksi_ab / gamma1_a theta deds[i],     z2 = the same with a <-> b, deds[j]
//   embedding term f2 = (z1 eta2_b + z2 eta2_a) + (z1 + z2) acut theta x
//   F_i = sum_j (g - f2) d / r,     deds[a] = dE_c / dsigma1 (0 for an isolated atom)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emt_deds_kernel(BatchView b, int total_atoms, const EmtTable* __restrict__ tp,
                const double* __restrict__ sigma1, double* __restrict__ deds) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total_atoms) return;
    const EmtTable& t = *tp;
    const EmtSpecies& p = t.sp[tag_species(tag_of(b.upos[a]))];
    const double s1 = sigma1[a];
    double v = 0.0;
    if (s1 > 0.0) {
        const double ds = -log(s1 / 12.0) / (t.beta * p.eta2);
        const double x = p.lambda * ds;
        const double y = exp(-x);
        const double z = 6.0 * p.V0 * exp(-p.kappa * ds);
        v = (x * y * p.E0 * p.lambda + p.kappa * z) / (s1 * t.beta * p.eta2);
    }
    deds[a] = v;
}

__global__ void __launch_bounds__(256)
emt_forces_kernel(BatchView b, const int* __restrict__ atom_off, int total_atoms,
                  const EmtTable* __restrict__ tp, const double* __restrict__ deds,
                  double* __restrict__ forces) {
    const EmtTable& t = *tp;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const double rcl = t.rc_list;
    for (int a = warp; a < total_atoms; a += nwarps) {
        const int r = find_replica(atom_off, b.R, a);
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int i = a - off;
        const double4 me = b.upos[a];
        const int sa = tag_species(tag_of(me));
        const EmtSpecies& pa = t.sp[sa];
        const double dea = deds[a];
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for_each_candidate(g, b.spos + off, cstart, me.x, me.y, me.z, lane, 32,
            [&](const Cand& c) {
                const double rr = sqrt(c.r2);
                const int j = tag_index(c.tag);
                if (rr < rcl && !(c.same_image && j == i)) {
                    const int sb = tag_species(c.tag);
                    const EmtSpecies& pb = t.sp[sb];
                    const double kab = t.ksi[sa][sb], kba = t.ksi[sb][sa];
                    const double x = exp(t.acut * (rr - t.rc));
                    const double theta = fast_rcp(1.0 + x);
                    const double cut = t.acut * theta * x;
                    const double y1 = 0.5 * pa.V0 * exp(-pb.kappa * (rr * t.inv_beta - pb.s0)) * kab * pa.inv_gamma2 * theta;
                    const double y2 = 0.5 * pb.V0 * exp(-pa.kappa * (rr * t.inv_beta - pa.s0)) * kba * pb.inv_gamma2 * theta;
                    const double gp = (y1 * pb.kappa + y2 * pa.kappa) * t.inv_beta + (y1 + y2) * cut;
                    const double z1 = exp(-pb.eta2 * (rr - t.beta * pb.s0)) * kab * pa.inv_gamma1 * theta * dea;
                    const double z2 = exp(-pa.eta2 * (rr - t.beta * pa.s0)) * kba * pb.inv_gamma1 * theta * deds[off + j];
                    const double f2 = (z1 * pb.eta2 + z2 * pa.eta2) + (z1 + z2) * cut;
                    const double w = (gp - f2) / rr;
                    fx = fma(w, c.dx, fx); fy = fma(w, c.dy, fy); fz = fma(w, c.dz, fz);
                }
            });
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
        if (lane == 0) { forces[3 * (size_t)a] = fx; forces[3 * (size_t)a + 1] = fy; forces[3 * (size_t)a + 2] = fz; }
    }
}

// ---------------------------------------------------------------------------
// EMT trial delta-E.  Needs the resident sigma1[] and e_atom[] of the batch.
//   dE = sum_{new a} (e_pair_a + embed(sigma1_a))  -  sum_{old b} e_atom[b]
//      + sum_{unmoved j near a site} [ d e_pair_j + embed(sigma1_j + d sigma1_j) - embed(sigma1_j) ]
// Each unmoved atom j is handled once, at the first site (old sites first, then
// new) that has it within rc_list, with the contributions of ALL sites summed
// before the non-linear embedding.  This requires that no two periodic images
// of one atom can both be within rc_list of the moved set; when the box is too
// thin for that the trial reports NaN (the caller then recomputes totals).
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void emt_trial_delta_body(const BatchView& b, const TrialBatch& tb,
                                                     const EmtTable* __restrict__ tp,
                                                     const int* __restrict__ z_to_slot,
                                                     const double* __restrict__ sigma1,
                                                     const double* __restrict__ e_atom, double* __restrict__ dE,
                                                     int* __restrict__ npairs, double* s_part, int* s_parti,
                                                     const int* __restrict__ only = nullptr, int n_only = 0) {
    // `only` != nullptr: just the trials only[0 .. n_only) (the ones the list-driven kernel handed back)
    const EmtTable& t = *tp;
    const double rcl = t.rc_list;
    const int gl = threadIdx.x % G;
    // a CTA-wide group walks a site's stencil row-wise: warp w takes the rows w, w + 8, ... with its 32 lanes
    const int TL = G > 32 ? 32 : G, tl = G > 32 ? (gl & 31) : gl;
    const int trows = G > 32 ? G / 32 : 1, trow0 = G > 32 ? (gl >> 5) : 0;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ngroups = (gridDim.x * blockDim.x) / G;
    const int ntr = only ? n_only : tb.T;
    for (int kk = group; kk < ntr; kk += ngroups) {
        const int tr = only ? only[kk] : kk;
        const int r = tb.rep ? tb.rep[tr] : 0;
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const double4* spos = b.spos + off;
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        const int ob = tb.old_off[tr], oe = tb.old_off[tr + 1];
        const int nb = tb.new_off[tr], ne = tb.new_off[tr + 1];
        const int kold = oe - ob, knew = ne - nb, nsites = kold + knew;
        auto excluded = [&](int j) {
            bool ex = false;
            for (int e = ob; e < oe; ++e) ex |= (tb.old_idx[e] == j);
            return ex;
        };
        // site s: position, species, sign (+1 new, -1 old); old sites come first
        auto site = [&](int s, double& x, double& y, double& z, int& sp, double& sign) {
            if (s < kold) {
                const double4 q = b.upos[off + tb.old_idx[ob + s]];
                x = q.x; y = q.y; z = q.z; sp = tag_species(tag_of(q)); sign = -1.0;
            } else {
                const int k = nb + (s - kold);
                x = tb.new_pos[3 * k]; y = tb.new_pos[3 * k + 1]; z = tb.new_pos[3 * k + 2];
                sp = z_to_slot[tb.new_Z[k]]; sign = 1.0;
            }
        };
        // ---- thin-box guard: min periodic height must exceed 2 rc_list + diameter of the site set
        bool ok = true;
        if (nsites > 0) {
            double spread = 0.0;
            for (int s = 0; s < nsites; ++s) {
                double x0, y0, z0, sg; int sp0;
                site(s, x0, y0, z0, sp0, sg);
                for (int s2 = s + 1; s2 < nsites; ++s2) {
                    double x, y, z; int sp;
                    site(s2, x, y, z, sp, sg);
                    double f[3];
                    to_frac(g, x - x0, y - y0, z - z0, f);
                    for (int d = 0; d < 3; ++d) if (g.pbc[d]) f[d] -= rint(f[d]);
                    const double dx = f[0] * g.cell[0] + f[1] * g.cell[3] + f[2] * g.cell[6];
                    const double dy = f[0] * g.cell[1] + f[1] * g.cell[4] + f[2] * g.cell[7];
                    const double dz = f[0] * g.cell[2] + f[1] * g.cell[5] + f[2] * g.cell[8];
                    spread = fmax(spread, sqrt(dx * dx + dy * dy + dz * dz));
                }
            }
            for (int d = 0; d < 3; ++d)
                if (g.pbc[d] && !(g.height[d] > 2.0 * rcl + spread + 1e-9)) ok = false;
        }
        double acc = 0.0;
        int cnt = 0;
        if (ok) {
            // ---- (1) the moved atoms themselves
            for (int s = kold; s < nsites; ++s) {
                double px, py, pz, sg; int spa;
                site(s, px, py, pz, spa, sg);
                double sig = 0.0, ep = 0.0;
                for_each_candidate(g, spos, cstart, px, py, pz, tl, TL,
                    [&](const Cand& c) {
                        const double rr = sqrt(c.r2);
                        if (rr < rcl && !(kold && excluded(tag_index(c.tag)))) {
                            double ds, de;
                            emt_row_terms(t, spa, tag_species(c.tag), rr, ds, de);
                            sig += ds; ep += de; ++cnt;
                        }
                    }, trow0, trows);
                // other new atoms (and periodic self images): directed terms into this atom
                int nim[3];
                for (int d = 0; d < 3; ++d) nim[d] = g.pbc[d] ? (int)ceil(rcl * (1.0 + 1e-12) / g.height[d]) : 0;
                const int w0 = 2 * nim[0] + 1, w1 = 2 * nim[1] + 1, w2 = 2 * nim[2] + 1;
                const int nimg = w0 * w1 * w2;
                for (int item = gl; item < knew * nimg; item += G) {
                    const int a2 = item / nimg, im = item - a2 * nimg;
                    int s0 = im % w0 - nim[0], s1 = (im / w0) % w1 - nim[1], s2 = im / (w0 * w1) - nim[2];
                    double qx, qy, qz, sg2; int spb;
                    site(kold + a2, qx, qy, qz, spb, sg2);
                    if (kold + a2 == s) {
                        if (s0 == 0 && s1 == 0 && s2 == 0) continue;
                    } else {
                        double f[3];
                        to_frac(g, qx - px, qy - py, qz - pz, f);
                        if (g.pbc[0]) s0 -= (int)rint(f[0]);
                        if (g.pbc[1]) s1 -= (int)rint(f[1]);
                        if (g.pbc[2]) s2 -= (int)rint(f[2]);
                    }
                    double sx, sy, sz, dx, dy, dz;
                    shift_vec(g, s0, s1, s2, sx, sy, sz);
                    const double rr = sqrt(image_r2(qx, qy, qz, sx, sy, sz, px, py, pz, dx, dy, dz));
                    if (rr < rcl) {
                        double ds, de;
                        emt_row_terms(t, spa, spb, rr, ds, de);
                        sig += ds; ep += de; ++cnt;
                    }
                }
                sig = group_sum<G>(sig, s_part, gl);
                ep = group_sum<G>(ep, s_part, gl);
                if (gl == 0) acc += ep + emt_embed(t, spa, sig);
            }
            if (gl == 0)
                for (int e = ob; e < oe; ++e) acc -= e_atom[off + tb.old_idx[e]];
            // ---- (2) unmoved neighbours of any site
            for (int s = 0; s < nsites; ++s) {
                double px, py, pz, sgn; int sps;
                site(s, px, py, pz, sps, sgn);
                for_each_candidate(g, spos, cstart, px, py, pz, tl, TL,
                    [&](const Cand& c) {
                        const double rr = sqrt(c.r2);
                        if (!(rr < rcl)) return;
                        const int j = tag_index(c.tag);
                        if (kold && excluded(j)) return;
                        // handled at an earlier site already?
                        for (int s2 = 0; s2 < s; ++s2) {
                            double qx, qy, qz, sg2; int sp2;
                            site(s2, qx, qy, qz, sp2, sg2);
                            double dx, dy, dz;
                            const double r2b = image_r2(c.q.x, c.q.y, c.q.z, c.ax, c.ay, c.az, qx, qy, qz, dx, dy, dz);
                            if (sqrt(r2b) < rcl) return;
                        }
                        const int spj = tag_species(c.tag);
                        double dsig = 0.0, dep = 0.0;
                        for (int s2 = s; s2 < nsites; ++s2) {
                            double qx, qy, qz, sg2; int sp2;
                            site(s2, qx, qy, qz, sp2, sg2);
                            double rb = rr;
                            if (s2 != s) {
                                double dx, dy, dz;
                                rb = sqrt(image_r2(c.q.x, c.q.y, c.q.z, c.ax, c.ay, c.az, qx, qy, qz, dx, dy, dz));
                            }
                            if (rb < rcl) {
                                double ds, de;
                                emt_row_terms(t, spj, sp2, rb, ds, de);
                                dsig += sg2 * ds; dep += sg2 * de; ++cnt;
                            }
                        }
                        const double s_old = sigma1[off + j];
                        acc += dep + (emt_embed(t, spj, s_old + dsig) - emt_embed(t, spj, s_old));
                    }, trow0, trows);
            }
        }
        const double tot = group_sum<G>(acc, s_part, gl);
        const int c = group_sum_int<G>(cnt, s_parti, gl);
        if (gl == 0) {
            dE[tr] = ok ? tot : __longlong_as_double(0x7ff8000000000000LL);
            if (npairs) npairs[tr] = c;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(256)
emt_trial_delta_kernel(BatchView b, TrialBatch tb, const EmtTable* __restrict__ tp,
                       const int* __restrict__ z_to_slot, const double* __restrict__ sigma1,
                       const double* __restrict__ e_atom, double* __restrict__ dE,
                       int* __restrict__ npairs) {
    __shared__ double s_part[8];
    __shared__ int s_parti[8];
    emt_trial_delta_body<G>(b, tb, tp, z_to_slot, sigma1, e_atom, dE, npairs, s_part, s_parti);
}

// One trial on the latency-bound single-chain path (mcpyb_tracked_energy_host): the whole CTA works on
// it (G = 256), its description rides in the kernel parameters and dE is written straight to pinned
// host memory -- one launch and one stream synchronisation per call, no H2D / D2H copies.
#define EMT_TINY_MAX_MOVED 8
struct EmtTinyTrial {
    int old_off[2], new_off[2];
    int old_idx[EMT_TINY_MAX_MOVED];
    int new_Z[EMT_TINY_MAX_MOVED];
    double new_pos[3 * EMT_TINY_MAX_MOVED];
};

__global__ void __launch_bounds__(256)
emt_trial_tiny_kernel(BatchView b, const __grid_constant__ EmtTinyTrial tt, const EmtTable* __restrict__ tp,
                      const int* __restrict__ z_to_slot, const double* __restrict__ sigma1,
                      const double* __restrict__ e_atom, double* __restrict__ dE_host,
                      volatile unsigned int* done_flag, unsigned int seq) {
    __shared__ double s_part[8];
    __shared__ int s_parti[8];
    TrialBatch tb;
    tb.T = 1; tb.rep = nullptr; tb.old_off = tt.old_off; tb.old_idx = tt.old_idx;
    tb.new_off = tt.new_off; tb.new_pos = tt.new_pos; tb.new_Z = tt.new_Z;
    emt_trial_delta_body<256>(b, tb, tp, z_to_slot, sigma1, e_atom, dE_host, nullptr, s_part, s_parti);
    // the host polls this word instead of waiting for the kernel to retire (engine.cu, wait_flag)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && done_flag) { *done_flag = seq; __threadfence_system(); }
}

// Note (round 1): a lane-per-site variant of these kernels (the scheme of site_kernels.cuh) was
// built and measured: with cell edge = rc_list only ~17 of 32 lanes hold an atom and the per-hit
// functor (three exp, two embeddings) is register-hungry; it ran 185 us against 100 us for
// emt_pair_kernel on the 10 825-atom S5 configuration and 4x slower on trial batches, so EMT stays
// on the warp-per-site traversal (profiles/README.md).
