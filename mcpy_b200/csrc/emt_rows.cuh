// EMT trial-move dE, second generation: one warp per trial over the PRUNED per-cell stencil lists.
//
// The first-generation kernel (emt_kernels.cuh: emt_trial_delta_body) walks the 27-cell stencil of every site with
// its lanes over the candidates -- 460 candidates per site for 54 hits on the CuPd particle, ~10 k warp instructions
// per trial, most of them the walk itself, and the exponentials of a hit run with a third of the lanes
// (profiles/README.md: issue active 27 %, 21 % of the stalls on instruction fetch).  Here
//
//   * a site's candidates are the records of its cell's stencil list (rows_block.cuh: image shift applied, pruned to
//     what lies within rc_list of the cell's box; .w = atom index | species << 24): one coalesced 1 KB read per 32
//     candidates, three subtractions and the unfused r^2 of the oracle per candidate, nothing else;
//   * the hits (sqrt(r^2) < rc_list, the oracle's predicate) are compacted through a ballot into a per-warp queue in
//     shared memory, and the expensive part -- three exp per directed pair term, a log and two exp per re-embedding --
//     runs over the queue with every lane busy;
//   * one walk per site serves both halves of the first generation's sum: the row of a new atom (its own density and
//     pair energy) and the change of every unmoved neighbour (handled once, at the first site that sees it, with its
//     summed density change re-embedded); the cutoff exponential of a pair is shared by its two directions, and the
//     neighbour's old embedding energy is read (e_emb, written by K5) instead of recomputed;
//   * exp / log are calls (er_exp, er_log), not inline expansions: inlined, the kernel was 5 k instructions and 4 of
//     its 10 stall cycles per issue were instruction fetch (profiles/README.md, r2 EMT rows v1).
//
// A trial whose site falls in a cell without a list (thin box, oversized stencil) is handed back through `redo` and
// evaluated by the first-generation body in a second, small launch.
#pragma once
#include "emt_kernels.cuh"
#include "rows_block.cuh"

#define ER_WARPS 4
#define ER_QCAP 128
#define ER_MAXS 8            // sites of a trial staged in shared memory; larger moves go to the first-generation body

struct ErQueue {
    double rr[ER_QCAP];          // r^2 of the hit
    int gi[ER_QCAP];
};

struct ErSites {
    double x[ER_MAXS], y[ER_MAXS], z[ER_MAXS];
    int sp[ER_MAXS];
    int old[ER_MAXS];
};

__device__ __noinline__ double er_exp(double v) { return exp(v); }
__device__ __noinline__ double er_log(double v) { return log(v); }
__device__ __noinline__ double er_sqrt(double v) { return sqrt(v); }

// emt_row_terms with the cutoff factor theta = 1 / (1 + exp(acut (r - rc))) handed in; returns (dsigma, depair)
__device__ __noinline__ double2 er_row_terms(const EmtTable& t, int a, int b, double r, double theta) {
    const EmtSpecies& pa = t.sp[a];
    const EmtSpecies& pb = t.sp[b];
    const double ksi = t.ksi[a][b];
    double2 o;
    o.y = -(0.5 * pa.V0 * er_exp(-pb.kappa * (r * t.inv_beta - pb.s0)) * ksi * pa.inv_gamma2 * theta);
    o.x = er_exp(-pb.eta2 * (r - t.beta * pb.s0)) * ksi * theta * pa.inv_gamma1;
    return o;
}

__device__ __forceinline__ double er_theta(const EmtTable& t, double r) {
    return fast_rcp(1.0 + er_exp(t.acut * (r - t.rc)));
}

__device__ __noinline__ double er_embed(const EmtTable& t, int a, double sigma1) {
    const EmtSpecies& p = t.sp[a];
    if (!(sigma1 > 0.0)) return -p.E0;
    const double ds = -er_log(sigma1 / 12.0) / (t.beta * p.eta2);
    const double x = p.lambda * ds;
    const double y = er_exp(-x);
    const double z = 6.0 * p.V0 * er_exp(-p.kappa * ds);
    return p.E0 * ((1.0 + x) * y - 1.0) + z;
}

// The smallest double v with sqrt(v) >= rc: "sqrt(r2) < rc" (the oracle's predicate) is "r2 < v", bit for bit, because
// the correctly rounded square root is monotone.
__device__ __forceinline__ double er_r2_limit(double rc) {
    double v = rc * rc;
    while (sqrt(v) >= rc) v = __longlong_as_double(__double_as_longlong(v) - 1);
    while (sqrt(v) < rc) v = __longlong_as_double(__double_as_longlong(v) + 1);
    return v;
}

__device__ __forceinline__ double er_warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int MINB>
__global__ void __launch_bounds__(32 * ER_WARPS, MINB)
emt_trial_rows_kernel(BatchView b, TrialBatch tb, const EmtTable* __restrict__ tp, const int* __restrict__ z_to_slot,
                      const double* __restrict__ sigma1, const double* __restrict__ e_atom,
                      const double* __restrict__ e_emb, StencilLists x, int bin_stride, double* __restrict__ dE,
                      int* __restrict__ npairs, int* __restrict__ redo) {
    __shared__ ErQueue s_qu[ER_WARPS];
    __shared__ ErSites s_st[ER_WARPS];
    __shared__ EmtTable s_t;
    for (int i = threadIdx.x; i < (int)(sizeof(EmtTable) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&s_t)[i] = reinterpret_cast<const int*>(tp)[i];
    __syncthreads();
    const EmtTable& t = s_t;
    const double rcl = t.rc_list;
    const double r2lim = er_r2_limit(rcl);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ErQueue& qu = s_qu[wid];
    ErSites& st = s_st[wid];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int tr = warp; tr < tb.T; tr += nwarps) {
        const int r = tb.rep ? tb.rep[tr] : 0;
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const double4* spos = b.spos + off;
        const int bin_base = r * bin_stride;
        const int ob = tb.old_off[tr], nb = tb.new_off[tr];
        const int kold = tb.old_off[tr + 1] - ob, knew = tb.new_off[tr + 1] - nb, nsites = kold + knew;
        bool listed = nsites <= ER_MAXS;
        // ---- the sites of the trial (old ones first) into shared memory
        __syncwarp();
        if (listed && lane < nsites) {
            if (lane < kold) {
                const int j = tb.old_idx[ob + lane];
                const double4 q = b.upos[off + j];
                st.x[lane] = q.x; st.y[lane] = q.y; st.z[lane] = q.z;
                st.sp[lane] = tag_species(tag_of(q));
                st.old[lane] = j;
            } else {
                const int k = nb + (lane - kold);
                st.x[lane] = tb.new_pos[3 * k]; st.y[lane] = tb.new_pos[3 * k + 1]; st.z[lane] = tb.new_pos[3 * k + 2];
                st.sp[lane] = z_to_slot[tb.new_Z[k]];
                st.old[lane] = -1;
            }
        }
        __syncwarp();
        // ---- thin-box guard: min periodic height must exceed 2 rc_list + diameter of the site set
        bool ok = true;
        if (listed && nsites > 0 && (g.pbc[0] | g.pbc[1] | g.pbc[2])) {
            double spread = 0.0;
            #pragma unroll 1
            for (int p = lane; p < nsites * nsites; p += 32) {
                const int s = p / nsites, s2 = p - s * nsites;
                if (s2 > s) {
                    double f[3];
                    to_frac(g, st.x[s2] - st.x[s], st.y[s2] - st.y[s], st.z[s2] - st.z[s], f);
                    for (int d = 0; d < 3; ++d) if (g.pbc[d]) f[d] -= rint(f[d]);
                    const double dx = f[0] * g.cell[0] + f[1] * g.cell[3] + f[2] * g.cell[6];
                    const double dy = f[0] * g.cell[1] + f[1] * g.cell[4] + f[2] * g.cell[7];
                    const double dz = f[0] * g.cell[2] + f[1] * g.cell[5] + f[2] * g.cell[8];
                    spread = fmax(spread, er_sqrt(dx * dx + dy * dy + dz * dz));
                }
            }
            spread = er_warp_max(spread);
            for (int d = 0; d < 3; ++d)
                if (g.pbc[d] && !(g.height[d] > 2.0 * rcl + spread + 1e-9)) ok = false;
        }
        double acc = 0.0;
        int cnt = 0;
        if (ok && listed) {
            if (lane < kold) acc -= e_atom[off + st.old[lane]];
            #pragma unroll 1
            for (int s = 0; s < nsites && listed; ++s) {
                const double px = st.x[s], py = st.y[s], pz = st.z[s];
                const int sps = st.sp[s];
                const bool isnew = s >= kold;
                double sig = 0.0, ep = 0.0;            // the new atom's own row
                // ---- this site's list
                int c[3], wr[3];
                locate(g, px, py, pz, c, wr);
                const int wsite = pack_wraps(wr[0], wr[1], wr[2]);
                const int2 head = x.head[bin_base + (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0]];
                if (head.x < 0) { listed = false; break; }
                MCPYB_CHECK(head.y >= 0 && (long long)head.x + head.y <= x.cap);
                const double4* xq = x.q + head.x;
                const int2* xi = x.info + head.x;
                // the list is written in the frame of the primary cell: a site stored outside it is shifted in, and
                // the shift is added back to the image positions compared with other, unshifted sites
                double wx = 0.0, wy = 0.0, wz = 0.0;
                const bool wrapped = wsite != pack_wraps(0, 0, 0);
                if (wrapped) shift_vec(g, wr[0], wr[1], wr[2], wx, wy, wz);
                const double qx = px - wx, qy = py - wy, qz = pz - wz;
                int nq = 0;
                auto drain = [&]() {
                    __syncwarp();
                    #pragma unroll 1
                    for (int i = lane; i < nq; i += 32) {
                        const double4 q = xq[qu.gi[i]];
                        const double rr = er_sqrt(qu.rr[i]);
                        const int lo = __double2loint(q.w);
                        const int j = lo & 0xFFFFFF, spj = (lo >> 24) & 0xFF;
                        bool ex = false;
                        #pragma unroll 1
                        for (int e = 0; e < kold; ++e) ex |= st.old[e] == j;
                        if (ex) continue;
                        const double cx = q.x + wx, cy = q.y + wy, cz = q.z + wz;
                        const double theta = er_theta(t, rr);
                        if (isnew) {
                            const double2 o = er_row_terms(t, sps, spj, rr, theta);
                            sig += o.x; ep += o.y; ++cnt;
                        }
                        // the neighbour's own change: once, at the first site that sees it
                        bool seen = false;
                        double dx, dy, dz;
                        #pragma unroll 1
                        for (int s2 = 0; s2 < s; ++s2)
                            seen |= er_sqrt(image_r2(cx, cy, cz, 0.0, 0.0, 0.0, st.x[s2], st.y[s2], st.z[s2], dx, dy, dz)) < rcl;
                        if (seen) continue;
                        double dsig = 0.0, dep = 0.0;
                        #pragma unroll 1
                        for (int s2 = s; s2 < nsites; ++s2) {
                            double rb = rr, th = theta;
                            if (s2 != s) {
                                rb = er_sqrt(image_r2(cx, cy, cz, 0.0, 0.0, 0.0, st.x[s2], st.y[s2], st.z[s2], dx, dy, dz));
                                if (!(rb < rcl)) continue;
                                th = er_theta(t, rb);
                            }
                            const double sg2 = s2 < kold ? -1.0 : 1.0;
                            const double2 o = er_row_terms(t, spj, st.sp[s2], rb, th);
                            dsig += sg2 * o.x; dep += sg2 * o.y; ++cnt;
                        }
                        acc += dep + (er_embed(t, spj, sigma1[off + j] + dsig) - e_emb[off + j]);
                    }
                    __syncwarp();
                    nq = 0;
                };
                auto test = [&](int gi, const double4& q, double& r2) -> bool {
                    const double dx = __dsub_rn(q.x, qx), dy = __dsub_rn(q.y, qy), dz = __dsub_rn(q.z, qz);
                    r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    // a shifted site: within rounding of the cutoff the oracle's own expression decides
                    if (wrapped && fabs(r2 - r2lim) <= r2lim * 1.0e-9) r2 = rp_exact_r2(g, spos, xi[gi], px, py, pz, wsite);
                    return r2 < r2lim;
                };
                auto push = [&](bool hit, int gi, double rr) {
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int at = nq + __popc(bal & ((1u << lane) - 1u));
                        qu.gi[at] = gi;
                        qu.rr[at] = rr;
                    }
                    nq += __popc(bal);
                };
                #pragma unroll 1
                for (int base = 0; ; base += 64) {
                    const bool more = base < head.y;
                    if (more) {
                        // two chunks in flight
                        const int g0 = base + lane, g1 = g0 + 32;
                        double4 q0 = make_double4(1e150, 1e150, 1e150, 0.0), q1 = q0;
                        if (g0 < head.y) q0 = xq[g0];
                        if (g1 < head.y) q1 = xq[g1];
                        double r0, r1;
                        const bool h0 = test(g0, q0, r0);
                        const bool h1 = test(g1, q1, r1);
                        push(h0, g0, r0);
                        push(h1, g1, r1);
                    }
                    if (!more || nq > ER_QCAP - 64) drain();
                    if (!more) break;
                }
                if (isnew) {
                    // other new atoms (and periodic self images): directed terms into this atom
                    int nim[3];
                    for (int d = 0; d < 3; ++d) nim[d] = g.pbc[d] ? (int)ceil(rcl * (1.0 + 1e-12) / g.height[d]) : 0;
                    const int w0 = 2 * nim[0] + 1, w1 = 2 * nim[1] + 1, w2 = 2 * nim[2] + 1;
                    const int nimg = w0 * w1 * w2;
                    #pragma unroll 1
                    for (int item = lane; item < knew * nimg; item += 32) {
                        const int a2 = item / nimg, im = item - a2 * nimg;
                        int s0 = im % w0 - nim[0], s1 = (im / w0) % w1 - nim[1], s2 = im / (w0 * w1) - nim[2];
                        const double ux = st.x[kold + a2], uy = st.y[kold + a2], uz = st.z[kold + a2];
                        if (kold + a2 == s) {
                            if (s0 == 0 && s1 == 0 && s2 == 0) continue;
                        } else {
                            double f[3];
                            to_frac(g, ux - px, uy - py, uz - pz, f);
                            if (g.pbc[0]) s0 -= (int)rint(f[0]);
                            if (g.pbc[1]) s1 -= (int)rint(f[1]);
                            if (g.pbc[2]) s2 -= (int)rint(f[2]);
                        }
                        double sx, sy, sz, dx, dy, dz;
                        shift_vec(g, s0, s1, s2, sx, sy, sz);
                        const double rr = er_sqrt(image_r2(ux, uy, uz, sx, sy, sz, px, py, pz, dx, dy, dz));
                        if (rr < rcl) {
                            const double2 o = er_row_terms(t, sps, st.sp[kold + a2], rr, er_theta(t, rr));
                            sig += o.x; ep += o.y; ++cnt;
                        }
                    }
                    sig = warp_sum(sig);
                    ep = warp_sum(ep);
                    if (lane == 0) acc += ep + er_embed(t, sps, sig);
                }
            }
        }
        const double tot = warp_sum(acc);
        const int cn = warp_sum_int(cnt);
        if (lane == 0) {
            if (!listed) {
                // a site without a list, or too many sites: the first-generation body takes this trial (second launch)
                const int k = atomicAdd(&redo[0], 1);
                redo[1 + k] = tr;
            } else {
                dE[tr] = ok ? tot : __longlong_as_double(0x7ff8000000000000LL);
                if (npairs) npairs[tr] = cn;
            }
        }
    }
}

// The trials the list-driven kernel handed back (redo[0] of them, indices in redo[1 ..]; the host zeroes the counter
// before every batch).
__global__ void __launch_bounds__(256)
emt_trial_redo_kernel(BatchView b, TrialBatch tb, const EmtTable* __restrict__ tp, const int* __restrict__ z_to_slot,
                      const double* __restrict__ sigma1, const double* __restrict__ e_atom, double* __restrict__ dE,
                      int* __restrict__ npairs, int* __restrict__ redo) {
    __shared__ double s_part[8];
    __shared__ int s_parti[8];
    const int n = redo[0];
    if (n > 0)
        emt_trial_delta_body<32>(b, tb, tp, z_to_slot, sigma1, e_atom, dE, npairs, s_part, s_parti, redo + 1, n);
}
