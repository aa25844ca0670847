// Lane-per-site row sums (K2 / K3 / K6, second generation).
//
// The first-generation kernels gave every site (an atom, or the old / new position
// of a moved atom) a whole warp whose lanes strode over the candidate atoms.  ncu
// showed them integer- and control-bound: ~215 warp instructions per 32 candidates,
// FP64 pipe 17 % busy (profiles/r1_lj_trial_v1.md).  Here the roles are swapped:
//
//   * sites are grouped by the cell they fall in; a warp takes up to 32 sites of ONE
//     cell, one site per lane;
//   * the candidate atoms of that cell's stencil are the same for all 32 lanes, so the
//     warp walks them with UNIFORM addresses (one broadcast 32-byte load per candidate)
//     while each lane evaluates its own distance -- full lane utilisation, no
//     shuffle reduction, the stencil bookkeeping amortised over 32 sites;
//   * each lane accumulates its row in candidate order, so a site's result does not
//     depend on which other sites share its warp (batch-independent bits).
//
// The stencil of a cell is (2 m1 + 1)(2 m2 + 1) rows; a row is at most two
// contiguous pieces of the cell-sorted array (the second one is the periodic wrap
// along x), worked out by one lane per row and broadcast with shuffles.
#pragma once
#include "common.cuh"
#include "stencil.cuh"
#include "lj_kernels.cuh"

// Visit every piece [beg, end) of cell (c0, c1, c2)'s stencil with its image shift.
// body(beg, end, S0, S1, S2) is called warp-uniformly.
template <class Body>
__device__ __forceinline__ void for_each_piece(const GridDesc& g, const int* __restrict__ cstart,
                                               int c0, int c1, int c2, int lane, Body&& body) {
    const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
    const int w1 = 2 * m1 + 1, w2 = 2 * m2 + 1;
    const int nrows = w1 * w2;
    const bool px = g.pbc[0] != 0;
    const bool simple_x = !px || n0 >= 2 * m0 + 1;
    for (int base = 0; base < nrows; base += 32) {
        const int row = base + lane;
        int begA = 0, endA = 0, begB = 0, endB = 0, SxB = 0, Sy = 0, Sz = 0, rowoff = 0;
        bool valid = row < nrows;
        if (valid) {
            int cy = c1 + (row % w1) - m1, cz = c2 + (row / w1) - m2;
            if (g.pbc[1]) { Sy = floor_div(cy, n1); cy -= Sy * n1; }
            else if (cy < 0 || cy >= n1) valid = false;
            if (g.pbc[2]) { Sz = floor_div(cz, n2); cz -= Sz * n2; }
            else if (cz < 0 || cz >= n2) valid = false;
            if (valid) {
                rowoff = (cz * n1 + cy) * n0;
                if (simple_x) {
                    const int lo = c0 - m0, hi = c0 + m0;
                    const int a0 = max(lo, 0), a1 = min(hi, n0 - 1);
                    begA = cstart[rowoff + a0];
                    endA = cstart[rowoff + a1 + 1];
                    if (px) {
                        if (lo < 0) { begB = cstart[rowoff + lo + n0]; endB = cstart[rowoff + n0]; SxB = -1; }
                        else if (hi >= n0) { begB = cstart[rowoff]; endB = cstart[rowoff + hi - n0 + 1]; SxB = 1; }
                    }
                }
            }
        }
        const int nr = min(32, nrows - base);
        for (int rr = 0; rr < nr; ++rr) {
            if (!__shfl_sync(0xffffffffu, (int)valid, rr)) continue;
            const int sy = __shfl_sync(0xffffffffu, Sy, rr), sz = __shfl_sync(0xffffffffu, Sz, rr);
            if (simple_x) {
                const int bA = __shfl_sync(0xffffffffu, begA, rr), eA = __shfl_sync(0xffffffffu, endA, rr);
                const int bB = __shfl_sync(0xffffffffu, begB, rr), eB = __shfl_sync(0xffffffffu, endB, rr);
                const int sB = __shfl_sync(0xffffffffu, SxB, rr);
                if (bA < eA) body(bA, eA, 0, sy, sz);
                if (bB < eB) body(bB, eB, sB, sy, sz);
            } else {
                // box thinner than the stencil along x: every offset is its own image
                const int ro = __shfl_sync(0xffffffffu, rowoff, rr);
                for (int ox = -m0; ox <= m0; ++ox) {
                    int cx = c0 + ox;
                    const int sx = floor_div(cx, n0);
                    cx -= sx * n0;
                    const int b = cstart[ro + cx], e = cstart[ro + cx + 1];
                    if (b < e) body(b, e, sx, sy, sz);
                }
            }
        }
    }
}

// Row sum of one site per lane over the stencil of cell (c0,c1,c2), three phases per item:
//
//   stage  : one lane per stencil row works out the row's (at most two) contiguous pieces of the
//            cell-sorted array; the warp copies their fp32 screening records -- with the piece's
//            image shift already added -- into a dense shared-memory array.
//   screen : every staged candidate is read with ONE broadcast LDS and tested in fp32 against
//            (rcut + margin)^2.  The margin (cell_list.cuh) bounds the fp32 rounding, so the screen
//            never drops a pair the exact test would keep.  Survivors are appended to the lane's
//            FIFO in shared memory (slot of the atom + slot of the piece).
//   exact  : when a FIFO is nearly full (and at the end) all lanes drain theirs: fp64 image vector
//            in the oracle's operation order, exact cutoff test, exclusions, pair energy,
//            accumulation in candidate order.
//
// The screen keeps the warp converged on cheap single-issue instructions; the expensive fp64 path
// runs on ~22 % of the candidates with most lanes busy instead of diverging on every candidate.
//   active : this lane holds a site
//   excl(j): true if atom j must be skipped in EVERY image (moved atoms)
//   self   : atom index to skip in the zero-shift image only (-1: none)
#define SITE_QCAP 32              // FIFO entries per lane (one byte each)
#define SITE_UNROLL 4
#define SITE_WARPS 4              // warps per block of the row-sum kernels
#define SITE_MIN_BLOCKS 6         // resident blocks per SM the compiler must allow
#define SITE_CAND 128             // candidates staged per chunk
#define SITE_ROWS 8               // stencil rows worked out per batch (one lane each)
#define SITE_PIECES (2 * SITE_ROWS) // two pieces per stencil row

struct SiteScratch {              // per warp, in shared memory (~8 KB)
    double4 candd[SITE_CAND];     // staged atoms: original (x, y, z, tag) -- the exact path reads these
    float4 candf[SITE_CAND];      // fp32 screening copy: wrapped - origin + piece shift; .w = piece slot
    unsigned char fifo[SITE_QCAP * 32];   // lane-private columns of staged-candidate indices
    int4 pieces[SITE_PIECES];     // beg, end, unused, kind (0 empty, 1 valid)
    double dshift[SITE_PIECES][3];// S @ cell of the piece
    int4 rawS[SITE_PIECES];       // S of the piece
};

// within(r2) -> bool : the exact cutoff test (LJ: r2 <= rc^2, EMT: sqrt(r2) < rc_list)
// hit(q, ax, ay, az, r2): one neighbour image inside the cutoff; q = its atom record, (ax,ay,az) =
//                         S @ cell of the image (absolute image position = q + a)
// plane_done()  : all rows of one stencil z-plane have been evaluated (FIFO drained) -- callers
//                 fold their per-plane partial sums here, which makes a row sum independent of how
//                 the planes are distributed over work items
template <class Within, class Hit, class PlaneDone>
__device__ __forceinline__ void site_rowsum(const GridDesc& g, const double4* __restrict__ spos,
                                            const float4* __restrict__ fpos,
                                            const int* __restrict__ cstart,
                                            int c0, int c1, int c2, int lane, bool active,
                                            double px, double py, double pz, int wsite, int self,
                                            int plane, int nplanes,
                                            SiteScratch& sc, Within&& within, Hit&& hit, PlaneDone&& plane_done) {
    int w0, w1, w2;
    unpack_wraps(wsite, w0, w1, w2);
    const float3 pf = screen_pos(g, px, py, pz, w0, w1, w2);
    const float sr2 = g.screen_r2;
    int qn = 0;

    // exact evaluation of one atom record against this lane's site; `slot` = its stencil piece
    auto exact = [&](const double4& q, int slot) {
        const long long tag = tag_of(q);
        const int j = tag_index(tag);
        const int wj = tag_wraps(tag);
        double r2, dx, dy, dz;
        if (wj == wsite) {
            const int4 S = sc.rawS[slot];
            if ((S.x | S.y | S.z) == 0) {       // q + 0.0 == q: skip the three additions, same bits
                dx = __dsub_rn(q.x, px); dy = __dsub_rn(q.y, py); dz = __dsub_rn(q.z, pz);
                r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (within(r2) && j != self) hit(q, 0.0, 0.0, 0.0, r2);
            } else {
                const double ax = sc.dshift[slot][0], ay = sc.dshift[slot][1], az = sc.dshift[slot][2];
                r2 = image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
                if (within(r2)) hit(q, ax, ay, az, r2);
            }
        } else {                                 // rare: atom or site outside the primary cell
            const int4 S = sc.rawS[slot];
            int jx, jy, jz;
            unpack_wraps(wj, jx, jy, jz);
            const int T0 = S.x + w0 - jx, T1 = S.y + w1 - jy, T2 = S.z + w2 - jz;
            double ax, ay, az;
            shift_vec(g, T0, T1, T2, ax, ay, az);
            r2 = image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
            const bool same = (T0 | T1 | T2) == 0;
            if (within(r2) && !(same && j == self)) hit(q, ax, ay, az, r2);
        }
    };
    auto drain = [&]() {
        const int nmax = __reduce_max_sync(0xffffffffu, qn);
        // two entries per trip: their loads and fp64 chains are independent and overlap
        for (int i = 0; i < nmax; i += 2) {
            const bool ha = i < qn, hb = i + 1 < qn;
            const int ia = ha ? sc.fifo[i * 32 + lane] : 0;
            const int ib = hb ? sc.fifo[(i + 1) * 32 + lane] : 0;
            const double4 qa = sc.candd[ia], qb = sc.candd[ib];
            const int sa = __float_as_int(sc.candf[ia].w), sb = __float_as_int(sc.candf[ib].w);
            if (ha) exact(qa, sa);
            if (hb) exact(qb, sb);
        }
        qn = 0;
    };
    auto test = [&](int i) {
        const float4 c = sc.candf[i];                              // uniform address: one broadcast LDS
        const float dx = c.x - pf.x, dy = c.y - pf.y, dz = c.z - pf.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (active && r2 <= sr2) { sc.fifo[qn * 32 + lane] = (unsigned char)i; ++qn; }
    };
    // screen the staged chunk, then drain: FIFO entries index the chunk, so it must be empty
    // before the chunk is overwritten
    auto screen_staged = [&](int n) {
        int i = 0;
        for (; i + SITE_UNROLL <= n; i += SITE_UNROLL) {
            if (__any_sync(0xffffffffu, qn > SITE_QCAP - SITE_UNROLL)) drain();
#pragma unroll
            for (int u = 0; u < SITE_UNROLL; ++u) test(i + u);
        }
        if (i < n) {
            if (__any_sync(0xffffffffu, qn > SITE_QCAP - SITE_UNROLL)) drain();
            for (; i < n; ++i) test(i);
        }
        drain();
    };

    const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
    const int wr1 = 2 * m1 + 1, wr2 = 2 * m2 + 1;
    // this work item owns the z-planes plane, plane + nplanes, ... of the stencil (normally one)
    const bool perx = g.pbc[0] != 0;
    const bool simple_x = !perx || n0 >= 2 * m0 + 1;
    const int nrows = wr1;                       // rows of one z-plane
    for (int zp = plane; zp < wr2; zp += nplanes) {
    for (int base = 0; base < nrows; base += SITE_ROWS) {
        const int row = base + lane;
        int begA = 0, endA = 0, begB = 0, endB = 0, SxB = 0, Sy = 0, Sz = 0, rowoff = 0;
        bool valid = lane < SITE_ROWS && row < nrows;
        if (valid) {
            int cy = c1 + row - m1, cz = c2 + zp - m2;
            if (g.pbc[1]) { Sy = floor_div(cy, n1); cy -= Sy * n1; }
            else if (cy < 0 || cy >= n1) valid = false;
            if (g.pbc[2]) { Sz = floor_div(cz, n2); cz -= Sz * n2; }
            else if (cz < 0 || cz >= n2) valid = false;
            if (valid) {
                rowoff = (cz * n1 + cy) * n0;
                if (simple_x) {
                    const int lo = c0 - m0, hi = c0 + m0;
                    const int a0 = max(lo, 0), a1 = min(hi, n0 - 1);
                    begA = cstart[rowoff + a0];
                    endA = cstart[rowoff + a1 + 1];
                    if (perx) {
                        if (lo < 0) { begB = cstart[rowoff + lo + n0]; endB = cstart[rowoff + n0]; SxB = -1; }
                        else if (hi >= n0) { begB = cstart[rowoff]; endB = cstart[rowoff + hi - n0 + 1]; SxB = 1; }
                    }
                }
            }
        }
        const int nr = min(SITE_ROWS, nrows - base);
        if (simple_x) {
          if (lane < SITE_ROWS) {
            // publish: slot 2*lane = piece A (Sx = 0), slot 2*lane + 1 = piece B (Sx = SxB)
            double ax, ay, az, bx, by, bz;
            shift_vec(g, 0, Sy, Sz, ax, ay, az);
            shift_vec(g, SxB, Sy, Sz, bx, by, bz);
            const int kind = valid ? 1 : 0;
            sc.pieces[2 * lane] = make_int4(begA, valid ? endA : begA, 0, kind);
            sc.pieces[2 * lane + 1] = make_int4(begB, valid ? endB : begB, 0, kind);
            sc.dshift[2 * lane][0] = ax; sc.dshift[2 * lane][1] = ay; sc.dshift[2 * lane][2] = az;
            sc.dshift[2 * lane + 1][0] = bx; sc.dshift[2 * lane + 1][1] = by; sc.dshift[2 * lane + 1][2] = bz;
            sc.rawS[2 * lane] = make_int4(0, Sy, Sz, 0);
            sc.rawS[2 * lane + 1] = make_int4(SxB, Sy, Sz, 0);
          }
            __syncwarp();
            int nst = 0;
            for (int sl = 0; sl < 2 * nr; ++sl) {
                const int4 pc = sc.pieces[sl];
                if (pc.x >= pc.y) continue;
                const float fx = (float)sc.dshift[sl][0], fy = (float)sc.dshift[sl][1], fz = (float)sc.dshift[sl][2];
                int beg = pc.x;
                while (beg < pc.y) {
                    const int take = min(SITE_CAND - nst, pc.y - beg);
                    for (int k = beg + lane; k < beg + take; k += 32) {
                        const float4 f = __ldg(&fpos[k]);
                        sc.candd[nst + (k - beg)] = spos[k];
                        sc.candf[nst + (k - beg)] = make_float4(f.x + fx, f.y + fy, f.z + fz, __int_as_float(sl));
                    }
                    nst += take;
                    beg += take;
                    if (nst == SITE_CAND) {
                        __syncwarp();
                        screen_staged(nst);
                        nst = 0;
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
            screen_staged(nst);
            __syncwarp();
        } else {
            // box thinner than the stencil along x: every offset is its own image; exact path
            // directly (rare, tiny systems)
            for (int rr = 0; rr < nr; ++rr) {
                if (!__shfl_sync(0xffffffffu, (int)valid, rr)) continue;
                const int sy = __shfl_sync(0xffffffffu, Sy, rr), sz = __shfl_sync(0xffffffffu, Sz, rr);
                const int ro = __shfl_sync(0xffffffffu, rowoff, rr);
                for (int ox = -m0; ox <= m0; ++ox) {
                    int cx = c0 + ox;
                    const int sx = floor_div(cx, n0);
                    cx -= sx * n0;
                    const int bb = cstart[ro + cx], ee = cstart[ro + cx + 1];
                    if (bb >= ee) continue;
                    if (lane == 0) {
                        sc.rawS[0] = make_int4(sx, sy, sz, 0);
                        double ax, ay, az;
                        shift_vec(g, sx, sy, sz, ax, ay, az);
                        sc.dshift[0][0] = ax; sc.dshift[0][1] = ay; sc.dshift[0][2] = az;
                    }
                    __syncwarp();
                    if (active) for (int k = bb; k < ee; ++k) exact(spos[k], 0);
                    __syncwarp();
                }
            }
        }
    }
    plane_done();
    }
}

__device__ __forceinline__ int upper_cell_of_item(const int* __restrict__ item_start, int ncells, int item) {
    int lo = 0, hi = ncells;           // item_start[lo] <= item < item_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_start[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------
// K2 (second generation): per-atom energies.  blockIdx.y = replica; warps stride over
// the replica's work items (cell, chunk of <= 32 atoms of that cell).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * SITE_WARPS, SITE_MIN_BLOCKS)
lj_cell_energy_kernel(BatchView b, LJParams p, int nplanes, double* __restrict__ e_part,
                      int* __restrict__ cnt_part) {
    __shared__ SiteScratch s_scratch[SITE_WARPS];
    const int r = blockIdx.y;
    const GridDesc& g = b.grid[r];
    const int off = g.atom_off;
    const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
    const int* item_start = cstart + (b.max_cells + 1);
    const double4* spos = b.spos + off;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n0 = g.ncell[0], n1 = g.ncell[1];
    const int nwork = g.nitems * nplanes;
    for (int wi = warp; wi < nwork; wi += nwarps) {
        const int item = wi / nplanes, plane = wi - item * nplanes;
        const int c = upper_cell_of_item(item_start, g.ncells, item);
        const int chunk = item - item_start[c];
        const int k = cstart[c] + chunk * 32 + lane;
        const bool active = k < cstart[c + 1];
        double4 me = make_double4(0.0, 0.0, 0.0, 0.0);
        if (active) me = spos[k];
        const long long tag = tag_of(me);
        const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
        double acc = 0.0, tot = 0.0;
        int cnt = 0;
        // idle lanes carry the common wrap value so they never take the slow image path
        site_rowsum(g, spos, b.fpos + off, cstart, c0, c1, c2, lane, active, me.x, me.y, me.z,
                    active ? tag_wraps(tag) : pack_wraps(0, 0, 0), active ? tag_index(tag) : -1,
                    plane, nplanes, s_scratch[threadIdx.x >> 5],
                    [&](double r2) { return r2 <= p.rc2; },
                    [&](const double4&, double, double, double, double r2) { acc += lj_pair(p, r2); ++cnt; },
                    [&]() { tot += acc; acc = 0.0; });
        if (active) {
            const size_t o = (size_t)(off + tag_index(tag)) * nplanes + plane;
            e_part[o] = tot;
            cnt_part[o] = cnt;
        }
    }
}

// e_atom[a] = 0.5 * sum_planes e_part[a][plane] (fixed order); one thread per atom
__global__ void __launch_bounds__(256)
lj_fold_planes_kernel(int total_atoms, int nplanes, const double* __restrict__ e_part,
                      const int* __restrict__ cnt_part, double* __restrict__ e_atom,
                      int* __restrict__ pair_count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total_atoms) return;
    double s = 0.0;
    int c = 0;
    for (int pl = 0; pl < nplanes; ++pl) { s += e_part[(size_t)a * nplanes + pl]; c += cnt_part[(size_t)a * nplanes + pl]; }
    e_atom[a] = 0.5 * s;
    pair_count[a] = c;
}

// ---------------------------------------------------------------------------
// K3 / K6: batched trial moves in five launches (a handful of sites: one, lj_trial_small_kernel).
//   1 trial_sites_kernel   one thread per trial: its sites' positions, bin = (replica, cell), rank in the bin
//   2 trial_scan_kernel    one CTA: exclusive scans of the bin histogram -> bin_start, work items
//   3 trial_scatter_kernel site records to their bin
//   4 lj_trial_rows_block_kernel (rows_block.cuh)  one CTA per (bin, <= 64 sites): row sums, slice by slice
//   5 lj_trial_combine_kernel  one thread per trial: sum(new rows) - sum(old rows) + moved-set pairs
// ---------------------------------------------------------------------------
#define TRIAL_HIST_STRIDE 8

struct TrialWork {
    int nsites;            // n_old_total + n_new_total; sites [0, n_old) are old, the rest new
    int n_old;
    int n_new;
    int nbins;             // R * bin_stride
    int bin_stride;        // max_cells
    int* site_trial;       // [nsites]
    int4* site_a;          // [nsites] (bin, rank inside the bin, first excluded atom or -1, old_off[t])
    double4* site_pos;     // [nsites] (x, y, z, .w = packed wraps | moved-atom count << 32 | stamp << 38)
    int stamp;             // 1 .. 2^26-1, new for every launch: sites left over from earlier launches are ignored
    int* hist;             // [nbins * TRIAL_HIST_STRIDE] site histogram, one 32-byte sector per bin (atomics on
                           // neighbouring bins of one sector serialise in L2: 8 us of a 19 us kernel); zero on
                           // entry, re-zeroed by the scan kernel once it has read it
    int* bin_count;        // [nbins + 1]  bin_start (exclusive prefix of the histogram)
    int* item_start;       // [nbins + 1]
    int* sorted;           // [nsites] site ids grouped by bin
    int nplanes;           // partial sums per site: stencil z-planes (warp kernel) or list slices (CTA kernel)
    int item_slices;       // slices one work item of the CTA kernel covers (divides nplanes)
    int item_shift;        // sites per work item = 1 << item_shift (5: warp items, RB_SITES_SHIFT: CTA items)
    double* site_E;        // [nsites][nplanes] partial row sums
    int* site_cnt;         // [nsites][nplanes] in-cutoff pairs
    int* totals;           // [0] = number of work items, [1] = work counter of the persistent row-sum CTAs
    // device-side validation: first offending trial << 8 | code (1 replica, 2 list length / offsets,
    // 3 atom index); ~0 = none.  The scan kernel moves it to err_out and re-arms it.
    unsigned long long* err;
    unsigned long long* err_out;   // may be nullptr
    // CTA row-sum kernel (rows_block.cuh): its prologue reads these instead of chasing pointers
    int2* item_tab;        // [items] (bin, group of the bin's sites); nullptr: not wanted
    // warp row-sum kernel (rows_warp.cuh)
    int4* items;           // [2 * items] (bin, first site in bin order, sites, replica), (list first, list length, -, -)
    unsigned long long* chain;   // [2 R] per-replica totals of the chained scan: (stamp << 36) | sites, | items
    double4* sorted_pos;   // [nsites] site_pos in bin order (.w = packed wraps)
    int4* sorted_meta;     // [nsites] (site id, first excluded atom or -1, old_off[t], old_off[t+1]) in bin order
    // direct placement (dense batches; trial_sites_direct_kernel): slot_cap > 0 means sorted_pos / sorted_meta hold
    // slot_cap records per bin -- the site that drew rank k in the histogram of bin c sits at c * slot_cap + k -- and
    // item_tab lists (bin, group of 16 ranks) in the order the groups were opened (totals[0] of them), so neither the
    // scan nor the scatter runs; sites beyond a bin's capacity go to site_pos / site_a [0 .. ovf[0])
    int slot_cap;
    int* ovf;
};

__device__ __forceinline__ int owner_of(const int* __restrict__ off, int T, int k) {
    int lo = 0, hi = T;                 // off[lo] <= k < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= k) lo = mid; else hi = mid;
    }
    return lo;
}

// One thread per trial: its old sites are [old_off[t], old_off[t+1]) and its new sites
// n_old + [new_off[t], new_off[t+1]) in the global site numbering.  The histogram increment
// returns the site's rank inside its bin, so the scatter needs no second atomic.  Arguments are
// validated here (the host variant used to spend 330 us per 65 536 trials doing it): an invalid
// trial records itself in *w.err and contributes no site.
// 0: trial t is well formed; 1: replica out of range; 2: bad offsets / list lengths; 3: atom index.
__device__ __forceinline__ int trial_invalid(const BatchView& b, const TrialBatch& tb, const TrialWork& w,
                                             int r, int ob, int oe, int nb, int ne) {
    if (r < 0 || r >= b.R) return 1;
    if (ob < 0 || oe < ob || oe > w.n_old || nb < 0 || ne < nb || ne > w.n_new ||
        oe - ob > MCPYB_MAX_MOVED || ne - nb > MCPYB_MAX_MOVED) return 2;
    const int n = b.grid[r].n;
    int bad = 0;
    for (int i = ob; i < oe; ++i) { const int a = tb.old_idx[i]; if (a < 0 || a >= n) bad = 3; }
    return bad;
}

__device__ __forceinline__ void trial_sites_body(const BatchView& b, const TrialBatch& tb, const TrialWork& w, int t) {
    const int r = tb.rep ? tb.rep[t] : 0;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    const int bad = trial_invalid(b, tb, w, r, ob, oe, nb, ne);
    if (bad) {
        atomicMin(w.err, ((unsigned long long)t << 8) | (unsigned long long)bad);
        return;
    }
    const GridDesc& g = b.grid[r];
    const int ex0 = oe > ob ? tb.old_idx[ob] : -1;
    const long long cnt_bits = ((long long)(oe - ob) << 32) | ((long long)w.stamp << 38);
    auto bin_of = [&](double x, double y, double z, int wr[3]) {
        int c[3];
        locate(g, x, y, z, c, wr);
        return r * w.bin_stride + (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
    };
    auto record = [&](int s, int bin, int rank, double x, double y, double z, const int wr[3]) {
        w.site_trial[s] = t;
        w.site_a[s] = make_int4(bin, rank, ex0, ob);
        w.site_pos[s] = make_double4(x, y, z, __longlong_as_double((long long)pack_wraps(wr[0], wr[1], wr[2]) | cnt_bits));
    };
    for (int i = ob; i < oe + (ne - nb); ++i) {
        int s;
        double x, y, z;
        if (i < oe) {
            s = i;
            const double4 q = b.upos[g.atom_off + tb.old_idx[i]];
            x = q.x; y = q.y; z = q.z;
        } else {
            const int k = nb + (i - oe);
            s = w.n_old + k;
            x = tb.new_pos[3 * k]; y = tb.new_pos[3 * k + 1]; z = tb.new_pos[3 * k + 2];
        }
        int wr[3];
        const int bin = bin_of(x, y, z, wr);
        const int rank = atomicAdd(&w.hist[(size_t)bin * TRIAL_HIST_STRIDE], 1);
        record(s, bin, rank, x, y, z, wr);
    }
}

// Direct placement: the rank a site draws from its bin's histogram IS its slot.  Record formats are those of
// sorted_pos / sorted_meta; the bin rides in the upper half of .w (the row-sum kernel reads the lower one).
// The two sites of an ordinary trial (one removed, one added atom) are handled side by side, so their gathers and
// their atomics are in flight together.
__device__ __forceinline__ void trial_place(const TrialWork& w, int s, int bin, int rank, double x, double y, double z,
                                            const int wr[3], int ex0, int ob, int oe) {
    const double4 rp = make_double4(x, y, z, __longlong_as_double((long long)(unsigned int)pack_wraps(wr[0], wr[1], wr[2]) |
                                                                  ((long long)bin << 32)));
    const int4 rm = make_int4(s, ex0, ob, oe);
    MCPYB_CHECK(bin >= 0 && bin < w.nbins && rank >= 0);
    if (rank < w.slot_cap) {
        const size_t k = (size_t)bin * w.slot_cap + rank;
        w.sorted_pos[k] = rp;
        w.sorted_meta[k] = rm;
        // the first site of every group of 16 announces the group as a work item of the row-sum kernel
        if ((rank & ((1 << 4) - 1)) == 0) w.item_tab[atomicAdd(&w.totals[0], 1)] = make_int2(bin, rank >> 4);
    } else {
        const int k = atomicAdd(w.ovf, 1);
        w.site_pos[k] = rp;
        w.site_a[k] = rm;
    }
}

// (Two trials per thread, walked phase by phase so that the loads and atomics of both are in flight together, were
// measured slower: 59 us against 41 us for 524 288 trials -- 74 registers halve the resident warps.  47 % of this
// kernel's stall samples sit on the histogram atomic, waiting for the chain offsets -> index -> position -> bin.)
__global__ void __launch_bounds__(256)
trial_sites_direct_kernel(BatchView b, TrialBatch tb, TrialWork w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (t == 0) w.totals[2] = 0;                       // "some trial needs the general combine": raised by the combine kernel
    if (t >= tb.T) return;
    const int r = tb.rep ? tb.rep[t] : 0;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    const int bad = trial_invalid(b, tb, w, r, ob, oe, nb, ne);
    if (bad) {
        atomicMin(w.err, ((unsigned long long)t << 8) | (unsigned long long)bad);
        return;
    }
    const GridDesc& g = b.grid[r];
    const int ex0 = oe > ob ? tb.old_idx[ob] : -1;
    auto bin_of = [&](double x, double y, double z, int wr[3]) {
        int c[3];
        locate(g, x, y, z, c, wr);
        return r * w.bin_stride + (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
    };
    const int ko = oe - ob, kn = ne - nb;
    if (ko <= 1 && kn <= 1) {
        double xo = 0.0, yo = 0.0, zo = 0.0, xn = 0.0, yn = 0.0, zn = 0.0;
        if (ko) { const double4 q = b.upos[g.atom_off + ex0]; xo = q.x; yo = q.y; zo = q.z; }
        if (kn) { xn = tb.new_pos[3 * nb]; yn = tb.new_pos[3 * nb + 1]; zn = tb.new_pos[3 * nb + 2]; }
        int wo[3] = {0, 0, 0}, wn[3] = {0, 0, 0}, bo = 0, bn = 0, ro = 0, rn = 0;
        if (ko) bo = bin_of(xo, yo, zo, wo);
        if (kn) bn = bin_of(xn, yn, zn, wn);
        if (ko) ro = atomicAdd(&w.hist[(size_t)bo * TRIAL_HIST_STRIDE], 1);
        if (kn) rn = atomicAdd(&w.hist[(size_t)bn * TRIAL_HIST_STRIDE], 1);
        if (ko) trial_place(w, ob, bo, ro, xo, yo, zo, wo, ex0, ob, oe);
        if (kn) trial_place(w, w.n_old + nb, bn, rn, xn, yn, zn, wn, ex0, ob, oe);
        return;
    }
    for (int i = ob; i < oe + kn; ++i) {
        int s;
        double x, y, z;
        if (i < oe) {
            s = i;
            const double4 q = b.upos[g.atom_off + tb.old_idx[i]];
            x = q.x; y = q.y; z = q.z;
        } else {
            const int k = nb + (i - oe);
            s = w.n_old + k;
            x = tb.new_pos[3 * k]; y = tb.new_pos[3 * k + 1]; z = tb.new_pos[3 * k + 2];
        }
        int wr[3];
        const int bin = bin_of(x, y, z, wr);
        const int rank = atomicAdd(&w.hist[(size_t)bin * TRIAL_HIST_STRIDE], 1);
        trial_place(w, s, bin, rank, x, y, z, wr, ex0, ob, oe);
    }
}

__global__ void __launch_bounds__(256)
trial_sites_kernel(BatchView b, TrialBatch tb, TrialWork w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (t == 0) w.totals[2] = 0;                       // "some trial needs the general combine": raised by the combine kernel
    if (t >= tb.T) return;
    trial_sites_body(b, tb, w, t);
}

// One CTA of NT threads: exclusive scans of the bin histogram -> bin starts, work items; re-zeroes the histogram.
template <int NT>
__device__ __forceinline__ void trial_scan_body(const TrialWork& w, int* s_a, int* s_b) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // each thread owns a contiguous run of bins; the counts of short runs (the usual case: two bins per
    // thread for a 2 000-atom box) stay in registers, so the histogram is read once
    const int per = (w.nbins + NT - 1) / NT;
    const int beg = min(tid * per, w.nbins), end = min(beg + per, w.nbins);
    unsigned long long ev = ~0ULL;
    if (tid == 0) ev = *w.err;                       // independent of the scan: in flight alongside it
    int sa = 0, sb = 0, c4[4] = {0, 0, 0, 0};
    const int ish = w.item_shift, iround = (1 << ish) - 1;
    for (int i = beg; i < end; ++i) {
        const int c = w.hist[(size_t)i * TRIAL_HIST_STRIDE];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i - beg == k) c4[k] = c;
        sa += c; sb += (c + iround) >> ish;
    }
    // block-wide exclusive scan of (sa, sb): shuffles inside the warps, one warp over the warp totals
    int ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb2 = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb2; }
    }
    if (lane == 31) { s_a[wid] = ia; s_b[wid] = ib; }
    __syncthreads();
    if (wid == 0) {
        int va = lane < NT / 32 ? s_a[lane] : 0, vb = lane < NT / 32 ? s_b[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, va, o), tb2 = __shfl_up_sync(0xffffffffu, vb, o);
            if (lane >= o) { va += ta; vb += tb2; }
        }
        if (lane < NT / 32) { s_a[lane] = va; s_b[lane] = vb; }        // inclusive totals of warps 0 .. lane
    }
    __syncthreads();
    int ra = (wid ? s_a[wid - 1] : 0) + ia - sa, rb = (wid ? s_b[wid - 1] : 0) + ib - sb;   // exclusive prefix of this run
    for (int i = beg; i < end; ++i) {
        int c = 0;
        if (i - beg < 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i - beg == k) c = c4[k];
        } else {
            c = w.hist[(size_t)i * TRIAL_HIST_STRIDE];
        }
        if (c) w.hist[(size_t)i * TRIAL_HIST_STRIDE] = 0;
        w.bin_count[i] = ra; w.item_start[i] = rb;
        const int ni = (c + iround) >> ish;
        if (w.item_tab) for (int gq = 0; gq < ni; ++gq) w.item_tab[rb + gq] = make_int2(i, gq);
        ra += c; rb += ni;
    }
    if (tid == NT - 1) {
        const int ta = s_a[NT / 32 - 1], tb2 = s_b[NT / 32 - 1];
        w.bin_count[w.nbins] = ta; w.item_start[w.nbins] = tb2; w.totals[0] = tb2; w.totals[1] = 0;
    }
    if (tid == 0) {
        if (w.err_out) *w.err_out = ev;
        *w.err = ~0ULL;
    }
}

__global__ void __launch_bounds__(1024, 1)
trial_scan_kernel(TrialWork w) {
    __shared__ int s_a[1024], s_b[1024];
    pdl_wait();
    pdl_trigger();
    trial_scan_body<1024>(w, s_a, s_b);
}

__device__ __forceinline__ void trial_scatter_body(const TrialWork& w, int s) {
    const int4 a = w.site_a[s];
    const double4 ps = w.site_pos[s];
    const long long wbits = __double_as_longlong(ps.w);
    if ((int)((unsigned long long)wbits >> 38) != w.stamp) return;   // not written by this launch (rejected trial)
    const int count = (int)(wbits >> 32) & 63;
    MCPYB_CHECK(a.x >= 0 && a.x < w.nbins);
    const int k = w.bin_count[a.x] + a.y;
    MCPYB_CHECK(k >= 0 && k < w.nsites);
    w.sorted[k] = s;
    if (w.sorted_pos) {
        w.sorted_pos[k] = make_double4(ps.x, ps.y, ps.z, __longlong_as_double(wbits & 0xFFFFFFFFLL));
        w.sorted_meta[k] = make_int4(s, a.z, a.w, a.w + count);
    }
}

__global__ void __launch_bounds__(256)
trial_scatter_kernel(TrialBatch tb, TrialWork w) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (s >= w.nsites) return;
    trial_scatter_body(w, s);
}

// dE of one trial from its sites' partial row sums: sum(new rows) - sum(old rows) + moved-set pairs.
__device__ __forceinline__ void combine_trial(const BatchView& b, const TrialBatch& tb, const TrialWork& w,
                                              const LJParams& p, int t, double* __restrict__ dE,
                                              int* __restrict__ npairs) {
    const int r = tb.rep ? tb.rep[t] : 0;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    if (trial_invalid(b, tb, w, r, ob, oe, nb, ne)) {              // rejected by trial_sites_kernel
        dE[t] = __longlong_as_double(0x7ff8000000000000LL);
        if (npairs) npairs[t] = 0;
        return;
    }
    const GridDesc& g = b.grid[r];
    const int off = g.atom_off;
    double plus = 0.0, minus = 0.0;
    int cnt = 0;
    auto row_of = [&](int sidx, double& e) {           // planes in fixed order
        double row = 0.0;
        for (int pl = 0; pl < w.nplanes; ++pl) {
            row += w.site_E[(size_t)sidx * w.nplanes + pl];
            cnt += w.site_cnt[(size_t)sidx * w.nplanes + pl];
        }
        e += row;
    };
    for (int k = nb; k < ne; ++k) row_of(w.n_old + k, plus);
    for (int k = ob; k < oe; ++k) row_of(k, minus);
    // pairs inside the moved sets (and their periodic self images); a single moved atom can only
    // see its own image when the box is thinner than the cutoff
    if (!g.thin && (ne - nb) <= 1 && (oe - ob) <= 1) {
        dE[t] = plus - minus;
        if (npairs) npairs[t] = cnt;
        return;
    }
    intra_set_lj(g, p, ne - nb,
        [&](int a, double& x, double& y, double& z) {
            x = tb.new_pos[3 * (nb + a)]; y = tb.new_pos[3 * (nb + a) + 1]; z = tb.new_pos[3 * (nb + a) + 2];
        }, 0, 1, plus, cnt);
    intra_set_lj(g, p, oe - ob,
        [&](int a, double& x, double& y, double& z) {
            const double4 q = b.upos[off + tb.old_idx[ob + a]];
            x = q.x; y = q.y; z = q.z;
        }, 0, 1, minus, cnt);
    dE[t] = plus - minus;
    if (npairs) npairs[t] = cnt;
}

// The batched combine in two launches: the first answers the common case (one removed and / or one added atom in a box
// wider than the cutoff) from a dozen registers at full occupancy and raises w.totals[2] when it meets anything else;
// the second, which returns at once when the flag is down, runs combine_trial on exactly those trials.
__device__ __forceinline__ bool combine_simple(const BatchView& b, const TrialBatch& tb, const TrialWork& w, int t,
                                               double* __restrict__ dE, int* __restrict__ npairs) {
    // every load issued before the first use: dE = row(new) - row(old)
    const int r = tb.rep ? tb.rep[t] : 0;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    const int ko = oe - ob, kn = ne - nb;
    if (w.nplanes == 1 && r >= 0 && r < b.R && ob >= 0 && nb >= 0 && oe <= w.n_old && ne <= w.n_new &&
        ko >= 0 && ko <= 1 && kn >= 0 && kn <= 1) {
        const int a = ko ? tb.old_idx[ob] : 0;
        const double e_old = ko ? w.site_E[ob] : 0.0, e_new = kn ? w.site_E[w.n_old + nb] : 0.0;
        const int c_old = ko ? w.site_cnt[ob] : 0, c_new = kn ? w.site_cnt[w.n_old + nb] : 0;
        const GridDesc& g = b.grid[r];
        if (!g.thin && a >= 0 && a < g.n) {
            if (dE) {
                // (0.0 + e_new) - (0.0 + e_old): the sums combine_trial forms, term for term
                dE[t] = (0.0 + e_new) - (0.0 + e_old);
                if (npairs) npairs[t] = c_new + c_old;
            }
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(256, 8)
lj_trial_combine_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p,
                        double* __restrict__ dE, int* __restrict__ npairs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (w.slot_cap > 0 && t == 0) {
        // direct placement has no scan kernel: the validation word is published, and the counters re-armed, here
        const unsigned long long ev = *w.err;
        if (w.err_out) *w.err_out = ev;
        *w.err = ~0ULL;
        w.totals[0] = 0;
        w.totals[1] = 0;
        w.ovf[0] = 0;
    }
    if (t >= tb.T) return;
    if (!combine_simple(b, tb, w, t, dE, npairs)) w.totals[2] = 1;        // (same value from every writer)
}

__global__ void __launch_bounds__(128)
lj_trial_combine_rest_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p,
                             double* __restrict__ dE, int* __restrict__ npairs) {
    pdl_wait();
    pdl_trigger();
    if (w.totals[2] == 0) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < tb.T; t += gridDim.x * blockDim.x)
        if (!combine_simple(b, tb, w, t, nullptr, nullptr)) combine_trial(b, tb, w, p, t, dE, npairs);
}

// One launch for everything (the single-configuration chain and other small batches).
__global__ void __launch_bounds__(128)
lj_trial_combine_all_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p,
                            double* __restrict__ dE, int* __restrict__ npairs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (t >= tb.T) return;
    combine_trial(b, tb, w, p, t, dE, npairs);
}
