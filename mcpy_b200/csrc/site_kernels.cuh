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

// Row sum of one site per lane over the stencil of cell (c0,c1,c2), two phases:
//
//   screen : every candidate of the stencil is tested in fp32 on the wrapped, origin-relative
//            copies (fpos) against (rcut + margin)^2.  The margin (cell_list.cuh) bounds the fp32
//            rounding, so the screen never drops a pair the exact test would keep.  Candidates
//            that pass are appended to the lane's FIFO in shared memory (index only).
//   exact  : when a FIFO is nearly full (and at the end) all lanes drain theirs: fp64 image
//            vector in the oracle's operation order, exact cutoff test, exclusions, pair energy,
//            accumulation in candidate order.
//
// The screen keeps the warp converged on cheap single-issue instructions; the expensive fp64 path
// runs on ~22 % of the candidates with most lanes busy instead of diverging on every candidate.
//   active : this lane holds a site
//   excl(j): true if atom j must be skipped in EVERY image (moved atoms)
//   self   : atom index to skip in the zero-shift image only (-1: none)
#define SITE_QCAP 32              // FIFO entries per lane
#define SITE_UNROLL 4
#define SITE_WARPS 4               // warps per block of the row-sum kernels

template <class Excl>
__device__ __forceinline__ void lj_site_rowsum(const GridDesc& g, const double4* __restrict__ spos,
                                               const float4* __restrict__ fpos,
                                               const int* __restrict__ cstart, const LJParams& p,
                                               int c0, int c1, int c2, int lane, bool active,
                                               double px, double py, double pz, int wsite, int self,
                                               int* __restrict__ fifo /* [SITE_QCAP][32] of this warp */,
                                               int4* __restrict__ pieces /* [64] */, float4* __restrict__ shifts /* [64] */,
                                               int4* __restrict__ rawS /* [32] */,
                                               Excl&& excl, double& acc, int& cnt) {
    int w0, w1, w2;
    unpack_wraps(wsite, w0, w1, w2);
    const float3 pf = screen_pos(g, px, py, pz, w0, w1, w2);
    const float sr2 = g.screen_r2;
    int qn = 0;

    // drain this lane's FIFO: entries are (candidate slot k, packed stencil shift)
    auto drain = [&]() {
        const int nmax = __reduce_max_sync(0xffffffffu, qn);
        for (int i = 0; i < nmax; ++i) {
            if (i < qn) {
                const int ent = fifo[i * 32 + lane];
                const int k = ent & 0xFFFFFF;
                const int S0 = ((ent >> 24) & 7) - 3, S1 = ((ent >> 27) & 7) - 3;
                int S2 = (int)(((unsigned)ent) >> 30);     // 2 bits: 0,1,2 -> -1,0,+1 (3: escape)
                const double4 q = spos[k];
                const long long tag = tag_of(q);
                const int j = tag_index(tag);
                int jx, jy, jz;
                unpack_wraps(tag_wraps(tag), jx, jy, jz);
                S2 -= 1;
                const int T0 = S0 + w0 - jx, T1 = S1 + w1 - jy, T2 = S2 + w2 - jz;
                double r2, dx, dy, dz;
                if ((T0 | T1 | T2) == 0) {          // q + 0.0 == q: skip the three additions, same bits
                    dx = __dsub_rn(q.x, px); dy = __dsub_rn(q.y, py); dz = __dsub_rn(q.z, pz);
                    r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    if (r2 <= p.rc2 && j != self && !excl(j)) { acc += lj_pair(p, r2); ++cnt; }
                } else {
                    double ax, ay, az;
                    shift_vec(g, T0, T1, T2, ax, ay, az);
                    r2 = image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
                    if (r2 <= p.rc2 && !excl(j)) { acc += lj_pair(p, r2); ++cnt; }
                }
            }
        }
        qn = 0;
    };

    // exact path for one piece whose shift does not fit the packed FIFO entry (very thin boxes)
    auto exact_piece = [&](int beg, int end, int S0, int S1, int S2) {
        drain();
        for (int k = beg; k < end; ++k) {
            const double4 q = spos[k];
            const long long tag = tag_of(q);
            const int j = tag_index(tag);
            int jx, jy, jz;
            unpack_wraps(tag_wraps(tag), jx, jy, jz);
            const int T0 = S0 + w0 - jx, T1 = S1 + w1 - jy, T2 = S2 + w2 - jz;
            double ax, ay, az, dx, dy, dz;
            shift_vec(g, T0, T1, T2, ax, ay, az);
            const double r2 = image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
            const bool same = (T0 | T1 | T2) == 0;
            if (active && r2 <= p.rc2 && !(same && j == self) && !excl(j)) { acc += lj_pair(p, r2); ++cnt; }
        }
    };
    // fp32 screen of one piece; (fx,fy,fz) = float(S @ cell), spack = packed S for the FIFO
    auto screen_piece = [&](int beg, int end, float fx, float fy, float fz, int spack) {
        const float qx = pf.x - fx, qy = pf.y - fy, qz = pf.z - fz;    // d = f - (pf - s)
        auto test = [&](int k) {
            const float4 f = __ldg(&fpos[k]);                      // uniform address: one broadcast
            const float dx = f.x - qx, dy = f.y - qy, dz = f.z - qz;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (active && r2 <= sr2) { fifo[qn * 32 + lane] = k | spack; ++qn; }
        };
        int k = beg;
        for (; k + SITE_UNROLL <= end; k += SITE_UNROLL) {
            if (__any_sync(0xffffffffu, qn > SITE_QCAP - SITE_UNROLL)) drain();
#pragma unroll
            for (int u = 0; u < SITE_UNROLL; ++u) test(k + u);
        }
        if (k < end) {
            if (__any_sync(0xffffffffu, qn > SITE_QCAP - SITE_UNROLL)) drain();
            for (; k < end; ++k) test(k);
        }
    };

    // ---- stencil rows: one lane per row works out its (at most two) contiguous pieces and
    // publishes them in shared memory; the warp then walks the list with broadcast reads.
    const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
    const int wr1 = 2 * m1 + 1, wr2 = 2 * m2 + 1;
    const int nrows = wr1 * wr2;
    const bool perx = g.pbc[0] != 0;
    const bool simple_x = !perx || n0 >= 2 * m0 + 1;
    for (int base = 0; base < nrows; base += 32) {
        const int row = base + lane;
        int begA = 0, endA = 0, begB = 0, endB = 0, SxB = 0, Sy = 0, Sz = 0, rowoff = 0;
        bool valid = row < nrows;
        if (valid) {
            int cy = c1 + (row % wr1) - m1, cz = c2 + (row / wr1) - m2;
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
        const bool packable = (Sy >= -3 && Sy <= 3 && Sz >= -1 && Sz <= 1);
        if (simple_x) {
            // publish: slot 2*lane = piece A (Sx = 0), slot 2*lane+1 = piece B (Sx = SxB)
            double ax, ay, az, bx, by, bz;
            shift_vec(g, 0, Sy, Sz, ax, ay, az);
            shift_vec(g, SxB, Sy, Sz, bx, by, bz);
            const int flag = valid ? (packable ? 1 : 2) : 0;
            pieces[2 * lane] = make_int4(begA, (valid ? endA : begA), ((0 + 3) << 24) | ((Sy + 3) << 27) | ((Sz + 1) << 30), flag);
            pieces[2 * lane + 1] = make_int4(begB, (valid ? endB : begB), ((SxB + 3) << 24) | ((Sy + 3) << 27) | ((Sz + 1) << 30), flag);
            shifts[2 * lane] = make_float4((float)ax, (float)ay, (float)az, 0.f);
            shifts[2 * lane + 1] = make_float4((float)bx, (float)by, (float)bz, 0.f);
            // raw S for the unpackable case
            rawS[lane] = make_int4(SxB, Sy, Sz, 0);
            __syncwarp();
            const int nr = min(32, nrows - base);
            for (int sl = 0; sl < 2 * nr; ++sl) {
                const int4 pc = pieces[sl];
                if (pc.x >= pc.y) continue;
                if (pc.w == 1) {
                    const float4 sh = shifts[sl];
                    screen_piece(pc.x, pc.y, sh.x, sh.y, sh.z, pc.z);
                } else {
                    const int4 rs = rawS[sl >> 1];
                    exact_piece(pc.x, pc.y, (sl & 1) ? rs.x : 0, rs.y, rs.z);
                }
            }
            __syncwarp();
        } else {
            // box thinner than the stencil along x: every offset is its own image (rare, exact path)
            const int nr = min(32, nrows - base);
            for (int rr = 0; rr < nr; ++rr) {
                if (!__shfl_sync(0xffffffffu, (int)valid, rr)) continue;
                const int sy = __shfl_sync(0xffffffffu, Sy, rr), sz = __shfl_sync(0xffffffffu, Sz, rr);
                const int ro = __shfl_sync(0xffffffffu, rowoff, rr);
                for (int ox = -m0; ox <= m0; ++ox) {
                    int cx = c0 + ox;
                    const int sx = floor_div(cx, n0);
                    cx -= sx * n0;
                    const int bb = cstart[ro + cx], ee = cstart[ro + cx + 1];
                    if (bb < ee) exact_piece(bb, ee, sx, sy, sz);
                }
            }
        }
    }
    drain();
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
__global__ void __launch_bounds__(32 * SITE_WARPS)
lj_cell_energy_kernel(BatchView b, LJParams p, double* __restrict__ e_atom, int* __restrict__ pair_count) {
    __shared__ int s_fifo[SITE_WARPS][SITE_QCAP * 32];
    __shared__ int4 s_pieces[SITE_WARPS][64];
    __shared__ float4 s_shifts[SITE_WARPS][64];
    __shared__ int4 s_rawS[SITE_WARPS][32];
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
    for (int item = warp; item < g.nitems; item += nwarps) {
        const int c = upper_cell_of_item(item_start, g.ncells, item);
        const int chunk = item - item_start[c];
        const int k = cstart[c] + chunk * 32 + lane;
        const bool active = k < cstart[c + 1];
        double4 me = make_double4(0.0, 0.0, 0.0, 0.0);
        if (active) me = spos[k];
        const long long tag = tag_of(me);
        const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
        double acc = 0.0;
        int cnt = 0;
        // idle lanes carry the common wrap value so they never take the slow image path
        lj_site_rowsum(g, spos, b.fpos + off, cstart, p, c0, c1, c2, lane, active, me.x, me.y, me.z,
                       active ? tag_wraps(tag) : pack_wraps(0, 0, 0), active ? tag_index(tag) : -1,
                       s_fifo[threadIdx.x >> 5], s_pieces[threadIdx.x >> 5], s_shifts[threadIdx.x >> 5],
                       s_rawS[threadIdx.x >> 5], [](int) { return false; }, acc, cnt);
        if (active) {
            e_atom[off + tag_index(tag)] = 0.5 * acc;
            if (pair_count) pair_count[off + tag_index(tag)] = cnt;
        }
    }
}

// ---------------------------------------------------------------------------
// K3 / K6 (second generation): batched trial moves in five small launches.
//   1 trial_sites_kernel   one thread per site: position, owning trial, bin = (replica, cell)
//   2 trial_scan_kernel    one CTA: exclusive scans of the bin histogram -> bin_start, item_start
//   3 trial_scatter_kernel sites to their bin
//   4 lj_trial_rows_kernel one warp per (bin, chunk of 32 sites): row sums, one site per lane
//   5 lj_trial_combine_kernel  one thread per trial: sum(new rows) - sum(old rows) + moved-set pairs
// ---------------------------------------------------------------------------
struct TrialWork {
    int nsites;            // n_old_total + n_new_total; sites [0, n_old) are old, the rest new
    int n_old;
    int nbins;             // R * bin_stride
    int bin_stride;        // max_cells
    int* site_trial;       // [nsites]
    int* site_bin;         // [nsites]
    double4* site_pos;     // [nsites] (x, y, z, packed wraps in .w)
    int* bin_count;        // [nbins + 1]  histogram, then bin_start (exclusive)
    int* item_start;       // [nbins + 1]
    int* bin_cursor;       // [nbins]
    int* sorted;           // [nsites] site ids grouped by bin
    double* site_E;        // [nsites] row sums
    int* site_cnt;         // [nsites] in-cutoff pairs of the row
    int* totals;           // [0] = number of work items
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
// n_old + [new_off[t], new_off[t+1]) in the global site numbering.
__global__ void __launch_bounds__(256)
trial_sites_kernel(BatchView b, TrialBatch tb, TrialWork w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tb.T) return;
    const int r = tb.rep ? tb.rep[t] : 0;
    const GridDesc& g = b.grid[r];
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
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
        int c[3], wr[3];
        locate(g, x, y, z, c, wr);
        const int bin = r * w.bin_stride + (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
        w.site_trial[s] = t;
        w.site_bin[s] = bin;
        w.site_pos[s] = make_double4(x, y, z, __longlong_as_double((long long)pack_wraps(wr[0], wr[1], wr[2])));
        atomicAdd(&w.bin_count[bin], 1);
    }
}

__global__ void __launch_bounds__(1024, 1)
trial_scan_kernel(TrialWork w) {
    __shared__ int s_a[1024], s_b[1024];
    const int tid = threadIdx.x;
    // each thread owns a contiguous run of bins so the scan is two passes over registers/global
    const int per = (w.nbins + 1023) / 1024;
    const int beg = min(tid * per, w.nbins), end = min(beg + per, w.nbins);
    int sa = 0, sb = 0;
    for (int i = beg; i < end; ++i) { const int c = w.bin_count[i]; sa += c; sb += (c + 31) >> 5; }
    s_a[tid] = sa; s_b[tid] = sb;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int ta = (tid >= o) ? s_a[tid - o] : 0, tb2 = (tid >= o) ? s_b[tid - o] : 0;
        __syncthreads();
        s_a[tid] += ta; s_b[tid] += tb2;
        __syncthreads();
    }
    int ra = s_a[tid] - sa, rb = s_b[tid] - sb;      // exclusive prefix of this thread's run
    for (int i = beg; i < end; ++i) {
        const int c = w.bin_count[i];
        w.bin_count[i] = ra; w.item_start[i] = rb; w.bin_cursor[i] = 0;
        ra += c; rb += (c + 31) >> 5;
    }
    if (tid == 1023) { w.bin_count[w.nbins] = s_a[1023]; w.item_start[w.nbins] = s_b[1023]; w.totals[0] = s_b[1023]; }
}

__global__ void __launch_bounds__(256)
trial_scatter_kernel(TrialWork w) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= w.nsites) return;
    const int bin = w.site_bin[s];
    w.sorted[w.bin_count[bin] + atomicAdd(&w.bin_cursor[bin], 1)] = s;
}

__global__ void __launch_bounds__(32 * SITE_WARPS)
lj_trial_rows_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p) {
    __shared__ int s_fifo[SITE_WARPS][SITE_QCAP * 32];
    __shared__ int4 s_pieces[SITE_WARPS][64];
    __shared__ float4 s_shifts[SITE_WARPS][64];
    __shared__ int4 s_rawS[SITE_WARPS][32];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nitems = w.totals[0];
    for (int item = warp; item < nitems; item += nwarps) {
        const int bin = upper_cell_of_item(w.item_start, w.nbins, item);
        const int chunk = item - w.item_start[bin];
        const int r = bin / w.bin_stride, c = bin - r * w.bin_stride;
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        const int k = w.bin_count[bin] + chunk * 32 + lane;
        const bool active = k < w.bin_count[bin + 1];
        int s = 0, ob = 0, oe = 0, ex0 = -1;
        double4 me = make_double4(0.0, 0.0, 0.0, 0.0);
        if (active) {
            s = w.sorted[k];
            me = w.site_pos[s];
            const int t = w.site_trial[s];
            ob = tb.old_off[t]; oe = tb.old_off[t + 1];
            if (oe > ob) ex0 = tb.old_idx[ob];
        }
        const int n0 = g.ncell[0], n1 = g.ncell[1];
        const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
        const int wsite = active ? (int)__double_as_longlong(me.w) : pack_wraps(0, 0, 0);
        double acc = 0.0;
        int cnt = 0;
        const bool multi = (oe - ob) > 1;
        lj_site_rowsum(g, b.spos + off, b.fpos + off, cstart, p, c0, c1, c2, lane, active, me.x, me.y, me.z,
                       wsite, -1, s_fifo[threadIdx.x >> 5], s_pieces[threadIdx.x >> 5], s_shifts[threadIdx.x >> 5],
                       s_rawS[threadIdx.x >> 5],
                       [&](int j) {
                           if (j == ex0) return true;
                           if (multi) for (int e = ob + 1; e < oe; ++e) if (tb.old_idx[e] == j) return true;
                           return false;
                       }, acc, cnt);
        if (active) { w.site_E[s] = acc; w.site_cnt[s] = cnt; }
    }
}

__global__ void __launch_bounds__(128)
lj_trial_combine_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p,
                        double* __restrict__ dE, int* __restrict__ npairs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tb.T) return;
    const int r = tb.rep ? tb.rep[t] : 0;
    const GridDesc& g = b.grid[r];
    const int off = g.atom_off;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    double plus = 0.0, minus = 0.0;
    int cnt = 0;
    for (int k = nb; k < ne; ++k) { plus += w.site_E[w.n_old + k]; cnt += w.site_cnt[w.n_old + k]; }
    for (int k = ob; k < oe; ++k) { minus += w.site_E[k]; cnt += w.site_cnt[k]; }
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
