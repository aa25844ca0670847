// K3 / K6, third generation: trial-move row sums, one CTA per (cell, group of <= RB_SITES sites).
//
// The second-generation kernel (site_kernels.cuh: fp32 screen + per-lane FIFO + exact drain) was
// bound by instruction issue and latency, not by a math pipe: 20 k warp instructions per 32 sites,
// issue-active 52-57 %, FP64 pipe 15 % (profiles/r1_v8_*).  A tighter FIFO version of it (38 M -> 31 M
// instructions) ran no faster.  On B200 the FP64 pipe is only 2x narrower than the FP32 one, so this
// kernel drops the screen altogether and feeds the FP64 pipe directly:
//
//   * the CTA stages the whole stencil of its cell ONCE in shared memory: exact records with the image
//     shift already added in fp64, (q + S@cell) in the oracle's operation order;
//   * each warp holds 32 sites of that cell, one per lane, and visits every staged candidate with two
//     broadcast LDS.128 and ~15 FP64-pipe instructions: d = qs - p, fused r2, reciprocal, c6, and two
//     masked accumulators (sum c12, sum c6).  No divergence, no FIFO, four candidates in flight;
//   * membership r2 <= rc^2 must match the oracle bit for bit, energies only to 1e-10: the loop decides
//     on a fused r2 and re-derives r2 in the oracle's unfused arithmetic (rb_exact_r2) for the rare
//     candidate within a few ulps of the cutoff.  Sites outside the primary periodic cell are wrapped
//     first and get a wider re-check band, so they cost nothing extra.
//
// The staged-candidate list of a cell (z-plane, row, main piece, wrapped piece, cell-sorted order) is cut
// into w.nplanes equal slices; a work item walks w.item_slices of them one after the other behind one
// header.  A site's row sum is the fixed-order fold over the slices of sequentially accumulated sums -- a
// function of the site and the configuration only, independent of the lane, the warp, the batch the site
// arrives in and of how slices are grouped into work items.
//
// From the second batch against a resident configuration the stencil is not rebuilt from the cell list
// but read from the per-configuration StencilLists below (one ordered, coalesced run per cell).
//
// Boxes thinner than the stencil along x, or stencils with more than RB_MAXPIECES pieces, take the
// generic per-lane traversal (tiny systems).
#pragma once
#include "common.cuh"
#include "stencil.cuh"
#include "lj_kernels.cuh"
#include "site_kernels.cuh"

// Two warps per CTA, twelve CTAs per SM (measured against 4 x 6: 48.8 -> 44.8 us): a cell's sites rarely fill four
// warps, and a warp without sites holds registers through the whole loop of its CTA.
#ifndef RB_WARPS
#define RB_WARPS 2
#define RB_SITES_SHIFT 6
#define RB_CAND 256               // staged candidates per chunk
#define RB_MIN_BLOCKS 12
#endif
#define RB_THREADS (32 * RB_WARPS)
#define RB_SITES RB_THREADS       // sites per work item
#ifndef RB_UNROLL
#define RB_UNROLL 4
#endif
#define RB_MAXPIECES 128

template <int N> struct IntC { static constexpr int value = N; };

// The stencil of a cell as contiguous pieces of the cell-sorted array: rows (y, z offsets, z-plane major),
// each a main piece (image shift 0 along x) and the piece that wraps around the periodic x boundary.
struct RbPieces {
    int pbeg[RB_MAXPIECES];                     // first cell-sorted slot of the piece
    int pstart[RB_MAXPIECES + 1];               // offset of the piece in the stencil's candidate list
    int pS[RB_MAXPIECES];                       // image shift of the piece (pack_wraps)
    double pshift[RB_MAXPIECES][3];             // S @ cell
    int total;                                  // candidates in the stencil
};

struct RowsBlockSmem {
    double4 candd[RB_CAND];                     // q + S@cell (exact, unfused), .w = tag
    int2 cinfo[RB_CAND];                        // (cell-sorted slot, packed image shift of the piece): borderline re-check
    RbPieces pc;
    float wbox[RB_WARPS][6];                    // per-warp bounding boxes of the sites
    int wcnt[RB_WARPS];                         // survivors per warp of a staging round
};

// Lane l of the calling warp owns rows 2l, 2l+1 = pieces 4l .. 4l+3 (needs 2 * rows <= RB_MAXPIECES and
// a box at least as wide as the stencil along x).
__device__ __forceinline__ void rb_build_pieces(RbPieces& pc, const GridDesc& g, const int* __restrict__ cstart,
                                        int c0, int c1, int c2, int lane) {
    const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
    const int wr1 = 2 * m1 + 1, wr2 = 2 * m2 + 1, nrows = wr1 * wr2;
    const bool perx = g.pbc[0] != 0;

    int beg[4] = {0, 0, 0, 0}, len[4] = {0, 0, 0, 0}, Sx[4] = {0, 0, 0, 0}, Sy[2] = {0, 0}, Sz[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = 2 * lane + j;
        if (row < nrows) {
            int cy = c1 + (row % wr1) - m1, cz = c2 + (row / wr1) - m2;
            bool valid = true;
            if (g.pbc[1]) { Sy[j] = floor_div(cy, n1); cy -= Sy[j] * n1; }
            else if (cy < 0 || cy >= n1) valid = false;
            if (g.pbc[2]) { Sz[j] = floor_div(cz, n2); cz -= Sz[j] * n2; }
            else if (cz < 0 || cz >= n2) valid = false;
            if (valid) {
                const int rowoff = (cz * n1 + cy) * n0;
                const int lo = c0 - m0, hi = c0 + m0;
                const int a0 = max(lo, 0), a1 = min(hi, n0 - 1);
                beg[2 * j] = cstart[rowoff + a0];
                len[2 * j] = cstart[rowoff + a1 + 1] - beg[2 * j];
                if (perx) {
                    if (lo < 0) {
                        beg[2 * j + 1] = cstart[rowoff + lo + n0];
                        len[2 * j + 1] = cstart[rowoff + n0] - beg[2 * j + 1];
                        Sx[2 * j + 1] = -1;
                    } else if (hi >= n0) {
                        beg[2 * j + 1] = cstart[rowoff];
                        len[2 * j + 1] = cstart[rowoff + hi - n0 + 1] - beg[2 * j + 1];
                        Sx[2 * j + 1] = 1;
                    }
                }
            }
        }
    }
    const int mine = len[0] + len[1] + len[2] + len[3];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int base = incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ip = 4 * lane + j;
        if (ip < 2 * nrows) {
            pc.pstart[ip] = base;
            pc.pbeg[ip] = beg[j];
            pc.pS[ip] = pack_wraps(Sx[j], Sy[j >> 1], Sz[j >> 1]);
            double ax, ay, az;
            shift_vec(g, Sx[j], Sy[j >> 1], Sz[j >> 1], ax, ay, az);
            pc.pshift[ip][0] = ax; pc.pshift[ip][1] = ay; pc.pshift[ip][2] = az;
        }
        base += len[j];
    }
    if (lane == 31) { pc.pstart[2 * nrows] = incl; pc.total = incl; }
}

// r2 of staged candidate i against a site in the oracle's exact arithmetic: image shift
// T = S_piece + w_site - w_j rebuilt from the raw record, unfused (dx*dx + dy*dy) + dz*dz.
__device__ __noinline__ double rb_exact_r2(const GridDesc& g, const double4* __restrict__ spos,
                                           const RbPieces& pc, int packed, double px, double py, double pz,
                                           int wsite) {
    const double4 q = spos[packed & 0xFFFFFF];
    int sx, sy, sz, jx, jy, jz, w0, w1, w2;
    unpack_wraps(pc.pS[packed >> 24], sx, sy, sz);
    unpack_wraps(tag_wraps(tag_of(q)), jx, jy, jz);
    unpack_wraps(wsite, w0, w1, w2);
    double ax, ay, az, dx, dy, dz;
    shift_vec(g, sx + w0 - jx, sy + w1 - jy, sz + w2 - jz, ax, ay, az);
    return image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
}

// The site wrapped into the primary periodic cell and the width (ulps of rc^2) of the band in which the
// fused r2 of the loops below is re-checked in the oracle's arithmetic.  ONE out-of-line copy: every row-sum
// kernel (batched, pipelined, single-launch) and the pipelined kernel's bounding box see the same bits.
__device__ __noinline__ void rb_wrap_site(const GridDesc& g, double rc, double px, double py, double pz, int wsite,
                                          double& qx, double& qy, double& qz, double& band) {
    int w0, w1, w2;
    unpack_wraps(wsite, w0, w1, w2);
    const double a = (double)w0, bb = (double)w1, cc = (double)w2;
    qx = px - (a * g.cell[0] + bb * g.cell[3] + cc * g.cell[6]);
    qy = py - (a * g.cell[1] + bb * g.cell[4] + cc * g.cell[7]);
    qz = pz - (a * g.cell[2] + bb * g.cell[5] + cc * g.cell[8]);
    band = 16.0 + 64.0 * (fabs(px) + fabs(py) + fabs(pz) + fabs(qx) + fabs(qy) + fabs(qz)) / rc;
}

// Generic traversal for one site (thin boxes, oversized stencils): exact arithmetic directly.
// Returns the same accumulators as the fast path: sum c6^2, sum c6, pairs.
__device__ __noinline__ void rb_generic_row(const GridDesc& g, const double4* __restrict__ spos,
                                            const int* __restrict__ cstart, const LJParams& p,
                                            double px, double py, double pz,
                                            const int* __restrict__ old_idx, int ob, int oe,
                                            double& acc12, double& acc6, int& cnt) {
    double a12 = 0.0, a6 = 0.0;
    int c = 0;
    for_each_candidate(g, spos, cstart, px, py, pz, 0, 1, [&](const Cand& cd) {
        if (cd.r2 <= p.rc2) {
            const int j = tag_index(cd.tag);
            bool ex = false;
            for (int e = ob; e < oe; ++e) ex |= (old_idx[e] == j);
            if (!ex) {
                const double t = p.sigma2 * fast_rcp(cd.r2);
                const double c6 = t * t * t;
                a12 = fma(c6, c6, a12); a6 += c6; ++c;
            }
        }
    });
    acc12 = a12; acc6 = a6; cnt = c;
}

// Pruning rule of the stencil lists: a candidate image farther than rc from the BOX of the home cell is farther
// than rc from every site binned into that cell (sites of an edge cell that lie beyond the grid on a non-periodic
// axis are on the far side of the cell from every atom, so the bound holds for them too).  Per axis the distance
// to the cell's slab is (fractional distance) x (perpendicular height of the box); for a diagonal lattice the
// three combine to the exact box distance, otherwise only their maximum is a bound.  A function of the
// configuration alone: every kernel that walks a cell's stencil prunes it identically.
__device__ __forceinline__ bool rb_keep_candidate(const GridDesc& g, int c0, int c1, int c2,
                                                  double x, double y, double z, double reach) {
    double f[3];
    to_frac(g, x, y, z, f);
    const int c[3] = {c0, c1, c2};
    double e2 = 0.0, emax = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double flo = g.lo[d] + (double)c[d] / g.scale[d], fhi = g.lo[d] + (double)(c[d] + 1) / g.scale[d];
        const double e = fmax(fmax(flo - f[d], f[d] - fhi), 0.0) * g.height[d];
        e2 = fma(e, e, e2);
        emax = fmax(emax, e);
    }
    return g.ortho ? (e2 <= reach * reach) : (emax <= reach);
}
__device__ __forceinline__ double rb_keep_reach(double rc) { return rc * 1.000001 + 1.0e-9; }

// The stencil of every cell, written out once per resident configuration (expand_stencils_kernel): the
// candidate records the row-sum kernel would otherwise re-derive for every work item of every launch
// (piece table, a binary search and a gather per candidate, image shift) as one contiguous, ordered run
// per cell.  40 bytes per record, (2m+1)^3 cells' worth of atoms per cell: 10 MB for the 2 000-atom LJ box,
// L2-resident.  A cell whose list does not fit the buffer (or whose box needs the generic traversal) has
// head.x < 0 and is staged on the fly as before.
struct StencilLists {
    double4* q;                                // q + S@cell (exact, unfused), .w low word = atom index
    int2* info;                                // (cell-sorted slot, packed image shift of the piece)
    int2* head;                                // [R * bin_stride] (first record, records) of the cell; nullptr: no lists
    unsigned long long* cursor;                // allocation cursor, zeroed before the expansion
    long long cap;                             // records the buffers hold
};

// Candidate gi of the stencil described by the piece table: the staged record and its re-check info.
__device__ __forceinline__ void rp_stage_candidate(const GridDesc& g, const double4* __restrict__ spos,
                                                   const RbPieces& pc, int npieces, int gi, double4& q, int2& info) {
    int lo = 0, hi = npieces;                  // pstart[lo] <= gi < pstart[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pc.pstart[mid] <= gi) lo = mid; else hi = mid;
    }
    const int k = pc.pbeg[lo] + (gi - pc.pstart[lo]);
    q = spos[k];
    double ax = pc.pshift[lo][0], ay = pc.pshift[lo][1], az = pc.pshift[lo][2];
    const int pS = pc.pS[lo];
    const long long tag = tag_of(q);
    const int wj = tag_wraps(tag);
    if (wj != pack_wraps(0, 0, 0)) {           // atom stored outside the primary cell
        int sx, sy, sz, jx, jy, jz;
        unpack_wraps(pS, sx, sy, sz);
        unpack_wraps(wj, jx, jy, jz);
        shift_vec(g, sx - jx, sy - jy, sz - jz, ax, ay, az);
    }
    q.x = __dadd_rn(q.x, ax); q.y = __dadd_rn(q.y, ay); q.z = __dadd_rn(q.z, az);
    // low word = atom index | species slot << 24 (the slot is 0 for LJ, so the LJ loops read the index as it is)
    q.w = __longlong_as_double((long long)tag_index(tag) | ((long long)tag_species(tag) << 24));
    info = make_int2(k, pS);
}

#define EX_WARPS 4
// One warp per (replica, cell): piece table, then the whole stencil in list order.
__global__ void __launch_bounds__(32 * EX_WARPS)
expand_stencils_kernel(BatchView b, StencilLists x, int nbins, int bin_stride, double rc) {
    __shared__ RbPieces pcs[EX_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RbPieces& pc = pcs[wid];
    for (int bin = blockIdx.x * EX_WARPS + wid; bin < nbins; bin += gridDim.x * EX_WARPS) {
        const int r = bin / bin_stride, c = bin - r * bin_stride;
        const GridDesc& g = b.grid[r];
        int2 head = make_int2(-1, 0);
        const int n0 = g.ncell[0], n1 = g.ncell[1];
        const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
        const int nrows = (2 * m1 + 1) * (2 * m2 + 1);
        const bool fast = (!g.pbc[0] || n0 >= 2 * m0 + 1) && 2 * nrows <= RB_MAXPIECES;
        if (c < g.ncells && fast) {
            const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
            const double4* spos = b.spos + g.atom_off;
            const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
            __syncwarp();                           // the previous bin's table is no longer read
            rb_build_pieces(pc, g, cstart, c0, c1, c2, lane);
            __syncwarp();
            const int total = pc.total, npieces = 2 * nrows;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(x.cursor, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned long long)total <= (unsigned long long)x.cap) {
                // the run is reserved for the whole stencil and filled, in list order, with the candidates that
                // can be within rc of a site of this cell (the unused tail is never read)
                const double reach = rb_keep_reach(rc);
                int kept = 0;
                for (int g0 = 0; g0 < total; g0 += 32) {
                    const int gi = g0 + lane;
                    double4 q = make_double4(0.0, 0.0, 0.0, 0.0);
                    int2 info = make_int2(0, 0);
                    bool keep = false;
                    if (gi < total) {
                        rp_stage_candidate(g, spos, pc, npieces, gi, q, info);
                        keep = rb_keep_candidate(g, c0, c1, c2, q.x, q.y, q.z, reach);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const unsigned long long at = base + kept + __popc(bal & ((1u << lane) - 1u));
                        MCPYB_CHECK(at < (unsigned long long)x.cap);
                        x.q[at] = q;
                        x.info[at] = info;
                    }
                    kept += __popc(bal);
                }
                head = make_int2((int)base, kept);
            }
        }
        if (lane == 0) x.head[bin] = head;
    }
}

// rb_exact_r2 with the piece's image shift taken from the staged record instead of the piece table
__device__ __noinline__ double rp_exact_r2(const GridDesc& g, const double4* __restrict__ spos, int2 info,
                                           double px, double py, double pz, int wsite) {
    const double4 q = spos[info.x];
    int sx, sy, sz, jx, jy, jz, w0, w1, w2;
    unpack_wraps(info.y, sx, sy, sz);
    unpack_wraps(tag_wraps(tag_of(q)), jx, jy, jz);
    unpack_wraps(wsite, w0, w1, w2);
    double ax, ay, az, dx, dy, dz;
    shift_vec(g, sx + w0 - jx, sy + w1 - jy, sz + w2 - jz, ax, ay, az);
    return image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, dx, dy, dz);
}

template <int RBU, int MINB>
__global__ void __launch_bounds__(RB_THREADS, MINB)
lj_trial_rows_block_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p, StencilLists x) {
    __shared__ RowsBlockSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_wait();
    pdl_trigger();
    const int nq = w.nplanes;                     // slices of the candidate list, one work item each
    // a work item covers item_slices consecutive slices of its cell's list (one header -- sites, bounding box, list
    // head -- for all of them); every slice is still accumulated and stored on its own, so the bits of a row sum do
    // not depend on how the slices are grouped into items
    const int isl = w.item_slices, ngr = nq / isl;
    const int nwork = w.totals[0] * ngr;
    // persistent CTAs: the first work item is the CTA's own index, the following ones come from a device
    // counter (zeroed by the scan kernel), claimed one item AHEAD so that the atomic's round trip hides
    // behind the current item.  The grid is sized by the SM count, not by the host's upper bound on the
    // item count, so no CTA is launched only to find nothing to do, and the tail is one item long.
    __shared__ int s_wi;
    int next_wi = blockIdx.x;
    for (;;) {
        if (tid == 0) s_wi = next_wi;
        __syncthreads();
        const int wi = s_wi;
        if (wi >= nwork) break;
        if (tid == 0) next_wi = (int)gridDim.x + atomicAdd(&w.totals[1], 1);
        const int item = wi / ngr, slice0 = (wi - item * ngr) * isl;
        const int2 it = w.item_tab[item];
        const int bin = it.x, group = it.y;
        const int r = bin / w.bin_stride, c = bin - r * w.bin_stride;
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        const double4* spos = b.spos + off;
        const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
        const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
        const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
        const int wr1 = 2 * m1 + 1, wr2 = 2 * m2 + 1, nrows = wr1 * wr2;
        const bool perx = g.pbc[0] != 0;
        const bool fast = (!perx || n0 >= 2 * m0 + 1) && 2 * nrows <= RB_MAXPIECES;

        // ---- this thread's site
        const int kbeg = w.bin_count[bin] + group * RB_SITES, kend = w.bin_count[bin + 1];
        const bool active = kbeg + tid < kend;
        const bool warp_active = kbeg + wid * 32 < kend;
        int s = 0, ob = 0, oe = 0, ex0 = -1, wsite = pack_wraps(0, 0, 0);
        double px = 0.0, py = 0.0, pz = 0.0;
        if (active) {
            const double4 me = w.sorted_pos[kbeg + tid];
            const int4 mt = w.sorted_meta[kbeg + tid];
            px = me.x; py = me.y; pz = me.z;
            wsite = (int)__double_as_longlong(me.w);
            s = mt.x; ex0 = mt.y; ob = mt.z; oe = mt.w;
        }
        double acc12 = 0.0, acc6 = 0.0;
        int cnt = 0;
        auto store_row = [&](int slice) {
            if (active) {
                // row = 4 eps (sum c12 - sum c6) - e0 * pairs
                const double row = fma(p.eps4, acc12 - acc6, -p.e0 * (double)cnt);
                w.site_E[(size_t)s * nq + slice] = row;
                w.site_cnt[(size_t)s * nq + slice] = cnt;
            }
        };

        if (!fast) {
            for (int slice = slice0; slice < slice0 + isl; ++slice) {
                acc12 = 0.0; acc6 = 0.0; cnt = 0;
                if (active && slice == 0)
                    rb_generic_row(g, spos, cstart, p, px, py, pz, tb.old_idx, ob, oe, acc12, acc6, cnt);
                store_row(slice);
            }
        } else {
            // ---- the cell's stencil: its expanded list if there is one, else the piece table (warp 0)
            const int2 head = x.head ? x.head[bin] : make_int2(-1, 0);
            const bool listed = head.x >= 0;
            if (!listed) {
                if (wid == 0) rb_build_pieces(sm.pc, g, cstart, c0, c1, c2, lane);
                __syncthreads();
            }
            const int total = listed ? head.y : sm.pc.total, npieces = 2 * nrows;
            const double4* xq = x.q + (listed ? head.x : 0);
            const int2* xi = x.info + (listed ? head.x : 0);

            // ---- per-lane constants.  The loop works with the site wrapped into the primary cell
            // (qx, qy, qz) and a fused r2; both differ from the oracle's value by rounding only.
            // `in` is decided on the bit pattern of r2 (positive doubles order like integers):
            // below lo_bits -> inside, above hi_bits -> outside, in between -> exact re-check.
            double qx = 3.0e150, qy = 3.0e150, qz = 3.0e150;     // idle lanes: r2 overflows, never inside
            double band = 16.0;                                    // ulps, relative to rc^2
            if (active) {
                qx = px; qy = py; qz = pz;
                if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(g, p.rc, px, py, pz, wsite, qx, qy, qz, band);
            }
            const double bw = p.rc2 * band * 2.220446049250313e-16;
            const long long lo_bits = __double_as_longlong(p.rc2 - bw), hi_bits = __double_as_longlong(p.rc2 + bw);
            // moves of up to two atoms (atomic moves, diatomic molecules) exclude through two registers;
            // larger rigid molecules walk their exclusion list on the slow path
            const bool multi = (oe - ob) > 2;
            const int ex1 = (oe - ob) > 1 ? tb.old_idx[ob + 1] : -1;

            // ---- bounding box of this CTA's (wrapped) sites, rounded outward to fp32.  A candidate
            // farther than rc from the box is farther than rc from every site: it is dropped at staging
            // (its masked contribution would be +0.0 to every accumulator, so no sum changes).
            {
                float lo3[3] = {__double2float_rd(qx), __double2float_rd(qy), __double2float_rd(qz)};
                float hi3[3] = {__double2float_ru(qx), __double2float_ru(qy), __double2float_ru(qz)};
                if (!active) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { lo3[d] = 3.0e38f; hi3[d] = -3.0e38f; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        lo3[d] = fminf(lo3[d], __shfl_xor_sync(0xffffffffu, lo3[d], o));
                        hi3[d] = fmaxf(hi3[d], __shfl_xor_sync(0xffffffffu, hi3[d], o));
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { sm.wbox[wid][d] = lo3[d]; sm.wbox[wid][3 + d] = hi3[d]; }
                }
            }
            __syncthreads();
            // The pruning test runs in fp32: candidates rounded to nearest, the box rounded outward, and the
            // radius carries a margin for the rounding of coordinates as large as any that can pass.
            float blo[3], bhi[3], maxabs = 0.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float l = sm.wbox[0][d], h = sm.wbox[0][3 + d];
#pragma unroll
                for (int ww = 1; ww < RB_WARPS; ++ww) { l = fminf(l, sm.wbox[ww][d]); h = fmaxf(h, sm.wbox[ww][3 + d]); }
                blo[d] = l; bhi[d] = h;
                maxabs = fmaxf(maxabs, fmaxf(fabsf(l), fabsf(h)));
            }
            const float rcf = __double2float_ru(p.rc);
            const float keep_r = rcf * 1.000002f + 1.0e-6f * (maxabs + rcf);
            const float keep_r2 = keep_r * keep_r * 1.000001f;
            const int lo_hi = (int)(lo_bits >> 32), hi_hi = (int)(hi_bits >> 32);

            for (int slice = slice0; slice < slice0 + isl; ++slice) {
                acc12 = 0.0; acc6 = 0.0; cnt = 0;
                // slice of the stencil's candidate list: a function of the cell alone
                const int cfirst = (int)(((long long)total * slice) / nq), clast = (int)(((long long)total * (slice + 1)) / nq);
                int gi0 = cfirst;
                while (gi0 < clast) {
                    // ---- stage: test RB_THREADS candidates per round, keep the survivors in list order
                    int nst = 0;
                    while (gi0 < clast && nst + RB_THREADS <= RB_CAND) {
                        const int gi = gi0 + tid;
                        bool keep = false;
                        double4 q = make_double4(0.0, 0.0, 0.0, 0.0);
                        int2 info = make_int2(0, 0);
                        if (gi < clast) {
                            if (listed) { q = xq[gi]; info = xi[gi]; }
                            else rp_stage_candidate(g, spos, sm.pc, npieces, gi, q, info);
                            const float xf = __double2float_rn(q.x), yf = __double2float_rn(q.y), zf = __double2float_rn(q.z);
                            const float ex = fmaxf(fmaxf(blo[0] - xf, xf - bhi[0]), 0.0f);
                            const float ey = fmaxf(fmaxf(blo[1] - yf, yf - bhi[1]), 0.0f);
                            const float ez = fmaxf(fmaxf(blo[2] - zf, zf - bhi[2]), 0.0f);
                            keep = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= keep_r2;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, keep);
                        if (lane == 0) sm.wcnt[wid] = __popc(bal);
                        __syncthreads();
                        int before = 0, all = 0;
    #pragma unroll
                        for (int ww = 0; ww < RB_WARPS; ++ww) {
                            const int cw = sm.wcnt[ww];
                            before += (ww < wid) ? cw : 0;
                            all += cw;
                        }
                        if (keep) {
                            const int slot = nst + before + __popc(bal & ((1u << lane) - 1u));
                            MCPYB_CHECK(slot >= 0 && slot < RB_CAND);
                            sm.candd[slot] = q;
                            sm.cinfo[slot] = info;
                        }
                        nst += all;
                        gi0 += RB_THREADS;
                        __syncthreads();                               // wcnt is rewritten next round
                    }
                    const int n = nst;
                    if (warp_active) {
                        // U candidates per trip, branch-free, so the U dependent FP64 chains interleave.  The
                        // trip is handed to `rare` when some lane has a candidate whose high word of r2 falls
                        // in the window [hi(lo_bits), hi(hi_bits)] (full compare, exact re-check) or belongs
                        // to a multi-atom move (exclusion list).
                        const unsigned amb_width = (unsigned)(hi_hi - lo_hi);
                        auto rare = [&](int i, int U) {
                            for (int u = 0; u < U; ++u) {
                                const double4 q = sm.candd[i + u];
                                const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                                const long long rb = __double_as_longlong(r2);
                                bool in = rb < lo_bits;
                                if (!in && rb <= hi_bits)        // within rounding of the cutoff: the oracle's arithmetic decides
                                    in = rp_exact_r2(g, spos, sm.cinfo[i + u], px, py, pz, wsite) <= p.rc2;
                                const int j = __double2loint(q.w);
                                in = in && (j != ex0) && (j != ex1);
                                if (multi && in)
                                    for (int e = ob + 2; e < oe; ++e) if (tb.old_idx[e] == j) in = false;
                                const double rm = __hiloint2double(in ? __double2hiint(r2) : 0x7fe00000, __double2loint(r2));
                                const double t = p.sigma2 * fast_rcp(rm);
                                const double c6 = t * t * t;
                                acc12 = fma(c6, c6, acc12);
                                acc6 += c6;
                                cnt += in ? 1 : 0;
                            }
                        };
                        auto visit = [&](int i, auto U_) {
                            constexpr int U = decltype(U_)::value;
                            double r2[U];
                            int jj[U], rh[U];
                            unsigned amb = multi ? 0u : 0xffffffffu;   // min over the trip of (hi(r2) - lo_hi), unsigned
    #pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const double4 q = sm.candd[i + u];    // uniform address: two broadcast LDS.128
                                const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                                r2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                                rh[u] = __double2hiint(r2[u]);        // positive doubles order like their bits
                                amb = min(amb, (unsigned)(rh[u] - lo_hi));
                                jj[u] = __double2loint(q.w);
                            }
                            if (__any_sync(0xffffffffu, amb <= amb_width)) { rare(i, U); return; }
    #pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const bool hit = rh[u] < lo_hi && jj[u] != ex0 && jj[u] != ex1;
                                // a miss is evaluated at r2 ~ 9e307: its reciprocal flushes to zero, so c6 = 0
                                const double rm = __hiloint2double(hit ? rh[u] : 0x7fe00000, __double2loint(r2[u]));
                                const double t = p.sigma2 * fast_rcp(rm);
                                const double c6 = t * t * t;
                                acc12 = fma(c6, c6, acc12);
                                acc6 += c6;
                                if (hit) ++cnt;
                            }
                        };
                        int i = 0;
                        for (; i + RBU <= n; i += RBU) visit(i, IntC<RBU>());
                        for (; i < n; ++i) visit(i, IntC<1>());
                    }
                    __syncthreads();
                }
                store_row(slice);
            }
        }
        __syncthreads();                                          // the piece table and s_wi are rewritten next
    }
}

// ---------------------------------------------------------------------------
// Small batches (the one-call-per-trial pattern of GrandCanonicalEnsemble.do_gcmc_step, a handful of
// sites): ONE launch.  One CTA per site spreads the stencil's candidates over its threads, evaluates
// them with the arithmetic of the kernel above (wrapped site, fused r2, bit-pattern membership with
// the exact re-check, same reciprocal and c6), leaves c6 of every candidate (0 for a miss) in shared
// memory in list order, and one thread per slice then runs the same sequential accumulation
// acc12 = fma(c6, c6, acc12), acc6 += c6 -- so a site's partial sums carry the same bits as in a large
// batch.  The last CTA to finish (threadfence + counter) combines the rows into dE for every trial and
// publishes the validation status: no histogram, scan, scatter or separate combine launch.
// ---------------------------------------------------------------------------
#define RS_THREADS 128
#define RS_MAX_SITES 1024          // batches up to this many sites take the single-launch kernel
#define RS_CAND 1024              // candidates of one stencil held in shared memory per chunk

struct RowsSmallSmem {
    RbPieces pc;
    double c6[RS_CAND];
    unsigned char hit[RS_CAND];
    double part12[8], part6[8];
    int partc[8];
    int last;
};

__device__ __forceinline__ void lj_trial_small_body(RowsSmallSmem& sm, const BatchView& b, const TrialBatch& tb,
                                                    const TrialWork& w, const LJParams& p,
                                                    double* __restrict__ dE, int* __restrict__ npairs,
                                                    int* __restrict__ done_counter, const StencilLists& x,
                                                    volatile unsigned int* done_flag = nullptr, unsigned int seq = 0,
                                                    int site = -1, int nctas = -1) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int s = site >= 0 ? site : (int)blockIdx.x;  // site: [0, n_old) old, then new
    const int ncta = nctas >= 0 ? nctas : (int)gridDim.x;
    const int nq = w.nplanes;
    // ---- owning trial (T is small: binary search over the offsets)
    const bool is_old = s < w.n_old;
    const int t = is_old ? owner_of(tb.old_off, tb.T, s) : owner_of(tb.new_off, tb.T, s - w.n_old);
    const int r = tb.rep ? tb.rep[t] : 0;
    const int ob = tb.old_off[t], oe = tb.old_off[t + 1];
    const int nb = tb.new_off[t], ne = tb.new_off[t + 1];
    bool ok = !trial_invalid(b, tb, w, r, ob, oe, nb, ne);
    ok = ok && (is_old ? (s >= ob && s < oe) : (s - w.n_old >= nb && s - w.n_old < ne));
    // the row sum is the fold ((a0 + a1) + (a2 + a3)) of four class sums, class k = the candidates at positions
    // = k (mod 4) of the cell's PRUNED stencil list, each accumulated in list order (rows_warp.cuh): thread k < 4
    // carries class k across the chunks
    double a12 = 0.0, a6 = 0.0;
    int cc = 0, kpos = 0;                              // kpos: pruned-list position of the next kept candidate
    bool generic = false;
    if (ok) {
        const GridDesc& g = b.grid[r];
        const int off = g.atom_off;
        const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
        const double4* spos = b.spos + off;
        double px, py, pz;
        if (is_old) { const double4 q = b.upos[off + tb.old_idx[s]]; px = q.x; py = q.y; pz = q.z; }
        else { const int k = s - w.n_old; px = tb.new_pos[3 * k]; py = tb.new_pos[3 * k + 1]; pz = tb.new_pos[3 * k + 2]; }
        int c[3], wr[3];
        locate(g, px, py, pz, c, wr);
        const int wsite = pack_wraps(wr[0], wr[1], wr[2]);
        const int n0 = g.ncell[0], m0 = g.m[0];
        const int nrows = (2 * g.m[1] + 1) * (2 * g.m[2] + 1);
        const bool fast = (!g.pbc[0] || n0 >= 2 * m0 + 1) && 2 * nrows <= RB_MAXPIECES;
        // the cell's stencil: its per-configuration list when there is one (one coalesced read, no piece
        // table, no search), else built here from the cell list and pruned by the same rule
        const int cell_bin = r * w.bin_stride + (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
        const int2 head = (fast && x.head) ? x.head[cell_bin] : make_int2(-1, 0);
        const bool listed = head.x >= 0;
        if (!fast || (x.head && !listed)) {
            // thin box / oversized stencil / list buffer full: the batched kernel takes the exact per-lane
            // traversal for such a cell, and so does this one
            generic = true;
            if (tid == 0) rb_generic_row(g, spos, cstart, p, px, py, pz, tb.old_idx, ob, oe, a12, a6, cc);
        } else {
            if (!listed) {
                if (wid == 0) rb_build_pieces(sm.pc, g, cstart, c[0], c[1], c[2], lane);
                __syncthreads();
            }
            const int total = listed ? head.y : sm.pc.total, npieces = 2 * nrows;
            const double4* xq = x.q + (listed ? head.x : 0);
            const int2* xi = x.info + (listed ? head.x : 0);
            const double reach = rb_keep_reach(p.rc);
            // per-site constants, as in lj_trial_rows_warp_kernel
            double qx = px, qy = py, qz = pz, band = 16.0;
            if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(g, p.rc, px, py, pz, wsite, qx, qy, qz, band);
            const double bw = p.rc2 * band * 2.220446049250313e-16;
            const long long lo_bits = __double_as_longlong(p.rc2 - bw), hi_bits = __double_as_longlong(p.rc2 + bw);
            for (int cbase = 0; cbase < total; cbase += RS_CAND) {
                const int n = min(RS_CAND, total - cbase);
                // c6 of the chunk's KEPT candidates, compacted in list order (a listed cell holds kept ones only)
                int kept = 0;
                for (int i0 = 0; i0 < n; i0 += RS_THREADS) {
                    const int i = i0 + tid;
                    bool keep = false, in = false;
                    double c6 = 0.0;
                    if (i < n) {
                        const int gi = cbase + i;
                        double4 q;
                        int2 info = make_int2(0, 0);
                        keep = true;
                        if (listed) q = xq[gi];
                        else {
                            rp_stage_candidate(g, spos, sm.pc, npieces, gi, q, info);
                            keep = rb_keep_candidate(g, c[0], c[1], c[2], q.x, q.y, q.z, reach);
                        }
                        const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        const long long rb = __double_as_longlong(r2);
                        in = keep && rb < lo_bits;
                        if (keep && !in && rb <= hi_bits) in = rp_exact_r2(g, spos, listed ? xi[gi] : info, px, py, pz, wsite) <= p.rc2;
                        const int j = __double2loint(q.w) & 0xFFFFFF;
                        for (int e = ob; e < oe; ++e) if (tb.old_idx[e] == j) in = false;
                        if (in) { const double tt = p.sigma2 * fast_rcp(r2); c6 = tt * tt * tt; }
                    }
                    if (listed) {
                        if (i < n) { sm.c6[i] = c6; sm.hit[i] = in ? 1 : 0; }
                    } else {
                        const unsigned bal = __ballot_sync(0xffffffffu, keep);
                        if (lane == 0) sm.partc[wid] = __popc(bal);
                        __syncthreads();
                        int before = 0, all = 0;
                        for (int k = 0; k < RS_THREADS / 32; ++k) { const int v = sm.partc[k]; all += v; if (k < wid) before += v; }
                        if (keep) {
                            const int at = kept + before + __popc(bal & ((1u << lane) - 1u));
                            sm.c6[at] = c6;
                            sm.hit[at] = in ? 1 : 0;
                        }
                        kept += all;
                        __syncthreads();
                    }
                }
                if (listed) kept = n;
                __syncthreads();
                // class k = the kept candidates at list positions = k (mod 4), each summed in list order
                if (tid < 4) {
                    for (int i = (tid - kpos) & 3; i < kept; i += 4) {
                        const double c6 = sm.c6[i];
                        a12 = fma(c6, c6, a12);
                        a6 += c6;
                        cc += sm.hit[i];
                    }
                }
                kpos += kept;
                __syncthreads();
            }
        }
        // fold the classes in the fixed order ((a0 + a1) + (a2 + a3))
        if (!generic) {
            if (tid < 4) { sm.part12[tid] = a12; sm.part6[tid] = a6; sm.partc[tid] = cc; }
            __syncthreads();
            if (tid == 0) {
                a12 = (sm.part12[0] + sm.part12[1]) + (sm.part12[2] + sm.part12[3]);
                a6 = (sm.part6[0] + sm.part6[1]) + (sm.part6[2] + sm.part6[3]);
                cc = (sm.partc[0] + sm.partc[1]) + (sm.partc[2] + sm.partc[3]);
            }
        }
        if (tid == 0) {
            w.site_E[s] = fma(p.eps4, a12 - a6, -p.e0 * (double)cc);
            w.site_cnt[s] = cc;
        }
    }
    // ---- the last CTA to arrive combines every trial
    __threadfence();
    __syncthreads();
    if (tid == 0) sm.last = (atomicAdd(done_counter, 1) == ncta - 1) ? 1 : 0;
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    unsigned long long first_bad = ~0ULL;
    for (int tt = tid; tt < tb.T; tt += RS_THREADS) {
        const int rr = tb.rep ? tb.rep[tt] : 0;
        const int bad = trial_invalid(b, tb, w, rr, tb.old_off[tt], tb.old_off[tt + 1], tb.new_off[tt], tb.new_off[tt + 1]);
        if (bad) first_bad = min(first_bad, ((unsigned long long)tt << 8) | (unsigned long long)bad);
        combine_trial(b, tb, w, p, tt, dE, npairs);
    }
    if (first_bad != ~0ULL) atomicMin(w.err, first_bad);
    if (done_flag) __threadfence_system();        // dE (pinned host memory) before the sequence word below
    __syncthreads();
    if (tid == 0) {
        const unsigned long long ev = *w.err;
        if (w.err_out) *w.err_out = ev;
        *w.err = ~0ULL;
        *done_counter = 0;
        if (done_flag) {
            // the host polls this word instead of waiting for the kernel to retire (engine.cu, wait_flag)
            __threadfence_system();
            *done_flag = seq;
            __threadfence_system();
        }
    }
}

__global__ void __launch_bounds__(RS_THREADS)
lj_trial_small_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p,
                      double* __restrict__ dE, int* __restrict__ npairs, int* __restrict__ done_counter, StencilLists x,
                      volatile unsigned int* done_flag, unsigned int seq) {
    // done_flag != nullptr: dE / npairs / the status word are pinned host memory and the host polls done_flag
    __shared__ RowsSmallSmem sm;
    lj_trial_small_body(sm, b, tb, w, p, dE, npairs, done_counter, x, done_flag, seq);
}

// One trial whose description travels in the kernel's parameter space and whose result is written
// straight into pinned host memory: no H2D or D2H copy on the latency-bound single-chain path
// (mcpyb_tracked_energy_host).
#define TINY_MAX_MOVED 8
struct TinyTrial {
    int old_off[2], new_off[2];
    int old_idx[TINY_MAX_MOVED];
    double new_pos[3 * TINY_MAX_MOVED];
};

__global__ void __launch_bounds__(RS_THREADS)
lj_trial_tiny_kernel(BatchView b, const __grid_constant__ TinyTrial tt, TrialWork w, LJParams p,
                     double* __restrict__ dE_host, int* __restrict__ done_counter, StencilLists x,
                     volatile unsigned int* done_flag, unsigned int seq) {
    __shared__ RowsSmallSmem sm;
    TrialBatch tb;
    tb.T = 1; tb.rep = nullptr; tb.old_off = tt.old_off; tb.old_idx = tt.old_idx;
    tb.new_off = tt.new_off; tb.new_pos = tt.new_pos; tb.new_Z = nullptr;
    lj_trial_small_body(sm, b, tb, w, p, dE_host, nullptr, done_counter, x, done_flag, seq);
}
