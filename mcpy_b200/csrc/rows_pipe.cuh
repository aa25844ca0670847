// K3 / K6, fourth generation: the CTA-per-cell row sums of rows_block.cuh as a producer / consumer pipeline.
//
// profiles/r1_v12_*: the third-generation kernel spends a third of its warp-stall samples in the loop and
// the rest before and after it.  One work item is ~3 us of FP64-pipe time behind ~6 us of dependent
// latency (work counter -> item table -> bin bounds -> site records -> piece table -> 2-3 staging rounds,
// each a binary search + a gather + a block-wide compaction), every CTA runs its items one after the
// other, and six CTAs per SM are not enough to hide that chain: FP64 pipe 35 %, kernel 50 us for 14 us
// of pipe time.
//
// Here a CTA is four CONSUMER warps (lane = site, the loop of rows_block.cuh unchanged: same arithmetic,
// same order, same bits) and one PRODUCER warp that claims the work items and stages them -- site
// records, bounding box, piece table, pruned candidate list -- into a two-stage ring in shared memory
// while the consumers are still looping over the previous stage.  Hand-over is by named barriers
// (bar.arrive / bar.sync with 160 participants): full[s] producer -> consumers, empty[s] back.  A stencil
// that does not fit one stage is handed over in chunks; the consumers carry their accumulators across the
// chunks of an item, so a site's row sum is still the sequential fold over its slice of the list.
#pragma once
#include "rows_block.cuh"

#define RP_CWARPS RB_WARPS                     // consumer warps = 128 sites per item, as the scan kernel cuts them
#define RP_THREADS (32 * (RP_CWARPS + 1))
#define RP_CAND 384                            // staged candidates per stage
#define RP_MIN_BLOCKS 4

#define RP_FIRST 1                             // first chunk of a work item: load the site, reset the sums
#define RP_LAST 2                              // last chunk: write the row
#define RP_EXIT 4
#define RP_GENERIC 8                           // thin box / oversized stencil: per-lane generic traversal

struct RpStage {
    double4 candd[RP_CAND];                    // q + S@cell (exact, unfused), .w low word = atom index
    int2 cinfo[RP_CAND];                       // (cell-sorted slot, packed image shift of the piece): exact re-check
    double4 spos[RB_SITES];                    // site records of the item (x, y, z, packed wraps)
    int4 smeta[RB_SITES];                      // (site id, first excluded atom or -1, old_off[t], old_off[t+1])
    int n;                                     // staged candidates
    int flags;
    int nsites;                                // sites of the item (<= RB_SITES)
    int slice;
    int rep;                                   // replica
    int pad[3];
};

struct RowsPipeSmem {
    RpStage st[2];
    RbPieces pc;                               // the producer's piece table
};

__device__ __forceinline__ void rp_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(RP_THREADS) : "memory"); }
__device__ __forceinline__ void rp_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(RP_THREADS) : "memory"); }

template <int RBU, int MINB>
__global__ void __launch_bounds__(RP_THREADS, MINB)
lj_trial_rows_pipe_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p, StencilLists x) {
    extern __shared__ __align__(16) unsigned char rp_raw[];
    RowsPipeSmem& sm = *reinterpret_cast<RowsPipeSmem*>(rp_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_wait();                                   // the scatter kernel's output (programmatic dependent launch)
    pdl_trigger();
    const int nq = w.nplanes;
    const int nwork = w.totals[0] * nq;

    if (wid == RP_CWARPS) {
        // ================================================================ producer warp
        int nfill = 0;
        auto acquire = [&]() -> RpStage& {
            const int s = nfill & 1;
            if (nfill >= 2) rp_bar_sync(3 + s);   // the consumers are done with what this stage held
            return sm.st[s];
        };
        auto publish = [&]() {
            __threadfence_block();
            __syncwarp();
            rp_bar_arrive(1 + (nfill & 1));
            ++nfill;
        };
        int wi_next = 0;
        if (lane == 0) wi_next = atomicAdd(&w.totals[1], 1);
        for (;;) {
            const int wi = __shfl_sync(0xffffffffu, wi_next, 0);
            if (wi >= nwork) break;
            if (lane == 0) wi_next = atomicAdd(&w.totals[1], 1);      // next claim: its latency hides behind this item
            const int item = wi / nq, slice = wi - item * nq;
            const int2 it = w.item_tab[item];
            const int bin = it.x, group = it.y;
            const int r = bin / w.bin_stride, c = bin - r * w.bin_stride;
            const GridDesc& g = b.grid[r];
            const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
            const double4* spos = b.spos + g.atom_off;
            const int n0 = g.ncell[0], n1 = g.ncell[1];
            const int m0 = g.m[0], m1 = g.m[1], m2 = g.m[2];
            const int c0 = c % n0, c1 = (c / n0) % n1, c2 = c / (n0 * n1);
            const int nrows = (2 * m1 + 1) * (2 * m2 + 1);
            const bool fast = (!g.pbc[0] || n0 >= 2 * m0 + 1) && 2 * nrows <= RB_MAXPIECES;
            const int kbeg = w.bin_count[bin] + group * RB_SITES;
            const int nsites = min(w.bin_count[bin + 1] - kbeg, RB_SITES);

            RpStage* S = &acquire();
            // ---- site records and the bounding box of the (wrapped) sites, rounded outward to fp32
            float lo3[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi3[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
            for (int j = 0; j < RB_SITES / 32; ++j) {
                const int k = lane + 32 * j;
                if (k < nsites) {
                    const double4 me = w.sorted_pos[kbeg + k];
                    S->spos[k] = me;
                    S->smeta[k] = w.sorted_meta[kbeg + k];
                    double qx = me.x, qy = me.y, qz = me.z, band;
                    const int wsite = (int)__double_as_longlong(me.w);
                    if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(g, p.rc, me.x, me.y, me.z, wsite, qx, qy, qz, band);
                    lo3[0] = fminf(lo3[0], __double2float_rd(qx)); hi3[0] = fmaxf(hi3[0], __double2float_ru(qx));
                    lo3[1] = fminf(lo3[1], __double2float_rd(qy)); hi3[1] = fmaxf(hi3[1], __double2float_ru(qy));
                    lo3[2] = fminf(lo3[2], __double2float_rd(qz)); hi3[2] = fmaxf(hi3[2], __double2float_ru(qz));
                }
            }
            if (lane == 0) { S->nsites = nsites; S->slice = slice; S->rep = r; }
            if (!fast) {
                if (lane == 0) { S->n = 0; S->flags = RP_FIRST | RP_LAST | RP_GENERIC; }
                publish();
                continue;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    lo3[d] = fminf(lo3[d], __shfl_xor_sync(0xffffffffu, lo3[d], o));
                    hi3[d] = fmaxf(hi3[d], __shfl_xor_sync(0xffffffffu, hi3[d], o));
                }
            }
            // The pruning test runs in fp32 (a lone warp is bound by its instruction count, and an FP64 max is
            // half a dozen instructions): candidates are rounded to nearest, the box was rounded outward, and
            // the radius carries a margin for the rounding of coordinates as large as any that can pass.
            float maxabs = 0.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) maxabs = fmaxf(maxabs, fmaxf(fabsf(lo3[d]), fabsf(hi3[d])));
            const float rcf = __double2float_ru(p.rc);
            const float keep_r = rcf * 1.000002f + 1.0e-6f * (maxabs + rcf);
            const float keep_r2 = keep_r * keep_r * 1.000001f;

            // ---- the cell's stencil: the expanded list if there is one, else pieces built here
            const int2 head = x.head ? x.head[bin] : make_int2(-1, 0);
            const bool listed = head.x >= 0;
            int total = head.y;
            const int npieces = 2 * nrows;
            if (!listed) {
                __syncwarp();
                rb_build_pieces(sm.pc, g, cstart, c0, c1, c2, lane);
                __syncwarp();
                total = sm.pc.total;
            }
            const int cfirst = (int)(((long long)total * slice) / nq), clast = (int)(((long long)total * (slice + 1)) / nq);
            const double4* xq = x.q + (listed ? head.x : 0);
            const int2* xi = x.info + (listed ? head.x : 0);

            // ---- candidates of the slice, 128 per round (four independent loads in flight), survivors kept
            // in list order
            int nst = 0, flags = RP_FIRST;
            for (int gi0 = cfirst; gi0 < clast; gi0 += 128) {
                if (nst + 128 > RP_CAND) {                     // stage full: hand it over, go on in the other one
                    if (lane == 0) { S->n = nst; S->flags = flags; }
                    publish();
                    S = &acquire();
                    nst = 0; flags = 0;
                }
                double4 q[4];
                int2 info[4];
                bool keep[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int gi = gi0 + 32 * h + lane;
                    keep[h] = gi < clast;
                    q[h] = make_double4(3.0e30, 3.0e30, 3.0e30, 0.0);
                    info[h] = make_int2(0, 0);
                    if (keep[h]) {
                        if (listed) { q[h] = xq[gi]; info[h] = xi[gi]; }
                        else rp_stage_candidate(g, spos, sm.pc, npieces, gi, q[h], info[h]);
                    }
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float xf = __double2float_rn(q[h].x), yf = __double2float_rn(q[h].y), zf = __double2float_rn(q[h].z);
                    const float ex = fmaxf(fmaxf(lo3[0] - xf, xf - hi3[0]), 0.0f);
                    const float ey = fmaxf(fmaxf(lo3[1] - yf, yf - hi3[1]), 0.0f);
                    const float ez = fmaxf(fmaxf(lo3[2] - zf, zf - hi3[2]), 0.0f);
                    keep[h] = keep[h] && fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= keep_r2;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep[h]);
                    if (keep[h]) {
                        const int slot = nst + __popc(bal & ((1u << lane) - 1u));
                        S->candd[slot] = q[h];
                        S->cinfo[slot] = info[h];
                    }
                    nst += __popc(bal);
                }
            }
            if (lane == 0) { S->n = nst; S->flags = flags | RP_LAST; }
            publish();
        }
        RpStage& S = acquire();
        if (lane == 0) { S.n = 0; S.flags = RP_EXIT; }
        publish();
        return;
    }

    // ==================================================================== consumer warps: lane = site
    int s = 0, ob = 0, oe = 0, ex0 = -1, ex1 = -1, wsite = pack_wraps(0, 0, 0), cnt = 0, lo_hi = 0, rep = 0, slice = 0;
    unsigned amb_width = 0;
    bool active = false, warp_active = false, multi = false;
    double px = 0.0, py = 0.0, pz = 0.0, qx = 3.0e150, qy = 3.0e150, qz = 3.0e150, acc12 = 0.0, acc6 = 0.0;
    long long lo_bits = 0, hi_bits = 0;
    for (int nuse = 0;; ++nuse) {
        const int sidx = nuse & 1;
        rp_bar_sync(1 + sidx);
        const RpStage& S = sm.st[sidx];
        const int flags = S.flags;
        if (flags & RP_EXIT) break;
        if (flags & RP_FIRST) {
            rep = S.rep; slice = S.slice;
            active = tid < S.nsites;
            warp_active = wid * 32 < S.nsites;
            acc12 = 0.0; acc6 = 0.0; cnt = 0;
            // idle lanes: r2 overflows, never inside
            qx = 3.0e150; qy = 3.0e150; qz = 3.0e150; px = 0.0; py = 0.0; pz = 0.0;
            wsite = pack_wraps(0, 0, 0); ex0 = -1; ex1 = -1; ob = 0; oe = 0; s = 0;
            double band = 16.0;
            if (active) {
                const double4 me = S.spos[tid];
                const int4 mt = S.smeta[tid];
                px = me.x; py = me.y; pz = me.z;
                wsite = (int)__double_as_longlong(me.w);
                s = mt.x; ex0 = mt.y; ob = mt.z; oe = mt.w;
                qx = px; qy = py; qz = pz;
                if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(b.grid[rep], p.rc, px, py, pz, wsite, qx, qy, qz, band);
            }
            // membership on the bit pattern of the fused r2: below lo_bits inside, above hi_bits outside,
            // in between the oracle's arithmetic decides (rows_block.cuh)
            const double bw = p.rc2 * band * 2.220446049250313e-16;
            lo_bits = __double_as_longlong(p.rc2 - bw); hi_bits = __double_as_longlong(p.rc2 + bw);
            lo_hi = (int)(lo_bits >> 32);
            amb_width = (unsigned)((int)(hi_bits >> 32) - lo_hi);
            multi = (oe - ob) > 2;
            ex1 = (oe - ob) > 1 ? tb.old_idx[ob + 1] : -1;
        }
        if (flags & RP_GENERIC) {
            if (active && slice == 0) {
                const GridDesc& g = b.grid[rep];
                rb_generic_row(g, b.spos + g.atom_off, b.cell_start + (size_t)rep * b.cs_stride, p, px, py, pz,
                               tb.old_idx, ob, oe, acc12, acc6, cnt);
            }
        } else if (warp_active) {
            const int n = S.n;
            auto rare = [&](int i, int U) {
                const GridDesc& g = b.grid[rep];
                const double4* spos = b.spos + g.atom_off;
                for (int u = 0; u < U; ++u) {
                    const double4 q = S.candd[i + u];
                    const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const long long rb = __double_as_longlong(r2);
                    bool in = rb < lo_bits;
                    if (!in && rb <= hi_bits)        // within rounding of the cutoff: the oracle's arithmetic decides
                        in = rp_exact_r2(g, spos, S.cinfo[i + u], px, py, pz, wsite) <= p.rc2;
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
                    const double4 q = S.candd[i + u];     // uniform address: two broadcast LDS.128
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
        if ((flags & RP_LAST) && active) {
            // row = 4 eps (sum c12 - sum c6) - e0 * pairs
            const double row = fma(p.eps4, acc12 - acc6, -p.e0 * (double)cnt);
            w.site_E[(size_t)s * nq + slice] = row;
            w.site_cnt[(size_t)s * nq + slice] = cnt;
        }
        __threadfence_block();
        __syncwarp();
        rp_bar_arrive(3 + sidx);
    }
}
