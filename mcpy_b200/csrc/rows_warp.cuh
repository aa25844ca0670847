// K3 / K6, fourth generation: trial-move row sums, one WARP per (cell, <= 16 sites), no block-wide barrier.
//
// What the round-1 captures said about the CTA-per-cell kernel (profiles/README.md): its candidate loop is healthy
// (math-pipe throttle, 15 FP64-pipe instructions per candidate), but the loop is only half of the kernel -- every
// work item first walks a chain of dependent loads (claim -> item -> bin bounds -> sites -> list head), then stages
// and compacts its stencil behind two block-wide barriers, and with few sites per cell (replica batches: 8 x fewer
// sites per cell than one configuration with the same number of proposals) most lanes of the 64-site CTA idle.
// This kernel removes the fixed part instead of tuning the loop:
//
//   * the per-configuration stencil lists are PRUNED when they are written (expand_stencils_kernel): a candidate
//     image farther than rc from the box of the home cell is farther than rc from every site binned into that cell,
//     so it never enters the list (343 -> ~185 records per cell for the 2 000-atom LJ fluid).  The row-sum kernel
//     streams the records through a warp-private double buffer -- nothing to compact, no block-wide barrier;
//   * the scan kernel writes one fat 32-byte record per work item (bin, first site, site count, replica, list head
//     and length): the item's header is ONE load;
//   * a work item is (cell, <= 16 of its sites); a cell's sites are split into equal items.  The warp runs as 2 lane
//     groups of 16 (or 4 of 8 for items with <= 8 sites) that share the sites and split the candidates, so a replica
//     batch with a dozen sites per cell keeps its lanes busy and items stay short (the tail of the launch is one item).
//
// Bits.  A site's row sum is DEFINED as the fold ((a0 + a1) + (a2 + a3)) of four interleaved class sums, class k =
// the candidates at list positions = k (mod 4), each class accumulated sequentially in list order
// (acc12 = fma(c6, c6, acc12), acc6 += c6).  Two lane groups hold two classes each, four groups one each; the fold
// runs through shuffles in the same association order.  The single-launch kernels of the one-call-per-trial
// chain (rows_block.cuh) use the same definition, so a trial's dE does not depend on the batch it arrives in.
// Membership r2 <= rc^2 is decided as before: on the bit pattern of a fused r2, with the oracle's unfused
// arithmetic re-derived for a candidate within a few ulps of the cutoff.
#pragma once
#include "common.cuh"
#include "stencil.cuh"
#include "lj_kernels.cuh"
#include "site_kernels.cuh"
#include "rows_block.cuh"

#define RW_WARPS 4
#define RW_MIN_BLOCKS 4
#define RW_STAGES 4                  // 1 KB chunks of a cell's list in flight or in use per warp

__device__ __forceinline__ void rw_cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void rw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rw_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#define RW_SITES_SHIFT 4             // sites per work item: 16 (two lane groups) or <= 8 (four)

// ---------------------------------------------------------------------------------------------------------
// Work items: CTAs of 256 threads scan the bin histogram, one bin per thread, replica by replica; the totals of the
// CTAs are chained through `chain` (a CTA only ever waits for CTAs with a smaller index, which the hardware has
// scheduled before it).
//   chain[2 q], chain[2 q + 1] = (stamp << 36) | sites (items) of CTA q
//   items[2 k], items[2 k + 1] = (bin, first site in bin order, sites, replica), (list first, list length, -, -)
// ---------------------------------------------------------------------------------------------------------
#define RW_SCAN_THREADS 256

// CTA q = r * cpr + k scans bins [k, k + 1) * RW_SCAN_THREADS of replica r, one bin per thread; the totals of all
// CTAs before it in that order come through `chain`.
__global__ void __launch_bounds__(RW_SCAN_THREADS)
trial_scan_items_kernel(BatchView b, TrialWork w, StencilLists x, int cpr) {
    __shared__ int s_a[RW_SCAN_THREADS / 32], s_b[RW_SCAN_THREADS / 32];
    __shared__ long long s_base[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = blockIdx.x, r = q / cpr, k = q - r * cpr;
    pdl_wait();
    pdl_trigger();
    const int ncells = min(b.grid[r].ncells, w.bin_stride);
    const int i = k * RW_SCAN_THREADS + tid;                  // this thread's bin inside the replica
    const int bin = r * w.bin_stride + i;
    unsigned long long ev = ~0ULL;
    if (q == 0 && tid == 0) ev = *w.err;
    const int c = i < ncells ? w.hist[(size_t)bin * TRIAL_HIST_STRIDE] : 0;
    const int ni = (c + (1 << RW_SITES_SHIFT) - 1) >> RW_SITES_SHIFT;
    int2 head = make_int2(-1, 0);
    if (ni && x.head) head = x.head[bin];                    // in flight during the scan
    int ia = c, ib = ni;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb2 = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb2; }
    }
    if (lane == 31) { s_a[wid] = ia; s_b[wid] = ib; }
    __syncthreads();
    int wa = 0, wb = 0, ta_all = 0, tb_all = 0;
#pragma unroll
    for (int j = 0; j < RW_SCAN_THREADS / 32; ++j) {
        if (j < wid) { wa += s_a[j]; wb += s_b[j]; }
        ta_all += s_a[j]; tb_all += s_b[j];
    }
    const unsigned long long stamp = (unsigned long long)w.stamp << 36;
    if (tid == 0) {
        volatile unsigned long long* ch = w.chain;
        ch[2 * q] = stamp | (unsigned long long)ta_all;
        ch[2 * q + 1] = stamp | (unsigned long long)tb_all;
        __threadfence();
    }
    if (wid == 0) {
        long long base_a = 0, base_b = 0;
        volatile unsigned long long* ch = w.chain;
        for (int j = lane; j < q; j += 32) {
            unsigned long long va, vb;
            do { va = ch[2 * j]; } while ((va >> 36) != (unsigned long long)w.stamp);
            do { vb = ch[2 * j + 1]; } while ((vb >> 36) != (unsigned long long)w.stamp);
            base_a += (long long)(va & 0xFFFFFFFFFULL); base_b += (long long)(vb & 0xFFFFFFFFFULL);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            base_a += __shfl_xor_sync(0xffffffffu, base_a, o);
            base_b += __shfl_xor_sync(0xffffffffu, base_b, o);
        }
        if (lane == 0) { s_base[0] = base_a; s_base[1] = base_b; }
    }
    __syncthreads();
    const int ra = (int)s_base[0] + wa + ia - c, rb = (int)s_base[1] + wb + ib - ni;     // exclusive prefixes
    if (i < ncells) {
        if (c) w.hist[(size_t)bin * TRIAL_HIST_STRIDE] = 0;
        w.bin_count[bin] = ra;
        for (int j = 0; j < ni; ++j) {                           // equal shares of the cell's sites
            const int first = (int)(((long long)c * j) / ni), last = (int)(((long long)c * (j + 1)) / ni);
            w.items[2 * (rb + j)] = make_int4(bin, ra + first, last - first, r);
            w.items[2 * (rb + j) + 1] = make_int4(head.x, head.y, 0, 0);
        }
    }
    if (i == ncells) w.bin_count[bin] = ra;
    if (q == (int)gridDim.x - 1 && tid == 0) {
        w.totals[0] = (int)s_base[1] + tb_all;                   // work items of the batch
        w.totals[1] = 0;                                         // claim counter of the persistent warps
    }
    if (q == 0 && tid == 0) {
        if (w.err_out) *w.err_out = ev;
        *w.err = ~0ULL;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The candidate loop of one work item for G lane groups (G = 2, 4); `grp` = this lane's group.  The warp streams the
// cell's list through a private ring of RW_STAGES 1 KB chunks in shared memory, filled by cp.async (every lane copies
// one 32-byte record of a chunk; RW_STAGES - 1 chunks are in flight while one is evaluated, no registers held), so the
// loop never waits for a memory round trip per candidate -- the first build of this kernel read each record with a
// warp-uniform global load and ran at the latency of four loads in flight per warp (82 us against 20 us of FP64
// issue); a register double buffer (one chunk ahead) still left 2.2 warps per issue slot on the long scoreboard
// (profiles/r2_rows_warp_v1_full.csv).  A trip covers 4 G
// consecutive list positions; the lane evaluates positions i + grp + G u, u = 0..3 (classes (grp + G u) mod 4,
// ascending inside each class) into its 4 / G accumulator pairs.  The tail of the last chunk is padded with records
// that lie 1e150 away: misses, never ambiguous, not on any exclusion list.
// ---------------------------------------------------------------------------------------------------------
// F32 = true: fp32 pair math (north_star's second precision).  The distance vector is still formed in fp64
// (q - site: both are exact records) and rounded once; r2, the reciprocal and c6 run on the FP32 pipe; the class sums
// of a trip are added up in fp32 and go into the fp64 accumulators once per trip.  Membership stays bit-exact: a
// candidate whose fp32 r2 falls within 4e-7 of rc^2 is decided -- and evaluated -- by the fp64 path.
template <int G, bool F32>
__device__ __noinline__ void rw_candidate_loop(
        const double4* __restrict__ xq, const int2* __restrict__ xi, int len, int lane, int grp, double4 (*buf)[32],   // buf[RW_STAGES][32]
        double qx, double qy, double qz, int ksite, const double4* __restrict__ sorted_pos,
        long long lo_bits, long long hi_bits, int ex0, int ex1, bool multi, int ob, int oe,
        const GridDesc& g, const double4* __restrict__ spos, const TrialBatch& tb, const double sigma2, const double rc2,
        double& out12, double& out6, int& outc) {
    constexpr int NA = 4 / G;
    double a12[NA], a6[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) { a12[k] = 0.0; a6[k] = 0.0; }
    int cnt = 0;
    const int lo_hi = (int)(lo_bits >> 32), hi_hi = (int)(hi_bits >> 32);
    const unsigned amb_width = (unsigned)(hi_hi - lo_hi);
    // fp32 mode: the window around rc^2 inside which the fp64 arithmetic decides (fp32 rounding of d and r2: < 3e-7)
    const float lo_f = __double2float_rd(rc2 * (1.0 - 1.0e-6)), hi_f = __double2float_ru(rc2 * (1.0 + 1.0e-6));
    const float s2f = (float)sigma2;
    const double4 pad = make_double4(1.0e150, 1.0e150, 1.0e150, __longlong_as_double(-2LL));
    const int nchunks = (len + 31) >> 5;
    auto fetch = [&](int c) {                              // chunk c -> its ring slot (one commit group per call)
        if (c < nchunks) {
            const int k = (c << 5) + lane;
            double4* dst = &buf[c % RW_STAGES][lane];
            if (k < len) {
                rw_cp_async16(dst, xq + k);
                rw_cp_async16((char*)dst + 16, (const char*)(xq + k) + 16);
            } else {
                *dst = pad;
            }
        }
        rw_cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < RW_STAGES - 1; ++c) fetch(c);
    for (int c = 0; c < nchunks; ++c) {
        rw_cp_async_wait<RW_STAGES - 2>();                 // this lane's copies of chunk c have landed ...
        __syncwarp();                                      // ... and so have everybody else's; chunk c - 1 is read out
        fetch(c + RW_STAGES - 1);                          // refill the slot chunk c - 1 occupied
        const double4* cur = buf[c % RW_STAGES];
        const int here = min(32, len - (c << 5));
        for (int i = 0; i < here; i += 4 * G) {
            if (F32) {
                float r2f[4];
                int jj[4];
                bool amb = multi;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double4 q = cur[i + grp + G * u];
                    const float fx = (float)(q.x - qx), fy = (float)(q.y - qy), fz = (float)(q.z - qz);
                    r2f[u] = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
                    amb = amb || (r2f[u] >= lo_f && r2f[u] <= hi_f);
                    jj[u] = __double2loint(q.w);
                }
                if (!__any_sync(0xffffffffu, amb)) {
                    float p12[NA], p6[NA];
#pragma unroll
                    for (int k = 0; k < NA; ++k) { p12[k] = 0.0f; p6[k] = 0.0f; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool hit = r2f[u] < lo_f && jj[u] != ex0 && jj[u] != ex1;
                        const float t = s2f * __frcp_rn(r2f[u]);
                        float c6 = t * t * t;
                        c6 = hit ? c6 : 0.0f;
                        p12[u % NA] = fmaf(c6, c6, p12[u % NA]);
                        p6[u % NA] += c6;
                        if (hit) ++cnt;
                    }
#pragma unroll
                    for (int k = 0; k < NA; ++k) { a12[k] += (double)p12[k]; a6[k] += (double)p6[k]; }
                    continue;
                }
            }
            double r2[4];
            int jj[4], rh[4];
            unsigned amb = multi ? 0u : 0xffffffffu;       // min over the trip of (hi(r2) - lo_hi), unsigned
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double4 q = cur[i + grp + G * u];    // one address per lane group: broadcast LDS.128 x 2
                const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                r2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                rh[u] = __double2hiint(r2[u]);             // positive doubles order like their bits
                amb = min(amb, (unsigned)(rh[u] - lo_hi));
                jj[u] = __double2loint(q.w);
            }
            if (F32 || __any_sync(0xffffffffu, amb <= amb_width)) {
                // rare: some lane has a candidate within rounding of the cutoff (the oracle's arithmetic decides) or
                // a move of more than two atoms (exclusion list)
                double px = 0.0, py = 0.0, pz = 0.0;
                int wsite = pack_wraps(0, 0, 0);
                if (ksite >= 0) {
                    const double4 me = sorted_pos[ksite];
                    px = me.x; py = me.y; pz = me.z; wsite = (int)__double_as_longlong(me.w);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pos = (c << 5) + i + grp + G * u;
                    const long long rb = __double_as_longlong(r2[u]);
                    bool in = rb < lo_bits;
                    if (!in && rb <= hi_bits && pos < len) in = rp_exact_r2(g, spos, xi[pos], px, py, pz, wsite) <= rc2;
                    in = in && (jj[u] != ex0) && (jj[u] != ex1);
                    if (multi && in)
                        for (int e = ob + 2; e < oe; ++e) if (tb.old_idx[e] == jj[u]) in = false;
                    const double rm = __hiloint2double(in ? rh[u] : 0x7fe00000, __double2loint(r2[u]));
                    const double t = sigma2 * fast_rcp(rm);
                    const double c6 = t * t * t;
                    a12[u % NA] = fma(c6, c6, a12[u % NA]);
                    a6[u % NA] += c6;
                    cnt += in ? 1 : 0;
                }
                continue;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool hit = rh[u] < lo_hi && jj[u] != ex0 && jj[u] != ex1;
                // a miss is evaluated at r2 ~ 9e307: its reciprocal flushes to zero, so c6 = 0
                const double rm = __hiloint2double(hit ? rh[u] : 0x7fe00000, __double2loint(r2[u]));
                const double t = sigma2 * fast_rcp(rm);
                const double c6 = t * t * t;
                a12[u % NA] = fma(c6, c6, a12[u % NA]);
                a6[u % NA] += c6;
                if (hit) ++cnt;
            }
        }
    }
    rw_cp_async_wait<0>();
    __syncwarp();                                          // the ring is refilled by the next item
    // fold the four classes in the fixed order ((a0 + a1) + (a2 + a3)); the classes of a site live in G lane groups
    if (G == 2) {
        const double t12 = a12[0] + __shfl_xor_sync(0xffffffffu, a12[0], 16);
        const double u12 = a12[1 % NA] + __shfl_xor_sync(0xffffffffu, a12[1 % NA], 16);
        const double t6 = a6[0] + __shfl_xor_sync(0xffffffffu, a6[0], 16);
        const double u6 = a6[1 % NA] + __shfl_xor_sync(0xffffffffu, a6[1 % NA], 16);
        out12 = t12 + u12; out6 = t6 + u6;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
    } else {
        double t12 = a12[0] + __shfl_xor_sync(0xffffffffu, a12[0], 8);
        double t6 = a6[0] + __shfl_xor_sync(0xffffffffu, a6[0], 8);
        t12 += __shfl_xor_sync(0xffffffffu, t12, 16);
        t6 += __shfl_xor_sync(0xffffffffu, t6, 16);
        out12 = t12; out6 = t6;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
    }
    outc = cnt;
}

// One work item: sites [kbeg, kbeg + n) of sorted_pos / sorted_meta (n <= 16, all in one cell of replica r) against
// the cell's list `head` (first record, length; first < 0: no list).
template <bool F32>
__device__ __forceinline__ void rw_item(const BatchView& b, const TrialBatch& tb, const TrialWork& w, const LJParams& p,
                                        const StencilLists& x, double4 (*buf)[32], int lane,
                                        const double4* __restrict__ rec_pos, const int4* __restrict__ rec_meta,
                                        int kbeg, int n, int r, int2 head) {
    const GridDesc& g = b.grid[r];
    const double4* spos = b.spos + g.atom_off;
    // lane groups: sites sl = 0 .. SP-1 in each of G = 32 / SP groups
    const int gshift = n <= 8 ? 3 : 4;
    const int sl = lane & ((1 << gshift) - 1), grp = lane >> gshift;
    const bool active = sl < n;
    int s = 0, ob = 0, oe = 0, ex0 = -1, wsite = pack_wraps(0, 0, 0);
    double px = 0.0, py = 0.0, pz = 0.0;
    if (active) {
        const double4 me = rec_pos[(size_t)kbeg + sl];
        const int4 mt = rec_meta[(size_t)kbeg + sl];
        px = me.x; py = me.y; pz = me.z;
        wsite = (int)__double_as_longlong(me.w);
        s = mt.x; ex0 = mt.y; ob = mt.z; oe = mt.w;
    }
    double acc12 = 0.0, acc6 = 0.0;
    int cnt = 0;
    if (head.x < 0) {
        // no list for this cell (thin box, oversized stencil, list buffer full): exact per-lane traversal
        if (active && grp == 0) {
            const int* cstart = b.cell_start + (size_t)r * b.cs_stride;
            rb_generic_row(g, spos, cstart, p, px, py, pz, tb.old_idx, ob, oe, acc12, acc6, cnt);
        }
    } else {
        // the site wrapped into the primary cell and the band (ulps of rc^2) in which the fused r2 is re-checked
        double qx = 3.0e150, qy = -3.0e150, qz = 3.0e150;    // idle lanes: r2 overflows, never inside
        double band = 16.0;
        if (active) {
            qx = px; qy = py; qz = pz;
            if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(g, p.rc, px, py, pz, wsite, qx, qy, qz, band);
        }
        const double bw = p.rc2 * band * 2.220446049250313e-16;
        const long long lo_bits = __double_as_longlong(p.rc2 - bw), hi_bits = __double_as_longlong(p.rc2 + bw);
        const bool multi = (oe - ob) > 2;
        const int ex1 = (oe - ob) > 1 ? tb.old_idx[ob + 1] : -1;
        MCPYB_CHECK(head.x >= 0 && head.y >= 0 && (long long)head.x + head.y <= x.cap);
        const double4* xq = x.q + head.x;
        const int2* xi = x.info + head.x;
        const int ksite = active ? kbeg + sl : -1;
        MCPYB_CHECK(ksite < (w.slot_cap > 0 ? w.nbins * w.slot_cap : w.nsites));
        if (gshift == 4)
            rw_candidate_loop<2, F32>(xq, xi, head.y, lane, grp, buf, qx, qy, qz, ksite, rec_pos, lo_bits, hi_bits,
                                 ex0, ex1, multi, ob, oe, g, spos, tb, p.sigma2, p.rc2, acc12, acc6, cnt);
        else
            rw_candidate_loop<4, F32>(xq, xi, head.y, lane, grp, buf, qx, qy, qz, ksite, rec_pos, lo_bits, hi_bits,
                                 ex0, ex1, multi, ob, oe, g, spos, tb, p.sigma2, p.rc2, acc12, acc6, cnt);
    }
    if (active && grp == 0) {
        // row = 4 eps (sum c12 - sum c6) - e0 * pairs
        w.site_E[s] = fma(p.eps4, acc12 - acc6, -p.e0 * (double)cnt);
        w.site_cnt[s] = cnt;
    }
}

template <int MINB, bool F32>
__global__ void __launch_bounds__(32 * RW_WARPS, MINB)
lj_trial_rows_warp_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p, StencilLists x) {
    __shared__ double4 s_buf[RW_WARPS][RW_STAGES][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    pdl_wait();
    pdl_trigger();
    const int nwork = w.totals[0];
    // persistent warps: the first item is the warp's own index, the following ones come from a device counter
    // (zeroed by the scan kernel), claimed one item AHEAD so that the atomic's round trip hides behind the loop
    int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int next_wi = nwork;
    if (wi < nwork) {
        if (lane == 0) next_wi = nwarps + atomicAdd(&w.totals[1], 1);
        next_wi = __shfl_sync(0xffffffffu, next_wi, 0);
    }
    while (wi < nwork) {
        const int4 h0 = w.items[2 * wi], h1 = w.items[2 * wi + 1];
        int claimed = nwork;
        if (next_wi < nwork && lane == 0) claimed = nwarps + atomicAdd(&w.totals[1], 1);
        rw_item<F32>(b, tb, w, p, x, s_buf[wid], lane, w.sorted_pos, w.sorted_meta, h0.y, h0.z, h0.w, make_int2(h1.x, h1.y));
        wi = next_wi;
        next_wi = __shfl_sync(0xffffffffu, claimed, 0);
    }
}

// Direct placement (TrialWork::slot_cap > 0): the work items are the (bin, group of 16 ranks) pairs the site kernel
// announced; a group's sites are slots [16 g, min(16 g + 16, count)) of its bin, the count is the bin's histogram
// word.  No scan, no scatter.
template <int MINB>
__global__ void __launch_bounds__(32 * RW_WARPS, MINB)
lj_trial_rows_direct_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p, StencilLists x) {
    __shared__ double4 s_buf[RW_WARPS][RW_STAGES][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    pdl_wait();
    pdl_trigger();
    const int nwork = w.totals[0];
    int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int next_wi = nwork;
    if (wi < nwork) {
        if (lane == 0) next_wi = nwarps + atomicAdd(&w.totals[1], 1);
        next_wi = __shfl_sync(0xffffffffu, next_wi, 0);
    }
    while (wi < nwork) {
        const int2 it = w.item_tab[wi];
        int claimed = nwork;
        if (next_wi < nwork && lane == 0) claimed = nwarps + atomicAdd(&w.totals[1], 1);
        const int c = min(w.hist[(size_t)it.x * TRIAL_HIST_STRIDE], w.slot_cap);
        const int2 head = x.head ? x.head[it.x] : make_int2(-1, 0);
        const int first = it.y << RW_SITES_SHIFT;
        rw_item<false>(b, tb, w, p, x, s_buf[wid], lane, w.sorted_pos, w.sorted_meta, it.x * w.slot_cap + first,
                       min(1 << RW_SITES_SHIFT, c - first), it.x / w.bin_stride, head);
        wi = next_wi;
        next_wi = __shfl_sync(0xffffffffu, claimed, 0);
    }
}

// Sites that found their bin's slots full (records in site_pos / site_a, the bin in the upper half of .w): one warp
// per site through the same item code, so such a site's row carries the bits of every other listed row.
__global__ void __launch_bounds__(32 * RW_WARPS)
lj_trial_overflow_kernel(BatchView b, TrialBatch tb, TrialWork w, LJParams p, StencilLists x) {
    __shared__ double4 s_buf[RW_WARPS][RW_STAGES][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    pdl_wait();
    pdl_trigger();
    const int n = w.ovf[0];
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nwarps) {
        const int bin = (int)(__double_as_longlong(w.site_pos[k].w) >> 32);
        const int2 head = x.head ? x.head[bin] : make_int2(-1, 0);
        rw_item<false>(b, tb, w, p, x, s_buf[wid], lane, w.site_pos, w.site_a, k, 1, bin / w.bin_stride, head);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2 for large batches: the total energy of every resident configuration from the SAME row-sum pipeline --
// every atom is the removed atom of one trial (its row is the sum over its neighbours), e_atom = row / 2.
// A 64 x 2 000-atom batch takes the five launches of the pipeline (~0.1 ms) instead of 1.28 ms in the
// lane-per-atom kernel with its fp32 screen and per-lane FIFOs (profiles/README.md, round 2).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
self_trials_kernel(const int* __restrict__ atom_off, int R, int total, int* __restrict__ rep,
                   int* __restrict__ old_off, int* __restrict__ old_idx, int* __restrict__ new_off) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > total) return;
    old_off[t] = t;
    new_off[t] = 0;
    if (t < total) {
        const int r = find_replica(atom_off, R, t);
        rep[t] = r;
        old_idx[t] = t - atom_off[r];
    }
}

__global__ void __launch_bounds__(256)
rows_to_atom_energy_kernel(int total, const double* __restrict__ dE, const int* __restrict__ pairs,
                           double* __restrict__ e_atom, int* __restrict__ pair_count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    e_atom[t] = -0.5 * dE[t];                     // dE of removing atom t = -(its row)
    pair_count[t] = pairs[t];
}
