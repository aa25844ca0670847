// f2, first slice: the one-call-per-trial chain WITHOUT a launch per trial.
//
// An unmodified mcpy run asks for one energy per trial move, sequentially (GrandCanonicalEnsemble.do_gcmc_step,
// grand_canonical_ensemble.py:244-308).  Round 1 answered each call with ONE kernel launch (resident base + the
// trial in the kernel parameters + the result in pinned memory) and measured 31 us per call, most of it launch and
// retirement latency plus a chain of dependent loads.  Here ONE CTA stays resident while the chain is hot and is fed
// through a mailbox in page-locked host memory:
//
//   host   writes the trial -- the positions of the removed and of the added atoms (it mirrors the base, so the device
//          looks nothing up) and the removed atoms' indices -- into 64-byte lines of the mailbox, each line ending
//          in the sequence number (x86 stores are ordered, so a line that shows the new number is complete)
//   device warp 0 polls the lines with one coalesced read per poll; each warp then takes sites of the trial, spreads
//          the cell's candidate list over its lanes with the arithmetic of the single-launch kernel and folds the
//          four classes in list order -- same bits as every other row-sum path -- and thread 0 combines the rows
//          (plus the pairs inside the moved sets, as combine_trial does) and answers with ONE 16-byte store:
//          dE and (status, sequence number)
//   host   polls the sequence number of the answer
//
// The CTA leaves on its own after CHAIN_IDLE_US without a request (so a device-wide synchronisation elsewhere in the
// process waits a millisecond at most, and a host that died does not leave a spinning kernel behind), when the host
// raises `stop` (before anything the kernel reads is rebuilt: upload, re-base, new stencil lists, destroy), or after
// CHAIN_LIFETIME_US in any case; the host relaunches it on the next call.  Served: trials of up to CHAIN_MAX_SITES
// sites (the difference between the resident base and the incoming configuration: accepted moves accumulate until
// the next re-base) against a box with stencil lists; everything else takes the launch path (status
// CHAIN_ST_FALLBACK tells the host so).
#pragma once
#include "rows_block.cuh"
#include "emt_kernels.cuh"

#define CHAIN_THREADS 256
#define CHAIN_WARPS (CHAIN_THREADS / 32)
#define CHAIN_MAX_OLD 6
#define CHAIN_MAX_SITES 12
#define CHAIN_LINES 7                     // line 0: counts and indices; lines 1..6: two positions each
#define CHAIN_CAND 256                    // candidates of one site evaluated per pass of a warp
#define CHAIN_IDLE_US 1500
#define CHAIN_LIFETIME_US 2000000
#define CHAIN_ST_OK 0u
#define CHAIN_ST_FALLBACK 1u

struct __align__(64) ChainLine0 { int n_old, n_new; int old_idx[CHAIN_MAX_OLD]; int new_Z[CHAIN_MAX_OLD]; unsigned int pad[1]; unsigned int seq; };
struct __align__(64) ChainLineP { double pos[2][3]; unsigned int pad[3]; unsigned int seq; };

struct __align__(64) ChainMailbox {
    // request (host -> device).  LJ: sites are the removed atoms' (base) positions first, then the added ones;
    // EMT: the added atoms' positions only (the kernel body reads the removed ones from the resident batch)
    ChainLine0 l0;
    ChainLineP lp[CHAIN_LINES - 1];
    // control
    volatile unsigned int stop;
    volatile unsigned int alive;          // 1 while the CTA is resident; cleared when it leaves
    unsigned int pad1[14];
    // response (device -> host): one 16-byte store
    volatile double dE;
    volatile unsigned long long status_seq;   // (status << 32) | seq
};

__device__ __forceinline__ unsigned long long chain_clock_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct ChainSmem {
    GridDesc g;
    unsigned int req[16 * CHAIN_LINES];   // the request lines
    double c6[CHAIN_WARPS][CHAIN_CAND];
    unsigned char hit[CHAIN_WARPS][CHAIN_CAND];
    double row[CHAIN_MAX_SITES];
    int row_cnt[CHAIN_MAX_SITES];
    int status;
    int cmd;
};

__global__ void __launch_bounds__(CHAIN_THREADS, 1)
lj_chain_kernel(BatchView b, LJParams p, StencilLists x, ChainMailbox* mb, unsigned int start_seq) {
    __shared__ ChainSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // the base does not change while this CTA is resident: its grid descriptor is read once
    {
        const unsigned int* src = (const unsigned int*)&b.grid[0];
        unsigned int* dst = (unsigned int*)&sm.g;
        for (int k = tid; k < (int)(sizeof(GridDesc) / 4); k += CHAIN_THREADS) dst[k] = src[k];
    }
    __syncthreads();
    const GridDesc& g = sm.g;
    const double4* spos = b.spos + g.atom_off;
    const int nrows = (2 * g.m[1] + 1) * (2 * g.m[2] + 1);
    const bool fast = (!g.pbc[0] || g.ncell[0] >= 2 * g.m[0] + 1) && 2 * nrows <= RB_MAXPIECES && x.head != nullptr;
    unsigned int last = start_seq;
    const unsigned long long t_start = chain_clock_ns();
    unsigned long long t_idle = t_start;
    const volatile unsigned int* lines = (const volatile unsigned int*)&mb->l0;
    for (;;) {
        // ---- poll: warp 0 reads the request lines (16 words each) and checks the sequence numbers of the used ones
        if (wid == 0) {
            int cmd = -1;
            while (cmd < 0) {
                unsigned int wv[(16 * CHAIN_LINES + 31) / 32];
#pragma unroll
                for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k)
                    wv[k] = (32 * k + lane) < 16 * CHAIN_LINES ? lines[32 * k + lane] : 0u;
                const unsigned int s0 = __shfl_sync(0xffffffffu, wv[0], 15);
                bool fresh = s0 != last;
                if (fresh) {
                    const int n_old = (int)__shfl_sync(0xffffffffu, wv[0], 0), n_new = (int)__shfl_sync(0xffffffffu, wv[0], 1);
                    int nsites = n_old + n_new;
                    nsites = max(0, min(nsites, CHAIN_MAX_SITES));
                    const int used = 1 + (nsites + 1) / 2;                  // lines that carry this trial
#pragma unroll
                    for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k) {
                        // the sequence words sit at word 15 of every line: lanes 15 and 31 of each 32-word load
                        const unsigned int sa = __shfl_sync(0xffffffffu, wv[k], 15), sb = __shfl_sync(0xffffffffu, wv[k], 31);
                        if (2 * k < used && sa != s0) fresh = false;
                        if (2 * k + 1 < used && sb != s0) fresh = false;
                    }
                }
                if (fresh) {
#pragma unroll
                    for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k)
                        if ((32 * k + lane) < 16 * CHAIN_LINES) sm.req[32 * k + lane] = wv[k];
                    cmd = 0;
                    break;
                }
                int leave = 0;
                if (lane == 0) {
                    const unsigned long long now = chain_clock_ns();
                    leave = (mb->stop || now - t_idle > (unsigned long long)CHAIN_IDLE_US * 1000ULL ||
                             now - t_start > (unsigned long long)CHAIN_LIFETIME_US * 1000ULL) ? 1 : 0;
                }
                leave = __shfl_sync(0xffffffffu, leave, 0);
                if (leave) { cmd = 1; break; }
                __nanosleep(64);
            }
            if (lane == 0) sm.cmd = cmd;
        }
        __syncthreads();
        if (sm.cmd == 1) break;
        const ChainLine0& l0 = *(const ChainLine0*)&sm.req[0];
        const ChainLineP* lp = (const ChainLineP*)&sm.req[16];
        const unsigned int seq = l0.seq;
        const int n_old = l0.n_old, n_new = l0.n_new, nsites = n_old + n_new;
        const bool serve = fast && n_old >= 0 && n_new >= 0 && nsites >= 1 && nsites <= CHAIN_MAX_SITES && n_old <= CHAIN_MAX_OLD;
        if (tid == 0) sm.status = serve ? (int)CHAIN_ST_OK : (int)CHAIN_ST_FALLBACK;
        __syncthreads();
        if (serve) {
            // ---- one warp per site
            for (int s = wid; s < nsites; s += CHAIN_WARPS) {
                const double* ps = lp[s >> 1].pos[s & 1];
                const double px = ps[0], py = ps[1], pz = ps[2];
                int c[3], wr[3];
                locate(g, px, py, pz, c, wr);
                const int wsite = pack_wraps(wr[0], wr[1], wr[2]);
                const int cell_bin = (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
                const int2 head = x.head[cell_bin];
                if (head.x < 0) { if (lane == 0) sm.status = (int)CHAIN_ST_FALLBACK; continue; }
                MCPYB_CHECK(head.y >= 0 && (long long)head.x + head.y <= x.cap);
                const double4* xq = x.q + head.x;
                const int2* xi = x.info + head.x;
                double qx = px, qy = py, qz = pz, band = 16.0;
                if (wsite != pack_wraps(0, 0, 0)) rb_wrap_site(g, p.rc, px, py, pz, wsite, qx, qy, qz, band);
                const double bw = p.rc2 * band * 2.220446049250313e-16;
                const long long lo_bits = __double_as_longlong(p.rc2 - bw), hi_bits = __double_as_longlong(p.rc2 + bw);
                // class k = list positions = k (mod 4), each folded in list order (rows_warp.cuh); lane k < 4 carries it
                double a12 = 0.0, a6 = 0.0;
                int cc = 0;
                for (int cbase = 0; cbase < head.y; cbase += CHAIN_CAND) {
                    const int n = min(CHAIN_CAND, head.y - cbase);
                    for (int i = lane; i < n; i += 32) {
                        const int gi = cbase + i;
                        const double4 q = xq[gi];
                        const double dx = q.x - qx, dy = q.y - qy, dz = q.z - qz;
                        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        const long long rb = __double_as_longlong(r2);
                        bool in = rb < lo_bits;
                        if (!in && rb <= hi_bits) in = rp_exact_r2(g, spos, xi[gi], px, py, pz, wsite) <= p.rc2;
                        const int j = __double2loint(q.w);
                        for (int e = 0; e < n_old; ++e) if (l0.old_idx[e] == j) in = false;   // every removed atom
                        double c6 = 0.0;
                        if (in) { const double tt = p.sigma2 * fast_rcp(r2); c6 = tt * tt * tt; }
                        sm.c6[wid][i] = c6;
                        sm.hit[wid][i] = in ? 1 : 0;
                    }
                    __syncwarp();
                    if (lane < 4) {
                        for (int i = ((lane - cbase) & 3); i < n; i += 4) {
                            const double c6 = sm.c6[wid][i];
                            a12 = fma(c6, c6, a12);
                            a6 += c6;
                            cc += sm.hit[wid][i];
                        }
                    }
                    __syncwarp();
                }
                // fold the classes in the fixed order ((a0 + a1) + (a2 + a3))
                const double b12 = __shfl_sync(0xffffffffu, a12, 1), c12 = __shfl_sync(0xffffffffu, a12, 2), d12 = __shfl_sync(0xffffffffu, a12, 3);
                const double b6 = __shfl_sync(0xffffffffu, a6, 1), c6s = __shfl_sync(0xffffffffu, a6, 2), d6 = __shfl_sync(0xffffffffu, a6, 3);
                const int bc = __shfl_sync(0xffffffffu, cc, 1), ccn = __shfl_sync(0xffffffffu, cc, 2), dc = __shfl_sync(0xffffffffu, cc, 3);
                if (lane == 0) {
                    const double t12 = (a12 + b12) + (c12 + d12), t6 = (a6 + b6) + (c6s + d6);
                    const int tc = (cc + bc) + (ccn + dc);
                    sm.row[s] = fma(p.eps4, t12 - t6, -p.e0 * (double)tc);
                    sm.row_cnt[s] = tc;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            double dE = 0.0;
            if (sm.status == (int)CHAIN_ST_OK) {
                // combine_trial: sum(new rows) - sum(old rows) + the pairs inside the moved sets
                double plus = 0.0, minus = 0.0;
                int cnt = 0;
                for (int k = n_old; k < nsites; ++k) plus += sm.row[k];
                for (int k = 0; k < n_old; ++k) minus += sm.row[k];
                if (g.thin || n_new > 1 || n_old > 1) {
                    intra_set_lj(g, p, n_new,
                        [&](int a, double& xx, double& yy, double& zz) {
                            const double* q = lp[(n_old + a) >> 1].pos[(n_old + a) & 1];
                            xx = q[0]; yy = q[1]; zz = q[2];
                        }, 0, 1, plus, cnt);
                    intra_set_lj(g, p, n_old,
                        [&](int a, double& xx, double& yy, double& zz) {
                            const double* q = lp[a >> 1].pos[a & 1];
                            xx = q[0]; yy = q[1]; zz = q[2];
                        }, 0, 1, minus, cnt);
                }
                dE = plus - minus;
            }
            const unsigned long long ss = ((unsigned long long)(unsigned int)sm.status << 32) | (unsigned long long)seq;
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::
                         "l"(&mb->dE), "l"(__double_as_longlong(dE)), "l"(ss) : "memory");
            __threadfence_system();
        }
        last = seq;
        t_idle = chain_clock_ns();
        __syncthreads();
    }
    if (tid == 0) {
        __threadfence_system();
        mb->alive = 0;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same for EMT: the CTA runs the body of emt_trial_tiny_kernel (the whole CTA on one trial: moved atoms' rows,
// every affected neighbour re-embedded once with its summed density change) on what the mailbox holds.
// ---------------------------------------------------------------------------------------------------------
struct EmtChainSmem {
    unsigned int req[16 * CHAIN_LINES];
    EmtTinyTrial tt;
    double s_part[8];
    int s_parti[8];
    double dE;
    int cmd;
};

__global__ void __launch_bounds__(256, 1)
emt_chain_kernel(BatchView b, const EmtTable* __restrict__ tp, const int* __restrict__ z_to_slot,
                 const double* __restrict__ sigma1, const double* __restrict__ e_atom, ChainMailbox* mb,
                 unsigned int start_seq) {
    __shared__ EmtChainSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned int last = start_seq;
    const unsigned long long t_start = chain_clock_ns();
    unsigned long long t_idle = t_start;
    const volatile unsigned int* lines = (const volatile unsigned int*)&mb->l0;
    for (;;) {
        if (wid == 0) {
            int cmd = -1;
            while (cmd < 0) {
                unsigned int wv[(16 * CHAIN_LINES + 31) / 32];
#pragma unroll
                for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k)
                    wv[k] = (32 * k + lane) < 16 * CHAIN_LINES ? lines[32 * k + lane] : 0u;
                const unsigned int s0 = __shfl_sync(0xffffffffu, wv[0], 15);
                bool fresh = s0 != last;
                if (fresh) {
                    int n_new = (int)__shfl_sync(0xffffffffu, wv[0], 1);
                    n_new = max(0, min(n_new, CHAIN_MAX_OLD));
                    const int used = 1 + (n_new + 1) / 2;
#pragma unroll
                    for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k) {
                        const unsigned int sa = __shfl_sync(0xffffffffu, wv[k], 15), sb = __shfl_sync(0xffffffffu, wv[k], 31);
                        if (2 * k < used && sa != s0) fresh = false;
                        if (2 * k + 1 < used && sb != s0) fresh = false;
                    }
                }
                if (fresh) {
#pragma unroll
                    for (int k = 0; k < (16 * CHAIN_LINES + 31) / 32; ++k)
                        if ((32 * k + lane) < 16 * CHAIN_LINES) sm.req[32 * k + lane] = wv[k];
                    cmd = 0;
                    break;
                }
                int leave = 0;
                if (lane == 0) {
                    const unsigned long long now = chain_clock_ns();
                    leave = (mb->stop || now - t_idle > (unsigned long long)CHAIN_IDLE_US * 1000ULL ||
                             now - t_start > (unsigned long long)CHAIN_LIFETIME_US * 1000ULL) ? 1 : 0;
                }
                leave = __shfl_sync(0xffffffffu, leave, 0);
                if (leave) { cmd = 1; break; }
                __nanosleep(64);
            }
            if (lane == 0) sm.cmd = cmd;
        }
        __syncthreads();
        if (sm.cmd == 1) break;
        const ChainLine0& l0 = *(const ChainLine0*)&sm.req[0];
        const ChainLineP* lp = (const ChainLineP*)&sm.req[16];
        const unsigned int seq = l0.seq;
        const int n_old = l0.n_old, n_new = l0.n_new;
        const bool serve = n_old >= 0 && n_new >= 0 && n_old <= CHAIN_MAX_OLD && n_new <= CHAIN_MAX_OLD && n_old + n_new >= 1;
        if (serve) {
            if (tid == 0) { sm.tt.old_off[0] = 0; sm.tt.old_off[1] = n_old; sm.tt.new_off[0] = 0; sm.tt.new_off[1] = n_new; }
            if (tid < n_old) sm.tt.old_idx[tid] = l0.old_idx[tid];
            if (tid < n_new) {
                sm.tt.new_Z[tid] = l0.new_Z[tid];
                const double* q = lp[tid >> 1].pos[tid & 1];
                sm.tt.new_pos[3 * tid] = q[0]; sm.tt.new_pos[3 * tid + 1] = q[1]; sm.tt.new_pos[3 * tid + 2] = q[2];
            }
            __syncthreads();
            TrialBatch tb;
            tb.T = 1; tb.rep = nullptr; tb.old_off = sm.tt.old_off; tb.old_idx = sm.tt.old_idx;
            tb.new_off = sm.tt.new_off; tb.new_pos = sm.tt.new_pos; tb.new_Z = sm.tt.new_Z;
            emt_trial_delta_body<256>(b, tb, tp, z_to_slot, sigma1, e_atom, &sm.dE, nullptr, sm.s_part, sm.s_parti);
            __syncthreads();
        }
        if (tid == 0) {
            const unsigned int status = serve ? CHAIN_ST_OK : CHAIN_ST_FALLBACK;
            const unsigned long long ss = ((unsigned long long)status << 32) | (unsigned long long)seq;
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::
                         "l"(&mb->dE), "l"(__double_as_longlong(serve ? sm.dE : 0.0)), "l"(ss) : "memory");
            __threadfence_system();
        }
        last = seq;
        t_idle = chain_clock_ns();
        __syncthreads();
    }
    if (tid == 0) {
        __threadfence_system();
        mb->alive = 0;
        __threadfence_system();
    }
}
