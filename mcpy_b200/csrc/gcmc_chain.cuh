// f2 -- device-resident GCMC chains: the loop of GrandCanonicalEnsemble.do_gcmc_step
// (mcpy/ensembles/grand_canonical_ensemble.py:244-308) without a host round trip per trial.
//
// ONE CTA (17 warps) runs one Markov chain.  For the length of a launch the chain's whole state image lives in the CTA's
// shared memory (2 000 atoms with room for 4 000: 96 KB of positions, 24 KB of order / cell-of / unsettled lists, 23 KB
// of cell lists, 25 KB of Mersenne-Twister states: 174 KB of the 227 KB; larger chains run from global memory):
//   configuration part (moves with the configuration in a replica swap)
//     px, py, pz [slot]      positions by storage slot (a slot is stable for the life of an atom)
//     order [index]          atom index -> slot: the reference's atom order (del atoms[i] shifts it, insertion appends);
//                            order[n..] holds the free slots
//     cell_of, unstable, uflag   cell of a slot; the atoms whose last wrap() still changed them
//     cell_cnt / cell_slots  cell list of edge >= rc/2 with a fixed number of slots per cell
//   slot part (stays with the replica slot)
//     mt  [stream][625]      MT19937 states of random.Random: MoveSelector.rng, each move's rng, rng_acceptance
//     mt2 [stream][624]      the NEXT state of each stream, computed ahead by the whole CTA (gc_mt_refill)
// Lane 0 of warp 0 is the Markov chain: the serial section (acceptance of trial t, commit, proposal t+1) and, while the
// other 16 warps sum the rows of a trial, the look-ahead (the draws of the next proposal, the de Broglie factor and the
// acceptance uniform of the running trial -- looked at, never consumed before their turn).  Warps 1..16 are the energy
// evaluator: a quad of lanes per (site, stencil cell), candidates by minimum image, the LJ pair sum folded in a fixed
// order.  O(N) commit work (order shift after a deletion, many-atom wrap, best-configuration copy) and the full energy
// re-sum at the end of a launch are done by the whole CTA.
//
// What is restated, with the reference's draw order (the contract is an identical accept / reject sequence):
//   MoveSelector.__get_index__ / do_trial_move           moves/move_selector.py:66-87
//   InsertionMove.do_trial_move (no min_insert)          moves/insertion_move.py:30-74   + Cell.get_random_point cell.py:34-41
//   DeletionMove.do_trial_move                           moves/deletion_move.py:19-47
//   DisplacementMove.do_trial_move (n_steps = 1)         moves/displacement_move.py:51-84
//   GrandCanonicalEnsemble._acceptance_condition         ensembles/grand_canonical_ensemble.py:197-242
//   Atoms.wrap() on acceptance (:296-297)                ase.geometry.wrap_positions, diagonal cell
//   BaseEnsemble._record_minimum                         ensembles/base_ensemble.py:142-150
//   random.Random: random(), _randbelow(), choice()      CPython Lib/random.py, Modules/_randommodule.c (MT19937)
#pragma once
#include "common.cuh"
#include "lj_kernels.cuh"
#include "pcg64.cuh"
#include "../../include/mcpy_b200.h"

#define GC_ROW_THREADS 512            // warps 1..16: the energy evaluator
#define GC_THREADS (GC_ROW_THREADS + 32)   // warp 0: the Markov chain (its lane 0), busy while the others sum rows
#define GC_QUADS (GC_ROW_THREADS / 4)
#define GC_NONE 0xFFFF
#define GC_SERIAL_WRAP 6          // unstable lists up to this length are wrapped by thread 0 itself

struct GcLayout {                 // byte offsets of the arrays inside a chain's state image
    // [0, mt): what belongs to the CONFIGURATION (moves with it in a replica swap); [mt, bytes): what belongs to the SLOT
    int px, py, pz, order, cell_of, unstable, uflag, cell_cnt, cell_slots, mt, mt2, bytes;
};

struct GcChain {                  // read-only description of one chain
    mcpyb_gcmc_desc d;
    double inv_box[3];
    LJParams lj;
    int ncell[3], span[3];
    int ncells, ccap, acap, nmt;
    GcLayout lay;
    unsigned char* blob;          // global memory: the configuration part of the state image, [0, lay.mt)
    unsigned char* slot;          // global memory: the slot part, [lay.mt, lay.bytes)
    double* best_pos;             // [acap][3], atom order
    mcpyb_gcmc_trial* trace;
    long long trace_cap;
    int in_smem;
};

struct GcDyn {                    // read-write scalars of one chain
    mcpyb_gcmc_scalars s;
    long long trace_len;
    long long cyc_serial, cyc_total, rounds;    // of the last launch: thread 0's serial section, the whole loop, loop trips
    long long cyc_pre;                          // ... and thread 0's look-ahead proposals (hidden behind the row sums)
    long long cyc_rows, cyc_wait;               // the row sums (first row thread); thread 0: end of look-ahead -> next serial section
    long long cyc_gap1;                         // thread 0: end of serial section -> start of look-ahead
    long long acc_since_resync;                 // accepted moves added onto E since it was last summed from scratch
    long long resyncs;
    double e_scale;                             // max |E| since then: bounds the rounding carried by E
    int n_unstable;
    int pad;
};

static inline int gc_align16(int x) { return (x + 15) & ~15; }
static inline GcLayout gc_layout(int acap, int ncells, int ccap, int nmt) {
    GcLayout l;
    int o = 0;
    l.px = o; o += gc_align16(8 * acap);
    l.py = o; o += gc_align16(8 * acap);
    l.pz = o; o += gc_align16(8 * acap);
    l.order = o; o += gc_align16(2 * acap);
    l.cell_of = o; o += gc_align16(2 * acap);
    l.unstable = o; o += gc_align16(2 * acap);
    l.uflag = o; o += gc_align16(acap);
    l.cell_cnt = o; o += gc_align16(2 * ncells);
    l.cell_slots = o; o += gc_align16(2 * ncells * ccap);
    l.mt = o; o += gc_align16(4 * MCPYB_GCMC_MT_WORDS * nmt);
    l.mt2 = o; o += gc_align16(4 * 624 * nmt);           // scratch: the NEXT state of every stream, computed ahead
    l.bytes = o;
    return l;
}

// ---- MT19937 as CPython drives it --------------------------------------------------------------------------
// A stream's state is 624 words + the position (random.Random.getstate()[1]).  During a launch the position lives
// in s_mt_idx, the words in one of two buffers: the canonical one (lay.mt) or the scratch (lay.mt2).  The NEXT
// state is a function of the current words alone, so the CTA computes it ahead, out of place, in three parallel
// phases (gc_mt_refill); when thread 0 has used up a state it just flips the buffers.  Without a precomputed state
// (two states used up within one trial) thread 0 regenerates in place, serially, as genrand_uint32 does.
#define GC_MAX_STREAMS (MCPYB_GCMC_MAX_MOVES + 2)
struct GcMt {
    uint32_t* a;                  // lay.mt : [stream][625]
    uint32_t* b;                  // lay.mt2: [stream][624]
    uint32_t* cur[GC_MAX_STREAMS];           // where the current state of a stream lives (a or b)
    int idx[GC_MAX_STREAMS];
    int flip_mask;                // bit k: the current state of stream k lives in b
    int ready_mask;               // bit k: the next state of stream k is complete in the other buffer
};
__device__ __forceinline__ uint32_t* gc_mt_cur(GcMt& M, int k) { return M.cur[k]; }
__device__ __forceinline__ uint32_t* gc_mt_other(GcMt& M, int k) { return (M.flip_mask >> k & 1) ? M.a + k * MCPYB_GCMC_MT_WORDS : M.b + k * 624; }
__device__ __forceinline__ uint32_t gc_mt_twist(uint32_t cur, uint32_t nxt) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t gc_mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// all threads: next state of stream k, current buffer -> other buffer (new[j] = old[j+397] ^ twist(old[j], old[j+1]) for
// j < 227, new[j] = new[j-227] ^ twist(old[j], old[j+1]) after that, the last word pairs old[623] with new[0])
__device__ __forceinline__ void gc_mt_refill(GcMt& M, int k, int tid) {
    const uint32_t* o = gc_mt_cur(M, k);
    uint32_t* w = gc_mt_other(M, k);
    if (tid < 227) w[tid] = o[tid + 397] ^ gc_mt_twist(o[tid], o[tid + 1]);
    __syncthreads();
    if (tid < 227) w[tid + 227] = w[tid] ^ gc_mt_twist(o[tid + 227], o[tid + 228]);
    __syncthreads();
    if (tid < 169) w[tid + 454] = w[tid + 227] ^ gc_mt_twist(o[tid + 454], o[tid + 455]);
    if (tid == 169) w[623] = w[396] ^ gc_mt_twist(o[623], w[0]);
    __syncthreads();
}
// thread 0: the state of stream k is used up
__device__ __noinline__ void gc_mt_advance(GcMt& M, int k) {
    if (M.ready_mask >> k & 1) {
        M.cur[k] = gc_mt_other(M, k); M.flip_mask ^= 1 << k; M.ready_mask &= ~(1 << k); M.idx[k] = 0;
        return;
    }
    uint32_t* s = gc_mt_cur(M, k);
    uint32_t cur = s[0];
    for (int j = 0; j < 623; ++j) {
        const int jm = (j + 397 >= 624) ? j + 397 - 624 : j + 397;
        const uint32_t nxt = s[j + 1];
        s[j] = s[jm] ^ gc_mt_twist(cur, nxt);
        cur = nxt;
    }
    s[623] = s[396] ^ gc_mt_twist(cur, s[0]);
    M.idx[k] = 0;
}
// K consecutive outputs of stream k; the common case loads and tempers them side by side
template <int K>
__device__ __forceinline__ void gc_mt_take(GcMt& M, int k, uint32_t (&out)[K]) {
    int i = M.idx[k];
    MCPYB_CHECK(k >= 0 && k < GC_MAX_STREAMS && i >= 0 && i <= 624);
    if (i + K <= 624) {
        const uint32_t* s = gc_mt_cur(M, k) + i;
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = s[j];
        M.idx[k] = i + K;
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = gc_mt_temper(out[j]);
        return;
    }
    for (int j = 0; j < K; ++j) {
        if (M.idx[k] >= 624) gc_mt_advance(M, k);
        out[j] = gc_mt_temper(gc_mt_cur(M, k)[M.idx[k]]);
        M.idx[k] += 1;
    }
}
// random.random(): (a * 2^26 + b) / 2^53 with a = 27 bits, b = 26 bits (exact in fp64)
__device__ __forceinline__ double gc_mt_double(uint32_t w0, uint32_t w1) {
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double gc_mt_random(GcMt& M, int k) {
    uint32_t w[2];
    gc_mt_take<2>(M, k, w);
    return gc_mt_double(w[0], w[1]);
}
// Random._randbelow_with_getrandbits(n), n >= 1: k = n.bit_length(); r = getrandbits(k) until r < n.  Four words are
// looked at side by side; only those up to the accepted one are consumed.
__device__ __forceinline__ uint32_t gc_mt_randbelow(GcMt& M, int k, uint32_t n) {
    const int sh = __clz(n);                     // 32 - bit_length
    while (true) {
        const int i = M.idx[k];
        if (i + 4 <= 624) {
            const uint32_t* s = gc_mt_cur(M, k) + i;
            const uint32_t r0 = gc_mt_temper(s[0]) >> sh, r1 = gc_mt_temper(s[1]) >> sh,
                           r2 = gc_mt_temper(s[2]) >> sh, r3 = gc_mt_temper(s[3]) >> sh;
            if (r0 < n) { M.idx[k] = i + 1; return r0; }
            if (r1 < n) { M.idx[k] = i + 2; return r1; }
            if (r2 < n) { M.idx[k] = i + 3; return r2; }
            M.idx[k] = i + 4;
            if (r3 < n) return r3;
        } else {
            uint32_t w[1];
            gc_mt_take<1>(M, k, w);
            const uint32_t r = w[0] >> sh;
            if (r < n) return r;
        }
    }
}

// Looking ahead without consuming (the proposal prepared while the CTA sums the rows of the trial before it): the words
// at offset `off` past the position of stream k, as long as the current state has them.
__device__ __forceinline__ bool gc_peek_randbelow(GcMt& M, int k, int& off, uint32_t n, uint32_t& r) {
    const int sh = __clz(n);
    const uint32_t* s = gc_mt_cur(M, k) + M.idx[k];
    const int avail = 624 - M.idx[k];
    while (off < avail) {
        const uint32_t v = gc_mt_temper(s[off]) >> sh;
        ++off;
        if (v < n) { r = v; return true; }
    }
    return false;
}
struct GcPre {                    // a proposal drawn ahead: valid as long as the generators have not moved and n == n_spec
    int valid, mv, kind, n_spec, used_mv, index, has_point;
    double a, b, c;               // displacement: the unit-ball vector; insertion: the point
    unsigned long long st_hi, st_lo;         // insertion: the PCG64 state after the point
};

// ase.geometry.wrap_positions for a diagonal cell and a periodic axis (center 0.5, eps 1e-7), operation for operation:
// fractional = x * (1/L) - shift, shift = -1e-7 (np.linalg.solve on a diagonal matrix multiplies by the reciprocal);
// fractional %= 1.0 (numpy remainder); fractional += shift; x' = fractional * L.
__device__ __forceinline__ double gc_wrap(double x, double inv, double L) {
    const double f = __dadd_rn(__dmul_rn(x, inv), 1e-7);
    double m = __dsub_rn(f, trunc(f));                 // fmod(f, 1.0), exact
    if (m != 0.0) { if (m < 0.0) m = __dadd_rn(m, 1.0); } else m = 0.0;
    return __dmul_rn(__dadd_rn(m, -1e-7), L);
}

__device__ __forceinline__ int gc_cell_coord(double x, double inv, int nc) {
    double f = x * inv;
    f -= floor(f);
    int c = (int)(f * (double)nc);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

struct GcTrial {                  // thread 0 -> CTA
    double sx[2], sy[2], sz[2];   // site positions: [0] removed (sign -1) or inserted; [1] the new position of a displacement
    double sign[2];
    int home[2][3];
    int nsites;
    int excl;                     // slot left out of the sums (the moved / deleted atom), GC_NONE for an insertion
    int done;
    int par;                      // parallel commit work: bit 0 shift the order, bit 1 wrap the unstable list, bit 2 copy the best
    int shift_from, shift_n;      // order[shift_from .. shift_n - 2] <- order[shift_from + 1 .. shift_n - 1]
    int freed_slot;
    int n;                        // atoms after the commit (for the best copy)
    int n_unstable;
    int refill;                   // bit k: the CTA computes the next state of Mersenne-Twister stream k in this round
    int resync;
};

#define GC_PAR_SHIFT 1
#define GC_PAR_WRAP 2
#define GC_PAR_BEST 4

__global__ void __launch_bounds__(GC_THREADS, 1)
gcmc_chain_kernel(const GcChain* __restrict__ chains, GcDyn* __restrict__ dyn, const long long* __restrict__ n_trials) {
    extern __shared__ __align__(16) unsigned char gc_smem[];
    __shared__ GcTrial T;
    __shared__ double s_part[GC_THREADS / 32];
    __shared__ unsigned long long s_pcg[MCPYB_GCMC_MAX_MOVES][4];
    __shared__ long long s_cnt[3][MCPYB_GCMC_MAX_MOVES];      // attempts, accepted, failed
    __shared__ GcMt M;
    __shared__ mcpyb_gcmc_desc s_d;                           // the chain's description, read by thread 0 all the time
    __shared__ signed char s_st[125][4];                      // stencil cell -> offsets along the three axes

    const GcChain& C = chains[blockIdx.x];
    GcDyn& D = dyn[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GcLayout lay = C.lay;
    unsigned char* base = C.in_smem ? gc_smem : C.blob;
    unsigned char* sbase = C.in_smem ? gc_smem + lay.mt : C.slot;
    if (C.in_smem) {
        const uint4* src = reinterpret_cast<const uint4*>(C.blob);
        uint4* dst = reinterpret_cast<uint4*>(gc_smem);
        for (int i = tid; i < lay.mt / 16; i += GC_THREADS) dst[i] = src[i];
        src = reinterpret_cast<const uint4*>(C.slot);
        dst = reinterpret_cast<uint4*>(gc_smem + lay.mt);
        for (int i = tid; i < (lay.bytes - lay.mt) / 16; i += GC_THREADS) dst[i] = src[i];
    }
    double* px = reinterpret_cast<double*>(base + lay.px);
    double* py = reinterpret_cast<double*>(base + lay.py);
    double* pz = reinterpret_cast<double*>(base + lay.pz);
    uint32_t* mt = reinterpret_cast<uint32_t*>(sbase);
    uint32_t* mt2 = reinterpret_cast<uint32_t*>(sbase + (lay.mt2 - lay.mt));
    unsigned short* order = reinterpret_cast<unsigned short*>(base + lay.order);
    unsigned short* cell_of = reinterpret_cast<unsigned short*>(base + lay.cell_of);
    unsigned short* unstable = reinterpret_cast<unsigned short*>(base + lay.unstable);
    unsigned char* uflag = base + lay.uflag;
    unsigned short* cell_cnt = reinterpret_cast<unsigned short*>(base + lay.cell_cnt);
    unsigned short* cell_slots = reinterpret_cast<unsigned short*>(base + lay.cell_slots);

    for (int i = tid; i < (int)(sizeof(mcpyb_gcmc_desc) / 4); i += GC_THREADS)
        reinterpret_cast<uint32_t*>(&s_d)[i] = reinterpret_cast<const uint32_t*>(&C.d)[i];
    const double L0 = C.d.box[0], L1 = C.d.box[1], L2 = C.d.box[2];
    const double i0 = C.inv_box[0], i1 = C.inv_box[1], i2 = C.inv_box[2];
    const int nc0 = C.ncell[0], nc1 = C.ncell[1], nc2 = C.ncell[2];
    const int sp0 = C.span[0], sp1 = C.span[1], sp2 = C.span[2];
    const int ns = sp0 * sp1 * sp2;
    const int ccap = C.ccap;
    const LJParams lj = C.lj;
    const int nm = C.d.n_move_types;
    for (int sc = tid; sc < ns; sc += GC_THREADS) {
        s_st[sc][2] = (signed char)(sc % sp2);
        s_st[sc][1] = (signed char)((sc / sp2) % sp1);
        s_st[sc][0] = (signed char)(sc / (sp2 * sp1));
    }
    __syncthreads();                                        // the state block is in shared memory: thread 0 reads it below

    // thread 0's chain state
    int n = 0, n_unst = 0, pending = 0, cur_move = 0, cur_kind = 0, cur_index = 0, have_rows = 0, drew = 0;
    long long t_done = 0, t_goal = 0, trace_len = 0;
    double E = 0.0, best_score = 0.0, best_E = 0.0, e_scale = 0.0;
    long long acc_since = 0, resyncs = 0;
    int best_n = -1;
    GcPre pre;
    pre.valid = 0;
    double acc_db = 1.0, acc_u = 0.0;
    bool acc_u_ok = false;
    const int all_streams = (1 << (nm + 2)) - 1;
    if (tid == 0) {
        n = D.s.n_atoms; n_unst = D.n_unstable; E = D.s.energy; best_score = D.s.best_score; best_E = D.s.best_energy;
        best_n = D.s.best_n; t_done = D.s.trials_done; t_goal = t_done + n_trials[blockIdx.x]; trace_len = 0;
        for (int k = 0; k < MCPYB_GCMC_MAX_MOVES; ++k)
            for (int j = 0; j < 4; ++j) s_pcg[k][j] = D.s.pcg[k][j];
        for (int k = 0; k < MCPYB_GCMC_MAX_MOVES; ++k) { s_cnt[0][k] = D.s.attempts[k]; s_cnt[1][k] = D.s.accepted[k]; s_cnt[2][k] = D.s.failed[k]; }
        T.done = 0; T.par = 0; T.nsites = 0; T.refill = 0; T.resync = 0;
        for (int a = 0; a < 2; ++a) { T.sx[a] = T.sy[a] = T.sz[a] = 0.0; T.sign[a] = 0.0; T.home[a][0] = T.home[a][1] = T.home[a][2] = 0; }
        D.s.status = MCPYB_GCMC_DONE;
        M.a = mt; M.b = mt2;
        for (int k = 0; k < nm + 2; ++k) { M.idx[k] = (int)mt[k * MCPYB_GCMC_MT_WORDS + 624]; M.cur[k] = mt + k * MCPYB_GCMC_MT_WORDS; }
        M.flip_mask = 0; M.ready_mask = 0;
        e_scale = fmax(D.e_scale, fabs(E)); acc_since = D.acc_since_resync; resyncs = D.resyncs;
    }
    __syncthreads();                                        // (the state block, s_d and the stencil table are in place)

    long long cyc_serial = 0, rounds = 0, cyc_pre = 0, cyc_rows = 0, cyc_wait = 0, cyc_gap1 = 0, mark_serial_end = 0, mark_pre_end = 0;
    const long long cyc_begin = clock64();
    while (true) {
        if (tid == 0) {
            const long long cyc_in = clock64();
            if (mark_pre_end) { cyc_wait += cyc_in - mark_pre_end; mark_pre_end = 0; }
            // ---------------- the Markov chain proper (serial) ----------------
            bool proposing = true;
            if (T.refill) { M.ready_mask |= T.refill; T.refill = 0; }   // completed by the CTA in the last parallel section
            if (T.par) { T.par = 0; }                       // the parallel commit of the last trial is done
            else if (have_rows) {
                have_rows = 0;
                double dE = 0.0;
                for (int w = 0; w < GC_ROW_THREADS / 32; ++w) dE += s_part[w];
                // _acceptance_condition; n is the pre-move atom count (self.n_atoms)
                bool acc;
                drew = 0;
                const int mt_acc = nm + 1;
                // rng_acceptance.get_uniform(): the two words looked at behind the row sums, consumed here
                auto draw = [&]() {
                    drew = 1;
                    if (acc_u_ok) { M.idx[mt_acc] += 2; return acc_u; }
                    return gc_mt_random(M, mt_acc);
                };
                if (cur_kind == MCPYB_GCMC_DISPLACE) {
                    if (dE <= 0.0) acc = true;
                    else { const double p = exp(__dmul_rn(-dE, s_d.beta)); acc = p > draw(); }
                } else {
                    double p;
                    if (cur_kind == MCPYB_GCMC_INSERT) p = __dmul_rn(acc_db, exp(__dmul_rn(-s_d.beta, __dsub_rn(dE, s_d.mu))));
                    else p = __dmul_rn(acc_db, exp(__dmul_rn(-s_d.beta, __dadd_rn(dE, s_d.mu))));
                    if (p > 1.0) acc = true;
                    else acc = p > draw();
                }
                int par = 0;
                if (acc) {
                    s_cnt[1][cur_move] += 1;
                    int moved = GC_NONE;
                    if (cur_kind == MCPYB_GCMC_INSERT) {
                        MCPYB_CHECK(n >= 0 && n < C.acap);
                        const int slot = order[n];
                        MCPYB_CHECK(slot < C.acap);
                        px[slot] = T.sx[0]; py[slot] = T.sy[0]; pz[slot] = T.sz[0];
                        const int c = (T.home[0][0] * nc1 + T.home[0][1]) * nc2 + T.home[0][2];
                        MCPYB_CHECK(c >= 0 && c < C.ncells);
                        const int cnt = cell_cnt[c];
                        MCPYB_CHECK(cnt < ccap);
                        cell_slots[c * ccap + cnt] = (unsigned short)slot;
                        cell_cnt[c] = (unsigned short)(cnt + 1);
                        cell_of[slot] = (unsigned short)c;
                        if (cnt + 1 >= ccap) pending = MCPYB_GCMC_CELL_FULL;
                        ++n;
                        if (n >= C.acap) pending = MCPYB_GCMC_ATOMS_FULL;
                        moved = slot;
                    } else {
                        const int slot = T.excl;
                        const int c_old = cell_of[slot];
                        const int c_new = cur_kind == MCPYB_GCMC_DISPLACE
                                              ? (T.home[1][0] * nc1 + T.home[1][1]) * nc2 + T.home[1][2] : -1;
                        if (c_new != c_old) {
                            const int cnt = cell_cnt[c_old];
                            unsigned short* lst = cell_slots + c_old * ccap;
                            int k = 0;
                            while (k < cnt && lst[k] != slot) ++k;
                            MCPYB_CHECK(cnt > 0 && k < cnt);             // the atom was in the cell it is filed under
                            lst[k] = lst[cnt - 1];
                            cell_cnt[c_old] = (unsigned short)(cnt - 1);
                            if (c_new >= 0) {
                                const int cn = cell_cnt[c_new];
                                MCPYB_CHECK(c_new < C.ncells && cn < ccap);
                                cell_slots[c_new * ccap + cn] = (unsigned short)slot;
                                cell_cnt[c_new] = (unsigned short)(cn + 1);
                                cell_of[slot] = (unsigned short)c_new;
                                if (cn + 1 >= ccap) pending = MCPYB_GCMC_CELL_FULL;
                            }
                        }
                        if (cur_kind == MCPYB_GCMC_DISPLACE) {
                            px[slot] = T.sx[1]; py[slot] = T.sy[1]; pz[slot] = T.sz[1];
                            moved = slot;
                        } else {
                            // del atoms[i]: the order shifts down (in parallel), the slot joins the free ones
                            par |= GC_PAR_SHIFT;
                            T.shift_from = cur_index; T.shift_n = n; T.freed_slot = slot;
                            --n;
                            if (uflag[slot]) {
                                uflag[slot] = 0;
                                int k = 0;
                                while (k < n_unst && unstable[k] != slot) ++k;
                                if (k < n_unst) { unstable[k] = unstable[n_unst - 1]; --n_unst; }
                            }
                        }
                    }
                    if (moved != GC_NONE && !uflag[moved]) { uflag[moved] = 1; unstable[n_unst++] = (unsigned short)moved; }
                    MCPYB_CHECK(n_unst >= 0 && n_unst <= C.acap && n >= 0 && n < C.acap + 1);
                    E += dE;
                    e_scale = fmax(e_scale, fabs(E)); ++acc_since;
                    // atoms.wrap(): only atoms whose last wrap still changed a bit can change again
                    if (n_unst <= GC_SERIAL_WRAP) {
                        int keep = 0;
                        for (int k = 0; k < n_unst; ++k) {
                            const int s = unstable[k];
                            const double x = px[s], y = py[s], z = pz[s];
                            const double xw = gc_wrap(x, i0, L0), yw = gc_wrap(y, i1, L1), zw = gc_wrap(z, i2, L2);
                            if (xw != x || yw != y || zw != z) { px[s] = xw; py[s] = yw; pz[s] = zw; unstable[keep++] = (unsigned short)s; }
                            else uflag[s] = 0;
                        }
                        n_unst = keep;
                    } else {
                        par |= GC_PAR_WRAP;
                        T.n_unstable = n_unst;
                    }
                    // _record_minimum: score = E - mu * N
                    const double score = __dsub_rn(E, __dmul_rn(s_d.mu, (double)n));
                    if (score < best_score) { best_score = score; best_E = E; best_n = n; par |= GC_PAR_BEST; }
                    T.n = n;
                }
                if (C.trace) {
                    mcpyb_gcmc_trial r;
                    r.dE = dE; r.move = cur_move; r.flags = (acc ? 1 : 0) | (drew ? 4 : 0); r.n_atoms = n; r.index = cur_index;
                    C.trace[trace_len++] = r;
                }
                ++t_done;
                if (par) { T.par = par; proposing = false; }
            }
            // ---------------- next proposal ----------------
            while (proposing) {
                if (pending) { D.s.status = pending; T.done = 1; break; }
                if (t_done >= t_goal) { T.done = 1; break; }
                if (C.trace && trace_len >= C.trace_cap) { D.s.status = MCPYB_GCMC_TRACE_FULL; T.done = 1; break; }
                // a proposal drawn ahead during the last row sums stands for the draws below if nothing it assumed changed
                const bool use = pre.valid && (pre.kind != MCPYB_GCMC_DISPLACE || pre.n_spec == n);
                pre.valid = 0;
                int mv = 0;
                if (use) { mv = pre.mv; M.idx[0] += 2; }
                else {
                    // MoveSelector.__get_index__: searchsorted(rho, u * rho_total, side='right'), clamped
                    const double v = __dmul_rn(gc_mt_random(M, 0), s_d.move_cum_weight[nm - 1]);
                    while (mv < nm && s_d.move_cum_weight[mv] <= v) ++mv;
                    if (mv >= nm) mv = nm - 1;
                }
                cur_move = mv; cur_kind = s_d.move_kind[mv];
                const int mt_mv = 1 + mv;
                bool failed = false;
                if (cur_kind == MCPYB_GCMC_DISPLACE && n == 0) { D.s.status = MCPYB_GCMC_EMPTY_DISPLACE; T.done = 1; break; }
                s_cnt[0][mv] += 1;
                if (cur_kind == MCPYB_GCMC_INSERT) {
                    if (use) M.idx[mt_mv] += pre.used_mv;
                    else gc_mt_randbelow(M, mt_mv, 1u);                              // rng.random.choice(self.species)
                    cur_index = n;
                    if (s_d.move_limit[mv] >= 0 && n >= s_d.move_limit[mv]) failed = true;
                    else {
                        unsigned long long* ps = s_pcg[s_d.move_pcg[mv]];
                        if (use && pre.has_point) {
                            ps[0] = pre.st_hi; ps[1] = pre.st_lo;
                            T.sx[0] = pre.a; T.sy[0] = pre.b; T.sz[0] = pre.c;
                        } else {
                            u128 st = ((u128)ps[0] << 64) | ps[1];
                            const u128 inc = ((u128)ps[2] << 64) | ps[3];
                            const double u0 = pcg_next_double(st, inc), u1 = pcg_next_double(st, inc), u2 = pcg_next_double(st, inc);
                            ps[0] = (unsigned long long)(st >> 64); ps[1] = (unsigned long long)st;
                            T.sx[0] = __dmul_rn(u0, L0); T.sy[0] = __dmul_rn(u1, L1); T.sz[0] = __dmul_rn(u2, L2);
                        }
                        T.sign[0] = 1.0; T.nsites = 1; T.excl = GC_NONE;
                    }
                } else if (cur_kind == MCPYB_GCMC_DELETE) {
                    if (use) M.idx[mt_mv] += pre.used_mv;
                    else gc_mt_randbelow(M, mt_mv, 1u);
                    cur_index = -1;
                    if ((s_d.move_limit[mv] >= 0 && n <= s_d.move_limit[mv]) || n == 0) failed = true;
                    else {
                        cur_index = (int)gc_mt_randbelow(M, mt_mv, (uint32_t)n);     // rng.random.choice(indices_of_species)
                        const int slot = order[cur_index];
                        T.sx[0] = px[slot]; T.sy[0] = py[slot]; T.sz[0] = pz[slot];
                        T.sign[0] = -1.0; T.nsites = 1; T.excl = slot;
                    }
                } else {
                    double rx, ry, rz, rsq;
                    if (use) { M.idx[mt_mv] += pre.used_mv; cur_index = pre.index; rx = pre.a; ry = pre.b; rz = pre.c; }
                    else {
                        cur_index = (int)gc_mt_randbelow(M, mt_mv, (uint32_t)n);     // rng.random.choice(to_move)
                        do {
                            uint32_t w[6];
                            gc_mt_take<6>(M, mt_mv, w);
                            rx = __dsub_rn(__dmul_rn(2.0, gc_mt_double(w[0], w[1])), 1.0);
                            ry = __dsub_rn(__dmul_rn(2.0, gc_mt_double(w[2], w[3])), 1.0);
                            rz = __dsub_rn(__dmul_rn(2.0, gc_mt_double(w[4], w[5])), 1.0);
                            rsq = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                        } while (rsq > 1.0);
                    }
                    const int slot = order[cur_index];
                    const double md = s_d.move_max_displacement[mv];
                    T.sx[0] = px[slot]; T.sy[0] = py[slot]; T.sz[0] = pz[slot];
                    T.sx[1] = __dadd_rn(T.sx[0], __dmul_rn(rx, md));
                    T.sy[1] = __dadd_rn(T.sy[0], __dmul_rn(ry, md));
                    T.sz[1] = __dadd_rn(T.sz[0], __dmul_rn(rz, md));
                    T.sign[0] = -1.0; T.sign[1] = 1.0; T.nsites = 2; T.excl = slot;
                }
                if (failed) {
                    s_cnt[2][mv] += 1;
                    if (C.trace) {
                        mcpyb_gcmc_trial r;
                        r.dE = __longlong_as_double(0x7ff8000000000000LL); r.move = mv; r.flags = 2; r.n_atoms = n; r.index = cur_index;
                        C.trace[trace_len++] = r;
                    }
                    ++t_done;
                    continue;
                }
                for (int s = 0; s < T.nsites; ++s) {
                    T.home[s][0] = gc_cell_coord(T.sx[s], i0, nc0);
                    T.home[s][1] = gc_cell_coord(T.sy[s], i1, nc1);
                    T.home[s][2] = gc_cell_coord(T.sz[s], i2, nc2);
                }
                have_rows = 1;
                break;
            }
            if (!T.done) T.refill = ~M.ready_mask & all_streams;
            mark_serial_end = clock64();
            cyc_serial += mark_serial_end - cyc_in;
            ++rounds;
        }
        __syncthreads();
        const int ctl_done = T.done, ctl_refill = T.refill, ctl_par = T.par;     // three loads in flight together
        if (ctl_done) break;
        if (ctl_refill) {                                   // next Mersenne-Twister states, computed ahead by the whole CTA
            const int need = ctl_refill;
            for (int k = 0; k < nm + 2; ++k) if (need >> k & 1) gc_mt_refill(M, k, tid);
        }
        if (ctl_par) {
            const int par = ctl_par;
            if (par & GC_PAR_SHIFT) {
                const int from = T.shift_from, last = T.shift_n - 1;        // order[from .. last-1] <- order[from+1 .. last]
                for (int b0 = from; b0 < last; b0 += GC_THREADS) {
                    const int j = b0 + tid;
                    unsigned short v = 0;
                    if (j < last) v = order[j + 1];
                    __syncthreads();
                    if (j < last) order[j] = v;
                    __syncthreads();
                }
                if (tid == 0) order[last] = (unsigned short)T.freed_slot;
            }
            if (par & GC_PAR_WRAP) {
                const int nu = T.n_unstable;
                for (int k = tid; k < nu; k += GC_THREADS) {
                    const int s = unstable[k];
                    const double x = px[s], y = py[s], z = pz[s];
                    const double xw = gc_wrap(x, i0, L0), yw = gc_wrap(y, i1, L1), zw = gc_wrap(z, i2, L2);
                    if (xw != x || yw != y || zw != z) { px[s] = xw; py[s] = yw; pz[s] = zw; }
                    else { uflag[s] = 0; unstable[k] = GC_NONE; }
                }
                __syncthreads();
                if (tid == 0) {
                    int keep = 0;
                    for (int k = 0; k < nu; ++k) { const unsigned short s = unstable[k]; if (s != GC_NONE) unstable[keep++] = s; }
                    n_unst = keep;
                }
            }
            if (par & GC_PAR_BEST) {
                __syncthreads();
                const int nn = T.n;
                for (int j = tid; j < nn; j += GC_THREADS) {
                    const int s = order[j];
                    C.best_pos[3 * j] = px[s]; C.best_pos[3 * j + 1] = py[s]; C.best_pos[3 * j + 2] = pz[s];
                }
            }
            __syncthreads();
            continue;
        }
        // ---------------- row sums: a quad of lanes per (site, stencil cell) ----------------
        if (tid >= 32) {
            const long long cyc_in = clock64();
            const int rt = tid - 32;
            double acc = 0.0;
            const int ntasks = T.nsites * ns;
            const int excl = T.excl;
            for (int task = rt >> 2; task < ntasks; task += GC_QUADS) {
                const int site = task >= ns ? 1 : 0;
                const int sc = task - site * ns;
                const int a0 = s_st[sc][0], a1 = s_st[sc][1], a2 = s_st[sc][2];
                int c0 = a0, c1 = a1, c2 = a2;
                if (sp0 != nc0) { c0 = T.home[site][0] - 2 + a0; c0 += c0 < 0 ? nc0 : 0; c0 -= c0 >= nc0 ? nc0 : 0; }
                if (sp1 != nc1) { c1 = T.home[site][1] - 2 + a1; c1 += c1 < 0 ? nc1 : 0; c1 -= c1 >= nc1 ? nc1 : 0; }
                if (sp2 != nc2) { c2 = T.home[site][2] - 2 + a2; c2 += c2 < 0 ? nc2 : 0; c2 -= c2 >= nc2 ? nc2 : 0; }
                const int cell = (c0 * nc1 + c1) * nc2 + c2;
                MCPYB_CHECK(sc >= 0 && sc < ns && cell >= 0 && cell < C.ncells);
                const int cnt = cell_cnt[cell];
                MCPYB_CHECK(cnt >= 0 && cnt <= ccap);
                const double sx = T.sx[site], sy = T.sy[site], sz = T.sz[site], sg = T.sign[site];
                for (int k = rt & 3; k < cnt; k += 4) {
                    const int slot = cell_slots[cell * ccap + k];
                    MCPYB_CHECK(slot < C.acap);
                    if (slot == excl) continue;
                    const double x = px[slot], y = py[slot], z = pz[slot];
                    // the image of the neighbour the reference's neighbour list pairs with the site: (pos_j + S * L) - pos_i
                    const double S0 = -rint((x - sx) * i0), S1 = -rint((y - sy) * i1), S2 = -rint((z - sz) * i2);
                    const double dx = __dsub_rn(__dadd_rn(x, __dmul_rn(S0, L0)), sx);
                    const double dy = __dsub_rn(__dadd_rn(y, __dmul_rn(S1, L1)), sy);
                    const double dz = __dsub_rn(__dadd_rn(z, __dmul_rn(S2, L2)), sz);
                    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    if (r2 <= lj.rc2) acc += sg * lj_pair(lj, r2);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) s_part[warp - 1] = acc;
            cyc_rows += clock64() - cyc_in;
        } else if (tid == 0) {
            // ---------------- meanwhile: the draws of the NEXT proposal, looked at without consuming them ----------------
            // MoveSelector.rng, the move's rng and the cell's PCG64 are not touched by the acceptance test (its own stream),
            // so what the next proposal draws is known now -- except that an index is drawn below n, which the running
            // trial may change.  Drawn for the current n; used only if n is still that (or not needed) when it is its turn.
            const long long cyc_in = clock64();
            cyc_gap1 += cyc_in - mark_serial_end;
            // what the acceptance test of the running trial needs besides dE: the de Broglie factor (a division) and the
            // next uniform of the acceptance stream (looked at, consumed only if the test draws)
            if (cur_kind == MCPYB_GCMC_INSERT) acc_db = s_d.move_volume[cur_move] / __dmul_rn((double)(n + 1), s_d.lambda3);
            else if (cur_kind == MCPYB_GCMC_DELETE) acc_db = __dmul_rn(s_d.lambda3, (double)n) / s_d.move_volume[cur_move];
            acc_u_ok = M.idx[nm + 1] + 2 <= 624;
            if (acc_u_ok) {
                const uint32_t* sa = gc_mt_cur(M, nm + 1) + M.idx[nm + 1];
                acc_u = gc_mt_double(gc_mt_temper(sa[0]), gc_mt_temper(sa[1]));
            }
            pre.valid = 0;
            const uint32_t* s0 = gc_mt_cur(M, 0) + M.idx[0];
            if (t_done + 1 < t_goal && !pending && M.idx[0] + 2 <= 624) {
                const double v = __dmul_rn(gc_mt_double(gc_mt_temper(s0[0]), gc_mt_temper(s0[1])), s_d.move_cum_weight[nm - 1]);
                int mv = 0;
                while (mv < nm && s_d.move_cum_weight[mv] <= v) ++mv;
                if (mv >= nm) mv = nm - 1;
                const int kind = s_d.move_kind[mv], k = 1 + mv;
                pre.mv = mv; pre.kind = kind; pre.n_spec = n; pre.has_point = 0;
                int off = 0;
                uint32_t r = 0;
                if (kind != MCPYB_GCMC_DISPLACE) {
                    if (gc_peek_randbelow(M, k, off, 1u, r)) {           // rng.random.choice(self.species)
                        pre.used_mv = off;
                        pre.valid = 1;
                        if (kind == MCPYB_GCMC_INSERT && s_d.move_limit[mv] < 0) {
                            const unsigned long long* ps = s_pcg[s_d.move_pcg[mv]];
                            u128 st = ((u128)ps[0] << 64) | ps[1];
                            const u128 inc = ((u128)ps[2] << 64) | ps[3];
                            const double u0 = pcg_next_double(st, inc), u1 = pcg_next_double(st, inc), u2 = pcg_next_double(st, inc);
                            pre.st_hi = (unsigned long long)(st >> 64); pre.st_lo = (unsigned long long)st;
                            pre.a = __dmul_rn(u0, L0); pre.b = __dmul_rn(u1, L1); pre.c = __dmul_rn(u2, L2);
                            pre.has_point = 1;
                        }
                    }
                } else if (n > 0 && gc_peek_randbelow(M, k, off, (uint32_t)n, r)) {   // rng.random.choice(to_move)
                    pre.index = (int)r;
                    const uint32_t* s = gc_mt_cur(M, k) + M.idx[k];
                    const int avail = 624 - M.idx[k];
                    while (off + 6 <= avail) {
                        const double rx = __dsub_rn(__dmul_rn(2.0, gc_mt_double(gc_mt_temper(s[off]), gc_mt_temper(s[off + 1]))), 1.0);
                        const double ry = __dsub_rn(__dmul_rn(2.0, gc_mt_double(gc_mt_temper(s[off + 2]), gc_mt_temper(s[off + 3]))), 1.0);
                        const double rz = __dsub_rn(__dmul_rn(2.0, gc_mt_double(gc_mt_temper(s[off + 4]), gc_mt_temper(s[off + 5]))), 1.0);
                        off += 6;
                        const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                        if (!(rsq > 1.0)) { pre.a = rx; pre.b = ry; pre.c = rz; pre.used_mv = off; pre.valid = 1; break; }
                    }
                }
            }
            mark_pre_end = clock64();
            cyc_pre += mark_pre_end - cyc_in;
        }
        __syncthreads();
    }

    // ---------------- E summed from scratch when the running sum may carry too much rounding ----------------
    // E = E_start + sum of accepted dE: every addition rounds at the size of E at that time, so a chain that came down
    // from a large |E| (overlapping start, a hot rung) keeps an absolute error the reference's per-trial full energy
    // does not have.  Bound: accepted moves x 2^-53 x max |E| since the last full sum.
    if (tid == 0) {
        T.n = n;
        T.resync = (double)acc_since * 1.2e-16 * e_scale > 1e-12 * fmax(1.0, fabs(E)) ? 1 : 0;
    }
    __syncthreads();
    if (T.resync) {
        const int nn = T.n;
        double acc = 0.0;
        for (int i = tid; i < nn; i += GC_THREADS) {
            const int slot = order[i];
            const double sx = px[slot], sy = py[slot], sz = pz[slot];
            int c = cell_of[slot];
            const int h2 = c % nc2; c /= nc2;
            const int h1 = c % nc1;
            const int h0 = c / nc1;
            for (int sc = 0; sc < ns; ++sc) {
                int c0 = s_st[sc][0], c1 = s_st[sc][1], c2 = s_st[sc][2];
                if (sp0 != nc0) { c0 += h0 - 2; c0 += c0 < 0 ? nc0 : 0; c0 -= c0 >= nc0 ? nc0 : 0; }
                if (sp1 != nc1) { c1 += h1 - 2; c1 += c1 < 0 ? nc1 : 0; c1 -= c1 >= nc1 ? nc1 : 0; }
                if (sp2 != nc2) { c2 += h2 - 2; c2 += c2 < 0 ? nc2 : 0; c2 -= c2 >= nc2 ? nc2 : 0; }
                const int cell = (c0 * nc1 + c1) * nc2 + c2;
                const int cnt = cell_cnt[cell];
                for (int q = 0; q < cnt; ++q) {
                    const int other = cell_slots[cell * ccap + q];
                    if (other == slot) continue;
                    const double x = px[other], y = py[other], z = pz[other];
                    const double S0 = -rint((x - sx) * i0), S1 = -rint((y - sy) * i1), S2 = -rint((z - sz) * i2);
                    const double dx = __dsub_rn(__dadd_rn(x, __dmul_rn(S0, L0)), sx);
                    const double dy = __dsub_rn(__dadd_rn(y, __dmul_rn(S1, L1)), sy);
                    const double dz = __dsub_rn(__dadd_rn(z, __dmul_rn(S2, L2)), sz);
                    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    if (r2 <= lj.rc2) acc += lj_pair(lj, r2);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double e2 = 0.0;
            for (int w = 0; w < GC_THREADS / 32; ++w) e2 += s_part[w];
            E = 0.5 * e2;                                   // every pair was met from both ends
            e_scale = fabs(E); acc_since = 0; ++resyncs;
        }
    }
    // ---------------- hand the state back ----------------
    // Mersenne-Twister streams whose current state sits in the scratch buffer go back to the canonical one
    for (int k = 0; k < nm + 2; ++k) {
        if (M.flip_mask >> k & 1) for (int i = tid; i < 624; i += GC_THREADS) mt[k * MCPYB_GCMC_MT_WORDS + i] = mt2[k * 624 + i];
    }
    if (tid < nm + 2) mt[tid * MCPYB_GCMC_MT_WORDS + 624] = (uint32_t)M.idx[tid];
    if (tid == 32) D.cyc_rows = cyc_rows;
    if (tid == 0) {
        D.e_scale = e_scale; D.acc_since_resync = acc_since; D.resyncs = resyncs;
        D.s.n_atoms = n; D.n_unstable = n_unst; D.s.energy = E; D.s.best_score = best_score; D.s.best_energy = best_E;
        D.s.best_n = best_n; D.s.last_move = cur_move; D.s.trials_done = t_done; D.trace_len = trace_len;
        D.cyc_serial = cyc_serial; D.cyc_total = clock64() - cyc_begin; D.rounds = rounds; D.cyc_pre = cyc_pre; D.cyc_wait = cyc_wait; D.cyc_gap1 = cyc_gap1;
        for (int k = 0; k < MCPYB_GCMC_MAX_MOVES; ++k)
            for (int j = 0; j < 4; ++j) D.s.pcg[k][j] = s_pcg[k][j];
        for (int k = 0; k < MCPYB_GCMC_MAX_MOVES; ++k) { D.s.attempts[k] = s_cnt[0][k]; D.s.accepted[k] = s_cnt[1][k]; D.s.failed[k] = s_cnt[2][k]; }
    }
    __syncthreads();
    if (C.in_smem) {
        uint4* dst = reinterpret_cast<uint4*>(C.blob);
        const uint4* src = reinterpret_cast<const uint4*>(gc_smem);
        for (int i = tid; i < lay.mt / 16; i += GC_THREADS) dst[i] = src[i];
        dst = reinterpret_cast<uint4*>(C.slot);
        src = reinterpret_cast<const uint4*>(gc_smem + lay.mt);
        for (int i = tid; i < (lay.bytes - lay.mt) / 16; i += GC_THREADS) dst[i] = src[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side (C ABI: include/mcpy_b200.h, "device-resident GCMC chains")
// ---------------------------------------------------------------------------------------------------------------
struct mcpyb_gcmc {
    mcpyb_engine* e = nullptr;
    int n = 0;
    std::vector<GcChain> chains;
    std::vector<GcDyn> dyn;
    GcChain* d_chains = nullptr;
    GcDyn* d_dyn = nullptr;
    long long* d_ntr = nullptr;
    bool dirty = false;                   // host copies of chains / dyn are ahead of the device's (a swap)
    long long* h_ntr = nullptr;           // page-locked staging of the trial counts and of the scalars coming back:
    GcDyn* h_dyn = nullptr;               // true asynchronous copies, one synchronisation per run
    size_t smem = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    long long launches = 0;
};

namespace {

int gc_cell_coord_host(double x, double inv, int nc) {
    double f = x * inv;
    f -= floor(f);
    int c = (int)(f * (double)nc);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

}  // namespace

extern "C" {

int mcpyb_gcmc_create(mcpyb_engine* e, int n_chains, const mcpyb_gcmc_desc* desc, int64_t trace_capacity, mcpyb_gcmc** out) {
    if (!e || !out) return MCPYB_ERR_ARG;
    *out = nullptr;
    if (n_chains < 1 || !desc) return e->fail(MCPYB_ERR_ARG, "gcmc: need at least one chain");
    cudaSetDevice(e->device);
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device), "query shared memory");
    mcpyb_gcmc* g = new mcpyb_gcmc();
    g->e = e; g->n = n_chains;
    g->chains.resize(n_chains);
    g->dyn.resize(n_chains);
    memset(g->dyn.data(), 0, sizeof(GcDyn) * n_chains);
    auto bail = [&](int code) { mcpyb_gcmc_destroy(g); return code; };
    for (int c = 0; c < n_chains; ++c) {
        GcChain& C = g->chains[c];
        memset(&C, 0, sizeof C);
        C.d = desc[c];
        const mcpyb_gcmc_desc& d = C.d;
        if (d.n_move_types < 1 || d.n_move_types > MCPYB_GCMC_MAX_MOVES) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: 1..%d moves", c, MCPYB_GCMC_MAX_MOVES));
        if (d.atom_capacity < 1 || d.atom_capacity > 65535) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: atom_capacity must be 1..65535", c));
        if (!(d.lj_sigma > 0.0) || !std::isfinite(d.lj_epsilon)) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: sigma must be > 0 and epsilon finite", c));
        const double rc = d.lj_rc > 0.0 ? d.lj_rc : 3.0 * d.lj_sigma;
        for (int m = 0; m < d.n_move_types; ++m) {
            if (d.move_kind[m] < 0 || d.move_kind[m] > 2) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: unknown kind of move %d", c, m));
            if (!(d.move_cum_weight[m] >= (m ? d.move_cum_weight[m - 1] : 0.0))) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: cumulative move weights must not decrease", c));
            if (d.move_kind[m] == MCPYB_GCMC_INSERT && (d.move_pcg[m] < 0 || d.move_pcg[m] >= MCPYB_GCMC_MAX_MOVES)) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: move %d names no PCG64 stream", c, m));
            if (d.move_kind[m] != MCPYB_GCMC_DISPLACE && !(d.move_volume[m] > 0.0)) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: move %d needs a positive volume", c, m));
        }
        C.ncells = 1;
        for (int k = 0; k < 3; ++k) {
            if (!(d.box[k] > 2.0 * rc)) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: box edge %d (%g) must exceed 2 rc (%g): the resident chain pairs an atom with ONE image of each neighbour", c, k, d.box[k], 2.0 * rc));
            C.inv_box[k] = 1.0 / d.box[k];
            int nc = (int)floor(d.box[k] / (0.5 * rc * (1.0 + 1e-6)));
            if (nc > 40) nc = 40;
            C.ncell[k] = nc;
            C.span[k] = nc >= 5 ? 5 : nc;
            C.ncells *= nc;
        }
        C.lj.eps4 = 4.0 * d.lj_epsilon; C.lj.sigma2 = d.lj_sigma * d.lj_sigma; C.lj.rc2 = rc * rc; C.lj.rc = rc;
        C.lj.e0 = 4.0 * d.lj_epsilon * (pow(d.lj_sigma / rc, 12) - pow(d.lj_sigma / rc, 6));
        C.acap = d.atom_capacity;
        C.ccap = d.cell_capacity > 0 ? d.cell_capacity : 16;
        if (C.ccap > 4096) return bail(e->fail(MCPYB_ERR_ARG, "gcmc chain %d: cell_capacity too large", c));
        C.nmt = d.n_move_types + 2;
        C.lay = gc_layout(C.acap, C.ncells, C.ccap, C.nmt);
        C.in_smem = (size_t)C.lay.bytes + 4096 <= (size_t)optin;
        if (C.in_smem && (size_t)C.lay.bytes > g->smem) g->smem = C.lay.bytes;
        if (cudaMalloc(&C.blob, C.lay.mt) != cudaSuccess || cudaMalloc(&C.slot, C.lay.bytes - C.lay.mt) != cudaSuccess)
            return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: cudaMalloc state"));
        if (cudaMalloc(&C.best_pos, sizeof(double) * 3 * C.acap) != cudaSuccess) return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: cudaMalloc best"));
        C.trace_cap = trace_capacity > 0 ? trace_capacity : 0;
        if (C.trace_cap && cudaMalloc(&C.trace, sizeof(mcpyb_gcmc_trial) * C.trace_cap) != cudaSuccess) return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: cudaMalloc trace"));
        g->dyn[c].s.best_n = -1;
    }
    if (cudaMalloc(&g->d_chains, sizeof(GcChain) * n_chains) != cudaSuccess || cudaMalloc(&g->d_dyn, sizeof(GcDyn) * n_chains) != cudaSuccess ||
        cudaMalloc(&g->d_ntr, sizeof(long long) * n_chains) != cudaSuccess)
        return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: cudaMalloc descriptors"));
    if (cudaMemcpy(g->d_chains, g->chains.data(), sizeof(GcChain) * n_chains, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(g->d_dyn, g->dyn.data(), sizeof(GcDyn) * n_chains, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: upload descriptors"));
    if (g->smem && cudaFuncSetAttribute(gcmc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem) != cudaSuccess)
        return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: shared memory opt-in of %zu bytes refused", g->smem));
    cudaEventCreate(&g->ev0); cudaEventCreate(&g->ev1);
    if (cudaMallocHost(&g->h_ntr, sizeof(long long) * n_chains) != cudaSuccess || cudaMallocHost(&g->h_dyn, sizeof(GcDyn) * n_chains) != cudaSuccess)
        return bail(e->fail(MCPYB_ERR_CUDA, "gcmc: cudaMallocHost"));
    *out = g;
    return MCPYB_OK;
}

void mcpyb_gcmc_destroy(mcpyb_gcmc* g) {
    if (!g) return;
    cudaSetDevice(g->e->device);
    cudaStreamSynchronize(g->e->stream);
    for (auto& C : g->chains) { if (C.blob) cudaFree(C.blob); if (C.slot) cudaFree(C.slot); if (C.best_pos) cudaFree(C.best_pos); if (C.trace) cudaFree(C.trace); }
    if (g->d_chains) cudaFree(g->d_chains);
    if (g->d_dyn) cudaFree(g->d_dyn);
    if (g->d_ntr) cudaFree(g->d_ntr);
    if (g->h_ntr) cudaFreeHost(g->h_ntr);
    if (g->h_dyn) cudaFreeHost(g->h_dyn);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    delete g;
}

int mcpyb_gcmc_set_state(mcpyb_gcmc* g, int chain, const double* positions, const mcpyb_gcmc_scalars* s, const uint32_t* mt) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (chain < 0 || chain >= g->n || !s || !mt) return e->fail(MCPYB_ERR_ARG, "gcmc set_state: bad argument");
    const GcChain& C = g->chains[chain];
    const int n = s->n_atoms;
    if (n < 0 || n >= C.acap) return e->fail(MCPYB_ERR_CAPACITY, "gcmc chain %d: %d atoms do not fit a capacity of %d (one slot stays free)", chain, n, C.acap);
    if (n > 0 && !positions) return e->fail(MCPYB_ERR_ARG, "gcmc set_state: positions missing");
    cudaSetDevice(e->device);
    std::vector<unsigned char> blob(C.lay.bytes, 0);
    double* px = reinterpret_cast<double*>(blob.data() + C.lay.px);
    double* py = reinterpret_cast<double*>(blob.data() + C.lay.py);
    double* pz = reinterpret_cast<double*>(blob.data() + C.lay.pz);
    uint32_t* mts = reinterpret_cast<uint32_t*>(blob.data() + C.lay.mt);
    unsigned short* order = reinterpret_cast<unsigned short*>(blob.data() + C.lay.order);
    unsigned short* cell_of = reinterpret_cast<unsigned short*>(blob.data() + C.lay.cell_of);
    unsigned short* unstable = reinterpret_cast<unsigned short*>(blob.data() + C.lay.unstable);
    unsigned char* uflag = blob.data() + C.lay.uflag;
    unsigned short* cell_cnt = reinterpret_cast<unsigned short*>(blob.data() + C.lay.cell_cnt);
    unsigned short* cell_slots = reinterpret_cast<unsigned short*>(blob.data() + C.lay.cell_slots);
    for (int i = 0; i < C.acap; ++i) order[i] = (unsigned short)i;
    for (int i = 0; i < n; ++i) {
        const double x = positions[3 * i], y = positions[3 * i + 1], z = positions[3 * i + 2];
        if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) return e->fail(MCPYB_ERR_ARG, "gcmc chain %d: atom %d has a non-finite position", chain, i);
        px[i] = x; py[i] = y; pz[i] = z;
        const int c = (gc_cell_coord_host(x, C.inv_box[0], C.ncell[0]) * C.ncell[1] + gc_cell_coord_host(y, C.inv_box[1], C.ncell[1])) * C.ncell[2]
                      + gc_cell_coord_host(z, C.inv_box[2], C.ncell[2]);
        if (cell_cnt[c] + 1 >= C.ccap) return e->fail(MCPYB_ERR_CAPACITY, "gcmc chain %d: more than %d atoms in one cell of the chain's cell list", chain, C.ccap - 2);
        cell_slots[c * C.ccap + cell_cnt[c]] = (unsigned short)i;
        cell_cnt[c] += 1;
        cell_of[i] = (unsigned short)c;
        unstable[i] = (unsigned short)i;
        uflag[i] = 1;
    }
    for (int k = 0; k < C.nmt; ++k) {
        const uint32_t* src = mt + (size_t)k * MCPYB_GCMC_MT_WORDS;
        if (src[624] > 624) return e->fail(MCPYB_ERR_ARG, "gcmc chain %d: Mersenne-Twister stream %d has position %u (> 624)", chain, k, src[624]);
        memcpy(mts + (size_t)k * MCPYB_GCMC_MT_WORDS, src, sizeof(uint32_t) * MCPYB_GCMC_MT_WORDS);
    }
    GcDyn& D = g->dyn[chain];
    D.s = *s;
    D.s.status = MCPYB_GCMC_DONE;
    D.trace_len = 0;
    D.n_unstable = n;
    D.e_scale = fabs(s->energy);             // the caller's energy is a full sum
    D.acc_since_resync = 0;
    if (g->dirty) {                          // swapped chains first: the descriptors below must not be overwritten later
        CK(cudaMemcpyAsync(g->d_chains, g->chains.data(), sizeof(GcChain) * g->n, cudaMemcpyHostToDevice, e->stream), "gcmc upload descriptors");
        CK(cudaMemcpyAsync(g->d_dyn, g->dyn.data(), sizeof(GcDyn) * g->n, cudaMemcpyHostToDevice, e->stream), "gcmc upload scalars");
        g->dirty = false;
    }
    CK(cudaMemcpyAsync(C.blob, blob.data(), C.lay.mt, cudaMemcpyHostToDevice, e->stream), "gcmc upload state");
    CK(cudaMemcpyAsync(C.slot, blob.data() + C.lay.mt, C.lay.bytes - C.lay.mt, cudaMemcpyHostToDevice, e->stream), "gcmc upload state");
    CK(cudaMemcpyAsync(g->d_dyn + chain, &D, sizeof(GcDyn), cudaMemcpyHostToDevice, e->stream), "gcmc upload scalars");
    CK(cudaStreamSynchronize(e->stream), "gcmc set_state sync");
    return MCPYB_OK;
}

int mcpyb_gcmc_get_state(mcpyb_gcmc* g, int chain, double* positions_out, mcpyb_gcmc_scalars* s_out, uint32_t* mt_out, double* best_positions_out) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (chain < 0 || chain >= g->n) return e->fail(MCPYB_ERR_ARG, "gcmc get_state: no chain %d", chain);
    const GcChain& C = g->chains[chain];
    cudaSetDevice(e->device);
    std::vector<unsigned char> blob(C.lay.bytes);
    GcDyn& D = g->dyn[chain];
    CK(cudaMemcpyAsync(blob.data(), C.blob, C.lay.mt, cudaMemcpyDeviceToHost, e->stream), "gcmc download state");
    CK(cudaMemcpyAsync(blob.data() + C.lay.mt, C.slot, C.lay.bytes - C.lay.mt, cudaMemcpyDeviceToHost, e->stream), "gcmc download state");
    if (!g->dirty) CK(cudaMemcpyAsync(&D, g->d_dyn + chain, sizeof(GcDyn), cudaMemcpyDeviceToHost, e->stream), "gcmc download scalars");
    CK(cudaStreamSynchronize(e->stream), "gcmc get_state sync");
    if (s_out) *s_out = D.s;
    const int n = D.s.n_atoms;
    if (positions_out) {
        const double* px = reinterpret_cast<const double*>(blob.data() + C.lay.px);
        const double* py = reinterpret_cast<const double*>(blob.data() + C.lay.py);
        const double* pz = reinterpret_cast<const double*>(blob.data() + C.lay.pz);
        const unsigned short* order = reinterpret_cast<const unsigned short*>(blob.data() + C.lay.order);
        for (int i = 0; i < n; ++i) {
            const int s = order[i];
            positions_out[3 * i] = px[s]; positions_out[3 * i + 1] = py[s]; positions_out[3 * i + 2] = pz[s];
        }
    }
    if (mt_out) memcpy(mt_out, blob.data() + C.lay.mt, sizeof(uint32_t) * MCPYB_GCMC_MT_WORDS * C.nmt);
    if (best_positions_out && D.s.best_n > 0) {
        CK(cudaMemcpyAsync(best_positions_out, C.best_pos, sizeof(double) * 3 * D.s.best_n, cudaMemcpyDeviceToHost, e->stream), "gcmc download best");
        CK(cudaStreamSynchronize(e->stream), "gcmc get_state sync");
    }
    return MCPYB_OK;
}

int mcpyb_gcmc_run(mcpyb_gcmc* g, const int64_t* n_trials, int32_t* status_out) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (!n_trials) return e->fail(MCPYB_ERR_ARG, "gcmc run: n_trials missing");
    for (int c = 0; c < g->n; ++c) if (n_trials[c] < 0) return e->fail(MCPYB_ERR_ARG, "gcmc run: negative trial count");
    cudaSetDevice(e->device);
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t");
    if (g->dirty) {
        CK(cudaMemcpyAsync(g->d_chains, g->chains.data(), sizeof(GcChain) * g->n, cudaMemcpyHostToDevice, e->stream), "gcmc upload descriptors");
        CK(cudaMemcpyAsync(g->d_dyn, g->dyn.data(), sizeof(GcDyn) * g->n, cudaMemcpyHostToDevice, e->stream), "gcmc upload scalars");
        g->dirty = false;
    }
    memcpy(g->h_ntr, n_trials, sizeof(long long) * g->n);
    CK(cudaMemcpyAsync(g->d_ntr, g->h_ntr, sizeof(long long) * g->n, cudaMemcpyHostToDevice, e->stream), "gcmc upload trial counts");
    CK(cudaEventRecord(g->ev0, e->stream), "event");
    gcmc_chain_kernel<<<g->n, GC_THREADS, g->smem, e->stream>>>(g->d_chains, g->d_dyn, g->d_ntr);
    LAUNCH_CHECK("gcmc_chain_kernel");
    CK(cudaEventRecord(g->ev1, e->stream), "event");
    CK(cudaMemcpyAsync(g->h_dyn, g->d_dyn, sizeof(GcDyn) * g->n, cudaMemcpyDeviceToHost, e->stream), "gcmc download scalars");
    CK(cudaStreamSynchronize(e->stream), "gcmc run");
    memcpy(g->dyn.data(), g->h_dyn, sizeof(GcDyn) * g->n);
    cudaEventElapsedTime(&g->last_ms, g->ev0, g->ev1);
    g->launches++;
    if (status_out) for (int c = 0; c < g->n; ++c) status_out[c] = g->dyn[c].s.status;
    return MCPYB_OK;
}

int mcpyb_gcmc_get_trace(mcpyb_gcmc* g, int chain, mcpyb_gcmc_trial* out, int64_t capacity, int64_t* n_out) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (chain < 0 || chain >= g->n || !n_out) return e->fail(MCPYB_ERR_ARG, "gcmc get_trace: bad argument");
    const GcChain& C = g->chains[chain];
    long long len = g->dyn[chain].trace_len;
    if (len > capacity) len = capacity;
    *n_out = len;
    if (len > 0 && out) {
        cudaSetDevice(e->device);
        CK(cudaMemcpyAsync(out, C.trace, sizeof(mcpyb_gcmc_trial) * len, cudaMemcpyDeviceToHost, e->stream), "gcmc download trace");
        CK(cudaStreamSynchronize(e->stream), "gcmc get_trace sync");
    }
    return MCPYB_OK;
}

int mcpyb_gcmc_get_scalars(mcpyb_gcmc* g, int chain, mcpyb_gcmc_scalars* s_out) {
    if (!g) return MCPYB_ERR_ARG;
    if (chain < 0 || chain >= g->n || !s_out) return g->e->fail(MCPYB_ERR_ARG, "gcmc get_scalars: bad argument");
    *s_out = g->dyn[chain].s;                 // the host copy: mcpyb_gcmc_run downloads every chain's scalars
    return MCPYB_OK;
}

int mcpyb_gcmc_swap_configurations(mcpyb_gcmc* g, int i, int j) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (i < 0 || j < 0 || i >= g->n || j >= g->n) return e->fail(MCPYB_ERR_ARG, "gcmc swap: no such chain");
    if (i == j) return MCPYB_OK;
    GcChain& A = g->chains[i];
    GcChain& B = g->chains[j];
    if (memcmp(&A.lay, &B.lay, sizeof(GcLayout)) != 0 || A.acap != B.acap || A.ccap != B.ccap || A.ncells != B.ncells ||
        memcmp(A.ncell, B.ncell, sizeof A.ncell) != 0 || memcmp(A.d.box, B.d.box, sizeof A.d.box) != 0)
        return e->fail(MCPYB_ERR_CAPACITY, "gcmc swap: chains %d and %d have different boxes or storage layouts", i, j);
    // the configuration part of the state changes owner (positions, atom order, cell lists, wrap bookkeeping) together
    // with the energy and the particle number; the slot part (the Mersenne-Twister streams) stays.  Host copies only:
    // the next mcpyb_gcmc_run / set_state uploads the descriptors and scalars of all chains in two copies.
    std::swap(A.blob, B.blob);
    GcDyn& DA = g->dyn[i];
    GcDyn& DB = g->dyn[j];
    std::swap(DA.s.energy, DB.s.energy);
    std::swap(DA.s.n_atoms, DB.s.n_atoms);
    std::swap(DA.n_unstable, DB.n_unstable);
    std::swap(DA.e_scale, DB.e_scale);
    std::swap(DA.acc_since_resync, DB.acc_since_resync);
    g->dirty = true;
    return MCPYB_OK;
}

int mcpyb_gcmc_get_block_scalars(mcpyb_gcmc* g, double* energy, int32_t* n_atoms, int64_t* trials_done, double* best2) {
    if (!g) return MCPYB_ERR_ARG;
    if (!energy || !n_atoms || !trials_done || !best2) return g->e->fail(MCPYB_ERR_ARG, "gcmc get_block_scalars: bad argument");
    for (int c = 0; c < g->n; ++c) {
        const mcpyb_gcmc_scalars& s = g->dyn[c].s;
        energy[c] = s.energy; n_atoms[c] = s.n_atoms; trials_done[c] = s.trials_done;
        best2[2 * c] = s.best_score; best2[2 * c + 1] = s.best_energy;
    }
    return MCPYB_OK;
}

int mcpyb_gcmc_config_ptr(mcpyb_gcmc* g, int chain, void** dev_ptr, int64_t* bytes, double* scalars5) {
    if (!g) return MCPYB_ERR_ARG;
    if (chain < 0 || chain >= g->n || !dev_ptr || !bytes) return g->e->fail(MCPYB_ERR_ARG, "gcmc config_ptr: bad argument");
    const GcChain& C = g->chains[chain];
    *dev_ptr = C.blob;
    *bytes = C.lay.mt;
    if (scalars5) {
        const GcDyn& D = g->dyn[chain];
        scalars5[0] = D.s.energy; scalars5[1] = (double)D.s.n_atoms; scalars5[2] = (double)D.n_unstable;
        scalars5[3] = D.e_scale; scalars5[4] = (double)D.acc_since_resync;
    }
    return MCPYB_OK;
}

int mcpyb_gcmc_set_config_scalars(mcpyb_gcmc* g, int chain, const double* scalars5) {
    if (!g) return MCPYB_ERR_ARG;
    mcpyb_engine* e = g->e;
    if (chain < 0 || chain >= g->n || !scalars5) return e->fail(MCPYB_ERR_ARG, "gcmc set_config_scalars: bad argument");
    const GcChain& C = g->chains[chain];
    const int n = (int)scalars5[1], nu = (int)scalars5[2];
    if (n < 0 || n >= C.acap || nu < 0 || nu > n) return e->fail(MCPYB_ERR_ARG, "gcmc set_config_scalars: %d atoms (%d unsettled) do not fit chain %d", n, nu, chain);
    GcDyn& D = g->dyn[chain];
    D.s.energy = scalars5[0]; D.s.n_atoms = n; D.n_unstable = nu;
    D.e_scale = scalars5[3]; D.acc_since_resync = (long long)scalars5[4];
    g->dirty = true;
    return MCPYB_OK;
}

int mcpyb_gcmc_info(mcpyb_gcmc* g, int chain, int32_t* out6) {
    if (!g) return MCPYB_ERR_ARG;
    if (chain < 0 || chain >= g->n || !out6) return g->e->fail(MCPYB_ERR_ARG, "gcmc info: bad argument");
    const GcChain& C = g->chains[chain];
    out6[0] = C.ncell[0]; out6[1] = C.ncell[1]; out6[2] = C.ncell[2]; out6[3] = C.ccap; out6[4] = C.lay.bytes; out6[5] = C.in_smem;
    return MCPYB_OK;
}

int mcpyb_gcmc_cycles(mcpyb_gcmc* g, int chain, int64_t* out3 /* [8] */) {
    if (!g) return MCPYB_ERR_ARG;
    if (chain < 0 || chain >= g->n || !out3) return g->e->fail(MCPYB_ERR_ARG, "gcmc cycles: bad argument");
    out3[0] = g->dyn[chain].cyc_serial; out3[1] = g->dyn[chain].cyc_total; out3[2] = g->dyn[chain].rounds;
    out3[3] = g->dyn[chain].resyncs; out3[4] = g->dyn[chain].cyc_pre;
    out3[5] = g->dyn[chain].cyc_rows; out3[6] = g->dyn[chain].cyc_wait; out3[7] = g->dyn[chain].cyc_gap1;
    return MCPYB_OK;
}

int mcpyb_gcmc_last_run_ms(mcpyb_gcmc* g, double* ms_out) {
    if (!g || !ms_out) return MCPYB_ERR_ARG;
    *ms_out = (double)g->last_ms;
    return MCPYB_OK;
}

}  // extern "C"
