// Host side of the engine and the C ABI declared in include/mcpy_b200.h.
// One translation unit: every kernel is included here so the library is a single
// nvcc -shared build for sm_100a (see mcpy_b200/build.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "../../include/mcpy_b200.h"
#include "common.cuh"
#include "cell_list.cuh"
#include "stencil.cuh"
#include "lj_kernels.cuh"
#include "site_kernels.cuh"
#include "rows_block.cuh"
#include "rows_warp.cuh"
#include "chain.cuh"
#include "emt_rows.cuh"
#include "emt_kernels.cuh"
#include "free_volume.cuh"
#include "pcg64.cuh"

// FP64 pipe microbenchmark: 8 independent DFMA chains per thread (roofline denominator
// for the pair kernels; MEASURED_PEAKS.json only has HBM and bf16 numbers).
__global__ void __launch_bounds__(256)
fp64_fma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t want = cap ? cap : 256;
        while (want < n) want *= 2;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    unsigned char* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t want = cap ? cap : 4096;
        while (want < n) want *= 2;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost((void**)&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A few helper threads for the large host-side passes of the batched calls (user arrays -> pinned staging
// and back; the row-by-row comparison of 64 incoming configurations with their resident bases): one core
// moves ~20 GB/s, which made such a pass the longest single item of a call.  Workers poll for the next job for a
// short while after each one, then sleep on a condition variable.
class CopyPool {
public:
    explicit CopyPool(int nworkers, int spin_us) : n_(nworkers), spin_us_(spin_us) {
        for (int k = 0; k < nworkers; ++k) threads_.emplace_back([this, k]() { loop(k); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_.store(true); gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int workers() const { return n_; }
    // memcpy(dst, src, n) split over the workers and the caller
    void copy(void* dst, const void* src, size_t n) {
        const size_t min_part = 128u << 10;
        int parts = (int)(n / min_part);
        if (parts > n_ + 1) parts = n_ + 1;
        if (parts <= 1) { memcpy(dst, src, n); return; }
        struct Ctx { unsigned char* dst; const unsigned char* src; size_t n, each; int parts; };
        Ctx c{(unsigned char*)dst, (const unsigned char*)src, n, ((n / parts) + 63) & ~(size_t)63, parts};
        parallel_for(parts, [](void* v, int p) {
            Ctx& x = *(Ctx*)v;
            const size_t beg = x.each * p;
            if (beg < x.n) memcpy(x.dst + beg, x.src + beg, (p == x.parts - 1 || beg + x.each > x.n) ? x.n - beg : x.each);
        }, &c);
    }
    // fn(ctx, i) for i in [0, n): index i always runs on the same thread (i mod parts; the caller takes part 0), so
    // whatever a thread touched for index i in the last call is still in its cache
    void parallel_for(int n, void (*fn)(void*, int), void* ctx) {
        int parts = n_ + 1;
        if (parts > n) parts = n;
        if (parts <= 1) { for (int i = 0; i < n; ++i) fn(ctx, i); return; }
        fn_ = fn; ctx_ = ctx; count_ = n; parts_ = parts;
        pending_.store(n_, std::memory_order_relaxed);
        bool wake;
        { std::lock_guard<std::mutex> lk(m_); gen_.fetch_add(1, std::memory_order_release); wake = sleepers_ > 0; }
        if (wake) cv_.notify_all();
        for (int i = 0; i < n; i += parts) fn(ctx, i);
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    }
private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // Every worker acknowledges every generation (pending_ counts all of them), so the job fields are never rewritten
    // while a worker may still read them.  After a job a worker polls for the next one for spin_us_ before it goes to
    // sleep on the condition variable: a Monte-Carlo loop calls every ~100 us, and a futex wake-up costs 10-20 us.
    void loop(int k) {
        unsigned long long seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned it = 1; gen_.load(std::memory_order_acquire) == seen; ++it) {
                cpu_relax();
                if ((it & 0xFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) {
                    std::unique_lock<std::mutex> lk(m_);
                    ++sleepers_;
                    cv_.wait(lk, [&]() { return gen_.load(std::memory_order_acquire) != seen; });
                    --sleepers_;
                    break;
                }
            }
            if (stop_.load()) return;
            seen = gen_.load(std::memory_order_acquire);
            const int first = k + 1;
            if (first < parts_) for (int i = first; i < count_; i += parts_) fn_(ctx_, i);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    const int n_, spin_us_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<unsigned long long> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> stop_{false};
    int sleepers_ = 0;
    void (*fn_)(void*, int) = nullptr;
    void* ctx_ = nullptr;
    int count_ = 0, parts_ = 1;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- EMT constants (ase.calculators.emt) -----------------
struct EmtRow { int Z; double E0, s0, V0, eta2, kappa, lambda, n0; };
const EmtRow kEmtRows[] = {
    {1,  -3.21, 1.31, 0.132, 2.652, 2.790, 3.892, 0.00547},   // H
    {6,  -3.50, 1.81, 0.332, 1.652, 2.790, 1.892, 0.01322},   // C
    {7,  -5.10, 1.88, 0.132, 1.652, 2.790, 1.892, 0.01222},   // N
    {8,  -4.60, 1.95, 0.332, 1.652, 2.790, 1.892, 0.00850},   // O
    {13, -3.28, 3.00, 1.493, 1.240, 2.000, 1.169, 0.00700},   // Al
    {28, -4.44, 2.60, 3.673, 1.669, 2.757, 1.948, 0.01030},   // Ni
    {29, -3.51, 2.67, 2.476, 1.652, 2.740, 1.906, 0.00910},   // Cu
    {46, -3.90, 2.87, 2.773, 1.818, 3.107, 2.155, 0.00688},   // Pd
    {47, -2.96, 3.01, 2.132, 1.652, 2.790, 1.892, 0.00547},   // Ag
    {78, -5.85, 2.90, 4.067, 1.812, 3.145, 2.192, 0.00802},   // Pt
    {79, -3.80, 3.00, 2.321, 1.674, 2.873, 2.182, 0.00703},   // Au
};
const double kBohr = 0.52917721067;
const double kEmtBeta = 1.809;

}  // namespace

struct mcpyb_engine {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    // potential
    int pot_kind = POT_NONE;
    int subdiv = 1;                       // cells per cutoff length of the RESIDENT cell list
    int subdiv_trials = 1;                // ... wanted when trial moves will be scored against it
    int subdiv_energy = 1;                // ... wanted for upload + total energies in one call
    int rb_slices = 2;                    // candidate-list slices per work item of the CTA row-sum kernel
    // tuning switches (environment, read once at creation; the defaults are the measured best):
    //   MCPYB_PDL=0        ordinary stream-ordered launches instead of programmatic dependent launch
    //   MCPYB_RB_SLICES=n  candidate-list slices per work item (1..8)
    //   MCPYB_EXPAND=0     no per-configuration stencil lists (every work item stages from the cell list)
    //   MCPYB_ROWS=block   the CTA-per-cell row-sum kernel of round 1 instead of the warp-per-item one
    bool use_pdl = true;
    bool use_expand = true;
    bool use_warp_rows = true;            // MCPYB_ROWS=block: the round-1 CTA-per-cell row-sum kernel (A/B runs only)
    int rw_minb = RW_MIN_BLOCKS;          // MCPYB_RW_MINB=4|5|6: resident CTAs per SM the warp row-sum kernel is built for
    int l2_persist_mb = 0;                // MCPYB_L2PERSIST=<MB>: keep the stencil lists in a persisting L2 window
    bool fp32_pairs = false;              // mcpyb_set_precision: fp32 pair math in the batched LJ row sums
    bool use_direct = true;               // MCPYB_DIRECT=0: dense LJ batches keep the scan + scatter pipeline
    int slot_cap_override = 0;            // MCPYB_SLOT_CAP: slots per bin of the direct placement (tests: force overflow)
    bool hist_dirty = false;              // the trial histogram holds the counts of a direct-placement launch
    double cells_est = 0.0;               // cells of the resident batch's grids, summed over replicas (periodic boxes)
    DevBuf<int> tw_ovf;                   // [0] sites that found their bin full, [2] bin claim counter (direct placement)
    int er_minb = 6;                      // MCPYB_ER_MINB: resident CTAs per SM the list-driven EMT trial kernel is compiled for
    bool use_emt_lists = true;            // MCPYB_EMT_LISTS=0: the first-generation EMT trial kernel for every batch
    // stencil lists of the resident batch (rows_block.cuh): built when a SECOND trial batch is scored against
    // the same resident configuration -- one batch per upload (the end-to-end pattern) does not amortise them
    bool exp_valid = false;
    int exp_batches = 0;                  // batched trial launches since the last upload
    DevBuf<double4> exp_q;
    DevBuf<int2> exp_info, exp_head;
    DevBuf<unsigned long long> exp_cursor;
    double rcut = 0.0;
    LJParams lj{};
    EmtTable emt_host{};
    DevBuf<EmtTable> emt_dev;
    int z_to_slot_host[120];
    DevBuf<int> z_to_slot;

    // resident batch
    int R = 0;
    int total_atoms = 0;
    int max_n = 0;                        // largest configuration of the resident batch
    int max_cells = 0;
    bool have_batch = false;
    bool emt_state_valid = false;     // sigma1 / e_atom describe the resident batch
    bool status_checked = false;      // the build status of the resident batch has been read back
    std::vector<int> atom_off_host;
    DevBuf<GridDesc> grid;
    DevBuf<double4> upos, spos, stmp;
    DevBuf<float4> fpos;
    DevBuf<int> cid, cell_start;
    const int* d_atom_off = nullptr;      // points into stage_dev (header of the last upload)
    DevBuf<CellGeom> geom;
    DevBuf<double> bbox;                  // cluster build scratch
    DevBuf<double> e_atom, sigma1, e_emb, e_part, deds, forces;
    DevBuf<int> cnt_part;
    DevBuf<double> E;                     // one block: E[R] | pairs[R] (int64) | status[R] (int32)
    DevBuf<int> pair_count;
    long long* pairs_ptr() { return (long long*)(E.p + R); }
    int* status_ptr() { return (int*)(E.p + 2 * (size_t)R); }
    DevBuf<unsigned char> stage_dev;      // H2D landing zone: header | positions | numbers
    PinBuf stage_pin, out_pin;

    // incremental drop-in path (mcpyb_tracked_energy_host): host mirror of the resident base
    bool trk_valid = false;
    std::vector<double> trk_pos;
    std::vector<int32_t> trk_Z;
    double trk_cell[9];
    uint8_t trk_pbc[3];
    double trk_E = 0.0;
    long long trk_full = 0, trk_incremental = 0;
    bool trkb_timing = false;             // MCPYB_TRK_TIMING=1: host-side split of the batched tracked call, printed at destroy
    double trkb_t_compare = 0.0, trkb_t_build = 0.0, trkb_t_trial = 0.0, trkb_t_stage = 0.0, trkb_t_launch = 0.0, trkb_t_wait = 0.0;
    long long trkb_calls = 0;
    // ... and of a resident BATCH of bases (mcpyb_tracked_energies_host): slot r holds rows
    // [trkb_off[r], trkb_off[r+1]) of trkb_pos / trkb_Z
    bool trkb_valid = false;
    std::vector<double> trkb_pos, trkb_cell, trkb_E;
    std::vector<int32_t> trkb_Z;
    std::vector<int64_t> trkb_off;
    std::vector<uint8_t> trkb_pbc;

    // instrumentation: CUDA events between the launches of the trial pipeline (mcpyb_stage_timing)
    bool stage_timing = false;
    cudaEvent_t stage_ev[6] = {};
    int stage_n = 0;                      // events recorded by the last launch_trials
    double stage_ms[5] = {0, 0, 0, 0, 0};
    long long stage_calls = 0;

    // trial staging
    CopyPool* copy_pool = nullptr;        // created on the first large batched call
    bool trial_staging_busy = false;      // a host trial call returned early with copies in flight
    cudaEvent_t trial_ev = nullptr;
    unsigned int tiny_seq = 0;            // completion word of the single-launch trial kernels (wait_flag)
    // large trial batches travel on their own stream, so that their H2D copy overlaps whatever the engine's
    // stream is still doing (the upload + cell-list build of the configuration they are scored against)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev = nullptr;
    DevBuf<unsigned char> trial_dev;
    DevBuf<unsigned char> mask_dev;          // K10 region mask: positions | numbers | species | mask, one block
    PinBuf mask_pin;
    PinBuf trial_pin, trial_out_pin;
    DevBuf<double> dE;
    DevBuf<int> tpairs;
    // second-generation trial pipeline (site_kernels.cuh)
    DevBuf<int> tw_int;                   // site_trial | sorted | site_cnt | totals
    DevBuf<int> tw_bins;                  // bin_count | item_start
    DevBuf<int4> tw_a;
    DevBuf<int> tw_hist;
    DevBuf<unsigned long long> tw_err;    // [0] armed error word, [1] its value after the last launch
    int tw_stamp = 0;
    DevBuf<double4> tw_pos, tw_spos;
    DevBuf<int4> tw_meta;
    DevBuf<int2> tw_items;
    DevBuf<int4> tw_items4;
    DevBuf<unsigned long long> tw_chain;
    DevBuf<int> emt_redo;                 // [1 + T] trials handed back by the list-driven EMT trial kernel
    DevBuf<int> self_trials;              // rep | old_off | old_idx | new_off of the all-atoms trial batch (K2 via rows)
    DevBuf<double> tw_E;

    // resident chain (chain.cuh): mailbox in pinned host memory, private work arrays, its own stream
    bool use_chain = true;                // MCPYB_CHAIN=0: one launch per call, as in round 1
    cudaStream_t chain_stream = nullptr;
    cudaEvent_t chain_ev = nullptr;
    ChainMailbox* chain_mb = nullptr;
    unsigned int chain_seq = 0;
    bool chain_launched = false;          // a CTA has been launched and not been waited for since
    long long chain_launches = 0, chain_trials = 0;

    // neighbour list scratch
    DevBuf<int> nl_i, nl_j, nl_S;
    DevBuf<double> nl_r2;
    DevBuf<unsigned long long> nl_cursor;

    // free volume
    DevBuf<FvGrid> fv_grid;
    DevBuf<double4> fv_spheres;
    DevBuf<int> fv_cell_start, fv_cursor, fv_cid;
    DevBuf<double> fv_points, fv_pos, fv_radii, fv_small;   // fv_small: offset(3) + shifts(81)
    DevBuf<unsigned long long> fv_counts;
    DevBuf<unsigned char> fv_mask;
    PinBuf fv_pin, fv_atoms_pin, fv_out_pin;
    DevBuf<int> pcg_count;
    DevBuf<long long> pcg_prefix;
    DevBuf<SphereCtl> pcg_ctl;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        return fail(MCPYB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    BatchView view() {
        BatchView v;
        v.grid = grid.p; v.upos = upos.p; v.spos = spos.p; v.stmp = stmp.p; v.fpos = fpos.p; v.cell_start = cell_start.p;
        v.cid = cid.p; v.R = R; v.max_cells = max_cells; v.cs_stride = 2 * (max_cells + 1);
        return v;
    }
};

#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return e->cuda_fail(_e, what); } while (0)
#define LAUNCH_CHECK(what) do { e->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return e->cuda_fail(_e, what); } while (0)

namespace {

// The resident chain grid reads the batch arrays, the stencil lists and its private work arrays: it has to be gone
// before any of them is rebuilt or freed.
int chain_stop(mcpyb_engine* e) {
    if (!e->chain_mb || !e->chain_launched) return MCPYB_OK;
    e->chain_mb->stop = 1;
    CK(cudaStreamSynchronize(e->chain_stream), "sync resident chain");
    e->chain_mb->stop = 0;
    e->chain_mb->alive = 0;
    e->chain_launched = false;
    return MCPYB_OK;
}

// Lattice bookkeeping done once per upload on the host (9 numbers per replica).
int make_geom(mcpyb_engine* e, const double* cell, const uint8_t* pbc, CellGeom* g) {
    double c[3][3];
    for (int i = 0; i < 9; ++i) c[i / 3][i % 3] = cell[i];
    bool zero[3];
    for (int d = 0; d < 3; ++d) {
        zero[d] = (c[d][0] == 0.0 && c[d][1] == 0.0 && c[d][2] == 0.0);
        if (zero[d] && pbc[d]) return e->fail(MCPYB_ERR_GEOMETRY, "periodic axis %d has a zero lattice vector", d);
    }
    auto cross = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto norm = [](const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    // complete missing rows with unit vectors orthogonal to the others
    int nzero = zero[0] + zero[1] + zero[2];
    if (nzero == 3) {
        for (int d = 0; d < 3; ++d) c[d][d] = 1.0;
    } else if (nzero == 2) {
        int k = !zero[0] ? 0 : (!zero[1] ? 1 : 2);
        int a = (k + 1) % 3, b = (k + 2) % 3;
        int imin = 0;
        for (int i = 1; i < 3; ++i) if (fabs(c[k][i]) < fabs(c[k][imin])) imin = i;
        double t[3] = {0, 0, 0};
        t[imin] = 1.0;
        double v[3];
        cross(c[k], t, v);
        double nv = norm(v);
        for (int i = 0; i < 3; ++i) c[a][i] = v[i] / nv;
        cross(c[k], c[a], v);
        nv = norm(v);
        for (int i = 0; i < 3; ++i) c[b][i] = v[i] / nv;
    } else if (nzero == 1) {
        int k = zero[0] ? 0 : (zero[1] ? 1 : 2);
        double v[3];
        cross(c[(k + 1) % 3], c[(k + 2) % 3], v);
        double nv = norm(v);
        if (nv == 0.0) return e->fail(MCPYB_ERR_GEOMETRY, "singular cell");
        for (int i = 0; i < 3; ++i) c[k][i] = v[i] / nv;
    }
    const double det = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1])
                     - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0])
                     + c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
    if (!(fabs(det) > 0.0) || !isfinite(det)) return e->fail(MCPYB_ERR_GEOMETRY, "singular cell");
    // inverse: inv[j][d] with frac_d = sum_j p_j inv[j][d]  (p = frac @ cell)
    double inv[3][3];
    inv[0][0] = (c[1][1] * c[2][2] - c[1][2] * c[2][1]) / det;
    inv[0][1] = (c[0][2] * c[2][1] - c[0][1] * c[2][2]) / det;
    inv[0][2] = (c[0][1] * c[1][2] - c[0][2] * c[1][1]) / det;
    inv[1][0] = (c[1][2] * c[2][0] - c[1][0] * c[2][2]) / det;
    inv[1][1] = (c[0][0] * c[2][2] - c[0][2] * c[2][0]) / det;
    inv[1][2] = (c[0][2] * c[1][0] - c[0][0] * c[1][2]) / det;
    inv[2][0] = (c[1][0] * c[2][1] - c[1][1] * c[2][0]) / det;
    inv[2][1] = (c[0][1] * c[2][0] - c[0][0] * c[2][1]) / det;
    inv[2][2] = (c[0][0] * c[1][1] - c[0][1] * c[1][0]) / det;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { g->cell[3 * i + j] = c[i][j]; g->inv[3 * i + j] = inv[i][j]; }
    for (int d = 0; d < 3; ++d) {
        double v[3];
        cross(c[(d + 1) % 3], c[(d + 2) % 3], v);
        g->height[d] = fabs(det) / norm(v);
        g->pbc[d] = pbc[d] ? 1 : 0;
    }
    g->ortho = (c[0][1] == 0 && c[0][2] == 0 && c[1][0] == 0 && c[1][2] == 0 && c[2][0] == 0 && c[2][1] == 0);
    return MCPYB_OK;
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// Reserve every per-batch device array.
int reserve_batch(mcpyb_engine* e, int R, int total, int max_n) {
    int mc = next_pow2(max_n > 0 ? max_n : 1);
    if (mc < 512) mc = 512;
    if (mc > 32768) mc = 32768;
    if (mc > e->max_cells) e->max_cells = mc;
    const size_t cs = 2 * (size_t)(e->max_cells + 1);
    CK(e->grid.reserve(R), "cudaMalloc grid");
    CK(e->geom.reserve(R), "cudaMalloc geom");
    CK(e->cell_start.reserve(cs * R), "cudaMalloc cell_start");
    CK(e->E.reserve(3 * (size_t)R), "cudaMalloc E");
    const size_t na = total > 0 ? total : 1;
    CK(e->upos.reserve(na), "cudaMalloc upos");
    CK(e->spos.reserve(na), "cudaMalloc spos");
    CK(e->stmp.reserve(na), "cudaMalloc stmp");
    CK(e->fpos.reserve(na), "cudaMalloc fpos");
    CK(e->cid.reserve(na), "cudaMalloc cid");
    CK(e->e_atom.reserve(na), "cudaMalloc e_atom");
    CK(e->sigma1.reserve(na), "cudaMalloc sigma1");
    CK(e->e_emb.reserve(na), "cudaMalloc e_emb");
    CK(e->pair_count.reserve(na), "cudaMalloc pair_count");
    return MCPYB_OK;
}

// Shared tail of the two upload paths: geometry + offsets to the device, one build launch.
int upload_common(mcpyb_engine* e, int R, const int64_t* atom_off, const double* cells,
                  const uint8_t* pbc, const double* d_pos, const int32_t* d_Z,
                  const unsigned char* d_header) {
    // d_header: [int atom_off[R+1]] padded to 16 | [CellGeom geom[R]]
    const size_t off_bytes = align_up((size_t)(R + 1) * sizeof(int), 16);
    const int* d_off = (const int*)d_header;
    const CellGeom* d_geom = (const CellGeom*)(d_header + off_bytes);
    (void)atom_off; (void)cells; (void)pbc;
    { int rc = chain_stop(e); if (rc != MCPYB_OK) return rc; }
    if (e->max_n >= 4096) {
        // large configurations: one 8-CTA thread-block cluster per replica
        CK(e->bbox.reserve((size_t)R * 8 * 6), "cudaMalloc bbox");
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(R * 8); cfg.blockDim = dim3(CL_THREADS); cfg.stream = e->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, build_cell_list_kernel<8>, e->view(), d_pos, (const int*)d_Z, d_off, d_geom,
                              (const int*)e->z_to_slot.p, e->rcut, e->subdiv, e->bbox.p), "launch build (cluster)");
        e->launches++;
    } else {
        build_cell_list_kernel<1><<<R, CL_THREADS, 0, e->stream>>>(
            e->view(), d_pos, d_Z, d_off, d_geom, e->z_to_slot.p, e->rcut, e->subdiv, nullptr);
        LAUNCH_CHECK("build_cell_list_kernel");
    }
    e->d_atom_off = d_off;          // stays valid in stage_dev until the next upload
    e->have_batch = true;
    e->emt_state_valid = false;
    e->status_checked = false;
    e->exp_valid = false;
    e->exp_batches = 0;
    return MCPYB_OK;
}

int validate_batch_args(mcpyb_engine* e, int R, const int64_t* atom_off, int* total, int* max_n) {
    if (e->pot_kind == POT_NONE) return e->fail(MCPYB_ERR_STATE, "no potential set (call mcpyb_set_lj / mcpyb_set_emt)");
    if (R < 0) return e->fail(MCPYB_ERR_ARG, "R must be >= 0");
    if (R > 0 && !atom_off) return e->fail(MCPYB_ERR_ARG, "atom_off is NULL");
    int mx = 0;
    for (int r = 0; r < R; ++r) {
        const int64_t n = atom_off[r + 1] - atom_off[r];
        if (n < 0) return e->fail(MCPYB_ERR_ARG, "atom_off must be non-decreasing");
        if (n > 0xFFFFFF) return e->fail(MCPYB_ERR_CAPACITY, "more than 16777215 atoms in one configuration");
        if (n > mx) mx = (int)n;
    }
    if (R > 0 && atom_off[0] != 0) return e->fail(MCPYB_ERR_ARG, "atom_off[0] must be 0");
    const int64_t tot = R > 0 ? atom_off[R] : 0;
    if (tot > 0x7FFFFFF0) return e->fail(MCPYB_ERR_CAPACITY, "batch too large");
    *total = (int)tot;
    *max_n = mx;
    return MCPYB_OK;
}

// Fill the pinned header (atom_off as int32 + geometry) for an upload.
int fill_header(mcpyb_engine* e, int R, const int64_t* atom_off, const double* cells,
                const uint8_t* pbc, unsigned char* hdr) {
    int* off32 = (int*)hdr;
    for (int r = 0; r <= R; ++r) off32[r] = (int)atom_off[r];
    CellGeom* g = (CellGeom*)(hdr + align_up((size_t)(R + 1) * sizeof(int), 16));
    double cells_est = 0.0;
    for (int r = 0; r < R; ++r) {
        int rc = make_geom(e, cells + 9 * r, pbc + 3 * r, g + r);
        if (rc != MCPYB_OK) return rc;
        // the grid build_cell_list_kernel will choose for a periodic box (cell_list.cuh); an open direction spans the
        // atoms' extent, which only the device knows: taken as the full height (more cells, smaller mean)
        const double rc_safe = e->rcut * (1.0 + 1e-12);
        double want[3], total = 1.0;
        for (int d = 0; d < 3; ++d) {
            double nc = floor(g[r].height[d] * (double)e->subdiv / rc_safe);
            if (!(nc >= 1.0)) nc = 1.0;
            if (nc > 1024.0) nc = 1024.0;
            want[d] = nc; total *= nc;
        }
        if (total > (double)e->max_cells) {
            const double f = cbrt((double)e->max_cells / total);
            total = 1.0;
            for (int d = 0; d < 3; ++d) total *= fmax(1.0, floor(want[d] * f));
        }
        cells_est += total < (double)e->max_cells ? total : (double)e->max_cells;
    }
    e->cells_est = cells_est;
    e->atom_off_host.assign(off32, off32 + R + 1);
    return MCPYB_OK;
}

size_t header_bytes(int R) {
    return align_up(align_up((size_t)(R + 1) * sizeof(int), 16) + (size_t)R * sizeof(CellGeom), 64);
}

int launch_trials_fwd(mcpyb_engine* e, int T, const int32_t* d_rep, const int32_t* d_old_off, const int32_t* d_old_idx,
                      const int32_t* d_new_off, int n_old_total, double* d_dE, int32_t* d_pairs);

int launch_energies(mcpyb_engine* e) {
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch (call mcpyb_upload_* first)");
    if (e->pot_kind == POT_EMT) { int rc = chain_stop(e); if (rc != MCPYB_OK) return rc; }   // sigma1 / e_atom are rewritten
    const int total = e->total_atoms;
    if (total >= 16384 && e->pot_kind == POT_LJ && e->use_warp_rows && e->use_expand && e->R > 0) {
        // large LJ batches: every atom's row from the trial pipeline (rows_warp.cuh)
        const size_t tp = (size_t)total + 1;
        CK(e->self_trials.reserve(4 * tp), "cudaMalloc self trials");
        CK(e->dE.reserve(tp), "cudaMalloc dE");
        CK(e->tpairs.reserve(tp), "cudaMalloc trial pairs");
        int* rep = e->self_trials.p; int* oo = rep + tp; int* oi = oo + tp; int* no = oi + tp;
        self_trials_kernel<<<(total + 256) / 256, 256, 0, e->stream>>>(e->d_atom_off, e->R, total, rep, oo, oi, no);
        LAUNCH_CHECK("self_trials_kernel");
        int rc = launch_trials_fwd(e, total, rep, oo, oi, no, total, e->dE.p, e->tpairs.p);
        if (rc != MCPYB_OK) return rc;
        rows_to_atom_energy_kernel<<<(total + 255) / 256, 256, 0, e->stream>>>(total, e->dE.p, e->tpairs.p, e->e_atom.p, e->pair_count.p);
        LAUNCH_CHECK("rows_to_atom_energy_kernel");
        reduce_replica_kernel<<<e->R, e->max_n > 1024 ? 1024 : 256, 0, e->stream>>>(
            e->e_atom.p, e->d_atom_off, e->E.p, e->pair_count.p, e->pairs_ptr(), e->grid.p, e->status_ptr(), nullptr, nullptr, 1);
        LAUNCH_CHECK("reduce_replica_kernel");
        return MCPYB_OK;
    }
    if (total > 0) {
        const int warps_per_block = 8;
        int blocks = (total + warps_per_block - 1) / warps_per_block;
        if (blocks > 148 * 64) blocks = 148 * 64;
        if (e->pot_kind == POT_LJ) {
            // lane-per-atom: work items = (cell, chunk of 32 atoms, stencil z-plane); grid.y = replica
            const int nplanes = 2 * e->subdiv + 1;
            CK(e->e_part.reserve((size_t)total * nplanes), "cudaMalloc e_part");
            CK(e->cnt_part.reserve((size_t)total * nplanes), "cudaMalloc cnt_part");
            long long est = ((long long)(e->max_n < e->max_cells ? e->max_n : e->max_cells) + e->max_n / 32 + 1) * nplanes;
            int gx = (int)((est + SITE_WARPS - 1) / SITE_WARPS);
            if (gx > 148 * 16) gx = 148 * 16;
            if (gx < 1) gx = 1;
            lj_cell_energy_kernel<<<dim3(gx, e->R), 32 * SITE_WARPS, 0, e->stream>>>(e->view(), e->lj, nplanes,
                                                                                   e->e_part.p, e->cnt_part.p);
            LAUNCH_CHECK("lj_cell_energy_kernel");
        } else {
            // EMT keeps the warp-per-atom traversal: with ~17 atoms per cell (edge = rc_list) the
            // lane-per-atom kernel (emt_cell_pair_kernel) leaves half the lanes idle and measured
            // 185 us against 100 us on the 10 825-atom S5 configuration (profiles/README.md).
            emt_pair_kernel<<<blocks, 256, 0, e->stream>>>(
                e->view(), e->d_atom_off, total, e->emt_dev.p, e->e_atom.p, e->sigma1.p, e->pair_count.p);
            LAUNCH_CHECK("emt_pair_kernel");
            const int eb = (total + 255) / 256;
            emt_embed_kernel<<<eb, 256, 0, e->stream>>>(e->view(), total, e->emt_dev.p, e->sigma1.p, e->e_atom.p, e->e_emb.p);
            LAUNCH_CHECK("emt_embed_kernel");
            e->emt_state_valid = true;
        }
    }
    if (e->R > 0) {
        reduce_replica_kernel<<<e->R, e->max_n > 1024 ? 1024 : 256, 0, e->stream>>>(
            e->e_atom.p, e->d_atom_off, e->E.p, e->pair_count.p, e->pairs_ptr(), e->grid.p, e->status_ptr(),
            e->pot_kind == POT_LJ && e->total_atoms > 0 ? e->e_part.p : nullptr, e->cnt_part.p, 2 * e->subdiv + 1);
        LAUNCH_CHECK("reduce_replica_kernel");
    }
    return MCPYB_OK;
}

int status_to_error(mcpyb_engine* e, int st, int r) {
    if (st & 2) return e->fail(MCPYB_ERR_SPECIES, "configuration %d contains an atomic number without parameters for this potential", r);
    if (st & 4) return e->fail(MCPYB_ERR_GEOMETRY, "configuration %d has an atom more than 511 periodic boxes from the cell", r);
    if (st) return e->fail(MCPYB_ERR_CAPACITY, "configuration %d: internal capacity exceeded (status %d)", r, st);
    return MCPYB_OK;
}

}  // namespace

extern "C" {

const char* mcpyb_version(void) { return "mcpy_b200 0.1.0 (sm_100a)"; }

int mcpyb_create(int device, mcpyb_engine** out) {
    if (!out) return MCPYB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || device < 0 || device >= count) return MCPYB_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return MCPYB_ERR_CUDA;
    mcpyb_engine* e = new mcpyb_engine();
    e->device = device;
    { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) e->num_sms = sms; }
    if (const char* v = getenv("MCPYB_PDL")) e->use_pdl = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_EXPAND")) e->use_expand = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_ROWS")) e->use_warp_rows = strcmp(v, "block") != 0;
    if (const char* v = getenv("MCPYB_CHAIN")) e->use_chain = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_EMT_LISTS")) e->use_emt_lists = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_ER_MINB")) e->er_minb = atoi(v);
    if (const char* v = getenv("MCPYB_TRK_TIMING")) e->trkb_timing = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_DIRECT")) e->use_direct = atoi(v) != 0;
    if (const char* v = getenv("MCPYB_SLOT_CAP")) e->slot_cap_override = atoi(v);
    if (const char* v = getenv("MCPYB_RW_MINB")) { const int n = atoi(v); if (n >= 4 && n <= 6) e->rw_minb = n; }
    if (const char* v = getenv("MCPYB_L2PERSIST")) {
        e->l2_persist_mb = atoi(v);
        if (e->l2_persist_mb > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)e->l2_persist_mb << 20);
    }
    if (const char* v = getenv("MCPYB_RB_SLICES")) { const int n = atoi(v); if (n >= 1 && n <= 8) e->rb_slices = n; }
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return MCPYB_ERR_CUDA;
    }
    e->stream = e->own_stream;
    for (int i = 0; i < 120; ++i) e->z_to_slot_host[i] = -1;
    *out = e;
    return MCPYB_OK;
}

void mcpyb_destroy(mcpyb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    chain_stop(e);
    cudaStreamSynchronize(e->stream);
    if (e->trkb_timing && e->trkb_calls)
        fprintf(stderr, "[mcpyb] batched tracked calls %lld: compare %.1f us, build %.1f us, trial call %.1f us (stage + H2D %.1f, launch %.1f, wait %.1f) (means)\n", e->trkb_calls,
                1e6 * e->trkb_t_compare / e->trkb_calls, 1e6 * e->trkb_t_build / e->trkb_calls, 1e6 * e->trkb_t_trial / e->trkb_calls,
                1e6 * e->trkb_t_stage / e->trkb_calls, 1e6 * e->trkb_t_launch / e->trkb_calls, 1e6 * e->trkb_t_wait / e->trkb_calls);
    if (e->chain_mb) cudaFreeHost(e->chain_mb);
    if (e->chain_ev) cudaEventDestroy(e->chain_ev);
    if (e->chain_stream) cudaStreamDestroy(e->chain_stream);
    e->emt_dev.release(); e->z_to_slot.release(); e->grid.release(); e->upos.release(); e->spos.release(); e->stmp.release(); e->fpos.release();
    e->cid.release(); e->cell_start.release(); e->geom.release(); e->bbox.release();
    e->exp_q.release(); e->exp_info.release(); e->exp_head.release(); e->exp_cursor.release();
    e->e_atom.release(); e->sigma1.release(); e->e_emb.release(); e->e_part.release(); e->deds.release(); e->forces.release(); e->cnt_part.release(); e->E.release(); e->pair_count.release();
    e->stage_dev.release(); e->stage_pin.release(); e->out_pin.release();
    e->mask_dev.release(); e->mask_pin.release();
    e->trial_dev.release(); e->trial_pin.release(); e->trial_out_pin.release(); e->dE.release(); e->tpairs.release();
    e->tw_int.release(); e->tw_bins.release(); e->tw_pos.release(); e->tw_E.release(); e->tw_spos.release(); e->tw_meta.release(); e->tw_items.release(); e->tw_a.release(); e->tw_err.release(); e->tw_ovf.release(); e->tw_hist.release();
    e->nl_i.release(); e->nl_j.release(); e->nl_S.release(); e->nl_r2.release(); e->nl_cursor.release();
    e->fv_grid.release(); e->fv_spheres.release(); e->fv_cell_start.release(); e->fv_cursor.release();
    e->fv_cid.release(); e->fv_points.release(); e->fv_pos.release(); e->fv_radii.release(); e->fv_small.release();
    e->fv_counts.release(); e->fv_mask.release(); e->fv_pin.release(); e->fv_atoms_pin.release(); e->fv_out_pin.release();
    e->pcg_count.release(); e->pcg_prefix.release(); e->pcg_ctl.release();
    delete e->copy_pool;
    if (e->trial_ev) cudaEventDestroy(e->trial_ev);
    if (e->copy_ev) cudaEventDestroy(e->copy_ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

const char* mcpyb_last_error(const mcpyb_engine* e) { return e ? e->err.c_str() : "null engine"; }

int mcpyb_set_stream(mcpyb_engine* e, void* cuda_stream) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return MCPYB_OK;
}

int mcpyb_synchronize(mcpyb_engine* e) {
    if (!e) return MCPYB_ERR_ARG;
    CK(cudaStreamSynchronize(e->stream), "cudaStreamSynchronize");
    return MCPYB_OK;
}

int64_t mcpyb_launch_count(const mcpyb_engine* e) { return e ? e->launches : 0; }

int mcpyb_set_precision(mcpyb_engine* e, int fp32_pair_math) {
    if (!e) return MCPYB_ERR_ARG;
    e->fp32_pairs = fp32_pair_math != 0;
    return MCPYB_OK;
}

int mcpyb_resident_chain(mcpyb_engine* e, int enable, int64_t* launches_out, int64_t* trials_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (enable == 0) { int rc = chain_stop(e); if (rc != MCPYB_OK) return rc; e->use_chain = false; }
    else if (enable > 0) e->use_chain = true;
    if (launches_out) *launches_out = e->chain_launches;
    if (trials_out) *trials_out = e->chain_trials;
    return MCPYB_OK;
}
double mcpyb_cutoff(const mcpyb_engine* e) { return e ? e->rcut : 0.0; }

int mcpyb_set_lj(mcpyb_engine* e, double epsilon, double sigma, double rc) {
    if (!e) return MCPYB_ERR_ARG;
    if (!(sigma > 0.0) || !isfinite(epsilon)) return e->fail(MCPYB_ERR_ARG, "sigma must be > 0 and epsilon finite");
    if (!(rc > 0.0)) rc = 3.0 * sigma;
    cudaSetDevice(e->device);
    chain_stop(e);
    e->pot_kind = POT_LJ;
    // Trial batches put many sites in every cell, so fine cells (edge rc/2) pay: 343 instead of 844
    // candidates per site.  Total energies have one site per atom: coarse cells (edge rc) fill the
    // 32 lanes of a work item (31 atoms per cell at rho* = 0.6) where fine cells leave 29 idle.
    // cells of edge rc / 3: 135 pruned candidates per site instead of 185 at rc / 2 (row sums 218 -> 201 us for the
    // bench's 8 x 65 536 proposals; MCPYB_SUBDIV=2 restores the coarser grid for A/B runs)
    e->subdiv_trials = 3; e->subdiv_energy = 1; e->subdiv = 3;
    if (const char* v = getenv("MCPYB_SUBDIV")) { const int n = atoi(v); if (n >= 1 && n <= 3) { e->subdiv_trials = n; e->subdiv = n; } }
    e->rcut = rc;
    e->lj.eps4 = 4.0 * epsilon;
    e->lj.sigma2 = sigma * sigma;
    e->lj.rc2 = rc * rc;
    e->lj.rc = rc;
    e->lj.e0 = 4.0 * epsilon * (pow(sigma / rc, 12) - pow(sigma / rc, 6));
    for (int i = 0; i < 120; ++i) e->z_to_slot_host[i] = 0;       // LJ is species-blind
    CK(e->z_to_slot.reserve(120), "cudaMalloc z_to_slot");
    CK(cudaMemcpyAsync(e->z_to_slot.p, e->z_to_slot_host, sizeof e->z_to_slot_host, cudaMemcpyHostToDevice, e->stream), "upload z_to_slot");
    CK(cudaStreamSynchronize(e->stream), "sync");
    e->have_batch = false;
    return MCPYB_OK;
}

int mcpyb_set_emt(mcpyb_engine* e) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    EmtTable& t = e->emt_host;
    memset(&t, 0, sizeof t);
    double maxseq = 0.0;
    const int ns = (int)(sizeof kEmtRows / sizeof kEmtRows[0]);
    for (int k = 0; k < ns; ++k) maxseq = fmax(maxseq, kEmtRows[k].s0 * kBohr);
    const double rc = kEmtBeta * maxseq * 0.5 * (sqrt(3.0) + sqrt(4.0));
    const double rr = rc * 2 * sqrt(4.0) / (sqrt(3.0) + sqrt(4.0));
    const double acut = log(9999.0) / (rr - rc);
    t.rc = rc; t.acut = acut; t.beta = kEmtBeta; t.inv_beta = 1.0 / kEmtBeta; t.rc_list = rc + 0.5; t.nspecies = ns;
    for (int i = 0; i < 120; ++i) e->z_to_slot_host[i] = -1;
    for (int k = 0; k < ns; ++k) {
        const EmtRow& p = kEmtRows[k];
        EmtSpecies& s = t.sp[k];
        s.E0 = p.E0; s.s0 = p.s0 * kBohr; s.V0 = p.V0; s.eta2 = p.eta2 / kBohr; s.kappa = p.kappa / kBohr;
        s.lambda = p.lambda / kBohr; s.n0 = p.n0 / (kBohr * kBohr * kBohr);
        double g1 = 0.0, g2 = 0.0;
        const int nn[3] = {12, 6, 24};
        for (int i = 0; i < 3; ++i) {
            const double r = s.s0 * kEmtBeta * sqrt((double)(i + 1));
            const double x = nn[i] / (12 * (1.0 + exp(acut * (r - rc))));
            g1 += x * exp(-s.eta2 * (r - kEmtBeta * s.s0));
            g2 += x * exp(-s.kappa / kEmtBeta * (r - kEmtBeta * s.s0));
        }
        s.gamma1 = g1; s.gamma2 = g2; s.inv_gamma1 = 1.0 / g1; s.inv_gamma2 = 1.0 / g2;
        e->z_to_slot_host[p.Z] = k;
    }
    for (int a = 0; a < ns; ++a)
        for (int b = 0; b < ns; ++b) t.ksi[a][b] = t.sp[b].n0 / t.sp[a].n0;
    e->pot_kind = POT_EMT;
    e->subdiv_trials = 1; e->subdiv_energy = 1; e->subdiv = 1;
    e->rcut = t.rc_list;
    CK(e->emt_dev.reserve(1), "cudaMalloc emt table");
    CK(e->z_to_slot.reserve(120), "cudaMalloc z_to_slot");
    CK(cudaMemcpyAsync(e->emt_dev.p, &t, sizeof t, cudaMemcpyHostToDevice, e->stream), "upload emt table");
    CK(cudaMemcpyAsync(e->z_to_slot.p, e->z_to_slot_host, sizeof e->z_to_slot_host, cudaMemcpyHostToDevice, e->stream), "upload z_to_slot");
    CK(cudaStreamSynchronize(e->stream), "sync");
    e->have_batch = false;
    return MCPYB_OK;
}

int mcpyb_emt_species(const mcpyb_engine* e, int Z, double* out9) {
    if (!e || !out9 || e->pot_kind != POT_EMT || Z < 0 || Z >= 120 || e->z_to_slot_host[Z] < 0) return MCPYB_ERR_ARG;
    const EmtSpecies& s = e->emt_host.sp[e->z_to_slot_host[Z]];
    const double v[9] = {s.E0, s.s0, s.V0, s.eta2, s.kappa, s.lambda, s.n0, s.gamma1, s.gamma2};
    memcpy(out9, v, sizeof v);
    return MCPYB_OK;
}

// helper threads for host copies of a megabyte or more (nullptr: copy inline)
static CopyPool* copy_pool_for(mcpyb_engine* e, size_t bytes) {
    if (bytes < (1u << 20)) return nullptr;
    if (!e->copy_pool) {
        // helper threads: MCPYB_POOL_THREADS, else 7 of >= 16 cores, 3 of >= 8, 1 of >= 4
        const unsigned hw = std::thread::hardware_concurrency();
        int nw = hw >= 16 ? 7 : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0)), spin = 150;
        if (const char* v = getenv("MCPYB_POOL_THREADS")) { nw = atoi(v); if (nw < 0) nw = 0; if (nw > 63) nw = 63; }
        if (const char* v = getenv("MCPYB_POOL_SPIN_US")) { spin = atoi(v); if (spin < 0) spin = 0; }
        e->copy_pool = new CopyPool(nw, spin);
    }
    return e->copy_pool;
}

// Page-locked host memory (cudaHostAlloc / cudaHostRegister / mcpyb_alloc_pinned) can be handed to the copy
// engine as it is; anything else has to pass through a pinned staging block first.
static bool is_pinned_host(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int upload_host_impl(mcpyb_engine* e, int subdiv, int R, const int64_t* atom_off, const double* positions,
                            const int32_t* numbers, const double* cells, const uint8_t* pbc) {
    if (!e) return MCPYB_ERR_ARG;
    e->subdiv = subdiv;
    e->trk_valid = false;
    e->trkb_valid = false;
    cudaSetDevice(e->device);
    int total = 0, max_n = 0;
    int rc = validate_batch_args(e, R, atom_off, &total, &max_n);
    if (rc != MCPYB_OK) return rc;
    if (R > 0 && (!cells || !pbc)) return e->fail(MCPYB_ERR_ARG, "cells / pbc are NULL");
    if (total > 0 && !positions) return e->fail(MCPYB_ERR_ARG, "positions is NULL");
    if (total > 0 && !numbers && e->pot_kind == POT_EMT) return e->fail(MCPYB_ERR_ARG, "EMT needs atomic numbers");
    e->have_batch = false;
    rc = chain_stop(e);
    if (rc != MCPYB_OK) return rc;
    rc = reserve_batch(e, R > 0 ? R : 1, total, max_n);
    if (rc != MCPYB_OK) return rc;
    e->R = R; e->total_atoms = total; e->max_n = max_n;
    if (R == 0) { e->have_batch = true; e->atom_off_host.assign(1, 0); return MCPYB_OK; }
    // one pinned block, one H2D copy: header | positions | numbers
    const size_t hb = header_bytes(R);
    const size_t pb = align_up((size_t)total * 3 * sizeof(double), 64);
    const size_t zb = align_up((size_t)total * sizeof(int32_t), 64);
    const size_t bytes = hb + pb + zb;
    CK(cudaStreamSynchronize(e->stream), "sync before staging reuse");
    CK(e->stage_pin.reserve(bytes), "cudaMallocHost stage");
    CK(e->stage_dev.reserve(bytes), "cudaMalloc stage");
    rc = fill_header(e, R, atom_off, cells, pbc, e->stage_pin.p);
    if (rc != MCPYB_OK) return rc;
    if (total > 0) {
        CopyPool* pool = copy_pool_for(e, (size_t)total * 24);
        if (pool) pool->copy(e->stage_pin.p + hb, positions, (size_t)total * 3 * sizeof(double));
        else memcpy(e->stage_pin.p + hb, positions, (size_t)total * 3 * sizeof(double));
        if (numbers) memcpy(e->stage_pin.p + hb + pb, numbers, (size_t)total * sizeof(int32_t));
    }
    CK(cudaMemcpyAsync(e->stage_dev.p, e->stage_pin.p, bytes, cudaMemcpyHostToDevice, e->stream), "H2D batch");
    return upload_common(e, R, atom_off, cells, pbc, (const double*)(e->stage_dev.p + hb),
                         numbers ? (const int32_t*)(e->stage_dev.p + hb + pb) : nullptr, e->stage_dev.p);
}

int mcpyb_upload_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                      const int32_t* numbers, const double* cells, const uint8_t* pbc) {
    if (!e) return MCPYB_ERR_ARG;
    return upload_host_impl(e, e->subdiv_trials, R, atom_off, positions, numbers, cells, pbc);
}

int mcpyb_upload_dev(mcpyb_engine* e, int R, const int64_t* atom_off, const double* d_positions,
                     const int32_t* d_numbers, const double* cells, const uint8_t* pbc) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    e->subdiv = e->subdiv_trials;
    e->trk_valid = false;
    e->trkb_valid = false;
    int total = 0, max_n = 0;
    int rc = validate_batch_args(e, R, atom_off, &total, &max_n);
    if (rc != MCPYB_OK) return rc;
    if (R > 0 && (!cells || !pbc)) return e->fail(MCPYB_ERR_ARG, "cells / pbc are NULL");
    if (total > 0 && !d_positions) return e->fail(MCPYB_ERR_ARG, "positions is NULL");
    if (total > 0 && !d_numbers && e->pot_kind == POT_EMT) return e->fail(MCPYB_ERR_ARG, "EMT needs atomic numbers");
    e->have_batch = false;
    rc = chain_stop(e);
    if (rc != MCPYB_OK) return rc;
    rc = reserve_batch(e, R > 0 ? R : 1, total, max_n);
    if (rc != MCPYB_OK) return rc;
    e->R = R; e->total_atoms = total; e->max_n = max_n;
    if (R == 0) { e->have_batch = true; e->atom_off_host.assign(1, 0); return MCPYB_OK; }
    const size_t hb = header_bytes(R);
    // the previous upload may still be reading the pinned header: wait for it
    CK(cudaStreamSynchronize(e->stream), "sync before header reuse");
    CK(e->stage_pin.reserve(hb), "cudaMallocHost stage");
    CK(e->stage_dev.reserve(hb), "cudaMalloc stage");
    rc = fill_header(e, R, atom_off, cells, pbc, e->stage_pin.p);
    if (rc != MCPYB_OK) return rc;
    CK(cudaMemcpyAsync(e->stage_dev.p, e->stage_pin.p, hb, cudaMemcpyHostToDevice, e->stream), "H2D header");
    return upload_common(e, R, atom_off, cells, pbc, d_positions, d_numbers, e->stage_dev.p);
}

// ---- many small device copies in one launch ----------------------------------------------------
#define SEG_MAX 48
struct CopySegs {
    const void* src[SEG_MAX];
    void* dst[SEG_MAX];
    int count[SEG_MAX];
    unsigned char kind[SEG_MAX];          // 0: f64 -> f64, 1: i32 -> i32, 2: f64 -> i32, 3: i32 -> f64
    int n;
};

__global__ void __launch_bounds__(256)
copy_segments_kernel(const __grid_constant__ CopySegs sg) {
    const int k = blockIdx.y;
    const int n = sg.count[k], kind = sg.kind[k];
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    if (kind == 0) {
        const double* s = (const double*)sg.src[k]; double* d = (double*)sg.dst[k];
        for (int i = i0; i < n; i += step) d[i] = s[i];
    } else if (kind == 1) {
        const int* s = (const int*)sg.src[k]; int* d = (int*)sg.dst[k];
        for (int i = i0; i < n; i += step) d[i] = s[i];
    } else if (kind == 2) {
        const double* s = (const double*)sg.src[k]; int* d = (int*)sg.dst[k];
        for (int i = i0; i < n; i += step) d[i] = (int)s[i];
    } else {
        const int* s = (const int*)sg.src[k]; double* d = (double*)sg.dst[k];
        for (int i = i0; i < n; i += step) d[i] = (double)s[i];
    }
}

int mcpyb_copy_segments_dev(mcpyb_engine* e, int nseg, const uint64_t* src, const uint64_t* dst, const int64_t* count,
                            const uint8_t* kind) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (nseg < 0 || (nseg > 0 && (!src || !dst || !count || !kind))) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    for (int s0 = 0; s0 < nseg; s0 += SEG_MAX) {
        CopySegs sg;
        memset(&sg, 0, sizeof sg);
        sg.n = nseg - s0 < SEG_MAX ? nseg - s0 : SEG_MAX;
        int64_t most = 0;
        for (int k = 0; k < sg.n; ++k) {
            const int64_t c = count[s0 + k];
            if (c < 0 || c > 0x7fffffffLL || kind[s0 + k] > 3 || (c > 0 && (!src[s0 + k] || !dst[s0 + k])))
                return e->fail(MCPYB_ERR_ARG, "segment %d: bad count / kind / address", s0 + k);
            sg.src[k] = (const void*)(uintptr_t)src[s0 + k];
            sg.dst[k] = (void*)(uintptr_t)dst[s0 + k];
            sg.count[k] = (int)c; sg.kind[k] = kind[s0 + k];
            if (c > most) most = c;
        }
        if (most == 0) continue;
        int bx = (int)((most + 255) / 256);
        if (bx > 64) bx = 64;
        copy_segments_kernel<<<dim3(bx, sg.n), 256, 0, e->stream>>>(sg);
        LAUNCH_CHECK("copy_segments_kernel");
    }
    return MCPYB_OK;
}

int mcpyb_energies_dev(mcpyb_engine* e, double* d_energies_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    int rc = launch_energies(e);
    if (rc != MCPYB_OK) return rc;
    if (e->R > 0 && d_energies_out)
        CK(cudaMemcpyAsync(d_energies_out, e->E.p, (size_t)e->R * sizeof(double), cudaMemcpyDeviceToDevice, e->stream), "D2D energies");
    return MCPYB_OK;
}

int mcpyb_energies_host(mcpyb_engine* e, double* energies_out, int64_t* pairs_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    int rc = launch_energies(e);
    if (rc != MCPYB_OK) return rc;
    const int R = e->R;
    if (R == 0) return MCPYB_OK;
    if (!energies_out) return e->fail(MCPYB_ERR_ARG, "energies_out is NULL");
    // one device block -> one pinned block: E[R] | pairs[R] | status[R]
    const size_t eb = (size_t)R * sizeof(double), pb = (size_t)R * sizeof(long long), sb = (size_t)R * sizeof(int);
    CK(e->out_pin.reserve(eb + pb + sb), "cudaMallocHost out");
    CK(cudaMemcpyAsync(e->out_pin.p, e->E.p, eb + pb + sb, cudaMemcpyDeviceToHost, e->stream), "D2H energies");
    CK(cudaStreamSynchronize(e->stream), "sync energies");
    const int* st = (const int*)(e->out_pin.p + eb + pb);
    for (int r = 0; r < R; ++r) {
        rc = status_to_error(e, st[r], r);
        if (rc != MCPYB_OK) return rc;
    }
    memcpy(energies_out, e->out_pin.p, eb);
    if (pairs_out) memcpy(pairs_out, e->out_pin.p + eb, pb);
    e->status_checked = true;
    return MCPYB_OK;
}

int mcpyb_potential_energies_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                                  const int32_t* numbers, const double* cells, const uint8_t* pbc,
                                  double* energies_out) {
    if (!e) return MCPYB_ERR_ARG;
    // few atoms: latency matters, fine cells give more (shorter) work items; many atoms (replica
    // batches): throughput matters, coarse cells keep the lanes full
    const int64_t tot = (R > 0 && atom_off) ? atom_off[R] : 0;
    int rc = upload_host_impl(e, tot >= 8192 ? e->subdiv_energy : e->subdiv_trials, R, atom_off, positions,
                              numbers, cells, pbc);
    if (rc != MCPYB_OK) return rc;
    return mcpyb_energies_host(e, energies_out, nullptr);
}

int mcpyb_atom_energies_host(mcpyb_engine* e, double* e_atom_out, double* sigma1_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    int rc = launch_energies(e);
    if (rc != MCPYB_OK) return rc;
    const size_t nb = (size_t)e->total_atoms * sizeof(double);
    if (e_atom_out && nb) CK(cudaMemcpyAsync(e_atom_out, e->e_atom.p, nb, cudaMemcpyDeviceToHost, e->stream), "D2H e_atom");
    if (sigma1_out && nb) {
        if (e->pot_kind != POT_EMT) return e->fail(MCPYB_ERR_STATE, "sigma1 exists for EMT only");
        CK(cudaMemcpyAsync(sigma1_out, e->sigma1.p, nb, cudaMemcpyDeviceToHost, e->stream), "D2H sigma1");
    }
    CK(cudaStreamSynchronize(e->stream), "sync atom energies");
    return MCPYB_OK;
}

// ---- forces ------------------------------------------------------------------------
// Forces of the resident batch, [sum N][3], for the ASE Calculator protocol (`atoms.calc = calc`,
// optimisers in CanonicalEnsemble.relax, canonical_ensemble.py:115-123).  EMT needs sigma1 of the
// batch first (one energy pass if it is not resident).
static int launch_forces(mcpyb_engine* e) {
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch (call mcpyb_upload_* first)");
    const int total = e->total_atoms;
    CK(e->forces.reserve(3 * (size_t)(total > 0 ? total : 1)), "cudaMalloc forces");
    if (total == 0) return MCPYB_OK;
    int blocks = (total + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (e->pot_kind == POT_LJ) {
        lj_forces_kernel<<<blocks, 256, 0, e->stream>>>(e->view(), e->d_atom_off, total, e->lj, e->forces.p);
        LAUNCH_CHECK("lj_forces_kernel");
    } else {
        if (!e->emt_state_valid) {
            int rc = launch_energies(e);
            if (rc != MCPYB_OK) return rc;
        }
        CK(e->deds.reserve((size_t)total), "cudaMalloc deds");
        emt_deds_kernel<<<(total + 255) / 256, 256, 0, e->stream>>>(e->view(), total, e->emt_dev.p, e->sigma1.p, e->deds.p);
        LAUNCH_CHECK("emt_deds_kernel");
        emt_forces_kernel<<<blocks, 256, 0, e->stream>>>(e->view(), e->d_atom_off, total, e->emt_dev.p, e->deds.p, e->forces.p);
        LAUNCH_CHECK("emt_forces_kernel");
    }
    return MCPYB_OK;
}

int mcpyb_forces_host(mcpyb_engine* e, double* forces_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    int rc;
    if (!e->status_checked) {
        // species / geometry errors of the resident batch surface through the energy path's status
        std::vector<double> tmp((size_t)(e->R > 0 ? e->R : 1));
        rc = mcpyb_energies_host(e, tmp.data(), nullptr);
        if (rc != MCPYB_OK) return rc;
    }
    rc = launch_forces(e);
    if (rc != MCPYB_OK) return rc;
    const size_t nb = 3 * (size_t)e->total_atoms * sizeof(double);
    if (nb == 0) return MCPYB_OK;
    if (!forces_out) return e->fail(MCPYB_ERR_ARG, "forces_out is NULL");
    CK(e->out_pin.reserve(nb), "cudaMallocHost out");
    CK(cudaMemcpyAsync(e->out_pin.p, e->forces.p, nb, cudaMemcpyDeviceToHost, e->stream), "D2H forces");
    CK(cudaStreamSynchronize(e->stream), "sync forces");
    memcpy(forces_out, e->out_pin.p, nb);
    return MCPYB_OK;
}

// Upload + energies + forces in one call: what Calculator.calculate(atoms, ['energy', 'forces']) needs.
int mcpyb_energy_forces_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                             const int32_t* numbers, const double* cells, const uint8_t* pbc,
                             double* energies_out, double* forces_out) {
    if (!e) return MCPYB_ERR_ARG;
    int rc = mcpyb_potential_energies_host(e, R, atom_off, positions, numbers, cells, pbc, energies_out);
    if (rc != MCPYB_OK) return rc;
    return mcpyb_forces_host(e, forces_out);
}

// upload + energy + forces of ONE configuration with a single stream synchronisation (the optimiser loop below:
// mcpyb_energy_forces_host synchronises after the energies to check the build status before it launches the
// force kernels; from the second evaluation of a relaxation on, species and box are known to be good and only
// the geometry can still fail, which the energy kernels tolerate -- the status is checked after the fact).
static int energy_forces_one_sync(mcpyb_engine* e, const int64_t* off, const double* positions, const int32_t* numbers,
                                  const double* cell9, const uint8_t* pbc3, double* energy_out, double* forces_out) {
    int rc = upload_host_impl(e, off[1] >= 8192 ? e->subdiv_energy : e->subdiv_trials, 1, off, positions, numbers, cell9, pbc3);
    if (rc != MCPYB_OK) return rc;
    if (e->pot_kind == POT_LJ && e->total_atoms > 0 && e->total_atoms <= 8192) {
        // one pass over the pairs gives forces and energies (lj_energy_forces_kernel); the reduction adds them up
        const int total = e->total_atoms;
        CK(e->forces.reserve(3 * (size_t)total), "cudaMalloc forces");
        int blocks = total < e->num_sms * 16 ? total : e->num_sms * 16;
        lj_energy_forces_kernel<<<blocks, 32 * EF_WARPS, 0, e->stream>>>(e->view(), e->d_atom_off, total, e->lj,
                                                                        e->forces.p, e->e_atom.p, e->pair_count.p);
        LAUNCH_CHECK("lj_energy_forces_kernel");
        reduce_replica_kernel<<<e->R, e->max_n > 1024 ? 1024 : 256, 0, e->stream>>>(
            e->e_atom.p, e->d_atom_off, e->E.p, e->pair_count.p, e->pairs_ptr(), e->grid.p, e->status_ptr(),
            nullptr, nullptr, 0);
        LAUNCH_CHECK("reduce_replica_kernel");
    } else {
        rc = launch_energies(e);
        if (rc != MCPYB_OK) return rc;
        rc = launch_forces(e);
        if (rc != MCPYB_OK) return rc;
    }
    const size_t blk = align_up(sizeof(double) + sizeof(long long) + sizeof(int), 64);
    const size_t nb = 3 * (size_t)e->total_atoms * sizeof(double);
    CK(e->out_pin.reserve(blk + nb), "cudaMallocHost out");
    CK(cudaMemcpyAsync(e->out_pin.p, e->E.p, sizeof(double) + sizeof(long long) + sizeof(int), cudaMemcpyDeviceToHost, e->stream), "D2H energy");
    if (nb) CK(cudaMemcpyAsync(e->out_pin.p + blk, e->forces.p, nb, cudaMemcpyDeviceToHost, e->stream), "D2H forces");
    CK(cudaStreamSynchronize(e->stream), "sync energy + forces");
    rc = status_to_error(e, *(const int*)(e->out_pin.p + sizeof(double) + sizeof(long long)), 0);
    if (rc != MCPYB_OK) return rc;
    e->status_checked = true;
    memcpy(energy_out, e->out_pin.p, sizeof(double));
    if (nb) memcpy(forces_out, e->out_pin.p + blk, nb);
    return MCPYB_OK;
}

// ---- relax-then-energy ---------------------------------------------------------------
// BaseCalculator.get_potential_energy (mcpy/calculators/base_calculator.py:14-22) and CanonicalEnsemble.relax
// (mcpy/ensembles/canonical_ensemble.py:115-123) hand every trial configuration to ase.optimize.LBFGS before
// they read its energy.  Through the ASE protocol that is one Python round trip per optimiser step; here the
// optimiser loop lives next to the force kernels: each step is one upload + cell list + energy + forces of the
// current positions (mcpyb_energy_forces_host), and the two-loop recursion over 3n doubles runs on the host
// between them (a few microseconds at these sizes -- far below one kernel launch per dot product).
// The iteration is ASE 3.23's LBFGS without line search [recalled, ase/optimize/lbfgs.py; restated in
// oracle/lbfgs.py]: H0 = 1/alpha, at most `memory` (s, y) pairs, step p = -H g scaled so that no atom moves
// farther than maxstep, damping; converged when max_i |f_i| < fmax, checked before every step; at most
// max_steps steps.  Atoms with fixed[i] != 0 feel no force and do not move (FixAtoms).
int mcpyb_relax_lbfgs_host(mcpyb_engine* e, double* positions, const int32_t* numbers, int n, const double* cell9,
                           const uint8_t* pbc3, const uint8_t* fixed, double fmax, int max_steps, double maxstep,
                           int memory, double alpha, double damping, double* energy_out, int* steps_out,
                           double* fmax_out) {
    if (!e) return MCPYB_ERR_ARG;
    if (n < 0 || !cell9 || !pbc3 || !energy_out || (n > 0 && !positions)) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    if (!(fmax >= 0.0) || max_steps < 0 || !(maxstep > 0.0) || memory < 1 || !(alpha > 0.0) || !(damping > 0.0))
        return e->fail(MCPYB_ERR_ARG, "bad optimiser parameters");
    const int64_t off[2] = {0, n};
    const size_t m3 = 3 * (size_t)n;
    std::vector<double> f(m3), r0, f0, q(m3), z(m3), a;
    std::vector<std::vector<double>> S, Y;
    std::vector<double> rho;
    double E = 0.0;
    bool trusted = false;                        // species / box accepted by a first, fully checked evaluation
    auto evaluate = [&]() -> int {
        int rc = trusted ? energy_forces_one_sync(e, off, positions, numbers, cell9, pbc3, &E, f.data())
                         : mcpyb_energy_forces_host(e, 1, off, positions, numbers, cell9, pbc3, &E, f.data());
        if (rc != MCPYB_OK) return rc;
        trusted = true;
        if (fixed) for (int i = 0; i < n; ++i) if (fixed[i]) f[3 * i] = f[3 * i + 1] = f[3 * i + 2] = 0.0;
        return MCPYB_OK;
    };
    auto fmax2_of = [&]() {
        double m = 0.0;
        for (int i = 0; i < n; ++i) {
            const double v = f[3 * i] * f[3 * i] + f[3 * i + 1] * f[3 * i + 1] + f[3 * i + 2] * f[3 * i + 2];
            if (!(v <= m)) m = v;                 // NaN propagates: never converged, the step limit ends the loop
        }
        return m;
    };
    auto dot = [&](const double* u, const double* v) { double t = 0.0; for (size_t k = 0; k < m3; ++k) t += u[k] * v[k]; return t; };
    int rc = evaluate();
    if (rc != MCPYB_OK) return rc;
    int iteration = 0, nsteps = 0;
    while (n > 0 && nsteps < max_steps && !(fmax2_of() < fmax * fmax)) {
        // update(): curvature pair of the previous step
        if (iteration > 0) {
            std::vector<double> s0(m3), y0(m3);
            for (size_t k = 0; k < m3; ++k) { s0[k] = positions[k] - r0[k]; y0[k] = f0[k] - f[k]; }
            rho.push_back(1.0 / dot(y0.data(), s0.data()));
            S.push_back(std::move(s0)); Y.push_back(std::move(y0));
        }
        if (iteration > memory) { S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin()); }
        const int loopmax = (int)S.size() < (memory < iteration ? memory : iteration) ? (int)S.size() : (memory < iteration ? memory : iteration);
        a.assign((size_t)(loopmax > 0 ? loopmax : 1), 0.0);
        for (size_t k = 0; k < m3; ++k) q[k] = -f[k];
        for (int i = loopmax - 1; i >= 0; --i) {
            a[i] = rho[i] * dot(S[i].data(), q.data());
            for (size_t k = 0; k < m3; ++k) q[k] -= a[i] * Y[i][k];
        }
        const double H0 = 1.0 / alpha;
        for (size_t k = 0; k < m3; ++k) z[k] = H0 * q[k];
        for (int i = 0; i < loopmax; ++i) {
            const double bb = rho[i] * dot(Y[i].data(), z.data());
            for (size_t k = 0; k < m3; ++k) z[k] += S[i][k] * (a[i] - bb);
        }
        // determine_step(): p = -z, scaled so that the longest atomic step is maxstep
        double longest = 0.0;
        for (int i = 0; i < n; ++i) {
            const double l = sqrt(z[3 * i] * z[3 * i] + z[3 * i + 1] * z[3 * i + 1] + z[3 * i + 2] * z[3 * i + 2]);
            if (l > longest) longest = l;
        }
        const double scale = (longest >= maxstep ? maxstep / longest : 1.0) * damping;
        r0.assign(positions, positions + m3);
        f0 = f;
        for (size_t k = 0; k < m3; ++k) positions[k] = r0[k] + (-z[k]) * scale;
        ++iteration;
        rc = evaluate();
        if (rc != MCPYB_OK) return rc;
        ++nsteps;
    }
    *energy_out = E;
    if (steps_out) *steps_out = nsteps;
    if (fmax_out) *fmax_out = sqrt(fmax2_of());
    return MCPYB_OK;
}

// ---- trial moves ---------------------------------------------------------------
// Per-site partial sums, status words and the slicing shared by every LJ trial path.
static int prepare_rows_work(mcpyb_engine* e, int nsites, int n_old_total, int n_new_total,
                             unsigned long long* d_status, TrialWork* wp) {
    TrialWork& w = *wp;
    memset(&w, 0, sizeof w);
    w.nsites = nsites; w.n_old = n_old_total; w.n_new = n_new_total;
    w.bin_stride = e->max_cells; w.nbins = e->R * e->max_cells;
    const size_t ns = (size_t)(nsites > 0 ? nsites : 1), nbp = (size_t)w.nbins + 1;
    // every path cuts a cell's candidate list into the same slices, so a trial's bits do not depend on
    // the batch it arrives in
    if (e->use_warp_rows) {
        // one folded row sum per site: the four interleaved classes of rows_warp.cuh
        w.nplanes = 1; w.item_slices = 1; w.item_shift = RW_SITES_SHIFT;
    } else {
        w.nplanes = e->rb_slices;
        w.item_slices = e->rb_slices;        // one work item walks all slices of its cell (one header for both)
        w.item_shift = RB_SITES_SHIFT;
    }
    const size_t nsp = ns * (size_t)w.nplanes;
    CK(e->tw_int.reserve(2 * ns + nsp + 4), "cudaMalloc trial work");
    CK(e->tw_E.reserve(nsp), "cudaMalloc trial work");
    if (!e->tw_err.p) {
        // [0] armed error word, [1] its value after the last launch (both ~0 = none), [2] CTA counter
        CK(e->tw_err.reserve(3), "cudaMalloc trial status");
        CK(cudaMemsetAsync(e->tw_err.p, 0xFF, 2 * sizeof(unsigned long long), e->stream), "memset trial status");
        CK(cudaMemsetAsync(e->tw_err.p + 2, 0, sizeof(unsigned long long), e->stream), "memset trial status");
    }
    w.site_trial = e->tw_int.p; w.sorted = e->tw_int.p + ns;
    w.site_cnt = e->tw_int.p + 2 * ns; w.totals = e->tw_int.p + 2 * ns + nsp;
    w.site_E = e->tw_E.p;
    w.err = e->tw_err.p; w.err_out = d_status ? d_status : e->tw_err.p + 1;
    return MCPYB_OK;
}

// The per-configuration stencil lists of the resident batch (rows_block.cuh), built when a SECOND trial launch
// meets the same resident configuration (one launch per upload does not amortise the 15 us of
// expand_stencils_kernel).  *xl is left zeroed -- every work item stages from the cell list -- until they exist.
// `few_sites`: the single-launch kernels of the one-call-per-trial chain, which re-base every few dozen calls:
// worth it for the configurations such a chain runs on, not for a 100 000-atom batch.
static int stencil_lists(mcpyb_engine* e, bool few_sites, StencilLists* xl) {
    memset(xl, 0, sizeof *xl);
    if (!e->use_expand || e->pot_kind == POT_NONE || !e->have_batch || e->total_atoms <= 0) return MCPYB_OK;
    if (e->pot_kind == POT_EMT && few_sites) return MCPYB_OK;        // the EMT chain / single-launch kernels walk the cell list
    if (few_sites && e->total_atoms > 20000) return MCPYB_OK;
    // the warp row-sum kernel reads nothing but the lists; the single-launch kernels wait for a second call
    if (!e->exp_valid && ++e->exp_batches < 2 && (few_sites || (!e->use_warp_rows && e->pot_kind == POT_LJ))) return MCPYB_OK;
    const int nbins = e->R * e->max_cells;
    const long long m_st = 2 * (long long)e->subdiv + 1;
    long long cap = (long long)e->total_atoms * (m_st * m_st * m_st + 3);   // (2m+1)^3 cells see each atom (125 at m = 2)
    // a cell's run is reserved for its whole stencil and filled with the pruned part; 40 B per record: 1.8 GB for a
    // 64 x 2 000-atom batch, capped at 20 GB of the 180 GB (cells beyond the cap take the exact per-lane traversal)
    if (cap > (512LL << 20)) cap = 512LL << 20;
    CK(e->exp_q.reserve((size_t)cap), "cudaMalloc stencil lists");
    CK(e->exp_info.reserve((size_t)cap), "cudaMalloc stencil lists");
    CK(e->exp_head.reserve((size_t)nbins), "cudaMalloc stencil lists");
    CK(e->exp_cursor.reserve(1), "cudaMalloc stencil lists");
    xl->q = e->exp_q.p; xl->info = e->exp_info.p; xl->head = e->exp_head.p; xl->cursor = e->exp_cursor.p; xl->cap = cap;
    if (!e->exp_valid) {
        CK(cudaMemsetAsync(e->exp_cursor.p, 0, sizeof(unsigned long long), e->stream), "memset stencil cursor");
        int blocks_x = (nbins + EX_WARPS - 1) / EX_WARPS;
        if (blocks_x > e->num_sms * 8) blocks_x = e->num_sms * 8;
        expand_stencils_kernel<<<blocks_x, 32 * EX_WARPS, 0, e->stream>>>(e->view(), *xl, nbins, e->max_cells, e->rcut);
        LAUNCH_CHECK("expand_stencils_kernel");
        e->exp_valid = true;
        if (e->l2_persist_mb > 0) {
            // the lists are read by every trial launch against this batch: ask L2 to keep them across the streaming
            // proposal / result traffic
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            size_t span = (size_t)cap * sizeof(double4);
            int maxwin = 0;
            cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
            if (maxwin > 0 && span > (size_t)maxwin) span = (size_t)maxwin;
            av.accessPolicyWindow.base_ptr = e->exp_q.p;
            av.accessPolicyWindow.num_bytes = span;
            const double set_aside = (double)((size_t)e->l2_persist_mb << 20);
            av.accessPolicyWindow.hitRatio = (float)(set_aside >= (double)span ? 1.0 : set_aside / (double)span);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        }
    }
    return MCPYB_OK;
}

static int launch_trials(mcpyb_engine* e, int T, const int32_t* d_rep, const int32_t* d_old_off,
                         const int32_t* d_old_idx, const int32_t* d_new_off, const double* d_new_pos,
                         const int32_t* d_new_Z, int n_old_total, int n_new_total,
                         double* d_dE, int32_t* d_pairs, unsigned long long* d_status = nullptr,
                         volatile unsigned int* d_flag = nullptr, unsigned int flag_seq = 0) {
    // d_flag (single-launch LJ batches only): d_dE / d_pairs / d_status are pinned host memory, the kernel's last CTA
    // writes flag_seq to *d_flag after them
    if (T == 0) return MCPYB_OK;
    TrialBatch tb;
    tb.T = T; tb.rep = d_rep; tb.old_off = d_old_off; tb.old_idx = d_old_idx;
    tb.new_off = d_new_off; tb.new_pos = d_new_pos; tb.new_Z = d_new_Z;
    int blocks = (T + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (e->pot_kind == POT_EMT && !e->emt_state_valid) {
        int rc = launch_energies(e);           // sigma1 / e_atom of the resident batch
        if (rc != MCPYB_OK) return rc;
    }
    const int nsites = n_old_total + n_new_total;
    if (e->pot_kind == POT_EMT) {
        StencilLists xe;
        memset(&xe, 0, sizeof xe);
        if (T >= 64 && e->use_emt_lists) {
            int rc = stencil_lists(e, false, &xe);
            if (rc != MCPYB_OK) return rc;
        }
        if (xe.head) {
            // one warp per trial over the pruned per-cell lists, hits compacted (emt_rows.cuh); trials it cannot take
            // (cells without a list) come back through emt_redo and run on the first-generation body
            CK(e->emt_redo.reserve((size_t)T + 1), "cudaMalloc emt redo");
            CK(cudaMemsetAsync(e->emt_redo.p, 0, sizeof(int), e->stream), "memset emt redo");
            int gb = (T + ER_WARPS - 1) / ER_WARPS;
            if (gb > e->num_sms * e->er_minb) gb = e->num_sms * e->er_minb;
#define ER_LAUNCH(M) emt_trial_rows_kernel<M><<<gb, 32 * ER_WARPS, 0, e->stream>>>( \
                e->view(), tb, e->emt_dev.p, e->z_to_slot.p, e->sigma1.p, e->e_atom.p, e->e_emb.p, xe, e->max_cells, d_dE, d_pairs, e->emt_redo.p)
            if (e->er_minb >= 8) ER_LAUNCH(8); else if (e->er_minb >= 6) ER_LAUNCH(6); else ER_LAUNCH(4);
#undef ER_LAUNCH
            LAUNCH_CHECK("emt_trial_rows_kernel");
            emt_trial_redo_kernel<<<e->num_sms, 256, 0, e->stream>>>(
                e->view(), tb, e->emt_dev.p, e->z_to_slot.p, e->sigma1.p, e->e_atom.p, d_dE, d_pairs, e->emt_redo.p);
            LAUNCH_CHECK("emt_trial_redo_kernel");
            return MCPYB_OK;
        }
        // one warp per trial, lanes over the candidates (see the note at the end of emt_kernels.cuh)
        emt_trial_delta_kernel<32><<<blocks, 256, 0, e->stream>>>(
            e->view(), tb, e->emt_dev.p, e->z_to_slot.p, e->sigma1.p, e->e_atom.p, d_dE, d_pairs);
        LAUNCH_CHECK("emt_trial_delta_kernel");
        return MCPYB_OK;
    }
    TrialWork w;
    {
        int rc = prepare_rows_work(e, nsites, n_old_total, n_new_total, d_status, &w);
        if (rc != MCPYB_OK) return rc;
    }
    const size_t ns = (size_t)(nsites > 0 ? nsites : 1), nbp = (size_t)w.nbins + 1;
    e->stage_n = 0;
    if (nsites <= RS_MAX_SITES) {
        // a handful of sites (the one-call-per-trial chain): one launch does everything
        if (nsites > 0) {
            StencilLists xs;
            int rc = stencil_lists(e, true, &xs);
            if (rc != MCPYB_OK) return rc;
            lj_trial_small_kernel<<<nsites, RS_THREADS, 0, e->stream>>>(e->view(), tb, w, e->lj, d_dE, d_pairs,
                                                                         (int*)(e->tw_err.p + 2), xs, d_flag, flag_seq);
            LAUNCH_CHECK("lj_trial_small_kernel");
        } else {
            lj_trial_combine_all_kernel<<<(T + 127) / 128, 128, 0, e->stream>>>(e->view(), tb, w, e->lj, d_dE, d_pairs);
            LAUNCH_CHECK("lj_trial_combine_all_kernel");
            if (d_status) CK(cudaMemsetAsync(d_status, 0xFF, sizeof(unsigned long long), e->stream), "memset status");
        }
        return MCPYB_OK;
    }
    // dense batches: direct placement (site_kernels.cuh: trial_sites_direct_kernel) -- the rank a site draws from its
    // bin's histogram is its slot, the row-sum warps claim bins: no scan, no scatter, no item table
    int slot_cap = 0;
    if (e->use_direct && e->use_warp_rows && !e->fp32_pairs && w.nbins > 0) {
        // sites per occupied-grid cell: the grids are sized on the device, the host knows them for periodic boxes
        // (make_geom) and otherwise falls back on the stride (an underestimate of the mean: more overflow, same bits)
        const double cells = e->cells_est > 0.0 ? e->cells_est : (double)w.nbins;
        const double mean = (double)nsites / cells;
        if (mean >= 4.0 || e->slot_cap_override > 0) {
            slot_cap = e->slot_cap_override > 0 ? e->slot_cap_override : 16 * (int)ceil((1.6 * mean + 24.0) / 16.0);
            if ((double)w.nbins * slot_cap * 48.0 > 1.5e9) slot_cap = 0;
        }
    }
    if (slot_cap > 0) {
        // the histogram is zeroed here (the row-sum warps of several items read a bin's count, none of them may clear it)
        CK(e->tw_hist.reserve(nbp * TRIAL_HIST_STRIDE), "cudaMalloc trial histogram");
        CK(cudaMemsetAsync(e->tw_hist.p, 0, nbp * TRIAL_HIST_STRIDE * sizeof(int), e->stream), "memset histogram");
        e->hist_dirty = true;
        CK(e->tw_items.reserve((ns >> 4) + (ns < nbp ? ns : nbp) + 1), "cudaMalloc trial items");
        w.item_tab = e->tw_items.p;
        CK(e->tw_pos.reserve(ns), "cudaMalloc trial work");          // overflow records
        CK(e->tw_a.reserve(ns), "cudaMalloc trial work");
        CK(e->tw_spos.reserve((size_t)w.nbins * slot_cap), "cudaMalloc trial slots");
        CK(e->tw_meta.reserve((size_t)w.nbins * slot_cap), "cudaMalloc trial slots");
        if (!e->tw_ovf.p) {
            // [0] overflow count, [1] work items, [2] claim counter of the row-sum warps: zero on entry, re-zeroed by the
            // combine kernel
            CK(e->tw_ovf.reserve(4), "cudaMalloc trial overflow");
            CK(cudaMemsetAsync(e->tw_ovf.p, 0, 4 * sizeof(int), e->stream), "memset trial counters");
        }
        w.totals = e->tw_ovf.p + 1;
        w.hist = e->tw_hist.p; w.site_pos = e->tw_pos.p; w.site_a = e->tw_a.p;
        w.sorted_pos = e->tw_spos.p; w.sorted_meta = e->tw_meta.p;
        w.slot_cap = slot_cap; w.ovf = e->tw_ovf.p;
        auto mark = [&]() { if (e->stage_timing && e->stage_n < 6) cudaEventRecord(e->stage_ev[e->stage_n++], e->stream); };
        const bool pdl = e->use_pdl && !e->stage_timing;
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        auto config = [&](int grid, int block) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = e->stream;
            cfg.attrs = pdl_attr; cfg.numAttrs = pdl ? 1 : 0;
            return cfg;
        };
        const BatchView bv = e->view();
        StencilLists xl;
        { int rc = stencil_lists(e, false, &xl); if (rc != MCPYB_OK) return rc; }
        mark();
        { cudaLaunchConfig_t cfg = config((T + 255) / 256, 256);
          CK(cudaLaunchKernelEx(&cfg, trial_sites_direct_kernel, bv, tb, w), "launch trial_sites_direct_kernel"); e->launches++; }
        mark(); mark(); mark();                                   // (no scan, no scatter)
        {
            const long long max_items = (long long)(ns >> 4) + (long long)(ns < nbp ? ns : nbp) + 1;
            const long long want = (max_items + RW_WARPS - 1) / RW_WARPS;
            const int grid = (int)(want < (long long)e->num_sms * e->rw_minb ? want : (long long)e->num_sms * e->rw_minb);
            cudaLaunchConfig_t cfg = config(grid, 32 * RW_WARPS);
            if (e->rw_minb == 4)
                CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_direct_kernel<4>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_direct_kernel");
            else if (e->rw_minb == 5)
                CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_direct_kernel<5>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_direct_kernel");
            else
                CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_direct_kernel<6>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_direct_kernel");
            e->launches++;
        }
        mark();
        { cudaLaunchConfig_t cfg = config(e->num_sms * 2, 32 * RW_WARPS);
          CK(cudaLaunchKernelEx(&cfg, lj_trial_overflow_kernel, bv, tb, w, e->lj, xl), "launch lj_trial_overflow_kernel"); e->launches++; }
        { cudaLaunchConfig_t cfg = config((T + 255) / 256, 256);
          CK(cudaLaunchKernelEx(&cfg, lj_trial_combine_kernel, bv, tb, w, e->lj, d_dE, d_pairs), "launch lj_trial_combine_kernel");
          e->launches++; }
        { cudaLaunchConfig_t cfg = config(e->num_sms * 4, 128);
          CK(cudaLaunchKernelEx(&cfg, lj_trial_combine_rest_kernel, bv, tb, w, e->lj, d_dE, d_pairs), "launch lj_trial_combine_rest_kernel");
          e->launches++; }
        mark();
        return MCPYB_OK;
    }
    // batches: counting sort of the sites by cell, then one CTA per (cell, <= RB_SITES sites)
    CK(e->tw_bins.reserve(2 * nbp), "cudaMalloc trial bins");
    {
        const size_t had = e->tw_hist.cap;
        CK(e->tw_hist.reserve(nbp * TRIAL_HIST_STRIDE), "cudaMalloc trial histogram");
        if (e->tw_hist.cap != had || e->hist_dirty)      // fresh histogram, or one the direct placement left behind:
            CK(cudaMemsetAsync(e->tw_hist.p, 0, e->tw_hist.cap * sizeof(int), e->stream), "memset histogram");   // zero once, the scan kernel keeps it zero
        e->hist_dirty = false;
    }
    {
        const size_t had = e->tw_pos.cap;
        CK(e->tw_pos.reserve(ns), "cudaMalloc trial work");
        if (e->tw_pos.cap != had)            // stamp 0 marks "never written"
            CK(cudaMemsetAsync(e->tw_pos.p, 0, e->tw_pos.cap * sizeof(double4), e->stream), "memset trial sites");
    }
    CK(e->tw_a.reserve(ns), "cudaMalloc trial work");
    w.hist = e->tw_hist.p; w.bin_count = e->tw_bins.p; w.item_start = e->tw_bins.p + nbp;
    w.site_pos = e->tw_pos.p; w.site_a = e->tw_a.p;
    e->tw_stamp = (e->tw_stamp % 0x3FFFFFF) + 1;
    w.stamp = e->tw_stamp;
    const size_t max_items = (ns >> w.item_shift) + (ns < nbp ? ns : nbp) + 1;
    CK(e->tw_items.reserve(max_items), "cudaMalloc trial items");
    if (e->use_warp_rows) {
        CK(e->tw_items4.reserve(2 * max_items), "cudaMalloc trial items");
        const size_t had = e->tw_chain.cap;
        const size_t scan_ctas = (size_t)(e->R > 0 ? e->R : 1) * ((e->max_cells + RW_SCAN_THREADS) / RW_SCAN_THREADS);
        CK(e->tw_chain.reserve(2 * scan_ctas), "cudaMalloc scan chain");
        if (e->tw_chain.cap != had)          // stamp 0 = "not published"
            CK(cudaMemsetAsync(e->tw_chain.p, 0, e->tw_chain.cap * sizeof(unsigned long long), e->stream), "memset scan chain");
        w.items = e->tw_items4.p; w.chain = e->tw_chain.p;
    }
    CK(e->tw_spos.reserve(ns), "cudaMalloc trial work");
    CK(e->tw_meta.reserve(ns), "cudaMalloc trial work");
    w.item_tab = e->tw_items.p; w.sorted_pos = e->tw_spos.p; w.sorted_meta = e->tw_meta.p;
    auto mark = [&]() { if (e->stage_timing && e->stage_n < 6) cudaEventRecord(e->stage_ev[e->stage_n++], e->stream); };
    // Programmatic dependent launch: each kernel of the pipeline may be scheduled while its predecessor
    // drains (every one of them starts with pdl_wait()), which takes the launch gaps out of the step.
    // Off while the stages are being timed one by one.
    const bool pdl = e->use_pdl && !e->stage_timing;
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    auto config = [&](int grid, int block, size_t smem) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;
        cfg.attrs = pdl_attr; cfg.numAttrs = pdl ? 1 : 0;
        return cfg;
    };
    const BatchView bv = e->view();
    StencilLists xl;
    {
        int rc = stencil_lists(e, false, &xl);
        if (rc != MCPYB_OK) return rc;
    }
    mark();
    {
        { cudaLaunchConfig_t cfg = config((T + 255) / 256, 256, 0);
          CK(cudaLaunchKernelEx(&cfg, trial_sites_kernel, bv, tb, w), "launch trial_sites_kernel"); e->launches++; }
        mark();
        if (e->use_warp_rows) {
            // (max_cells + 1 bins per replica so that the end marker of a full grid has a thread)
            const int cpr = (e->max_cells + RW_SCAN_THREADS) / RW_SCAN_THREADS;
            cudaLaunchConfig_t cfg = config(e->R * cpr, RW_SCAN_THREADS, 0);
            CK(cudaLaunchKernelEx(&cfg, trial_scan_items_kernel, bv, w, xl, cpr), "launch trial_scan_items_kernel"); e->launches++;
        } else {
            cudaLaunchConfig_t cfg = config(1, 1024, 0);
            CK(cudaLaunchKernelEx(&cfg, trial_scan_kernel, w), "launch trial_scan_kernel"); e->launches++;
        }
        mark();
        { cudaLaunchConfig_t cfg = config((nsites + 255) / 256, 256, 0);
          CK(cudaLaunchKernelEx(&cfg, trial_scatter_kernel, tb, w), "launch trial_scatter_kernel"); e->launches++; }
    }
    mark();
    const long long est = (long long)max_items * (w.nplanes / w.item_slices);
    if (e->use_warp_rows) {
        const long long want = ((long long)max_items + RW_WARPS - 1) / RW_WARPS;
        const int minb = e->fp32_pairs ? 4 : e->rw_minb;
        const int grid = (int)(want < (long long)e->num_sms * minb ? want : (long long)e->num_sms * minb);
        cudaLaunchConfig_t cfg = config(grid, 32 * RW_WARPS, 0);
        if (e->fp32_pairs)
            CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_warp_kernel<4, true>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_warp_kernel");
        else if (e->rw_minb == 4)
            CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_warp_kernel<4, false>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_warp_kernel");
        else if (e->rw_minb == 5)
            CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_warp_kernel<5, false>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_warp_kernel");
        else
            CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_warp_kernel<6, false>, bv, tb, w, e->lj, xl), "launch lj_trial_rows_warp_kernel");
        e->launches++;
    } else {
        const int rb = (int)(est < e->num_sms * RB_MIN_BLOCKS ? est : e->num_sms * RB_MIN_BLOCKS);
        cudaLaunchConfig_t cfg = config(rb, RB_THREADS, 0);
        CK(cudaLaunchKernelEx(&cfg, lj_trial_rows_block_kernel<RB_UNROLL, RB_MIN_BLOCKS>, bv, tb, w, e->lj, xl),
           "launch lj_trial_rows_block_kernel");
        e->launches++;
    }
    mark();
    { cudaLaunchConfig_t cfg = config((T + 255) / 256, 256, 0);
      CK(cudaLaunchKernelEx(&cfg, lj_trial_combine_kernel, bv, tb, w, e->lj, d_dE, d_pairs), "launch lj_trial_combine_kernel");
      e->launches++; }
    { cudaLaunchConfig_t cfg = config(e->num_sms * 4, 128, 0);
      CK(cudaLaunchKernelEx(&cfg, lj_trial_combine_rest_kernel, bv, tb, w, e->lj, d_dE, d_pairs), "launch lj_trial_combine_rest_kernel");
      e->launches++; }
    mark();
    return MCPYB_OK;
}

extern "C++" {
namespace {
int launch_trials_fwd(mcpyb_engine* e, int T, const int32_t* d_rep, const int32_t* d_old_off, const int32_t* d_old_idx,
                      const int32_t* d_new_off, int n_old_total, double* d_dE, int32_t* d_pairs) {
    return launch_trials(e, T, d_rep, d_old_off, d_old_idx, d_new_off, nullptr, nullptr, n_old_total, 0, d_dE, d_pairs);
}
}  // namespace
}  // extern "C++"

int mcpyb_trial_delta_e_dev(mcpyb_engine* e, int T, const int32_t* d_rep, const int32_t* d_old_off,
                            const int32_t* d_old_idx, const int32_t* d_new_off, const double* d_new_pos,
                            const int32_t* d_new_Z, int n_old_total, int n_new_total,
                            double* d_dE_out, int32_t* d_pairs_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch (call mcpyb_upload_* first)");
    if (T < 0 || (T > 0 && (!d_old_off || !d_new_off || !d_dE_out))) return e->fail(MCPYB_ERR_ARG, "bad trial arguments");
    if (e->pot_kind == POT_EMT && n_new_total > 0 && !d_new_Z) return e->fail(MCPYB_ERR_ARG, "EMT trials need new_Z");
    if (n_old_total < 0 || n_new_total < 0) return e->fail(MCPYB_ERR_ARG, "negative site counts");
    return launch_trials(e, T, d_rep, d_old_off, d_old_idx, d_new_off, d_new_pos, d_new_Z, n_old_total, n_new_total,
                         d_dE_out, d_pairs_out);
}

static bool wait_flag(volatile unsigned int* flag, unsigned int seq);

int mcpyb_trial_delta_e_host(mcpyb_engine* e, int T, const int32_t* rep, const int32_t* old_off,
                             const int32_t* old_idx, const int32_t* new_off, const double* new_pos,
                             const int32_t* new_Z, double* dE_out, int32_t* pairs_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch (call mcpyb_upload_* first)");
    if (T < 0) return e->fail(MCPYB_ERR_ARG, "T must be >= 0");
    if (T == 0) return MCPYB_OK;
    if (!old_off || !new_off || !dE_out) return e->fail(MCPYB_ERR_ARG, "old_off / new_off / dE_out are NULL");
    const int n_old = old_off[T], n_new = new_off[T];
    if (n_old < 0 || n_new < 0 || old_off[0] != 0 || new_off[0] != 0) return e->fail(MCPYB_ERR_ARG, "bad trial offsets");
    if ((n_old > 0 && !old_idx) || (n_new > 0 && !new_pos)) return e->fail(MCPYB_ERR_ARG, "old_idx / new_pos are NULL");
    if (e->pot_kind == POT_EMT && n_new > 0 && !new_Z) return e->fail(MCPYB_ERR_ARG, "EMT trials need new_Z");
    // Argument checks.  The LJ pipeline validates every trial on the device (trial_sites_kernel) and
    // reports the first offender in the status word that travels back with dE; the host loop it
    // replaces cost 330 us per 65 536 trials.  The EMT kernel still relies on the host check.
    const bool dev_checked = e->pot_kind == POT_LJ;
    for (int t = 0; t < T && !dev_checked; ++t) {
        const int r = rep ? rep[t] : 0;
        if (r < 0 || r >= e->R) return e->fail(MCPYB_ERR_ARG, "trial %d: replica %d out of range", t, r);
        const int ko = old_off[t + 1] - old_off[t], kn = new_off[t + 1] - new_off[t];
        if (ko < 0 || kn < 0 || ko > MCPYB_MAX_MOVED || kn > MCPYB_MAX_MOVED)
            return e->fail(MCPYB_ERR_ARG, "trial %d: moved-atom lists must hold 0..%d atoms", t, MCPYB_MAX_MOVED);
        const int n = e->atom_off_host[r + 1] - e->atom_off_host[r];
        for (int k = old_off[t]; k < old_off[t + 1]; ++k)
            if (old_idx[k] < 0 || old_idx[k] >= n) return e->fail(MCPYB_ERR_ARG, "trial %d: atom index %d out of range (n=%d)", t, old_idx[k], n);
        if (new_Z)
            for (int k = new_off[t]; k < new_off[t + 1]; ++k)
                if (new_Z[k] < 0 || new_Z[k] >= 120 || e->z_to_slot_host[new_Z[k]] < 0)
                    return e->fail(MCPYB_ERR_SPECIES, "trial %d: atomic number %d has no EMT parameters", t, new_Z[k]);
    }
    // pinned block: rep[T] | old_off[T+1] | new_off[T+1] | old_idx | new_Z | new_pos
    size_t o_rep = 0;
    size_t o_oo = align_up(o_rep + (size_t)T * 4, 16);
    size_t o_no = align_up(o_oo + (size_t)(T + 1) * 4, 16);
    size_t o_oi = align_up(o_no + (size_t)(T + 1) * 4, 16);
    size_t o_nz = align_up(o_oi + (size_t)n_old * 4, 16);
    size_t o_np = align_up(o_nz + (size_t)n_new * 4, 16);
    size_t bytes = align_up(o_np + (size_t)n_new * 24, 64);
    // every host trial call ends with a stream synchronisation, so the staging buffers are free here
    // unless a call returned early
    if (e->trial_staging_busy) {
        CK(cudaStreamSynchronize(e->stream), "sync before trial staging reuse");
        if (e->copy_stream) CK(cudaStreamSynchronize(e->copy_stream), "sync before trial staging reuse");
    }
    e->trial_staging_busy = true;
    CopyPool* pool = copy_pool_for(e, bytes);
    auto hcopy = [&](void* dst, const void* src, size_t nbytes) {
        if (pool) pool->copy(dst, src, nbytes); else memcpy(dst, src, nbytes);
    };
    CK(e->trial_pin.reserve(bytes), "cudaMallocHost trials");
    CK(e->trial_dev.reserve(bytes), "cudaMalloc trials");
    CK(e->dE.reserve((size_t)T + 1), "cudaMalloc dE");          // dE[T] carries the device status word
    CK(e->tpairs.reserve(T), "cudaMalloc trial pairs");
    unsigned char* h = e->trial_pin.p;
    unsigned char* d = e->trial_dev.p;
    const bool small = bytes <= (64u << 10);          // latency path: one H2D, one D2H
    if (!small && !e->copy_stream) {
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate (copy)");
        CK(cudaEventCreateWithFlags(&e->copy_ev, cudaEventDisableTiming), "cudaEventCreate (copy)");
    }
    cudaStream_t cs = small ? e->stream : e->copy_stream;
    // A batch that goes to the row-sum pipeline against a freshly uploaded configuration needs the per-configuration
    // stencil lists first (expand_stencils_kernel: ~150 us for 8 x 2 000 atoms).  They depend on the configurations
    // only: launched here, on the compute stream, they are built while the copy stream brings the proposals in.
    if (!small && e->pot_kind == POT_LJ && e->use_warp_rows && n_old + n_new > RS_MAX_SITES) {
        StencilLists ahead;
        const int rc_lists = stencil_lists(e, false, &ahead);
        if (rc_lists != MCPYB_OK) return rc_lists;
    }
    // stage piece by piece: the H2D copy of one piece runs while the host copies the next one
    auto stage = [&](size_t off, const void* src, size_t nbytes) -> cudaError_t {
        const size_t piece = 512u << 10;
        for (size_t done = 0; done < nbytes; done += piece) {
            const size_t nb = nbytes - done < piece ? nbytes - done : piece;
            hcopy(h + off + done, (const unsigned char*)src + done, nb);
            cudaError_t ce = cudaMemcpyAsync(d + off + done, h + off + done, nb, cudaMemcpyHostToDevice, cs);
            if (ce != cudaSuccess) return ce;
        }
        return cudaSuccess;
    };
    // Caller's arrays already page-locked (mcpyb_alloc_pinned, cudaHostRegister): the copy engine reads
    // them where they are -- no staging copy, which for a 65 536-trial batch is a third of the call.
    const bool direct = !small && is_pinned_host(old_off) && is_pinned_host(new_off) && (!rep || is_pinned_host(rep)) &&
                        (!n_old || is_pinned_host(old_idx)) && (!n_new || is_pinned_host(new_pos)) &&
                        (!(n_new && new_Z && e->pot_kind == POT_EMT) || is_pinned_host(new_Z));
    const auto tq0 = std::chrono::steady_clock::now();
    if (small) {
        if (rep) memcpy(h + o_rep, rep, (size_t)T * 4);
        memcpy(h + o_oo, old_off, (size_t)(T + 1) * 4);
        memcpy(h + o_no, new_off, (size_t)(T + 1) * 4);
        if (n_old) memcpy(h + o_oi, old_idx, (size_t)n_old * 4);
        if (n_new && new_Z) memcpy(h + o_nz, new_Z, (size_t)n_new * 4);
        if (n_new) memcpy(h + o_np, new_pos, (size_t)n_new * 24);
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, e->stream), "H2D trials");
    } else if (direct) {
        if (rep) CK(cudaMemcpyAsync(d + o_rep, rep, (size_t)T * 4, cudaMemcpyHostToDevice, cs), "H2D trials");
        CK(cudaMemcpyAsync(d + o_oo, old_off, (size_t)(T + 1) * 4, cudaMemcpyHostToDevice, cs), "H2D trials");
        CK(cudaMemcpyAsync(d + o_no, new_off, (size_t)(T + 1) * 4, cudaMemcpyHostToDevice, cs), "H2D trials");
        if (n_old) CK(cudaMemcpyAsync(d + o_oi, old_idx, (size_t)n_old * 4, cudaMemcpyHostToDevice, cs), "H2D trials");
        if (n_new && new_Z && e->pot_kind == POT_EMT)
            CK(cudaMemcpyAsync(d + o_nz, new_Z, (size_t)n_new * 4, cudaMemcpyHostToDevice, cs), "H2D trials");
        if (n_new) CK(cudaMemcpyAsync(d + o_np, new_pos, (size_t)n_new * 24, cudaMemcpyHostToDevice, cs), "H2D trials");
    } else {
        if (rep) CK(stage(o_rep, rep, (size_t)T * 4), "H2D trials");
        CK(stage(o_oo, old_off, (size_t)(T + 1) * 4), "H2D trials");
        CK(stage(o_no, new_off, (size_t)(T + 1) * 4), "H2D trials");
        if (n_old) CK(stage(o_oi, old_idx, (size_t)n_old * 4), "H2D trials");
        if (n_new && new_Z && e->pot_kind == POT_EMT) CK(stage(o_nz, new_Z, (size_t)n_new * 4), "H2D trials");
        if (n_new) CK(stage(o_np, new_pos, (size_t)n_new * 24), "H2D trials");
    }
    if (!small) {
        CK(cudaEventRecord(e->copy_ev, cs), "cudaEventRecord (copy)");
        CK(cudaStreamWaitEvent(e->stream, e->copy_ev, 0), "cudaStreamWaitEvent (copy)");
    }
    const size_t ob = (size_t)(T + 1) * sizeof(double), pb = (size_t)T * sizeof(int);
    const size_t o_flag = align_up(ob + pb, 8);
    CK(e->trial_out_pin.reserve(o_flag + 8), "cudaMallocHost trial out");
    // A batch the single-launch LJ kernel takes: its last CTA writes dE, the pair counts and the status word straight
    // into the pinned block and then a sequence word the host polls -- no D2H copies, no stream synchronisation
    // (~12 us of a ~45 us call for the 64-replica batch of BatchedReplicaExchange).
    const bool flagged = small && e->pot_kind == POT_LJ && n_old + n_new > 0 && n_old + n_new <= RS_MAX_SITES;
    volatile unsigned int* h_flag = (volatile unsigned int*)(e->trial_out_pin.p + o_flag);
    unsigned int seq = 0;
    if (flagged) { seq = ++e->tiny_seq; *h_flag = seq - 1; }
    const auto tq1 = std::chrono::steady_clock::now();
    int rc = launch_trials(e, T, rep ? (const int32_t*)(d + o_rep) : nullptr, (const int32_t*)(d + o_oo), (const int32_t*)(d + o_oi),
                           (const int32_t*)(d + o_no), (const double*)(d + o_np),
                           (new_Z && e->pot_kind == POT_EMT) ? (const int32_t*)(d + o_nz) : nullptr, n_old, n_new,
                           flagged ? (double*)e->trial_out_pin.p : e->dE.p,
                           flagged ? (pairs_out ? (int32_t*)(e->trial_out_pin.p + ob) : nullptr) : e->tpairs.p,
                           flagged ? (unsigned long long*)(e->trial_out_pin.p + (size_t)T * sizeof(double)) : (unsigned long long*)(e->dE.p + T),
                           flagged ? h_flag : nullptr, seq);
    if (rc != MCPYB_OK) return rc;
    // status word first, then dE in two halves: the host copies half 1 out while half 2 is in flight
    const size_t half = ((size_t)T / 2) * sizeof(double), full = (size_t)T * sizeof(double);
    const bool out_direct = !small && is_pinned_host(dE_out) && (!pairs_out || is_pinned_host(pairs_out));
    const auto tq2 = std::chrono::steady_clock::now();
    if (flagged) {
        if (!wait_flag(h_flag, seq)) CK(cudaStreamSynchronize(e->stream), "sync trials");
        if (e->trkb_timing) {
            const auto tq3 = std::chrono::steady_clock::now();
            e->trkb_t_stage += std::chrono::duration<double>(tq1 - tq0).count();
            e->trkb_t_launch += std::chrono::duration<double>(tq2 - tq1).count();
            e->trkb_t_wait += std::chrono::duration<double>(tq3 - tq2).count();
        }
    } else if (small) {
        CK(cudaMemcpyAsync(e->trial_out_pin.p, e->dE.p, ob, cudaMemcpyDeviceToHost, e->stream), "D2H dE");
        if (pairs_out) CK(cudaMemcpyAsync(e->trial_out_pin.p + ob, e->tpairs.p, pb, cudaMemcpyDeviceToHost, e->stream), "D2H pairs");
        CK(cudaStreamSynchronize(e->stream), "sync trials");
    } else if (out_direct) {
        // page-locked result arrays: dE lands in the caller's buffer, only the status word is staged
        CK(cudaMemcpyAsync(e->trial_out_pin.p + full, e->dE.p + T, sizeof(double), cudaMemcpyDeviceToHost, e->stream), "D2H status");
        CK(cudaMemcpyAsync(dE_out, e->dE.p, full, cudaMemcpyDeviceToHost, e->stream), "D2H dE");
        if (pairs_out) CK(cudaMemcpyAsync(pairs_out, e->tpairs.p, pb, cudaMemcpyDeviceToHost, e->stream), "D2H pairs");
        CK(cudaStreamSynchronize(e->stream), "sync trials");
    } else {
        if (!e->trial_ev) CK(cudaEventCreateWithFlags(&e->trial_ev, cudaEventDisableTiming), "cudaEventCreate");
        CK(cudaMemcpyAsync(e->trial_out_pin.p + full, e->dE.p + T, sizeof(double), cudaMemcpyDeviceToHost, e->stream), "D2H status");
        CK(cudaMemcpyAsync(e->trial_out_pin.p, e->dE.p, half, cudaMemcpyDeviceToHost, e->stream), "D2H dE");
        CK(cudaEventRecord(e->trial_ev, e->stream), "cudaEventRecord");
        CK(cudaMemcpyAsync(e->trial_out_pin.p + half, (const unsigned char*)e->dE.p + half, full - half, cudaMemcpyDeviceToHost, e->stream), "D2H dE");
        if (pairs_out) CK(cudaMemcpyAsync(e->trial_out_pin.p + ob, e->tpairs.p, pb, cudaMemcpyDeviceToHost, e->stream), "D2H pairs");
        CK(cudaEventSynchronize(e->trial_ev), "sync trials (first half)");
    }
    if (dev_checked) {
        unsigned long long st;
        memcpy(&st, e->trial_out_pin.p + (size_t)T * sizeof(double), sizeof st);
        if (st != ~0ULL) {
            const int t = (int)(st >> 8), code = (int)(st & 0xFF);
            if (code == 1) return e->fail(MCPYB_ERR_ARG, "trial %d: replica %d out of range", t, rep ? rep[t] : 0);
            if (code == 2) return e->fail(MCPYB_ERR_ARG, "trial %d: moved-atom lists must hold 0..%d atoms (offsets non-decreasing)", t, MCPYB_MAX_MOVED);
            return e->fail(MCPYB_ERR_ARG, "trial %d: atom index out of range", t);
        }
    }
    if (!out_direct) {
        hcopy(dE_out, e->trial_out_pin.p, half);
        if (!small) CK(cudaStreamSynchronize(e->stream), "sync trials");
        hcopy((unsigned char*)dE_out + half, e->trial_out_pin.p + half, full - half);
        if (pairs_out) memcpy(pairs_out, e->trial_out_pin.p + ob, pb);
    }
    e->trial_staging_busy = false;
    return MCPYB_OK;
}

// Completion of a single-launch trial kernel: its last thread writes `seq` to a word in pinned host
// memory after dE, and the host polls that word instead of paying for the kernel's retirement and a
// stream synchronisation (~5 us of a ~35 us call).  Falls back to the synchronisation after 2 ms (a
// faulted kernel never writes the word; the synchronisation then reports the error).
static bool wait_flag(volatile unsigned int* flag, unsigned int seq) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned it = 1;; ++it) {
        if (*flag == seq) return true;
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((it & 0xFFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) return false;
    }
}

// ---- resident chain -------------------------------------------------------------------
// Post one trial against the resident base to the chain CTA and wait for its dE.  *answered stays false when the
// CTA cannot serve it (then the caller takes the one-launch path).
static int chain_launch(mcpyb_engine* e, const StencilLists& xs, unsigned int start_seq) {
    if (!e->chain_stream) {
        CK(cudaStreamCreateWithFlags(&e->chain_stream, cudaStreamNonBlocking), "cudaStreamCreate chain");
        CK(cudaEventCreateWithFlags(&e->chain_ev, cudaEventDisableTiming), "cudaEventCreate chain");
        CK(cudaHostAlloc((void**)&e->chain_mb, sizeof(ChainMailbox), cudaHostAllocDefault), "cudaHostAlloc mailbox");
        memset(e->chain_mb, 0, sizeof(ChainMailbox));
    }
    // everything the CTA reads was produced on the engine's stream
    CK(cudaEventRecord(e->chain_ev, e->stream), "record chain dependency");
    CK(cudaStreamWaitEvent(e->chain_stream, e->chain_ev, 0), "chain waits for the base");
    e->chain_mb->alive = 1;
    e->chain_mb->stop = 0;
    if (e->pot_kind == POT_LJ) {
        lj_chain_kernel<<<1, CHAIN_THREADS, 0, e->chain_stream>>>(e->view(), e->lj, xs, e->chain_mb, start_seq);
        LAUNCH_CHECK("lj_chain_kernel");
    } else {
        emt_chain_kernel<<<1, 256, 0, e->chain_stream>>>(e->view(), e->emt_dev.p, e->z_to_slot.p, e->sigma1.p, e->e_atom.p,
                                                         e->chain_mb, start_seq);
        LAUNCH_CHECK("emt_chain_kernel");
    }
    e->chain_launched = true;
    e->chain_launches++;
    return MCPYB_OK;
}

static int chain_trial(mcpyb_engine* e, int n_old, const int32_t* old_idx, int n_new, const double* new_pos,
                       const int32_t* new_Z, double* dE_out, bool* answered) {
    *answered = false;
    const bool lj = e->pot_kind == POT_LJ;
    if (n_old > CHAIN_MAX_OLD || e->R != 1) return MCPYB_OK;
    if (lj ? (n_old + n_new > CHAIN_MAX_SITES) : (n_new > CHAIN_MAX_OLD)) return MCPYB_OK;
    StencilLists xs;
    memset(&xs, 0, sizeof xs);
    if (lj) {
        // the lists appear with the second call against a base: the CTA is only used from then on
        int rc = stencil_lists(e, true, &xs);
        if (rc != MCPYB_OK) return rc;
        if (!xs.head) return MCPYB_OK;
    } else if (!e->emt_state_valid) {
        int rc = launch_energies(e);               // sigma1 / e_atom of the resident base
        if (rc != MCPYB_OK) return rc;
    }
    if (!e->chain_mb) { int rc = chain_launch(e, xs, e->chain_seq); if (rc != MCPYB_OK) return rc; }
    ChainMailbox* mb = e->chain_mb;
    const unsigned int seq = ++e->chain_seq;
    // LJ: positions of the removed atoms (from the host mirror of the base), then of the added ones; EMT: added only
    const int nsites = lj ? n_old + n_new : n_new, used = 1 + (nsites + 1) / 2;
    mb->l0.n_old = n_old; mb->l0.n_new = n_new;
    for (int k = 0; k < n_old; ++k) mb->l0.old_idx[k] = old_idx[k];
    for (int k = 0; k < n_new && k < CHAIN_MAX_OLD; ++k) mb->l0.new_Z[k] = new_Z ? new_Z[k] : 0;
    for (int s = 0; s < nsites; ++s) {
        const double* src = !lj ? new_pos + 3 * s
                                : (s < n_old ? &e->trk_pos[3 * (size_t)old_idx[s]] : new_pos + 3 * (s - n_old));
        double* dst = mb->lp[s >> 1].pos[s & 1];
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
    std::atomic_thread_fence(std::memory_order_release);     // payload before the sequence words (x86: store order)
    ((volatile unsigned int*)&mb->l0.seq)[0] = seq;
    for (int k = 0; k + 1 < used; ++k) ((volatile unsigned int*)&mb->lp[k].seq)[0] = seq;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (!e->chain_launched || !mb->alive) {
            // not resident (first call, idle timeout, stopped): launch it; it starts from the request before this one
            if (e->chain_launched) { CK(cudaStreamSynchronize(e->chain_stream), "sync resident chain"); e->chain_launched = false; }
            int rc = chain_launch(e, xs, seq - 1);
            if (rc != MCPYB_OK) return rc;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned it = 1;; ++it) {
            const unsigned long long ss = mb->status_seq;
            if ((unsigned int)(ss & 0xFFFFFFFFULL) == seq) {
                std::atomic_thread_fence(std::memory_order_acquire);
                if ((unsigned int)(ss >> 32) != CHAIN_ST_OK) return MCPYB_OK;        // not for the CTA: the caller launches
                *dE_out = mb->dE;
                *answered = true;
                e->chain_trials++;
                return MCPYB_OK;
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((it & 0x3FF) == 0) {
                if (!mb->alive) break;                                   // the CTA left before it saw the request
                if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) {
                    cudaError_t st = cudaStreamQuery(e->chain_stream);
                    if (st != cudaSuccess && st != cudaErrorNotReady) return e->cuda_fail(st, "resident chain");
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(500)) {
                        chain_stop(e);
                        return MCPYB_OK;                                 // give up on the CTA: the caller launches
                    }
                }
            }
        }
    }
    chain_stop(e);
    return MCPYB_OK;
}

// ---- incremental drop-in path ------------------------------------------------------
// calculator.get_potential_energy(atoms) is called once per trial move on a configuration that
// differs from the previous ones by a handful of atoms (grand_canonical_ensemble.py:244-308).
// The engine keeps one fully evaluated BASE configuration resident.  Each call aligns the incoming
// configuration with the base on the host (rows equal within 1e-10 A and same species match; ASE's
// `del` removes rows, `+=` appends, displacements keep the row), and returns
// E_base + dE(base -> incoming) from the trial kernels.  There is no accept / reject inference and
// no drift: every answer is one full evaluation plus ONE delta.  When the difference grows beyond a
// few atoms (accepted moves accumulate) the incoming configuration becomes the new base.
extern "C++" {
template <class ZT>
static int tracked_energy_core(mcpyb_engine* e, const double* positions, const ZT* numbers, int n,
                               const double* cell9, const uint8_t* pbc3, double* energy_out, int* path_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (n < 0 || !cell9 || !pbc3 || !energy_out || (n > 0 && !positions)) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    if (e->pot_kind == POT_NONE) return e->fail(MCPYB_ERR_STATE, "no potential set");
    if (e->pot_kind == POT_EMT && n > 0 && !numbers) return e->fail(MCPYB_ERR_ARG, "EMT needs atomic numbers");
    const double tol = 1e-10;
    const int kRebase = 16;                      // moved atoms (old + new) beyond which we re-base
    auto zof = [&](int j) -> int32_t { return numbers ? (int32_t)numbers[j] : 0; };
    bool same_box = e->trk_valid && memcmp(e->trk_cell, cell9, sizeof e->trk_cell) == 0 &&
                    memcmp(e->trk_pbc, pbc3, 3) == 0 && e->have_batch && e->R == 1;
    int32_t old_idx[2 * MCPYB_MAX_MOVED];
    int new_row[2 * MCPYB_MAX_MOVED];
    int n_old = 0, n_new = 0;
    bool incremental = same_box;
    if (incremental) {
        const int nB = (int)e->trk_Z.size();
        const double* B = e->trk_pos.data();
        auto match = [&](int i, int j) {
            return e->trk_Z[i] == zof(j) && fabs(B[3 * i] - positions[3 * j]) <= tol &&
                   fabs(B[3 * i + 1] - positions[3 * j + 1]) <= tol && fabs(B[3 * i + 2] - positions[3 * j + 2]) <= tol;
        };
        int i = 0, j = 0;
        // unchanged rows are bit-identical copies of the base's: skip them a block at a time (the per-row
        // tolerance test below costs ~5 ns a row, 10 us of a 35 us call at 2 000 atoms)
        const int kBlock = 64;
        auto same_block = [&](int i0, int j0) {
            if (memcmp(B + 3 * (size_t)i0, positions + 3 * (size_t)j0, (size_t)kBlock * 24) != 0) return false;
            if (numbers && sizeof(ZT) == sizeof(int32_t)) return memcmp(e->trk_Z.data() + i0, numbers + j0, (size_t)kBlock * 4) == 0;
            if (numbers) {
                int diff = 0;
                for (int k = 0; k < kBlock; ++k) diff |= (long long)e->trk_Z[i0 + k] != (long long)numbers[j0 + k];
                return diff == 0;
            }
            for (int k = 0; k < kBlock; ++k) if (e->trk_Z[i0 + k] != 0) return false;
            return true;
        };
        while (incremental && i < nB && j < n) {
            if (i + kBlock <= nB && j + kBlock <= n && same_block(i, j)) { i += kBlock; j += kBlock; continue; }
            if (match(i, j)) { ++i; ++j; continue; }
            if (n_old + n_new + 2 > kRebase) { incremental = false; break; }
            if (i + 1 < nB && match(i + 1, j)) { old_idx[n_old++] = i++; continue; }          // row deleted
            if (j + 1 < n && match(i, j + 1)) { new_row[n_new++] = j++; continue; }           // row inserted
            old_idx[n_old++] = i++; new_row[n_new++] = j++;                                  // row displaced
        }
        while (incremental && i < nB) { if (n_old + n_new + 1 > kRebase) incremental = false; else old_idx[n_old++] = i++; }
        while (incremental && j < n) { if (n_old + n_new + 1 > kRebase) incremental = false; else new_row[n_new++] = j++; }
    }
    if (incremental && n_old == 0 && n_new == 0) {
        *energy_out = e->trk_E;
        if (path_out) *path_out = 2;
        return MCPYB_OK;
    }
    if (incremental) {
        int32_t old_off[2] = {0, n_old}, new_off[2] = {0, n_new};
        double new_pos[3 * 2 * MCPYB_MAX_MOVED];
        int32_t new_Z[2 * MCPYB_MAX_MOVED];
        for (int k = 0; k < n_new; ++k) {
            for (int c = 0; c < 3; ++c) new_pos[3 * k + c] = positions[3 * new_row[k] + c];
            new_Z[k] = zof(new_row[k]);
        }
        double dE = 0.0;
        const bool keep = e->trk_valid;
        int rc;
        bool answered = false;
        if (e->use_chain && n_old + n_new >= 1) {
            // the resident chain CTA (chain.cuh): the trial goes into the pinned mailbox, no launch
            bool species_ok = true;
            if (e->pot_kind == POT_EMT)
                for (int k = 0; k < n_new; ++k)
                    species_ok = species_ok && new_Z[k] >= 0 && new_Z[k] < 120 && e->z_to_slot_host[new_Z[k]] >= 0;
            if (species_ok) {
                rc = chain_trial(e, n_old, old_idx, n_new, new_pos, new_Z, &dE, &answered);
                if (rc != MCPYB_OK) return rc;
            }
        }
        if (answered) {
        } else if (e->pot_kind == POT_LJ && n_old <= TINY_MAX_MOVED && n_new <= TINY_MAX_MOVED) {
            // the trial rides in the kernel parameters, the result lands in pinned host memory:
            // one launch and one stream synchronisation per call
            TinyTrial tt;
            memset(&tt, 0, sizeof tt);
            tt.old_off[1] = n_old; tt.new_off[1] = n_new;
            for (int k = 0; k < n_old; ++k) tt.old_idx[k] = old_idx[k];
            for (int k = 0; k < 3 * n_new; ++k) tt.new_pos[k] = new_pos[k];
            CK(e->trial_out_pin.reserve(64), "cudaMallocHost trial out");
            double* h_dE = (double*)e->trial_out_pin.p;
            unsigned long long* h_status = (unsigned long long*)(e->trial_out_pin.p + 8);
            TrialWork w;
            rc = prepare_rows_work(e, n_old + n_new, n_old, n_new, h_status, &w);
            if (rc != MCPYB_OK) return rc;
            volatile unsigned int* h_flag = (volatile unsigned int*)(e->trial_out_pin.p + 16);
            const unsigned int seq = ++e->tiny_seq;
            *h_flag = seq - 1;                        // whatever the (possibly fresh) pinned block held
            StencilLists xs;
            rc = stencil_lists(e, true, &xs);
            if (rc != MCPYB_OK) return rc;
            lj_trial_tiny_kernel<<<n_old + n_new, RS_THREADS, 0, e->stream>>>(e->view(), tt, w, e->lj, h_dE,
                                                                             (int*)(e->tw_err.p + 2), xs, h_flag, seq);
            LAUNCH_CHECK("lj_trial_tiny_kernel");
            if (!wait_flag(h_flag, seq)) CK(cudaStreamSynchronize(e->stream), "sync tiny trial");
            if (*h_status != ~0ULL) return e->fail(MCPYB_ERR_ARG, "tracked trial rejected on the device (status %llu)", *h_status);
            dE = *h_dE;
        } else if (e->pot_kind == POT_EMT && n_old <= EMT_TINY_MAX_MOVED && n_new <= EMT_TINY_MAX_MOVED) {
            bool species_ok = true;
            for (int k = 0; k < n_new; ++k)
                species_ok = species_ok && new_Z[k] >= 0 && new_Z[k] < 120 && e->z_to_slot_host[new_Z[k]] >= 0;
            if (!species_ok) return e->fail(MCPYB_ERR_SPECIES, "an inserted atom has no EMT parameters");
            EmtTinyTrial tt;
            memset(&tt, 0, sizeof tt);
            tt.old_off[1] = n_old; tt.new_off[1] = n_new;
            for (int k = 0; k < n_old; ++k) tt.old_idx[k] = old_idx[k];
            for (int k = 0; k < n_new; ++k) tt.new_Z[k] = new_Z[k];
            for (int k = 0; k < 3 * n_new; ++k) tt.new_pos[k] = new_pos[k];
            CK(e->trial_out_pin.reserve(64), "cudaMallocHost trial out");
            double* h_dE = (double*)e->trial_out_pin.p;
            if (!e->emt_state_valid) {
                rc = launch_energies(e);           // sigma1 / e_atom of the resident base
                if (rc != MCPYB_OK) return rc;
            }
            volatile unsigned int* h_flag = (volatile unsigned int*)(e->trial_out_pin.p + 16);
            const unsigned int seq = ++e->tiny_seq;
            *h_flag = seq - 1;                        // whatever the (possibly fresh) pinned block held
            emt_trial_tiny_kernel<<<1, 256, 0, e->stream>>>(e->view(), tt, e->emt_dev.p, e->z_to_slot.p,
                                                            e->sigma1.p, e->e_atom.p, h_dE, h_flag, seq);
            LAUNCH_CHECK("emt_trial_tiny_kernel");
            if (!wait_flag(h_flag, seq)) CK(cudaStreamSynchronize(e->stream), "sync tiny trial");
            dE = *h_dE;
        } else {
            rc = mcpyb_trial_delta_e_host(e, 1, nullptr, old_off, old_idx, new_off, new_pos,
                                          numbers ? new_Z : nullptr, &dE, nullptr);
            e->trk_valid = keep;
            if (rc != MCPYB_OK) return rc;
        }
        e->trk_valid = keep;
        if (isfinite(dE) || e->pot_kind == POT_LJ) {      // EMT reports NaN when the box is too thin
            *energy_out = e->trk_E + dE;
            if (path_out) *path_out = 1;
            e->trk_incremental++;
            return MCPYB_OK;
        }
    }
    // full evaluation; the incoming configuration becomes the base
    const int64_t off[2] = {0, n};
    std::vector<int32_t> z32;
    const int32_t* numbers32 = nullptr;
    if (numbers) {
        if (sizeof(ZT) == sizeof(int32_t)) numbers32 = (const int32_t*)numbers;
        else { z32.resize((size_t)n); for (int j = 0; j < n; ++j) z32[j] = (int32_t)numbers[j]; numbers32 = z32.data(); }
    }
    int rc = upload_host_impl(e, e->subdiv_trials, 1, off, positions, numbers32, cell9, pbc3);
    if (rc != MCPYB_OK) return rc;
    double E = 0.0;
    rc = mcpyb_energies_host(e, &E, nullptr);
    if (rc != MCPYB_OK) return rc;
    e->trk_pos.assign(positions, positions + 3 * (size_t)n);
    e->trk_Z.resize((size_t)n);
    for (int j = 0; j < n; ++j) e->trk_Z[j] = zof(j);
    memcpy(e->trk_cell, cell9, sizeof e->trk_cell);
    memcpy(e->trk_pbc, pbc3, 3);
    e->trk_E = E;
    e->trk_valid = true;
    e->trk_full++;
    *energy_out = E;
    if (path_out) *path_out = 0;
    return MCPYB_OK;
}
}  // extern "C++"

int mcpyb_tracked_energy_host(mcpyb_engine* e, const double* positions, const int32_t* numbers, int n,
                              const double* cell9, const uint8_t* pbc3, double* energy_out, int* path_out) {
    return tracked_energy_core<int32_t>(e, positions, numbers, n, cell9, pbc3, energy_out, path_out);
}

// The same call with the atomic numbers as ase.Atoms keeps them (int64): no conversion pass on the caller's side.
int mcpyb_tracked_energy_z64_host(mcpyb_engine* e, const double* positions, const int64_t* numbers, int n,
                                  const double* cell9, const uint8_t* pbc3, double* energy_out, int* path_out) {
    return tracked_energy_core<int64_t>(e, positions, numbers, n, cell9, pbc3, energy_out, path_out);
}

// Align an incoming configuration P with a base B: rows equal within 1e-10 A and of the same species
// match; unmatched base rows are removed, unmatched incoming rows added (any such description gives
// the right energy difference).  Returns false when more than `limit` atoms differ.
extern "C++" {
template <class ZT>
static bool align_rows(const double* B, const int32_t* BZ, int nB, const double* P, const ZT* PZ, int n,
                       int limit, int32_t* old_idx, int* new_row, int& n_old, int& n_new) {
    const double tol = 1e-10;
    auto match = [&](int i, int j) {
        return (long long)BZ[i] == (PZ ? (long long)PZ[j] : 0LL) && fabs(B[3 * i] - P[3 * j]) <= tol &&
               fabs(B[3 * i + 1] - P[3 * j + 1]) <= tol && fabs(B[3 * i + 2] - P[3 * j + 2]) <= tol;
    };
    n_old = n_new = 0;
    int i = 0, j = 0;
    while (i < nB && j < n) {
        // unchanged stretches are bit-identical almost always: compare 64 rows at a time
        {
            const int run = (nB - i < n - j ? nB - i : n - j);
            int k = 0;
            auto same_species = [&](int i0, int j0) {
                if (!PZ) return true;
                if (sizeof(ZT) == sizeof(int32_t)) return memcmp(BZ + i0, PZ + j0, 64 * sizeof(int32_t)) == 0;
                for (int q = 0; q < 64; ++q) if ((long long)BZ[i0 + q] != (long long)PZ[j0 + q]) return false;
                return true;
            };
            while (k + 64 <= run && memcmp(B + 3 * (size_t)(i + k), P + 3 * (size_t)(j + k), 64 * 24) == 0 &&
                   same_species(i + k, j + k)) k += 64;
            i += k; j += k;                       // (without numbers every stored species is 0 as well)
            if (i >= nB || j >= n) break;
        }
        if (match(i, j)) { ++i; ++j; continue; }
        if (n_old + n_new + 2 > limit) return false;
        if (i + 1 < nB && match(i + 1, j)) { old_idx[n_old++] = i++; continue; }          // row deleted
        if (j + 1 < n && match(i, j + 1)) { new_row[n_new++] = j++; continue; }           // row inserted
        old_idx[n_old++] = i++; new_row[n_new++] = j++;                                  // row displaced
    }
    while (i < nB) { if (n_old + n_new + 1 > limit) return false; old_idx[n_old++] = i++; }
    while (j < n) { if (n_old + n_new + 1 > limit) return false; new_row[n_new++] = j++; }
    return true;
}

// Batched form of mcpyb_tracked_energy_host for get_potential_energies(atoms_list)
// (batched_replica_exchange.py:277-278): the engine keeps R fully evaluated base configurations
// resident and answers each incoming configuration with E_base + dE(base -> incoming) from ONE batched
// trial launch.  An incoming configuration is matched with the base in its own slot first, then with any
// other base of similar size (replica exchange swaps configurations between slots).  When some
// configuration differs from every base by more than a few atoms, the whole incoming batch is
// evaluated in full and becomes the new set of bases.  paths_out[r] (may be NULL): 0 full,
// 1 incremental, 2 identical to its base.
// Core of the two entry points below.  Configuration r is n_of[r] rows at pos_of[r] (positions) and num_of[r]
// (atomic numbers, int32 or int64 -- ASE keeps int64; may be NULL for LJ): either slices of one concatenated
// array or the caller's own per-configuration arrays, read where they are.
template <class ZT>
static int tracked_energies_core(mcpyb_engine* e, int R, const double* const* pos_of, const ZT* const* num_of,
                                 const int64_t* n_of, const double* cells, const uint8_t* pbc,
                                 double* energies_out, int* paths_out) {
    const int kLimit = 16;                       // moved atoms (old + new) per configuration before a re-base
    const int Rb = e->trkb_valid ? (int)e->trkb_E.size() : 0;
    bool incremental = e->trkb_valid && e->have_batch && e->R == Rb && Rb > 0;
    std::vector<int32_t> t_rep, t_old_off(1, 0), t_new_off(1, 0), t_old_idx, t_new_Z;
    std::vector<double> t_new_pos;
    std::vector<int> slot_of(R, -1), trial_of(R, -1);
    const bool have_numbers = num_of != nullptr;
    // Every configuration against the base in its OWN slot first -- the usual case, independent per
    // configuration and memory-bound on the host (64 x 2 000 atoms: 6 MB touched), so spread over the helper
    // threads; the ones that do not match there are then tried against the other bases one by one.
    struct Own {
        int32_t old_idx[kLimit];
        int new_row[kLimit];
        int n_old = 0, n_new = 0;
        bool ok = false;
    };
    std::vector<Own> own(incremental ? (size_t)R : 0);
    const auto tk0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    auto try_slot = [&](int r, int q, int32_t* old_idx, int* new_row, int& n_old, int& n_new) -> bool {
        if (q >= Rb) return false;
        const int n = (int)n_of[r];
        const int nB = (int)(e->trkb_off[q + 1] - e->trkb_off[q]);
        if (abs(nB - n) > kLimit) return false;
        if (memcmp(&e->trkb_cell[9 * (size_t)q], cells + 9 * (size_t)r, 9 * sizeof(double)) != 0 ||
            memcmp(&e->trkb_pbc[3 * (size_t)q], pbc + 3 * (size_t)r, 3) != 0) return false;
        return align_rows<ZT>(&e->trkb_pos[3 * (size_t)e->trkb_off[q]], &e->trkb_Z[(size_t)e->trkb_off[q]], nB,
                              pos_of[r], have_numbers ? num_of[r] : nullptr, n, kLimit, old_idx, new_row, n_old, n_new);
    };
    if (incremental) {
        struct Ctx { decltype(try_slot)* f; std::vector<Own>* own; } ctx{&try_slot, &own};
        auto body = [](void* c, int r) {
            Ctx& x = *(Ctx*)c;
            Own& o = (*x.own)[r];
            o.ok = (*x.f)(r, r, o.old_idx, o.new_row, o.n_old, o.n_new);
        };
        size_t touched = 0;
        for (int r = 0; r < R; ++r) touched += (size_t)n_of[r] * 48;
        CopyPool* pool = copy_pool_for(e, touched);
        if (pool) pool->parallel_for(R, body, &ctx); else for (int r = 0; r < R; ++r) body(&ctx, r);
    }
    const double t_cmp = since(tk0);
    const auto tk1 = std::chrono::steady_clock::now();
    if (incremental) {
        int32_t old_idx_buf[kLimit];
        int new_row_buf[kLimit];
        for (int r = 0; r < R && incremental; ++r) {
            const double* P = pos_of[r];
            const ZT* PZ = have_numbers ? num_of[r] : nullptr;
            int n_old = own[r].n_old, n_new = own[r].n_new, found = own[r].ok ? r : -1;
            const int32_t* old_idx = own[r].old_idx;
            const int* new_row = own[r].new_row;
            for (int q = 0; q < Rb && found < 0; ++q) {                   // another slot's base (replica exchange)
                if (q == r) continue;
                if (try_slot(r, q, old_idx_buf, new_row_buf, n_old, n_new)) { found = q; old_idx = old_idx_buf; new_row = new_row_buf; }
            }
            if (found < 0) { incremental = false; break; }
            slot_of[r] = found;
            if (n_old == 0 && n_new == 0) continue;                       // identical to its base
            trial_of[r] = (int)t_rep.size();
            t_rep.push_back(found);
            for (int k = 0; k < n_old; ++k) t_old_idx.push_back(old_idx[k]);
            for (int k = 0; k < n_new; ++k) {
                for (int c = 0; c < 3; ++c) t_new_pos.push_back(P[3 * new_row[k] + c]);
                t_new_Z.push_back(PZ ? (int32_t)PZ[new_row[k]] : 0);
            }
            t_old_off.push_back((int32_t)t_old_idx.size());
            t_new_off.push_back((int32_t)t_new_Z.size());
        }
    }
    if (incremental) {
        const int T = (int)t_rep.size();
        std::vector<double> dE((size_t)(T > 0 ? T : 1), 0.0);
        bool finite = true;
        const double t_build = since(tk1);
        const auto tk2 = std::chrono::steady_clock::now();
        if (T > 0) {
            const bool keep = e->trkb_valid;
            int rc = mcpyb_trial_delta_e_host(e, T, t_rep.data(), t_old_off.data(), t_old_idx.data(), t_new_off.data(),
                                              t_new_pos.data(), have_numbers ? t_new_Z.data() : nullptr, dE.data(), nullptr);
            e->trkb_valid = keep;
            if (rc != MCPYB_OK) return rc;
            if (e->pot_kind == POT_EMT) for (int t = 0; t < T; ++t) finite = finite && isfinite(dE[t]);   // thin box -> NaN
        }
        if (finite) {
            for (int r = 0; r < R; ++r) {
                energies_out[r] = e->trkb_E[slot_of[r]] + (trial_of[r] >= 0 ? dE[trial_of[r]] : 0.0);
                if (paths_out) paths_out[r] = trial_of[r] >= 0 ? 1 : 2;
            }
            e->trk_incremental += T;
            if (e->trkb_timing) { e->trkb_t_compare += t_cmp; e->trkb_t_build += t_build; e->trkb_t_trial += since(tk2); e->trkb_calls++; }
            return MCPYB_OK;
        }
    }
    // full evaluation; the incoming batch becomes the set of bases (gathered straight into their host mirror)
    std::vector<int64_t> off((size_t)R + 1, 0);
    for (int r = 0; r < R; ++r) off[r + 1] = off[r] + n_of[r];
    const size_t total = (size_t)off[R];
    std::vector<double> pos_all(3 * total);
    std::vector<int32_t> z_all(total, 0);
    for (int r = 0; r < R; ++r) {
        if (n_of[r]) memcpy(&pos_all[3 * (size_t)off[r]], pos_of[r], (size_t)n_of[r] * 24);
        if (have_numbers) for (int64_t k = 0; k < n_of[r]; ++k) z_all[(size_t)(off[r] + k)] = (int32_t)num_of[r][k];
    }
    int rc = mcpyb_potential_energies_host(e, R, off.data(), pos_all.data(), have_numbers ? z_all.data() : nullptr,
                                           cells, pbc, energies_out);
    if (rc != MCPYB_OK) return rc;
    e->trkb_pos.swap(pos_all);
    e->trkb_Z.swap(z_all);
    e->trkb_off.swap(off);
    e->trkb_cell.assign(cells, cells + 9 * (size_t)R);
    e->trkb_pbc.assign(pbc, pbc + 3 * (size_t)R);
    e->trkb_E.assign(energies_out, energies_out + R);
    e->trkb_valid = true;
    e->trk_full += R;
    if (paths_out) for (int r = 0; r < R; ++r) paths_out[r] = 0;
    return MCPYB_OK;
}

}  // extern "C++"

int mcpyb_tracked_energies_host(mcpyb_engine* e, int R, const int64_t* atom_off, const double* positions,
                                const int32_t* numbers, const double* cells, const uint8_t* pbc,
                                double* energies_out, int* paths_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (R < 0 || (R > 0 && (!atom_off || !cells || !pbc || !energies_out))) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    if (e->pot_kind == POT_NONE) return e->fail(MCPYB_ERR_STATE, "no potential set");
    if (R == 0) return MCPYB_OK;
    std::vector<const double*> pos_of((size_t)R);
    std::vector<const int32_t*> num_of((size_t)R);
    std::vector<int64_t> n_of((size_t)R);
    for (int r = 0; r < R; ++r) {
        pos_of[r] = positions + 3 * atom_off[r];
        num_of[r] = numbers ? numbers + atom_off[r] : nullptr;
        n_of[r] = atom_off[r + 1] - atom_off[r];
        if (n_of[r] < 0) return e->fail(MCPYB_ERR_ARG, "atom_off must be non-decreasing");
    }
    return tracked_energies_core<int32_t>(e, R, pos_of.data(), numbers ? num_of.data() : nullptr, n_of.data(), cells, pbc,
                                          energies_out, paths_out);
}

int mcpyb_tracked_energies_rows_host(mcpyb_engine* e, int R, const double* const* positions_of,
                                     const void* const* numbers_of, int numbers_are_int64, const int64_t* n_atoms,
                                     const double* cells, const uint8_t* pbc, double* energies_out, int* paths_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (R < 0 || (R > 0 && (!positions_of || !n_atoms || !cells || !pbc || !energies_out))) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    if (e->pot_kind == POT_NONE) return e->fail(MCPYB_ERR_STATE, "no potential set");
    if (R == 0) return MCPYB_OK;
    for (int r = 0; r < R; ++r) {
        if (n_atoms[r] < 0 || (n_atoms[r] > 0 && (!positions_of[r] || (numbers_of && !numbers_of[r]))))
            return e->fail(MCPYB_ERR_ARG, "configuration %d: bad row pointers", r);
    }
    if (e->pot_kind == POT_EMT && !numbers_of) return e->fail(MCPYB_ERR_ARG, "EMT needs atomic numbers");
    if (numbers_are_int64)
        return tracked_energies_core<int64_t>(e, R, positions_of, (const int64_t* const*)numbers_of, n_atoms, cells, pbc,
                                              energies_out, paths_out);
    return tracked_energies_core<int32_t>(e, R, positions_of, (const int32_t* const*)numbers_of, n_atoms, cells, pbc,
                                          energies_out, paths_out);
}

// ---- neighbour pairs -------------------------------------------------------------
int mcpyb_neighbor_pairs_host(mcpyb_engine* e, int rep, double cutoff, int mode, int64_t max_pairs,
                              int32_t* i_out, int32_t* j_out, int32_t* S_out, double* r2_out, int64_t* npairs) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch");
    if (rep < 0 || rep >= e->R) return e->fail(MCPYB_ERR_ARG, "replica out of range");
    if (!(cutoff > 0.0) || cutoff > e->rcut * (1.0 + 1e-12)) return e->fail(MCPYB_ERR_ARG, "cutoff must be in (0, mcpyb_cutoff()]");
    if (mode < 0 || mode > 2 || max_pairs < 0 || !npairs) return e->fail(MCPYB_ERR_ARG, "bad arguments");
    const size_t cap = (size_t)(max_pairs > 0 ? max_pairs : 1);
    CK(e->nl_i.reserve(cap), "cudaMalloc nl"); CK(e->nl_j.reserve(cap), "cudaMalloc nl");
    CK(e->nl_S.reserve(3 * cap), "cudaMalloc nl"); CK(e->nl_r2.reserve(cap), "cudaMalloc nl");
    CK(e->nl_cursor.reserve(1), "cudaMalloc nl");
    CK(cudaMemsetAsync(e->nl_cursor.p, 0, sizeof(unsigned long long), e->stream), "memset cursor");
    const int n = e->atom_off_host[rep + 1] - e->atom_off_host[rep];
    if (n > 0) {
        int blocks = (n + 7) / 8;
        if (blocks > 148 * 16) blocks = 148 * 16;
        neighbor_pairs_kernel<<<blocks, 256, 0, e->stream>>>(e->view(), rep, cutoff, mode, (long long)max_pairs,
            e->nl_i.p, e->nl_j.p, e->nl_S.p, e->nl_r2.p, e->nl_cursor.p);
        LAUNCH_CHECK("neighbor_pairs_kernel");
    }
    unsigned long long found = 0;
    CK(cudaMemcpyAsync(&found, e->nl_cursor.p, sizeof found, cudaMemcpyDeviceToHost, e->stream), "D2H cursor");
    CK(cudaStreamSynchronize(e->stream), "sync pairs");
    *npairs = (int64_t)found;
    const size_t w = (size_t)((int64_t)found < max_pairs ? (int64_t)found : max_pairs);
    if (w) {
        CK(cudaMemcpy(i_out, e->nl_i.p, w * 4, cudaMemcpyDeviceToHost), "D2H i");
        CK(cudaMemcpy(j_out, e->nl_j.p, w * 4, cudaMemcpyDeviceToHost), "D2H j");
        CK(cudaMemcpy(S_out, e->nl_S.p, w * 12, cudaMemcpyDeviceToHost), "D2H S");
        if (r2_out) CK(cudaMemcpy(r2_out, e->nl_r2.p, w * 8, cudaMemcpyDeviceToHost), "D2H r2");
    }
    return MCPYB_OK;
}

// ---- free volume -----------------------------------------------------------------
// Bin the exclusion spheres (one CTA); afterwards fv_grid / fv_spheres / fv_cell_start are valid.
static int fv_build(mcpyb_engine* e, const double* d_pos, const double* d_radii, int64_t M,
                    const double* offset3, const double* shifts, int nshift) {
    if (nshift < 1 || nshift > 27) return e->fail(MCPYB_ERR_ARG, "nshift must be in 1..27");
    if (M * nshift > 0x7FFFFFF0LL) return e->fail(MCPYB_ERR_CAPACITY, "too many exclusion spheres");
    const int fv_max_cells = 1 << 18;
    const size_t nsph = (size_t)(M * nshift > 0 ? M * nshift : 1);
    CK(e->fv_grid.reserve(1), "cudaMalloc fv"); CK(e->fv_spheres.reserve(nsph), "cudaMalloc fv");
    CK(e->fv_cid.reserve(nsph), "cudaMalloc fv");
    CK(e->fv_cell_start.reserve(fv_max_cells + 1), "cudaMalloc fv");
    CK(e->fv_cursor.reserve(fv_max_cells), "cudaMalloc fv");
    if (!e->fv_small.p) {
        // offset(3) | shifts(81) | dims(9) | six zeros (offset and the single shift of the plain case)
        CK(e->fv_small.reserve(3 + 81 + 9 + 6), "cudaMalloc fv");
        CK(cudaMemsetAsync(e->fv_small.p, 0, e->fv_small.cap * sizeof(double), e->stream), "memset fv small");
    }
    const double* d_offset = e->fv_small.p + 93;
    const double* d_shifts = e->fv_small.p + 96;
    if (offset3 || shifts) {
        // offset + shifts: tiny pinned block
        CK(cudaStreamSynchronize(e->stream), "sync before fv staging reuse");
        CK(e->fv_pin.reserve((3 + 81 + 9) * sizeof(double)), "cudaMallocHost fv");
        double* hs = (double*)e->fv_pin.p;
        for (int i = 0; i < 3; ++i) hs[i] = offset3 ? offset3[i] : 0.0;
        for (int i = 0; i < 3 * nshift; ++i) hs[3 + i] = shifts ? shifts[i] : 0.0;
        CK(cudaMemcpyAsync(e->fv_small.p, hs, (3 + 3 * (size_t)nshift) * sizeof(double), cudaMemcpyHostToDevice, e->stream), "H2D fv small");
        d_offset = e->fv_small.p;
        d_shifts = e->fv_small.p + 3;
    }
    if (M * nshift >= 4096) {
        // large sphere sets: one 8-CTA thread-block cluster
        CK(e->bbox.reserve(64), "cudaMalloc fv partials");
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(FV_BUILD_THREADS); cfg.stream = e->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, fv_build_kernel<8>, d_pos, d_radii, (int)M, d_offset, d_shifts, nshift,
                              e->fv_grid.p, e->fv_spheres.p, (int)nsph, e->fv_cell_start.p, e->fv_cursor.p,
                              fv_max_cells, e->fv_cid.p, e->bbox.p), "launch fv build (cluster)");
        e->launches++;
    } else {
        fv_build_kernel<1><<<1, FV_BUILD_THREADS, 0, e->stream>>>(d_pos, d_radii, (int)M, d_offset, d_shifts, nshift,
            e->fv_grid.p, e->fv_spheres.p, (int)nsph, e->fv_cell_start.p, e->fv_cursor.p, fv_max_cells, e->fv_cid.p, nullptr);
        LAUNCH_CHECK("fv_build_kernel");
    }
    return MCPYB_OK;
}

static int fv_launch(mcpyb_engine* e, const double* d_points, int64_t P, const double* d_pos,
                     const double* d_radii, int64_t M, const double* offset3, const double* shifts,
                     int nshift, unsigned long long* d_counts, unsigned char* d_mask) {
    int rc = fv_build(e, d_pos, d_radii, M, offset3, shifts, nshift);
    if (rc != MCPYB_OK) return rc;
    CK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), e->stream), "memset counts");
    if (P > 0) {
        long long blocks = (P + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        fv_count_kernel<<<(int)blocks, 256, 0, e->stream>>>(d_points, P, e->fv_grid.p, e->fv_spheres.p,
                                                            e->fv_cell_start.p, d_counts, d_mask);
        LAUNCH_CHECK("fv_count_kernel");
    }
    return MCPYB_OK;
}

static int fv_upload_atoms(mcpyb_engine* e, const double* positions, const double* radii, int64_t M) {
    CK(e->fv_pos.reserve((size_t)(M > 0 ? 3 * M : 1)), "cudaMalloc fv pos");
    CK(e->fv_radii.reserve((size_t)(M > 0 ? M : 1)), "cudaMalloc fv radii");
    CK(e->fv_counts.reserve(2), "cudaMalloc fv counts");
    if (M) {
        // pinned staging: a pageable source makes cudaMemcpyAsync stage and block inside the driver
        CK(cudaStreamSynchronize(e->stream), "sync before fv atom staging reuse");
        CK(e->fv_atoms_pin.reserve((size_t)M * 32), "cudaMallocHost fv atoms");
        memcpy(e->fv_atoms_pin.p, positions, (size_t)M * 24);
        memcpy(e->fv_atoms_pin.p + (size_t)M * 24, radii, (size_t)M * 8);
        CK(cudaMemcpyAsync(e->fv_pos.p, e->fv_atoms_pin.p, (size_t)M * 24, cudaMemcpyHostToDevice, e->stream), "H2D fv pos");
        CK(cudaMemcpyAsync(e->fv_radii.p, e->fv_atoms_pin.p + (size_t)M * 24, (size_t)M * 8, cudaMemcpyHostToDevice, e->stream), "H2D fv radii");
    }
    return MCPYB_OK;
}

// counts and the grid's error flag come back through one pinned block (a pageable destination would
// make each copy synchronous); `ctl_out` additionally fetches the sphere sampler's control block
static int fv_read_counts(mcpyb_engine* e, int64_t* counts_out, SphereCtl* ctl_out = nullptr) {
    CK(e->fv_out_pin.reserve(256), "cudaMallocHost fv out");
    unsigned long long* c = (unsigned long long*)e->fv_out_pin.p;
    FvGrid* g = (FvGrid*)(e->fv_out_pin.p + 16);
    SphereCtl* ctl = (SphereCtl*)(e->fv_out_pin.p + 16 + ((sizeof(FvGrid) + 15) / 16) * 16);
    CK(cudaMemcpyAsync(c, e->fv_counts.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream), "D2H counts");
    CK(cudaMemcpyAsync(g, e->fv_grid.p, sizeof(FvGrid), cudaMemcpyDeviceToHost, e->stream), "D2H fv grid");
    if (ctl_out) CK(cudaMemcpyAsync(ctl, e->pcg_ctl.p, sizeof(SphereCtl), cudaMemcpyDeviceToHost, e->stream), "D2H ctl");
    CK(cudaStreamSynchronize(e->stream), "sync free volume");
    if (ctl_out) *ctl_out = *ctl;
    if (g->error) return e->fail(MCPYB_ERR_CAPACITY, "free-volume sphere capacity exceeded");
    counts_out[0] = (int64_t)c[0];
    counts_out[1] = (int64_t)c[1];
    return MCPYB_OK;
}

int mcpyb_free_volume_count_dev(mcpyb_engine* e, const double* d_points, int64_t P, const double* d_positions,
                                const double* d_radii, int64_t M, const double* offset3, const double* shifts,
                                int nshift, uint64_t* d_counts2) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (P < 0 || M < 0 || !d_counts2 || (P > 0 && !d_points) || (M > 0 && (!d_positions || !d_radii)))
        return e->fail(MCPYB_ERR_ARG, "bad free-volume arguments");
    return fv_launch(e, d_points, P, d_positions, d_radii, M, offset3, shifts, nshift,
                     (unsigned long long*)d_counts2, nullptr);
}

int mcpyb_free_volume_count_host(mcpyb_engine* e, const double* points, int64_t P, const double* positions,
                                 const double* radii, int64_t M, const double* offset3, const double* shifts,
                                 int nshift, int64_t* counts_out, uint8_t* covered_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (P < 0 || M < 0 || !counts_out || (P > 0 && !points) || (M > 0 && (!positions || !radii)))
        return e->fail(MCPYB_ERR_ARG, "bad free-volume arguments");
    CK(e->fv_points.reserve((size_t)(P > 0 ? 3 * P : 1)), "cudaMalloc fv points");
    if (covered_out) CK(e->fv_mask.reserve((size_t)(P > 0 ? P : 1)), "cudaMalloc fv mask");
    if (P) CK(cudaMemcpyAsync(e->fv_points.p, points, (size_t)P * 24, cudaMemcpyHostToDevice, e->stream), "H2D points");
    int rc = fv_upload_atoms(e, positions, radii, M);
    if (rc != MCPYB_OK) return rc;
    rc = fv_launch(e, e->fv_points.p, P, e->fv_pos.p, e->fv_radii.p, M, offset3, shifts, nshift,
                   e->fv_counts.p, covered_out ? e->fv_mask.p : nullptr);
    if (rc != MCPYB_OK) return rc;
    if (covered_out && P) CK(cudaMemcpyAsync(covered_out, e->fv_mask.p, (size_t)P, cudaMemcpyDeviceToHost, e->stream), "D2H mask");
    return fv_read_counts(e, counts_out);
}

static int sphere_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo,
                                    int64_t P, double radius, const double* center3, double zmin,
                                    const double* positions, const double* radii, int64_t M,
                                    int64_t* counts_out, int64_t* draws_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (zmin != zmin || P < 0 || M < 0 || !state_hi_lo || !inc_hi_lo || !counts_out || !(radius > 0.0) ||
        (M > 0 && (!positions || !radii)))
        return e->fail(MCPYB_ERR_ARG, "bad sphere free-volume arguments");
    int rc = fv_upload_atoms(e, positions, radii, M);
    if (rc != MCPYB_OK) return rc;
    rc = fv_build(e, e->fv_pos.p, e->fv_radii.p, M, nullptr, nullptr, 1);
    if (rc != MCPYB_OK) return rc;
    Pcg64 rng;
    rng.state = ((u128)state_hi_lo[0] << 64) | (u128)state_hi_lo[1];
    rng.inc = ((u128)inc_hi_lo[0] << 64) | (u128)inc_hi_lo[1];
    const long long Mc = (long long)(1.5 * (double)P) + 1;          // int(1.5 * n) + 1 candidates per pass
    // short chunks while the problem is small (more threads to hide the stencil-walk latency), long ones
    // once there are threads to spare
    const int chunk = Mc >= (4LL << 20) ? 8 : 2;
    const PcgStrideTable tbl = pcg_stride_table(rng.inc, chunk);
    const long long nthreads = (Mc + chunk - 1) / chunk;
    const int blocks = (int)((nthreads + PCG_BLOCK - 1) / PCG_BLOCK);
    CK(e->pcg_count.reserve(2 * (size_t)blocks), "cudaMalloc pcg");
    CK(e->pcg_prefix.reserve(2 * (size_t)blocks), "cudaMalloc pcg");
    CK(e->pcg_ctl.reserve(1), "cudaMalloc pcg");
    CK(cudaMemsetAsync(e->pcg_ctl.p, 0, sizeof(SphereCtl), e->stream), "memset ctl");
    CK(cudaMemsetAsync(e->fv_counts.p, 0, 2 * sizeof(unsigned long long), e->stream), "memset counts");
    const double cx = center3 ? center3[0] : 0.0, cy = center3 ? center3[1] : 0.0, cz = center3 ? center3[2] : 0.0;
    SphereCtl ctl;
    memset(&ctl, 0, sizeof ctl);
    int rounds = 0;
    const dim3 grid2(blocks, 2);
    while (P > 0) {
        // one round = two passes of the reference loop (enough in all but astronomically unlikely
        // cases): count, scan + bookkeeping, evaluate; then look
        if (chunk == 8) pcg_sphere_count_kernel<8><<<grid2, PCG_BLOCK, 0, e->stream>>>(rng, tbl, Mc, radius, P, e->pcg_ctl.p, e->pcg_count.p);
        else pcg_sphere_count_kernel<2><<<grid2, PCG_BLOCK, 0, e->stream>>>(rng, tbl, Mc, radius, P, e->pcg_ctl.p, e->pcg_count.p);
        LAUNCH_CHECK("pcg_sphere_count_kernel");
        pcg_scan_kernel<<<1, 1024, 0, e->stream>>>(e->pcg_count.p, e->pcg_prefix.p, blocks, P, e->pcg_ctl.p);
        LAUNCH_CHECK("pcg_scan_kernel");
        if (chunk == 8) pcg_sphere_eval_kernel<8><<<grid2, PCG_BLOCK, 0, e->stream>>>(rng, tbl, Mc, radius, cx, cy, cz, zmin, e->pcg_ctl.p,
            e->pcg_prefix.p, e->fv_grid.p, e->fv_spheres.p, e->fv_cell_start.p, e->fv_counts.p);
        else pcg_sphere_eval_kernel<2><<<grid2, PCG_BLOCK, 0, e->stream>>>(rng, tbl, Mc, radius, cx, cy, cz, zmin, e->pcg_ctl.p,
            e->pcg_prefix.p, e->fv_grid.p, e->fv_spheres.p, e->fv_cell_start.p, e->fv_counts.p);
        LAUNCH_CHECK("pcg_sphere_eval_kernel");
        rc = fv_read_counts(e, counts_out, &ctl);
        if (rc != MCPYB_OK) return rc;
        if (ctl.filled >= P) break;
        if (++rounds > 32) return e->fail(MCPYB_ERR_STATE, "sphere sampler did not converge");
    }
    if (draws_out) *draws_out = 3 * Mc * ctl.passes;
    if (P > 0) return MCPYB_OK;
    return fv_read_counts(e, counts_out);
}

int mcpyb_sphere_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo,
                                   int64_t P, double radius, const double* center3, const double* positions,
                                   const double* radii, int64_t M, int64_t* counts_out, int64_t* draws_out) {
    return sphere_free_volume_pcg64(e, state_hi_lo, inc_hi_lo, P, radius, center3, -HUGE_VAL, positions, radii, M,
                                    counts_out, draws_out);
}

int mcpyb_dome_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo,
                                 int64_t P, double radius, const double* center3, double bottom_z,
                                 const double* positions, const double* radii, int64_t M,
                                 int64_t* counts_out, int64_t* draws_out) {
    return sphere_free_volume_pcg64(e, state_hi_lo, inc_hi_lo, P, radius, center3, bottom_z, positions, radii, M,
                                    counts_out, draws_out);
}

int mcpyb_box_free_volume_pcg64(mcpyb_engine* e, const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo,
                                int64_t P, const double* dims9, const double* positions, const double* radii,
                                int64_t M, const double* offset3, const double* shifts, int nshift,
                                int64_t* counts_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (P < 0 || M < 0 || !state_hi_lo || !inc_hi_lo || !counts_out || !dims9 || (M > 0 && (!positions || !radii)))
        return e->fail(MCPYB_ERR_ARG, "bad box free-volume arguments");
    int rc = fv_upload_atoms(e, positions, radii, M);
    if (rc != MCPYB_OK) return rc;
    rc = fv_build(e, e->fv_pos.p, e->fv_radii.p, M, offset3, shifts, nshift);
    if (rc != MCPYB_OK) return rc;
    // dims ride in the tail of the small block; wait for the H2D that still reads its head
    CK(cudaStreamSynchronize(e->stream), "sync before dims staging");
    CK(e->fv_pin.reserve((3 + 81 + 9) * sizeof(double)), "cudaMallocHost fv");
    double* hs = (double*)e->fv_pin.p;
    for (int i = 0; i < 9; ++i) hs[84 + i] = dims9[i];
    CK(cudaMemcpyAsync(e->fv_small.p + 84, hs + 84, 9 * sizeof(double), cudaMemcpyHostToDevice, e->stream), "H2D dims");
    Pcg64 rng;
    rng.state = ((u128)state_hi_lo[0] << 64) | (u128)state_hi_lo[1];
    rng.inc = ((u128)inc_hi_lo[0] << 64) | (u128)inc_hi_lo[1];
    CK(cudaMemsetAsync(e->fv_counts.p, 0, 2 * sizeof(unsigned long long), e->stream), "memset counts");
    if (P > 0) {
        const int chunk = P >= (4LL << 20) ? 8 : 2;
        const PcgStrideTable tbl = pcg_stride_table(rng.inc, chunk);
        const long long nthreads = (P + chunk - 1) / chunk;
        const int nblk = (int)((nthreads + PCG_BLOCK - 1) / PCG_BLOCK);
        if (chunk == 8) pcg_box_eval_kernel<8><<<nblk, PCG_BLOCK, 0, e->stream>>>(rng, tbl, P, e->fv_small.p + 84,
            e->fv_grid.p, e->fv_spheres.p, e->fv_cell_start.p, e->fv_counts.p);
        else pcg_box_eval_kernel<2><<<nblk, PCG_BLOCK, 0, e->stream>>>(rng, tbl, P, e->fv_small.p + 84,
            e->fv_grid.p, e->fv_spheres.p, e->fv_cell_start.p, e->fv_counts.p);
        LAUNCH_CHECK("pcg_box_eval_kernel");
    }
    return fv_read_counts(e, counts_out);
}

void* mcpyb_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes > 0 ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void mcpyb_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int mcpyb_fp64_peak_tflops(mcpyb_engine* e, double* tflops_out) {
    if (!e || !tflops_out) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
    DevBuf<double> out;
    CK(out.reserve((size_t)blocks * threads), "cudaMalloc peak");
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0), "event"); CK(cudaEventCreate(&t1), "event");
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(t0, e->stream), "record");
        fp64_fma_peak_kernel<<<blocks, threads, 0, e->stream>>>(out.p, iters, 1.0000001, 1e-9);
        LAUNCH_CHECK("fp64_fma_peak_kernel");
        CK(cudaEventRecord(t1, e->stream), "record");
        CK(cudaEventSynchronize(t1), "sync");
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1), "elapsed");
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    out.release();
    *tflops_out = best;
    return MCPYB_OK;
}

// Stage timing of the batched LJ trial pipeline: sites | scan | scatter | rows | combine.
// enable != 0 switches event recording on; a later call folds the events of the most recent trial
// launch into the running sums (after synchronising the stream) and returns the mean ms per stage.
int mcpyb_stage_timing(mcpyb_engine* e, int enable, double* mean_ms_out5, int64_t* calls_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (enable && !e->stage_timing) {
        for (int i = 0; i < 6; ++i) CK(cudaEventCreate(&e->stage_ev[i]), "cudaEventCreate");
        e->stage_timing = true;
        e->stage_n = 0;
    }
    if (e->stage_timing && e->stage_n == 6) {
        CK(cudaStreamSynchronize(e->stream), "sync stage timing");
        for (int i = 0; i < 5; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e->stage_ev[i], e->stage_ev[i + 1]), "cudaEventElapsedTime");
            e->stage_ms[i] += ms;
        }
        e->stage_calls++;
        e->stage_n = 0;
    }
    if (mean_ms_out5)
        for (int i = 0; i < 5; ++i) mean_ms_out5[i] = e->stage_calls ? e->stage_ms[i] / (double)e->stage_calls : 0.0;
    if (calls_out) *calls_out = e->stage_calls;
    if (!enable && e->stage_timing) {
        for (int i = 0; i < 6; ++i) cudaEventDestroy(e->stage_ev[i]);
        e->stage_timing = false;
    }
    return MCPYB_OK;
}

int mcpyb_min_dist2_host(mcpyb_engine* e, int rep, const double* points, int64_t C,
                         const uint8_t* mask, double* out_d2) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (!e->have_batch) return e->fail(MCPYB_ERR_STATE, "no resident batch (call mcpyb_upload_* first)");
    if (rep < 0 || rep >= e->R) return e->fail(MCPYB_ERR_ARG, "replica out of range");
    if (C < 0 || (C > 0 && (!points || !out_d2))) return e->fail(MCPYB_ERR_ARG, "bad min-distance arguments");
    if (C == 0) return MCPYB_OK;
    const int n = e->atom_off_host[rep + 1] - e->atom_off_host[rep];
    DevBuf<double> d_pts, d_out;
    DevBuf<unsigned char> d_mask;
    CK(d_pts.reserve((size_t)3 * C), "cudaMalloc points"); CK(d_out.reserve((size_t)C), "cudaMalloc out");
    CK(cudaMemcpyAsync(d_pts.p, points, (size_t)C * 24, cudaMemcpyHostToDevice, e->stream), "H2D points");
    if (mask && n > 0) {
        CK(d_mask.reserve((size_t)n), "cudaMalloc mask");
        CK(cudaMemcpyAsync(d_mask.p, mask, (size_t)n, cudaMemcpyHostToDevice, e->stream), "H2D mask");
    }
    long long blocks = (C + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    min_dist2_kernel<<<(int)blocks, 256, 0, e->stream>>>(e->view(), rep, d_pts.p, (int)C,
                                                         (mask && n > 0) ? d_mask.p : nullptr, e->rcut, d_out.p);
    LAUNCH_CHECK("min_dist2_kernel");
    CK(cudaMemcpyAsync(out_d2, d_out.p, (size_t)C * 8, cudaMemcpyDeviceToHost, e->stream), "D2H min d2");
    CK(cudaStreamSynchronize(e->stream), "sync min d2");
    d_pts.release(); d_out.release(); d_mask.release();
    return MCPYB_OK;
}

// K10: `species AND inside` of SphericalCell / DomeCell.get_atoms_specie_inside_cell
// (spherical_cell.py:65-80, dome_cell.py).  The inside test is numpy's, operation for operation:
// diff = positions - center; einsum('ij,ij->i', diff, diff) sums (dx^2 + dz^2) + dy^2 (its unrolled pair + tail order,
// no fused multiply-add); `<= radius ** 2`; DomeCell adds `z >= bottom_z`.
__global__ void __launch_bounds__(256)
region_mask_kernel(const double* __restrict__ pos, const int* __restrict__ Z, int n, double cx, double cy, double cz,
                   double R2, int clip_z, double bottom_z, const int* __restrict__ species, int n_species,
                   unsigned char* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int z = Z[i];
    bool wanted = false;
    for (int k = 0; k < n_species; ++k) wanted |= species[k] == z;
    const double x = pos[3 * i], y = pos[3 * i + 1], zz = pos[3 * i + 2];
    const double dx = __dsub_rn(x, cx), dy = __dsub_rn(y, cy), dz = __dsub_rn(zz, cz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy));
    mask[i] = (wanted && d2 <= R2 && (!clip_z || zz >= bottom_z)) ? 1 : 0;
}

int mcpyb_region_mask_host(mcpyb_engine* e, const double* positions, const int32_t* numbers, int n, const double* center3,
                           double radius_squared, int clip_z, double bottom_z, const int32_t* species_Z, int n_species,
                           uint8_t* mask_out) {
    if (!e) return MCPYB_ERR_ARG;
    cudaSetDevice(e->device);
    if (n < 0 || n_species < 0 || n_species > 64 || !center3 || (n > 0 && (!positions || !numbers || !mask_out)) ||
        (n_species > 0 && !species_Z))
        return e->fail(MCPYB_ERR_ARG, "bad region-mask arguments");
    if (n == 0) return MCPYB_OK;
    // one device block and one page-locked block, kept with the engine: [positions 24 n | numbers 4 n | species | mask n]
    const size_t o_Z = (size_t)24 * n, o_sp = o_Z + (size_t)4 * n, o_mask = o_sp + 4 * 64, bytes = o_mask + (size_t)n;
    CK(e->mask_dev.reserve(bytes), "cudaMalloc region mask");
    CK(e->mask_pin.reserve(bytes), "cudaMallocHost region mask");
    unsigned char* h = e->mask_pin.p;
    unsigned char* d = e->mask_dev.p;
    memcpy(h, positions, (size_t)24 * n);
    memcpy(h + o_Z, numbers, (size_t)4 * n);
    if (n_species) memcpy(h + o_sp, species_Z, (size_t)4 * n_species);
    CK(cudaMemcpyAsync(d, h, o_mask, cudaMemcpyHostToDevice, e->stream), "H2D region mask");
    region_mask_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>((const double*)d, (const int*)(d + o_Z), n, center3[0], center3[1], center3[2],
                                                                radius_squared, clip_z, bottom_z, (const int*)(d + o_sp), n_species, d + o_mask);
    LAUNCH_CHECK("region_mask_kernel");
    CK(cudaMemcpyAsync(h + o_mask, d + o_mask, (size_t)n, cudaMemcpyDeviceToHost, e->stream), "D2H mask");
    CK(cudaStreamSynchronize(e->stream), "sync region mask");
    memcpy(mask_out, h + o_mask, (size_t)n);
    return MCPYB_OK;
}

int mcpyb_pcg64_host_doubles(const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo, uint64_t skip,
                             int64_t n, double* out) {
    if (!state_hi_lo || !inc_hi_lo || n < 0 || (n > 0 && !out)) return MCPYB_ERR_ARG;
    const u128 inc = ((u128)inc_hi_lo[0] << 64) | (u128)inc_hi_lo[1];
    u128 st = pcg_advance(((u128)state_hi_lo[0] << 64) | (u128)state_hi_lo[1], inc, skip);
    for (int64_t i = 0; i < n; ++i) out[i] = pcg_next_double(st, inc);
    return MCPYB_OK;
}

}  // extern "C"

// f2: device-resident GCMC chains (kernel + its C ABI); needs mcpyb_engine, CK and LAUNCH_CHECK from above
#include "gcmc_chain.cuh"
