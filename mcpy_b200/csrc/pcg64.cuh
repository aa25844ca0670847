// K8 -- numpy's PCG64 stream regenerated on the GPU, fused with the free-volume count.
//
// The reference draws its sample points from np.random.default_rng(seed)
// (mcpy/cell/cell.py:25): PCG64 = 128-bit LCG, XSL-RR 128/64 output, and
// Generator.random() = (next_uint64 >> 11) * 2^-53, filled in row-major order.
//   SphericalCell._sample_sphere_points (spherical_cell.py:88-106): passes of
//     M = int(1.5 P) + 1 candidates, pts = (u - 0.5) * 2.0 * R, kept when
//     einsum('ij,ij->i') <= R^2 -- numpy evaluates that as (x*x + z*z) + y*y, unfused
//     (tests/test_pcg64.py pins this against numpy itself) -- first P kept, in stream order.
//   CustomCell.calculate_volume (custom_cell.py:74-75): rng.random((P,3)) @ dimensions,
//     evaluated by the BLAS as fma(u2, d2c, fma(u1, d1c, u0*d0c)) (exact for diagonal cells).
// Each thread jumps the LCG to its chunk of the stream (two-level jump-ahead, below),
// so no random numbers cross the PCIe bus or touch HBM; the host generator is advanced
// by the same number of draws (Generator.bit_generator.advance) to stay in lock-step.
#pragma once
#include "common.cuh"
#include "free_volume.cuh"
#include "cell_list.cuh"      // block_inclusive_scan

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state, inc;
};

__host__ __device__ __forceinline__ u128 pcg_mult() {
    return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}

// state after `delta` steps (Brown's LCG jump-ahead, as in pcg_advance_lcg_128)
__host__ __device__ inline u128 pcg_advance(u128 state, u128 inc, unsigned long long delta) {
    u128 cur_mult = pcg_mult(), cur_plus = inc, acc_mult = 1, acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

__host__ __device__ __forceinline__ double pcg_next_double(u128& state, u128 inc) {
    state = state * pcg_mult() + inc;
    const unsigned long long hi = (unsigned long long)(state >> 64), lo = (unsigned long long)state;
    const unsigned long long x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    const unsigned long long out = (x >> rot) | (x << ((64 - rot) & 63));
    return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

#define PCG_BLOCK 256         // threads per block; a thread owns CHUNK consecutive candidates / points

// Two-level jump-ahead.  The full O(log n) jump to a block's first candidate is done by ONE thread per
// block; every thread then hops tid * 3 * CHUNK draws further with at most eight multiply-adds from a
// table of (A^n, plus_n), n = (3 * CHUNK) << k, prepared on the host for this stream's increment.  The
// per-thread cost no longer grows with the stream position, which is what lets small problems use
// short chunks (more threads in flight to hide the latency of the stencil walk).
struct PcgStrideTable {
    u128 mult[8], plus[8];
};

// coefficients of `delta` LCG steps: state' = mult * state + plus
__host__ __device__ inline void pcg_jump_coeffs(u128 inc, unsigned long long delta, u128& mult, u128& plus) {
    u128 cur_mult = pcg_mult(), cur_plus = inc, acc_mult = 1, acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    mult = acc_mult; plus = acc_plus;
}

inline PcgStrideTable pcg_stride_table(u128 inc, int chunk) {
    PcgStrideTable t;
    for (int k = 0; k < 8; ++k) pcg_jump_coeffs(inc, (unsigned long long)(3 * chunk) << k, t.mult[k], t.plus[k]);
    return t;
}

// state at the first draw of this thread's chunk; block_first_draw = draws consumed before the block's
// first candidate.  All threads of the block must call it (one __syncthreads).
__device__ __forceinline__ u128 pcg_thread_state(const Pcg64& rng, const PcgStrideTable& tbl,
                                                 unsigned long long block_first_draw, u128* s_base) {
    if (threadIdx.x == 0) *s_base = pcg_advance(rng.state, rng.inc, block_first_draw);
    __syncthreads();
    u128 st = *s_base;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((threadIdx.x >> k) & 1) st = st * tbl.mult[k] + tbl.plus[k];
    return st;
}

// Control block of the sphere sampler.  One round of launches covers TWO passes of the reference's
// `while len(pts) < n` loop (spherical_cell.py:93-106; at 1.5x oversampling the second pass is needed
// unless n is tiny): grid.y = pass of the round.  A pass that is not needed consumes no random numbers.
struct SphereCtl {
    long long filled;         // points accepted so far (all rounds)
    long long passes;         // passes that consumed random numbers (all rounds)
    long long base_pass;      // `passes` at the start of the current round
    long long need[2];        // points still wanted from pass 0 / 1 of the current round (<= 0: pass unused)
};

__device__ __forceinline__ bool sphere_candidate(u128& st, u128 inc, double R, double R2,
                                                 double& x, double& y, double& z) {
    const double u0 = pcg_next_double(st, inc), u1 = pcg_next_double(st, inc), u2 = pcg_next_double(st, inc);
    x = __dmul_rn(__dmul_rn(__dsub_rn(u0, 0.5), 2.0), R);
    y = __dmul_rn(__dmul_rn(__dsub_rn(u1, 0.5), 2.0), R);
    z = __dmul_rn(__dmul_rn(__dsub_rn(u2, 0.5), 2.0), R);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(z, z)), __dmul_rn(y, y));
    return d2 <= R2;
}

__device__ __forceinline__ int block_sum_int(int v, int* s_w /* [32] */) {
    v = warp_sum_int(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    return t;
}

// pass A: accepted candidates per BLOCK of PCG_BLOCK * CHUNK candidates; blockIdx.y = pass of the round
template <int CHUNK>
__global__ void __launch_bounds__(PCG_BLOCK)
pcg_sphere_count_kernel(Pcg64 rng, PcgStrideTable tbl, long long M, double R, long long P,
                        const SphereCtl* __restrict__ ctl, int* __restrict__ block_count) {
    __shared__ int s_w[32];
    __shared__ u128 s_base;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0 = t * CHUNK;
    int n = 0;
    const bool wanted = ctl->filled < P;                          // grid-uniform
    const long long pass = ctl->passes + blockIdx.y;
    u128 st = 0;
    if (wanted) st = pcg_thread_state(rng, tbl, (unsigned long long)(3 * (pass * M + (long long)blockIdx.x * PCG_BLOCK * CHUNK)), &s_base);
    if (c0 < M && wanted) {
        const long long c1 = min(c0 + (long long)CHUNK, M);
        const double R2 = __dmul_rn(R, R);
        for (long long c = c0; c < c1; ++c) {
            double x, y, z;
            n += sphere_candidate(st, rng.inc, R, R2, x, y, z) ? 1 : 0;
        }
    }
    const int tot = block_sum_int(n, s_w);
    if (threadIdx.x == 0) block_count[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// single block: per-pass exclusive scans of block_count -> block_prefix, then the bookkeeping of the
// round: how many points each pass still has to deliver, how many passes consumed random numbers
__global__ void __launch_bounds__(1024, 1)
pcg_scan_kernel(const int* __restrict__ block_count, long long* __restrict__ block_prefix, long long nblocks,
                long long P, SphereCtl* __restrict__ ctl) {
    __shared__ long long s[1024];
    __shared__ long long s_total[2];
    const int tid = threadIdx.x;
    const long long per = (nblocks + 1023) / 1024;
    const long long beg = min((long long)tid * per, nblocks), end = min(beg + per, nblocks);
    for (int pass = 0; pass < 2; ++pass) {
        const int* bc = block_count + (size_t)pass * nblocks;
        long long* bp = block_prefix + (size_t)pass * nblocks;
        long long sum = 0;
        for (long long i = beg; i < end; ++i) sum += bc[i];
        s[tid] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long v = (tid >= o) ? s[tid - o] : 0;
            __syncthreads();
            s[tid] += v;
            __syncthreads();
        }
        long long run = s[tid] - sum;
        for (long long i = beg; i < end; ++i) { bp[i] = run; run += bc[i]; }
        if (tid == 1023) s_total[pass] = s[1023];
        __syncthreads();
    }
    if (tid == 0) {
        const long long need0 = P - ctl->filled;
        const long long take0 = need0 > 0 ? min(s_total[0], need0) : 0;
        const long long need1 = need0 - take0;                    // > 0: the second pass is drawn too
        const long long take1 = need1 > 0 ? min(s_total[1], need1) : 0;
        ctl->base_pass = ctl->passes;
        ctl->need[0] = need0;
        ctl->need[1] = need1;
        ctl->passes += (need0 > 0 ? 1 : 0) + (need1 > 0 ? 1 : 0);
        ctl->filled += take0 + take1;
    }
}

// pass B: regenerate, rank the accepted candidates in stream order, evaluate those still needed
template <int CHUNK>
__global__ void __launch_bounds__(PCG_BLOCK)
pcg_sphere_eval_kernel(Pcg64 rng, PcgStrideTable tbl, long long M, double R, double cx, double cy, double cz,
                       double zmin /* DomeCell: points with z < zmin are drawn but not counted; -inf otherwise */,
                       const SphereCtl* __restrict__ ctl, const long long* __restrict__ block_prefix,
                       const FvGrid* __restrict__ gridp, const double4* __restrict__ spheres,
                       const int* __restrict__ cell_start, unsigned long long* __restrict__ counts) {
    __shared__ int s_cnt[8];
    __shared__ int s_scan[32];
    __shared__ u128 s_base;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0 = t * CHUNK;
    const long long need = ctl->need[blockIdx.y];
    const long long first = block_prefix[(size_t)blockIdx.y * gridDim.x + blockIdx.x];
    int covered_local = 0, total_local = 0;
    if (need > 0 && first < need) {                               // block-uniform
        double px[CHUNK], py[CHUNK], pz[CHUNK];
        unsigned acc_mask = 0;
        int n = 0;
        const long long pass = ctl->base_pass + blockIdx.y;
        u128 st = pcg_thread_state(rng, tbl, (unsigned long long)(3 * (pass * M + (long long)blockIdx.x * PCG_BLOCK * CHUNK)), &s_base);
        if (c0 < M) {
            const long long c1 = min(c0 + (long long)CHUNK, M);
            const double R2 = __dmul_rn(R, R);
#pragma unroll
            for (int k = 0; k < CHUNK; ++k) {
                if (c0 + k < c1 && sphere_candidate(st, rng.inc, R, R2, px[k], py[k], pz[k])) {
                    acc_mask |= 1u << k;
                    ++n;
                }
            }
        }
        // exclusive rank of this thread's first accepted candidate inside the block
        const int incl = block_inclusive_scan(n, s_scan);
        long long rank = first + (incl - n);
        if (rank < need) {
            const FvGrid g = *gridp;
#pragma unroll
            for (int k = 0; k < CHUNK; ++k) {
                if ((acc_mask >> k) & 1u) {
                    const double z = __dadd_rn(pz[k], cz);
                    if (rank < need && z >= zmin) {
                        const bool cov = fv_point_covered(g, spheres, cell_start, __dadd_rn(px[k], cx),
                                                          __dadd_rn(py[k], cy), z);
                        covered_local += cov ? 1 : 0;
                        total_local += 1;
                    }
                    ++rank;
                }
            }
        }
    }
    fv_block_accumulate(covered_local, total_local, s_cnt, counts);
}

// CustomCell: point i = rng.random(3) @ dims, one kernel
template <int CHUNK>
__global__ void __launch_bounds__(PCG_BLOCK)
pcg_box_eval_kernel(Pcg64 rng, PcgStrideTable tbl, long long P, const double* __restrict__ dims /* 9, device */,
                    const FvGrid* __restrict__ gridp, const double4* __restrict__ spheres,
                    const int* __restrict__ cell_start, unsigned long long* __restrict__ counts) {
    __shared__ int s_cnt[8];
    __shared__ u128 s_base;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = t * CHUNK;
    int covered_local = 0, total_local = 0;
    u128 st = pcg_thread_state(rng, tbl, (unsigned long long)(3 * (long long)blockIdx.x * PCG_BLOCK * CHUNK), &s_base);
    if (i0 < P) {
        const FvGrid g = *gridp;
        const long long i1 = min(i0 + (long long)CHUNK, P);
        double d[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k] = dims[k];
        for (long long i = i0; i < i1; ++i) {
            const double u0 = pcg_next_double(st, rng.inc), u1 = pcg_next_double(st, rng.inc),
                         u2 = pcg_next_double(st, rng.inc);
            const double x = fma(u2, d[6], fma(u1, d[3], __dmul_rn(u0, d[0])));
            const double y = fma(u2, d[7], fma(u1, d[4], __dmul_rn(u0, d[1])));
            const double z = fma(u2, d[8], fma(u1, d[5], __dmul_rn(u0, d[2])));
            covered_local += fv_point_covered(g, spheres, cell_start, x, y, z) ? 1 : 0;
            total_local += 1;
        }
    }
    fv_block_accumulate(covered_local, total_local, s_cnt, counts);
}
