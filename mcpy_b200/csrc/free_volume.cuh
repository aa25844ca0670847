// K7 -- Monte-Carlo free volume: integer count of sample points outside every
// exclusion sphere.  Replaces the cKDTree loop of
// mcpy/cell/spherical_cell.py:127-140 and mcpy/cell/custom_cell.py:80-89
// (periodic images: custom_cell.py:93-107).  A point is covered iff for some
// atom with radius r > 0:  (dx*dx + dy*dy) + dz*dz < r*r  in unfused fp64 -- the
// strict comparison cKDTree.query(distance_upper_bound=r) implements
// (SURVEY.md 8c), so the counts are bit-exact integers.
//
// fv_build_kernel : one CTA (or an 8-CTA cluster for large sets) bins the (imaged)
//                   exclusion spheres on a grid of edge >= r_max (counting sort,
//                   order inside a cell is irrelevant for an any()).
// fv_count_kernel : one thread per sample point, 27-cell stencil, early exit,
//                   warp ballot -> block sum -> one 64-bit atomicAdd per CTA.
#pragma once
#include "common.cuh"
#include "cell_list.cuh"      // block_reduce_minmax, block_inclusive_scan

struct FvGrid {
    double lo[3];
    double inv_edge[3];
    int ncell[3];
    int ncells;
    int natoms;        // expanded spheres with r > 0
    int error;
};

#define FV_BUILD_THREADS 1024

// pos [M][3], radii [M]; images: (pos - offset) + shifts[s], s < nshift.
// spheres out: double4 (x, y, z, r*r) in cell order; cell_start [max_cells+1]; cursor [max_cells]
//
// CS = 1 : one CTA.
// CS = 8 : one thread-block cluster of 8 CTAs for large sphere sets (the 10 825-atom S5 build took 97 us
//          in one CTA).  The per-sphere phases are strided over all 8 x 1024 threads and meet at cluster
//          barriers (release / acquire at cluster scope; everything exchanged lives in global memory);
//          the scan runs on CTA 0.  part[8][8] is scratch for the bounding-box partials.
template <int CS>
__global__ void __launch_bounds__(FV_BUILD_THREADS, 1)
fv_build_kernel(const double* __restrict__ pos, const double* __restrict__ radii, int M,
                const double* __restrict__ offset, const double* __restrict__ shifts, int nshift,
                FvGrid* __restrict__ grid, double4* __restrict__ spheres, int sphere_cap,
                int* __restrict__ cell_start, int* __restrict__ cursor, int max_cells,
                int* __restrict__ cid, double* __restrict__ part) {
    __shared__ double s_red[32];
    __shared__ int s_scan[FV_BUILD_THREADS];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    const int rank = blockIdx.x % CS;
    const long long gtid = (long long)rank * FV_BUILD_THREADS + tid;
    const long long gstride = (long long)CS * FV_BUILD_THREADS;
    auto sync_all = [&]() {
        if (CS == 1) { __syncthreads(); }
        else {
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        }
    };
    const double ox = offset[0], oy = offset[1], oz = offset[2];
    const long long total = (long long)M * nshift;

    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    double rmax = 0.0;
    int mine = 0;
    for (long long e = gtid; e < total; e += gstride) {
        const int a = (int)(e % M), s = (int)(e / M);
        const double r = radii[a];
        if (!(r > 0.0)) continue;
        const double x = __dadd_rn(__dsub_rn(pos[3 * a + 0], ox), shifts[3 * s + 0]);
        const double y = __dadd_rn(__dsub_rn(pos[3 * a + 1], oy), shifts[3 * s + 1]);
        const double z = __dadd_rn(__dsub_rn(pos[3 * a + 2], oz), shifts[3 * s + 2]);
        lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
        lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
        lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
        rmax = fmax(rmax, r);
        ++mine;
    }
    double glo[3], ghi[3];
    for (int d = 0; d < 3; ++d) {
        glo[d] = block_reduce_minmax(lo[d], true, s_red);
        ghi[d] = block_reduce_minmax(hi[d], false, s_red);
    }
    rmax = block_reduce_minmax(rmax, false, s_red);
    // sphere count of this CTA
    s_scan[tid] = mine;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (tid < o) s_scan[tid] += s_scan[tid + o];
        __syncthreads();
    }
    int nsph = s_scan[0];
    __syncthreads();
    if (CS > 1) {
        if (tid == 0) {
            double* pp = part + 8 * rank;
            for (int d = 0; d < 3; ++d) { pp[d] = glo[d]; pp[3 + d] = ghi[d]; }
            pp[6] = rmax; pp[7] = (double)nsph;
        }
        sync_all();
        nsph = 0;
        for (int k = 0; k < CS; ++k) {
            const double* pp = part + 8 * k;
            for (int d = 0; d < 3; ++d) { glo[d] = fmin(glo[d], pp[d]); ghi[d] = fmax(ghi[d], pp[3 + d]); }
            rmax = fmax(rmax, pp[6]);
            nsph += (int)pp[7];
        }
    }

    if (rank == 0 && tid == 0) {
        grid->natoms = nsph;
        grid->error = (nsph > sphere_cap) ? 1 : 0;
        const double edge_min = rmax * (1.0 + 1e-12);
        double want[3];
        for (int d = 0; d < 3; ++d) {
            const double width = (nsph > 0) ? (ghi[d] - glo[d]) : 0.0;
            double nc = (edge_min > 0.0) ? floor(width / edge_min) : 1.0;
            if (!(nc >= 1.0)) nc = 1.0;
            if (nc > 512.0) nc = 512.0;
            want[d] = nc;
            grid->lo[d] = (nsph > 0) ? glo[d] : 0.0;
        }
        double tot = want[0] * want[1] * want[2];
        if (tot > (double)max_cells) {
            const double f = cbrt((double)max_cells / tot);
            for (int d = 0; d < 3; ++d) want[d] = fmax(1.0, floor(want[d] * f));
            while (want[0] * want[1] * want[2] > (double)max_cells) {
                int k = 0;
                if (want[1] > want[k]) k = 1;
                if (want[2] > want[k]) k = 2;
                want[k] -= 1.0;
            }
        }
        for (int d = 0; d < 3; ++d) {
            grid->ncell[d] = (int)want[d];
            const double width = (nsph > 0) ? (ghi[d] - glo[d]) : 0.0;
            grid->inv_edge[d] = (width > 0.0) ? want[d] / width : 0.0;
        }
        grid->ncells = grid->ncell[0] * grid->ncell[1] * grid->ncell[2];
    }
    sync_all();
    if (grid->error) return;                 // uniform over the cluster
    const int ncells = grid->ncells;
    for (long long c = gtid; c <= ncells; c += gstride) cell_start[c] = 0;
    for (long long c = gtid; c < ncells; c += gstride) cursor[c] = 0;
    sync_all();

    auto cell_of = [&](double x, double y, double z) {
        int cx = (int)floor((x - grid->lo[0]) * grid->inv_edge[0]);
        int cy = (int)floor((y - grid->lo[1]) * grid->inv_edge[1]);
        int cz = (int)floor((z - grid->lo[2]) * grid->inv_edge[2]);
        cx = max(0, min(grid->ncell[0] - 1, cx));
        cy = max(0, min(grid->ncell[1] - 1, cy));
        cz = max(0, min(grid->ncell[2] - 1, cz));
        return (cz * grid->ncell[1] + cy) * grid->ncell[0] + cx;
    };
    for (long long e = gtid; e < total; e += gstride) {
        const int a = (int)(e % M), s = (int)(e / M);
        const double r = radii[a];
        if (!(r > 0.0)) { cid[e] = -1; continue; }
        const double x = __dadd_rn(__dsub_rn(pos[3 * a + 0], ox), shifts[3 * s + 0]);
        const double y = __dadd_rn(__dsub_rn(pos[3 * a + 1], oy), shifts[3 * s + 1]);
        const double z = __dadd_rn(__dsub_rn(pos[3 * a + 2], oz), shifts[3 * s + 2]);
        const int id = cell_of(x, y, z);
        cid[e] = id;
        atomicAdd(&cell_start[id + 1], 1);
    }
    sync_all();
    if (rank == 0) {
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < ncells; base += blockDim.x) {
            const int c = base + tid;
            const int scanned = block_inclusive_scan((c < ncells) ? cell_start[c + 1] : 0, s_scan);
            const int incl = scanned + s_carry;
            __syncthreads();
            if (c < ncells) cell_start[c + 1] = incl;
            if (tid == blockDim.x - 1) s_carry = incl;
            __syncthreads();
        }
    }
    sync_all();
    for (long long e = gtid; e < total; e += gstride) {
        const int id = cid[e];
        if (id < 0) continue;
        const int a = (int)(e % M), s = (int)(e / M);
        const double r = radii[a];
        const double x = __dadd_rn(__dsub_rn(pos[3 * a + 0], ox), shifts[3 * s + 0]);
        const double y = __dadd_rn(__dsub_rn(pos[3 * a + 1], oy), shifts[3 * s + 1]);
        const double z = __dadd_rn(__dsub_rn(pos[3 * a + 2], oz), shifts[3 * s + 2]);
        const int slot = cell_start[id] + atomicAdd(&cursor[id], 1);
        spheres[slot] = make_double4(x, y, z, __dmul_rn(r, r));
    }
}

// true iff some exclusion sphere strictly contains the point
__device__ __forceinline__ bool fv_point_covered(const FvGrid& g, const double4* __restrict__ spheres,
                                                 const int* __restrict__ cell_start,
                                                 double px, double py, double pz) {
    if (g.natoms <= 0) return false;
    // raw (unclamped) cell coordinates; the stencil is clipped to the grid
    const int cx = (int)floor(fmax(-2.0, fmin((px - g.lo[0]) * g.inv_edge[0], (double)g.ncell[0] + 1.0)));
    const int cy = (int)floor(fmax(-2.0, fmin((py - g.lo[1]) * g.inv_edge[1], (double)g.ncell[1] + 1.0)));
    const int cz = (int)floor(fmax(-2.0, fmin((pz - g.lo[2]) * g.inv_edge[2], (double)g.ncell[2] + 1.0)));
    // nine rows of three x-adjacent cells, the point's own row first (a covered point is usually covered
    // by an atom of its own neighbourhood -- the answer is an any(), so the order is free); all eighteen
    // row bounds are fetched before the first sphere so their latencies overlap
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.ncell[0] - 1);
    if (x0 > x1) return false;
    int beg[9], end[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        // order: (0,0) (0,-1) (0,1) (-1,0) (1,0) (-1,-1) (-1,1) (1,-1) (1,1) as (oy, oz)
        const int oy = (r == 3 || r == 5 || r == 6) ? -1 : ((r == 4 || r == 7 || r == 8) ? 1 : 0);
        const int oz = (r == 1 || r == 5 || r == 7) ? -1 : ((r == 2 || r == 6 || r == 8) ? 1 : 0);
        const int y = cy + oy, z = cz + oz;
        beg[r] = 0; end[r] = 0;
        if (y >= 0 && y < g.ncell[1] && z >= 0 && z < g.ncell[2]) {
            const int row = (z * g.ncell[1] + y) * g.ncell[0];
            beg[r] = cell_start[row + x0];
            end[r] = cell_start[row + x1 + 1];
        }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        for (int k = beg[r]; k < end[r]; ++k) {
            const double4 s = spheres[k];
            const double dx = __dsub_rn(px, s.x), dy = __dsub_rn(py, s.y), dz = __dsub_rn(pz, s.z);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d2 < s.w) return true;
        }
    }
    return false;
}

// block sum of (covered, total) -> one 64-bit atomicAdd pair per CTA; all threads must call
__device__ __forceinline__ void fv_block_accumulate(int covered_local, int total_local, int* s_cnt /* [8] */,
                                                    unsigned long long* __restrict__ counts) {
    const int c = warp_sum_int(covered_local);
    const int t = warp_sum_int(total_local);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    int cb = 0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) cb += s_cnt[w];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tb = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tb += s_cnt[w];
        if (tb) {
            atomicAdd(&counts[0], (unsigned long long)(tb - cb));
            atomicAdd(&counts[1], (unsigned long long)cb);
        }
    }
}

// points [P][3]; counts[0] += free points, counts[1] += covered points.
__global__ void __launch_bounds__(256)
fv_count_kernel(const double* __restrict__ points, long long P,
                const FvGrid* __restrict__ gridp, const double4* __restrict__ spheres,
                const int* __restrict__ cell_start, unsigned long long* __restrict__ counts,
                unsigned char* __restrict__ covered_mask) {
    __shared__ int s_cnt[8];
    const FvGrid g = *gridp;
    int covered_local = 0, total_local = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P;
         i += (long long)gridDim.x * blockDim.x) {
        const bool covered = fv_point_covered(g, spheres, cell_start, points[3 * i], points[3 * i + 1],
                                              points[3 * i + 2]);
        if (covered_mask) covered_mask[i] = covered ? 1 : 0;
        covered_local += covered ? 1 : 0;
        total_local += 1;
    }
    fv_block_accumulate(covered_local, total_local, s_cnt, counts);
}
