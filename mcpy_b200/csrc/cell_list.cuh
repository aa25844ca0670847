// K1 -- counting-sort cell list, one CTA per replica.
//
// Replaces the neighbour search structure the reference rebuilds inside
// ase.neighborlist.NeighborList (ASE 3.23.0, reached from
// mcpy/ensembles/base_ensemble.py:190-191) whenever the atom count changes.
// Steps, all inside one kernel so a rebuild costs one launch:
//   pack   : raw (x,y,z) + atomic number -> double4 with tag (index, species)
//   bbox   : fractional min/max along non-periodic axes (block reduction)
//   grid   : cells per axis = floor(width / rcut'), rcut' = rcut*(1+1e-12)
//   hash   : per-atom cell id + periodic wrap counts, shared/global histogram
//   scan   : exclusive prefix sum of the histogram -> cell_start
//   scatter: atoms to their cell segment
//   order  : each segment sorted by atom index, so the layout (and therefore
//            every floating-point sum taken over it) is deterministic.
#pragma once
#include "common.cuh"

#define CL_THREADS 1024

__device__ __forceinline__ double block_reduce_minmax(double v, bool is_min, double* sm) {
    // sm: at least 32 doubles
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fmin(v, u) : fmax(v, u);
    }
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = sm[0];
    for (int i = 1; i < nw; ++i) r = is_min ? fmin(r, sm[i]) : fmax(r, sm[i]);
    return r;
}

// Inclusive block scan (warp shuffles + one partial per warp); s_w holds >= 32 ints.
// Every thread of the block must call it; ends with the block synchronised.
__device__ __forceinline__ int block_inclusive_scan(int v, int* s_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();                       // s_w may still be read from a previous call
    if (lane == 31) s_w[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int x = (lane < nw) ? s_w[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        s_w[lane] = x;
    }
    __syncthreads();
    if (wid > 0) v += s_w[wid - 1];
    return v;
}

// raw_pos: [sum N][3] doubles, raw_Z: [sum N] int32, atom_off: [R+1] ints, geom: [R]
//
// CS = 1 : one CTA per replica (many small replicas: R CTAs run side by side).
// CS = 8 : one thread-block CLUSTER of 8 CTAs per replica for large configurations.  The per-atom
//          phases are strided over all 8 x 1024 threads; phases meet at cluster barriers
//          (barrier.cluster, release/acquire at cluster scope -- the arrays exchanged between CTAs
//          live in global memory), the two short scans run on CTA 0.  bbox[r][8][6] is scratch.
template <int CS>
__global__ void __launch_bounds__(CL_THREADS, 1)
build_cell_list_kernel(BatchView b, const double* __restrict__ raw_pos,
                       const int* __restrict__ raw_Z, const int* __restrict__ atom_off,
                       const CellGeom* __restrict__ geom, const int* __restrict__ z_to_slot,
                       double rcut, int subdiv, double* __restrict__ bbox) {
    __shared__ double s_red[32];
    __shared__ int s_scan[CL_THREADS];
    __shared__ int s_carry;
    __shared__ int s_err;
    const int r = blockIdx.x / CS;
    const int rank = blockIdx.x % CS;                 // == cluster block rank (1-D clusters)
    const int tid = threadIdx.x;
    const int gtid = rank * CL_THREADS + tid;
    const int gstride = CS * CL_THREADS;
    auto sync_all = [&]() {
        if (CS == 1) { __syncthreads(); }
        else {
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        }
    };
    GridDesc& g = b.grid[r];
    const int off = atom_off[r];
    const int n = atom_off[r + 1] - off;
    const CellGeom& cg = geom[r];
    if (tid == 0) s_err = 0;

    // ---- lattice into the device descriptor (every CTA keeps what it needs in registers too)
    if (rank == 0) {
        if (tid < 9) { g.cell[tid] = cg.cell[tid]; g.inv[tid] = cg.inv[tid]; }
        if (tid < 3) { g.height[tid] = cg.height[tid]; g.pbc[tid] = cg.pbc[tid]; }
        if (tid == 0) { g.ortho = cg.ortho; g.n = n; g.atom_off = off; g.error = 0; }
    }
    sync_all();

    // ---- pack + bounding box in fractional coordinates (non-periodic axes)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = gtid; i < n; i += gstride) {
        const double x = raw_pos[3 * (size_t)(off + i) + 0];
        const double y = raw_pos[3 * (size_t)(off + i) + 1];
        const double z = raw_pos[3 * (size_t)(off + i) + 2];
        int sp = 0;
        if (raw_Z != nullptr && z_to_slot != nullptr) {
            const int Z = raw_Z[off + i];
            sp = (Z >= 0 && Z < 120) ? z_to_slot[Z] : -1;
            if (sp < 0) { atomicOr(&s_err, 2); sp = 0; }
        }
        b.upos[off + i] = make_double4(x, y, z, make_tag(i, sp, pack_wraps(0, 0, 0)));
        double f[3];
        to_frac(g, x, y, z, f);
#pragma unroll
        for (int d = 0; d < 3; ++d) { lo[d] = fmin(lo[d], f[d]); hi[d] = fmax(hi[d], f[d]); }
    }
    double glo[3], ghi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        glo[d] = block_reduce_minmax(lo[d], true, s_red);
        ghi[d] = block_reduce_minmax(hi[d], false, s_red);
    }
    if (CS > 1) {
        if (tid == 0) {
            double* bb = bbox + ((size_t)r * CS + rank) * 6;
            for (int d = 0; d < 3; ++d) { bb[d] = glo[d]; bb[3 + d] = ghi[d]; }
        }
        sync_all();
        for (int k = 0; k < CS; ++k) {
            const double* bb = bbox + ((size_t)r * CS + k) * 6;
            for (int d = 0; d < 3; ++d) { glo[d] = fmin(glo[d], bb[d]); ghi[d] = fmax(ghi[d], bb[3 + d]); }
        }
    }

    // ---- binning grid (one thread), capped at max_cells.  Cell edge >= rcut' / subdiv, so the
    // stencil half width is m = ceil(rcut' / edge) <= subdiv cells (more only for boxes thinner
    // than the cutoff, where the extra offsets are periodic images of the single cell).
    if (rank == 0 && tid == 0) {
        const double rc_safe = rcut * (1.0 + 1e-12);
        double want[3], width[3], width_frac[3];
        for (int d = 0; d < 3; ++d) {
            if (g.pbc[d]) { g.lo[d] = 0.0; width_frac[d] = 1.0; }
            else if (n > 0) { g.lo[d] = glo[d]; width_frac[d] = ghi[d] - glo[d]; }
            else { g.lo[d] = 0.0; width_frac[d] = 0.0; }
            width[d] = width_frac[d] * g.height[d];               // Angstrom, perpendicular
            double nc = floor(width[d] * (double)subdiv / rc_safe);
            if (!(nc >= 1.0)) nc = 1.0;
            if (nc > 1024.0) nc = 1024.0;
            want[d] = nc;
        }
        double total = want[0] * want[1] * want[2];
        if (total > (double)b.max_cells) {
            const double f = cbrt((double)b.max_cells / total);
            for (int d = 0; d < 3; ++d) want[d] = fmax(1.0, floor(want[d] * f));
            // flooring can still overshoot by rounding; shave the largest axis
            while (want[0] * want[1] * want[2] > (double)b.max_cells) {
                int k = 0;
                if (want[1] > want[k]) k = 1;
                if (want[2] > want[k]) k = 2;
                want[k] -= 1.0;
            }
        }
        for (int d = 0; d < 3; ++d) {
            const int nc = (int)want[d];
            g.ncell[d] = nc;
            g.scale[d] = (width_frac[d] > 0.0) ? want[d] / width_frac[d] : 0.0;
            int m;
            if (g.pbc[d]) {
                m = (int)ceil(rc_safe * want[d] / width[d]);
            } else {
                m = (width[d] > 0.0) ? (int)ceil(rc_safe * want[d] / width[d]) : 0;
                if (m > nc - 1) m = nc - 1;                      // offsets beyond the grid are clipped anyway
            }
            if (m < 0) m = 0;
            if (m > 64) { m = 64; s_err |= 8; }
            g.m[d] = m;
        }
        g.ncells = g.ncell[0] * g.ncell[1] * g.ncell[2];
        // fp32 screening: reference point = grid origin, margin from the largest coordinate
        double ext = 0.0;
        int thin = 0;
        for (int d = 0; d < 3; ++d) {
            g.origin[d] = g.lo[0] * g.cell[d] + g.lo[1] * g.cell[3 + d] + g.lo[2] * g.cell[6 + d];
            ext += width_frac[d] * sqrt(g.cell[3 * d] * g.cell[3 * d] + g.cell[3 * d + 1] * g.cell[3 * d + 1]
                                        + g.cell[3 * d + 2] * g.cell[3 * d + 2]);
            if (g.pbc[d] && !(g.height[d] > rc_safe)) thin = 1;
        }
        g.thin = thin;
        const double margin = 64.0 * 5.9604644775390625e-08 * (ext + 2.0 * rc_safe) + 1e-30;
        const double rs = rc_safe + margin;
        g.screen_r2 = (float)(rs * rs * (1.0 + 1e-6));
    }
    sync_all();

    const int ncells = g.ncells;
    int* cstart = b.cell_start + (size_t)r * b.cs_stride;
    int* cursor = cstart + (b.max_cells + 1);        // second half: scatter cursors, later item_start
    for (int c = gtid; c <= ncells; c += gstride) { cstart[c] = 0; if (c < ncells) cursor[c] = 0; }
    sync_all();

    // ---- hash: cell id + wrap counts, histogram in cstart[c + 1]
    for (int i = gtid; i < n; i += gstride) {
        double4 q = b.upos[off + i];
        int c[3], w[3];
        if (!locate(g, q.x, q.y, q.z, c, w)) atomicOr(&s_err, 4);
        const long long t = tag_of(q);
        q.w = make_tag(i, tag_species(t), pack_wraps(w[0], w[1], w[2]));
        b.upos[off + i] = q;
        const int id = (c[2] * g.ncell[1] + c[1]) * g.ncell[0] + c[0];
        b.cid[off + i] = id;
        atomicAdd(&cstart[id + 1], 1);
    }
    sync_all();

    // ---- exclusive scan of cstart[1..ncells] (chunked block scan on CTA 0)
    if (rank == 0) {
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < ncells; base += blockDim.x) {
            const int c = base + tid;
            const int v = (c < ncells) ? cstart[c + 1] : 0;
            const int scanned = block_inclusive_scan(v, s_scan);
            const int incl = scanned + s_carry;
            __syncthreads();
            if (c < ncells) cstart[c + 1] = incl;          // inclusive -> start of next cell
            if (tid == blockDim.x - 1) s_carry = incl;
            __syncthreads();
        }
    }
    sync_all();
    // cstart[c] is now the first slot of cell c (cstart[0] = 0, cstart[ncells] = n)

    // ---- scatter (unordered) into the staging array
    for (int i = gtid; i < n; i += gstride) {
        const int id = b.cid[off + i];
        const int slot = cstart[id] + atomicAdd(&cursor[id], 1);
        b.stmp[off + slot] = b.upos[off + i];
    }
    sync_all();

    // ---- order each segment by atom index: rank = #atoms of the segment with a smaller
    // index (segments hold a few dozen atoms, reads are broadcast-friendly, no serial chain)
    for (int k = gtid; k < n; k += gstride) {
        const double4 me = b.stmp[off + k];
        const int mi = tag_index(tag_of(me));
        const int id = b.cid[off + mi];
        const int s = cstart[id], e = cstart[id + 1];
        int rk = 0;
        for (int a = s; a < e; ++a) rk += (tag_index(tag_of(b.stmp[off + a])) < mi) ? 1 : 0;
        b.spos[off + s + rk] = me;
        int w0, w1, w2;
        unpack_wraps(tag_wraps(tag_of(me)), w0, w1, w2);
        const float3 f = screen_pos(g, me.x, me.y, me.z, w0, w1, w2);
        b.fpos[off + s + rk] = make_float4(f.x, f.y, f.z, __int_as_float(mi));
    }
    sync_all();      // every CTA is done with the cursors before they become item_start

    // ---- work items of the lane-per-atom kernels: cell c contributes ceil(count/32) chunks of up
    // to 32 atoms; item_start (exclusive prefix, ncells+1 entries) reuses the cursor region.
    if (rank == 0) {
        int* item_start = cursor;
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < ncells; base += blockDim.x) {
            const int c = base + tid;
            const int v = (c < ncells) ? (cstart[c + 1] - cstart[c] + 31) / 32 : 0;
            const int scanned = block_inclusive_scan(v, s_scan);
            const int incl = scanned + s_carry;
            __syncthreads();
            if (c < ncells) item_start[c] = incl - v;
            if (tid == blockDim.x - 1) s_carry = incl;
            __syncthreads();
        }
        if (tid == 0) { item_start[ncells] = s_carry; g.nitems = s_carry; }
    }
    __syncthreads();
    if (tid == 0 && s_err) atomicOr(&g.error, s_err);
}
