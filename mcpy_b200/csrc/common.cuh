// Shared device-side types and helpers of the mcpy-b200 engine (sm_100a only).
//
// Data layout in HBM.  A "batch" is R resident
// configurations (replicas; R = 1 for a single GCMC run).  All replicas'
// atoms are concatenated:
//   upos[i]  double4  position in ORIGINAL atom order  (x, y, z, tag)
//   spos[i]  double4  the same atoms in cell-sorted order, per replica
//   cell_start[r*(max_cells+1) + c]  first sorted slot of cell c of replica r
//   grid[r]  GridDesc: lattice, inverse, binning of replica r (device-written)
// tag (64 bits stored in the .w lane) packs
//   bits  0..23  atom index inside its replica      (<= 16,777,215)
//   bits 24..31  species slot in the potential table (<= 255)
//   bits 32..61  three 10-bit periodic wrap counts, biased by 512
// so one 32-byte load delivers everything the pair loop needs.
#pragma once
// Bounds checks of the hand-computed indices (shared-memory staging slots, stencil-list runs, site slots, chain state):
// compiled in by `python -m mcpy_b200.build --bounds` (-DMCPYB_BOUNDS -> libmcpy_b200_bounds.so, selected with
// MCPYB_LIBRARY=...), absent from the product build.  A failing check stops the kernel with cudaErrorAssert and prints
// file:line; the GPU test-suite is run under this build as a stand-in for compute-sanitizer (closed on the pool).
#ifdef MCPYB_BOUNDS
#include <cassert>
#define MCPYB_CHECK(cond) assert(cond)
#else
#define MCPYB_CHECK(cond) ((void)0)
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#define MCPYB_MAX_SPECIES 16
#define MCPYB_WRAP_BIAS 512
#define MCPYB_MAX_MOVED 32           // atoms per trial move (rigid molecule size)

struct GridDesc {
    double cell[9];      // lattice rows a0,a1,a2 (rows of zero-length/non-periodic axes completed)
    double inv[9];       // frac_d = x*inv[0+d] + y*inv[3+d] + z*inv[6+d]
    double height[3];    // perpendicular width of the (completed) cell along axis d
    double lo[3];        // fractional origin of the binning grid (0 on periodic axes)
    double scale[3];     // cells per unit fractional coordinate
    int ncell[3];        // cells per axis
    int m[3];            // stencil half width, in cells (1 unless the box is thinner than rcut)
    int pbc[3];
    int ortho;           // lattice is diagonal: shift = S_d * cell[d][d]
    int n;               // atoms in this replica
    int atom_off;        // first atom of this replica in upos/spos
    int ncells;          // ncell[0]*ncell[1]*ncell[2]
    int nitems;          // sum over cells of ceil(count / 32): work items of the lane-per-atom kernels
    int thin;            // some periodic height <= rcut': moved atoms can see their own images
    double origin[3];    // cartesian reference point of the fp32 screening copies (grid origin)
    float screen_r2;     // (rcut + fp32 error margin)^2: fp32 screen never drops a true pair
    int error;           // nonzero: capacity exceeded / wrap out of range
};

// Host-supplied geometry of one replica (lattice only; binning is derived on device).
struct CellGeom {
    double cell[9];
    double inv[9];
    double height[3];
    int pbc[3];
    int ortho;
};

enum PotKind { POT_NONE = 0, POT_LJ = 1, POT_EMT = 2 };

struct EmtSpecies {      // derived constants of ase.calculators.emt
    double E0, s0, V0, eta2, kappa, lambda, n0, gamma1, gamma2;
    double inv_gamma1, inv_gamma2;
};

struct Potential {
    int kind;
    int nspecies;                 // EMT: species present in z_to_slot
    double rcut;                  // neighbour cutoff: LJ rc, EMT rc_list
    // LJ (ase.calculators.lj.LennardJones, smooth=False)
    double lj_eps, lj_sigma2, lj_rc2, lj_e0;
    // EMT
    double emt_rc, emt_acut, emt_beta;
    EmtSpecies emt[MCPYB_MAX_SPECIES];
    double emt_ksi[MCPYB_MAX_SPECIES][MCPYB_MAX_SPECIES];   // n0_b / n0_a
    int z_to_slot[120];           // atomic number -> species slot, -1 = unsupported
};

struct BatchView {
    GridDesc* grid;               // [R]
    double4* upos;                // [cap_atoms]
    double4* spos;                // [cap_atoms]
    double4* stmp;                // [cap_atoms] scratch of the counting sort
    float4* fpos;                 // [cap_atoms] fp32 screening copy of spos: wrapped position - origin, .w = index bits
    int* cell_start;              // [R * cs_stride]: cell_start (max_cells+1) then scatter cursors
    int* cid;                     // [cap_atoms] scratch: cell id per atom
    int R;
    int max_cells;
    int cs_stride;                // 2 * (max_cells + 1)
};

__device__ __forceinline__ long long tag_of(const double4& q) { return __double_as_longlong(q.w); }
__device__ __forceinline__ int tag_index(long long t) { return (int)(t & 0xFFFFFF); }
__device__ __forceinline__ int tag_species(long long t) { return (int)((t >> 24) & 0xFF); }
__device__ __forceinline__ int tag_wraps(long long t) { return (int)((t >> 32) & 0x3FFFFFFF); }
__device__ __forceinline__ int pack_wraps(int wx, int wy, int wz) {
    return (wx + MCPYB_WRAP_BIAS) | ((wy + MCPYB_WRAP_BIAS) << 10) | ((wz + MCPYB_WRAP_BIAS) << 20);
}
__device__ __forceinline__ void unpack_wraps(int w, int& wx, int& wy, int& wz) {
    wx = (w & 1023) - MCPYB_WRAP_BIAS;
    wy = ((w >> 10) & 1023) - MCPYB_WRAP_BIAS;
    wz = ((w >> 20) & 1023) - MCPYB_WRAP_BIAS;
}
__device__ __forceinline__ double make_tag(int index, int species, int wraps) {
    long long t = (long long)(index & 0xFFFFFF) | ((long long)(species & 0xFF) << 24)
                | ((long long)(wraps & 0x3FFFFFFF) << 32);
    return __longlong_as_double(t);
}

// Fractional coordinates: frac_d = (x*inv[d] + y*inv[3+d]) + z*inv[6+d]
__device__ __forceinline__ void to_frac(const GridDesc& g, double x, double y, double z, double f[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) f[d] = x * g.inv[d] + y * g.inv[3 + d] + z * g.inv[6 + d];
}

// Shift vector S @ cell in the oracle's order: (S0*a0 + S1*a1) + S2*a2, unfused.
__device__ __forceinline__ void shift_vec(const GridDesc& g, int S0, int S1, int S2,
                                          double& sx, double& sy, double& sz) {
    const double a = (double)S0, b = (double)S1, c = (double)S2;
    sx = __dadd_rn(__dadd_rn(__dmul_rn(a, g.cell[0]), __dmul_rn(b, g.cell[3])), __dmul_rn(c, g.cell[6]));
    sy = __dadd_rn(__dadd_rn(__dmul_rn(a, g.cell[1]), __dmul_rn(b, g.cell[4])), __dmul_rn(c, g.cell[7]));
    sz = __dadd_rn(__dadd_rn(__dmul_rn(a, g.cell[2]), __dmul_rn(b, g.cell[5])), __dmul_rn(c, g.cell[8]));
}

// r2 of the image vector d = (q + shift) - p, in the bit-exact association order
// (dx*dx + dy*dy) + dz*dz with no FMA contraction (oracle/neighbors.py).
__device__ __forceinline__ double image_r2(double qx, double qy, double qz,
                                           double sx, double sy, double sz,
                                           double px, double py, double pz,
                                           double& dx, double& dy, double& dz) {
    dx = __dsub_rn(__dadd_rn(qx, sx), px);
    dy = __dsub_rn(__dadd_rn(qy, sy), py);
    dz = __dsub_rn(__dadd_rn(qz, sz), pz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Cell coordinates and periodic wrap counts of a point.  Returns false when a
// wrap count does not fit the 10-bit field (atom > 511 boxes away).
__device__ __forceinline__ bool locate(const GridDesc& g, double x, double y, double z,
                                       int c[3], int w[3]) {
    double f[3];
    to_frac(g, x, y, z, f);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double fw = f[d];
        w[d] = 0;
        if (g.pbc[d]) {
            double fl = floor(fw);
            fw -= fl;                       // in [0, 1]; == 1.0 only by rounding of tiny negatives
            if (!(fl >= -511.0 && fl <= 511.0)) { ok = false; fl = 0.0; }
            w[d] = (int)fl;
            if (fw >= 1.0) { fw = 0.0; w[d] += 1; }
        }
        int ci = (int)floor((fw - g.lo[d]) * g.scale[d]);
        ci = max(0, min(g.ncell[d] - 1, ci));
        c[d] = ci;
    }
    return ok;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Wrapped position relative to the grid origin, rounded to fp32 (screening copy).
__device__ __forceinline__ float3 screen_pos(const GridDesc& g, double x, double y, double z, int w0, int w1, int w2) {
    const double a = (double)w0, b = (double)w1, c = (double)w2;
    const double xw = x - (a * g.cell[0] + b * g.cell[3] + c * g.cell[6]) - g.origin[0];
    const double yw = y - (a * g.cell[1] + b * g.cell[4] + c * g.cell[7]) - g.origin[1];
    const double zw = z - (a * g.cell[2] + b * g.cell[5] + c * g.cell[8]) - g.origin[2];
    return make_float3((float)xw, (float)yw, (float)zw);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining; pdl_wait() blocks until the
// predecessor has completed and its writes are visible, pdl_trigger() lets the successor start launching.
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#define MCPYB_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)
