// Stencil traversal shared by every pair kernel (K2..K6, K9).
//
// A "site" is a point p (an atom of the configuration, or the old / new
// position of a moved atom).  for_each_candidate() visits every image (j, S) of
// every atom whose cell lies within the stencil of p's cell and hands the
// functor the image vector d = (pos_j + S @ cell) - p and r2 = |d|^2, computed in
// the bit-exact order of oracle/neighbors.py.  Lanes of the cooperating group
// stride over the atoms of each contiguous x-run of cells, so global loads are
// 32-byte double4 records at unit stride.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int floor_div(int a, int n) {
    int q = a / n;
    if ((a % n != 0) && ((a < 0) != (n < 0))) --q;
    return q;
}

// One visited image.  same_image == true  <=>  S == 0, i.e. the candidate is the
// atom itself when the indices match.  (ax, ay, az) = S @ cell, so the absolute
// position of the image is (q + a) and other sites can be tested against it.
struct Cand {
    double4 q;
    long long tag;
    double r2, dx, dy, dz, ax, ay, az;
    bool same_image;
};

// F: void f(const Cand& c)
template <class F>
__device__ __forceinline__ void for_each_candidate(
        const GridDesc& g, const double4* __restrict__ spos, const int* __restrict__ cstart,
        double px, double py, double pz, int lane, int nlanes, F&& f,
        int row0 = 0, int rowstep = 1) {
    // row0 / rowstep: visit only the (y, z) rows ri with ri % rowstep == row0, so that several warps
    // can share one site row-wise (each with nlanes = 32) instead of queueing on the rows one by one
    int ri = -1;
    int c[3], w[3];
    locate(g, px, py, pz, c, w);
    const int wsite = pack_wraps(w[0], w[1], w[2]);
    const int n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    for (int oz = -g.m[2]; oz <= g.m[2]; ++oz) {
        int cz = c[2] + oz, Sz = 0;
        if (g.pbc[2]) { Sz = floor_div(cz, n2); cz -= Sz * n2; }
        else if (cz < 0 || cz >= n2) continue;
        for (int oy = -g.m[1]; oy <= g.m[1]; ++oy) {
            int cy = c[1] + oy, Sy = 0;
            if (g.pbc[1]) { Sy = floor_div(cy, n1); cy -= Sy * n1; }
            else if (cy < 0 || cy >= n1) continue;
            ++ri;
            if (rowstep > 1 && (ri % rowstep) != row0) continue;
            const int row = (cz * n1 + cy) * n0;
            // walk the x offsets, merging consecutive cells with the same image shift
            int seg_first = 0, seg_last = -2, seg_S = 0;
            for (int ox = -g.m[0]; ox <= g.m[0] + 1; ++ox) {
                int cx = c[0] + ox, Sx = 0;
                bool valid = ox <= g.m[0];
                if (valid) {
                    if (g.pbc[0]) { Sx = floor_div(cx, n0); cx -= Sx * n0; }
                    else if (cx < 0 || cx >= n0) valid = false;
                }
                if (valid && seg_last >= 0 && Sx == seg_S && cx == seg_last + 1) {
                    seg_last = cx;
                    continue;
                }
                if (seg_last >= 0) {
                    // flush [seg_first, seg_last] with shift (seg_S, Sy, Sz)
                    const int beg = cstart[row + seg_first];
                    const int end = cstart[row + seg_last + 1];
                    double sx, sy, sz;
                    shift_vec(g, seg_S, Sy, Sz, sx, sy, sz);
                    const bool zero_shift = (seg_S == 0 && Sy == 0 && Sz == 0);
                    for (int k = beg + lane; k < end; k += nlanes) {
                        const double4 q = spos[k];
                        const long long tag = tag_of(q);
                        const int wj = tag_wraps(tag);
                        double ax = sx, ay = sy, az = sz;
                        bool same = zero_shift;
                        if (wj != wsite) {
                            // rare: neighbour or site sits outside the primary cell
                            int jx, jy, jz;
                            unpack_wraps(wj, jx, jy, jz);
                            const int T0 = seg_S + w[0] - jx, T1 = Sy + w[1] - jy, T2 = Sz + w[2] - jz;
                            shift_vec(g, T0, T1, T2, ax, ay, az);
                            same = (T0 == 0 && T1 == 0 && T2 == 0);
                        }
                        Cand cd;
                        cd.q = q; cd.tag = tag; cd.ax = ax; cd.ay = ay; cd.az = az; cd.same_image = same;
                        cd.r2 = image_r2(q.x, q.y, q.z, ax, ay, az, px, py, pz, cd.dx, cd.dy, cd.dz);
                        f(cd);
                    }
                    seg_last = -2;
                }
                if (valid) { seg_first = cx; seg_last = cx; seg_S = Sx; }
            }
        }
    }
}
