/* Plain-C restatement of the CPU oracle (TEST INFRASTRUCTURE -- NOT PRODUCT CODE).
 *
 * Same definitions as oracle/neighbors.py, oracle/lj.py, oracle/emt.py and
 * oracle/free_volume.py (which cite the reference files they follow): loops over
 * every (i, j, image) in ascending order.  By default the j loop skips atoms whose
 * bin (edge >= cutoff / 2) is more than two bins from i's bin in any periodic image -- such a j
 * cannot pass the cutoff test, the visited (i, j, S) sequence and every sum stay
 * bit-identical to the plain O(N^2 * images) loop (oracle_set_bruteforce(1) selects
 * it; tests/test_oracle_lj.py compares the two), and the 64 x 2 000-atom and
 * 10 000-atom cases of BASELINE.json finish in seconds instead of minutes.  Compiled with -ffp-contract=off: the squared distance
 * (dx*dx + dy*dy) + dz*dz must be evaluated without FMA to stay bit-identical to
 * numpy and to the CUDA kernels.  tests/test_oracle_c.py pins this file to the
 * numpy oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double cell[9];   /* completed lattice rows */
    double inv[9];    /* frac_d = sum_j p_j inv[3*j+d] */
    double h[3];
    int pbc[3];
    int nimg[3];
} geom_t;

static void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

static int make_geom(const double* cell, const uint8_t* pbc, double cutoff, geom_t* g) {
    double c[3][3];
    int zero[3], nzero = 0;
    for (int i = 0; i < 9; ++i) c[i / 3][i % 3] = cell[i];
    for (int d = 0; d < 3; ++d) {
        zero[d] = (c[d][0] == 0 && c[d][1] == 0 && c[d][2] == 0);
        nzero += zero[d];
        if (zero[d] && pbc[d]) return -1;
    }
    if (nzero == 3) { for (int d = 0; d < 3; ++d) c[d][d] = 1.0; }
    else if (nzero == 2) {
        int k = !zero[0] ? 0 : (!zero[1] ? 1 : 2), a = (k + 1) % 3, b = (k + 2) % 3, imin = 0;
        for (int i = 1; i < 3; ++i) if (fabs(c[k][i]) < fabs(c[k][imin])) imin = i;
        double t[3] = {0, 0, 0}, v[3];
        t[imin] = 1.0;
        cross3(c[k], t, v);
        double nv = norm3(v);
        for (int i = 0; i < 3; ++i) c[a][i] = v[i] / nv;
        cross3(c[k], c[a], v);
        nv = norm3(v);
        for (int i = 0; i < 3; ++i) c[b][i] = v[i] / nv;
    } else if (nzero == 1) {
        int k = zero[0] ? 0 : (zero[1] ? 1 : 2);
        double v[3];
        cross3(c[(k + 1) % 3], c[(k + 2) % 3], v);
        double nv = norm3(v);
        if (nv == 0) return -1;
        for (int i = 0; i < 3; ++i) c[k][i] = v[i] / nv;
    }
    double det = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1]) - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0])
               + c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
    if (det == 0) return -1;
    double inv[3][3];
    inv[0][0] = (c[1][1] * c[2][2] - c[1][2] * c[2][1]) / det; inv[0][1] = (c[0][2] * c[2][1] - c[0][1] * c[2][2]) / det;
    inv[0][2] = (c[0][1] * c[1][2] - c[0][2] * c[1][1]) / det; inv[1][0] = (c[1][2] * c[2][0] - c[1][0] * c[2][2]) / det;
    inv[1][1] = (c[0][0] * c[2][2] - c[0][2] * c[2][0]) / det; inv[1][2] = (c[0][2] * c[1][0] - c[0][0] * c[1][2]) / det;
    inv[2][0] = (c[1][0] * c[2][1] - c[1][1] * c[2][0]) / det; inv[2][1] = (c[0][1] * c[2][0] - c[0][0] * c[2][1]) / det;
    inv[2][2] = (c[0][0] * c[1][1] - c[0][1] * c[1][0]) / det;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { g->cell[3 * i + j] = c[i][j]; g->inv[3 * i + j] = inv[i][j]; }
    for (int d = 0; d < 3; ++d) {
        double v[3];
        cross3(c[(d + 1) % 3], c[(d + 2) % 3], v);
        g->h[d] = fabs(det) / norm3(v);
        g->pbc[d] = pbc[d] ? 1 : 0;
        g->nimg[d] = pbc[d] ? (int)ceil(cutoff / g->h[d]) : 0;
    }
    return 0;
}

/* callback-free pair loop: visit(i, j, S, r2) for every image within the cutoff test */
typedef struct {
    int mode;            /* 0: r2 < rc2, 1: r2 <= rc2, 2: sqrt(r2) < rc */
    double rc, rc2;
} cut_t;

static inline int in_cut(const cut_t* c, double r2) {
    return c->mode == 0 ? (r2 < c->rc2) : c->mode == 1 ? (r2 <= c->rc2) : (sqrt(r2) < c->rc);
}

static int* wrap_counts(const double* pos, int n, const geom_t* g) {
    int* w = (int*)calloc((size_t)(n > 0 ? n : 1) * 3, sizeof(int));
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
            if (g->pbc[d]) {
                double f = pos[3 * i] * g->inv[d] + pos[3 * i + 1] * g->inv[3 + d] + pos[3 * i + 2] * g->inv[6 + d];
                w[3 * i + d] = (int)floor(f);
            }
    return w;
}

/* ---- candidate pruning: bins of perpendicular thickness >= cutoff over the wrapped fractional coordinates ---- */
static int g_bruteforce = 0;
void oracle_set_bruteforce(int on) { g_bruteforce = on; }

typedef struct {
    int nb[3];
    int* bin_start;   /* [nbins + 1] */
    int* bin_atoms;   /* atoms sorted by bin, ascending index inside a bin */
    int* bin_of;      /* [n][3] */
    int* cand;        /* scratch: candidates of the current atom, ascending */
    int ncand;
    int n;
    double* f;        /* wrapped fractional coordinates [n][3] */
    double reach[3];  /* an image farther than this (fractional, per axis) cannot be within the cutoff */
} bins_t;

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

static void bins_build(bins_t* b, const double* pos, int n, const geom_t* g, const int* w, double cutoff) {
    b->n = n;
    b->cand = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    b->bin_of = (int*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(int));
    double* f = (double*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
    double fmin[3] = {0, 0, 0}, fmax[3] = {1, 1, 1};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double fd = pos[3 * i] * g->inv[d] + pos[3 * i + 1] * g->inv[3 + d] + pos[3 * i + 2] * g->inv[6 + d];
            if (g->pbc[d]) fd -= (double)w[3 * i + d];
            f[3 * i + d] = fd;
        }
    for (int d = 0; d < 3; ++d) {
        if (!g->pbc[d] && n > 0) {
            fmin[d] = fmax[d] = f[d];
            for (int i = 1; i < n; ++i) {
                if (f[3 * i + d] < fmin[d]) fmin[d] = f[3 * i + d];
                if (f[3 * i + d] > fmax[d]) fmax[d] = f[3 * i + d];
            }
        }
        /* bins of fractional width wd with wd * h >= cutoff / 2 * (1 + 1e-9) */
        const double span = (fmax[d] - fmin[d]) * g->h[d];
        int nb = (int)floor(span / (0.5 * cutoff * (1.0 + 1e-9)));   /* two bins per cutoff: visit +-2 */
        if (nb < 1) nb = 1;
        if (nb > 64) nb = 64;
        b->nb[d] = nb;
    }
    const int nbins = b->nb[0] * b->nb[1] * b->nb[2];
    b->bin_start = (int*)calloc((size_t)nbins + 1, sizeof(int));
    b->bin_atoms = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) {
            int k = (int)floor((f[3 * i + d] - fmin[d]) / (fmax[d] - fmin[d] > 0 ? (fmax[d] - fmin[d]) : 1.0) * b->nb[d]);
            if (k < 0) k = 0;
            if (k >= b->nb[d]) k = b->nb[d] - 1;
            b->bin_of[3 * i + d] = k;
        }
        b->bin_start[(b->bin_of[3 * i] * b->nb[1] + b->bin_of[3 * i + 1]) * b->nb[2] + b->bin_of[3 * i + 2] + 1]++;
    }
    for (int k = 0; k < nbins; ++k) b->bin_start[k + 1] += b->bin_start[k];
    int* cur = (int*)malloc((size_t)nbins * sizeof(int));
    memcpy(cur, b->bin_start, (size_t)nbins * sizeof(int));
    for (int i = 0; i < n; ++i)
        b->bin_atoms[cur[(b->bin_of[3 * i] * b->nb[1] + b->bin_of[3 * i + 1]) * b->nb[2] + b->bin_of[3 * i + 2]]++] = i;
    free(cur);
    b->f = f;
    /* |df_d| * h_d is the distance between the lattice planes through the two points: a lower bound of the
     * pair distance.  1e-6 relative slack covers the rounding of f (|f| <= 512). */
    for (int d = 0; d < 3; ++d) b->reach[d] = g_bruteforce ? 1e300 : cutoff * (1.0 + 1e-6) / g->h[d] + 1e-9;
}

/* ascending list of the atoms j that can have an image within the cutoff of atom i */
static void bins_candidates(bins_t* b, const geom_t* g, int i) {
    b->ncand = 0;
    if (g_bruteforce) { for (int j = 0; j < b->n; ++j) b->cand[j] = j; b->ncand = b->n; return; }
    int lo[3], cnt[3];
    for (int d = 0; d < 3; ++d) {
        if (b->nb[d] <= 5 && g->pbc[d]) { lo[d] = 0; cnt[d] = b->nb[d]; }          /* every bin once */
        else if (g->pbc[d]) { lo[d] = b->bin_of[3 * i + d] - 2; cnt[d] = 5; }       /* wraps around */
        else {
            lo[d] = b->bin_of[3 * i + d] - 2; cnt[d] = 5;
            if (lo[d] < 0) { cnt[d] += lo[d]; lo[d] = 0; }
            if (lo[d] + cnt[d] > b->nb[d]) cnt[d] = b->nb[d] - lo[d];
        }
    }
    for (int a = 0; a < cnt[0]; ++a) for (int bb = 0; bb < cnt[1]; ++bb) for (int c = 0; c < cnt[2]; ++c) {
        const int k0 = ((lo[0] + a) % b->nb[0] + b->nb[0]) % b->nb[0], k1 = ((lo[1] + bb) % b->nb[1] + b->nb[1]) % b->nb[1],
                  k2 = ((lo[2] + c) % b->nb[2] + b->nb[2]) % b->nb[2];
        const int k = (k0 * b->nb[1] + k1) * b->nb[2] + k2;
        for (int q = b->bin_start[k]; q < b->bin_start[k + 1]; ++q) b->cand[b->ncand++] = b->bin_atoms[q];
    }
    qsort(b->cand, (size_t)b->ncand, sizeof(int), cmp_int);
}

static void bins_free(bins_t* b) { free(b->bin_start); free(b->bin_atoms); free(b->bin_of); free(b->cand); free(b->f); }

#define PAIR_LOOP_BEGIN \
    bins_t bins; bins_build(&bins, pos, n, &g, w, c.rc); \
    for (int i = 0; i < n; ++i) { bins_candidates(&bins, &g, i); \
    for (int jj = 0; jj < bins.ncand; ++jj) { const int j = bins.cand[jj]; \
    for (int s0 = -g.nimg[0]; s0 <= g.nimg[0]; ++s0) { \
    if (fabs(bins.f[3 * j] + s0 - bins.f[3 * i]) > bins.reach[0]) continue; \
    for (int s1 = -g.nimg[1]; s1 <= g.nimg[1]; ++s1) { \
    if (fabs(bins.f[3 * j + 1] + s1 - bins.f[3 * i + 1]) > bins.reach[1]) continue; \
    for (int s2 = -g.nimg[2]; s2 <= g.nimg[2]; ++s2) { \
        if (fabs(bins.f[3 * j + 2] + s2 - bins.f[3 * i + 2]) > bins.reach[2]) continue; \
        const int S0 = s0 + w[3 * i] - w[3 * j], S1 = s1 + w[3 * i + 1] - w[3 * j + 1], S2 = s2 + w[3 * i + 2] - w[3 * j + 2]; \
        if (i == j && S0 == 0 && S1 == 0 && S2 == 0) continue; \
        const double sx = ((double)S0 * g.cell[0] + (double)S1 * g.cell[3]) + (double)S2 * g.cell[6]; \
        const double sy = ((double)S0 * g.cell[1] + (double)S1 * g.cell[4]) + (double)S2 * g.cell[7]; \
        const double sz = ((double)S0 * g.cell[2] + (double)S1 * g.cell[5]) + (double)S2 * g.cell[8]; \
        const double dx = (pos[3 * j] + sx) - pos[3 * i], dy = (pos[3 * j + 1] + sy) - pos[3 * i + 1], \
                     dz = (pos[3 * j + 2] + sz) - pos[3 * i + 2]; \
        const double r2 = (dx * dx + dy * dy) + dz * dz; \
        if (!in_cut(&c, r2)) continue;
#define PAIR_LOOP_END } } } } } bins_free(&bins);

int64_t oracle_neighbor_pairs(const double* pos, int n, const double* cell, const uint8_t* pbc, double cutoff,
                              int mode, int64_t max_pairs, int32_t* oi, int32_t* oj, int32_t* oS, double* or2) {
    geom_t g;
    if (make_geom(cell, pbc, cutoff, &g)) return -1;
    cut_t c = {mode, cutoff, cutoff * cutoff};
    int* w = wrap_counts(pos, n, &g);
    int64_t k = 0;
    PAIR_LOOP_BEGIN
        if (k < max_pairs) { oi[k] = i; oj[k] = j; oS[3 * k] = S0; oS[3 * k + 1] = S1; oS[3 * k + 2] = S2; or2[k] = r2; }
        ++k;
    PAIR_LOOP_END
    free(w);
    return k;
}

/* ASE LennardJones(smooth=False): energies[i] = 0.5 sum_j (4 eps (c12 - c6) - e0), r2 <= rc^2 */
int oracle_lj_energy(const double* pos, int n, const double* cell, const uint8_t* pbc, double sigma, double eps,
                     double rc, double* energy, int64_t* npairs) {
    geom_t g;
    if (make_geom(cell, pbc, rc, &g)) return -1;
    cut_t c = {1, rc, rc * rc};
    int* w = wrap_counts(pos, n, &g);
    const double e0 = 4 * eps * (pow(sigma / rc, 12) - pow(sigma / rc, 6));
    /* per-atom sums as in ASE; long double accumulators keep the O(N^2) loop's rounding out of
     * the comparison (the quantity defined is the exact sum of the fp64 pair terms) */
    long double* en = (long double*)calloc((size_t)(n > 0 ? n : 1), sizeof(long double));
    int64_t k = 0;
    PAIR_LOOP_BEGIN
        const double t = sigma * sigma / r2, c6 = t * t * t, c12 = c6 * c6;
        en[i] += 0.5 * (4 * eps * (c12 - c6) - e0);
        ++k;
    PAIR_LOOP_END
    long double E = 0;
    for (int i = 0; i < n; ++i) E += en[i];
    *energy = (double)E;
    if (npairs) *npairs = k / 2;
    free(en); free(w);
    return 0;
}

/* EMT: table rows = E0,s0,V0,eta2,kappa,lambda,n0,gamma1,gamma2 per species slot; slot[i] per atom */
int oracle_emt_energy(const double* pos, const int32_t* slot, int n, const double* cell, const uint8_t* pbc,
                      const double* table, double rc, double rc_list, double acut, double beta,
                      double* energy, double* sigma1_out, int64_t* npairs) {
    geom_t g;
    if (make_geom(cell, pbc, rc_list, &g)) return -1;
    cut_t c = {2, rc_list, rc_list * rc_list};
    int* w = wrap_counts(pos, n, &g);
    double* sig = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    long double E = 0;          /* long double: see oracle_lj_energy */
    int64_t k = 0;
    PAIR_LOOP_BEGIN
        const double* pa = table + 9 * slot[i];
        const double* pb = table + 9 * slot[j];
        const double r = sqrt(r2), ksi = pb[6] / pa[6];
        const double theta = 1.0 / (1.0 + exp(acut * (r - rc)));
        E -= 0.5 * pa[2] * exp(-pb[4] * (r / beta - pb[1])) * ksi / pa[8] * theta;
        sig[i] += exp(-pb[3] * (r - beta * pb[1])) * ksi * theta / pa[7];
        ++k;
    PAIR_LOOP_END
    for (int i = 0; i < n; ++i) {
        const double* p = table + 9 * slot[i];
        if (!(sig[i] > 0)) { E -= p[0]; continue; }
        const double ds = -log(sig[i] / 12) / (beta * p[3]), x = p[5] * ds;
        E += p[0] * ((1 + x) * exp(-x) - 1) + 6 * p[2] * exp(-p[4] * ds);
    }
    *energy = (double)E;
    if (sigma1_out) memcpy(sigma1_out, sig, (size_t)n * sizeof(double));
    if (npairs) *npairs = k / 2;
    free(sig); free(w);
    return 0;
}

/* covered[p] = any sphere with (dx*dx+dy*dy)+dz*dz < r*r ; spheres = positions with radii > 0 */
int64_t oracle_fv_covered(const double* pts, int64_t P, const double* pos, const double* radii, int64_t M,
                          uint8_t* covered) {
    int64_t ncov = 0;
    for (int64_t p = 0; p < P; ++p) {
        uint8_t cv = 0;
        for (int64_t a = 0; a < M && !cv; ++a) {
            if (!(radii[a] > 0)) continue;
            const double dx = pts[3 * p] - pos[3 * a], dy = pts[3 * p + 1] - pos[3 * a + 1], dz = pts[3 * p + 2] - pos[3 * a + 2];
            if ((dx * dx + dy * dy) + dz * dz < radii[a] * radii[a]) cv = 1;
        }
        if (covered) covered[p] = cv;
        ncov += cv;
    }
    return ncov;
}
