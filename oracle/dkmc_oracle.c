/* TEST INFRASTRUCTURE ONLY — CPU restatement of the DeviceKMC field-and-rate hot path.
 * See dkmc_oracle.h for the contract, the pinning statement and the reference citations.
 * Built by oracle/Makefile with -ffp-contract=off so that no FMA contraction changes the
 * rounding relative to the reference's x86-64 baseline build (g++ -O3, no -march).
 * The product path (devicekmc_b200/) never links or calls this file. */
#include "dkmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ site_dist */
/* utils.cpp:100-137.  pbc: minimum image in y,z through round(); x never wraps. */
double orc_site_dist(double x1, double y1, double z1, double x2, double y2, double z2,
                     const double *lattice, int pbc) {
    if (pbc == 1) {
        double dx = x1 - x2;
        double fy = (y1 - y2) / lattice[1];
        fy -= round(fy);
        double fz = (z1 - z2) / lattice[2];
        fz -= round(fz);
        double dy = fy * lattice[1];
        double dz = fz * lattice[2];
        return sqrt(dx * dx + dy * dy + dz * dz);
    }
    double ax = x2 - x1, ay = y2 - y1, az = z2 - z1; /* pow(.,2) == exact square */
    return sqrt(ax * ax + ay * ay + az * az);
}

/* ------------------------------------------------------------------ neighbours */
typedef struct {
    int ncx, ncy, ncz;
    double minx, miny, minz, wx, wy, wz;
    int *cell_start; /* ncell+1 */
    int *cell_sites; /* N, ascending inside each cell */
    int *site_cell;
} cell_grid;

static double wrap_frac(double v, double L) {
    double f = v / L;
    f -= floor(f);
    if (f >= 1.0) f = 0.0;
    return f;
}

static void grid_build(cell_grid *g, int N, const double *x, const double *y, const double *z,
                       const double *lattice, int pbc, double cutoff) {
    double w = cutoff * 1.001 + 1e-9;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < N; ++i) {
        if (x[i] < lo[0]) lo[0] = x[i];
        if (x[i] > hi[0]) hi[0] = x[i];
        if (y[i] < lo[1]) lo[1] = y[i];
        if (y[i] > hi[1]) hi[1] = y[i];
        if (z[i] < lo[2]) lo[2] = z[i];
        if (z[i] > hi[2]) hi[2] = z[i];
    }
    g->minx = lo[0]; g->miny = lo[1]; g->minz = lo[2];
    g->ncx = (int)floor((hi[0] - lo[0]) / w) + 1;
    if (pbc) {
        g->ncy = (int)floor(lattice[1] / w); if (g->ncy < 1) g->ncy = 1;
        g->ncz = (int)floor(lattice[2] / w); if (g->ncz < 1) g->ncz = 1;
        g->wy = lattice[1] / g->ncy; g->wz = lattice[2] / g->ncz;
    } else {
        g->ncy = (int)floor((hi[1] - lo[1]) / w) + 1;
        g->ncz = (int)floor((hi[2] - lo[2]) / w) + 1;
        g->wy = w; g->wz = w;
    }
    g->wx = w;
    long ncell = (long)g->ncx * g->ncy * g->ncz;
    g->cell_start = (int *)calloc(ncell + 1, sizeof(int));
    g->cell_sites = (int *)malloc((size_t)N * sizeof(int));
    g->site_cell = (int *)malloc((size_t)N * sizeof(int));
    for (int i = 0; i < N; ++i) {
        int cx = (int)floor((x[i] - g->minx) / g->wx);
        int cy, cz;
        if (pbc) {
            cy = (int)(wrap_frac(y[i], lattice[1]) * g->ncy); if (cy >= g->ncy) cy = g->ncy - 1;
            cz = (int)(wrap_frac(z[i], lattice[2]) * g->ncz); if (cz >= g->ncz) cz = g->ncz - 1;
        } else {
            cy = (int)floor((y[i] - g->miny) / g->wy);
            cz = (int)floor((z[i] - g->minz) / g->wz);
        }
        if (cx >= g->ncx) cx = g->ncx - 1;
        if (cy >= g->ncy) cy = g->ncy - 1;
        if (cz >= g->ncz) cz = g->ncz - 1;
        int c = (cx * g->ncy + cy) * g->ncz + cz;
        g->site_cell[i] = c;
        g->cell_start[c + 1]++;
    }
    for (long c = 0; c < ncell; ++c) g->cell_start[c + 1] += g->cell_start[c];
    int *fill = (int *)malloc((size_t)ncell * sizeof(int));
    memcpy(fill, g->cell_start, (size_t)ncell * sizeof(int));
    for (int i = 0; i < N; ++i) g->cell_sites[fill[g->site_cell[i]]++] = i;
    free(fill);
}

static void grid_free(cell_grid *g) {
    free(g->cell_start); free(g->cell_sites); free(g->site_cell);
}

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* neighbours of site i through the cell grid, ascending; returns count (list may be NULL) */
static int grid_neighbors(const cell_grid *g, int i, const double *x, const double *y,
                          const double *z, const double *lattice, int pbc, double nn_dist,
                          int *list, int cap) {
    int c = g->site_cell[i];
    int cz = c % g->ncz, cy = (c / g->ncz) % g->ncy, cx = c / (g->ncz * g->ncy);
    int cnt = 0;
    int ylist[3], zlist[3], ny = 0, nz = 0;
    for (int d = -1; d <= 1; ++d) {
        int yy = cy + d, zz = cz + d;
        if (pbc) {
            yy = ((yy % g->ncy) + g->ncy) % g->ncy;
            zz = ((zz % g->ncz) + g->ncz) % g->ncz;
        }
        int dup = 0;
        if (yy >= 0 && yy < g->ncy) { for (int k = 0; k < ny; ++k) if (ylist[k] == yy) dup = 1; if (!dup) ylist[ny++] = yy; }
        dup = 0;
        if (zz >= 0 && zz < g->ncz) { for (int k = 0; k < nz; ++k) if (zlist[k] == zz) dup = 1; if (!dup) zlist[nz++] = zz; }
    }
    for (int dx = -1; dx <= 1; ++dx) {
        int xx = cx + dx;
        if (xx < 0 || xx >= g->ncx) continue;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nz; ++b) {
                int cc = (xx * g->ncy + ylist[a]) * g->ncz + zlist[b];
                for (int s = g->cell_start[cc]; s < g->cell_start[cc + 1]; ++s) {
                    int j = g->cell_sites[s];
                    if (j == i) continue;
                    double d = orc_site_dist(x[i], y[i], z[i], x[j], y[j], z[j], lattice, pbc);
                    if (d < nn_dist) {
                        if (list && cnt < cap) list[cnt] = j;
                        ++cnt;
                    }
                }
            }
    }
    if (list) qsort(list, cnt < cap ? cnt : cap, sizeof(int), cmp_int);
    return cnt;
}

int orc_neighbor_degrees(int N, const double *x, const double *y, const double *z,
                         const double *lattice, int pbc, double nn_dist, int *deg, int method) {
    int maxd = 0;
    if (method == 0) {
#pragma omp parallel for schedule(dynamic, 64) reduction(max : maxd)
        for (int i = 0; i < N; ++i) {
            int c = 0;
            for (int j = 0; j < N; ++j)
                if (i != j && orc_site_dist(x[i], y[i], z[i], x[j], y[j], z[j], lattice, pbc) < nn_dist) ++c;
            deg[i] = c;
            if (c > maxd) maxd = c;
        }
        return maxd;
    }
    cell_grid g;
    grid_build(&g, N, x, y, z, lattice, pbc, nn_dist);
#pragma omp parallel for schedule(dynamic, 256) reduction(max : maxd)
    for (int i = 0; i < N; ++i) {
        int c = grid_neighbors(&g, i, x, y, z, lattice, pbc, nn_dist, NULL, 0);
        deg[i] = c;
        if (c > maxd) maxd = c;
    }
    grid_free(&g);
    return maxd;
}

void orc_neighbor_fill(int N, const double *x, const double *y, const double *z,
                       const double *lattice, int pbc, double nn_dist, int nn, int *neigh_idx,
                       int method) {
    if (method == 0) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < N; ++i) {
            int c = 0;
            int *row = neigh_idx + (size_t)i * nn;
            for (int j = 0; j < N; ++j)
                if (i != j && orc_site_dist(x[i], y[i], z[i], x[j], y[j], z[j], lattice, pbc) < nn_dist) {
                    if (c < nn) row[c] = j;
                    ++c;
                }
            for (; c < nn; ++c) row[c] = -1;
        }
        return;
    }
    cell_grid g;
    grid_build(&g, N, x, y, z, lattice, pbc, nn_dist);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < N; ++i) {
        int *row = neigh_idx + (size_t)i * nn;
        int c = grid_neighbors(&g, i, x, y, z, lattice, pbc, nn_dist, row, nn);
        for (; c < nn; ++c) row[c] = -1;
    }
    grid_free(&g);
}

/* ------------------------------------------------------------------ layers */
void orc_site_layers(int N, const double *x, int n_layers, const double *start_x,
                     const double *end_x, int *site_layer) {
    for (int i = 0; i < N; ++i) {
        int id = -1;
        for (int l = 0; l < n_layers; ++l)
            if (start_x[l] <= x[i] && x[i] <= end_x[l]) id = l;
        site_layer[i] = id;
    }
}

/* ------------------------------------------------------------------ charge */
static int in_list(const int *list, int n, int v) {
    for (int k = 0; k < n; ++k) if (list[k] == v) return 1;
    return 0;
}

void orc_update_charge(int N, int nn, const int *neigh_idx, const int *element,
                       const int *metals, int num_metals, int *charge) {
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        const int *row = neigh_idx + (size_t)i * nn;
        if (element[i] == ORC_VACANCY) {
            int Vnn = 0;
            charge[i] = 2;
            for (int s = 0; s < nn && row[s] >= 0; ++s) {
                int ej = element[row[s]];
                if (ej == ORC_VACANCY) ++Vnn;
                if (in_list(metals, num_metals, ej)) { charge[i] = 0; break; }
                if (Vnn >= 2) { charge[i] = 0; break; }
            }
        }
        if (element[i] == ORC_OXYGEN_DEFECT) {
            charge[i] = -2;
            for (int s = 0; s < nn && row[s] >= 0; ++s)
                if (in_list(metals, num_metals, element[row[s]])) { charge[i] = 0; break; }
        }
    }
}

/* ------------------------------------------------------------------ CSR structure */
void orc_csr_row_ptr(int N, int nn, const int *neigh_idx, int NL, int NR, int *row_ptr,
                     int *left_row_ptr, int *right_row_ptr) {
    int m = N - NL - NR;
    row_ptr[0] = 0; left_row_ptr[0] = 0; right_row_ptr[0] = 0;
    for (int r = 0; r < m; ++r) {
        const int *row = neigh_idx + (size_t)(r + NL) * nn;
        int ci = 1, cl = 0, cr = 0; /* diagonal always present */
        for (int s = 0; s < nn && row[s] >= 0; ++s) {
            int j = row[s];
            if (j < NL) ++cl; else if (j >= N - NR) ++cr; else ++ci;
        }
        row_ptr[r + 1] = row_ptr[r] + ci;
        left_row_ptr[r + 1] = left_row_ptr[r] + cl;
        right_row_ptr[r + 1] = right_row_ptr[r] + cr;
    }
}

void orc_csr_fill(int N, int nn, const int *neigh_idx, int NL, int NR, const int *row_ptr,
                  int *col, const int *left_row_ptr, int *left_col, const int *right_row_ptr,
                  int *right_col) {
    int m = N - NL - NR;
#pragma omp parallel for
    for (int r = 0; r < m; ++r) {
        int i = r + NL;
        const int *row = neigh_idx + (size_t)i * nn;
        int pi = row_ptr[r], pl = left_row_ptr[r], pr = right_row_ptr[r];
        int diag_done = 0;
        for (int s = 0; s < nn && row[s] >= 0; ++s) {
            int j = row[s];
            if (j < NL) left_col[pl++] = j;
            else if (j >= N - NR) right_col[pr++] = j - (N - NR);
            else {
                if (!diag_done && j > i) { col[pi++] = r; diag_done = 1; }
                col[pi++] = j - NL;
            }
        }
        if (!diag_done) col[pi++] = r;
    }
}

/* ------------------------------------------------------------------ K assembly */
/* rule 0: background potential, potential_solver.cpp:325-346 — high_G iff (metal & metal) or
 *         (uncharged vacancy & uncharged vacancy);
 * rule 1: CB-edge Laplace solve, potential_solver.cpp:58-70 — high_G iff (metal || metal). */
static int g_rule = 0;
#pragma omp threadprivate(g_rule)
static double conductance(int ei, int qi, int ej, int qj, const int *metals, int nm,
                          double high_G, double low_G) {
    int metal1 = in_list(metals, nm, ei), metal2 = in_list(metals, nm, ej);
    if (g_rule == 1) return (metal1 || metal2) ? high_G : low_G;
    int cv1 = (ei == ORC_VACANCY && qi == 0), cv2 = (ej == ORC_VACANCY && qj == 0);
    return ((metal1 && metal2) || (cv1 && cv2)) ? high_G : low_G;
}

static void assemble_rule(int rule, double VL, double VR, int N, int nn, const int *neigh_idx, int NL, int NR,
                          const int *element, const int *charge, const int *metals, int num_metals, double high_G,
                          double low_G, const int *row_ptr, const int *col, double *val, double *rhs);

void orc_assemble_K(int N, int nn, const int *neigh_idx, int NL, int NR, const int *element,
                    const int *charge, const int *metals, int num_metals, double high_G,
                    double low_G, double Vd, const int *row_ptr, const int *col, double *val,
                    double *rhs) {
    assemble_rule(0, -Vd / 2, Vd / 2, N, nn, neigh_idx, NL, NR, element, charge, metals, num_metals, high_G, low_G,
                  row_ptr, col, val, rhs);
}

static void assemble_rule(int rule, double VL, double VR, int N, int nn, const int *neigh_idx, int NL, int NR,
                          const int *element, const int *charge, const int *metals, int num_metals, double high_G,
                          double low_G, const int *row_ptr, const int *col, double *val, double *rhs) {
    int m = N - NL - NR;
#pragma omp parallel for
    for (int r = 0; r < m; ++r) {
        g_rule = rule;
        int i = r + NL;
        const int *row = neigh_idx + (size_t)i * nn;
        /* diagonal: K[i][i] += -1*K[i][j] over ascending j (zeros add exactly), :350-359 */
        double diag = 0.0, ksub = 0.0;
        for (int s = 0; s < nn && row[s] >= 0; ++s) {
            int j = row[s];
            double kij = -conductance(element[i], charge[i], element[j], charge[j], metals,
                                      num_metals, high_G, low_G);
            diag += -1 * kij;
        }
        /* Ksub: left block then right block, each ascending j, :362-372 */
        for (int s = 0; s < nn && row[s] >= 0; ++s) {
            int j = row[s];
            if (j < NL)
                ksub += -conductance(element[i], charge[i], element[j], charge[j], metals, num_metals, high_G, low_G) * VL;
        }
        for (int s = 0; s < nn && row[s] >= 0; ++s) {
            int j = row[s];
            if (j >= N - NR)
                ksub += -conductance(element[i], charge[i], element[j], charge[j], metals, num_metals, high_G, low_G) * VR;
        }
        /* reference solves D*y = Ksub and sets phi = -y  (:379,396)  <=>  D*phi = -Ksub */
        rhs[r] = -ksub;
        for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
            int c = col[p];
            if (c == r) val[p] = diag;
            else {
                int j = c + NL;
                val[p] = -conductance(element[i], charge[i], element[j], charge[j], metals,
                                      num_metals, high_G, low_G);
            }
        }
    }
}

/* ------------------------------------------------------------------ solver */
static void spmv(int m, const int *rp, const int *ci, const double *v, const double *x, double *y) {
#pragma omp parallel for
    for (int r = 0; r < m; ++r) {
        double s = 0.0;
        for (int p = rp[r]; p < rp[r + 1]; ++p) s += v[p] * x[ci[p]];
        y[r] = s;
    }
}
static double dot(int m, const double *a, const double *b) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (int i = 0; i < m; ++i) s += a[i] * b[i];
    return s;
}

/* scaled CG on As = D^-1/2 A D^-1/2 (values pre-scaled in vs), recurrences of
 * iterative_solvers_gpu.cu:411-448: r = As y - b; p = -r; alpha = rr/(p As p); y += alpha p;
 * r += alpha As p; beta = rr'/rr; p = beta p - r.  Stops at ||r|| <= tol*||b||. */
static int cg_scaled(int m, const int *rp, const int *ci, const double *vs, const double *b,
                     double *y, double tol, int max_iter, double *relres) {
    double *r = (double *)malloc(sizeof(double) * m), *p = (double *)malloc(sizeof(double) * m),
           *t = (double *)malloc(sizeof(double) * m);
    spmv(m, rp, ci, vs, y, r);
    for (int i = 0; i < m; ++i) { r[i] -= b[i]; p[i] = -r[i]; }
    double bb = dot(m, b, b);
    double rr = dot(m, r, r);
    double stop = tol * tol * bb;
    int it = 0;
    while (rr > stop && it < max_iter && rr > 0.0) {
        spmv(m, rp, ci, vs, p, t);
        double alpha = rr / dot(m, p, t);
#pragma omp parallel for
        for (int i = 0; i < m; ++i) { y[i] += alpha * p[i]; r[i] += alpha * t[i]; }
        double rn = dot(m, r, r);
        double beta = rn / rr;
#pragma omp parallel for
        for (int i = 0; i < m; ++i) p[i] = beta * p[i] - r[i];
        rr = rn;
        ++it;
    }
    if (relres) *relres = bb > 0 ? sqrt(rr / bb) : sqrt(rr);
    free(r); free(p); free(t);
    return it;
}

int orc_solve(int m, const int *row_ptr, const int *col, const double *val, const double *rhs,
              double *x, double tol, int max_iter, int refine, double *info) {
    int nnz = row_ptr[m];
    double *dinv = (double *)malloc(sizeof(double) * m);
    double *vs = (double *)malloc(sizeof(double) * (size_t)nnz);
    double *b = (double *)malloc(sizeof(double) * m), *y = (double *)malloc(sizeof(double) * m);
    double *res = (double *)malloc(sizeof(double) * m);
    for (int r = 0; r < m; ++r) {
        double d = 1.0;
        for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) if (col[p] == r) d = val[p];
        dinv[r] = 1.0 / sqrt(d);
    }
#pragma omp parallel for
    for (int r = 0; r < m; ++r)
        for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) vs[p] = val[p] * dinv[r] * dinv[col[p]];
    for (int r = 0; r < m; ++r) { b[r] = rhs[r] * dinv[r]; y[r] = x[r] / dinv[r]; }
    double relres = 0.0;
    int iters = cg_scaled(m, row_ptr, col, vs, b, y, tol, max_iter, &relres);
    for (int r = 0; r < m; ++r) x[r] = y[r] * dinv[r];
    double resinf = 0.0;
    for (int round = 0; round <= refine; ++round) {
        /* residual of the UNSCALED system accumulated in binary128 */
        resinf = 0.0;
#pragma omp parallel for reduction(max : resinf)
        for (int r = 0; r < m; ++r) {
            __float128 s = (__float128)rhs[r];
            for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
                s -= (__float128)val[p] * (__float128)x[col[p]];
            res[r] = (double)s;
            double a = fabs(res[r]);
            if (a > resinf) resinf = a;
        }
        if (round == refine) break;
        for (int r = 0; r < m; ++r) { b[r] = res[r] * dinv[r]; y[r] = 0.0; }
        double rr2;
        iters += cg_scaled(m, row_ptr, col, vs, b, y, tol, max_iter, &rr2);
        for (int r = 0; r < m; ++r) x[r] += y[r] * dinv[r];
    }
    if (info) { info[0] = iters; info[1] = relres; info[2] = resinf; }
    free(dinv); free(vs); free(b); free(y); free(res);
    return iters;
}

void orc_background_potential(int N, int nn, const int *neigh_idx, int NL, int NR,
                              const int *element, const int *charge, const int *metals,
                              int num_metals, double high_G, double low_G, double Vd,
                              double *site_potential_boundary, double tol, int max_iter,
                              int refine, double *info) {
    int m = N - NL - NR;
    int *rp = (int *)malloc(sizeof(int) * (m + 1)), *lrp = (int *)malloc(sizeof(int) * (m + 1)),
        *rrp = (int *)malloc(sizeof(int) * (m + 1));
    orc_csr_row_ptr(N, nn, neigh_idx, NL, NR, rp, lrp, rrp);
    int *ci = (int *)malloc(sizeof(int) * (size_t)(rp[m] + 1)), *lci = (int *)malloc(sizeof(int) * (size_t)(lrp[m] + 1)),
        *rci = (int *)malloc(sizeof(int) * (size_t)(rrp[m] + 1));
    orc_csr_fill(N, nn, neigh_idx, NL, NR, rp, ci, lrp, lci, rrp, rci);
    double *val = (double *)malloc(sizeof(double) * (size_t)rp[m]), *rhs = (double *)malloc(sizeof(double) * m);
    orc_assemble_K(N, nn, neigh_idx, NL, NR, element, charge, metals, num_metals, high_G, low_G, Vd, rp, ci, val, rhs);
    orc_solve(m, rp, ci, val, rhs, site_potential_boundary + NL, tol, max_iter, refine, info);
    for (int i = 0; i < NL; ++i) site_potential_boundary[i] = -Vd / 2;
    for (int i = N - NR; i < N; ++i) site_potential_boundary[i] = Vd / 2;
    free(rp); free(lrp); free(rrp); free(ci); free(lci); free(rci); free(val); free(rhs);
}

/* Device::setLaplacePotential, CPU branch (potential_solver.cpp:4-139): the same Kirchhoff system with
 * the rule "high_G iff either site is a metal", contacts at +q Vd / 2 (left) and -q Vd / 2 (right);
 * site_CB_edge = contact values | -inv(D) Ksub.  `charge` does not enter. */
void orc_laplace_cb_edge(int N, int nn, const int *neigh_idx, int NL, int NR, const int *element,
                         const int *metals, int num_metals, double high_G, double low_G, double Vd,
                         double q, double *site_CB_edge, double tol, int max_iter, int refine,
                         double *info) {
    int m = N - NL - NR;
    int *rp = (int *)malloc(sizeof(int) * (m + 1)), *lrp = (int *)malloc(sizeof(int) * (m + 1)),
        *rrp = (int *)malloc(sizeof(int) * (m + 1));
    orc_csr_row_ptr(N, nn, neigh_idx, NL, NR, rp, lrp, rrp);
    int *ci = (int *)malloc(sizeof(int) * (size_t)(rp[m] + 1)), *lci = (int *)malloc(sizeof(int) * (size_t)(lrp[m] + 1)),
        *rci = (int *)malloc(sizeof(int) * (size_t)(rrp[m] + 1));
    orc_csr_fill(N, nn, neigh_idx, NL, NR, rp, ci, lrp, lci, rrp, rci);
    double *val = (double *)malloc(sizeof(double) * (size_t)rp[m]), *rhs = (double *)malloc(sizeof(double) * m);
    int *zero_charge = (int *)calloc((size_t)N, sizeof(int));
    const double VL = q * Vd / 2, VR = -q * Vd / 2;
    assemble_rule(1, VL, VR, N, nn, neigh_idx, NL, NR, element, zero_charge, metals, num_metals, high_G, low_G, rp, ci,
                  val, rhs);
    orc_solve(m, rp, ci, val, rhs, site_CB_edge + NL, tol, max_iter, refine, info);
    for (int i = 0; i < NL; ++i) site_CB_edge[i] = VL;
    for (int i = N - NR; i < N; ++i) site_CB_edge[i] = VR;
    free(rp); free(lrp); free(rrp); free(ci); free(lci); free(rci); free(val); free(rhs); free(zero_charge);
}

/* ------------------------------------------------------------------ pairwise Coulomb */
static const double ORC_Q = 1.60217663e-19; /* Device.h:113, KMCProcess.h:36 */
static const double ORC_KB = 8.617333262e-5; /* KMCProcess.h:35 */

static double v_solve(double r_dist, int charge, double sigma, double k, double q) {
    return (double)charge * erfc(r_dist / (sigma * sqrt(2))) * k * q / r_dist; /* utils.h:102 */
}

void orc_poisson_gridless_rows(int N, const double *x, const double *y, const double *z,
                               const double *lattice, int pbc, const int *charge, double sigma,
                               double k, int row_begin, int row_end, double *out_rows) {
    /* compaction of the charged sites keeps ascending j, i.e. the reference's summation order */
    int nc = 0;
    int *cj = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int j = 0; j < N; ++j) if (charge[j] != 0) cj[nc++] = j;
#pragma omp parallel for schedule(static)
    for (int i = row_begin; i < row_end; ++i) {
        double V = 0.0;
        for (int t = 0; t < nc; ++t) {
            int j = cj[t];
            if (i != j) {
                double r = (1e-10) * orc_site_dist(x[i], y[i], z[i], x[j], y[j], z[j], lattice, pbc);
                V += v_solve(r, charge[j], sigma, k, ORC_Q);
            }
        }
        out_rows[i - row_begin] = V;
    }
    free(cj);
}

void orc_poisson_gridless(int N, const double *x, const double *y, const double *z,
                          const double *lattice, int pbc, const int *charge, double sigma,
                          double k, double *site_potential_charge) {
    orc_poisson_gridless_rows(N, x, y, z, lattice, pbc, charge, sigma, k, 0, N, site_potential_charge);
}

/* ------------------------------------------------------------------ rate table */
void orc_rate_table(int N, int nn, const int *neigh_idx, const int *site_layer,
                    const double *lattice, int pbc, double T_bg, double freq, double sigma,
                    double k, const double *x, const double *y, const double *z,
                    const double *pb, const double *pc, const int *element, const int *charge,
                    const double *E_gen, const double *E_rec, const double *E_Vdiff,
                    const double *E_Odiff, int *event_type, double *event_prob) {
    long total = (long)N * nn;
    const double kB = ORC_KB, q = ORC_Q;
#pragma omp parallel for schedule(static)
    for (long idx = 0; idx < total; ++idx) {
        int et = ORC_NULL_EVENT;
        double P = 0;
        int i = (int)(idx / nn);
        int j = neigh_idx[idx];
        if (j >= 0 && j < N) {
            double r_dist = (1e-10) * orc_site_dist(x[i], y[i], z[i], x[j], y[j], z[j], lattice, pbc);
            if (element[i] == ORC_DEFECT && element[j] == ORC_O_EL) {
                double E = 2 * ((pb[i] + pc[i]) - (pb[j] + pc[j]));
                double zf = E_gen[site_layer[j]];
                et = ORC_VACANCY_GENERATION;
                double Ekin = 0;
                double EA = zf - E - Ekin;
                P = exp(-1 * EA / (kB * T_bg)) * freq;
            }
            if (element[i] == ORC_OXYGEN_DEFECT && element[j] == ORC_VACANCY) {
                double self_int_V = v_solve(r_dist, 2, sigma, k, q);
                int cs = charge[i] - charge[j];
                double E = cs * ((pb[i] + pc[i]) - (pb[j] + pc[j]) + (cs / 2) * self_int_V);
                double zf = E_rec[site_layer[j]];
                et = ORC_VACANCY_RECOMBINATION;
                double Ekin = 0;
                double EA = zf - E - Ekin;
                P = exp(-1 * EA / (kB * T_bg)) * freq;
            }
            if (element[i] == ORC_VACANCY && element[j] == ORC_O_EL) {
                double self_int_V = 0.0;
                if (charge[i] != 0) self_int_V = v_solve(r_dist, charge[i], sigma, k, q);
                et = ORC_VACANCY_DIFFUSION;
                double E = (charge[i] - charge[j]) * ((pb[i] + pc[i]) - (pb[j] + pc[j]) + self_int_V);
                double zf = E_Vdiff[site_layer[i]]; /* CPU path: layer of i (KMCProcess.cpp:134) */
                double Ekin = 0;
                double EA = zf - E - Ekin;
                P = exp(-1 * EA / (kB * T_bg)) * freq;
            }
            if (element[i] == ORC_OXYGEN_DEFECT && element[j] == ORC_DEFECT) {
                double self_int_V = 0.0;
                if (charge[i] != 0) self_int_V = v_solve(r_dist, 2, sigma, k, q);
                double E = (charge[i] - charge[j]) * ((pb[i] + pc[i]) - (pb[j] + pc[j]) - self_int_V);
                double zf = E_Odiff[site_layer[j]];
                et = ORC_ION_DIFFUSION;
                double Ekin = 0;
                double EA = zf - E - Ekin;
                P = exp(-1 * EA / (kB * T_bg)) * freq;
            }
        }
        event_type[idx] = et;
        event_prob[idx] = P;
    }
}

/* ------------------------------------------------------------------ RNG */
void orc_rng_seed(orc_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
static uint32_t mt_next(orc_rng *r) {
    if (r->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t yv = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
            r->mt[i] = r->mt[(i + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t yv = r->mt[r->idx++];
    yv ^= yv >> 11;
    yv ^= (yv << 7) & 0x9d2c5680u;
    yv ^= (yv << 15) & 0xefc60000u;
    yv ^= yv >> 18;
    return yv;
}
double orc_rng_uniform(orc_rng *r) {
    /* libstdc++ generate_canonical<double,53>(mt19937): k = 2 draws, R = 2^32 */
    double sum = (double)mt_next(r);
    sum += (double)mt_next(r) * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* ------------------------------------------------------------------ event loop */
long orc_select_event(long n, const double *event_prob, double u, double *Psum) {
    double acc = 0.0;
    int first = 1;
    for (long i = 0; i < n; ++i) {
        if (first) { acc = event_prob[i]; first = 0; } else acc = acc + event_prob[i];
    }
    *Psum = acc;
    double number = u * acc;
    double c = 0.0;
    for (long i = 0; i < n; ++i) {
        c = (i == 0) ? event_prob[0] : c + event_prob[i];
        if (c > number) return i;
    }
    return n;
}

static void apply_event(int type, int i, int j, int *element, int *charge) {
    int t;
    switch (type) { /* KMCProcess.cpp:187-256 */
    case ORC_VACANCY_GENERATION:
        element[i] = ORC_OXYGEN_DEFECT; charge[i] = -2; element[j] = ORC_VACANCY; charge[j] = 2; break;
    case ORC_VACANCY_RECOMBINATION:
        element[i] = ORC_DEFECT; charge[i] = 0; element[j] = ORC_O_EL; charge[j] = 0; break;
    case ORC_VACANCY_DIFFUSION:
    case ORC_ION_DIFFUSION:
        t = element[i]; element[i] = element[j]; element[j] = t;
        t = charge[i]; charge[i] = charge[j]; charge[j] = t; break;
    default: break;
    }
}

/* benchmark aid (bench.py's bounded CPU sample): stop the residence-time loop after this many
 * executed events; 0 = run the step to its end as the reference does */
static int g_event_limit = 0;
void orc_set_event_limit(int n) { g_event_limit = n > 0 ? n : 0; }

int orc_kmc_events(int N, int nn, const int *neigh_idx, int *event_type, double *event_prob,
                   int *element, int *charge, double freq, orc_rng *rng, double *event_time_out,
                   int *events, int max_events) {
    long total = (long)N * nn;
    /* Adding an exact 0.0 never changes a partial sum, so the strict left-to-right sum of
     * utils.h:91-99 over the whole table equals the same sum over the non-zero entries. */
    long nz = 0;
    for (long t = 0; t < total; ++t) if (event_prob[t] != 0.0) ++nz;
    long *nzi = (long *)malloc(sizeof(long) * (size_t)(nz > 0 ? nz : 1));
    double *cum = (double *)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1));
    nz = 0;
    for (long t = 0; t < total; ++t) if (event_prob[t] != 0.0) nzi[nz++] = t;
    double event_time = 0.0;
    int ne = 0;
    while (event_time < 1 / freq) {
        if (g_event_limit > 0 && ne >= g_event_limit) break;
        long live = 0;
        double acc = 0.0;
        for (long t = 0; t < nz; ++t) {
            double p = event_prob[nzi[t]];
            if (p == 0.0) continue;
            acc = (live == 0) ? p : acc + p;
            nzi[live] = nzi[t];
            cum[live] = acc;
            ++live;
        }
        nz = live;
        double Psum = acc;
        double number = orc_rng_uniform(rng) * Psum;
        long lo = 0, hi = nz; /* std::upper_bound: first cum > number */
        while (lo < hi) { long mid = lo + (hi - lo) / 2; if (cum[mid] > number) hi = mid; else lo = mid + 1; }
        if (lo < nz) {
            long idx = nzi[lo];
            int i = (int)(idx / nn), j = neigh_idx[idx];
            int type = event_type[idx];
            if (ne < max_events) { events[4 * ne] = (int)idx; events[4 * ne + 1] = i; events[4 * ne + 2] = j; events[4 * ne + 3] = type; }
            ++ne;
            apply_event(type, i, j, element, charge);
            /* conflicts (KMCProcess.cpp:330-352): rows i and j, plus every entry whose
             * neighbour slot holds i or j — by symmetry of the graph those live in the rows
             * of the neighbours of i and j. */
            for (int s = 0; s < nn; ++s) {
                event_type[(long)i * nn + s] = ORC_NULL_EVENT; event_prob[(long)i * nn + s] = 0.0;
                event_type[(long)j * nn + s] = ORC_NULL_EVENT; event_prob[(long)j * nn + s] = 0.0;
            }
            for (int side = 0; side < 2; ++side) {
                int c = side ? j : i;
                for (int s = 0; s < nn; ++s) {
                    int r = neigh_idx[(long)c * nn + s];
                    if (r < 0) break;
                    for (int s2 = 0; s2 < nn; ++s2) {
                        int jj = neigh_idx[(long)r * nn + s2];
                        if (jj == i || jj == j) { event_type[(long)r * nn + s2] = ORC_NULL_EVENT; event_prob[(long)r * nn + s2] = 0.0; }
                    }
                }
            }
        }
        event_time = -log(orc_rng_uniform(rng)) / Psum;
    }
    *event_time_out = event_time;
    free(nzi); free(cum);
    return ne;
}
