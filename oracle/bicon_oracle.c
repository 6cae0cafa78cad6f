/*
 * bicon_oracle.c -- CPU oracle (plain C) for the BiConMP centroidal biconvex solve.
 *
 * TEST INFRASTRUCTURE ONLY (see bicon_oracle.h).  PARITY UNPINNED against real Eigen.
 *
 * Build: gcc -O3 -ffp-contract=off -fPIC -shared  (see oracle/Makefile).  -ffp-contract=off is
 * REQUIRED: every multiply and add below is a separate IEEE-754 binary64 operation unless it is
 * written as MAD() with use_fma=1.
 *
 * Reference files restated here (all under /root/reference/iterative_supervised_learning/):
 *   [C]  src/dynamics/centroidal.cpp     [CH] include/dynamics/centroidal.hpp
 *   [P]  src/solvers/problem.cpp         [F]  src/solvers/fista.cpp   [FH] include/solvers/fista.hpp
 *   [B]  src/motion_planner/biconvex.cpp [BH] include/motion_planner/biconvex.hpp
 *
 * ---------------------------------------------------------------------------------------------
 * Canonical evaluation order (what "bit-exact" means for the CUDA path)
 * ---------------------------------------------------------------------------------------------
 * The reference writes Eigen expressions; Eigen's kernels fix an order of floating-point
 * operations that is not visible in the reference sources and cannot be checked here.  The order
 * below is the one Eigen 3.4's generic (non-vectorised-reduction) kernels are understood to use
 * for sparse work, plus a fixed reduction tree for dense norms/dots:
 *
 *  (1) sparse * dense vector, column-major A ([P]:55, [P]:47-48, [B]:98):  res_i accumulates
 *      A_ij * y_j over the structural non-zeros of row i in ASCENDING column order; the first
 *      product initialises the sum.  An empty row gives 0.0.  Explicit (structural) zeros take part.
 *  (2) gradient = (ATA*y) + ATbPk  ([P]:55): the product is formed first, then ATbPk is added.
 *  (3) ATA = 2*(Q + rho*A^T*A) ([P]:36):  S_ij = sum over rows k shared by columns i and j, in
 *      ASCENDING k, of (rho*A_ki)*A_kj; ATA_ij = 2*(Q_ij + S_ij) on the diagonal, 2*S_ij elsewhere.
 *  (4) ATbPk = 2*rho*A^T*bPk + q ([P]:38):  ((2*rho)*A_ki)*bPk_k summed over ASCENDING k, then + q_i.
 *  (5) dense reductions (norm, squaredNorm, dot, row*vec; [F]:16-18, [P]:47-48, [B]:102-111):
 *      grouped_sum() below.  The leaves are first summed in TRIPLES, (l0 + l1) + l2:
 *        - a vector indexed by contact forces (length 3*e*n): triple s = leaves 3s, 3s+1, 3s+2 (one
 *          3-D force vector);
 *        - a vector indexed by states or by constraint rows (length 9*(n+1)): triple s = 3t+a holds
 *          leaves 9t+a, 9t+3+a, 9t+6+a (component a of the CoM, of the velocity and of the angular
 *          momentum of knot t; for rows: the three dynamics rows of knot t that belong to axis a);
 *        - any other vector: every leaf is its own group.
 *      The group sums, in index order, are then summed by tree_sum(): blocks of 32, zero-padded,
 *      radix-2 tree with strides 16,8,4,2,1; block sums are combined the same way in groups of 32.
 *      (One group is what one GPU thread owns, one block is one warp.)  params.reduction = 32 selects
 *      the plain tree over the leaves instead -- a rounding VARIANT kept for the sensitivity test.
 *  (6) x.transpose()*Q*d with diagonal Q ([P]:47):  ((x_i*Q_ii)*d_i) summed by (5).
 *  (7) cwiseMin(ub).cwiseMax(lb) ([F]:10):  t = (ub < u) ? ub : u;  y = (t < lb) ? lb : t.
 *  (8) everything else is evaluated exactly as the C++ expression parses (left to right).
 * The GPU kernels (bunmpc_b200/csrc) implement the same order; tests require identical bits.
 */
#include "bicon_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define GRAV 9.81           /* literal in [C]:62,104 */
#define X_BLOCK 32
#define KIND_PLAIN 0        /* reduction layouts of rule (5) */
#define KIND_FORCE 1
#define KIND_STATE 2

#define ALWAYS_INLINE static inline __attribute__((always_inline))
#define MAD(FM, acc, a, b) ((FM) ? fma((a), (b), (acc)) : ((acc) + (a) * (b)))

void bicon_default_params(bicon_params *p)
{
    p->max_outer = 100;  /* abstract_cyclic_gen.py:663 kd.optimize(q, v, 100, 1) */
    p->max_inner = 150;  /* [BH]:154-156 */
    p->tol = 1e-5;       /* [BH]:158 */
    p->exit_tol = 1e-3;  /* [BH]:160 */
    p->beta = 1.5;       /* [FH]:54 */
    p->mu = 1.0;         /* [FH]:60 */
    p->use_fma = 0;
    p->reduction = 0;
    p->storage = 0;
}

/* ------------------------------------------------------------------ sparse patterns --------- */

typedef struct {
    int rows, cols, nnz;
    int *cp, *ri;          /* CSC: column pointers, row indices ascending within a column */
    int *rp, *cj, *map;    /* CSR view: row pointers, column indices ascending, CSR pos -> CSC pos */
} pattern;

typedef struct { int r, c; } rc_pair;

static int cmp_cr(const void *a, const void *b)
{
    const rc_pair *x = (const rc_pair *)a, *y = (const rc_pair *)b;
    if (x->c != y->c) return x->c - y->c;
    return x->r - y->r;
}

/* builds CSC+CSR from an unsorted list of (row, col) pairs, duplicates merged (coeffRef semantics) */
static void pattern_build(pattern *p, int rows, int cols, rc_pair *e, int ne)
{
    qsort(e, (size_t)ne, sizeof(rc_pair), cmp_cr);
    int nnz = 0;
    for (int i = 0; i < ne; ++i)
        if (i == 0 || e[i].r != e[i - 1].r || e[i].c != e[i - 1].c) e[nnz++] = e[i];
    p->rows = rows; p->cols = cols; p->nnz = nnz;
    p->cp = (int *)calloc((size_t)cols + 1, sizeof(int));
    p->ri = (int *)malloc(sizeof(int) * (size_t)nnz);
    p->rp = (int *)calloc((size_t)rows + 1, sizeof(int));
    p->cj = (int *)malloc(sizeof(int) * (size_t)nnz);
    p->map = (int *)malloc(sizeof(int) * (size_t)nnz);
    for (int i = 0; i < nnz; ++i) { p->cp[e[i].c + 1]++; p->ri[i] = e[i].r; p->rp[e[i].r + 1]++; }
    for (int c = 0; c < cols; ++c) p->cp[c + 1] += p->cp[c];
    for (int r = 0; r < rows; ++r) p->rp[r + 1] += p->rp[r];
    int *fill = (int *)calloc((size_t)rows, sizeof(int));
    for (int c = 0; c < cols; ++c)           /* ascending column => ascending within each row */
        for (int q = p->cp[c]; q < p->cp[c + 1]; ++q) {
            int r = p->ri[q], pos = p->rp[r] + fill[r]++;
            p->cj[pos] = c; p->map[pos] = q;
        }
    free(fill);
}

static void pattern_free(pattern *p)
{
    free(p->cp); free(p->ri); free(p->rp); free(p->cj); free(p->map);
}

/* position of (r, c) in the CSC arrays */
static int pattern_pos(const pattern *p, int r, int c)
{
    for (int q = p->cp[c]; q < p->cp[c + 1]; ++q) if (p->ri[q] == r) return q;
    return -1;
}

/* symbolic A^T A: CSR pattern of H plus, per entry, the list of (pos of A_ki, pos of A_kj), k ascending */
typedef struct {
    int nv, nnz;
    int *rp, *cj;
    int *pp, *pa, *pb;
} gram;

static void gram_build(gram *g, const pattern *A)
{
    int nv = A->cols;
    g->nv = nv;
    g->rp = (int *)calloc((size_t)nv + 1, sizeof(int));
    char *mark = (char *)calloc((size_t)nv, 1);
    int cap = 16 * nv + 16, nnz = 0, pcap = 64 * nv + 64, np = 0;
    g->cj = (int *)malloc(sizeof(int) * (size_t)cap);
    g->pp = (int *)malloc(sizeof(int) * ((size_t)cap + 1));
    g->pa = (int *)malloc(sizeof(int) * (size_t)pcap);
    g->pb = (int *)malloc(sizeof(int) * (size_t)pcap);
    for (int i = 0; i < nv; ++i) {
        /* candidate columns j: every column sharing a row with column i, plus the diagonal */
        mark[i] = 1;
        for (int q = A->cp[i]; q < A->cp[i + 1]; ++q) {
            int k = A->ri[q];
            for (int s = A->rp[k]; s < A->rp[k + 1]; ++s) mark[A->cj[s]] = 1;
        }
        for (int j = 0; j < nv; ++j) {
            if (!mark[j]) continue;
            mark[j] = 0;
            if (nnz + 1 >= cap) {
                cap *= 2;
                g->cj = (int *)realloc(g->cj, sizeof(int) * (size_t)cap);
                g->pp = (int *)realloc(g->pp, sizeof(int) * ((size_t)cap + 1));
            }
            g->cj[nnz] = j; g->pp[nnz] = np;
            int a = A->cp[i], ae = A->cp[i + 1], b = A->cp[j], be = A->cp[j + 1];
            while (a < ae && b < be) {                 /* merge: shared rows, ascending k */
                if (A->ri[a] < A->ri[b]) ++a;
                else if (A->ri[a] > A->ri[b]) ++b;
                else {
                    if (np >= pcap) {
                        pcap *= 2;
                        g->pa = (int *)realloc(g->pa, sizeof(int) * (size_t)pcap);
                        g->pb = (int *)realloc(g->pb, sizeof(int) * (size_t)pcap);
                    }
                    g->pa[np] = a; g->pb[np] = b; ++np; ++a; ++b;
                }
            }
            ++nnz;
        }
        g->rp[i + 1] = nnz;
    }
    g->pp[nnz] = np;
    g->nnz = nnz;
    free(mark);
}

static void gram_free(gram *g)
{
    free(g->rp); free(g->cj); free(g->pp); free(g->pa); free(g->pb);
}

/* ------------------------------------------------------------------ workspace --------------- */

struct bicon_ws {
    int n, e, nx, nf;
    pattern Ax, Af;
    gram Gf, Gx;                 /* Gram of A_x (force problem) and of A_f (state problem) */
    double *Ax_val, *Af_val;     /* CSC values */
    double *Ax_csr, *Af_csr;     /* CSR-ordered copies, refreshed by set_data */
    double *b_x, *b_f;
    double *Hf, *Hx;             /* ATA_ values in Gram CSR order */
    double *hf, *hx;             /* ATbPk_ */
    double *w;                   /* bPk_ */
    double *X, *F, *P;           /* prob_data_x.x_k, prob_data_f.x_k, P_k_ */
    double *y, *y1, *x1, *g, *d, *ynext;   /* max(nx, nf) each */
    double *e0, *e1, *e2, *e3, *e4, *e5;   /* reduction leaves */
    double *viol;
    /* positions of the A_f entries rewritten by compute_f_mat / update_x_init */
    int *pos_dt;     /* [n][3]  (9t+l, 9(t+1)+l+3) */
    int *pos_cr;     /* [n][6]  (6,1) (6,2) (7,0) (7,2) (8,0) (8,1) */
    int *pos_init;   /* [9] */
    /* positions of the A_x entries: [n][e][9]: vel k=0..2, then (6,by)(6,bz)(7,bx)(7,bz)(8,bx)(8,by) */
    int *pos_ax;
};

bicon_ws *bicon_ws_create(int n, int e)
{
    if (n < 1 || e < 1) return NULL;
    bicon_ws *ws = (bicon_ws *)calloc(1, sizeof(bicon_ws));
    int nx = 9 * (n + 1), nf = 3 * e * n;
    ws->n = n; ws->e = e; ws->nx = nx; ws->nf = nf;

    /* A_x structural pattern, [C]:67-82 */
    rc_pair *el = (rc_pair *)malloc(sizeof(rc_pair) * (size_t)(9 * e * n + 30 * n + 64));
    int ne = 0;
    for (int t = 0; t < n; ++t)
        for (int j = 0; j < e; ++j) {
            int base = 3 * e * t + 3 * j;
            for (int k = 0; k < 3; ++k) { el[ne].r = 9 * t + 3 + k; el[ne].c = base + k; ++ne; }
            el[ne].r = 9 * t + 6; el[ne].c = base + 1; ++ne;
            el[ne].r = 9 * t + 6; el[ne].c = base + 2; ++ne;
            el[ne].r = 9 * t + 7; el[ne].c = base + 0; ++ne;
            el[ne].r = 9 * t + 7; el[ne].c = base + 2; ++ne;
            el[ne].r = 9 * t + 8; el[ne].c = base + 0; ++ne;
            el[ne].r = 9 * t + 8; el[ne].c = base + 1; ++ne;
        }
    pattern_build(&ws->Ax, nx, nf, el, ne);

    /* A_f structural pattern: [C]:14-25 (identities), [C]:89-100 (dt, cross terms), [CH]:22-27 */
    ne = 0;
    for (int t = 0; t < n; ++t) {
        for (int l = 0; l < 9; ++l) {
            el[ne].r = 9 * t + l; el[ne].c = 9 * t + l; ++ne;
            el[ne].r = 9 * t + l; el[ne].c = 9 * (t + 1) + l; ++ne;
        }
        for (int l = 0; l < 3; ++l) { el[ne].r = 9 * t + l; el[ne].c = 9 * (t + 1) + l + 3; ++ne; }
        el[ne].r = 9 * t + 6; el[ne].c = 9 * t + 1; ++ne;
        el[ne].r = 9 * t + 6; el[ne].c = 9 * t + 2; ++ne;
        el[ne].r = 9 * t + 7; el[ne].c = 9 * t + 0; ++ne;
        el[ne].r = 9 * t + 7; el[ne].c = 9 * t + 2; ++ne;
        el[ne].r = 9 * t + 8; el[ne].c = 9 * t + 0; ++ne;
        el[ne].r = 9 * t + 8; el[ne].c = 9 * t + 1; ++ne;
    }
    for (int k = 0; k < 9; ++k) { el[ne].r = 9 * n + k; el[ne].c = k; ++ne; }
    pattern_build(&ws->Af, nx, nx, el, ne);
    free(el);

    gram_build(&ws->Gf, &ws->Ax);
    gram_build(&ws->Gx, &ws->Af);

    ws->Ax_val = (double *)calloc((size_t)ws->Ax.nnz, sizeof(double));
    ws->Af_val = (double *)calloc((size_t)ws->Af.nnz, sizeof(double));
    ws->Ax_csr = (double *)calloc((size_t)ws->Ax.nnz, sizeof(double));
    ws->Af_csr = (double *)calloc((size_t)ws->Af.nnz, sizeof(double));
    ws->b_x = (double *)calloc((size_t)nx, sizeof(double));
    ws->b_f = (double *)calloc((size_t)nx, sizeof(double));
    ws->Hf = (double *)calloc((size_t)ws->Gf.nnz, sizeof(double));
    ws->Hx = (double *)calloc((size_t)ws->Gx.nnz, sizeof(double));
    int mv = nx > nf ? nx : nf;
    ws->hf = (double *)calloc((size_t)nf, sizeof(double));
    ws->hx = (double *)calloc((size_t)nx, sizeof(double));
    ws->w = (double *)calloc((size_t)nx, sizeof(double));
    ws->X = (double *)calloc((size_t)nx, sizeof(double));
    ws->F = (double *)calloc((size_t)nf, sizeof(double));
    ws->P = (double *)calloc((size_t)nx, sizeof(double));
    double **v[] = { &ws->y, &ws->y1, &ws->x1, &ws->g, &ws->d, &ws->ynext,
                     &ws->e0, &ws->e1, &ws->e2, &ws->e3, &ws->e4, &ws->e5, &ws->viol };
    for (size_t i = 0; i < sizeof(v) / sizeof(v[0]); ++i) *v[i] = (double *)calloc((size_t)mv, sizeof(double));

    /* constant entries of A_f, [C]:14-25 */
    for (int t = 0; t < n; ++t)
        for (int l = 0; l < 9; ++l) {
            ws->Af_val[pattern_pos(&ws->Af, 9 * t + l, 9 * t + l)] = 1.0;
            ws->Af_val[pattern_pos(&ws->Af, 9 * t + l, 9 * (t + 1) + l)] = -1.0;
        }
    ws->pos_dt = (int *)malloc(sizeof(int) * (size_t)(3 * n));
    ws->pos_cr = (int *)malloc(sizeof(int) * (size_t)(6 * n));
    ws->pos_init = (int *)malloc(sizeof(int) * 9);
    ws->pos_ax = (int *)malloc(sizeof(int) * (size_t)(9 * e * n));
    for (int t = 0; t < n; ++t) {
        for (int l = 0; l < 3; ++l) ws->pos_dt[3 * t + l] = pattern_pos(&ws->Af, 9 * t + l, 9 * (t + 1) + l + 3);
        static const int cr_r[6] = { 6, 6, 7, 7, 8, 8 }, cr_c[6] = { 1, 2, 0, 2, 0, 1 };
        for (int q = 0; q < 6; ++q) ws->pos_cr[6 * t + q] = pattern_pos(&ws->Af, 9 * t + cr_r[q], 9 * t + cr_c[q]);
        for (int j = 0; j < e; ++j) {
            int base = 3 * e * t + 3 * j, *pa = ws->pos_ax + 9 * (e * t + j);
            for (int k = 0; k < 3; ++k) pa[k] = pattern_pos(&ws->Ax, 9 * t + 3 + k, base + k);
            for (int q = 0; q < 6; ++q) pa[3 + q] = pattern_pos(&ws->Ax, 9 * t + cr_r[q], base + cr_c[q]);
        }
    }
    for (int k = 0; k < 9; ++k) ws->pos_init[k] = pattern_pos(&ws->Af, 9 * n + k, k);
    return ws;
}

void bicon_ws_destroy(bicon_ws *ws)
{
    if (!ws) return;
    pattern_free(&ws->Ax); pattern_free(&ws->Af); gram_free(&ws->Gf); gram_free(&ws->Gx);
    free(ws->Ax_val); free(ws->Af_val); free(ws->Ax_csr); free(ws->Af_csr); free(ws->b_x); free(ws->b_f);
    free(ws->Hf); free(ws->Hx); free(ws->hf); free(ws->hx); free(ws->w); free(ws->X); free(ws->F); free(ws->P);
    free(ws->y); free(ws->y1); free(ws->x1); free(ws->g); free(ws->d); free(ws->ynext);
    free(ws->e0); free(ws->e1); free(ws->e2); free(ws->e3); free(ws->e4); free(ws->e5); free(ws->viol);
    free(ws->pos_dt); free(ws->pos_cr); free(ws->pos_init); free(ws->pos_ax);
    free(ws);
}

/* ------------------------------------------------------------------ reductions -------------- */

/* rule (5): one block of up to 32 leaves, zero padded, strides 16,8,4,2,1 */
static double block32(const double *v, int cnt)
{
    double s[32];
    for (int l = 0; l < 32; ++l) s[l] = l < cnt ? v[l] : 0.0;
    for (int off = 16; off > 0; off >>= 1)
        for (int l = 0; l < off; ++l) s[l] = s[l] + s[l + off];
    return s[0];
}

static double tree_sum(const double *v, int n, int blk)
{
    double part[1024];
    int nb = (n + blk - 1) / blk;
    if (nb < 1) return 0.0;
    for (int b = 0; b < nb; ++b) {
        int cnt = n - b * blk; if (cnt > blk) cnt = blk;
        part[b] = block32(v + (size_t)b * blk, cnt);
    }
    do {                                   /* block sums are always combined once more in groups of 32 */
        int nb2 = (nb + 31) / 32;
        for (int b = 0; b < nb2; ++b) {
            int cnt = nb - 32 * b; if (cnt > 32) cnt = 32;
            part[b] = block32(part + 32 * b, cnt);
        }
        nb = nb2;
    } while (nb > 1);
    return part[0];
}

/* rule (5): triples first, then the tree over the group sums */
static double grouped_sum(const double *v, int len, int kind)
{
    double grp[4096];
    if (kind == KIND_FORCE) {
        int ns = len / 3;
        for (int s = 0; s < ns; ++s) grp[s] = (v[3 * s] + v[3 * s + 1]) + v[3 * s + 2];
        return tree_sum(grp, ns, 32);
    }
    if (kind == KIND_STATE) {
        int nk = len / 9;
        for (int t = 0; t < nk; ++t)
            for (int a = 0; a < 3; ++a) grp[3 * t + a] = (v[9 * t + a] + v[9 * t + 3 + a]) + v[9 * t + 6 + a];
        return tree_sum(grp, 3 * nk, 32);
    }
    return tree_sum(v, len, 32);
}

/* rule (1): row i of a CSR matrix times y */
ALWAYS_INLINE double row_dot(const int FM, const int *rp, const int *cj, const double *val, int i, const double *y)
{
    int s = rp[i], e = rp[i + 1];
    if (s == e) return 0.0;
    double acc = val[s] * y[cj[s]];
    for (int q = s + 1; q < e; ++q) acc = MAD(FM, acc, val[q], y[cj[q]]);
    return acc;
}

/* ------------------------------------------------------------------ problem data ------------ */

/* params.storage = 1 (the GPU's BUNMPC_ARITH_MIXED mode): the entries of ATA_ and of A_ are kept in binary32
 * (round to nearest even) once set_data has formed them in binary64; every operation stays binary64. */
ALWAYS_INLINE double store_round(int storage, double v) { return storage ? (double)(float)v : v; }

/* ProblemData::set_data, [P]:31-39.  Q diagonal. */
ALWAYS_INLINE void set_data(const int FM, const int storage, const pattern *A, const double *Aval, double *Acsr,
                            const gram *G, double *H, double *h, double *w,
                            const double *b, const double *P, const double *Q, const double *q, double rho)
{
    for (int p = 0; p < A->nnz; ++p) Acsr[p] = Aval[A->map[p]];
    /* ATA_ = 2*(Q_ + rho_*A^T*A), rule (3) */
    for (int i = 0; i < G->nv; ++i)
        for (int p = G->rp[i]; p < G->rp[i + 1]; ++p) {
            int s = G->pp[p], e = G->pp[p + 1];
            double acc = 0.0;
            if (s < e) {
                acc = (rho * Aval[G->pa[s]]) * Aval[G->pb[s]];
                for (int r = s + 1; r < e; ++r) acc = MAD(FM, acc, rho * Aval[G->pa[r]], Aval[G->pb[r]]);
            }
            if (G->cj[p] == i) acc = Q[i] + acc;
            H[p] = 2 * acc;
        }
    /* bPk_ = -b_ + P_k_ */
    for (int k = 0; k < A->rows; ++k) w[k] = -b[k] + P[k];
    /* ATbPk_ = 2.0*rho_*A^T*bPk_ + q_, rule (4) */
    double two_rho = 2.0 * rho;
    for (int i = 0; i < A->cols; ++i) {
        int s = A->cp[i], e = A->cp[i + 1];
        double acc = 0.0;
        if (s < e) {
            acc = (two_rho * Aval[s]) * w[A->ri[s]];
            for (int r = s + 1; r < e; ++r) acc = MAD(FM, acc, two_rho * Aval[r], w[A->ri[r]]);
        }
        h[i] = acc + q[i];
    }
    if (storage) {
        for (int p = 0; p < G->nnz; ++p) H[p] = store_round(storage, H[p]);
        for (int p = 0; p < A->nnz; ++p) Acsr[p] = store_round(storage, Acsr[p]);
    }
}

typedef struct {
    int nv, nr, vkind, rkind, cone;   /* reduction layouts of the variable- and row-indexed sums, rule (5) */
    const int *hrp, *hcj; const double *H, *h;
    const int *arp, *acj; const double *A, *w;
    const double *Q, *q, *lb, *ub;
    double rho, mu, beta;
} fista_data;

/* FISTA::optimize + compute_step_length + SoC_projection, [F]:6-70.
 * x: prob_data.x_k (in: warm start, out: solution).  L: the FISTA object's L_ (in/out). */
ALWAYS_INLINE void fista(const int FM, bicon_ws *ws, const fista_data *D, double *x, double *L,
                         int max_iters, double tol, int *n_iters, int *n_ls)
{
    const int nv = D->nv, nr = D->nr;
    double *y = ws->y, *y1 = ws->y1, *x1 = ws->x1, *g = ws->g, *d = ws->d, *yn = ws->ynext;
    const double mu = D->mu;
    memcpy(y, x, sizeof(double) * (size_t)nv);                       /* [F]:30 */
    double t_k = 1.0;                                                /* [F]:31 */
    for (int it = 0; it < max_iters; ++it) {
        /* ---- compute_step_length, [F]:6-27 ---- */
        for (int i = 0; i < nv; ++i) g[i] = row_dot(FM, D->hrp, D->hcj, D->H, i, y) + D->h[i];   /* [P]:55 */
        double G_k_norm;
        for (;;) {
            if (!D->cone) {                                          /* [F]:10, rule (7) */
                for (int i = 0; i < nv; ++i) {
                    double u = y[i] - g[i] / (*L);
                    double tt = (D->ub[i] < u) ? D->ub[i] : u;
                    y1[i] = (tt < D->lb[i]) ? D->lb[i] : tt;
                }
            } else {                                                 /* SoC_projection, [F]:52-70 */
                for (int i = 0; i < nv; ++i) y1[i] = y[i] - g[i] / (*L);
                for (int i = 0; i < nv; i += 3) {
                    double a = y1[i], b = y1[i + 1];
                    double soc_norm = a * a + b * b;                 /* squaredNorm of the 2-segment */
                    double z = y1[i + 2];
                    if (soc_norm * mu < -z || z < 0) {
                        y1[i] = 0.0; y1[i + 1] = 0.0; y1[i + 2] = 0.0;
                    } else if (soc_norm > mu * z) {
                        double sc = ((mu * mu) * soc_norm + (mu * z)) / (((mu * mu) + 1) * soc_norm);
                        y1[i] = a * sc; y1[i + 1] = b * sc;
                        y1[i + 2] = (mu * soc_norm + z) / ((mu * mu) + 1);
                    }
                }
            }
            for (int i = 0; i < nv; ++i) {
                d[i] = y1[i] - y[i];                                 /* [F]:15 */
                ws->e0[i] = d[i] * d[i];
                ws->e1[i] = ((y1[i] + y[i]) * D->Q[i]) * (y1[i] - y[i]);     /* [P]:47, rule (6) */
                ws->e2[i] = D->q[i] * (y1[i] - y[i]);
                ws->e5[i] = g[i] * d[i];
            }
            G_k_norm = sqrt(grouped_sum(ws->e0, nv, D->vkind));           /* [F]:16 */
            for (int k = 0; k < nr; ++k) {                           /* [P]:48 */
                double r1 = row_dot(FM, D->arp, D->acj, D->A, k, y1) + D->w[k];
                double r0 = row_dot(FM, D->arp, D->acj, D->A, k, y) + D->w[k];
                ws->e3[k] = r1 * r1; ws->e4[k] = r0 * r0;
            }
            double t1 = grouped_sum(ws->e1, nv, D->vkind), t2 = grouped_sum(ws->e2, nv, D->vkind);
            double n1 = grouped_sum(ws->e3, nr, D->rkind), n0 = grouped_sum(ws->e4, nr, D->rkind);
            double obj = t1 + t2 + (D->rho) * (n1 - n0);             /* [P]:47-48 */
            double gd = grouped_sum(ws->e5, nv, D->vkind);
            if (obj > gd + ((*L) / 2) * (G_k_norm * G_k_norm)) {     /* [F]:17-19 */
                *L = D->beta * (*L);
                ++*n_ls;
            } else {
                memcpy(x1, y1, sizeof(double) * (size_t)nv);         /* [F]:23 */
                break;
            }
        }
        ++*n_iters;
        /* ---- [F]:34-48 ---- */
        double t_k_1 = 1.0 + sqrt(1 + 4 * t_k * t_k) / 2.0;          /* sic */
        double coef = (t_k - 1) / t_k_1;
        for (int i = 0; i < nv; ++i) yn[i] = MAD(FM, x1[i], coef, x1[i] - x[i]);
        memcpy(x, x1, sizeof(double) * (size_t)nv);
        if (G_k_norm < tol) break;
        memcpy(y, yn, sizeof(double) * (size_t)nv);
        t_k = t_k_1;
    }
}

/* CentroidalDynamics::compute_x_mat, [C]:57-84 */
static void compute_x_mat(bicon_ws *ws, const bicon_problem *p, const double *X)
{
    const int n = ws->n, e = ws->e;
    for (int t = 0; t < n; ++t) {
        double dt = p->dt[t];
        double *b = ws->b_x + 9 * t;
        const double *X0 = X + 9 * t, *X1 = X + 9 * (t + 1);
        b[3] = X1[3] - X0[3];
        b[4] = X1[4] - X0[4];
        b[5] = X1[5] - X0[5] + GRAV * dt;
        b[6] = X1[6] - X0[6];
        b[7] = X1[7] - X0[7];
        b[8] = X1[8] - X0[8];
        for (int j = 0; j < e; ++j) {
            const double *cp = p->cnt_plan + 4 * (e * t + j);
            double c = cp[0];
            const int *pa = ws->pos_ax + 9 * (e * t + j);
            ws->Ax_val[pa[0]] = c * (dt / p->m);
            ws->Ax_val[pa[1]] = c * (dt / p->m);
            ws->Ax_val[pa[2]] = c * (dt / p->m);
            ws->Ax_val[pa[3]] = c * (X0[2] - cp[3]) * dt;       /* (6, by) */
            ws->Ax_val[pa[4]] = -c * (X0[1] - cp[2]) * dt;      /* (6, bz) */
            ws->Ax_val[pa[5]] = -c * (X0[2] - cp[3]) * dt;      /* (7, bx) */
            ws->Ax_val[pa[6]] = c * (X0[0] - cp[1]) * dt;       /* (7, bz) */
            ws->Ax_val[pa[7]] = c * (X0[1] - cp[2]) * dt;       /* (8, bx) */
            ws->Ax_val[pa[8]] = -c * (X0[0] - cp[1]) * dt;      /* (8, by) */
        }
    }
}

/* CentroidalDynamics::compute_f_mat, [C]:86-127 */
static void compute_f_mat(bicon_ws *ws, const bicon_problem *p, const double *F)
{
    const int n = ws->n, e = ws->e;
    const double m = p->m;
    for (int t = 0; t < n; ++t) {
        double dt = p->dt[t];
        double *b = ws->b_f + 9 * t;
        const double *Ft = F + 3 * e * t;
        const double *cp = p->cnt_plan + 4 * (e * t);
        double *A = ws->Af_val;
        const int *pd = ws->pos_dt + 3 * t, *pc = ws->pos_cr + 6 * t;
        A[pd[0]] = dt; A[pd[1]] = dt; A[pd[2]] = dt;
        double c = cp[0];
        A[pc[0]] = -c * Ft[2] * dt;
        A[pc[1]] = c * Ft[1] * dt;
        A[pc[2]] = c * Ft[2] * dt;
        A[pc[3]] = -c * Ft[0] * dt;
        A[pc[4]] = -c * Ft[1] * dt;
        A[pc[5]] = c * Ft[0] * dt;
        b[3] = -c * Ft[0] * dt / m;
        b[4] = -c * Ft[1] * dt / m;
        b[5] = -c * Ft[2] * dt / m + GRAV * dt;
        b[6] = (c * Ft[1] * cp[3] - c * Ft[2] * cp[2]) * dt;
        b[7] = (c * Ft[2] * cp[1] - c * Ft[0] * cp[3]) * dt;
        b[8] = (c * Ft[0] * cp[2] - c * Ft[1] * cp[1]) * dt;
        for (int j = 1; j < e; ++j) {
            const double *f = Ft + 3 * j, *cq = cp + 4 * j;
            c = cq[0];
            A[pc[0]] += -c * f[2] * dt;
            A[pc[1]] += c * f[1] * dt;
            A[pc[2]] += c * f[2] * dt;
            A[pc[3]] += -c * f[0] * dt;
            A[pc[4]] += -c * f[1] * dt;
            A[pc[5]] += c * f[0] * dt;
            b[3] += -c * f[0] * dt / m;
            b[4] += -c * f[1] * dt / m;
            b[5] += -c * f[2] * dt / m;
            b[6] += (c * f[1] * cq[3] - c * f[2] * cq[2]) * dt;
            b[7] += (c * f[2] * cq[1] - c * f[0] * cq[3]) * dt;
            b[8] += (c * f[0] * cq[2] - c * f[1] * cq[1]) * dt;
        }
    }
}

/* BiConvexMP::optimize, [B]:80-120 */
ALWAYS_INLINE int solve_impl(const int FM, bicon_ws *ws, const bicon_problem *p, const bicon_params *prm,
                             bicon_result *out)
{
    const int n = ws->n, nx = ws->nx, nf = ws->nf;
    memset(ws->b_x, 0, sizeof(double) * (size_t)nx);                 /* [C]:29-30 */
    memset(ws->b_f, 0, sizeof(double) * (size_t)nx);                 /* [C]:12-13 */
    for (int k = 0; k < 9; ++k) {                                    /* update_x_init, [CH]:22-27 */
        ws->Af_val[ws->pos_init[k]] = 1.0;
        ws->b_f[9 * n + k] = p->x_init[k];
    }
    memcpy(ws->X, p->X0, sizeof(double) * (size_t)nx);               /* set_warm_start_vars, [BH]:66-70 */
    memcpy(ws->F, p->F0, sizeof(double) * (size_t)nf);
    memcpy(ws->P, p->P0, sizeof(double) * (size_t)nx);
    double L_f = p->L_f, L_x = p->L_x;
    int it_f = 0, it_x = 0, ls_f = 0, ls_x = 0, outer = 0, status = 1;
    double vnorm = 0.0;

    const int plain = prm->reduction == 32;
    const int kf = plain ? KIND_PLAIN : KIND_FORCE, kx = plain ? KIND_PLAIN : KIND_STATE;
    fista_data Df = { nf, nx, kf, kx, 1, ws->Gf.rp, ws->Gf.cj, ws->Hf, ws->hf,
                      ws->Ax.rp, ws->Ax.cj, ws->Ax_csr, ws->w, p->Qf, p->qf, NULL, NULL,
                      p->rho, prm->mu, prm->beta };
    fista_data Dx = { nx, nx, kx, kx, 0, ws->Gx.rp, ws->Gx.cj, ws->Hx, ws->hx,
                      ws->Af.rp, ws->Af.cj, ws->Af_csr, ws->w, p->Qx, p->qx, p->lbx, p->ubx,
                      p->rho, prm->mu, prm->beta };

    for (int i = 0; i < prm->max_outer; ++i) {
        /* optimizing for F, [B]:89-91 */
        compute_x_mat(ws, p, ws->X);
        set_data(FM, prm->storage, &ws->Ax, ws->Ax_val, ws->Ax_csr, &ws->Gf, ws->Hf, ws->hf, ws->w,
                 ws->b_x, ws->P, p->Qf, p->qf, p->rho);
        fista(FM, ws, &Df, ws->F, &L_f, prm->max_inner, prm->tol, &it_f, &ls_f);
        /* optimizing for X, [B]:94-96 */
        compute_f_mat(ws, p, ws->F);
        set_data(FM, prm->storage, &ws->Af, ws->Af_val, ws->Af_csr, &ws->Gx, ws->Hx, ws->hx, ws->w,
                 ws->b_f, ws->P, p->Qx, p->qx, p->rho);
        fista(FM, ws, &Dx, ws->X, &L_x, prm->max_inner, prm->tol, &it_x, &ls_x);
        /* dyn_violation = A_f * x_k - b_f;  P_k_ += dyn_violation, [B]:98-99 */
        for (int k = 0; k < nx; ++k) {
            ws->viol[k] = row_dot(FM, ws->Af.rp, ws->Af.cj, ws->Af_csr, k, ws->X) - ws->b_f[k];
            ws->P[k] += ws->viol[k];
            ws->e0[k] = ws->viol[k] * ws->viol[k];
        }
        vnorm = sqrt(grouped_sum(ws->e0, nx, kx));
        ++outer;
        if (out->viol_hist) out->viol_hist[i] = vnorm;               /* [B]:102-104 */
        if (isnan(vnorm)) { status = 2; break; }                      /* [B]:106-109 */
        if (vnorm < prm->exit_tol) { status = 0; break; }             /* [B]:111-114 */
    }
    memcpy(out->X, ws->X, sizeof(double) * (size_t)nx);
    memcpy(out->F, ws->F, sizeof(double) * (size_t)nf);
    memcpy(out->P, ws->P, sizeof(double) * (size_t)nx);
    out->L_f = L_f; out->L_x = L_x;
    out->outer_iters = outer; out->inner_f = it_f; out->inner_x = it_x; out->ls_f = ls_f; out->ls_x = ls_x;
    out->viol = vnorm; out->status = status;
    return 0;
}

static int solve_fm0(bicon_ws *ws, const bicon_problem *p, const bicon_params *prm, bicon_result *out)
{
    return solve_impl(0, ws, p, prm, out);
}
static int solve_fm1(bicon_ws *ws, const bicon_problem *p, const bicon_params *prm, bicon_result *out)
{
    return solve_impl(1, ws, p, prm, out);
}

int bicon_solve(bicon_ws *ws, const bicon_problem *p, const bicon_params *prm, bicon_result *out)
{
    if (!ws || !p || !prm || !out || p->n_col != ws->n || p->n_eff != ws->e) return -1;
    return prm->use_fma ? solve_fm1(ws, p, prm, out) : solve_fm0(ws, p, prm, out);
}

/* ------------------------------------------------------------------ builders ---------------- */

/* BiConvexMP::create_bound_constraints, [B]:27-58 (only the X box; the F box is dead code, fista.cpp:9-14) */
void bicon_create_bound_constraints(int n, int e, const double *cnt_plan, const double *b,
                                    double *lbx, double *ubx)
{
    int nx = 9 * (n + 1);
    for (int i = 0; i < nx; ++i) { lbx[i] = -1 * INFINITY * 1.0; ubx[i] = INFINITY * 1.0; }
    for (int i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int j = 0; j < e; ++j) sum += cnt_plan[4 * (e * i + j)];
        if (sum > 0) {
            for (int k = 0; k < 3; ++k) {
                double mx = cnt_plan[4 * (e * i) + 1 + k], mn = mx;
                for (int j = 1; j < e; ++j) {
                    double v = cnt_plan[4 * (e * i + j) + 1 + k];
                    if (v > mx) mx = v;
                    if (v < mn) mn = v;
                }
                lbx[9 * i + k] = mx + b[6 * i + k];
                ubx[9 * i + k] = mn + b[6 * i + 3 + k];
            }
        }
    }
}

/* BiConvexMP::create_cost_X, [B]:60-72 */
void bicon_create_cost_X(int n, const double *W_X, const double *W_X_ter, const double *X_ter,
                         const double *X_nom, double *Qx, double *qx)
{
    for (int i = 0; i < 9 * n; ++i) { Qx[i] = W_X[i]; qx[i] = -2 * (X_nom[i] * W_X[i]); }
    for (int k = 0; k < 9; ++k) { Qx[9 * n + k] = W_X_ter[k]; qx[9 * n + k] = -2 * (X_ter[k] * W_X_ter[k]); }
}

/* return_A_x / return_b_x, [BH]:30-38 */
void bicon_dense_x_mat(int n, int e, double m, const double *cnt_plan, const double *dt,
                       const double *X, double *A_x, double *b_x)
{
    bicon_ws *ws = bicon_ws_create(n, e);
    bicon_problem p; memset(&p, 0, sizeof(p));
    p.n_col = n; p.n_eff = e; p.m = m; p.cnt_plan = cnt_plan; p.dt = dt;
    compute_x_mat(ws, &p, X);
    memset(A_x, 0, sizeof(double) * (size_t)ws->nx * (size_t)ws->nf);
    for (int c = 0; c < ws->nf; ++c)
        for (int q = ws->Ax.cp[c]; q < ws->Ax.cp[c + 1]; ++q)
            A_x[(size_t)ws->Ax.ri[q] * ws->nf + c] = ws->Ax_val[q];
    memcpy(b_x, ws->b_x, sizeof(double) * (size_t)ws->nx);
    bicon_ws_destroy(ws);
}

/* return_A_f / return_b_f, [BH]:41-51 */
void bicon_dense_f_mat(int n, int e, double m, const double *cnt_plan, const double *dt,
                       const double *F, const double *x_init, double *A_f, double *b_f)
{
    bicon_ws *ws = bicon_ws_create(n, e);
    bicon_problem p; memset(&p, 0, sizeof(p));
    p.n_col = n; p.n_eff = e; p.m = m; p.cnt_plan = cnt_plan; p.dt = dt;
    compute_f_mat(ws, &p, F);
    for (int k = 0; k < 9; ++k) { ws->Af_val[ws->pos_init[k]] = 1.0; ws->b_f[9 * n + k] = x_init[k]; }
    memset(A_f, 0, sizeof(double) * (size_t)ws->nx * (size_t)ws->nx);
    for (int c = 0; c < ws->nx; ++c)
        for (int q = ws->Af.cp[c]; q < ws->Af.cp[c + 1]; ++q)
            A_f[(size_t)ws->Af.ri[q] * ws->nx + c] = ws->Af_val[q];
    memcpy(b_f, ws->b_f, sizeof(double) * (size_t)ws->nx);
    bicon_ws_destroy(ws);
}

/* ------------------------------------------------------------------ batch driver ------------ */

typedef struct {
    int lo, hi, n, e;
    const double *m, *rho, *x_init, *cnt_plan, *dt, *Qx, *qx, *Qf, *qf, *lbx, *ubx, *X0, *F0, *P0, *L_in;
    const bicon_params *prm;
    double *X, *F, *P, *L_out, *viol;
    int *iters, *status;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *J = (batch_job *)arg;
    int n = J->n, e = J->e, nx = 9 * (n + 1), nf = 3 * e * n;
    bicon_ws *ws = bicon_ws_create(n, e);
    for (int b = J->lo; b < J->hi; ++b) {
        bicon_problem p;
        p.n_col = n; p.n_eff = e; p.m = J->m[b]; p.rho = J->rho[b];
        p.x_init = J->x_init + 9 * (size_t)b;
        p.cnt_plan = J->cnt_plan + (size_t)4 * e * n * b;
        p.dt = J->dt + (size_t)n * b;
        p.Qx = J->Qx + (size_t)nx * b; p.qx = J->qx + (size_t)nx * b;
        p.Qf = J->Qf + (size_t)nf * b; p.qf = J->qf + (size_t)nf * b;
        p.lbx = J->lbx + (size_t)nx * b; p.ubx = J->ubx + (size_t)nx * b;
        p.X0 = J->X0 + (size_t)nx * b; p.F0 = J->F0 + (size_t)nf * b; p.P0 = J->P0 + (size_t)nx * b;
        p.L_f = J->L_in[2 * b]; p.L_x = J->L_in[2 * b + 1];
        bicon_result r;
        r.X = J->X + (size_t)nx * b; r.F = J->F + (size_t)nf * b; r.P = J->P + (size_t)nx * b;
        r.viol_hist = NULL;
        bicon_solve(ws, &p, J->prm, &r);
        J->L_out[2 * b] = r.L_f; J->L_out[2 * b + 1] = r.L_x;
        int *it = J->iters + 5 * (size_t)b;
        it[0] = r.outer_iters; it[1] = r.inner_f; it[2] = r.inner_x; it[3] = r.ls_f; it[4] = r.ls_x;
        J->viol[b] = r.viol; J->status[b] = r.status;
    }
    bicon_ws_destroy(ws);
    return NULL;
}

int bicon_solve_batch(int B, int n_col, int n_eff, const double *m, const double *rho,
                      const double *x_init, const double *cnt_plan, const double *dt,
                      const double *Qx, const double *qx, const double *Qf, const double *qf,
                      const double *lbx, const double *ubx,
                      const double *X0, const double *F0, const double *P0,
                      const double *L_in, const bicon_params *prm, int n_threads,
                      double *X, double *F, double *P, double *L_out, int *iters,
                      double *viol, int *status)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    batch_job *jobs = (batch_job *)calloc((size_t)n_threads, sizeof(batch_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        batch_job *J = &jobs[t];
        J->lo = (int)((long long)B * t / n_threads); J->hi = (int)((long long)B * (t + 1) / n_threads);
        J->n = n_col; J->e = n_eff;
        J->m = m; J->rho = rho; J->x_init = x_init; J->cnt_plan = cnt_plan; J->dt = dt;
        J->Qx = Qx; J->qx = qx; J->Qf = Qf; J->qf = qf; J->lbx = lbx; J->ubx = ubx;
        J->X0 = X0; J->F0 = F0; J->P0 = P0; J->L_in = L_in; J->prm = prm;
        J->X = X; J->F = F; J->P = P; J->L_out = L_out; J->viol = viol; J->iters = iters; J->status = status;
        if (n_threads == 1) batch_worker(J);
        else pthread_create(&th[t], NULL, batch_worker, J);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
    return 0;
}
