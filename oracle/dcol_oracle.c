/*
 * dcol_oracle.c — TEST INFRASTRUCTURE.  CPU restatement of the reference's proximity path.
 *
 * This file is the parity oracle for the CUDA kernels.  It restates, operation by operation
 * and in the same order, what the reference's Python/NumPy/SciPy code computes for one pair
 * of primitives; each function cites the reference file:line it follows (paths relative to
 * the reference root).  It is pinned against outputs of the reference itself, generated in
 * the build container by oracle/gen_golden.py and committed under tests/golden/
 * (tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load it; the product (dcol_trajectory_optimization_b200/) never does.
 *
 * Third-party arithmetic the reference relies on and that is restated here:
 *   - LAPACK dpotrf / dpotrs / dtrtrs through numpy.linalg.cholesky and
 *     scipy.linalg.{cholesky, cho_factor, cho_solve, solve_triangular} (SciPy 1.18.1 on
 *     OpenBLAS 0.3.30 in the build container; unpinned in the reference's requirements.txt:5).
 *     Unblocked Cholesky (n <= 8), pivot test `ajj <= 0` (OpenBLAS potf2: a NaN pivot is NOT
 *     an error), check_finite=True raising ValueError.
 *   - scipy.optimize.approx_fprime: 2-point forward difference, absolute step, divisor
 *     (x+h)-x (scipy/optimize/_numdiff.py approx_derivative / _dense_difference).
 * Summation order inside BLAS dot products / gemv is not reproducible bit for bit (SIMD
 * kernels); the restatement sums left to right.  SURVEY.md section 0 measured that 1-ulp
 * perturbations move alpha by <= 2e-14 and never flip an iteration count.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include "../include/dcol.h"

#define MAXN 8
#define MAXQ 4
#define MAXM (2 * DCOL_MAX_FACES + 2 * MAXQ)

/* ------------------------------------------------------------------------------------------ */
/* per-primitive blocks, dense, v = 4 + extras columns                                          */
typedef struct {
    int n_ort, n_soc, v;
    double G_ort[DCOL_MAX_FACES][MAXN], h_ort[DCOL_MAX_FACES];
    double G_soc[MAXQ][MAXN], h_soc[MAXQ];
} prim_blocks;

typedef struct {
    int n, m, n_ort, q1, q2;
    double c[MAXN], G[MAXM][MAXN], h[MAXM];
} conic_problem;

/* primitives/problem_matrices.py:213-251 */
static void dcm_from_mrp(const double p[3], double Q[3][3])
{
    double p1 = p[0], p2 = p[1], p3 = p[2];
    double t = p1 * p1 + p2 * p2 + p3 * p3 + 1.0;
    double den = t * t;
    double a = 4.0 * (p1 * p1) + 4.0 * (p2 * p2) + 4.0 * (p3 * p3) - 4.0;
    Q[0][0] = (-((8.0 * (p2 * p2) + 8.0 * (p3 * p3)) / den - 1.0)) * den;
    Q[0][1] = 8.0 * p1 * p2 + p3 * a;
    Q[0][2] = 8.0 * p1 * p3 - p2 * a;
    Q[1][0] = 8.0 * p1 * p2 - p3 * a;
    Q[1][1] = (-((8.0 * (p1 * p1) + 8.0 * (p3 * p3)) / den - 1.0)) * den;
    Q[1][2] = 8.0 * p2 * p3 + p1 * a;
    Q[2][0] = 8.0 * p1 * p3 + p2 * a;
    Q[2][1] = 8.0 * p2 * p3 - p1 * a;
    Q[2][2] = (-((8.0 * (p1 * p1) + 8.0 * (p2 * p2)) / den - 1.0)) * den;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Q[i][j] /= den;
}

static void mat3_mul(const double A[3][3], const double B[3][3], double C[3][3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}

/* Blocks of one primitive given the ADJUSTED pose r' = r + Q r_offset, Q' = Q Q_offset
 * (problem_matrices.py:272-364 computes those and dispatches to the builders below). */
static void blocks_from_adjusted(const dcol_shape* sh, const double* A, const double* b, const double r[3],
                                 const double Q[3][3], prim_blocks* out)
{
    memset(out, 0, sizeof(*out));
    const double bx[3] = { Q[0][0], Q[1][0], Q[2][0] }; /* n_Q_b @ [1,0,0] */
    switch (sh->type) {
    case DCOL_POLYTOPE: { /* problem_matrices.py:181-209 */
        int f = sh->n_faces;
        out->n_ort = f; out->n_soc = 0; out->v = 4;
        for (int i = 0; i < f; ++i) {
            const double* a = A + 3 * (sh->face_off + i);
            for (int j = 0; j < 3; ++j) out->G_ort[i][j] = a[0] * Q[j][0] + a[1] * Q[j][1] + a[2] * Q[j][2];
            out->G_ort[i][3] = -b[sh->face_off + i];
            out->h_ort[i] = out->G_ort[i][0] * r[0] + out->G_ort[i][1] * r[1] + out->G_ort[i][2] * r[2];
        }
        break;
    }
    case DCOL_SPHERE: /* problem_matrices.py:151-178 */
        out->n_ort = 0; out->n_soc = 4; out->v = 4;
        out->G_soc[0][3] = -sh->R;
        for (int j = 0; j < 3; ++j) { out->G_soc[1 + j][j] = -1.0; out->h_soc[1 + j] = -r[j]; }
        break;
    case DCOL_CONE: { /* problem_matrices.py:125-148 */
        double tb = tan(sh->beta);
        double E[3] = { tb, 1.0, 1.0 };
        out->n_ort = 1; out->n_soc = 3; out->v = 4;
        for (int i = 0; i < 3; ++i) {
            double EQt[3];
            for (int j = 0; j < 3; ++j) EQt[j] = E[i] * Q[j][i];
            for (int j = 0; j < 3; ++j) out->G_soc[i][j] = -EQt[j];
            out->h_soc[i] = (-EQt[0]) * r[0] + (-EQt[1]) * r[1] + (-EQt[2]) * r[2];
        }
        out->G_soc[0][3] = -(tb * 3.0 * sh->H / 4.0);
        out->G_soc[1][3] = -0.0; out->G_soc[2][3] = -0.0;
        for (int j = 0; j < 3; ++j) out->G_ort[0][j] = bx[j];
        out->G_ort[0][3] = -sh->H / 4.0;
        out->h_ort[0] = bx[0] * r[0] + bx[1] * r[1] + bx[2] * r[2];
        break;
    }
    case DCOL_CAPSULE:   /* problem_matrices.py:4-44  */
    case DCOL_CYLINDER: { /* problem_matrices.py:47-87 */
        out->n_soc = 4; out->v = 5;
        out->G_soc[0][3] = -sh->R;
        for (int j = 0; j < 3; ++j) { out->G_soc[1 + j][j] = -1.0; out->G_soc[1 + j][4] = bx[j]; out->h_soc[1 + j] = -r[j]; }
        out->G_ort[0][3] = -sh->L / 2.0; out->G_ort[0][4] = 1.0;
        out->G_ort[1][3] = -sh->L / 2.0; out->G_ort[1][4] = -1.0;
        out->n_ort = 2;
        if (sh->type == DCOL_CYLINDER) {
            double d = bx[0] * r[0] + bx[1] * r[1] + bx[2] * r[2];
            for (int j = 0; j < 3; ++j) { out->G_ort[2][j] = -bx[j]; out->G_ort[3][j] = bx[j]; }
            out->G_ort[2][3] = -sh->L / 2.0; out->G_ort[3][3] = -sh->L / 2.0;
            out->h_ort[2] = -d; out->h_ort[3] = d;
            out->n_ort = 4;
        }
        break;
    }
    case DCOL_ELLIPSOID: { /* EXTENSION, not in the reference's code (Report.pdf sec. 3.1.5 eq. 27): parity unpinned.
                            * || U Q'^T (x - r') || <= alpha with U = diag(1/R, 1/L, 1/H) */
        const double ia[3] = { 1.0 / sh->R, 1.0 / sh->L, 1.0 / sh->H };
        out->n_ort = 0; out->n_soc = 4; out->v = 4;
        out->G_soc[0][3] = -1.0;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) out->G_soc[1 + i][j] = -(ia[i] * Q[j][i]);
            out->h_soc[1 + i] = out->G_soc[1 + i][0] * r[0] + out->G_soc[1 + i][1] * r[1] + out->G_soc[1 + i][2] * r[2];
        }
        break;
    }
    case DCOL_POLYGON: { /* problem_matrices.py:90-120 */
        int f = sh->n_faces;
        out->n_ort = f; out->n_soc = 4; out->v = 6;
        for (int i = 0; i < f; ++i) {
            out->G_ort[i][3] = -b[sh->face_off + i];
            out->G_ort[i][4] = A[3 * (sh->face_off + i) + 0];
            out->G_ort[i][5] = A[3 * (sh->face_off + i) + 1];
        }
        out->G_soc[0][3] = -sh->R;
        for (int j = 0; j < 3; ++j) {
            out->G_soc[1 + j][j] = -1.0;
            out->G_soc[1 + j][4] = Q[j][0];
            out->G_soc[1 + j][5] = Q[j][1];
            out->h_soc[1 + j] = -r[j];
        }
        break;
    }
    default: break;
    }
}

/* problem_matrices.py:255-364: pose -> adjusted pose -> blocks */
static void problem_matrices(const dcol_shape* sh, const double* A, const double* b, const double r[3],
                             const double p[3], prim_blocks* out)
{
    double Q[3][3], Qo[3][3], Qp[3][3], rp[3];
    dcm_from_mrp(p, Q);
    memcpy(Qo, sh->Q_offset, sizeof(Qo));
    for (int i = 0; i < 3; ++i)
        rp[i] = r[i] + (Q[i][0] * sh->r_offset[0] + Q[i][1] * sh->r_offset[1] + Q[i][2] * sh->r_offset[2]);
    mat3_mul(Q, Qo, Qp);
    if (sh->type == DCOL_SPHERE) memcpy(Qp, Q, sizeof(Qp)); /* the sphere builder ignores orientation */
    blocks_from_adjusted(sh, A, b, rp, Qp, out);
}

/* EXTENSION switch (tests of DCOL_FIX_CASE4 only): assemble case 4 with the first primitive padded too */
static int g_fix_case4 = 0;
void dcol_oracle_set_fix_case4(int on) { g_fix_case4 = on; }

/* primitives/combine_problem_matrices.py:3-70; returns 0 or DCOL_STATUS_UNSUPPORTED (case 4 raises) */
static int combine_problem_matrices(const prim_blocks* B1, const prim_blocks* B2, conic_problem* P)
{
    int v1 = B1->v, v2 = B2->v;
    if (v1 > 4 && v2 > 4) {
        if (!g_fix_case4) return DCOL_STATUS_UNSUPPORTED; /* np.vstack of ragged widths -> ValueError */
        /* columns [x, alpha, extras1, extras2] (lines 58-67 place the second primitive's extras after the
         * first's; the fix is to pad the first primitive's blocks to the same width) */
        memset(P, 0, sizeof(*P));
        int e1 = v1 - 4, e2 = v2 - 4, row = 0;
        P->n = 4 + e1 + e2; P->n_ort = B1->n_ort + B2->n_ort; P->q1 = B1->n_soc; P->q2 = B2->n_soc;
        P->m = P->n_ort + P->q1 + P->q2;
        P->c[3] = 1.0;
        for (int i = 0; i < B1->n_ort; ++i, ++row) { memcpy(P->G[row], B1->G_ort[i], sizeof(double) * v1); P->h[row] = B1->h_ort[i]; }
        for (int i = 0; i < B2->n_ort; ++i, ++row) {
            memcpy(P->G[row], B2->G_ort[i], sizeof(double) * 4);
            memcpy(P->G[row] + 4 + e1, B2->G_ort[i] + 4, sizeof(double) * e2);
            P->h[row] = B2->h_ort[i];
        }
        for (int i = 0; i < B1->n_soc; ++i, ++row) { memcpy(P->G[row], B1->G_soc[i], sizeof(double) * v1); P->h[row] = B1->h_soc[i]; }
        for (int i = 0; i < B2->n_soc; ++i, ++row) {
            memcpy(P->G[row], B2->G_soc[i], sizeof(double) * 4);
            memcpy(P->G[row] + 4 + e1, B2->G_soc[i] + 4, sizeof(double) * e2);
            P->h[row] = B2->h_soc[i];
        }
        return 0;
    }
    memset(P, 0, sizeof(*P));
    int n = v1 > v2 ? v1 : v2;
    P->n = n; P->n_ort = B1->n_ort + B2->n_ort; P->q1 = B1->n_soc; P->q2 = B2->n_soc;
    P->m = P->n_ort + P->q1 + P->q2;
    P->c[3] = 1.0;
    int row = 0;
    /* rows [ort1; ort2; soc1; soc2]; each primitive's own columns are x(3), alpha, then ITS extras,
     * zero-padded on the right (cases 1-3, lines 34-56) */
    for (int i = 0; i < B1->n_ort; ++i, ++row) { memcpy(P->G[row], B1->G_ort[i], sizeof(double) * v1); P->h[row] = B1->h_ort[i]; }
    for (int i = 0; i < B2->n_ort; ++i, ++row) { memcpy(P->G[row], B2->G_ort[i], sizeof(double) * v2); P->h[row] = B2->h_ort[i]; }
    for (int i = 0; i < B1->n_soc; ++i, ++row) { memcpy(P->G[row], B1->G_soc[i], sizeof(double) * v1); P->h[row] = B1->h_soc[i]; }
    for (int i = 0; i < B2->n_soc; ++i, ++row) { memcpy(P->G[row], B2->G_soc[i], sizeof(double) * v2); P->h[row] = B2->h_soc[i]; }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                         */
static double dot(const double* a, const double* b, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
static int all_finite(const double* a, int n)
{
    for (int i = 0; i < n; ++i) if (!isfinite(a[i])) return 0;
    return 1;
}
static double py_min(double a, double b) { return (b < a) ? b : a; } /* Python min(a, b) */
static double py_max(double a, double b) { return (b > a) ? b : a; } /* Python max(a, b) */

/* Unblocked Cholesky, lower: A = L L^T (LAPACK dpotf2 'L' as OpenBLAS implements it).
 * Returns 0, or j+1 if the j-th pivot is <= 0 (NaN passes, as in OpenBLAS). */
static int chol_lower(double A[MAXN][MAXN], int n)
{
    for (int j = 0; j < n; ++j) {
        double ajj = A[j][j] - dot(A[j], A[j], j);
        if (ajj <= 0.0) return j + 1;
        ajj = sqrt(ajj);
        A[j][j] = ajj;
        for (int i = j + 1; i < n; ++i) A[i][j] = (A[i][j] - dot(A[i], A[j], j)) / ajj;
    }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) A[i][j] = 0.0; /* numpy.linalg.cholesky zeroes the upper triangle */
    return 0;
}

/* Unblocked Cholesky, upper: A = U^T U (scipy.linalg.cholesky / cho_factor default lower=False). */
static int chol_upper(double* A, int ld, int n)
{
    for (int j = 0; j < n; ++j) {
        double ajj = A[j * ld + j];
        for (int k = 0; k < j; ++k) ajj -= A[k * ld + j] * A[k * ld + j];
        if (ajj <= 0.0) return j + 1;
        ajj = sqrt(ajj);
        A[j * ld + j] = ajj;
        for (int i = j + 1; i < n; ++i) {
            double v = A[j * ld + i];
            for (int k = 0; k < j; ++k) v -= A[k * ld + j] * A[k * ld + i];
            A[j * ld + i] = v / ajj;
        }
    }
    return 0;
}

/* cho_solve((U, False), b): U^T y = b, U x = y */
static void cho_solve_upper(const double* U, int ld, int n, double* b)
{
    for (int i = 0; i < n; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= U[k * ld + i] * b[k];
        b[i] = v / U[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double v = b[i];
        for (int k = i + 1; k < n; ++k) v -= U[i * ld + k] * b[k];
        b[i] = v / U[i * ld + i];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cone algebra (proximity/pdip.py)                                                            */
typedef struct { int n_ort, q1, q2, m; } cone_idx;

/* pdip.py:7-22 */
static double ort_linesearch(const double* x, const double* dx, int n)
{
    double alpha = 1.0;
    for (int i = 0; i < n; ++i)
        if (dx[i] < 0.0) alpha = py_min(alpha, -x[i] / dx[i]);
    return alpha;
}

/* pdip.py:25-52 */
static double soc_linesearch(const double* y, const double* d, int q)
{
    double nu = py_max(y[0] * y[0] - dot(y + 1, y + 1, q - 1), 1e-25);
    double zeta = y[0] * d[0] - dot(y + 1, d + 1, q - 1);
    double snu = sqrt(nu);
    double rho0 = zeta / nu;
    double rho_v[MAXQ];
    double coef = (zeta / snu + d[0]) / (y[0] / snu + 1.0);
    for (int i = 1; i < q; ++i) rho_v[i - 1] = d[i] / snu - coef * (y[i] / nu);
    double nrm = sqrt(dot(rho_v, rho_v, q - 1));
    if (nrm > rho0) return py_min(1.0, 1.0 / (nrm - rho0));
    return 1.0;
}

/* pdip.py:55-85 */
static double linesearch(const double* x, const double* dx, const cone_idx* K)
{
    double alpha = 1.0;
    if (K->n_ort > 0) alpha = py_min(alpha, ort_linesearch(x, dx, K->n_ort));
    if (K->q1 > 0) alpha = py_min(alpha, soc_linesearch(x + K->n_ort, dx + K->n_ort, K->q1));
    if (K->q2 > 0) alpha = py_min(alpha, soc_linesearch(x + K->n_ort + K->q1, dx + K->n_ort + K->q1, K->q2));
    return alpha;
}

/* pdip.py:165-200 */
static void soc_cone_product(const double* u, const double* v, int q, double* out)
{
    if (q == 0) return;
    out[0] = dot(u, v, q);
    for (int i = 1; i < q; ++i) out[i] = u[0] * v[i] + v[0] * u[i];
}

/* pdip.py:335-370 */
static void cone_product(const double* s, const double* z, const cone_idx* K, double* out)
{
    for (int i = 0; i < K->n_ort; ++i) out[i] = s[i] * z[i];
    soc_cone_product(s + K->n_ort, z + K->n_ort, K->q1, out + K->n_ort);
    soc_cone_product(s + K->n_ort + K->q1, z + K->n_ort + K->q1, K->q2, out + K->n_ort + K->q1);
}

/* pdip.py:88-122 */
static void inverse_soc_cone_product(const double* u, const double* w, int q, double* out)
{
    if (q == 0) return;
    double rho = u[0] * u[0] - dot(u + 1, u + 1, q - 1);
    double nu = dot(u + 1, w + 1, q - 1);
    double scalar_part = u[0] * w[0] - nu;
    double inv = 1.0 / rho;
    out[0] = inv * scalar_part;
    double c1 = nu / u[0] - w[0], c2 = rho / u[0];
    for (int i = 1; i < q; ++i) out[i] = inv * (c1 * u[i] + c2 * w[i]);
}

/* pdip.py:125-162 */
static void inverse_cone_product(const double* lam, const double* v, const cone_idx* K, double* out)
{
    for (int i = 0; i < K->n_ort; ++i) out[i] = v[i] / lam[i];
    inverse_soc_cone_product(lam + K->n_ort, v + K->n_ort, K->q1, out + K->n_ort);
    inverse_soc_cone_product(lam + K->n_ort + K->q1, v + K->n_ort + K->q1, K->q2, out + K->n_ort + K->q1);
}

/* pdip.py:203-235 */
static void gen_e(const cone_idx* K, double* e)
{
    for (int i = 0; i < K->m; ++i) e[i] = 0.0;
    for (int i = 0; i < K->n_ort; ++i) e[i] = 1.0;
    if (K->q1 > 0) e[K->n_ort] = 1.0;
    if (K->q2 > 0) e[K->n_ort + K->q1] = 1.0;
}

/* pdip.py:237-287 */
static void bring2cone(double* r, const cone_idx* K)
{
    double alpha = -1.0;
    int any_nonpos = 0;
    double mn = INFINITY;
    int mn_nan = 0;
    for (int i = 0; i < K->n_ort; ++i) {
        if (r[i] <= 0.0) any_nonpos = 1;
        if (isnan(r[i])) mn_nan = 1;
        if (r[i] < mn) mn = r[i];
    }
    if (any_nonpos) alpha = mn_nan ? NAN : -mn; /* np.min propagates NaN */
    int off = K->n_ort;
    for (int k = 0; k < 2; ++k) {
        int q = k == 0 ? K->q1 : K->q2;
        if (q > 0) {
            double res = r[off] - sqrt(dot(r + off + 1, r + off + 1, q - 1));
            if (res <= 0.0) alpha = py_max(alpha, -res);
        }
        off += q;
    }
    if (alpha < 0.0) return;
    double e[MAXM];
    gen_e(K, e);
    for (int i = 0; i < K->m; ++i) r[i] = r[i] + (1.0 + alpha) * e[i];
}

/* ------------------------------------------------------------------------------------------ */
/* Nesterov-Todd scaling (proximity/NT/NT_scaling.py)                                          */
typedef struct {
    double ort[2 * DCOL_MAX_FACES];
    double soc[2][MAXQ * MAXQ];      /* W, dense row-major q x q      */
    double soc_fact[2][MAXQ * MAXQ]; /* upper Cholesky factor of W    */
} nt_scaling;

/* NT_scaling.py:7-16 */
static double soc_quad_J(const double* x, int q) { return x[0] * x[0] - dot(x + 1, x + 1, q - 1); }

/* NT_scaling.py:340-405 */
static void soc_NT_scaling(const double* s, const double* z, int q, double* W)
{
    double zb[MAXQ], sb[MAXQ], wb[MAXQ];
    double jz = sqrt(soc_quad_J(z, q)), js = sqrt(soc_quad_J(s, q));
    for (int i = 0; i < q; ++i) { zb[i] = z[i] / jz; sb[i] = s[i] / js; }
    double gamma = sqrt((1.0 + dot(zb, sb, q)) / 2.0);
    wb[0] = (sb[0] + zb[0]) / (2.0 * gamma);
    for (int i = 1; i < q; ++i) wb[i] = (sb[i] - zb[i]) / (2.0 * gamma);
    double b = 1.0 / (wb[0] + 1.0);
    double Jz = soc_quad_J(z, q);
    double eta = (Jz != 0.0) ? pow(soc_quad_J(s, q) / Jz, 0.25) : 1.0;
    for (int j = 0; j < q; ++j) W[j] = eta * wb[j];
    for (int i = 1; i < q; ++i) {
        W[i * q] = eta * wb[i];
        for (int j = 1; j < q; ++j) W[i * q + j] = eta * ((i == j ? 1.0 : 0.0) + b * (wb[i] * wb[j]));
    }
}

/* NT_scaling.py:407-463; returns 0 / DCOL_STATUS_NON_FINITE / DCOL_STATUS_NOT_PD (cho_factor) */
static int calc_NT_scalings(const double* s, const double* z, const cone_idx* K, nt_scaling* W)
{
    for (int i = 0; i < K->n_ort; ++i) W->ort[i] = sqrt(s[i] / z[i]);
    int off = K->n_ort;
    for (int k = 0; k < 2; ++k) {
        int q = k == 0 ? K->q1 : K->q2;
        if (q > 0) soc_NT_scaling(s + off, z + off, q, W->soc[k]);
        off += q;
    }
    for (int k = 0; k < 2; ++k) {
        int q = k == 0 ? K->q1 : K->q2;
        if (q > 0) {
            if (!all_finite(W->soc[k], q * q)) return DCOL_STATUS_NON_FINITE; /* cho_factor check_finite */
            memcpy(W->soc_fact[k], W->soc[k], sizeof(double) * q * q);
            if (chol_upper(W->soc_fact[k], q, q)) return DCOL_STATUS_NOT_PD;
        }
    }
    return 0;
}

/* NT_scaling.py:205-240 */
static void multiply_nt(const nt_scaling* W, const double* g, const cone_idx* K, double* out)
{
    for (int i = 0; i < K->n_ort; ++i) out[i] = g[i] * W->ort[i];
    int off = K->n_ort;
    for (int k = 0; k < 2; ++k) {
        int q = k == 0 ? K->q1 : K->q2;
        for (int i = 0; i < q; ++i) out[off + i] = dot(W->soc[k] + i * q, g + off, q);
        off += q;
    }
}

/* NT_scaling.py:75-126; returns 0 or DCOL_STATUS_NON_FINITE (cho_solve check_finite on the rhs) */
static int solve_nt(const nt_scaling* W, const double* g, const cone_idx* K, double* out)
{
    for (int i = 0; i < K->n_ort; ++i) out[i] = g[i] / W->ort[i];
    int off = K->n_ort;
    for (int k = 0; k < 2; ++k) {
        int q = k == 0 ? K->q1 : K->q2;
        if (q > 0) {
            if (!all_finite(g + off, q)) return DCOL_STATUS_NON_FINITE;
            for (int i = 0; i < q; ++i) out[off + i] = g[off + i];
            cho_solve_upper(W->soc_fact[k], q, q, out + off);
        }
        off += q;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pdip.py:291-332 */
static int initialize(const conic_problem* P, const cone_idx* K, double* x, double* s, double* z)
{
    int n = P->n, m = P->m;
    double F[MAXN][MAXN], rhs[MAXN], y[MAXN];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int r = 0; r < m; ++r) acc += P->G[r][i] * P->G[r][j];
            F[i][j] = acc;
        }
    if (chol_lower(F, n)) return DCOL_STATUS_NOT_PD; /* numpy.linalg.cholesky -> LinAlgError */
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int r = 0; r < m; ++r) acc += P->G[r][i] * P->h[r];
        rhs[i] = acc;
    }
    /* solve_triangular(check_finite=True) */
    for (int i = 0; i < n; ++i) if (!all_finite(F[i], n)) return DCOL_STATUS_NON_FINITE;
    if (!all_finite(rhs, n)) return DCOL_STATUS_NON_FINITE;
    for (int i = 0; i < n; ++i) { /* y = F^-1 rhs */
        double v = rhs[i];
        for (int k = 0; k < i; ++k) v -= F[i][k] * y[k];
        y[i] = v / F[i][i];
    }
    for (int i = n - 1; i >= 0; --i) { /* x_hat = F^-T y */
        double v = y[i];
        for (int k = i + 1; k < n; ++k) v -= F[k][i] * x[k];
        x[i] = v / F[i][i];
    }
    for (int r = 0; r < m; ++r) s[r] = dot(P->G[r], x, n) - P->h[r];
    bring2cone(s, K);
    /* quirk Q1 (pdip.py:326): solve_triangular(F, -c) without lower=True reads only the upper
     * triangle of the lower factor, i.e. its diagonal */
    double yx[MAXN], xd[MAXN];
    for (int i = 0; i < n; ++i) yx[i] = (-P->c[i]) / F[i][i];
    for (int i = n - 1; i >= 0; --i) { /* solve_triangular(F.T, y_x): proper back substitution */
        double v = yx[i];
        for (int k = i + 1; k < n; ++k) v -= F[k][i] * xd[k];
        xd[i] = v / F[i][i];
    }
    for (int r = 0; r < m; ++r) z[r] = dot(P->G[r], xd, n);
    bring2cone(z, K);
    return 0;
}

/* direction solve shared by the affine and corrector steps (pdip.py:427-440 and :452-460) */
static int solve_direction(const conic_problem* P, const cone_idx* K, const nt_scaling* W, const double Gt[MAXM][MAXN],
                           const double* F, const double* bx, const double* rz, const double* lambd_ds,
                           double* dx, double* ds, double* dz)
{
    int n = P->n, m = P->m;
    double t[MAXM], bzt[MAXM], rhs[MAXN];
    multiply_nt(W, lambd_ds, K, t);
    for (int i = 0; i < m; ++i) t[i] = -rz[i] - t[i];
    int st = solve_nt(W, t, K, bzt);
    if (st) return st;
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc += Gt[i][j] * bzt[i];
        rhs[j] = bx[j] + acc;
    }
    if (!all_finite(rhs, n)) return DCOL_STATUS_NON_FINITE; /* cho_solve check_finite */
    cho_solve_upper(F, MAXN, n, rhs);
    for (int j = 0; j < n; ++j) dx[j] = rhs[j];
    for (int i = 0; i < m; ++i) t[i] = dot(Gt[i], dx, n) - bzt[i];
    st = solve_nt(W, t, K, dz);
    if (st) return st;
    multiply_nt(W, dz, K, t);
    for (int i = 0; i < m; ++i) t[i] = lambd_ds[i] - t[i];
    multiply_nt(W, t, K, ds);
    return 0;
}

/* pdip.py:373-470.  iters = index of the pass whose mu test succeeded (number of Newton steps taken). */
static int solve_lp_pdip(const conic_problem* P, double tol, double* x, double* s, double* z, int* iters,
                         double* mu_trace)
{
    cone_idx K = { P->n_ort, P->q1, P->q2, P->m };
    int n = P->n, m = P->m;
    *iters = 0;
    int st = initialize(P, &K, x, s, z);
    if (st) return st;
    double e[MAXM];
    gen_e(&K, e);
    int cone_degree = K.n_ort + (K.q1 > 0) + (K.q2 > 0);
    static const int LOOP_CAP = DCOL_MAX_ITER; /* quirk Q2: max_iter is ignored, range(50) */
    for (int it = 0; it < LOOP_CAP; ++it) {
        nt_scaling W;
        *iters = it;
        st = calc_NT_scalings(s, z, &K, &W);
        if (st) return st;
        double lambd[MAXM], ll[MAXM], rx[MAXN], rz[MAXM], bx[MAXN], neg[MAXM], lambd_ds[MAXM];
        multiply_nt(&W, z, &K, lambd);
        cone_product(lambd, lambd, &K, ll);
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i) acc += P->G[i][j] * z[i];
            rx[j] = acc + P->c[j];
        }
        for (int i = 0; i < m; ++i) rz[i] = s[i] + dot(P->G[i], x, n) - P->h[i];
        double mu = dot(s, z, m) / cone_degree;
        if (mu_trace) mu_trace[it] = mu;
        if (mu < tol) return DCOL_STATUS_OK; /* quirk Q3: the only convergence test */
        for (int j = 0; j < n; ++j) bx[j] = -rx[j];
        for (int i = 0; i < m; ++i) neg[i] = -ll[i];
        inverse_cone_product(lambd, neg, &K, lambd_ds);

        /* G_tilde = W \ G column by column (NT_scaling.py:164-202) */
        double Gt[MAXM][MAXN], col[MAXM], sol[MAXM];
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i < m; ++i) col[i] = P->G[i][j];
            st = solve_nt(&W, col, &K, sol);
            if (st) return st;
            for (int i = 0; i < m; ++i) Gt[i][j] = sol[i];
        }
        /* the first solve_nt of the iteration (b_z_tilde, pdip.py:427) precedes G_tilde in the reference;
         * both raise the same ValueError class, so evaluation order does not change the status */
        double F[MAXN * MAXN];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int r = 0; r < m; ++r) acc += Gt[r][i] * Gt[r][j];
                F[i * MAXN + j] = acc;
            }
        for (int i = 0; i < n; ++i) if (!all_finite(F + i * MAXN, n)) return DCOL_STATUS_NON_FINITE; /* scipy cholesky check_finite */
        if (chol_upper(F, MAXN, n)) return DCOL_STATUS_NOT_PD;

        double dx[MAXN], ds[MAXM], dz[MAXM];
        st = solve_direction(P, &K, &W, Gt, F, bx, rz, lambd_ds, dx, ds, dz);
        if (st) return st;

        /* affine line search and centering, pdip.py:446-448 (quirk Q7: no 0.99 here) */
        double alpha = py_min(linesearch(s, ds, &K), linesearch(z, dz, &K));
        double num = 0.0;
        for (int i = 0; i < m; ++i) num += (s[i] + alpha * ds[i]) * (z[i] + alpha * dz[i]);
        double rho = num / dot(s, z, m);
        double sigma = pow(py_max(0.0, py_min(1.0, rho)), 3.0);

        /* corrector, pdip.py:450-460 */
        double a1[MAXM], a2[MAXM], cp[MAXM], dsv[MAXM];
        st = solve_nt(&W, ds, &K, a1);
        if (st) return st;
        multiply_nt(&W, dz, &K, a2);
        cone_product(a1, a2, &K, cp);
        for (int i = 0; i < m; ++i) dsv[i] = (-ll[i] - cp[i]) + (sigma * mu) * e[i];
        inverse_cone_product(lambd, dsv, &K, lambd_ds);
        st = solve_direction(P, &K, &W, Gt, F, bx, rz, lambd_ds, dx, ds, dz);
        if (st) return st;

        alpha = py_min(1.0, 0.99 * py_min(linesearch(s, ds, &K), linesearch(z, dz, &K)));
        for (int j = 0; j < n; ++j) x[j] += alpha * dx[j];
        for (int i = 0; i < m; ++i) { s[i] += alpha * ds[i]; z[i] += alpha * dz[i]; }
    }
    *iters = LOOP_CAP;
    return DCOL_STATUS_MAX_ITER; /* pdip.py:470, raised even if the 50th step converged */
}

/* ------------------------------------------------------------------------------------------ */
/* gradient                                                                                    */

/* proximity_gradient.py:8-47: z'(G(theta) x - h(theta)) for theta = [r1 p1 r2 p2] */
static double lag_con_part(const dcol_shape* s1, const dcol_shape* s2, const double* A, const double* b,
                           const double theta[12], const double* x, const double* z)
{
    prim_blocks B1, B2;
    conic_problem P;
    problem_matrices(s1, A, b, theta, theta + 3, &B1);
    problem_matrices(s2, A, b, theta + 6, theta + 9, &B2);
    combine_problem_matrices(&B1, &B2, &P);
    double t[MAXM];
    for (int i = 0; i < P.m; ++i) t[i] = dot(P.G[i], x, P.n) - P.h[i];
    return dot(z, t, P.m);
}

/* proximity_gradient.py:50-88 + scipy approx_fprime (2-point, abs_step = sqrt(eps) = 2^-26) */
static void obj_val_grad_fd(const dcol_shape* s1, const dcol_shape* s2, const double* A, const double* b,
                            const double pose1[6], const double pose2[6], const double* x, const double* z,
                            double grad[12])
{
    const double eps = 1.4901161193847656e-08; /* np.sqrt(np.finfo(float).eps) */
    double theta[12];
    memcpy(theta, pose1, 6 * sizeof(double));
    memcpy(theta + 6, pose2, 6 * sizeof(double));
    double f0 = lag_con_part(s1, s2, A, b, theta, x, z);
    for (int i = 0; i < 12; ++i) {
        double h = eps;
        volatile double xp = theta[i] + h;
        double dx = xp - theta[i];
        if (dx == 0.0) { /* _numdiff.py: fall back to a relative step */
            double sgn = theta[i] >= 0.0 ? 1.0 : -1.0;
            h = eps * sgn * fmax(1.0, fabs(theta[i]));
            xp = theta[i] + h;
            dx = xp - theta[i];
        }
        double save = theta[i];
        theta[i] = xp;
        double f1 = lag_con_part(s1, s2, A, b, theta, x, z);
        theta[i] = save;
        grad[i] = (f1 - f0) / dx;
    }
}

/* d dcm / d p_k by the quotient rule on Q = I + (8 S^2 + 4 (1-|p|^2) S) / (1+|p|^2)^2, S = [p x] */
static void dcm_derivative(const double p[3], double dQ[3][3][3])
{
    double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    double D = (1.0 + pp) * (1.0 + pp);
    double S[3][3] = { { 0, -p[2], p[1] }, { p[2], 0, -p[0] }, { -p[1], p[0], 0 } };
    double S2[3][3], N[3][3];
    mat3_mul(S, S, S2);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) N[i][j] = 8.0 * S2[i][j] + 4.0 * (1.0 - pp) * S[i][j];
    for (int k = 0; k < 3; ++k) {
        double Sk[3][3] = { { 0 } }, SkS[3][3], SSk[3][3];
        int a = (k + 1) % 3, c = (k + 2) % 3; /* [e_k x]: (c,a) = +1, (a,c) = -1 */
        Sk[c][a] = 1.0; Sk[a][c] = -1.0;
        mat3_mul(Sk, S, SkS);
        mat3_mul(S, Sk, SSk);
        double dD = 4.0 * (1.0 + pp) * p[k];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double dN = 8.0 * (SkS[i][j] + SSk[i][j]) - 8.0 * p[k] * S[i][j] + 4.0 * (1.0 - pp) * Sk[i][j];
                dQ[k][i][j] = dN / D - N[i][j] * dD / (D * D);
            }
    }
}

/* z_k'(G_k x - h_k) restricted to one primitive's own rows, as a function of the adjusted pose */
static double prim_lagrangian(const dcol_shape* sh, const double* A, const double* b, const double r[3],
                              const double Q[3][3], const double* xk /* x(3), alpha, own extras */,
                              const double* z_ort, const double* z_soc)
{
    prim_blocks B;
    blocks_from_adjusted(sh, A, b, r, Q, &B);
    double acc = 0.0;
    for (int i = 0; i < B.n_ort; ++i) acc += z_ort[i] * (dot(B.G_ort[i], xk, B.v) - B.h_ort[i]);
    for (int i = 0; i < B.n_soc; ++i) acc += z_soc[i] * (dot(B.G_soc[i], xk, B.v) - B.h_soc[i]);
    return acc;
}

/* EXTENSION (not in the reference): the exact derivative of the same frozen-(x,z) Lagrangian that
 * proximity_gradient.py:80-86 differentiates by finite differences.  The blocks are affine in Q'
 * for fixed r' and affine in r' for fixed Q' (h is bilinear), so the derivative is assembled from
 * differences of the block builder evaluated at (dQ', r'), (0, r'), (Q', dr'), (Q', 0) — a different
 * route from the closed forms the CUDA kernel uses, which is what makes it a useful cross-check. */
static void obj_val_grad_exact(const dcol_shape* s1, const dcol_shape* s2, const double* A, const double* b,
                               const double pose1[6], const double pose2[6], const conic_problem* P,
                               const double* x, const double* z, double grad[12])
{
    const dcol_shape* sh[2] = { s1, s2 };
    const double* pose[2] = { pose1, pose2 };
    int no1 = 0;
    {
        prim_blocks B;
        double I3[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } }, zero[3] = { 0, 0, 0 };
        blocks_from_adjusted(s1, A, b, zero, I3, &B);
        no1 = B.n_ort;
    }
    for (int k = 0; k < 2; ++k) {
        const double* r = pose[k];
        const double* p = pose[k] + 3;
        double Q[3][3], Qo[3][3], Qp[3][3], rp[3], dQ[3][3][3], Z3[3][3] = { { 0 } }, zero[3] = { 0, 0, 0 };
        dcm_from_mrp(p, Q);
        dcm_derivative(p, dQ);
        memcpy(Qo, sh[k]->Q_offset, sizeof(Qo));
        mat3_mul(Q, Qo, Qp);
        for (int i = 0; i < 3; ++i)
            rp[i] = r[i] + (Q[i][0] * sh[k]->r_offset[0] + Q[i][1] * sh[k]->r_offset[1] + Q[i][2] * sh[k]->r_offset[2]);
        /* this primitive's slice of x and z */
        double xk[MAXN] = { x[0], x[1], x[2], x[3] };
        int ne = (sh[k]->type == DCOL_CAPSULE || sh[k]->type == DCOL_CYLINDER) ? 1 : (sh[k]->type == DCOL_POLYGON ? 2 : 0);
        {
            int ne1 = (s1->type == DCOL_CAPSULE || s1->type == DCOL_CYLINDER) ? 1 : (s1->type == DCOL_POLYGON ? 2 : 0);
            int off = (k == 1) ? ne1 : 0; /* the second primitive's extras follow the first's */
            for (int j = 0; j < ne; ++j) xk[4 + j] = x[4 + off + j];
        }
        const double* z_ort = z + (k == 0 ? 0 : no1);
        const double* z_soc = z + P->n_ort + (k == 0 ? 0 : P->q1);
        double base_r0 = prim_lagrangian(sh[k], A, b, zero, Qp, xk, z_ort, z_soc);
        for (int j = 0; j < 3; ++j) { /* d/dr_j: L is affine in r' */
            double ej[3] = { 0, 0, 0 };
            ej[j] = 1.0;
            grad[6 * k + j] = prim_lagrangian(sh[k], A, b, ej, Qp, xk, z_ort, z_soc) - base_r0;
        }
        double base_Q0 = prim_lagrangian(sh[k], A, b, rp, Z3, xk, z_ort, z_soc);
        for (int j = 0; j < 3; ++j) { /* d/dp_j through Q' = Q Q_offset and r' = r + Q r_offset */
            double dQp[3][3], drp[3];
            mat3_mul(dQ[j], Qo, dQp);
            for (int i = 0; i < 3; ++i)
                drp[i] = dQ[j][i][0] * sh[k]->r_offset[0] + dQ[j][i][1] * sh[k]->r_offset[1] + dQ[j][i][2] * sh[k]->r_offset[2];
            double viaQ = prim_lagrangian(sh[k], A, b, rp, dQp, xk, z_ort, z_soc) - base_Q0;
            double viaR = prim_lagrangian(sh[k], A, b, drp, Qp, xk, z_ort, z_soc) - base_r0;
            grad[6 * k + 3 + j] = viaQ + viaR;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* exported entry points (ctypes)                                                              */

enum { DCOL_ORACLE_GRAD_NONE = 0, DCOL_ORACLE_GRAD_FD = 1, DCOL_ORACLE_GRAD_EXACT = 2 };

/* One pair.  x[8], s[72], z[72], grad[12], mu_trace[51] (may be NULL).  Returns the status word. */
int dcol_oracle_pair(const dcol_shape* shapes, const double* A, const double* b, int32_t i1, int32_t i2,
                     const double* pose1, const double* pose2, double tol, int grad_mode, double* alpha,
                     double* x, double* s, double* z, int32_t* n, int32_t* m, int32_t* iters, double* grad,
                     double* mu_trace)
{
    prim_blocks B1, B2;
    conic_problem P;
    double xs[MAXN], ss[MAXM], zs[MAXM];
    const dcol_shape* s1 = shapes + i1;
    const dcol_shape* s2 = shapes + i2;
    *alpha = NAN; *iters = 0; *n = 0; *m = 0;
    if (mu_trace) for (int i = 0; i <= DCOL_MAX_ITER; ++i) mu_trace[i] = NAN;
    if (grad) for (int i = 0; i < 12; ++i) grad[i] = NAN;
    problem_matrices(s1, A, b, pose1, pose1 + 3, &B1);
    problem_matrices(s2, A, b, pose2, pose2 + 3, &B2);
    int st = combine_problem_matrices(&B1, &B2, &P);
    if (st) return st;
    *n = P.n; *m = P.m;
    int it = 0;
    st = solve_lp_pdip(&P, tol, xs, ss, zs, &it, mu_trace);
    *iters = it;
    if (st) return st;
    *alpha = xs[3]; /* proximity.py:51 */
    if (x) memcpy(x, xs, sizeof(double) * P.n);
    if (s) memcpy(s, ss, sizeof(double) * P.m);
    if (z) memcpy(z, zs, sizeof(double) * P.m);
    if (grad && grad_mode == DCOL_ORACLE_GRAD_FD) obj_val_grad_fd(s1, s2, A, b, pose1, pose2, xs, zs, grad);
    if (grad && grad_mode == DCOL_ORACLE_GRAD_EXACT) obj_val_grad_exact(s1, s2, A, b, pose1, pose2, &P, xs, zs, grad);
    return DCOL_STATUS_OK;
}

/* A batch, pairs spread over `threads` POSIX threads (interleaved chunks of 64 pairs).
 * pose1/pose2 [B][6], contact [B][3] (may be NULL), grad [B][12] (may be NULL). */
typedef struct {
    const dcol_shape* shapes; const double *A, *b; const int32_t *idx1, *idx2; const double *pose1, *pose2;
    int64_t B; double tol; int grad_mode, tid, nthreads;
    double *alpha, *contact, *grad; int32_t *iters, *status;
} batch_job;

static void* batch_worker(void* arg)
{
    const batch_job* J = (const batch_job*)arg;
    const int64_t chunk = 64;
    for (int64_t c0 = (int64_t)J->tid * chunk; c0 < J->B; c0 += (int64_t)J->nthreads * chunk) {
        int64_t c1 = c0 + chunk < J->B ? c0 + chunk : J->B;
        for (int64_t k = c0; k < c1; ++k) {
            double x[MAXN], g[12], a;
            int32_t n, m, it;
            int st = dcol_oracle_pair(J->shapes, J->A, J->b, J->idx1[k], J->idx2[k], J->pose1 + 6 * k,
                                      J->pose2 + 6 * k, J->tol, J->grad ? J->grad_mode : DCOL_ORACLE_GRAD_NONE,
                                      &a, x, 0, 0, &n, &m, &it, g, 0);
            J->alpha[k] = a; J->iters[k] = it; J->status[k] = st;
            if (J->contact) for (int j = 0; j < 3; ++j) J->contact[3 * k + j] = st ? NAN : x[j];
            if (J->grad) for (int j = 0; j < 12; ++j) J->grad[12 * k + j] = g[j];
        }
    }
    return 0;
}

int dcol_oracle_batch(const dcol_shape* shapes, const double* A, const double* b, const int32_t* idx1,
                      const int32_t* idx2, const double* pose1, const double* pose2, int64_t B, double tol,
                      int grad_mode, int threads, double* alpha, double* contact, double* grad, int32_t* iters,
                      int32_t* status)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    batch_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        batch_job j = { shapes, A, b, idx1, idx2, pose1, pose2, B, tol, grad_mode, t, threads,
                        alpha, contact, grad, iters, status };
        jobs[t] = j;
    }
    if (threads == 1) { batch_worker(&jobs[0]); return 0; }
    for (int t = 0; t < threads; ++t) pthread_create(&tid[t], 0, batch_worker, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], 0);
    return 0;
}

/* assembled problem of one pair, for tests of the assembly alone: G row-major [m][8] */
int dcol_oracle_assemble(const dcol_shape* shapes, const double* A, const double* b, int32_t i1, int32_t i2,
                         const double* pose1, const double* pose2, double* c, double* G, double* h, int32_t* dims)
{
    prim_blocks B1, B2;
    conic_problem P;
    problem_matrices(shapes + i1, A, b, pose1, pose1 + 3, &B1);
    problem_matrices(shapes + i2, A, b, pose2, pose2 + 3, &B2);
    int st = combine_problem_matrices(&B1, &B2, &P);
    if (st) return st;
    memcpy(c, P.c, sizeof(P.c));
    for (int i = 0; i < P.m; ++i) { memcpy(G + i * MAXN, P.G[i], sizeof(double) * MAXN); h[i] = P.h[i]; }
    dims[0] = P.n; dims[1] = P.m; dims[2] = P.n_ort; dims[3] = P.q1; dims[4] = P.q2;
    return 0;
}

void dcol_oracle_dcm(const double* p, double* Q, double* dQ)
{
    double q[3][3], d[3][3][3];
    dcm_from_mrp(p, q);
    dcm_derivative(p, d);
    memcpy(Q, q, sizeof(q));
    if (dQ) memcpy(dQ, d, sizeof(d));
}
