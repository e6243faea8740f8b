/*
 * dcol_solver.cuh — one (primitive, primitive) proximity solve + pose gradient, owned by ONE thread.
 *
 * The whole fixed-cap primal-dual interior-point loop of the reference
 *     proximity/pdip.py:373-470 (solve_lp_pdip), :291-332 (initialize), :237-287 (bring2cone)
 *     proximity/NT/NT_scaling.py:340-463 (Nesterov-Todd scaling)
 *     primitives/problem_matrices.py:4-364, primitives/combine_problem_matrices.py:3-70 (assembly)
 *     proximity/proximity_gradient.py:50-88 (gradient of the frozen-(x,z) Lagrangian)
 * runs in registers, in FP64, specialised at compile time on the pair's kinds and face counts.
 *
 * It is the same algorithm, iterate for iterate (same initial point incl. the diag-only
 * triangular solve, same mu-only stopping test, same Mehrotra predictor / corrector, same
 * un-damped affine step and 0.99-damped final step, same 50-iteration cap) but NOT the same
 * arithmetic layout.  What is different, and why it is exact-arithmetic equivalent:
 *
 *  1. Body-frame constraint rows.  Every primitive's block of G x - h equals a CONSTANT matrix
 *     applied to the local vector (Q'^T (x - r'), alpha, extras): the pose enters only through
 *     that affine map.  So G is never stored: its coefficients are warp-uniform shape constants
 *     (kernel parameters -> constant bank operands), a thread keeps only Q' and r' of each
 *     primitive, and Gram matrices / adjoint products are accumulated in the local frame and
 *     rotated once.  Second-order-cone rows of ball-like primitives are carried in body-frame
 *     components (a rotation of the cone's vector part commutes with every cone operation).
 *  2. Scaled-space Newton step on the second-order cones.  With lambda = W z = W^-1 s, the
 *     directions are carried as ds~ = W^-1 ds and dz~ = W dz; both line searches then run against
 *     lambda (W is an automorphism of the cone) and the centring ratio uses <ds, dz> = <ds~, dz~>.
 *     On the ORTHANT the scaled directions are never formed: the relative steps ds_i / s_i and
 *     dz_i / z_i are the line-search measures themselves, every product with w_i or lambda_i
 *     collapses into 1 / s_i and z_i / s_i, and one rcp(s_i z_i) per row per iteration replaces
 *     the reference's ~8 divisions / square roots (pass_a .. pass_d).
 *  3. W^-1 of a second-order cone in closed form (1/eta) J Wbar J, no Cholesky of W; rz of a cone
 *     block stays unscaled and the slack is updated with the primal equation's own ds; the two
 *     affine line searches share one norm (soc_ls_affine).
 *  4. The Newton systems are factored as L D L^T (ldlt): same pivots, same tests, shorter
 *     dependent chains.  The initial point's normal equations come from a shape constant
 *     (ShapeConst::G0, init_accumulate).
 *
 * Rounding therefore differs from NumPy at the 1e-16 level per operation; SURVEY.md section 0
 * measured that such noise moves alpha by <= 2e-14 and never flips an iteration count.
 *
 * Compiles as CUDA device code and as plain C++ (tests build a host twin of this header to
 * debug the algorithm against the oracle without a GPU; the product never runs it on the CPU).
 */
#ifndef DCOL_SOLVER_CUH_
#define DCOL_SOLVER_CUH_

#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/dcol.h"

#if defined(__CUDACC__)
#define DCOL_HD __host__ __device__ __forceinline__
#define DCOL_UNROLL _Pragma("unroll")
#define DCOL_UNROLL_ROWS _Pragma("unroll (P::dyn ? 1 : 64)")
#else
#define DCOL_HD inline __attribute__((always_inline))
#define DCOL_UNROLL
#define DCOL_UNROLL_ROWS
#endif

namespace dcol {

/* ---------------------------------------------------------------------------------------- */
/* scalar helpers                                                                            */

/* 1/sqrt(x) and 1/x for the solver's scalars (products s_i z_i, Cholesky pivots, cone determinants).
 * On the device: the MUFU seed (rsqrt/rcp.approx.f64, ~2^-22 relative) and ONE third-order Newton step,
 *   e = 1 - x y^2,  y += y e (1/2 + 3/8 e)        |      e = 1 - x y,  y += y (e + e^2),
 * which brings the error to ~e^3 < 2^-60, i.e. rounding level, in 5 (resp. 3) FP64 instructions with no
 * branch; the CUDA library routines spend ~4x that on denormal / special-case paths this solver never
 * needs: a zero, negative, infinite or NaN argument still yields a non-finite result, which is all the
 * status logic relies on. */
DCOL_HD double rsqrt_(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
#else
    return 1.0 / sqrt(x);
#endif
}
DCOL_HD double rcp_(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
#else
    return 1.0 / x;
#endif
}
/* sqrt(x) for the cone norms of the line searches: x * rsqrt_(x), one rounding worse than the library's correctly
 * rounded routine at a third of its instructions; 0 -> 0, negative or NaN -> NaN */
DCOL_HD double sqrt_(double x)
{
#if defined(__CUDA_ARCH__)
    const double r = x * rsqrt_(x);
    return (x == 0.0) ? 0.0 : r;
#else
    return sqrt(x);
#endif
}
DCOL_HD double max_(double a, double b) { return (b > a) ? b : a; } /* Python max(a, b): keeps a on NaN */
DCOL_HD double min_(double a, double b) { return (b < a) ? b : a; }
/* 0 if x is finite, NaN otherwise (x * 0 is NaN for inf and NaN) */
DCOL_HD double nonfinite_probe(double x) { return x * 0.0; }

/* primitives/problem_matrices.py:213-251: direction cosine matrix of a modified Rodrigues vector,
 * Q = I + (8 S^2 + 4 (1 - |p|^2) S) / (1 + |p|^2)^2 with S = [p x] */
DCOL_HD void dcm_from_mrp(const double p[3], double Q[3][3])
{
    const double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const double t = 1.0 + pp;
    const double iden = rcp_(t * t);
    const double k8 = 8.0 * iden, k4 = 4.0 * (1.0 - pp) * iden;
    Q[0][0] = 1.0 - k8 * (p[1] * p[1] + p[2] * p[2]);
    Q[1][1] = 1.0 - k8 * (p[0] * p[0] + p[2] * p[2]);
    Q[2][2] = 1.0 - k8 * (p[0] * p[0] + p[1] * p[1]);
    Q[0][1] = k8 * p[0] * p[1] - k4 * p[2];
    Q[1][0] = k8 * p[0] * p[1] + k4 * p[2];
    Q[0][2] = k8 * p[0] * p[2] + k4 * p[1];
    Q[2][0] = k8 * p[0] * p[2] - k4 * p[1];
    Q[1][2] = k8 * p[1] * p[2] - k4 * p[0];
    Q[2][1] = k8 * p[1] * p[2] + k4 * p[0];
}

/* <M, dQ/dp_k> for k = 0..2, with dQ_k = (dN_k - (Q - I) dD_k) / D, N = 8 S^2 + 4 (1-pp) S,
 * D = (1+pp)^2.  (Analytic replacement for the reference's finite differences through
 * dcm_from_mrp, proximity_gradient.py:80-86.)
 * dN_k = 8 (S_k S + S S_k) + 4 (1-pp) S_k - 8 p_k S with S_k = [e_k x], and [a x][b x] = b a^T - (a.b) I gives
 * S_k S + S S_k = p e_k^T + e_k p^T - 2 p_k I, so the contraction with M needs only M p, M^T p, tr M, the
 * antisymmetric part m_a of M (<M, S> = p.m_a, <M, S_k> = (m_a)_k) and <M, Q - I>: ~50 operations instead of the
 * 27 entry-by-entry products. */
DCOL_HD void dcm_derivative_contract(const double p[3], const double Q[3][3], const double M[3][3], double out[3])
{
    const double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const double t = 1.0 + pp, iD = rcp_(t * t);
    const double tr = M[0][0] + M[1][1] + M[2][2];
    const double ma[3] = { M[2][1] - M[1][2], M[0][2] - M[2][0], M[1][0] - M[0][1] };
    const double ms = p[0] * ma[0] + p[1] * ma[1] + p[2] * ma[2]; /* <M, S> */
    double c0 = 0.0;                                              /* <M, Q - I> */
    DCOL_UNROLL
    for (int i = 0; i < 3; ++i) {
        DCOL_UNROLL
        for (int j = 0; j < 3; ++j) c0 += M[i][j] * (Q[i][j] - (i == j ? 1.0 : 0.0));
    }
    const double k4 = 4.0 * (1.0 - pp);
    const double kp = 16.0 * tr + 8.0 * ms + 4.0 * t * c0; /* everything that multiplies -p_k */
    DCOL_UNROLL
    for (int k = 0; k < 3; ++k) {
        const double mp = (M[k][0] + M[0][k]) * p[0] + (M[k][1] + M[1][k]) * p[1] + (M[k][2] + M[2][k]) * p[2];
        out[k] = (8.0 * mp + k4 * ma[k] - kp * p[k]) * iD;
    }
}

/* ---------------------------------------------------------------------------------------- */
/* shape constants: warp-uniform, passed by value as kernel parameters                       */

template <int FMAX>
struct ShapeConst {
    double R, L, H, tanb;   /* tanb = tan(beta)                       */
    double ia[3];           /* ellipsoid: inverse semi-axes           */
    double r_off[3];        /* body-frame origin offset               */
    double Q_off[3][3];     /* body-frame rotation offset             */
    int32_t nf;             /* faces actually present                 */
    int32_t no_offset;      /* r_off = 0 and Q_off = I (every shape of the reference's scenes): the products with them are skipped */
    /* G^T G of this shape's rows in its own local variables (upper triangle, row-major packed, NL <= 6): a constant of
     * the shape, which is all the initial point's normal equations need (Solver::init_accumulate); filled on the host
     * by fill_const_prim (dcol_classes.cuh) */
    double G0[21];
    double A[FMAX > 0 ? FMAX : 1][3];
    double b[FMAX > 0 ? FMAX : 1];
};

/* ---------------------------------------------------------------------------------------- */
/* primitive kinds.  Local variables of a primitive: xh = (y0, y1, y2, alpha, e0, e1) with     */
/* y = Q'^T (x - r') and e its own extra decision variables.  In these variables every row of  */
/* the primitive's block of G x - h has constant coefficients:                                 */
/*   polytope  face i   :  A_i . y - b_i alpha                 problem_matrices.py:181-209     */
/*   box       face i   :  +-y_(i mod 3) - b_i alpha           (polytope with A = [I; -I])     */
/*   cone      orthant  :  y0 - (H/4) alpha                    problem_matrices.py:125-148     */
/*             soc(3)   :  (-tanb y0 - (3/4) H tanb alpha, -y1, -y2)                           */
/*   capsule   orthant  :  -(L/2) alpha +- e0                  problem_matrices.py:4-44        */
/*             soc(4)   :  (-R alpha, -y0 + e0, -y1, -y2)                                      */
/*   cylinder  capsule + :  -+ y0 - (L/2) alpha                problem_matrices.py:47-87       */
/*   sphere    soc(4)   :  (-R alpha, -y)   (y in world axes)  problem_matrices.py:151-178     */
/*   polygon   face i   :  -b_i alpha + A_i0 e0 + A_i1 e1      problem_matrices.py:90-120      */
/*   ellipsoid soc(4)   :  (-alpha, -y0/a, -y1/b, -y2/c)       extension, Report.pdf eq. 27    */
/*             soc(4)   :  (-R alpha, -y0 + e0, -y1 + e1, -y2)                                 */

/* Internal kind: a polytope whose faces are exactly A = [I; -I] (create_rect_prism,
 * misc_primitive_constructor.py:91-142 — every box of the reference's scenes).  Same rows as the
 * polytope, but each has two non-zeros (+-y_k, -b_i alpha), known at compile time. */
constexpr int KIND_BOX = 100;

/* IDENT: the solver's working frame IS this primitive's body frame (Q' = I, r' = 0 are not stored and
 * every rotation / translation of this primitive's local vectors disappears at compile time). */
template <int KIND, int FC, bool IDENT = false>
struct Prim {
    static constexpr int kind = KIND;
    static constexpr int fc = FC;
    static constexpr bool ident = IDENT;
    static constexpr bool has_faces = (KIND == DCOL_POLYTOPE || KIND == DCOL_POLYGON || KIND == KIND_BOX);
    static constexpr bool dyn = has_faces && FC == 0;                       /* runtime face count      */
    static constexpr int FMAX = has_faces ? (FC ? FC : DCOL_MAX_FACES) : 0; /* sized for               */
    static constexpr int NO = has_faces ? FMAX : (KIND == DCOL_CONE ? 1 : KIND == DCOL_CAPSULE ? 2 : KIND == DCOL_CYLINDER ? 4 : 0);
    static constexpr int NOA = NO > 0 ? NO : 1;                             /* array extent            */
    static constexpr int Q = (KIND == DCOL_POLYTOPE || KIND == KIND_BOX) ? 0 : (KIND == DCOL_CONE ? 3 : 4);
    static constexpr int QA = Q > 0 ? Q : 1;
    static constexpr int NE = (KIND == DCOL_CAPSULE || KIND == DCOL_CYLINDER) ? 1 : (KIND == DCOL_POLYGON ? 2 : 0);
    static constexpr int NL = 4 + NE;
    static constexpr bool rot = KIND != DCOL_SPHERE;  /* the sphere's rows are written in world axes */
    static constexpr bool frot = rot && !IDENT;       /* local vectors are rotated by Qp                */
    static constexpr bool ball = Q == 4 && KIND != DCOL_ELLIPSOID; /* soc rows (-R alpha, -y + E e), world-frame duals */
    typedef ShapeConst<FMAX> Const;
    typedef Prim P;
    typedef Prim<KIND, FC, true> Ident; /* the same primitive as the solver's frame */

    double Qp[3][3]; /* Q' = Q(p) Q_offset                  */
    double rp[3];    /* r' = r + Q(p) r_offset              */

    DCOL_HD static int n_ort(const Const& c) { return dyn ? c.nf : NO; }

    /* structural non-zero of orthant row i at local column j */
    DCOL_HD static bool nz(int i, int j)
    {
        switch (KIND) {
        case DCOL_POLYTOPE: return j < 4;
        case KIND_BOX: return j == 3 || j == i % 3;
        case DCOL_POLYGON: return j >= 3;
        case DCOL_CONE: return j == 0 || j == 3;
        case DCOL_CAPSULE: return j == 3 || j == 4;
        case DCOL_CYLINDER: return i < 2 ? (j == 3 || j == 4) : (j == 0 || j == 3);
        default: return false;
        }
    }
    /* coefficient of orthant row i at local column j (where nz) */
    DCOL_HD static double g(const Const& c, int i, int j)
    {
        switch (KIND) {
        case DCOL_POLYTOPE: return j < 3 ? c.A[i][j] : -c.b[i];
        case KIND_BOX: return j == 3 ? -c.b[i] : (i < 3 ? 1.0 : -1.0);
        case DCOL_POLYGON: return j == 3 ? -c.b[i] : c.A[i][j - 4];
        case DCOL_CONE: return j == 0 ? 1.0 : -0.25 * c.H;
        case DCOL_CAPSULE: return j == 3 ? -0.5 * c.L : (i == 0 ? 1.0 : -1.0);
        case DCOL_CYLINDER:
            if (i < 2) return j == 3 ? -0.5 * c.L : (i == 0 ? 1.0 : -1.0);
            return j == 3 ? -0.5 * c.L : (i == 2 ? -1.0 : 1.0);
        default: return 0.0;
        }
    }
    /* soc block: local column j feeds exactly one cone row, soc_row(j), with coefficient soc_c(j) */
    DCOL_HD static int soc_row(int j)
    {
        if (KIND == DCOL_CONE) return j < 3 ? j : 0;
        return j < 3 ? 1 + j : (j == 3 ? 0 : j - 3); /* ball: y_j -> 1+j, alpha -> 0, e0 -> 1, e1 -> 2 */
    }
    DCOL_HD static double soc_c(const Const& c, int j)
    {
        if (KIND == DCOL_CONE) return j == 0 ? -c.tanb : (j == 3 ? -0.75 * c.H * c.tanb : -1.0);
        if (KIND == DCOL_ELLIPSOID) return j < 3 ? -c.ia[j] : -1.0;
        return j < 3 ? -1.0 : (j == 3 ? -c.R : 1.0);
    }

    /* pose -> (Q', r')   problem_matrices.py:272-364 */
    DCOL_HD void set_pose(const Const& c, const double r[3], const double Qm[3][3])
    {
        if (c.no_offset) { /* warp-uniform */
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                rp[i] = r[i];
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) Qp[i][j] = rot ? Qm[i][j] : (i == j ? 1.0 : 0.0);
            }
            return;
        }
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) {
            rp[i] = r[i] + (Qm[i][0] * c.r_off[0] + Qm[i][1] * c.r_off[1] + Qm[i][2] * c.r_off[2]);
            DCOL_UNROLL
            for (int j = 0; j < 3; ++j)
                Qp[i][j] = rot ? (Qm[i][0] * c.Q_off[0][j] + Qm[i][1] * c.Q_off[1][j] + Qm[i][2] * c.Q_off[2][j])
                               : (i == j ? 1.0 : 0.0);
        }
    }

    /* this primitive's pose expressed in the frame (Qf, rf) of the other one:  Q' <- Qf^T Q',  r' <- Qf^T (r' - rf) */
    DCOL_HD void make_relative(const double (&Qf)[3][3], const double (&rf)[3])
    {
        double d[3], Qn[3][3];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) d[i] = rp[i] - rf[i];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) {
            rp[i] = Qf[0][i] * d[0] + Qf[1][i] * d[1] + Qf[2][i] * d[2];
            DCOL_UNROLL
            for (int j = 0; j < 3; ++j) Qn[i][j] = Qf[0][i] * Qp[0][j] + Qf[1][i] * Qp[1][j] + Qf[2][i] * Qp[2][j];
        }
        if (rot) {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) Qp[i][j] = Qn[i][j];
            }
        }
    }

    /* world vector v[N] -> local xh[NL];  AFFINE subtracts r' (points), otherwise directions */
    template <int N, bool AFFINE>
    DCOL_HD void to_local(const double (&v)[N], int col_e, double (&xh)[NL]) const
    {
        double d[3];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) d[i] = (AFFINE && !IDENT) ? v[i] - rp[i] : v[i];
        DCOL_UNROLL
        for (int j = 0; j < 3; ++j) xh[j] = frot ? (Qp[0][j] * d[0] + Qp[1][j] * d[1] + Qp[2][j] * d[2]) : d[j];
        xh[3] = v[3];
        DCOL_UNROLL
        for (int j = 0; j < NE; ++j) xh[4 + j] = v[col_e + j];
    }
    /* out[N] += T^T acc  (adjoint of the linear part of to_local) */
    template <int N>
    DCOL_HD void from_local_add(const double (&acc)[NL], int col_e, double (&out)[N]) const
    {
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i)
            out[i] += frot ? (Qp[i][0] * acc[0] + Qp[i][1] * acc[1] + Qp[i][2] * acc[2]) : acc[i];
        out[3] += acc[3];
        DCOL_UNROLL
        for (int j = 0; j < NE; ++j) out[col_e + j] += acc[4 + j];
    }
    /* M[N][N] (upper triangle) += T^T Gl T for a symmetric local matrix Gl (upper triangle valid) */
    template <int N>
    DCOL_HD void gram_from_local(const double (&Gl)[NL][NL], int col_e, double (&M)[N][N]) const
    {
        if (frot) {
            double T[3][3]; /* T = Q' Gl_yy */
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) {
                    double t = 0.0;
                    DCOL_UNROLL
                    for (int k = 0; k < 3; ++k) t += Qp[i][k] * (k <= j ? Gl[k][j] : Gl[j][k]);
                    T[i][j] = t;
                }
            }
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = i; j < 3; ++j) M[i][j] += T[i][0] * Qp[j][0] + T[i][1] * Qp[j][1] + T[i][2] * Qp[j][2];
            }
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                M[i][3] += Qp[i][0] * Gl[0][3] + Qp[i][1] * Gl[1][3] + Qp[i][2] * Gl[2][3];
                DCOL_UNROLL
                for (int j = 0; j < NE; ++j)
                    M[i][col_e + j] += Qp[i][0] * Gl[0][4 + j] + Qp[i][1] * Gl[1][4 + j] + Qp[i][2] * Gl[2][4 + j];
            }
        } else {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = i; j < 3; ++j) M[i][j] += Gl[i][j];
                M[i][3] += Gl[i][3];
                DCOL_UNROLL
                for (int j = 0; j < NE; ++j) M[i][col_e + j] += Gl[i][4 + j];
            }
        }
        M[3][3] += Gl[3][3];
        DCOL_UNROLL
        for (int j = 0; j < NE; ++j) {
            M[3][col_e + j] += Gl[3][4 + j];
            DCOL_UNROLL
            for (int k = j; k < NE; ++k) M[col_e + j][col_e + k] += Gl[4 + j][4 + k];
        }
    }

    /* one orthant row: value on a local vector, rank-one Gram update, axpy into a local adjoint accumulator
     * (the passes of the solver fuse these per row so that per-row temporaries die immediately) */
    DCOL_HD static double row_dot(const Const& c, int i, const double (&xh)[NL], double t = 0.0)
    {
        DCOL_UNROLL
        for (int j = 0; j < NL; ++j)
            if (nz(i, j)) t += g(c, i, j) * xh[j];
        return t;
    }
    DCOL_HD static void row_gram_add(const Const& c, int i, double w, double (&Gl)[NL][NL])
    {
        DCOL_UNROLL
        for (int j = 0; j < NL; ++j) {
            if (!nz(i, j)) continue;
            const double wg = w * g(c, i, j);
            DCOL_UNROLL
            for (int k = j; k < NL; ++k)
                if (nz(i, k)) Gl[j][k] += wg * g(c, i, k);
        }
    }
    /* Gl += w row row^T and acc -= (w r) row in one sweep: the products w g_j are shared */
    DCOL_HD static void row_gram_axpy(const Const& c, int i, double w, double r, double (&Gl)[NL][NL], double (&acc)[NL])
    {
        DCOL_UNROLL
        for (int j = 0; j < NL; ++j) {
            if (!nz(i, j)) continue;
            const double wg = w * g(c, i, j);
            acc[j] = fma(-wg, r, acc[j]);
            DCOL_UNROLL
            for (int k = j; k < NL; ++k)
                if (nz(i, k)) Gl[j][k] += wg * g(c, i, k);
        }
    }
    DCOL_HD static void row_axpy(const Const& c, int i, double cf, double (&acc)[NL])
    {
        DCOL_UNROLL
        for (int j = 0; j < NL; ++j)
            if (nz(i, j)) acc[j] += g(c, i, j) * cf;
    }

    /* orthant rows applied to a local vector */
    DCOL_HD static void ort_apply(const Const& c, const double (&xh)[NL], double (&out)[NOA])
    {
        const int no = n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < NO; ++i) {
            if (dyn && i >= no) break;
            double t = 0.0;
            DCOL_UNROLL
            for (int j = 0; j < NL; ++j)
                if (nz(i, j)) t += g(c, i, j) * xh[j];
            out[i] = t;
        }
    }
    /* acc[NL] += sum_i cf[i] * row_i */
    DCOL_HD static void ort_apply_t(const Const& c, const double (&cf)[NOA], double (&acc)[NL])
    {
        const int no = n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < NO; ++i) {
            if (dyn && i >= no) break;
            DCOL_UNROLL
            for (int j = 0; j < NL; ++j)
                if (nz(i, j)) acc[j] += g(c, i, j) * cf[i];
        }
    }
    /* Gl (upper) += sum_i w[i] row_i row_i^T */
    DCOL_HD static void ort_gram(const Const& c, const double (&w)[NOA], double (&Gl)[NL][NL])
    {
        const int no = n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < NO; ++i) {
            if (dyn && i >= no) break;
            DCOL_UNROLL
            for (int j = 0; j < NL; ++j) {
                if (!nz(i, j)) continue;
                const double wg = w[i] * g(c, i, j);
                DCOL_UNROLL
                for (int k = j; k < NL; ++k)
                    if (nz(i, k)) Gl[j][k] += wg * g(c, i, k);
            }
        }
    }
    /* soc rows applied to a local vector */
    DCOL_HD static void soc_apply(const Const& c, const double (&xh)[NL], double (&out)[QA])
    {
        DCOL_UNROLL
        for (int r = 0; r < Q; ++r) {
            double t = 0.0;
            DCOL_UNROLL
            for (int j = 0; j < NL; ++j)
                if (soc_row(j) == r) t += soc_c(c, j) * xh[j];
            out[r] = t;
        }
    }
    DCOL_HD static void soc_apply_t(const Const& c, const double (&v)[QA], double (&acc)[NL])
    {
        if constexpr (Q > 0) {
            DCOL_UNROLL
            for (int j = 0; j < NL; ++j) acc[j] += soc_c(c, j) * v[soc_row(j)];
        }
    }
    /* Gl (upper) += G_soc^T W2 G_soc for a symmetric Q x Q matrix W2 (upper triangle valid) */
    DCOL_HD static void soc_gram(const Const& c, const double (&W2)[QA][QA], double (&Gl)[NL][NL])
    {
        if (Q == 0) return;
        DCOL_UNROLL
        for (int j = 0; j < NL; ++j) {
            DCOL_UNROLL
            for (int k = j; k < NL; ++k) {
                const int rj = soc_row(j), rk = soc_row(k);
                const double w = rj <= rk ? W2[rj][rk] : W2[rk][rj];
                Gl[j][k] += (soc_c(c, j) * soc_c(c, k)) * w;
            }
        }
    }
};

/* ---------------------------------------------------------------------------------------- */
/* per-primitive block of the iterate and of the per-iteration temporaries                   */

template <class P>
struct Block {
    /* iterate */
    double so[P::NOA], zo[P::NOA]; /* orthant slack / dual                                   */
    double sq[P::QA], zq[P::QA];   /* second-order-cone slack / dual (body-frame components) */
    /* orthant temporaries (relative steps: see pass_a) */
    double rinv[P::NOA]; /* 1 / (s_i z_i) = 1 / lambda_i^2                                  */
    double ta[P::NOA];   /* rz_i = s_i + (G x - h)_i, later ds_i / s_i                      */
    double tb[P::NOA];   /* (ds_i / s_i)(dz_i / z_i) of the affine step, later dz_i / z_i   */
    /* soc temporaries */
    double lam[P::QA]; /* lambda = W z                                                      */
    double tq[P::QA];  /* rz = s + (G x - h), later the corrector's ds (both unscaled)       */
    double kq[P::QA];  /* k = lambda^-1 o (ds~ o dz~), later the corrector's dz~            */
    double wh[P::QA]; /* J wbar = (wbar_0, -wbar_v): W^-1 = (1/eta) Wbar(wh), W = eta Wbar(J wh) */
    double eta, ieta, bw;
    double ls_isn, ls_inu, ls_c0; /* shared line-search terms of lambda, pdip.py:39-47      */
    double irho, il0;             /* 1 / J(lambda), 1 / lambda_0, pdip.py:108-118           */
};

template <int Q>
DCOL_HD double socJ(const double (&u)[Q > 0 ? Q : 1])
{
    double t = u[0] * u[0];
    DCOL_UNROLL
    for (int i = 1; i < Q; ++i) t -= u[i] * u[i];
    return t;
}

/* out = Wbar(w) v = (w.v, v_v + (v_0 + (w_v.v_v) b) w_v), b = 1 / (1 + w_0)   NT_scaling.py:380-392 */
template <int Q>
DCOL_HD void wbar_apply(const double (&w)[Q > 0 ? Q : 1], double b, double sign, const double (&v)[Q > 0 ? Q : 1],
                        double scale, double (&out)[Q > 0 ? Q : 1])
{
    /* sign = +1 uses w as given, -1 flips its vector part */
    double d = 0.0;
    DCOL_UNROLL
    for (int i = 1; i < Q; ++i) d += w[i] * v[i];
    d *= sign;
    out[0] = scale * (w[0] * v[0] + d);
    const double f = sign * (v[0] + d * b);
    DCOL_UNROLL
    for (int i = 1; i < Q; ++i) out[i] = scale * (v[i] + f * w[i]);
}

/* pdip.py:237-287 over both primitives' blocks.  Works on (so, sq) or (zo, zq) selected by SEL. */
template <class P, int SEL>
DCOL_HD void b2c_scan(const typename P::Const& c, const Block<P>& B, bool& any_nonpos, double& mn)
{
    const double* o = SEL == 0 ? B.so : B.zo;
    const int no = P::n_ort(c);
    DCOL_UNROLL_ROWS
    for (int i = 0; i < P::NO; ++i) {
        if (P::dyn && i >= no) break;
        if (o[i] <= 0.0) any_nonpos = true;
        mn = min_(mn, o[i]);
    }
}
template <class P, int SEL>
DCOL_HD void b2c_soc(const Block<P>& B, double& alpha)
{
    if (P::Q == 0) return;
    const double* q = SEL == 0 ? B.sq : B.zq;
    double nn = 0.0;
    DCOL_UNROLL
    for (int i = 1; i < P::Q; ++i) nn += q[i] * q[i];
    const double res = q[0] - sqrt(nn);
    if (res <= 0.0) alpha = max_(alpha, -res);
}
template <class P, int SEL>
DCOL_HD void b2c_shift(const typename P::Const& c, Block<P>& B, double shift)
{
    double* o = SEL == 0 ? B.so : B.zo;
    double* q = SEL == 0 ? B.sq : B.zq;
    const int no = P::n_ort(c);
    DCOL_UNROLL_ROWS
    for (int i = 0; i < P::NO; ++i) {
        if (P::dyn && i >= no) break;
        o[i] += shift;
    }
    if (P::Q > 0) q[0] += shift;
}

/* ---------------------------------------------------------------------------------------- */
/* results of one pair                                                                       */
template <int N>
struct PairResult {
    double x[N];
    double grad[12];
    int32_t iters, status;
};

/* optional per-iteration trace (debug entry point only) */
struct Trace {
    double* mu; /* [DCOL_MAX_ITER + 1] or null */
};

template <class P1, class P2>
struct Solver {
    static constexpr int N = 4 + P1::NE + P2::NE;
    static constexpr int CE1 = 4, CE2 = 4 + P1::NE;
    typedef typename P1::Const C1;
    typedef typename P2::Const C2;
    /* The solve runs in the body frame of the FIRST primitive: x~ = Q1'^T (x - r1').  The map is an orthogonal
     * change of the first three unknowns, under which the least-squares start, the Newton steps and even the
     * diag-only triangular solve of the dual start (it only involves the Schur complement of the x-block) are
     * invariant, so (s, z, alpha, extras) follow the same path; the first primitive then needs no rotation or
     * translation at all, and the second one carries its pose relative to the first. */
    /* ... unless the first primitive is a sphere (its rows need no rotation anyway) and the second is not:
     * then the second primitive's body frame is the working frame. */
    static constexpr bool kFrame2 = !P1::rot && P2::rot;
    typedef typename std::conditional<kFrame2, P1, typename P1::Ident>::type F1;
    typedef typename std::conditional<kFrame2, typename P2::Ident, P2>::type F2;

    F1 p1;
    F2 p2;
    Block<F1> b1;
    Block<F2> b2;
    double x[N]; /* (x~, alpha, extras) */

    /* ---- lower Cholesky of the symmetric M (upper triangle valid); Li holds 1/L_jj.
     * Unblocked, pivot test `<= 0` with NaN passing, as OpenBLAS potf2 behind numpy/scipy. */
    DCOL_HD static int chol(const double (&M)[N][N], double (&L)[N][N], double (&Li)[N])
    {
        int status = 0;
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) {
            double d = M[j][j];
            DCOL_UNROLL
            for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
            /* Branch-free (one basic block for the scheduler): the FIRST bad pivot decides the status, as the
             * early exits of potf2 / check_finite would; a non-finite entry of M reaches a pivot as NaN */
            if (status == 0 && !(d > 0.0)) status = (d != d) ? DCOL_STATUS_NON_FINITE : DCOL_STATUS_NOT_PD;
            const double ri = rsqrt_(d);
            L[j][j] = d * ri;
            Li[j] = ri;
            DCOL_UNROLL
            for (int i = j + 1; i < N; ++i) {
                double v = M[j][i];
                DCOL_UNROLL
                for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
                L[i][j] = v * ri;
            }
        }
        return status;
    }
    DCOL_HD static void chol_solve(const double (&L)[N][N], const double (&Li)[N], double (&v)[N])
    {
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) {
            double t = v[i];
            DCOL_UNROLL
            for (int k = 0; k < i; ++k) t -= L[i][k] * v[k];
            v[i] = t * Li[i];
        }
        DCOL_UNROLL
        for (int i = N - 1; i >= 0; --i) {
            double t = v[i];
            DCOL_UNROLL
            for (int k = i + 1; k < N; ++k) t -= L[k][i] * v[k];
            v[i] = t * Li[i];
        }
    }
    /* ---- the Newton systems use the square-root-free factor M = L D L^T of the same matrix (unit lower L; Di = 1 / d_j;
     * the strict upper triangle of L holds U_ik = L_ik d_k while factoring).  The pivots d_j are the squares of the
     * Cholesky pivots, so the `<= 0` / NaN tests decide exactly as above; what it buys is latency: a reciprocal instead
     * of a reciprocal square root per column, and substitutions that are pure multiply-add chains (the divisions by the
     * diagonal become one independent scaling between the two sweeps) — these chains are the serial part of an
     * iteration, which two warps per scheduler cannot hide.  The initial point keeps chol(): it needs the factor's
     * diagonal itself (pdip.py:326). */
    DCOL_HD static int ldlt(const double (&M)[N][N], double (&L)[N][N], double (&Di)[N])
    {
        int status = 0;
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) {
            double d = M[j][j];
            DCOL_UNROLL
            for (int k = 0; k < j; ++k) d -= L[j][k] * L[k][j];
            if (status == 0 && !(d > 0.0)) status = (d != d) ? DCOL_STATUS_NON_FINITE : DCOL_STATUS_NOT_PD;
            const double rd = rcp_(d);
            Di[j] = rd;
            DCOL_UNROLL
            for (int i = j + 1; i < N; ++i) {
                double v = M[j][i];
                DCOL_UNROLL
                for (int k = 0; k < j; ++k) v -= L[i][k] * L[k][j];
                L[j][i] = v;      /* U_ij = L_ij d_j */
                L[i][j] = v * rd;
            }
        }
        return status;
    }
    DCOL_HD static void ldlt_solve(const double (&L)[N][N], const double (&Di)[N], double (&v)[N])
    {
        DCOL_UNROLL
        for (int i = 1; i < N; ++i) {
            double t = v[i];
            DCOL_UNROLL
            for (int k = 0; k < i; ++k) t -= L[i][k] * v[k];
            v[i] = t;
        }
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) v[i] *= Di[i];
        DCOL_UNROLL
        for (int i = N - 2; i >= 0; --i) {
            double t = v[i];
            DCOL_UNROLL
            for (int k = i + 1; k < N; ++k) t -= L[k][i] * v[k];
            v[i] = t;
        }
    }
    DCOL_HD static double probe_vec(const double (&v)[N])
    {
        double t = 0.0;
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) t += nonfinite_probe(v[i]);
        return t;
    }

    /* ---- rows of G applied to a world vector: fills (ro, rq) = block of G v (- h if AFFINE) */
    template <class P, bool AFFINE>
    DCOL_HD static void rows(const P& p, const typename P::Const& c, int col_e, const double (&v)[N],
                             double (&ro)[P::NOA], double (&rq)[P::QA])
    {
        double xh[P::NL];
        p.template to_local<N, AFFINE>(v, col_e, xh);
        P::ort_apply(c, xh, ro);
        P::soc_apply(c, xh, rq);
    }
    /* out[N] += block of G^T (co, cq) */
    template <class P>
    DCOL_HD static void rows_t(const P& p, const typename P::Const& c, int col_e, const double (&co)[P::NOA],
                               const double (&cq)[P::QA], double (&out)[N])
    {
        double acc[P::NL];
        DCOL_UNROLL
        for (int j = 0; j < P::NL; ++j) acc[j] = 0.0;
        P::ort_apply_t(c, co, acc);
        P::soc_apply_t(c, cq, acc);
        p.template from_local_add<N>(acc, col_e, out);
    }

    /* ---- pdip.py:291-332, one primitive's share of M = G^T G and of G^T h.  In the primitive's local variables both
     * come from ONE constant of the shape, G0 = sum_i g_i g_i^T (+ the cone rows): G^T G = T^T G0 T, and because the rows
     * have no constant term in local variables, h_i = g_i[0:3] . u with u = Q'^T r' (minus the local coordinates of the
     * world origin), so G^T h = T^T (G0[:, 0:3] u). */
    template <class P>
    DCOL_HD static void init_accumulate(const P& p, const typename P::Const& c, int col_e, double (&M)[N][N],
                                        double (&gth)[N])
    {
        double Gl[P::NL][P::NL];
        {
            int at = 0;
            DCOL_UNROLL
            for (int i = 0; i < P::NL; ++i) {
                DCOL_UNROLL
                for (int j = i; j < P::NL; ++j) Gl[i][j] = c.G0[at++];
            }
        }
        p.template gram_from_local<N>(Gl, col_e, M);
        if constexpr (!P::ident) {
            double u[3], acc[P::NL];
            DCOL_UNROLL
            for (int j = 0; j < 3; ++j)
                u[j] = P::frot ? (p.Qp[0][j] * p.rp[0] + p.Qp[1][j] * p.rp[1] + p.Qp[2][j] * p.rp[2]) : p.rp[j];
            DCOL_UNROLL
            for (int j = 0; j < P::NL; ++j) {
                double t = 0.0;
                DCOL_UNROLL
                for (int k = 0; k < 3; ++k) t += (j <= k ? Gl[j][k] : Gl[k][j]) * u[k];
                acc[j] = t;
            }
            p.template from_local_add<N>(acc, col_e, gth);
        }
    }
    /* the shape constant G0 (host side, when the constant block of a launch is filled) */
    template <class P>
    static void gram0(typename P::Const& c)
    {
        double Gl[P::NL][P::NL];
        for (int i = 0; i < P::NL; ++i)
            for (int j = 0; j < P::NL; ++j) Gl[i][j] = 0.0;
        double ones[P::NOA];
        for (int i = 0; i < P::NOA; ++i) ones[i] = 1.0;
        P::ort_gram(c, ones, Gl);
        double I2[P::QA][P::QA];
        for (int i = 0; i < P::QA; ++i)
            for (int j = 0; j < P::QA; ++j) I2[i][j] = (i == j) ? 1.0 : 0.0;
        P::soc_gram(c, I2, Gl);
        int at = 0;
        for (int i = 0; i < 21; ++i) c.G0[i] = 0.0;
        for (int i = 0; i < P::NL; ++i)
            for (int j = i; j < P::NL; ++j) c.G0[at++] = Gl[i][j];
    }

    /* ---- NT scaling of one block + everything of pass A that does not need the Newton step.
     * Accumulates mu*deg, the Gram matrix G~^T G~, G^T z, G~^T (lambda - rho~) and G~^T (lambda^-1 o e). */
    template <class P>
    DCOL_HD static double nt_and_mu(const typename P::Const& c, Block<P>& B, double& bad)
    {
        double sz = 0.0;
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            const double pr = B.so[i] * B.zo[i];
            sz += pr;
            B.rinv[i] = rcp_(pr); /* NT_scaling.py:430: W_ort = sqrt(s/z), lambda = W z = sqrt(s z); only lambda^-2 is used */
        }
        if (P::Q > 0) {
            /* NT_scaling.py:340-405 */
            double dq = 0.0;
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) dq += B.sq[i] * B.zq[i];
            sz += dq;
            const double Js = socJ<P::Q>(B.sq), Jz = socJ<P::Q>(B.zq);
            const double rs = rsqrt_(Js), rz = rsqrt_(Jz);
            /* gamma = sqrt((1 + zbar.sbar) / 2);  wbar = (sbar + J zbar) / (2 gamma) */
            const double g2 = 0.5 * (1.0 + dq * (rs * rz));
            const double ig = 0.5 * rsqrt_(g2);
            B.wh[0] = (B.sq[0] * rs + B.zq[0] * rz) * ig;
            DCOL_UNROLL
            for (int i = 1; i < P::Q; ++i) B.wh[i] = -((B.sq[i] * rs - B.zq[i] * rz) * ig);
            B.bw = rcp_(B.wh[0] + 1.0);
            /* eta = (J(s)/J(z))^(1/4) */
            const double e2 = (Js * rs) * rz; /* sqrt(Js / Jz) */
            B.ieta = rsqrt_(e2);
            B.eta = e2 * B.ieta;
            /* one probe for the whole scaling: a sum is non-finite exactly when one of its terms is (a non-finite
             * ieta makes eta = e2 ieta non-finite as well) */
            double pr = B.eta + B.bw;
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) pr += B.wh[i];
            bad += nonfinite_probe(pr);
        }
        return sz;
    }

    template <class P>
    DCOL_HD static void pass_a(const P& p, const typename P::Const& c, int col_e, Block<P>& B, const double (&xx)[N],
                               double (&M)[N][N], double (&va)[N], double (&vl)[N])
    {
        double xh[P::NL], rq[P::QA], acc_a[P::NL], acc_l[P::NL];
        p.template to_local<N, true>(xx, col_e, xh); /* rows of G x - h are constant rows on xh */
        P::soc_apply(c, xh, rq);
        double Gl[P::NL][P::NL];
        DCOL_UNROLL
        for (int i = 0; i < P::NL; ++i) {
            acc_a[i] = acc_l[i] = 0.0;
            DCOL_UNROLL
            for (int j = 0; j < P::NL; ++j) Gl[i][j] = 0.0;
        }
        double qa[P::QA], ql[P::QA];
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            /* On the orthant the scaled directions are never formed: with the relative steps ds_i / s_i = ds~_i / lambda_i
             * and dz_i / z_i = dz~_i / lambda_i every product with w_i or lambda_i collapses into 1 / s_i = z_i / (s_i z_i)
             * and w_i^-2 = z_i / s_i, both line searches read the relative steps directly, and no square root is needed
             * (same Newton system and step lengths as NT_scaling.py:430 / pdip.py:7-22, 424-466 in exact arithmetic) */
            const double is = B.zo[i] * B.rinv[i];                        /* 1 / s_i                       */
            const double rz = P::row_dot(c, i, xh, B.so[i]);              /* rz_i = s_i + (G x - h)_i      */
            B.ta[i] = rz;
            /* Gram += w^-2 g g^T and acc_a -= w^-2 rz g  (b~_affine = lambda - rho~; G~^T lambda = G^T z cancels in bx) */
            P::row_gram_axpy(c, i, B.zo[i] * is, rz, Gl, acc_a);
            P::row_axpy(c, i, is, acc_l);                                 /* W^-1 (lambda^-1 o e) = g / s_i */
        }
        if (P::Q > 0) {
            /* lambda = W z */
            wbar_apply<P::Q>(B.wh, B.bw, -1.0, B.zq, B.eta, B.lam);
            /* rz = s + (G x - h) is kept UNSCALED (in tq): the Newton steps need W^-1 (G dx + rz) = G~ dx + rho~ and the
             * update of s needs ds = -(G dx + rz) itself (primal equation); rho~ = W^-1 rz is only a temporary below */
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) B.tq[i] = B.sq[i] + rq[i];
            /* W^-2 for the Gram block.  Wbar(w)^2 = 2 w w^T - J for w^T J w = 1, so
             * W^-2 = (2 wh wh^T - J) / eta^2: ten products instead of forming W^-1 and squaring it. */
            double W2[P::QA][P::QA];
            {
                const double ie2 = B.ieta * B.ieta, t2 = 2.0 * ie2;
                DCOL_UNROLL
                for (int i = 0; i < P::Q; ++i) {
                    const double ti = t2 * B.wh[i];
                    DCOL_UNROLL
                    for (int j = i; j < P::Q; ++j)
                        W2[i][j] = (i == j) ? fma(ti, B.wh[j], i == 0 ? -ie2 : ie2) : ti * B.wh[j];
                }
            }
            P::soc_gram(c, W2, Gl);
            /* shared scalars of lambda: inverse cone product (pdip.py:108-118) and line search (pdip.py:39-47) */
            const double Jl = socJ<P::Q>(B.lam);
            B.irho = rcp_(Jl);
            B.il0 = rcp_(B.lam[0]);
            const double nu = max_(Jl, 1e-25);
            B.ls_isn = rsqrt_(nu);
            B.ls_inu = B.ls_isn * B.ls_isn;
            B.ls_c0 = rcp_(B.lam[0] * B.ls_isn + 1.0);
            {
                /* -W^-1 rho~ by two applications of W^-1: the product with the W^-2 above would save one of them, but W^-2
                 * squares wh, which is large for iterates near the cone's boundary (measured on the host twin: the largest
                 * gradient deviation from the oracle over 20,000 pairs goes from 7e-8 to 1.2e-7) */
                double rho[P::QA], t1[P::QA];
                wbar_apply<P::Q>(B.wh, B.bw, 1.0, B.tq, B.ieta, rho); /* rho~ = W^-1 rz */
                DCOL_UNROLL
                for (int i = 0; i < P::Q; ++i) t1[i] = -rho[i];
                wbar_apply<P::Q>(B.wh, B.bw, 1.0, t1, B.ieta, qa);
            }
            /* W^-1 lambda^-1 for the centring term.  In exact arithmetic this is s^-1 = (s_0, -s_v) / J(s) ((W u)^-1 = W^-1 u^-1
             * and W lambda = s), which would save this W-apply (+2.8 % throughput, measured) — but J(s) cancels badly for a slack
             * near the cone's boundary, where J(lambda) = sqrt(J(s) J(z)) does not: over 8 M config-5 pairs the largest gradient
             * deviation from the oracle went from 4e-8 to 1.3e-6 (310 pairs above 1e-7 instead of none), so it is NOT used */
            {
                double t2[P::QA];
                DCOL_UNROLL
                for (int i = 0; i < P::Q; ++i) t2[i] = (i == 0 ? B.lam[i] : -B.lam[i]) * B.irho;
                wbar_apply<P::Q>(B.wh, B.bw, 1.0, t2, B.ieta, ql);
            }
        }
        P::soc_apply_t(c, qa, acc_a);
        P::soc_apply_t(c, ql, acc_l);
        p.template gram_from_local<N>(Gl, col_e, M);
        p.template from_local_add<N>(acc_a, col_e, va);
        p.template from_local_add<N>(acc_l, col_e, vl);
    }

    /* ---- line-search measure of a scaled direction against lambda: the step is 1/t for t > 1.
     * orthant pdip.py:7-22, soc pdip.py:25-52 */
    template <class P>
    DCOL_HD static double soc_ls(const Block<P>& B, const double (&d)[P::QA])
    {
        double zeta = B.lam[0] * d[0];
        DCOL_UNROLL
        for (int i = 1; i < P::Q; ++i) zeta -= B.lam[i] * d[i];
        const double rho0 = zeta * B.ls_inu;
        /* rho_v = d_v / sqrt(nu) - coef lambda_v / nu = (d_v - (coef / sqrt(nu)) lambda_v) / sqrt(nu): the common factor
         * 1 / sqrt(nu) > 0 leaves the norm once */
        const double cs = ((zeta * B.ls_isn + d[0]) * B.ls_c0) * B.ls_isn;
        double nn = 0.0;
        DCOL_UNROLL
        for (int i = 1; i < P::Q; ++i) {
            const double rv = fma(-cs, B.lam[i], d[i]);
            nn += rv * rv;
        }
        return sqrt_(nn) * B.ls_isn - rho0;
    }
    /* The two searches of the AFFINE step at once.  There ds~ + dz~ = -lambda, and with nu = J(lambda)
     * the vector part of rho(ds~) is minus that of rho(dz~): zeta_s = -nu - zeta_z, coef_s = -sqrt(nu) - coef_z, so
     * ds~_i / sqrt(nu) - coef_s lambda_i / nu = -(dz~_i / sqrt(nu) - coef_z lambda_i / nu).  One norm serves both; only
     * rho_0 differs.  Same values as two calls of soc_ls up to rounding — except below the 1e-25 floor that
     * pdip.py:39-47 puts under J(lambda), where nu is no longer J(lambda): that needs an iterate within 1e-25 of the
     * cone's boundary, i.e. one that has already broken down (J(lambda) = sqrt(J(s) J(z)) stays of the order of mu, and
     * the solver stops at mu < tol), and the failure statuses are decided elsewhere (finiteness of the scaling, pivots). */
    template <class P>
    DCOL_HD static void soc_ls_affine(const Block<P>& B, const double (&ds)[P::QA], const double (&dz)[P::QA], double& ms,
                                      double& mz)
    {
        double zeta_z = B.lam[0] * dz[0], zeta_s = B.lam[0] * ds[0];
        DCOL_UNROLL
        for (int i = 1; i < P::Q; ++i) {
            zeta_z -= B.lam[i] * dz[i];
            zeta_s -= B.lam[i] * ds[i];
        }
        const double cs = ((zeta_z * B.ls_isn + dz[0]) * B.ls_c0) * B.ls_isn;
        double nn = 0.0;
        DCOL_UNROLL
        for (int i = 1; i < P::Q; ++i) {
            const double rv = fma(-cs, B.lam[i], dz[i]);
            nn += rv * rv;
        }
        const double r = sqrt_(nn) * B.ls_isn;
        mz = r - zeta_z * B.ls_inu;
        ms = r - zeta_s * B.ls_inu;
    }

    /* ---- pass B: affine direction of one block, its line-search measures, the three dot products of
     * the centring ratio, k = lambda^-1 o (ds~ o dz~) and G~^T k */
    template <class P>
    DCOL_HD static void pass_b(const P& p, const typename P::Const& c, int col_e, Block<P>& B, const double (&dx)[N],
                               double (&tm)[2], double& d_sz, double (&vk)[N])
    {
        double xh[P::NL], rq[P::QA], acc_k[P::NL], qk[P::QA];
        p.template to_local<N, false>(dx, col_e, xh); /* G dx */
        P::soc_apply(c, xh, rq);
        DCOL_UNROLL
        for (int i = 0; i < P::NL; ++i) acc_k[i] = 0.0;
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            const double is = B.zo[i] * B.rinv[i];
            /* ds = -rz - G dx (primal equation) and s dz + z ds = -s z (complementarity), as relative steps */
            const double ds = -(is * P::row_dot(c, i, xh, B.ta[i])); /* ds_i / s_i */
            const double dz = -1.0 - ds;                              /* dz_i / z_i */
            /* both searches: max(-ds/s, -dz/z) */
            tm[i & 1] = max_(tm[i & 1], -min_(ds, dz));
            const double t = ds * dz;        /* ds~ dz~ / lambda^2 */
            const double zt = B.zo[i] * t;
            d_sz = fma(B.so[i], zt, d_sz);   /* <ds~, dz~> = sum s z t */
            B.tb[i] = t;
            P::row_axpy(c, i, zt, acc_k);    /* W^-1 (lambda^-1 o (ds~ o dz~)) = g z t */
        }
        if (P::Q > 0) {
            double g[P::QA], dz[P::QA], ds[P::QA], w[P::QA], v[P::QA];
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) v[i] = rq[i] + B.tq[i];
            wbar_apply<P::Q>(B.wh, B.bw, 1.0, v, B.ieta, g); /* G~ dx + rho~ = -ds~ (primal equation) */
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) {
                ds[i] = -g[i];
                dz[i] = g[i] - B.lam[i]; /* ds~ + dz~ = -lambda (complementarity) */
            }
            {
                double ms, mz;
                soc_ls_affine<P>(B, ds, dz, ms, mz);
                tm[0] = max_(tm[0], ms);
                tm[1] = max_(tm[1], mz);
            }
            /* w = ds~ o dz~ (pdip.py:165-200); k = lambda^-1 o w (pdip.py:88-122) */
            w[0] = 0.0;
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) w[0] += ds[i] * dz[i];
            d_sz += w[0]; /* <ds~, dz~> of this block is the scalar part of the cone product */
            DCOL_UNROLL
            for (int i = 1; i < P::Q; ++i) w[i] = ds[0] * dz[i] + dz[0] * ds[i];
            double nu = 0.0;
            DCOL_UNROLL
            for (int i = 1; i < P::Q; ++i) nu += B.lam[i] * w[i];
            B.kq[0] = B.irho * (B.lam[0] * w[0] - nu);
            const double c1 = nu * B.il0 - w[0], c2 = B.il0; /* (rho / lambda_0) / rho */
            DCOL_UNROLL
            for (int i = 1; i < P::Q; ++i) B.kq[i] = B.irho * (c1 * B.lam[i]) + c2 * w[i];
            wbar_apply<P::Q>(B.wh, B.bw, 1.0, B.kq, B.ieta, qk);
        }
        P::soc_apply_t(c, qk, acc_k);
        p.template from_local_add<N>(acc_k, col_e, vk);
    }

    /* ---- pass C: corrector direction of one block (orthant: relative steps in ta / tb; cone: unscaled ds in tq, dz~ in
     * kq) and its line-search measures */
    template <class P>
    DCOL_HD static void pass_c(const P& p, const typename P::Const& c, int col_e, Block<P>& B, const double (&dx)[N],
                               double sigmu, double (&tm)[2])
    {
        double xh[P::NL], rq[P::QA];
        p.template to_local<N, false>(dx, col_e, xh);
        P::soc_apply(c, xh, rq);
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            const double ip = B.rinv[i];
            const double is = B.zo[i] * ip;
            const double d = fma(sigmu, ip, -1.0 - B.tb[i]);          /* (lambda^-1 o ds) / lambda = -1 - t + sigma mu / (s z) */
            const double ds = -(is * P::row_dot(c, i, xh, B.ta[i])); /* ds_i / s_i (primal equation) */
            const double dz = d - ds;                                 /* dz_i / z_i */
            tm[i & 1] = max_(tm[i & 1], -min_(ds, dz));
            B.ta[i] = ds;
            B.tb[i] = dz;
        }
        if (P::Q > 0) {
            double g[P::QA], dz[P::QA], ds[P::QA], v[P::QA];
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) v[i] = rq[i] + B.tq[i];
            wbar_apply<P::Q>(B.wh, B.bw, 1.0, v, B.ieta, g); /* -ds~ (primal equation) */
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) {
                const double li = (i == 0 ? B.lam[i] : -B.lam[i]) * B.irho; /* lambda^-1 o e */
                const double d = -B.lam[i] - B.kq[i] + sigmu * li;          /* ds~ + dz~ */
                ds[i] = -g[i];
                dz[i] = d + g[i];
            }
            tm[0] = max_(tm[0], soc_ls<P>(B, ds));
            tm[1] = max_(tm[1], soc_ls<P>(B, dz));
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) {
                B.tq[i] = -v[i]; /* ds = W ds~, unscaled: -(G dx + rz) */
                B.kq[i] = dz[i];
            }
        }
    }

    /* ---- s += a W ds~,  z += a W^-1 dz~   pdip.py:464-466 */
    template <class P>
    DCOL_HD static void pass_d(const typename P::Const& c, Block<P>& B, double a)
    {
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            B.so[i] = fma(a * B.ta[i], B.so[i], B.so[i]); /* s (1 + a ds/s) */
            B.zo[i] = fma(a * B.tb[i], B.zo[i], B.zo[i]);
        }
        if (P::Q > 0) {
            double dz[P::QA];
            wbar_apply<P::Q>(B.wh, B.bw, 1.0, B.kq, B.ieta, dz);
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) {
                B.sq[i] += a * B.tq[i]; /* pass_c left the unscaled ds here */
                B.zq[i] += a * dz[i];
            }
        }
    }

    /* ---- gradient share of one primitive: d/d(r, p) of z^T (G(theta) x - h(theta)), (x, z) frozen.
     * proximity_gradient.py:8-88 by the chain rule instead of finite differences.  pw: the primitive at its
     * WORLD pose; xw: the solution in world coordinates; Rf: rotation of the working frame (world pose of the
     * first primitive), which carries the cone components of a primitive whose rows are not rotated (sphere). */
    template <class P, class BL>
    DCOL_HD static void grad_block(const P& pw, const typename P::Const& c, int col_e, const BL& B, const double (&xw)[N],
                                   const double (&Rf)[3][3], const double pm[3], const double Qm[3][3], double* g6)
    {
        /* u = y-part of the local adjoint product of the rows whose coefficients are constant in the body
         * frame (all orthant rows, the cone's and the ellipsoid's soc rows); the ball soc rows -(x - r') + Q' E e
         * are written in world axes by the reference, so their dual is frozen in world components: zw. */
        double acc[P::NL];
        DCOL_UNROLL
        for (int j = 0; j < P::NL; ++j) acc[j] = 0.0;
        P::ort_apply_t(c, B.zo, acc);
        if (P::Q > 0 && !P::ball) P::soc_apply_t(c, B.zq, acc);
        double d[3], gr[3], Mq[3][3];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) d[i] = xw[i] - pw.rp[i];
        /* gr = dL/dr' = -Q' u (+ zw);  Mq = dL/dQ(p) = d (Q_off u)^T + zw (Q_off ehat)^T + gr r_off^T */
        double zw[3] = { 0.0, 0.0, 0.0 }, eh[3] = { 0.0, 0.0, 0.0 };
        if (P::ball) {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i)
                zw[i] = P::rot ? (pw.Qp[i][0] * B.zq[1] + pw.Qp[i][1] * B.zq[2] + pw.Qp[i][2] * B.zq[3])
                               : (Rf[i][0] * B.zq[1] + Rf[i][1] * B.zq[2] + Rf[i][2] * B.zq[3]);
            DCOL_UNROLL
            for (int j = 0; j < P::NE; ++j) eh[j] = xw[col_e + j];
        }
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i)
            gr[i] = zw[i] - (P::rot ? (pw.Qp[i][0] * acc[0] + pw.Qp[i][1] * acc[1] + pw.Qp[i][2] * acc[2]) : acc[i]);
        if (c.no_offset) { /* warp-uniform: Q_off = I, r_off = 0 */
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) Mq[i][j] = P::rot ? d[i] * acc[j] + zw[i] * eh[j] : 0.0;
            }
        } else {
            double qu[3], qe[3];
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                qu[i] = c.Q_off[i][0] * acc[0] + c.Q_off[i][1] * acc[1] + c.Q_off[i][2] * acc[2];
                qe[i] = c.Q_off[i][0] * eh[0] + c.Q_off[i][1] * eh[1] + c.Q_off[i][2] * eh[2];
            }
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) Mq[i][j] = (P::rot ? d[i] * qu[j] + zw[i] * qe[j] : 0.0) + gr[i] * c.r_off[j];
            }
        }
        g6[0] = gr[0];
        g6[1] = gr[1];
        g6[2] = gr[2];
        dcm_derivative_contract(pm, Qm, Mq, g6 + 3);
    }

    /* ---- the solve in three pieces: solve() below runs them in sequence; the lane-refill kernel (dcol_kernels.cuh)
     * runs init_point for 32 new pairs at a time and scale_and_check / newton_step in lanes that are at different
     * iterations of different pairs.  Same operations in the same order either way. */
    static constexpr int kContinue = -1;

    /* pose = (r, p) -> working frame and the initial point of pdip.py:291-332.  Returns 0 or a failure status. */
    DCOL_HD int init_point(const C1& c1, const C2& c2, const double* pose1, const double* pose2)
    {
        {
            double Q1[3][3], Q2[3][3];
            dcm_from_mrp(pose1 + 3, Q1);
            dcm_from_mrp(pose2 + 3, Q2);
            P1 w1; /* world poses; one of them is the working frame, the other becomes relative to it */
            P2 w2;
            w1.set_pose(c1, pose1, Q1);
            w2.set_pose(c2, pose2, Q2);
            set_relative(w1, w2);
        }
        double L[N][N], Li[N];
        double M[N][N], gth[N];
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) {
            gth[i] = 0.0;
            DCOL_UNROLL
            for (int j = 0; j < N; ++j) M[i][j] = 0.0;
        }
        init_accumulate<F1>(p1, c1, CE1, M, gth);
        init_accumulate<F2>(p2, c2, CE2, M, gth);
        if (int bad = chol(M, L, Li)) return bad; /* numpy cholesky -> LinAlgError; check_finite */
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) x[j] = gth[j];
        chol_solve(L, Li, x); /* x_hat = (G^T G)^-1 G^T h */
        if (probe_vec(x) != 0.0) return DCOL_STATUS_NON_FINITE;
        rows<F1, true>(p1, c1, CE1, x, b1.so, b1.sq);
        rows<F2, true>(p2, c2, CE2, x, b2.so, b2.sq); /* s~ = G x_hat - h */
        /* solve_triangular(F, -c) reads the UPPER triangle of the lower factor, i.e. its diagonal
         * (pdip.py:326); then a proper back substitution with F^T */
        double xd[N];
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) xd[i] = (i == 3) ? -Li[3] : 0.0;
        DCOL_UNROLL
        for (int i = N - 1; i >= 0; --i) {
            double t = xd[i];
            DCOL_UNROLL
            for (int k = i + 1; k < N; ++k) t -= L[k][i] * xd[k];
            xd[i] = t * Li[i];
        }
        rows<F1, false>(p1, c1, CE1, xd, b1.zo, b1.zq);
        rows<F2, false>(p2, c2, CE2, xd, b2.zo, b2.zq); /* z~ = G x */
        bring2cone<0>(c1, c2);
        bring2cone<1>(c1, c2);
        return 0;
    }

    /* top of iteration `it` (pdip.py:396-422): NT scaling of the current iterate, sz = s'z, the finiteness checks and
     * the only convergence test, mu < tol.  Returns kContinue, DCOL_STATUS_OK (converged) or a failure status;
     * `iters` is what the batch API reports for that outcome. */
    DCOL_HD int scale_and_check(const C1& c1, const C2& c2, double tol, int it, int32_t& iters, double& sz,
                                const Trace* trace)
    {
        iters = it;
        double bad = 0.0;
        sz = nt_and_mu<F1>(c1, b1, bad);
        sz += nt_and_mu<F2>(c2, b2, bad);
        if (nonfinite_probe(sz) != 0.0) {
            /* the previous step left a non-finite iterate: the reference's check_finite raised inside it */
            iters = it > 0 ? it - 1 : 0;
            return DCOL_STATUS_NON_FINITE;
        }
        if (bad != 0.0) return DCOL_STATUS_NON_FINITE; /* cho_factor(W_soc) check_finite */
        const double mu = sz * ideg_of(c1, c2);
        if (trace && trace->mu) trace->mu[it] = mu;
        if (mu < tol) return DCOL_STATUS_OK; /* the only convergence test, pdip.py:418-422 */
        return kContinue;
    }

    DCOL_HD static double ideg_of(const C1& c1, const C2& c2)
    {
        const int deg = P1::n_ort(c1) + P2::n_ort(c2) + (P1::Q > 0) + (P2::Q > 0);
        return 1.0 / (double)deg;
    }

    /* one predictor-corrector step from the scaled iterate (pdip.py:424-466).  Returns 0 or a failure status. */
    DCOL_HD int newton_step(const C1& c1, const C2& c2, double sz)
    {
        const double mu = sz * ideg_of(c1, c2);
        double L[N][N], Li[N];
        double M[N][N], va[N], vl[N];
        DCOL_UNROLL
        for (int i = 0; i < N; ++i) {
            va[i] = (i == 3) ? -1.0 : 0.0; /* -c; pass_a adds -G^T W^-2 rz */
            vl[i] = 0.0;
            DCOL_UNROLL
            for (int j = 0; j < N; ++j) M[i][j] = 0.0;
        }
        pass_a<F1>(p1, c1, CE1, b1, x, M, va, vl);
        pass_a<F2>(p2, c2, CE2, b2, x, M, va, vl);
        double dx[N];
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) dx[j] = va[j]; /* bx + G~^T b~ */
        if (int bad = ldlt(M, L, Li)) return bad; /* scipy cholesky: check_finite, LinAlgError */
        ldlt_solve(L, Li, dx);

        /* affine step: un-damped line search, sigma = clip(rho, 0, 1)^3   pdip.py:446-448 */
        double tm[2] = { 0.0, 0.0 }, d_sz = 0.0, vk[N];
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) vk[j] = 0.0;
        pass_b<F1>(p1, c1, CE1, b1, dx, tm, d_sz, vk);
        pass_b<F2>(p2, c2, CE2, b2, dx, tm, d_sz, vk);
        double t = max_(tm[0], tm[1]);
        double a = t > 1.0 ? rcp_(t) : 1.0;
        /* (s + a ds)'(z + a dz) / s'z with ds~ + dz~ = -lambda for the affine direction: 1 - a + a^2 <ds~, dz~> / s'z */
        const double rho = (1.0 - a) + (a * a) * (d_sz * rcp_(sz));
        const double cl = max_(0.0, min_(1.0, rho));
        const double sigmu = (cl * cl * cl) * mu;

        /* corrector: rhs = rhs_affine + G~^T k - sigma mu G~^T (lambda^-1 o e), same factor   pdip.py:450-460 */
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) dx[j] = va[j] + vk[j] - sigmu * vl[j];
        ldlt_solve(L, Li, dx);
        tm[0] = 0.0;
        tm[1] = 0.0;
        pass_c<F1>(p1, c1, CE1, b1, dx, sigmu, tm);
        pass_c<F2>(p2, c2, CE2, b2, dx, sigmu, tm);
        t = max_(tm[0], tm[1]);
        a = min_(1.0, 0.99 * (t > 1.0 ? rcp_(t) : 1.0)); /* pdip.py:462 */
        DCOL_UNROLL
        for (int j = 0; j < N; ++j) x[j] += a * dx[j];
        pass_d<F1>(c1, b1, a);
        pass_d<F2>(c2, b2, a);
        return 0;
    }

    /* what the reference reports when the iteration cap is reached (pdip.py:470), or when the last step itself went
     * non-finite */
    DCOL_HD int cap_status(int max_iter, int32_t& iters) const
    {
        iters = max_iter;
        if (probe_vec(x) != 0.0) {
            iters = max_iter - 1;
            return DCOL_STATUS_NON_FINITE;
        }
        return DCOL_STATUS_MAX_ITER;
    }

    /* ---- the whole solve.  pose = (r, p).  Returns the status word; fills res. */
    DCOL_HD int solve(const C1& c1, const C2& c2, const double* pose1, const double* pose2, double tol, int max_iter,
                      bool want_grad, PairResult<N>& res, const Trace* trace)
    {
        res.iters = 0;
        if (int bad = init_point(c1, c2, pose1, pose2)) return res.status = bad;
        for (int it = 0; it < max_iter; ++it) {
            double sz;
            const int st = scale_and_check(c1, c2, tol, it, res.iters, sz, trace);
            if (st != kContinue) return res.status = st;
            if (int bad = newton_step(c1, c2, sz)) return res.status = bad;
        }
        return res.status = cap_status(max_iter, res.iters);
    }

    template <int SEL>
    DCOL_HD void bring2cone(const C1& c1, const C2& c2)
    {
        double alpha = -1.0, mn = INFINITY;
        bool any = false;
        b2c_scan<F1, SEL>(c1, b1, any, mn);
        b2c_scan<F2, SEL>(c2, b2, any, mn);
        if (any) alpha = -mn;
        b2c_soc<F1, SEL>(b1, alpha);
        b2c_soc<F2, SEL>(b2, alpha);
        if (alpha < 0.0) return;
        b2c_shift<F1, SEL>(c1, b1, 1.0 + alpha);
        b2c_shift<F2, SEL>(c2, b2, 1.0 + alpha);
    }

    DCOL_HD void set_relative(const P1& w1, const P2& w2)
    {
        if constexpr (kFrame2) {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                p1.rp[i] = w1.rp[i];
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) p1.Qp[i][j] = w1.Qp[i][j];
            }
            p1.make_relative(w2.Qp, w2.rp);
        } else {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                p2.rp[i] = w2.rp[i];
                DCOL_UNROLL
                for (int j = 0; j < 3; ++j) p2.Qp[i][j] = w2.Qp[i][j];
            }
            p2.make_relative(w1.Qp, w1.rp);
        }
    }

    /* the solution in world coordinates: x = Qf x~ + rf (alpha and the extras are frame independent) */
    DCOL_HD void world_frames(const C1& c1, const C2& c2, const double* pose1, const double* pose2, P1& w1, P2& w2,
                              double (&Q1)[3][3], double (&Q2)[3][3], double (&xw)[N]) const
    {
        dcm_from_mrp(pose1 + 3, Q1);
        dcm_from_mrp(pose2 + 3, Q2);
        w1.set_pose(c1, pose1, Q1);
        w2.set_pose(c2, pose2, Q2);
        const double (&Qf)[3][3] = kFrame2 ? w2.Qp : w1.Qp;
        const double (&rf)[3] = kFrame2 ? w2.rp : w1.rp;
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) xw[i] = rf[i] + (Qf[i][0] * x[0] + Qf[i][1] * x[1] + Qf[i][2] * x[2]);
        DCOL_UNROLL
        for (int j = 3; j < N; ++j) xw[j] = x[j];
    }

    /* contact point x[0:3] in world coordinates (proximity.py:52) */
    DCOL_HD void contact_point(const C1& c1, const C2& c2, const double* pose1, const double* pose2, double (&out)[3]) const
    {
        double Qm[3][3];
        dcm_from_mrp((kFrame2 ? pose2 : pose1) + 3, Qm);
        typename std::conditional<kFrame2, P2, P1>::type wf;
        if constexpr (kFrame2) wf.set_pose(c2, pose2, Qm);
        else wf.set_pose(c1, pose1, Qm);
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) out[i] = wf.rp[i] + (wf.Qp[i][0] * x[0] + wf.Qp[i][1] * x[1] + wf.Qp[i][2] * x[2]);
    }

    /* second = false: only grad[0:6] = d alpha / d [r1 p1] is filled (DCOL_WANT_GRAD1) */
    DCOL_HD void gradient(const C1& c1, const C2& c2, const double* pose1, const double* pose2, double* grad,
                          bool second = true) const
    {
        double Q1[3][3], Q2[3][3], xw[N];
        P1 w1;
        P2 w2;
        world_frames(c1, c2, pose1, pose2, w1, w2, Q1, Q2, xw);
        const double (&Rf)[3][3] = kFrame2 ? w2.Qp : w1.Qp;
        grad_block<P1>(w1, c1, CE1, b1, xw, Rf, pose1 + 3, Q1, grad);
        if (second) grad_block<P2>(w2, c2, CE2, b2, xw, Rf, pose2 + 3, Q2, grad + 6);
    }

    /* ------------------------------------------------------------------------------------------------
     * Solution Jacobian d(contact point, alpha) / d[r1 p1 r2 p2]  (Report.pdf
     * section 2.3 eq. 4-6: differentiate the relaxed KKT system at the returned central-path point).
     * EXTENSION: the reference only differentiates the frozen-(x, z) Lagrangian (proximity_gradient.py:8-88).
     *
     * With a symmetric linearisation of the complementarity condition, ds = -W^2 dz:
     *     G^T dz = -dG^T z,   G dx + ds = -(dG x - dh)   =>   dx = -M^-1 (dG^T z + G^T W^-2 (dG x - dh)),
     * M = G^T W^-2 G: the reduced KKT matrix of the Newton steps, rebuilt and factored at the final iterate.  Only
     * four rows of dx are wanted, so the solve is an adjoint one: y_k = M^-1 e_k (four back-substitutions with the
     * factor), v_k = W^-2 G y_k, and then row k is minus the pose derivative of the bilinear form
     *     z^T G(theta) y_k + v_k^T (G(theta) x - h(theta))            (x, y_k, z, v_k frozen),
     * i.e. two Lagrangian-like forms of exactly the kind gradient() differentiates analytically.
     *
     * Which W: on the orthant W^2 = s/z is the exact linearisation of s_i z_i = const.  On a second-order cone the
     * Nesterov-Todd W^2 = eta^2 (2 wbar wbar^T - J) is the exact linearisation of s o z = mu e only ON the central
     * path; the returned iterate is off-centre (the solver never re-centres, pdip.py:418-422), and there the NT
     * scale eta^2 = sqrt(J(s)/J(z)) misstates the curvature of the cone's tangent plane by sqrt(eps_s/eps_z)
     * (measured: contact-point rows off by factors of 2-6 under rotations).  The exact linearisation
     * Arw(z) ds + Arw(s) dz = 0 acts on that tangent plane as ds = -(s_0/z_0) dz; keeping the NT direction wbar
     * and replacing eta^2 by s_0/z_0 (equal to eta^2 on the central path, where both are mu/J(z)) gives a symmetric
     * positive-definite W^2 with the same limit, so the Cholesky machinery of the solver is reused unchanged.
     * Against central differences of 1e-12 solves it is as accurate as the unsymmetric exact linearisation
     * (tests/test_jacobian.py). */
    template <class P, class BL>
    DCOL_HD static void jac_v(const P& p, const typename P::Const& c, int col_e, const BL& B, const double (&yt)[N],
                              double (&vo)[P::NOA], double (&vq)[P::QA])
    {
        double rq[P::QA];
        rows<P, false>(p, c, col_e, yt, vo, rq); /* G y (working frame: same rows) */
        const int no = P::n_ort(c);
        DCOL_UNROLL_ROWS
        for (int i = 0; i < P::NO; ++i) {
            if (P::dyn && i >= no) break;
            vo[i] *= B.zo[i] * (B.zo[i] * B.rinv[i]); /* w_i^-2 = z_i / s_i = z_i^2 / (s_i z_i) */
        }
        if (P::Q > 0) {
            /* W^-2 = (2 wh wh^T - J) / eta^2, as in pass_a */
            const double ie2 = B.ieta * B.ieta;
            double d = 0.0;
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) d += B.wh[i] * rq[i];
            DCOL_UNROLL
            for (int i = 0; i < P::Q; ++i) vq[i] = ie2 * (2.0 * d * B.wh[i] + (i == 0 ? -rq[i] : rq[i]));
        }
    }
    /* g6 = d/d(r, p) [ z^T G(theta) yw + v^T (G(theta) xw - h(theta)) ] for one primitive at its WORLD pose pw
     * (same chain rule as grad_block; the linear form has no r' dependence) */
    template <class P, class BL>
    DCOL_HD static void jac_block(const P& pw, const typename P::Const& c, int col_e, const BL& B,
                                  const double (&vo)[P::NOA], const double (&vq)[P::QA], const double (&xw)[N],
                                  const double (&yw)[N], const double (&Rf)[3][3], const double pm[3],
                                  const double Qm[3][3], double* g6)
    {
        double az[P::NL], av[P::NL];
        DCOL_UNROLL
        for (int j = 0; j < P::NL; ++j) az[j] = av[j] = 0.0;
        P::ort_apply_t(c, B.zo, az);
        P::ort_apply_t(c, vo, av);
        if (P::Q > 0 && !P::ball) {
            P::soc_apply_t(c, B.zq, az);
            P::soc_apply_t(c, vq, av);
        }
        double d[3], gr[3], Mq[3][3];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) d[i] = xw[i] - pw.rp[i];
        double zwz[3] = { 0.0, 0.0, 0.0 }, zwv[3] = { 0.0, 0.0, 0.0 }, ex[3] = { 0.0, 0.0, 0.0 }, ey[3] = { 0.0, 0.0, 0.0 };
        if (P::ball) {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) {
                const double (&R)[3][3] = P::rot ? pw.Qp : Rf;
                zwz[i] = R[i][0] * B.zq[1] + R[i][1] * B.zq[2] + R[i][2] * B.zq[3];
                zwv[i] = R[i][0] * vq[1] + R[i][1] * vq[2] + R[i][2] * vq[3];
            }
            DCOL_UNROLL
            for (int j = 0; j < P::NE; ++j) {
                ex[j] = xw[col_e + j];
                ey[j] = yw[col_e + j];
            }
        }
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i)
            gr[i] = zwv[i] - (P::rot ? (pw.Qp[i][0] * av[0] + pw.Qp[i][1] * av[1] + pw.Qp[i][2] * av[2]) : av[i]);
        double quz[3], quv[3], qex[3], qey[3];
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) {
            quz[i] = c.Q_off[i][0] * az[0] + c.Q_off[i][1] * az[1] + c.Q_off[i][2] * az[2];
            quv[i] = c.Q_off[i][0] * av[0] + c.Q_off[i][1] * av[1] + c.Q_off[i][2] * av[2];
            qex[i] = c.Q_off[i][0] * ex[0] + c.Q_off[i][1] * ex[1] + c.Q_off[i][2] * ex[2];
            qey[i] = c.Q_off[i][0] * ey[0] + c.Q_off[i][1] * ey[1] + c.Q_off[i][2] * ey[2];
        }
        DCOL_UNROLL
        for (int i = 0; i < 3; ++i) {
            DCOL_UNROLL
            for (int j = 0; j < 3; ++j)
                Mq[i][j] = (P::rot ? d[i] * quv[j] + zwv[i] * qex[j] + yw[i] * quz[j] + zwz[i] * qey[j] : 0.0) +
                           gr[i] * c.r_off[j];
        }
        g6[0] = gr[0];
        g6[1] = gr[1];
        g6[2] = gr[2];
        dcm_derivative_contract(pm, Qm, Mq, g6 + 3);
    }
    /* jac[4][12]: rows (contact x, y, z, alpha), columns [r1 p1 r2 p2].  Call after a solve that returned
     * DCOL_STATUS_OK (the NT scaling of the final iterate is still in the blocks).  Returns 0, or the status of
     * the factorisation of M at the final iterate (jac is then left untouched). */
    DCOL_HD int jacobian(const C1& c1, const C2& c2, const double* pose1, const double* pose2, double* jac)
    {
        double L[N][N], Li[N];
        if (F1::Q > 0) { /* eta^2 <- s_0 / z_0 (see above); pass_a and jac_v read the scale from the block */
            b1.ieta = sqrt(b1.zq[0] / b1.sq[0]);
            b1.eta = 1.0 / b1.ieta;
        }
        if (F2::Q > 0) {
            b2.ieta = sqrt(b2.zq[0] / b2.sq[0]);
            b2.eta = 1.0 / b2.ieta;
        }
        {
            double M[N][N], va[N], vl[N];
            DCOL_UNROLL
            for (int i = 0; i < N; ++i) {
                va[i] = vl[i] = 0.0;
                DCOL_UNROLL
                for (int j = 0; j < N; ++j) M[i][j] = 0.0;
            }
            pass_a<F1>(p1, c1, CE1, b1, x, M, va, vl);
            pass_a<F2>(p2, c2, CE2, b2, x, M, va, vl);
            if (int bad = chol(M, L, Li)) return bad;
        }
        double Q1[3][3], Q2[3][3], xw[N];
        P1 w1;
        P2 w2;
        world_frames(c1, c2, pose1, pose2, w1, w2, Q1, Q2, xw);
        const double (&Rf)[3][3] = kFrame2 ? w2.Qp : w1.Qp;
        double Jt[4][12];
        for (int k = 0; k < 4; ++k) { /* not unrolled: four passes over the same code */
            double yt[N], yw[N];
            DCOL_UNROLL
            for (int j = 0; j < N; ++j) yt[j] = (j == k) ? 1.0 : 0.0;
            chol_solve(L, Li, yt);
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) yw[i] = Rf[i][0] * yt[0] + Rf[i][1] * yt[1] + Rf[i][2] * yt[2];
            DCOL_UNROLL
            for (int j = 3; j < N; ++j) yw[j] = yt[j];
            double vo1[F1::NOA], vq1[F1::QA], vo2[F2::NOA], vq2[F2::QA];
            jac_v<F1>(p1, c1, CE1, b1, yt, vo1, vq1);
            jac_v<F2>(p2, c2, CE2, b2, yt, vo2, vq2);
            jac_block<P1>(w1, c1, CE1, b1, vo1, vq1, xw, yw, Rf, pose1 + 3, Q1, Jt[k]);
            jac_block<P2>(w2, c2, CE2, b2, vo2, vq2, xw, yw, Rf, pose2 + 3, Q2, Jt[k] + 6);
        }
        DCOL_UNROLL
        for (int j = 0; j < 12; ++j) {
            DCOL_UNROLL
            for (int i = 0; i < 3; ++i) jac[12 * i + j] = -(Rf[i][0] * Jt[0][j] + Rf[i][1] * Jt[1][j] + Rf[i][2] * Jt[2][j]);
            jac[36 + j] = -Jt[3][j];
        }
        return 0;
    }

    /* world-frame (x, s, z) in the reference's row order [ort1; ort2; soc1; soc2] (debug entry point) */
    template <class P>
    DCOL_HD static void export_soc(const P& pw, const double (&Rf)[3][3], const double (&q)[P::QA], double* out)
    {
        if (P::Q == 0) return;
        if (P::ball) { /* the reference keeps these cone components in world axes */
            out[0] = q[0];
            for (int i = 0; i < 3; ++i)
                out[1 + i] = P::rot ? (pw.Qp[i][0] * q[1] + pw.Qp[i][1] * q[2] + pw.Qp[i][2] * q[3])
                                    : (Rf[i][0] * q[1] + Rf[i][1] * q[2] + Rf[i][2] * q[3]);
        } else {
            for (int i = 0; i < P::Q; ++i) out[i] = q[i];
        }
    }
    DCOL_HD int export_xsz(const C1& c1, const C2& c2, const double* pose1, const double* pose2, double* xout, double* s,
                           double* z) const
    {
        double Q1[3][3], Q2[3][3], xw[N];
        P1 w1;
        P2 w2;
        world_frames(c1, c2, pose1, pose2, w1, w2, Q1, Q2, xw);
        for (int j = 0; j < N; ++j) xout[j] = xw[j];
        int r = 0;
        for (int i = 0; i < P1::n_ort(c1); ++i, ++r) { s[r] = b1.so[i]; z[r] = b1.zo[i]; }
        for (int i = 0; i < P2::n_ort(c2); ++i, ++r) { s[r] = b2.so[i]; z[r] = b2.zo[i]; }
        const double (&Rf)[3][3] = kFrame2 ? w2.Qp : w1.Qp;
        export_soc<P1>(w1, Rf, b1.sq, s + r); export_soc<P1>(w1, Rf, b1.zq, z + r); r += P1::Q;
        export_soc<P2>(w2, Rf, b2.sq, s + r); export_soc<P2>(w2, Rf, b2.zq, z + r); r += P2::Q;
        return r;
    }
};

} /* namespace dcol */
#endif /* DCOL_SOLVER_CUH_ */
