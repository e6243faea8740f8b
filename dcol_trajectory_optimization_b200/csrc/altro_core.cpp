/*
 * altro_core.cpp — native host core of the batched AL-iLQR caller: the C ABI of include/dcol_altro.h.
 *
 * The reference's optimiser (ALTRO.py) is a Python loop over knots calling NumPy on 12 x 12 matrices; per pass it
 * spends its non-collision time in the RK4 rollouts of the line search (ALTRO.py:183-239), the forward-difference
 * dynamics Jacobians (ALTRO.py:77-100) and the Riccati recursion (ALTRO.py:242-338).  Those are sequential over
 * knots, so they stay on the host, but as straight C++ on stack arrays: no interpreter, no allocation, no BLAS call
 * overhead on matrices this small.  Every collision constraint still comes from the CUDA engine (include/dcol.h).
 * Formulas and operation order follow the reference (and the NumPy twin in altro/solver.py, which is kept for
 * user-defined dynamics and as the cross-check of this file).  Compiled with -ffp-contract=off.
 */
#include <math.h>
#include <string.h>

#include "../../include/dcol_altro.h"

namespace {

constexpr int NX = DCOL_ALTRO_MAX_NX, NU = DCOL_ALTRO_MAX_NU;
typedef dcol_altro_problem Prob;

inline void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* p_dot = ((1 + |p|^2) / 4) (I + 2 ([p x]^2 + [p x]) / (1 + |p|^2)) omega, written with cross products
 * (cone_through_wall.py:41-44, cluttered_hallway_quadrotor.py:60-61) */
inline void mrp_rate(const double* p, const double* w, double* o)
{
    const double pp = (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2];
    double pw[3], ppw[3];
    cross3(p, w, pw);
    cross3(p, pw, ppw);
    for (int i = 0; i < 3; ++i) o[i] = 0.25 * (1.0 + pp) * w[i] + 0.5 * (ppw[i] + pw[i]);
}

/* Q(p) v, Q = I + (8 [p x]^2 + 4 (1 - |p|^2) [p x]) / (1 + |p|^2)^2   (problem_matrices.py:213-251) */
inline void rotate_by_mrp(const double* p, const double* v, double* o)
{
    const double pp = (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2];
    double pv[3], ppv[3];
    cross3(p, v, pv);
    cross3(p, pv, ppv);
    const double den = (1.0 + pp) * (1.0 + pp);
    for (int i = 0; i < 3; ++i) o[i] = v[i] + (8.0 * ppv[i] + 4.0 * (1.0 - pp) * pv[i]) / den;
}

inline void euler_rates(const Prob& P, const double* w, const double* tau, double* wdot)
{
    const double Jw[3] = { P.inertia[0] * w[0], P.inertia[1] * w[1], P.inertia[2] * w[2] };
    double c[3];
    cross3(w, Jw, c);
    for (int i = 0; i < 3; ++i) wdot[i] = (tau[i] - c[i]) / P.inertia[i];
}

void dynamics(const Prob& P, const double* x, const double* u, double* xd)
{
    switch (P.system) {
    case DCOL_ALTRO_PIANO: /* piano_mover.py:5-23 */
        xd[0] = x[2];
        xd[1] = x[3];
        xd[2] = u[0];
        xd[3] = u[1];
        xd[4] = x[5];
        xd[5] = u[2] / 100.0;
        break;
    case DCOL_ALTRO_RIGID_BODY: /* cone_through_wall.py:19-47 */
        for (int i = 0; i < 3; ++i) {
            xd[i] = x[3 + i];
            xd[3 + i] = u[i] / P.mass;
        }
        mrp_rate(x + 6, x + 9, xd + 6);
        euler_rates(P, x + 9, u + 3, xd + 9);
        break;
    default: { /* cluttered_hallway_quadrotor.py:17-74 */
        double F[4], M[4];
        for (int i = 0; i < 4; ++i) {
            const double f = P.kf * u[i];
            F[i] = f > 0.0 ? f : 0.0;
            M[i] = P.km * u[i];
        }
        const double thrust[3] = { 0.0, 0.0, ((F[0] + F[1]) + F[2]) + F[3] };
        const double tau[3] = { P.arm * (F[1] - F[3]), P.arm * (F[2] - F[0]), ((M[0] - M[1]) + M[2]) - M[3] };
        double ft[3];
        rotate_by_mrp(x + 6, thrust, ft);
        const double grav[3] = { 0.0, 0.0, -9.81 };
        for (int i = 0; i < 3; ++i) {
            xd[i] = x[3 + i];
            xd[3 + i] = (P.mass * grav[i] + ft[i]) / P.mass;
        }
        mrp_rate(x + 6, x + 9, xd + 6);
        euler_rates(P, x + 9, tau, xd + 9);
    }
    }
}

/* discrete_dynamics of the system scripts (e.g. piano_mover.py:25-43) */
void rk4(const Prob& P, const double* x, const double* u, double* xn)
{
    const int nx = P.nx;
    double k1[NX], k2[NX], k3[NX], k4[NX], t[NX] = {};
    dynamics(P, x, u, k1);
    for (int i = 0; i < nx; ++i) { k1[i] *= P.dt; t[i] = x[i] + 0.5 * k1[i]; }
    dynamics(P, t, u, k2);
    for (int i = 0; i < nx; ++i) { k2[i] *= P.dt; t[i] = x[i] + 0.5 * k2[i]; }
    dynamics(P, t, u, k3);
    for (int i = 0; i < nx; ++i) { k3[i] *= P.dt; t[i] = x[i] + k3[i]; }
    dynamics(P, t, u, k4);
    for (int i = 0; i < nx; ++i) {
        k4[i] *= P.dt;
        xn[i] = x[i] + (1.0 / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
    }
}

bool bad_problem(const Prob* p)
{
    if (!p || p->nx < 1 || p->nx > NX || p->nu < 1 || p->nu > NU || p->N < 2 || p->n_obs < 0) return true;
    if (p->system == DCOL_ALTRO_PIANO) return p->nx != 6 || p->nu != 3;
    if (p->system == DCOL_ALTRO_RIGID_BODY) return p->nx != 12 || p->nu != 6;
    if (p->system == DCOL_ALTRO_QUADROTOR) return p->nx != 12 || p->nu != 4;
    return true;
}

/* lower Cholesky of the n x n SPD matrix M (row-major, leading dimension NU); false if a pivot is not positive */
bool chol_small(const double (*M)[NU], int n, double (*L)[NU])
{
    for (int j = 0; j < n; ++j) {
        double d = M[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) return false;
        L[j][j] = sqrt(d);
        for (int i = j + 1; i < n; ++i) {
            double v = M[i][j];
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v / L[j][j];
        }
    }
    return true;
}
void chol_solve_small(const double (*L)[NU], int n, double* v)
{
    for (int i = 0; i < n; ++i) {
        double t = v[i];
        for (int k = 0; k < i; ++k) t -= L[i][k] * v[k];
        v[i] = t / L[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double t = v[i];
        for (int k = i + 1; k < n; ++k) t -= L[k][i] * v[k];
        v[i] = t / L[i][i];
    }
}

/* mult . h + rho/2 h' I_mask h, mask = (mult > 0 or h > 0)   (eval_mask ALTRO.py:16-30, :127-133) */
inline double al_terms(const double* h, const double* mult, int n, double rho)
{
    double lin = 0.0, quad = 0.0;
    for (int i = 0; i < n; ++i) {
        lin += mult[i] * h[i];
        if (mult[i] > 0.0 || h[i] > 0.0) quad += h[i] * h[i];
    }
    return lin + 0.5 * rho * quad;
}

inline double quad_form(const double* M, const double* d, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double r = 0.0;
        for (int j = 0; j < n; ++j) r += M[i * n + j] * d[j];
        s += d[i] * r;
    }
    return s;
}

} /* namespace */

extern "C" {

const char* dcol_altro_version(void) { return "dcol-altro-core 0.1"; }

int dcol_altro_dynamics(const Prob* p, int64_t n, const double* X, const double* U, double* Xdot)
{
    if (bad_problem(p) || n < 0 || !X || !U || !Xdot) return DCOL_ALTRO_E_ARG;
    for (int64_t i = 0; i < n; ++i) dynamics(*p, X + i * p->nx, U + i * p->nu, Xdot + i * p->nx);
    return 0;
}

int dcol_altro_rk4(const Prob* p, int64_t n, const double* X, const double* U, double* Xnext)
{
    if (bad_problem(p) || n < 0 || !X || !U || !Xnext) return DCOL_ALTRO_E_ARG;
    for (int64_t i = 0; i < n; ++i) rk4(*p, X + i * p->nx, U + i * p->nu, Xnext + i * p->nx);
    return 0;
}

int dcol_altro_rollouts(const Prob* p, const double* X, const double* U, const double* K, const double* k,
                        const double* alphas, int32_t C, double* Xn, double* Un)
{
    if (bad_problem(p) || C < 0 || !X || !U || !K || !k || !alphas || !Xn || !Un) return DCOL_ALTRO_E_ARG;
    const int nx = p->nx, nu = p->nu, N = p->N;
    for (int c = 0; c < C; ++c) {
        double* xc = Xn + (int64_t)c * N * nx;
        double* uc = Un + (int64_t)c * (N - 1) * nu;
        memcpy(xc, X, sizeof(double) * nx);
        for (int t = 0; t < N - 1; ++t) {
            const double* Kt = K + (int64_t)t * nu * nx;
            double dx[NX];
            for (int j = 0; j < nx; ++j) dx[j] = xc[t * nx + j] - X[t * nx + j];
            for (int i = 0; i < nu; ++i) { /* ALTRO.py:219-220 */
                double fb = 0.0;
                for (int j = 0; j < nx; ++j) fb += Kt[i * nx + j] * dx[j];
                uc[t * nu + i] = (U[t * nu + i] - fb) - alphas[c] * k[t * nu + i];
            }
            rk4(*p, xc + t * nx, uc + t * nu, xc + (t + 1) * nx);
        }
    }
    return 0;
}

int dcol_altro_jacobians(const Prob* p, const double* X, const double* U, double delta, double* A, double* B)
{
    if (bad_problem(p) || !X || !U || !A || !B || !(delta > 0.0)) return DCOL_ALTRO_E_ARG;
    const int nx = p->nx, nu = p->nu, N = p->N;
    for (int t = 0; t < N - 1; ++t) { /* compute_jacobian, ALTRO.py:77-100: forward differences */
        double base[NX], xp[NX], up[NU], f[NX];
        rk4(*p, X + t * nx, U + t * nu, base);
        for (int i = 0; i < nx; ++i) {
            memcpy(xp, X + t * nx, sizeof(double) * nx);
            xp[i] += delta;
            rk4(*p, xp, U + t * nu, f);
            for (int r = 0; r < nx; ++r) A[((int64_t)t * nx + r) * nx + i] = (f[r] - base[r]) / delta;
        }
        for (int i = 0; i < nu; ++i) {
            memcpy(up, U + t * nu, sizeof(double) * nu);
            up[i] += delta;
            rk4(*p, X + t * nx, up, f);
            for (int r = 0; r < nx; ++r) B[((int64_t)t * nx + r) * nu + i] = (f[r] - base[r]) / delta;
        }
    }
    return 0;
}

int dcol_altro_backward_pass(const Prob* p, const double* X, const double* U, const double* hx, const double* ghx,
                             const double* mu, const double* mux, const double* lambd, double rho, double reg,
                             double* K, double* k, double* delta_J)
{
    if (bad_problem(p) || !X || !U || (p->n_obs > 0 && (!hx || !ghx || !mux)) || !mu || !lambd || !K || !k || !delta_J)
        return DCOL_ALTRO_E_ARG;
    const int nx = p->nx, nu = p->nu, N = p->N, no = p->n_obs;
    double Vx[NX], Vxx[NX][NX];
    double dJ = 0.0;
    /* cost-to-go of one knot's collision constraints: g += ghx^T (mux + rho mask hx), H += rho ghx^T mask ghx */
    auto add_constraints = [&](int t, double* g, double (*H)[NX]) {
        for (int o = 0; o < no; ++o) {
            const double h = hx[t * no + o], m = mux[t * no + o];
            const bool on = m > 0.0 || h > 0.0;
            const double* gr = ghx + ((int64_t)t * no + o) * nx;
            const double cf = m + rho * (on ? h : 0.0);
            for (int i = 0; i < nx; ++i) g[i] += gr[i] * cf;
            if (on)
                for (int i = 0; i < nx; ++i)
                    for (int j = 0; j < nx; ++j) H[i][j] += rho * gr[i] * gr[j];
        }
    };
    { /* terminal knot, ALTRO.py:265-282 */
        const double* xT = X + (N - 1) * nx;
        const double* xr = p->Xref + (N - 1) * nx;
        double d[NX];
        for (int i = 0; i < nx; ++i) d[i] = xT[i] - xr[i];
        for (int i = 0; i < nx; ++i) {
            double s = 0.0;
            for (int j = 0; j < nx; ++j) {
                s += p->Qf[i * nx + j] * d[j];
                Vxx[i][j] = p->Qf[i * nx + j];
            }
            Vx[i] = s;
        }
        add_constraints(N - 1, Vx, Vxx);
        for (int i = 0; i < nx; ++i) {
            Vx[i] += lambd[i] + rho * d[i]; /* goal constraint */
            Vxx[i][i] += rho;
        }
    }
    /* finite-difference Jacobians are produced knot by knot inside the sweep (same values as dcol_altro_jacobians) */
    for (int t = N - 2; t >= 0; --t) {
        const double* xt = X + t * nx;
        const double* ut = U + t * nu;
        double At[NX][NX], Bt[NX][NU];
        {
            const double delta = 1e-6;
            double base[NX], xp[NX], up[NU], f[NX];
            rk4(*p, xt, ut, base);
            for (int i = 0; i < nx; ++i) {
                memcpy(xp, xt, sizeof(double) * nx);
                xp[i] += delta;
                rk4(*p, xp, ut, f);
                for (int r = 0; r < nx; ++r) At[r][i] = (f[r] - base[r]) / delta;
            }
            for (int i = 0; i < nu; ++i) {
                memcpy(up, ut, sizeof(double) * nu);
                up[i] += delta;
                rk4(*p, xt, up, f);
                for (int r = 0; r < nx; ++r) Bt[r][i] = (f[r] - base[r]) / delta;
            }
        }
        double lx[NX], lu[NU], lxx[NX][NX], luu[NU][NU];
        for (int i = 0; i < nx; ++i) {
            double s = 0.0;
            for (int j = 0; j < nx; ++j) {
                s += p->Q[i * nx + j] * (xt[j] - p->Xref[t * nx + j]);
                lxx[i][j] = p->Q[i * nx + j];
            }
            lx[i] = s;
        }
        for (int i = 0; i < nu; ++i) {
            double s = 0.0;
            for (int j = 0; j < nu; ++j) {
                s += p->R[i * nu + j] * (ut[j] - p->Uref[t * nu + j]);
                luu[i][j] = p->R[i * nu + j];
            }
            lu[i] = s;
        }
        /* control bounds hu = [u - u_max; -u + u_min], Iu = [I; -I]   ALTRO.py:293-300 */
        for (int i = 0; i < nu; ++i) {
            const double hp = ut[i] - p->u_max[i], hm = -ut[i] + p->u_min[i];
            const double mp = mu[t * 2 * nu + i], mm = mu[t * 2 * nu + nu + i];
            const bool onp = mp > 0.0 || hp > 0.0, onm = mm > 0.0 || hm > 0.0;
            lu[i] += (mp + rho * (onp ? hp : 0.0)) - (mm + rho * (onm ? hm : 0.0));
            luu[i][i] += rho * ((onp ? 1.0 : 0.0) + (onm ? 1.0 : 0.0));
        }
        add_constraints(t, lx, lxx);
        /* Q-function, ALTRO.py:314-330 */
        double Vreg[NX][NX], VB[NX][NU], VA[NX][NX];
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) Vreg[i][j] = Vxx[i][j] + (i == j ? reg : 0.0);
        for (int i = 0; i < nx; ++i) {
            for (int j = 0; j < nu; ++j) {
                double s = 0.0;
                for (int l = 0; l < nx; ++l) s += Vreg[i][l] * Bt[l][j];
                VB[i][j] = s;
            }
            for (int j = 0; j < nx; ++j) {
                double s = 0.0;
                for (int l = 0; l < nx; ++l) s += Vreg[i][l] * At[l][j];
                VA[i][j] = s;
            }
        }
        double Qu[NU], Quu[NU][NU], Qux[NU][NX];
        for (int i = 0; i < nu; ++i) {
            double s = 0.0;
            for (int l = 0; l < nx; ++l) s += Bt[l][i] * Vx[l];
            Qu[i] = lu[i] + s;
            for (int j = 0; j < nu; ++j) {
                double q = 0.0;
                for (int l = 0; l < nx; ++l) q += Bt[l][i] * VB[l][j];
                Quu[i][j] = luu[i][j] + q;
            }
            for (int j = 0; j < nx; ++j) {
                double q = 0.0;
                for (int l = 0; l < nx; ++l) q += Bt[l][i] * VA[l][j];
                Qux[i][j] = q;
            }
        }
        double Lc[NU][NU];
        if (!chol_small(Quu, nu, Lc)) return DCOL_ALTRO_E_NOT_PD;
        double kt[NU], Kt[NU][NX];
        for (int i = 0; i < nu; ++i) kt[i] = Qu[i];
        chol_solve_small(Lc, nu, kt);
        for (int j = 0; j < nx; ++j) {
            double col[NU];
            for (int i = 0; i < nu; ++i) col[i] = Qux[i][j];
            chol_solve_small(Lc, nu, col);
            for (int i = 0; i < nu; ++i) Kt[i][j] = col[i];
        }
        /* value function, ALTRO.py:332-336:
         *   Vx  <- lx - K' lu + K' luu k + (A - B K)' (Vx - Vxx B k)
         *   Vxx <- lxx + K' luu K + (A - B K)' Vxx (A - B K)                       (un-regularised Vxx) */
        double Acl[NX][NX], Bk[NX], w[NX], luuk[NU], luuK[NU][NX];
        for (int i = 0; i < nx; ++i) {
            double s = 0.0;
            for (int l = 0; l < nu; ++l) s += Bt[i][l] * kt[l];
            Bk[i] = s;
            for (int j = 0; j < nx; ++j) {
                double q = 0.0;
                for (int l = 0; l < nu; ++l) q += Bt[i][l] * Kt[l][j];
                Acl[i][j] = At[i][j] - q;
            }
        }
        for (int i = 0; i < nx; ++i) {
            double s = 0.0;
            for (int l = 0; l < nx; ++l) s += Vxx[i][l] * Bk[l];
            w[i] = Vx[i] - s;
        }
        for (int i = 0; i < nu; ++i) {
            double s = 0.0;
            for (int l = 0; l < nu; ++l) s += luu[i][l] * kt[l];
            luuk[i] = s;
            for (int j = 0; j < nx; ++j) {
                double q = 0.0;
                for (int l = 0; l < nu; ++l) q += luu[i][l] * Kt[l][j];
                luuK[i][j] = q;
            }
        }
        double Vx_new[NX], VAcl[NX][NX], Vxx_new[NX][NX];
        for (int i = 0; i < nx; ++i) {
            double s = lx[i];
            for (int l = 0; l < nu; ++l) s += Kt[l][i] * (luuk[l] - lu[l]);
            for (int l = 0; l < nx; ++l) s += Acl[l][i] * w[l];
            Vx_new[i] = s;
            for (int j = 0; j < nx; ++j) {
                double q = 0.0;
                for (int l = 0; l < nx; ++l) q += Vxx[i][l] * Acl[l][j];
                VAcl[i][j] = q;
            }
        }
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) {
                double q = lxx[i][j];
                for (int l = 0; l < nu; ++l) q += Kt[l][i] * luuK[l][j];
                for (int l = 0; l < nx; ++l) q += Acl[l][i] * VAcl[l][j];
                Vxx_new[i][j] = q;
            }
        for (int i = 0; i < nx; ++i) {
            Vx[i] = Vx_new[i];
            for (int j = 0; j < nx; ++j) Vxx[i][j] = Vxx_new[i][j];
        }
        double qk = 0.0;
        for (int i = 0; i < nu; ++i) {
            qk += Qu[i] * kt[i];
            k[t * nu + i] = kt[i];
            for (int j = 0; j < nx; ++j) K[((int64_t)t * nu + i) * nx + j] = Kt[i][j];
        }
        dJ += qk;
    }
    *delta_J = dJ;
    return 0;
}

int dcol_altro_total_cost(const Prob* p, int32_t C, const double* X, const double* U, const double* hx, const double* mu,
                          const double* mux, const double* lambd, double rho, double* cost)
{
    if (bad_problem(p) || C < 0 || !X || !U || (p->n_obs > 0 && (!hx || !mux)) || !mu || !lambd || !cost)
        return DCOL_ALTRO_E_ARG;
    const int nx = p->nx, nu = p->nu, N = p->N, no = p->n_obs;
    for (int c = 0; c < C; ++c) { /* compute_total_cost, ALTRO.py:103-143, knot by knot in the reference's order */
        const double* Xc = X + (int64_t)c * N * nx;
        const double* Uc = U + (int64_t)c * (N - 1) * nu;
        const double* hc = hx + (int64_t)c * N * no;
        double J = 0.0, d[NX], du[NU], hu[2 * NU];
        for (int t = 0; t < N - 1; ++t) {
            for (int i = 0; i < nx; ++i) d[i] = Xc[t * nx + i] - p->Xref[t * nx + i];
            for (int i = 0; i < nu; ++i) {
                du[i] = Uc[t * nu + i] - p->Uref[t * nu + i];
                hu[i] = Uc[t * nu + i] - p->u_max[i];
                hu[nu + i] = -Uc[t * nu + i] + p->u_min[i];
            }
            J += 0.5 * quad_form(p->Q, d, nx) + 0.5 * quad_form(p->R, du, nu);
            J += al_terms(hu, mu + t * 2 * nu, 2 * nu, rho);
            J += al_terms(hc + t * no, mux + t * no, no, rho);
        }
        for (int i = 0; i < nx; ++i) d[i] = Xc[(N - 1) * nx + i] - p->Xref[(N - 1) * nx + i];
        J += 0.5 * quad_form(p->Qf, d, nx);
        J += al_terms(hc + (N - 1) * no, mux + (N - 1) * no, no, rho);
        double lg = 0.0, gg = 0.0;
        for (int i = 0; i < nx; ++i) {
            lg += lambd[i] * d[i];
            gg += d[i] * d[i];
        }
        cost[c] = J + lg + 0.5 * rho * gg;
    }
    return 0;
}

} /* extern "C" */
