/*
 * dcol_altro.h — C ABI of the native host core of the batched AL-iLQR caller (SURVEY.md section 8f, rows
 * N1/N2/N4: "dynamics Jacobians + Riccati").  Host code only (no CUDA): the per-pass work of the reference's
 * optimiser that is sequential over knots and therefore stays on the CPU —
 *
 *   dcol_altro_rollouts        forward_pass rollouts, ALTRO.py:183-239 (all line-search step sizes at once),
 *                              with discrete_dynamics (RK4) of the three system scripts:
 *                                systems/piano_mover.py:5-43, systems/cone_through_wall.py:19-65,
 *                                systems/cluttered_hallway_quadrotor.py:17-92
 *   dcol_altro_backward_pass   compute_jacobian ALTRO.py:77-100 (forward differences, delta = 1e-6) +
 *                              backward_pass ALTRO.py:242-338 (Riccati recursion with augmented-Lagrangian terms)
 *   dcol_altro_total_cost      compute_total_cost ALTRO.py:103-143 for a stack of trajectories
 *
 * — while every collision constraint h(x) = 1 - alpha and its gradient comes from the batched CUDA proximity
 * engine (include/dcol.h).  Same formulas, tolerances and update rules as the reference; the Python caller
 * (dcol_trajectory_optimization_b200/altro/solver.py) keeps a NumPy implementation of the same three functions
 * for user-defined dynamics and as the cross-check of this library.
 *
 * All arrays are row-major float64 host buffers.  Functions return 0, DCOL_ALTRO_E_ARG, or
 * DCOL_ALTRO_E_NOT_PD (Quu not positive definite: scipy's cho_factor raises LinAlgError there, ALTRO.py:321).
 */
#ifndef DCOL_ALTRO_H_
#define DCOL_ALTRO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DCOL_ALTRO_PIANO = 0,     /* planar double integrator with heading, nx 6, nu 3     piano_mover.py:5-23              */
    DCOL_ALTRO_RIGID_BODY = 1, /* 6-DOF rigid body, MRP attitude, nx 12, nu 6          cone_through_wall.py:19-47        */
    DCOL_ALTRO_QUADROTOR = 2  /* quadrotor, MRP attitude, nx 12, nu 4                  cluttered_hallway_quadrotor.py:17-74 */
};

enum { DCOL_ALTRO_E_ARG = -1, DCOL_ALTRO_E_NOT_PD = -3 };

#define DCOL_ALTRO_MAX_NX 12
#define DCOL_ALTRO_MAX_NU 6

typedef struct dcol_altro_problem {
    int32_t system;            /* DCOL_ALTRO_*                                             */
    int32_t nx, nu, N, n_obs;  /* states, controls, knots, collision constraints per knot  */
    int32_t reserved;
    double dt;
    double mass, inertia[3];   /* rigid body / quadrotor                                   */
    double arm, kf, km;        /* quadrotor: arm length, thrust and moment coefficients    */
    const double* Q;           /* [nx][nx]                                                 */
    const double* R;           /* [nu][nu]                                                 */
    const double* Qf;          /* [nx][nx]                                                 */
    const double* Xref;        /* [N][nx]                                                  */
    const double* Uref;        /* [N-1][nu]                                                */
    const double* u_min;       /* [nu]                                                     */
    const double* u_max;       /* [nu]                                                     */
} dcol_altro_problem;

const char* dcol_altro_version(void);

/* xdot[n][nx] = f(x[n], u[n]) and xnext[n][nx] = RK4 step (the systems' dynamics / discrete_dynamics) */
int dcol_altro_dynamics(const dcol_altro_problem* p, int64_t n, const double* X, const double* U, double* Xdot);
int dcol_altro_rk4(const dcol_altro_problem* p, int64_t n, const double* X, const double* U, double* Xnext);

/* Closed-loop rollouts for C step sizes: Un[c][t] = U[t] - K[t] (Xn[c][t] - X[t]) - alphas[c] k[t],
 * Xn[c][t+1] = rk4(Xn[c][t], Un[c][t]), Xn[c][0] = X[0].     K [N-1][nu][nx], k [N-1][nu] */
int dcol_altro_rollouts(const dcol_altro_problem* p, const double* X, const double* U, const double* K, const double* k,
                        const double* alphas, int32_t C, double* Xn /* [C][N][nx] */, double* Un /* [C][N-1][nu] */);

/* Forward-difference dynamics Jacobians of every knot: A [N-1][nx][nx], B [N-1][nx][nu] */
int dcol_altro_jacobians(const dcol_altro_problem* p, const double* X, const double* U, double delta, double* A, double* B);

/* Riccati recursion.  hx [N][n_obs] constraint values, ghx [N][n_obs][nx] their state gradients, mu [N-1][2 nu],
 * mux [N][n_obs], lambd [nx] multipliers; out K [N-1][nu][nx], k [N-1][nu], delta_J. */
int dcol_altro_backward_pass(const dcol_altro_problem* p, const double* X, const double* U, const double* hx,
                             const double* ghx, const double* mu, const double* mux, const double* lambd, double rho,
                             double reg, double* K, double* k, double* delta_J);

/* Augmented-Lagrangian cost of C trajectories: X [C][N][nx], U [C][N-1][nu], hx [C][N][n_obs] -> cost [C] */
int dcol_altro_total_cost(const dcol_altro_problem* p, int32_t C, const double* X, const double* U, const double* hx,
                          const double* mu, const double* mux, const double* lambd, double rho, double* cost);

#ifdef __cplusplus
}
#endif
#endif /* DCOL_ALTRO_H_ */
