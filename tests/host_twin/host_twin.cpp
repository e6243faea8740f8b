/*
 * host_twin.cpp — TEST INFRASTRUCTURE.  Compiles the product's device solver header
 * (dcol_trajectory_optimization_b200/csrc/dcol_solver.cuh) as plain C++ so that the algorithm the
 * CUDA threads run can be checked against the oracle in the CPU-only container.  The product never
 * loads this library: the shipped path is the CUDA extension and fails loudly without a GPU.
 */
#include <pthread.h>
#include <string.h>

#include "../../dcol_trajectory_optimization_b200/csrc/dcol_classes.cuh"

using namespace dcol;

namespace {

struct PairIn {
    const dcol_shape *s1, *s2;
    const double *A, *b, *pose1, *pose2;
    double tol;
    int max_iter, want_grad;
};
struct PairOut {
    double alpha, x[8], grad[12], s[2 * DCOL_MAX_FACES + 8], z[2 * DCOL_MAX_FACES + 8], mu[DCOL_MAX_ITER + 1];
    double jac[48];
    int n, m, iters, status, want_jac;
};

struct RunOne {
    const PairIn* in;
    PairOut* out;
    bool trace;
    template <int C1, int C2>
    void operator()()
    {
        typedef typename ClassPrim<C1>::type P1;
        typedef typename ClassPrim<C2>::type P2;
        typedef Solver<P1, P2> S;
        typename P1::Const c1;
        typename P2::Const c2;
        fill_const_prim<P1>(*in->s1, in->A, in->b, c1);
        fill_const_prim<P2>(*in->s2, in->A, in->b, c2);
        S* sv = new S();
        PairResult<S::N> res;
        Trace tr = { out->mu };
        int st = sv->solve(c1, c2, in->pose1, in->pose2, in->tol, in->max_iter, in->want_grad != 0, res,
                           trace ? &tr : nullptr);
        out->status = st;
        out->iters = res.iters;
        out->n = S::N;
        out->alpha = st == 0 ? sv->x[3] : NAN;
        for (int j = 0; j < 8; ++j) out->x[j] = NAN;
        for (int j = 0; j < 12; ++j) out->grad[j] = NAN;
        if (st == 0 && in->want_grad) sv->gradient(c1, c2, in->pose1, in->pose2, out->grad);
        out->m = st == 0 ? sv->export_xsz(c1, c2, in->pose1, in->pose2, out->x, out->s, out->z) : 0;
        for (int j = 0; j < 48; ++j) out->jac[j] = NAN;
        if (st == 0 && out->want_jac) { /* after the export: jacobian() reuses the blocks' temporaries */
            double J[48];
            if (sv->jacobian(c1, c2, in->pose1, in->pose2, J) == 0) memcpy(out->jac, J, sizeof(J));
        }
        delete sv;
    }
};

int g_fix_case4 = 0; /* mirrors the DCOL_FIX_CASE4 flag of the C ABI */

void run_pair(const PairIn& in, PairOut& out, bool trace)
{
    for (int i = 0; i <= DCOL_MAX_ITER; ++i) out.mu[i] = NAN;
    int c1 = shape_class(*in.s1, in.A), c2 = shape_class(*in.s2, in.A);
    RunOne f = { &in, &out, trace };
    const bool allowed = c1 >= 0 && c2 >= 0 && (g_fix_case4 || class_pair_supported(c1, c2));
    if (!allowed || !dispatch_classes(c1, c2, f)) {
        out.status = DCOL_STATUS_UNSUPPORTED;
        out.iters = 0;
        out.alpha = NAN;
        out.n = out.m = 0;
        for (int j = 0; j < 8; ++j) out.x[j] = NAN;
        for (int j = 0; j < 12; ++j) out.grad[j] = NAN;
        for (int j = 0; j < 48; ++j) out.jac[j] = NAN;
    }
}

struct Job {
    const dcol_shape* shapes;
    const double *A, *b;
    const int32_t *idx1, *idx2;
    const double *pose1, *pose2;
    int64_t B;
    double tol;
    int max_iter, want_grad, tid, nthreads;
    double *alpha, *contact, *grad;
    int32_t *iters, *status;
    double* jac;
};

void* worker(void* arg)
{
    const Job* J = (const Job*)arg;
    PairOut* out = new PairOut();
    out->want_jac = J->jac != 0;
    for (int64_t c0 = (int64_t)J->tid * 64; c0 < J->B; c0 += (int64_t)J->nthreads * 64) {
        int64_t c1 = c0 + 64 < J->B ? c0 + 64 : J->B;
        for (int64_t k = c0; k < c1; ++k) {
            PairIn in = { J->shapes + J->idx1[k], J->shapes + J->idx2[k], J->A, J->b, J->pose1 + 6 * k,
                          J->pose2 + 6 * k, J->tol, J->max_iter, J->want_grad };
            run_pair(in, *out, false);
            J->alpha[k] = out->alpha;
            J->iters[k] = out->iters;
            J->status[k] = out->status;
            if (J->contact) for (int j = 0; j < 3; ++j) J->contact[3 * k + j] = out->x[j];
            if (J->grad) for (int j = 0; j < 12; ++j) J->grad[12 * k + j] = out->grad[j];
            if (J->jac) for (int j = 0; j < 48; ++j) J->jac[48 * k + j] = out->jac[j];
        }
    }
    delete out;
    return 0;
}

} /* namespace */

extern "C" void dcol_twin_set_fix_case4(int on) { g_fix_case4 = on; }

static int twin_batch(const dcol_shape* shapes, const double* A, const double* b, const int32_t* idx1,
                      const int32_t* idx2, const double* pose1, const double* pose2, int64_t B, double tol,
                      int max_iter, int threads, double* alpha, double* contact, double* grad,
                      int32_t* iters, int32_t* status, double* jac)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    Job jobs[256];
    for (int t = 0; t < threads; ++t) {
        Job j = { shapes, A, b, idx1, idx2, pose1, pose2, B, tol, max_iter, grad != 0, t, threads,
                  alpha, contact, grad, iters, status, jac };
        jobs[t] = j;
    }
    if (threads == 1) { worker(&jobs[0]); return 0; }
    for (int t = 0; t < threads; ++t) pthread_create(&tid[t], 0, worker, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], 0);
    return 0;
}

extern "C" int dcol_twin_batch(const dcol_shape* shapes, const double* A, const double* b, const int32_t* idx1,
                               const int32_t* idx2, const double* pose1, const double* pose2, int64_t B, double tol,
                               int max_iter, int threads, double* alpha, double* contact, double* grad,
                               int32_t* iters, int32_t* status)
{
    return twin_batch(shapes, A, b, idx1, idx2, pose1, pose2, B, tol, max_iter, threads, alpha, contact, grad, iters,
                      status, 0);
}
/* ... plus the solution Jacobian jac[B][4][12] (dcol_proximity_batch_jacobian) */
extern "C" int dcol_twin_batch_jac(const dcol_shape* shapes, const double* A, const double* b, const int32_t* idx1,
                                   const int32_t* idx2, const double* pose1, const double* pose2, int64_t B, double tol,
                                   int max_iter, int threads, double* alpha, double* contact, double* grad,
                                   int32_t* iters, int32_t* status, double* jac)
{
    return twin_batch(shapes, A, b, idx1, idx2, pose1, pose2, B, tol, max_iter, threads, alpha, contact, grad, iters,
                      status, jac);
}

/* one pair with the mu trace and the world-frame (x, s, z) */
extern "C" int dcol_twin_pair(const dcol_shape* shapes, const double* A, const double* b, int32_t i1, int32_t i2,
                              const double* pose1, const double* pose2, double tol, double* alpha, double* x,
                              double* s, double* z, int32_t* n, int32_t* m, int32_t* iters, double* grad,
                              double* mu_trace)
{
    PairOut* out = new PairOut();
    out->want_jac = 0;
    PairIn in = { shapes + i1, shapes + i2, A, b, pose1, pose2, tol, DCOL_MAX_ITER, 1 };
    run_pair(in, *out, true);
    *alpha = out->alpha;
    *n = out->n;
    *m = out->m;
    *iters = out->iters;
    memcpy(x, out->x, sizeof(double) * 8);
    memcpy(s, out->s, sizeof(double) * out->m);
    memcpy(z, out->z, sizeof(double) * out->m);
    memcpy(grad, out->grad, sizeof(double) * 12);
    memcpy(mu_trace, out->mu, sizeof(out->mu));
    int st = out->status;
    delete out;
    return st;
}
