/*
 * dcol_kernels.cuh — the batched proximity kernel: one thread owns one pair and runs the whole
 * solve (dcol_solver.cuh) in registers.  One launch covers one GROUP of the plan: all pairs that
 * share the same two shape records, so every shape constant is a kernel parameter (constant bank,
 * warp-uniform) and a thread reads only its two poses (96 B) and writes its results (<= 136 B).
 *
 * Replaces the per-pair Python call chain proximity_gradient -> problem_matrices x2 ->
 * combine_problem_matrices -> solve_lp_pdip -> obj_val_grad
 * (proximity/proximity_gradient.py:91-138) evaluated in loops at
 * systems/cluttered_hallway_quadrotor.py:131-133,155 and friends.
 */
#ifndef DCOL_KERNELS_CUH_
#define DCOL_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdlib.h>

#include "dcol_classes.cuh"

namespace dcol {

#ifndef DCOL_THREADS
#define DCOL_THREADS 64
#endif
#ifndef DCOL_MAX_DEST
#define DCOL_MAX_DEST 8 /* destinations of a record-mode solve: the local buffer + up to 7 NVLink peers */
#endif
#ifndef DCOL_MIN_BLOCKS
#define DCOL_MIN_BLOCKS 4
#endif
constexpr int kThreads = DCOL_THREADS;      /* threads per CTA                                        */
constexpr int kMinBlocks = DCOL_MIN_BLOCKS; /* CTAs per SM the register allocation must leave room for */

/* everything of a launch that does not depend on the specialisation */
struct BatchArgs {
    const int32_t* perm; /* plan order -> pair index, or null for the identity */
    int64_t first, count;
    const double* pose1; /* [B][6] rows (r, p) */
    const double* pose2;
    double tol;
    int32_t max_iter;
    uint32_t flags;
    double* alpha;   /* [B]      */
    double* contact; /* [B][3]   */
    double* grad;    /* [B][12]  */
    int32_t* iters;  /* [B]      */
    int32_t* status; /* [B]      */
    struct TraceOut* trace; /* debug entry point only (count == 1): mu trace and world-frame (x, s, z) */
    /* record mode (n_dest > 0): instead of the separate arrays, every pair's 112-byte record
     * {alpha, grad[12], iters | status << 32} goes, in PLAN order, to dest[d] + 14 * (record_offset + first + t)
     * for each destination d; destinations may be peer-GPU memory (the all-gather is fused into the solve) */
    int32_t n_dest;
    int64_t record_offset;
    double* dest[DCOL_MAX_DEST];
    double* jac; /* [B][4][12] solution Jacobian (jacobian kernels only, otherwise null) */
    int32_t ahead; /* threads of one resident wave (SMs x CTAs/SM x kThreads): L2 prefetch distance, 0 = off */
};

/* one pair with the mu trace and the world-frame (x, s, z): the debug entry point */
struct TraceOut {
    double x[8], s[2 * DCOL_MAX_FACES + 8], z[2 * DCOL_MAX_FACES + 8], mu[DCOL_MAX_ITER + 1];
    int32_t n, m;
};

template <class P1, class P2>
struct GroupArgs {
    typename P1::Const c1;
    typename P2::Const c2;
    BatchArgs b;
};

constexpr int kRecordWords = 14;
constexpr int kStageBytes = (kThreads / 32) * 32 * kRecordWords * 8;

/* JAC: the specialisations that also write the solution Jacobian; separate kernels,
 * compiled in their own translation units (dcol_jac_*.cu), so the hot kernels keep their register allocation */
template <class P1, class P2, bool JAC = false>
__global__ void __launch_bounds__(kThreads, JAC ? 1 : kMinBlocks) pair_kernel(const __grid_constant__ GroupArgs<P1, P2> a)
{
    typedef Solver<P1, P2> S;
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    /* lanes of this warp that own a pair (a prefix of the warp); taken while all 32 lanes are still here */
    const unsigned lanes = __ballot_sync(0xffffffffu, t < a.b.count);
    if (t >= a.b.count) return;
    const int64_t k = a.b.perm ? (int64_t)a.b.perm[a.b.first + t] : a.b.first + t;
#ifndef DCOL_NO_PREFETCH
    /* A thread's only global reads are perm -> poses: two DEPENDENT DRAM latencies with nothing to overlap them
     * (ncu: 11 % of the warp samples of a launch sit on the first use of the poses).  So every thread pulls the
     * poses of the pair that the CTA taking over this slot one resident wave later will read into L2, and the perm
     * entries of the wave after that, in the shadow of its own loads; later waves then start from L2 hits. */
    int64_t ka = -1;
    if (a.b.ahead > 0 && t + a.b.ahead < a.b.count) {
        ka = a.b.perm ? (int64_t)a.b.perm[a.b.first + t + a.b.ahead] : a.b.first + t + a.b.ahead;
        if (a.b.perm && t + 2 * (int64_t)a.b.ahead < a.b.count)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b.perm + a.b.first + t + 2 * (int64_t)a.b.ahead));
    }
#endif

    double pose1[6], pose2[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        pose1[j] = __ldg(a.b.pose1 + 6 * k + j);
        pose2[j] = __ldg(a.b.pose2 + 6 * k + j);
    }
#ifndef DCOL_NO_PREFETCH
    if (ka >= 0) { /* a 48-byte row can straddle two 32-byte sectors of different lines: touch both ends */
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b.pose1 + 6 * ka));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b.pose1 + 6 * ka + 5));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b.pose2 + 6 * ka));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b.pose2 + 6 * ka + 5));
    }
#endif
    S sv;
    PairResult<S::N> res;
#ifdef DCOL_NO_RECORDS
    const bool records = false;
#else
    const bool records = a.b.n_dest > 0;
#endif
    const bool want_grad = records || (a.b.flags & DCOL_WANT_GRAD) != 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    Trace tr = { nullptr };
    if (a.b.trace) {
        tr.mu = a.b.trace->mu;
        for (int i = 0; i <= DCOL_MAX_ITER; ++i) tr.mu[i] = nan;
    }
    const int st = sv.solve(a.c1, a.c2, pose1, pose2, a.b.tol, a.b.max_iter, want_grad, res, &tr);
    const double alpha = st == DCOL_STATUS_OK ? sv.x[3] : nan; /* proximity.py:51 */
    if (a.b.flags & DCOL_WANT_CONTACT) {
        double cp[3] = { nan, nan, nan };
        if (st == DCOL_STATUS_OK) sv.contact_point(a.c1, a.c2, pose1, pose2, cp);
#pragma unroll
        for (int j = 0; j < 3; ++j) a.b.contact[3 * k + j] = cp[j]; /* proximity.py:52 */
    }
    double g[12];
    if (want_grad) {
        if (st == DCOL_STATUS_OK) {
            sv.gradient(a.c1, a.c2, pose1, pose2, g);
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) g[j] = nan;
        }
    }
    if (!records) {
        a.b.status[k] = st;
        a.b.iters[k] = res.iters;
        a.b.alpha[k] = alpha;
        if (want_grad) {
#pragma unroll
            for (int j = 0; j < 12; ++j) a.b.grad[12 * k + j] = g[j];
        }
    } else {
        /* Record mode.  A warp's records are contiguous in plan order: transpose them through shared memory
         * and write them with 16-byte stores that cover whole 128-byte lines, once per destination (local
         * memory or a peer GPU's over NVLink). */
        extern __shared__ __align__(16) double stage_all[]; /* kStageBytes, passed only by record-mode launches */
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        double* stage_w = stage_all + w * (32 * kRecordWords);
        double* mine = stage_w + lane * kRecordWords;
        mine[0] = alpha;
#pragma unroll
        for (int j = 0; j < 12; ++j) mine[1 + j] = g[j];
        mine[13] = __longlong_as_double((long long)(uint32_t)res.iters | ((long long)st << 32));
        __syncwarp(lanes);
        const int n_lanes = __popc(lanes);
        const int n_chunks = n_lanes * (kRecordWords / 2);
        const double2* src = reinterpret_cast<const double2*>(stage_w);
        const int64_t t0 = t - lane; /* first pair of this warp */
        if (a.b.flags & DCOL_DEST_MULTICAST) {
            /* dest[0] is a multicast address: ONE multimem store per 16 bytes, replicated by the NVSwitch into the
             * same offset of every rank's buffer (NVLink SHARP) — the all-gather costs this GPU 112 B/pair of egress
             * instead of 112 B x (world - 1) */
            const float4* src4 = reinterpret_cast<const float4*>(stage_w);
            float4* dst = reinterpret_cast<float4*>(a.b.dest[0] + kRecordWords * (a.b.record_offset + a.b.first + t0));
            for (int c = lane; c < n_chunks; c += n_lanes) {
                const float4 v = src4[c];
                asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y),
                             "f"(v.z), "f"(v.w)
                             : "memory");
            }
            /* no fence here: the grid's completion makes the stores visible, and the caller's handshake orders
             * it before any consumer */
        } else {
            for (int d = 0; d < a.b.n_dest; ++d) {
                double2* dst = reinterpret_cast<double2*>(a.b.dest[d] + kRecordWords * (a.b.record_offset + a.b.first + t0));
                for (int c = lane; c < n_chunks; c += n_lanes) dst[c] = src[c];
            }
        }
    }
    if (JAC) {
        /* d(contact, alpha) / d[r1 p1 r2 p2] by an adjoint solve with the factor of the final reduced KKT matrix */
        double* jp = a.b.jac + 48 * k;
        if (st != DCOL_STATUS_OK || sv.jacobian(a.c1, a.c2, pose1, pose2, jp) != 0) {
            for (int j = 0; j < 48; ++j) jp[j] = nan;
        }
    }
    if (a.b.trace) {
        TraceOut* out = a.b.trace;
        out->n = S::N;
        for (int j = 0; j < 8; ++j) out->x[j] = nan;
        out->m = st == DCOL_STATUS_OK ? sv.export_xsz(a.c1, a.c2, pose1, pose2, out->x, out->s, out->z) : 0;
    }
}

/* host-side description of one launch */
struct GroupLaunch {
    const dcol_shape* s1; /* host copies of the two shape records */
    const dcol_shape* s2;
    const double* A;      /* host copies of the packed faces      */
    const double* b;
    BatchArgs args;
};

/* threads of one resident wave of pair_kernel CTAs on the current device (the kernels are register-limited to
 * kMinBlocks CTAs per SM) */
inline int resident_wave_threads()
{
    static const int off = getenv("DCOL_NO_PREFETCH") ? 1 : 0; /* A/B switch */
    if (off) return 0;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return 0;
    return sms * kMinBlocks * kThreads;
}

template <int C1, int C2, bool JAC = false>
cudaError_t launch_pair(const GroupLaunch& g, cudaStream_t stream)
{
    typedef typename ClassPrim<C1>::type P1;
    typedef typename ClassPrim<C2>::type P2;
    GroupArgs<P1, P2> a;
    fill_const(*g.s1, g.A, g.b, a.c1);
    fill_const(*g.s2, g.A, g.b, a.c2);
    a.b = g.args;
    if (g.args.count <= 0) return cudaSuccess;
    a.b.ahead = g.args.trace ? 0 : resident_wave_threads();
    const int64_t blocks = (g.args.count + kThreads - 1) / kThreads;
    pair_kernel<P1, P2, JAC><<<(unsigned)blocks, kThreads, g.args.n_dest > 0 ? kStageBytes : 0, stream>>>(a);
    return cudaGetLastError();
}

/* one translation unit per first class (dcol_inst_*.cu) defines these */
template <int C1>
cudaError_t launch_first_class(int c2, const GroupLaunch& g, cudaStream_t stream);
template <> cudaError_t launch_first_class<CLS_POLY6>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_POLY8>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_POLYN>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_CAPSULE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_CYLINDER>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_CONE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_SPHERE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_PGON5>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_PGONN>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_BOX>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class<CLS_ELLIPSOID>(int, const GroupLaunch&, cudaStream_t);
/* the same dispatch for the Jacobian kernels (dcol_jac_*.cu) */
template <int C1>
cudaError_t launch_first_class_jac(int c2, const GroupLaunch& g, cudaStream_t stream);
template <> cudaError_t launch_first_class_jac<CLS_POLY6>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_POLY8>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_POLYN>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_CAPSULE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_CYLINDER>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_CONE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_SPHERE>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_PGON5>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_PGONN>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_BOX>(int, const GroupLaunch&, cudaStream_t);
template <> cudaError_t launch_first_class_jac<CLS_ELLIPSOID>(int, const GroupLaunch&, cudaStream_t);

#define DCOL_DEFINE_FIRST_CLASS(C1) DCOL_DEFINE_FIRST_CLASS_(C1, launch_first_class, false)
#define DCOL_DEFINE_FIRST_CLASS_JAC(C1) DCOL_DEFINE_FIRST_CLASS_(C1, launch_first_class_jac, true)
#define DCOL_DEFINE_FIRST_CLASS_(C1, FN, JAC)                                                    \
    namespace dcol {                                                                             \
    namespace {                                                                                  \
    struct LaunchFn_##C1 {                                                                       \
        const GroupLaunch* g;                                                                    \
        cudaStream_t stream;                                                                     \
        cudaError_t err;                                                                         \
        template <int A1, int A2>                                                                \
        void operator()()                                                                        \
        {                                                                                        \
            err = launch_pair<A1, A2, JAC>(*g, stream);                                          \
        }                                                                                        \
    };                                                                                           \
    }                                                                                            \
    template <>                                                                                  \
    cudaError_t FN<C1>(int c2, const GroupLaunch& g, cudaStream_t stream)                        \
    {                                                                                            \
        LaunchFn_##C1 f = { &g, stream, cudaSuccess };                                           \
        if (!dispatch_class2<C1>(c2, f)) return cudaErrorInvalidValue;                           \
        return f.err;                                                                            \
    }                                                                                            \
    }

} /* namespace dcol */
#endif /* DCOL_KERNELS_CUH_ */
