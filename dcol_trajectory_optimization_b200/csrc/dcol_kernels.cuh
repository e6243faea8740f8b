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
    int32_t per_warp; /* lane-refill path: records owned by one warp of trip_kernel */
    double* state;    /* lane-refill path: this group's scratch records (blocks of 32 pairs x StateLayout::NW doubles) */
    int64_t state_stride; /* doubles per pair the scratch was sized for */
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
            sv.gradient(a.c1, a.c2, pose1, pose2, g, !(a.b.flags & DCOL_WANT_GRAD1));
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
            if (a.b.flags & DCOL_WANT_GRAD1) {
#pragma unroll
                for (int j = 0; j < 6; ++j) a.b.grad[6 * k + j] = g[j];
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) a.b.grad[12 * k + j] = g[j];
            }
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

/* ---------------------------------------------------------------------------------------------------------------
 * Lane-refill path: three kernels per group instead of one.
 *
 * pair_kernel above gives a thread ONE pair, so a warp runs until its slowest pair converges: per-pair iteration counts
 * have a long right tail (config 4: mean 8.1, mean of the per-warp maximum 11.2) and 22-26 of 32 lanes are active on
 * average (profiles/r01_ncu_final_8kernels.md).  Iteration counts cannot be predicted from the geometry (sorting by
 * mu_0, mu_1 or alpha moves lane efficiency from 0.72 to 0.74), so the lanes are kept busy dynamically:
 *
 *   init_kernel   one thread per pair: pose -> working frame, initial point, NT scaling and tests of iteration 0; the
 *                 pair's loop state goes to a record of NW doubles in a scratch buffer (plan order);
 *   trip_kernel   a warp owns a contiguous chunk of records; every lane holds one pair and all lanes run the same
 *                 iteration body (scaling + tests, one Newton step); a lane whose pair has finished writes its result
 *                 (x, z, iteration count, status) back into the pair's record and loads the next record of the chunk,
 *                 so the lanes of a warp are at different iterations of different pairs and stay busy until the chunk
 *                 runs out;
 *   finish_kernel one thread per pair: alpha, contact point, gradient and the stores (arrays or records) from the
 *                 record's final (x, z).
 *
 * Each kernel is as small as its job (the iteration body alone is ~20 KB of code: warps of an SM that are at different
 * points of it still share the instruction cache, which a single kernel holding initialisation, iteration and epilogue
 * for desynchronised warps does not: measured, profiles/r02_lane_refill_history.md).  The price is the record traffic,
 * ~1.3 KB per pair against 212 B of inputs and outputs, at 15 % of the HBM bandwidth.  Per-pair arithmetic is the same
 * sequence of operations as Solver::solve.  Fixed-size classes only; runtime-face-count classes, the debug trace and the
 * Jacobian kernels use pair_kernel. */

template <class S>
struct StateLayout {
    typedef typename S::F1 F1;
    typedef typename S::F2 F2;
    static constexpr int N = S::N;
    static constexpr int X = 0;
    static constexpr int SO1 = X + N;
    static constexpr int ZO1 = SO1 + F1::NO;
    static constexpr int RI1 = ZO1 + F1::NO;
    static constexpr int SQ1 = RI1 + F1::NO;
    static constexpr int ZQ1 = SQ1 + F1::Q;
    static constexpr int WH1 = ZQ1 + F1::Q;
    static constexpr int SC1 = WH1 + F1::Q; /* eta, ieta, bw */
    static constexpr int SO2 = SC1 + (F1::Q > 0 ? 3 : 0);
    static constexpr int ZO2 = SO2 + F2::NO;
    static constexpr int RI2 = ZO2 + F2::NO;
    static constexpr int SQ2 = RI2 + F2::NO;
    static constexpr int ZQ2 = SQ2 + F2::Q;
    static constexpr int WH2 = ZQ2 + F2::Q;
    static constexpr int SC2 = WH2 + F2::Q;
    static constexpr int POSE = SC2 + (F2::Q > 0 ? 3 : 0); /* Qp[9], rp[3] of the primitive that is not the frame */
    static constexpr int SZ = POSE + 12;
    static constexpr int META = SZ + 1; /* iteration count | status << 8 | finished << 16 */
    static constexpr int POSES = META + 1; /* the pair's two input poses (12 doubles): init_kernel gathers them through the
                                              plan's permutation once, finish_kernel reads them back coalesced */
    static constexpr int NW = POSES + 12;
};

/* f(offset, value&, part_of_the_finished_state) over every field of one block that the iteration carries or needs */
template <class P, class F>
__device__ __forceinline__ void visit_block(Block<P>& B, int so, int zo, int ri, int sq, int zq, int wh, int sc, F&& f)
{
#pragma unroll
    for (int i = 0; i < P::NO; ++i) {
        f(so + i, B.so[i], false);
        f(zo + i, B.zo[i], true);
        f(ri + i, B.rinv[i], false);
    }
    if (P::Q > 0) {
#pragma unroll
        for (int i = 0; i < P::Q; ++i) {
            f(sq + i, B.sq[i], false);
            f(zq + i, B.zq[i], true);
            f(wh + i, B.wh[i], false);
        }
        f(sc + 0, B.eta, false);
        f(sc + 1, B.ieta, false);
        f(sc + 2, B.bw, false);
    }
}
template <class S, class F>
__device__ __forceinline__ void visit_state(S& sv, double& sz, F&& f)
{
    typedef StateLayout<S> Lay;
#pragma unroll
    for (int j = 0; j < S::N; ++j) f(Lay::X + j, sv.x[j], true);
    visit_block(sv.b1, Lay::SO1, Lay::ZO1, Lay::RI1, Lay::SQ1, Lay::ZQ1, Lay::WH1, Lay::SC1, f);
    visit_block(sv.b2, Lay::SO2, Lay::ZO2, Lay::RI2, Lay::SQ2, Lay::ZQ2, Lay::WH2, Lay::SC2, f);
    if constexpr (S::kFrame2) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) f(Lay::POSE + 3 * i + j, sv.p1.Qp[i][j], false);
            f(Lay::POSE + 9 + i, sv.p1.rp[i], false);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) f(Lay::POSE + 3 * i + j, sv.p2.Qp[i][j], false);
            f(Lay::POSE + 9 + i, sv.p2.rp[i], false);
        }
    }
    f(Lay::SZ, sz, false);
}

/* Records are stored in blocks of 32 pairs, field-major inside a block: field f of the pair at position pos of its group
 * lives at (pos / 32 * NW + f) * 32 + pos % 32, so that the one-thread-per-pair kernels (init, finish) read and write
 * whole 256-byte rows per instruction; the refilling lanes of trip_kernel take consecutive positions and share sectors. */
template <int NW>
__device__ __forceinline__ int64_t record_base(int64_t pos)
{
    return (pos >> 5) * (int64_t)(32 * NW) + (pos & 31);
}

__device__ __forceinline__ double pack_meta(int iters, int status, bool finished)
{
    return __longlong_as_double((long long)(iters & 0xff) | ((long long)(status & 0xff) << 8) | ((long long)(finished ? 1 : 0) << 16));
}

constexpr int kServiceThreads = 128; /* init / finish kernels: plain one-thread-per-pair kernels */

template <class P1, class P2>
__global__ void __launch_bounds__(kServiceThreads) init_kernel(const __grid_constant__ GroupArgs<P1, P2> a)
{
    typedef Solver<P1, P2> S;
    typedef StateLayout<S> Lay;
    const int64_t t = (int64_t)blockIdx.x * kServiceThreads + threadIdx.x;
    if (t >= a.b.count) return;
    const int64_t k = a.b.perm ? (int64_t)a.b.perm[a.b.first + t] : a.b.first + t;
    double pose1[6], pose2[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        pose1[j] = __ldg(a.b.pose1 + 6 * k + j);
        pose2[j] = __ldg(a.b.pose2 + 6 * k + j);
    }
    S sv;
    double sz = 0.0;
    int32_t iters = 0;
    int st = sv.init_point(a.c1, a.c2, pose1, pose2);
    if (st == 0) st = sv.scale_and_check(a.c1, a.c2, a.b.tol, 0, iters, sz, nullptr);
    double* rec = a.b.state + record_base<Lay::NW>(t);
    visit_state(sv, sz, [&](int off, double& v, bool) { rec[32 * off] = v; });
    rec[32 * Lay::META] = pack_meta(iters, st == S::kContinue ? 0 : st, st != S::kContinue);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        rec[32 * (Lay::POSES + j)] = pose1[j];
        rec[32 * (Lay::POSES + 6 + j)] = pose2[j];
    }
}

template <class P1, class P2>
__global__ void __launch_bounds__(kThreads, kMinBlocks) trip_kernel(const __grid_constant__ GroupArgs<P1, P2> a)
{
    typedef Solver<P1, P2> S;
    typedef StateLayout<S> Lay;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (warp_id * (int64_t)a.b.per_warp >= a.b.count) return;
    int32_t next = (int32_t)(warp_id * a.b.per_warp); /* a group holds fewer than 2^31 pairs (dcol_plan_create) */
    const int32_t end = (int64_t)next + a.b.per_warp < a.b.count ? next + a.b.per_warp : (int32_t)a.b.count;
    const unsigned lt = (1u << lane) - 1u;
    double* const base = a.b.state; /* this group's records */

    S sv;
    double sz = 0.0;
    bool has = false, fin = false;
    int it = 0, status = 0;
    int32_t iters = 0, pos = 0;
    for (;;) {
        /* scaling and tests of the current iterate, for every lane that went through a Newton step (a record that was
         * just loaded brings the scaling of its iteration 0 with it) */
        if (has && !fin) {
            if (it >= a.b.max_iter) {
                fin = true;
                status = sv.cap_status(a.b.max_iter, iters);
            } else if (it > 0) {
                const int st = sv.scale_and_check(a.c1, a.c2, a.b.tol, it, iters, sz, nullptr);
                if (st != S::kContinue) {
                    fin = true;
                    status = st;
                }
            }
        }
        /* refill: a finished lane writes (x, z, iterations, status) into its pair's record; it and the lanes without a
         * pair take the next records of the chunk (those that are already final — the initialisation failed or
         * converged at the initial point — are skipped) */
        for (;;) {
            const unsigned m_need = __ballot_sync(0xffffffffu, !has || fin);
            if (m_need == 0) break;
            if (has && fin) {
                double* rec = base + record_base<Lay::NW>(pos);
                visit_state(sv, sz, [&](int off, double& v, bool part) {
                    if (part) rec[32 * off] = v;
                });
                rec[32 * Lay::META] = pack_meta(iters, status, true);
                has = false;
                fin = false;
            }
            if (next >= end) break;
            if (!has) {
                const int32_t mine = next + __popc(m_need & lt);
                if (mine < end) {
                    const double* rec = base + record_base<Lay::NW>(mine);
                    const long long meta = __double_as_longlong(rec[32 * Lay::META]);
                    if (!((meta >> 16) & 1)) { /* needs iterations */
                        visit_state(sv, sz, [&](int off, double& v, bool) { v = rec[32 * off]; });
                        pos = mine;
                        it = 0;
                        has = true;
                    }
                }
            }
            const int n_need = __popc(m_need);
            next = end - next < n_need ? end : next + n_need;
        }
        if (!__any_sync(0xffffffffu, has)) break;
        if (next + 4 < end) { /* pull the block of records the next refills read towards this SM while the Newton step runs:
                                 256 bytes per field, lane l fetches the fields l, l + 32, l + 64 */
            const double* blk = base + record_base<Lay::NW>((next + 4) & ~31);
#pragma unroll
            for (int f = 0; f < Lay::NW; f += 32)
                if (f + lane < Lay::NW) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(blk + 32 * (f + lane)));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(blk + 32 * (f + lane) + 16));
                }
        }
        /* one Newton step for every lane that holds an unfinished pair */
        if (has && !fin) {
            const int st = sv.newton_step(a.c1, a.c2, sz);
            if (st != 0) {
                fin = true;
                status = st;
                iters = it;
            }
            ++it;
        }
    }
}

template <class P1, class P2>
__global__ void __launch_bounds__(kServiceThreads, 4) finish_kernel(const __grid_constant__ GroupArgs<P1, P2> a)
{
    typedef Solver<P1, P2> S;
    typedef StateLayout<S> Lay;
    const int64_t t = (int64_t)blockIdx.x * kServiceThreads + threadIdx.x;
    const unsigned lanes = __ballot_sync(0xffffffffu, t < a.b.count);
    if (t >= a.b.count) return;
    const int64_t k = a.b.perm ? (int64_t)a.b.perm[a.b.first + t] : a.b.first + t;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const bool records = a.b.n_dest > 0;
    const bool want_grad = records || (a.b.flags & DCOL_WANT_GRAD) != 0;
    S sv;
    double sz_unused = 0.0;
    const double* rec = a.b.state + record_base<Lay::NW>(t);
    double pose1[6], pose2[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        pose1[j] = rec[32 * (Lay::POSES + j)];
        pose2[j] = rec[32 * (Lay::POSES + 6 + j)];
    }
    visit_state(sv, sz_unused, [&](int off, double& v, bool part) {
        if (part) v = rec[32 * off];
    });
    const long long meta = __double_as_longlong(rec[32 * Lay::META]);
    const int iters = (int)(meta & 0xff), st = (int)((meta >> 8) & 0xff);
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
            sv.gradient(a.c1, a.c2, pose1, pose2, g, records || !(a.b.flags & DCOL_WANT_GRAD1));
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) g[j] = nan;
        }
    }
    if (!records) {
        a.b.status[k] = st;
        a.b.iters[k] = iters;
        a.b.alpha[k] = alpha;
        if (want_grad) {
            if (a.b.flags & DCOL_WANT_GRAD1) {
#pragma unroll
                for (int j = 0; j < 6; ++j) a.b.grad[6 * k + j] = g[j];
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) a.b.grad[12 * k + j] = g[j];
            }
        }
        return;
    }
    /* record mode, as in pair_kernel: a warp's records are contiguous in plan order; transposed through shared memory
     * and written with 16-byte stores that cover whole 128-byte lines, once per destination */
    __shared__ __align__(16) double stage_all[(kServiceThreads / 32) * 32 * kRecordWords];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* stage_w = stage_all + w * (32 * kRecordWords);
    double* mine = stage_w + lane * kRecordWords;
    mine[0] = alpha;
#pragma unroll
    for (int j = 0; j < 12; ++j) mine[1 + j] = g[j];
    mine[13] = __longlong_as_double((long long)(uint32_t)iters | ((long long)st << 32));
    __syncwarp(lanes);
    const int n_lanes = __popc(lanes);
    const int n_chunks = n_lanes * (kRecordWords / 2);
    const int64_t t0 = t - lane;
    if (a.b.flags & DCOL_DEST_MULTICAST) {
        const float4* src4 = reinterpret_cast<const float4*>(stage_w);
        float4* dst = reinterpret_cast<float4*>(a.b.dest[0] + kRecordWords * (a.b.record_offset + a.b.first + t0));
        for (int c = lane; c < n_chunks; c += n_lanes) {
            const float4 v = src4[c];
            asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y), "f"(v.z),
                         "f"(v.w)
                         : "memory");
        }
    } else {
        const double2* src = reinterpret_cast<const double2*>(stage_w);
        for (int d = 0; d < a.b.n_dest; ++d) {
            double2* dst = reinterpret_cast<double2*>(a.b.dest[d] + kRecordWords * (a.b.record_offset + a.b.first + t0));
            for (int c = lane; c < n_chunks; c += n_lanes) dst[c] = src[c];
        }
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

/* doubles per scratch record of the lane-refill path for a pair of classes (0: the pair does not use the path) */
template <int C1, int C2>
constexpr int state_words_of()
{
    typedef typename ClassPrim<C1>::type P1;
    typedef typename ClassPrim<C2>::type P2;
    if constexpr (P1::dyn || P2::dyn) return 0;
    else return StateLayout<Solver<P1, P2>>::NW;
}
struct StateWordsFn {
    int words = 0;
    template <int A1, int A2>
    void operator()()
    {
        words = state_words_of<A1, A2>();
    }
};
inline int state_words(int c1, int c2)
{
    StateWordsFn f;
    return dispatch_classes(c1, c2, f) ? f.words : 0;
}

/* lane-refill kernels as the default path: environment DCOL_REFILL=0/1 (A/B switch), else DCOL_REFILL_DEFAULT */
#ifndef DCOL_REFILL_DEFAULT
#define DCOL_REFILL_DEFAULT 0
#endif
inline bool refill_enabled()
{
    static const int on = getenv("DCOL_REFILL") ? atoi(getenv("DCOL_REFILL")) : DCOL_REFILL_DEFAULT;
    return on != 0;
}
/* plan positions per warp: 32 x generations.  More generations amortise the fill and the ragged drain of a warp's
 * chunk; fewer give the grid more warps to balance.  Small groups degenerate to one generation (one pair per lane). */
inline int32_t refill_pairs_per_warp(int64_t count)
{
    static const int forced = getenv("DCOL_REFILL_GEN") ? atoi(getenv("DCOL_REFILL_GEN")) : 0;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = 148;
    }
    /* about 0.7 resident waves of trip_kernel warps per group launch (several groups run concurrently on side streams):
     * measured best on config 4 (8 generations at 209,715 pairs per group) */
    int64_t gen = forced > 0 ? forced : (int64_t)((double)count / (32.0 * sms * kMinBlocks * (kThreads / 32) * 0.7) + 0.5);
    gen = gen < 1 ? 1 : (gen > 16 ? 16 : gen);
    if (forced > 0) gen = forced;
    return (int32_t)(32 * gen);
}

template <int C1, int C2, bool JAC = false>
cudaError_t launch_pair(const GroupLaunch& g, cudaStream_t stream)
{
    typedef typename ClassPrim<C1>::type P1;
    typedef typename ClassPrim<C2>::type P2;
    GroupArgs<P1, P2> a;
    fill_const_prim<P1>(*g.s1, g.A, g.b, a.c1);
    fill_const_prim<P2>(*g.s2, g.A, g.b, a.c2);
    a.b = g.args;
    if (g.args.count <= 0) return cudaSuccess;
    if constexpr (!JAC && !P1::dyn && !P2::dyn) {
        if (!g.args.trace && g.args.state && !(g.args.flags & DCOL_ONE_PAIR_PER_THREAD) &&
            ((g.args.flags & DCOL_LANE_REFILL) || refill_enabled())) {
            typedef StateLayout<Solver<P1, P2>> Lay;
            if (Lay::NW > g.args.state_stride) return cudaErrorInvalidValue;
            a.b.per_warp = refill_pairs_per_warp(g.args.count);
            const int64_t warps = (g.args.count + a.b.per_warp - 1) / a.b.per_warp;
            const unsigned service_blocks = (unsigned)((g.args.count + kServiceThreads - 1) / kServiceThreads);
            init_kernel<P1, P2><<<service_blocks, kServiceThreads, 0, stream>>>(a);
            trip_kernel<P1, P2><<<(unsigned)((warps + kThreads / 32 - 1) / (kThreads / 32)), kThreads, 0, stream>>>(a);
            finish_kernel<P1, P2><<<service_blocks, kServiceThreads, 0, stream>>>(a);
            return cudaGetLastError();
        }
    }
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
