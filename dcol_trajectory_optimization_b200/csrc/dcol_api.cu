/*
 * dcol_api.cu — the C ABI of include/dcol.h: shape tables, plans (grouping a batch by shape pair
 * with a device counting sort), the batched solve on device or host buffers, the single-pair debug
 * trace and the FP64 peak probe.  No torch types, no CPU fallback: without a CUDA device every
 * compute entry point fails with DCOL_E_NOGPU.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "dcol_kernels.cuh"

using namespace dcol;

/* ------------------------------------------------------------------------------------------ */
namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what)
{
    g_last_error = what;
    return code;
}
int fail_cuda(cudaError_t e, const char* where)
{
    g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}
#define DCOL_CUDA(call)                                        \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);    \
    } while (0)

/* Makes `device` current for the scope of an API call and restores the caller's device afterwards
 * (a host framework such as PyTorch must not find its current device changed). */
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err != cudaSuccess) prev = -1;
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
/* inside the chunk loops of the host entry points: record the error and leave the loop, so that the common epilogue
 * still synchronises every stream (copies into the caller's buffers must not outlive the call) */
#define DCOL_CUDA_BREAK(call)                                  \
    {                                                          \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) {                               \
            rc = fail_cuda(e_, #call);                         \
            break;                                             \
        }                                                      \
    }
#define DCOL_DEVICE(device)                                                   \
    DeviceGuard guard_(device);                                               \
    if (guard_.err != cudaSuccess) return fail_cuda(guard_.err, "cudaSetDevice")

constexpr uint32_t kSolveFlags = DCOL_WANT_CONTACT | DCOL_WANT_GRAD | DCOL_FIX_CASE4 | DCOL_ONE_PAIR_PER_THREAD | DCOL_WANT_GRAD1 | DCOL_LANE_REFILL;

struct Group {
    int32_t i1, i2;
    int64_t first, count;
    bool supported;
};

/* A process-wide cache of the buffers that tables and plans allocate (device memory and the small mapped host buffer of a
 * plan).  cudaMalloc / cudaFree / cudaHostAlloc / cudaFreeHost take 0.1-1 ms each on a good day and tens to hundreds of
 * milliseconds on a bad one (measured: tools/diag_teardown.py), and a caller that builds an engine per solve (every
 * altro_solve) paid ~25 of them.  Freed buffers are parked by (device, kind, size class) and handed out again; at most
 * kPoolMaxBytes stay parked.  pool_release_all() returns everything to the driver. */
struct BufPool {
    struct Key {
        int device, kind; /* kind 0: device memory, 1: mapped page-locked host memory */
        size_t bytes;
        bool operator<(const Key& o) const
        {
            return device != o.device ? device < o.device : (kind != o.kind ? kind < o.kind : bytes < o.bytes);
        }
    };
    std::mutex mu;
    std::multimap<Key, void*> parked;
    std::map<void*, Key> live;
    size_t parked_bytes = 0;
};
constexpr size_t kPoolMaxBytes = (size_t)3 << 30;
static BufPool& pool()
{
    static BufPool* p = new BufPool(); /* never destroyed: no CUDA calls from a static destructor at process exit */
    return *p;
}
static size_t pool_class(size_t bytes)
{
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
}
static cudaError_t pool_alloc(int device, int kind, size_t bytes, void** out)
{
    BufPool& P = pool();
    const BufPool::Key key = { device, kind, pool_class(bytes ? bytes : 1) };
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.parked.find(key);
        if (it != P.parked.end()) {
            *out = it->second;
            P.parked.erase(it);
            P.parked_bytes -= key.bytes;
            P.live[*out] = key;
            return cudaSuccess;
        }
    }
    cudaError_t e = kind == 0 ? cudaMalloc(out, key.bytes) : cudaHostAlloc(out, key.bytes, cudaHostAllocMapped);
    if (e != cudaSuccess) {
        *out = nullptr;
        return e;
    }
    std::lock_guard<std::mutex> lock(P.mu);
    P.live[*out] = key;
    return cudaSuccess;
}
/* The caller guarantees that no work that touches the buffer is still in flight (the destroy paths synchronise the device
 * first: a parked buffer may be handed to another owner at once, where cudaFree would have waited). */
static void pool_free(void* p)
{
    if (!p) return;
    BufPool& P = pool();
    BufPool::Key key;
    bool park = false;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.live.find(p);
        if (it == P.live.end()) return; /* not ours */
        key = it->second;
        P.live.erase(it);
        if (P.parked_bytes + key.bytes <= kPoolMaxBytes) {
            P.parked.insert({ key, p });
            P.parked_bytes += key.bytes;
            park = true;
        }
    }
    if (!park) {
        if (key.kind == 0) cudaFree(p);
        else cudaFreeHost(p);
    }
}
template <class T>
static cudaError_t pool_alloc_t(int device, size_t count, T** out)
{
    void* p = nullptr;
    cudaError_t e = pool_alloc(device, 0, sizeof(T) * count, &p);
    *out = (T*)p;
    return e;
}

/* scratch of the host-buffer entry point, cached per table and grown on demand */
struct HostScratch {
    int64_t cap = 0;
    int32_t *idx1 = nullptr, *idx2 = nullptr, *iters = nullptr, *status = nullptr;
    double *pose1 = nullptr, *pose2 = nullptr, *alpha = nullptr, *contact = nullptr, *grad = nullptr;
    void release() /* back to the pool: nothing may still be using the buffers */
    {
        pool_free(idx1); pool_free(idx2); pool_free(iters); pool_free(status);
        pool_free(pose1); pool_free(pose2); pool_free(alpha); pool_free(contact); pool_free(grad);
        *this = HostScratch();
    }
};

} /* namespace */

struct dcol_plan;
extern "C" void dcol_plan_destroy(dcol_plan* P);

struct dcol_shape_table {
    int device;
    std::vector<dcol_shape> shapes;
    std::vector<int> cls;
    std::vector<double> A, b;
    /* group launches of one solve fan out over side streams (fork/join on the caller's stream), so the
     * tail of one group's grid overlaps the head of the next and small groups run concurrently */
    static constexpr int kSide = 32;
    std::mutex launch_mu;
    cudaStream_t side[kSide] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[kSide] = {};
    /* host entry point state */
    std::mutex mu;
    /* chunk slots of the host pipelines: the pair-list entry point rotates through up to kSlots of them, so that the
     * copy-in and the counting sort of later chunks (and the host thread that waits for their histograms) run ahead of
     * the solve instead of being gated by the chunk before last; the scene entry point uses the first two */
    static constexpr int kSlots = 4;
    HostScratch scratch[kSlots];
    dcol_plan* plans[kSlots] = {};
    cudaStream_t streams[4] = { nullptr, nullptr, nullptr, nullptr }; /* h2d, plan, solve, d2h */
    /* pair-list entry point: one solve stream per slot.  A solve forks from / joins into its own stream, so consecutive
     * chunks have no barrier between them: each side stream starts chunk i+1's kernels right behind its kernels of chunk
     * i (a chunk's 40 small grids leave the SMs half empty at every join otherwise: tools/diag_small_solve.py) */
    cudaStream_t run[kSlots] = {};
    cudaEvent_t ev_in[kSlots] = {}, ev_plan[kSlots] = {}, ev_done[kSlots] = {}, ev_out[kSlots] = {};
    /* scene entry point: victim poses of a chunk (two slots), the obstacles of the call */
    double* scene_vic[2] = { nullptr, nullptr };
    int64_t scene_vic_cap = 0;
    double* scene_obs_pose = nullptr;
    int32_t* scene_obs_shape = nullptr;
    int64_t scene_obs_cap = 0;
    /* what plans[0] / plans[1] currently describe when a scene call built them: (victim, poses per chunk, obstacle shapes);
     * a caller that evaluates the same scene again (every AL-iLQR pass) then skips the counting sort and its host wait */
    std::vector<int32_t> scene_key[2];
};

struct dcol_plan {
    int64_t capacity_or_size() const { return capacity > B ? capacity : B; }
    const dcol_shape_table* table; /* borrowed: the table must outlive every solve / refine of this plan ...   */
    int device;                    /* ... but not its destruction: the plan remembers its device              */
    int64_t B, capacity;
    int32_t* d_perm;   /* [capacity] plan order -> pair index             */
    int32_t* d_counts; /* [n_shapes^2 + 1] histogram / cursor + error flag */
    int32_t* h_mapped; /* page-locked, device-mapped copy of the histogram: read back by a kernel's stores, so the
                          read-back never queues behind bulk copies on a copy engine */
    int32_t* d_mapped;
    std::vector<int32_t> h_counts;
    std::vector<Group> groups;
    int32_t n_launches;
    /* dcol_plan_refine: second permutation buffer (the two are swapped), group offsets and bin counters */
    int32_t* d_perm_alt = nullptr;
    int32_t* d_gstart = nullptr;   /* [n_groups + 1] first plan position of every group */
    int32_t* d_bins = nullptr;     /* [2][n_groups * kRefineBins] histogram, cursors       */
    int32_t refine_groups = -1;    /* group count d_gstart / d_bins were sized for          */
    /* lane-refill path: scratch records (allocated by the first solve that takes the path) */
    mutable double* d_state = nullptr;
    mutable int64_t state_stride = 0, state_pairs = 0, state_groups = 0;
};

/* ------------------------------------------------------------------------------------------ */
/* plan kernels: counting sort of the pairs by key = idx1 * n_shapes + idx2                     */
namespace {

constexpr int kPlanThreads = 256;
constexpr int kPlanTile = 4096;    /* consecutive pairs per CTA            */
constexpr int kPlanSmemKeys = 4096; /* keys a CTA can privatise in shared memory */

__global__ void plan_histogram(const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2, int64_t B,
                               int32_t n_shapes, int32_t n_keys, int32_t* __restrict__ counts, int32_t* __restrict__ err)
{
    __shared__ int32_t local[kPlanSmemKeys];
    const bool priv = n_keys <= kPlanSmemKeys;
    if (priv) {
        for (int i = threadIdx.x; i < n_keys; i += kPlanThreads) local[i] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * kPlanTile;
    for (int o = threadIdx.x; o < kPlanTile; o += kPlanThreads) {
        const int64_t k = base + o;
        if (k >= B) break;
        const int32_t a = idx1[k], c = idx2[k];
        if (a < 0 || a >= n_shapes || c < 0 || c >= n_shapes) {
            *err = 1;
            continue;
        }
        const int32_t key = a * n_shapes + c;
        if (priv) atomicAdd(&local[key], 1);
        else atomicAdd(&counts[key], 1);
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_keys; i += kPlanThreads)
            if (local[i]) atomicAdd(&counts[i], local[i]);
    }
}

/* cursor[key] starts at the group's first slot; every CTA reserves a contiguous range per key, so
 * pairs stay in increasing order inside a CTA's tile and nearly so across the group (coalescing) */
__global__ void plan_scatter(const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2, int64_t B,
                             int32_t n_shapes, int32_t n_keys, int32_t* __restrict__ cursor, int32_t* __restrict__ perm)
{
    __shared__ int32_t local[kPlanSmemKeys];
    __shared__ int32_t base_of[kPlanSmemKeys];
    const bool priv = n_keys <= kPlanSmemKeys;
    const int64_t base = (int64_t)blockIdx.x * kPlanTile;
    if (priv) {
        for (int i = threadIdx.x; i < n_keys; i += kPlanThreads) local[i] = 0;
        __syncthreads();
        for (int o = threadIdx.x; o < kPlanTile; o += kPlanThreads) {
            const int64_t k = base + o;
            if (k >= B) break;
            const int32_t a = idx1[k], c = idx2[k];
            if (a < 0 || a >= n_shapes || c < 0 || c >= n_shapes) continue;
            atomicAdd(&local[a * n_shapes + c], 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_keys; i += kPlanThreads) {
            base_of[i] = local[i] ? atomicAdd(&cursor[i], local[i]) : 0;
            local[i] = 0;
        }
        __syncthreads();
    }
    /* ordered inside the tile: thread-sequential chunks would need a scan; a shared-memory ticket is
     * enough because the solve only needs every pair exactly once */
    for (int o = threadIdx.x; o < kPlanTile; o += kPlanThreads) {
        const int64_t k = base + o;
        if (k >= B) break;
        const int32_t a = idx1[k], c = idx2[k];
        if (a < 0 || a >= n_shapes || c < 0 || c >= n_shapes) continue;
        const int32_t key = a * n_shapes + c;
        const int32_t pos = priv ? base_of[key] + atomicAdd(&local[key], 1) : atomicAdd(&cursor[key], 1);
        perm[pos] = (int32_t)k;
    }
}

/* ---- dcol_plan_refine: inside every group, order the pairs by the iteration count of their last solve ----
 * A warp runs until its slowest pair converges (22-26 of 32 lanes are active on random batches).  A caller that
 * re-solves the SAME pair list with slowly changing poses (every AL-iLQR pass, ALTRO.py:276-314) knows each pair's
 * iteration count from the previous solve, and it barely changes from pass to pass; grouping pairs of equal count
 * into the same warps removes most of the idle lanes.  Counting sort of the plan positions by
 * (group, iterations descending, clamped to 0..63): histogram -> one-CTA exclusive scan -> scatter, all on the stream, no
 * host sync. */
constexpr int kRefineBins = 64;
constexpr int kRefineMaxGroups = 1024;

__device__ __forceinline__ int refine_group_of(const int32_t* __restrict__ gstart, int lo, int hi, int64_t pos)
{
    while (lo < hi) { /* last group whose first position is <= pos, in [lo, hi] */
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)gstart[mid] <= pos) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

/* longest first: the CTAs that finish a group's grid are then the short ones (a sorted-ascending order would leave the
 * slowest warps for the tail of the grid) */
__device__ __forceinline__ int32_t refine_bin(int32_t it)
{
    it = it < 0 ? 0 : (it >= kRefineBins ? kRefineBins - 1 : it);
    return kRefineBins - 1 - it;
}

template <bool SCATTER>
__global__ void refine_pass(const int32_t* __restrict__ perm, const int32_t* __restrict__ iters,
                            const int32_t* __restrict__ gstart, int32_t n_groups, int64_t B, int32_t* __restrict__ bins,
                            int32_t* __restrict__ perm_out)
{
    __shared__ int32_t local[kPlanSmemKeys];
    __shared__ int32_t base_of[SCATTER ? kPlanSmemKeys : 1];
    const int64_t base = (int64_t)blockIdx.x * kPlanTile;
    const int64_t last = (base + kPlanTile <= B ? base + kPlanTile : B) - 1;
    const int g_lo = refine_group_of(gstart, 0, n_groups - 1, base);
    const int g_hi = refine_group_of(gstart, g_lo, n_groups - 1, last);
    const int span = (g_hi - g_lo + 1) * kRefineBins;
    const bool priv = span <= kPlanSmemKeys; /* a tile of consecutive plan positions touches few groups */
    const int32_t bin0 = g_lo * kRefineBins;
    if (priv) {
        for (int i = threadIdx.x; i < span; i += kPlanThreads) local[i] = 0;
        __syncthreads();
    }
    if (!SCATTER || priv) {
        for (int o = threadIdx.x; o < kPlanTile; o += kPlanThreads) {
            const int64_t pos = base + o;
            if (pos >= B) break;
            const int g = refine_group_of(gstart, g_lo, g_hi, pos);
            const int32_t bin = g * kRefineBins + refine_bin(iters[perm[pos]]);
            if (priv) atomicAdd(&local[bin - bin0], 1);
            else atomicAdd(&bins[bin], 1);
        }
    }
    if (!SCATTER) {
        if (priv) {
            __syncthreads();
            for (int i = threadIdx.x; i < span; i += kPlanThreads)
                if (local[i]) atomicAdd(&bins[bin0 + i], local[i]);
        }
        return;
    }
    if (priv) { /* reserve one contiguous range per bin for this tile, then hand out tickets inside it */
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += kPlanThreads) {
            base_of[i] = local[i] ? atomicAdd(&bins[bin0 + i], local[i]) : 0;
            local[i] = 0;
        }
        __syncthreads();
    }
    for (int o = threadIdx.x; o < kPlanTile; o += kPlanThreads) {
        const int64_t pos = base + o;
        if (pos >= B) break;
        const int g = refine_group_of(gstart, g_lo, g_hi, pos);
        const int32_t pair = perm[pos];
        const int32_t bin = g * kRefineBins + refine_bin(iters[pair]);
        const int32_t dst = priv ? base_of[bin - bin0] + atomicAdd(&local[bin - bin0], 1) : atomicAdd(&bins[bin], 1);
        perm_out[dst] = pair;
    }
}

/* exclusive scan of the n bin counts into cursors (absolute plan positions: the bins are ordered by (group, iters) and
 * the groups are contiguous); one CTA */
__global__ void refine_scan(const int32_t* __restrict__ counts, int32_t n, int32_t* __restrict__ cursor)
{
    __shared__ int32_t part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    int32_t s = 0;
    for (int i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int32_t run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        cursor[i] = run;
        run += counts[i];
    }
}

__global__ void plan_export_counts(const int32_t* __restrict__ counts, int32_t* __restrict__ mapped, int32_t n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mapped[i] = counts[i];
    __threadfence_system();
}

/* pairs the reference cannot assemble (combine_problem_matrices.py:58-67 raises ValueError) */
__global__ void fill_unsupported(BatchArgs b)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.count) return;
    const int64_t k = b.perm ? (int64_t)b.perm[b.first + t] : b.first + t;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (b.flags & DCOL_WANT_CONTACT)
        for (int j = 0; j < 3; ++j) b.contact[3 * k + j] = nan;
    if (b.n_dest > 0) {
        const double last = __longlong_as_double((long long)DCOL_STATUS_UNSUPPORTED << 32);
        for (int d = 0; d < b.n_dest; ++d) {
            double* r = b.dest[d] + kRecordWords * (b.record_offset + b.first + t);
            for (int j = 0; j < kRecordWords; ++j) {
                const double v = j < 13 ? nan : last;
                if (b.flags & DCOL_DEST_MULTICAST) asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(r + j), "d"(v) : "memory");
                else r[j] = v;
            }
        }
        return;
    }
    if (b.jac)
        for (int j = 0; j < 48; ++j) b.jac[48 * k + j] = nan;
    b.status[k] = DCOL_STATUS_UNSUPPORTED;
    b.iters[k] = 0;
    b.alpha[k] = nan;
    if (b.flags & DCOL_WANT_GRAD) {
        const int gw = (b.flags & DCOL_WANT_GRAD1) ? 6 : 12;
        for (int j = 0; j < gw; ++j) b.grad[gw * k + j] = nan;
    }
}

/* scene entry point: pair (m, j) = victim pose m against obstacle j, pairs ordered [m][j]; one thread per pose component */
__global__ void scene_expand(int32_t victim_shape, const int32_t* __restrict__ obs_shape, const double* __restrict__ vic_pose,
                             const double* __restrict__ obs_pose, int64_t n_pairs, int32_t n_obs, int32_t* __restrict__ idx1,
                             int32_t* __restrict__ idx2, double* __restrict__ pose1, double* __restrict__ pose2)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * n_pairs) return;
    const int64_t pair = t / 6;
    const int c = (int)(t - 6 * pair);
    const int64_t m = pair / n_obs;
    const int32_t j = (int32_t)(pair - m * n_obs);
    pose1[t] = vic_pose[6 * m + c];
    pose2[t] = obs_pose[6 * (int64_t)j + c];
    if (c == 0 && idx1) {
        idx1[pair] = victim_shape;
        idx2[pair] = obs_shape[j];
    }
}

/* FP64 FMA peak: 8 independent chains per thread, no memory traffic */
__global__ void fp64_peak_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s; /* never true: keeps the chains alive */
}

cudaError_t launch_group(const dcol_shape_table* T, int32_t i1, int32_t i2, const BatchArgs& args, cudaStream_t stream)
{
    GroupLaunch g = { &T->shapes[i1], &T->shapes[i2], T->A.data(), T->b.data(), args };
    const int c2 = T->cls[i2];
    if (args.jac) { /* solution-Jacobian kernels (dcol_jac_*.cu) */
        switch (T->cls[i1]) {
        case CLS_POLY6: return launch_first_class_jac<CLS_POLY6>(c2, g, stream);
        case CLS_POLY8: return launch_first_class_jac<CLS_POLY8>(c2, g, stream);
        case CLS_POLYN: return launch_first_class_jac<CLS_POLYN>(c2, g, stream);
        case CLS_CAPSULE: return launch_first_class_jac<CLS_CAPSULE>(c2, g, stream);
        case CLS_CYLINDER: return launch_first_class_jac<CLS_CYLINDER>(c2, g, stream);
        case CLS_CONE: return launch_first_class_jac<CLS_CONE>(c2, g, stream);
        case CLS_SPHERE: return launch_first_class_jac<CLS_SPHERE>(c2, g, stream);
        case CLS_PGON5: return launch_first_class_jac<CLS_PGON5>(c2, g, stream);
        case CLS_PGONN: return launch_first_class_jac<CLS_PGONN>(c2, g, stream);
        case CLS_BOX: return launch_first_class_jac<CLS_BOX>(c2, g, stream);
        case CLS_ELLIPSOID: return launch_first_class_jac<CLS_ELLIPSOID>(c2, g, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (T->cls[i1]) {
    case CLS_POLY6: return launch_first_class<CLS_POLY6>(c2, g, stream);
    case CLS_POLY8: return launch_first_class<CLS_POLY8>(c2, g, stream);
    case CLS_POLYN: return launch_first_class<CLS_POLYN>(c2, g, stream);
    case CLS_CAPSULE: return launch_first_class<CLS_CAPSULE>(c2, g, stream);
    case CLS_CYLINDER: return launch_first_class<CLS_CYLINDER>(c2, g, stream);
    case CLS_CONE: return launch_first_class<CLS_CONE>(c2, g, stream);
    case CLS_SPHERE: return launch_first_class<CLS_SPHERE>(c2, g, stream);
    case CLS_PGON5: return launch_first_class<CLS_PGON5>(c2, g, stream);
    case CLS_PGONN: return launch_first_class<CLS_PGONN>(c2, g, stream);
    case CLS_BOX: return launch_first_class<CLS_BOX>(c2, g, stream);
    case CLS_ELLIPSOID: return launch_first_class<CLS_ELLIPSOID>(c2, g, stream);
    default: return cudaErrorInvalidValue;
    }
}

int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(DCOL_E_NOGPU, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(DCOL_E_ARG, "device ordinal out of range");
    return 0;
}

} /* namespace */

/* ------------------------------------------------------------------------------------------ */
extern "C" {

const char* dcol_version(void) { return "dcol-b200 0.1 (abi 1, sm_100a)"; }
const char* dcol_last_error(void) { return g_last_error.c_str(); }

int dcol_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dcol_shape_table_create(const dcol_shape* shapes, int32_t n_shapes, const double* A, const double* b,
                            int32_t n_faces, int device, dcol_shape_table** out)
{
    if (!shapes || !out || n_shapes <= 0 || n_faces < 0 || (n_faces > 0 && (!A || !b)))
        return fail(DCOL_E_ARG, "dcol_shape_table_create: bad argument");
    if (n_shapes > 1024) return fail(DCOL_E_ARG, "dcol_shape_table_create: at most 1024 shapes per table");
    int rc = check_device(device);
    if (rc) return rc;
    dcol_shape_table* T = new dcol_shape_table();
    T->device = device;
    T->shapes.assign(shapes, shapes + n_shapes);
    T->A.assign(A, A + 3 * (size_t)n_faces);
    T->b.assign(b, b + (size_t)n_faces);
    if (n_faces == 0) {
        T->A.assign(3, 0.0);
        T->b.assign(1, 0.0);
    }
    T->cls.resize(n_shapes);
    for (int i = 0; i < n_shapes; ++i) {
        const dcol_shape& s = T->shapes[i];
        const bool faces_ok = !(s.type == DCOL_POLYTOPE || s.type == DCOL_POLYGON) ||
                              (s.face_off >= 0 && s.n_faces >= 0 && s.face_off + s.n_faces <= n_faces);
        const int c = faces_ok ? shape_class(s, T->A.data()) : -1;
        const bool faces = s.type == DCOL_POLYTOPE || s.type == DCOL_POLYGON;
        if (c < 0 || (faces && (s.face_off < 0 || s.face_off + s.n_faces > n_faces))) {
            delete T;
            return fail(DCOL_E_SHAPE, "dcol_shape_table_create: malformed shape record");
        }
        T->cls[i] = c;
    }
    *out = T;
    return 0;
}

void dcol_shape_table_destroy(dcol_shape_table* T)
{
    if (!T) return;
    DeviceGuard guard_(T->device);
    cudaDeviceSynchronize(); /* the buffers go back to the pool, which hands them out again without waiting */
    for (int i = 0; i < dcol_shape_table::kSlots; ++i) {
        T->scratch[i].release();
        dcol_plan_destroy(T->plans[i]);
        if (T->ev_in[i]) cudaEventDestroy(T->ev_in[i]);
        if (T->ev_plan[i]) cudaEventDestroy(T->ev_plan[i]);
        if (T->ev_done[i]) cudaEventDestroy(T->ev_done[i]);
        if (T->ev_out[i]) cudaEventDestroy(T->ev_out[i]);
    }
    pool_free(T->scene_vic[0]);
    pool_free(T->scene_vic[1]);
    pool_free(T->scene_obs_pose);
    pool_free(T->scene_obs_shape);
    for (int i = 0; i < 4; ++i)
        if (T->streams[i]) cudaStreamDestroy(T->streams[i]);
    for (int i = 0; i < dcol_shape_table::kSlots; ++i)
        if (T->run[i]) cudaStreamDestroy(T->run[i]);
    for (int i = 0; i < dcol_shape_table::kSide; ++i) {
        if (T->side[i]) cudaStreamDestroy(T->side[i]);
        if (T->join_ev[i]) cudaEventDestroy(T->join_ev[i]);
    }
    if (T->fork_ev) cudaEventDestroy(T->fork_ev);
    delete T;
}

/* Builds the grouping of B pairs into a plan whose device buffers (counts: n_keys + 1 ints, perm: B ints)
 * already exist.  Synchronises `stream` once, after the histogram, to read the group sizes; the scatter
 * is left in flight on `stream`. */
static int plan_build(dcol_plan* P, const int32_t* d_idx1, const int32_t* d_idx2, int64_t B, cudaStream_t stream)
{
    const dcol_shape_table* T = P->table;
    const int32_t ns = (int32_t)T->shapes.size();
    const int32_t nk = ns * ns;
    P->B = B;
    P->groups.clear();
    P->n_launches = 0;
    if (B == 0) return 0;
    (void)cudaGetLastError(); /* drop a stale error left by another library in this process */
    const unsigned blocks = (unsigned)((B + kPlanTile - 1) / kPlanTile);
    P->h_counts.resize((size_t)nk + 1);
    DCOL_CUDA(cudaMemsetAsync(P->d_counts, 0, sizeof(int32_t) * ((size_t)nk + 1), stream));
    plan_histogram<<<blocks, kPlanThreads, 0, stream>>>(d_idx1, d_idx2, B, ns, nk, P->d_counts, P->d_counts + nk);
    DCOL_CUDA(cudaGetLastError());
    plan_export_counts<<<(nk + 1 + 255) / 256 > 64 ? 64 : (nk + 1 + 255) / 256, 256, 0, stream>>>(P->d_counts, P->d_mapped, nk + 1);
    DCOL_CUDA(cudaGetLastError());
    DCOL_CUDA(cudaStreamSynchronize(stream));
    memcpy(P->h_counts.data(), P->h_mapped, sizeof(int32_t) * ((size_t)nk + 1));
    if (P->h_counts[nk] != 0) return fail(DCOL_E_INDEX, "shape index out of range");
    int64_t off = 0;
    for (int32_t key = 0; key < nk; ++key) {
        const int32_t cnt = P->h_counts[key];
        P->h_counts[key] = (int32_t)off; /* becomes the scatter cursor */
        if (cnt == 0) continue;
        Group g;
        g.i1 = key / ns;
        g.i2 = key % ns;
        g.first = off;
        g.count = cnt;
        g.supported = class_pair_supported(T->cls[g.i1], T->cls[g.i2]);
        P->groups.push_back(g);
        off += cnt;
    }
    P->n_launches = (int32_t)P->groups.size();
    /* cursors go back the same way: written into the mapped buffer, fetched by a kernel (no copy engine) */
    memcpy(P->h_mapped, P->h_counts.data(), sizeof(int32_t) * (size_t)nk);
    plan_export_counts<<<(nk + 255) / 256 > 64 ? 64 : (nk + 255) / 256, 256, 0, stream>>>(P->d_mapped, P->d_counts, nk);
    DCOL_CUDA(cudaGetLastError());
    plan_scatter<<<blocks, kPlanThreads, 0, stream>>>(d_idx1, d_idx2, B, ns, nk, P->d_counts, P->d_perm);
    DCOL_CUDA(cudaGetLastError());
    return 0;
}

/* The same grouping when the shape indices are (also) in HOST memory: the histogram is counted by the calling thread,
 * so nothing waits for the device (on the device the counting kernel only gets CTA slots when the pair kernels of the
 * chunk before drain — stream priorities do not help against a grid that owns every register — and the host thread
 * that waited for it could not enqueue the next chunk's copies; measured with DCOL_HOST_TRACE).  Only the scatter
 * runs on `stream`; it reads the cursors from the plan's mapped buffer when it executes, so the caller must not
 * rebuild this plan before that has happened (the host entry point waits on the slot's ev_plan). */
static int plan_build_host_counts(dcol_plan* P, const int32_t* h_idx1, const int32_t* h_idx2, const int32_t* d_idx1,
                                  const int32_t* d_idx2, int64_t B, cudaStream_t stream)
{
    const dcol_shape_table* T = P->table;
    const int32_t ns = (int32_t)T->shapes.size();
    const int32_t nk = ns * ns;
    P->B = B;
    P->groups.clear();
    P->n_launches = 0;
    if (B == 0) return 0;
    (void)cudaGetLastError();
    /* four interleaved histograms: consecutive pairs very often share a key (scenes), and one counter would then
     * serialise on its own store-to-load forwarding */
    const int lanes = nk <= 65536 ? 4 : 1;
    std::vector<int32_t>& cnt = P->h_counts;
    cnt.assign((size_t)nk * lanes + 1, 0);
    const uint32_t uns = (uint32_t)ns;
    bool bad = false;
    int64_t k = 0;
    if (lanes == 4) {
        int32_t* c0 = cnt.data();
        int32_t *c1 = c0 + nk, *c2 = c1 + nk, *c3 = c2 + nk;
        for (; k + 4 <= B; k += 4) {
            const uint32_t a0 = (uint32_t)h_idx1[k], b0 = (uint32_t)h_idx2[k], a1 = (uint32_t)h_idx1[k + 1], b1 = (uint32_t)h_idx2[k + 1];
            const uint32_t a2 = (uint32_t)h_idx1[k + 2], b2 = (uint32_t)h_idx2[k + 2], a3 = (uint32_t)h_idx1[k + 3], b3 = (uint32_t)h_idx2[k + 3];
            if ((a0 >= uns) | (b0 >= uns) | (a1 >= uns) | (b1 >= uns) | (a2 >= uns) | (b2 >= uns) | (a3 >= uns) | (b3 >= uns)) {
                bad = true;
                break;
            }
            ++c0[a0 * uns + b0];
            ++c1[a1 * uns + b1];
            ++c2[a2 * uns + b2];
            ++c3[a3 * uns + b3];
        }
    }
    for (; k < B && !bad; ++k) {
        const uint32_t a = (uint32_t)h_idx1[k], b = (uint32_t)h_idx2[k];
        if (a >= uns || b >= uns) bad = true;
        else ++cnt[(size_t)a * uns + b];
    }
    if (bad) return fail(DCOL_E_INDEX, "shape index out of range");
    int64_t off = 0;
    for (int32_t key = 0; key < nk; ++key) {
        int32_t c = cnt[key];
        for (int l = 1; l < lanes; ++l) c += cnt[(size_t)l * nk + key];
        P->h_mapped[key] = (int32_t)off; /* the scatter cursor */
        if (c == 0) continue;
        Group g;
        g.i1 = key / ns;
        g.i2 = key % ns;
        g.first = off;
        g.count = c;
        g.supported = class_pair_supported(T->cls[g.i1], T->cls[g.i2]);
        P->groups.push_back(g);
        off += c;
    }
    P->n_launches = (int32_t)P->groups.size();
    plan_export_counts<<<(nk + 255) / 256 > 64 ? 64 : (nk + 255) / 256, 256, 0, stream>>>(P->d_mapped, P->d_counts, nk);
    DCOL_CUDA(cudaGetLastError());
    const unsigned blocks = (unsigned)((B + kPlanTile - 1) / kPlanTile);
    plan_scatter<<<blocks, kPlanThreads, 0, stream>>>(d_idx1, d_idx2, B, ns, nk, P->d_counts, P->d_perm);
    DCOL_CUDA(cudaGetLastError());
    return 0;
}

static int plan_alloc(const dcol_shape_table* T, int64_t capacity, dcol_plan** out)
{
    const int32_t ns = (int32_t)T->shapes.size();
    dcol_plan* P = new dcol_plan();
    P->table = T;
    P->device = T->device;
    P->B = 0;
    P->capacity = capacity;
    P->d_perm = nullptr;
    P->d_counts = nullptr;
    P->h_mapped = nullptr;
    P->d_mapped = nullptr;
    P->n_launches = 0;
    cudaError_t e = pool_alloc_t(T->device, (size_t)ns * ns + 1, &P->d_counts);
    if (e == cudaSuccess) e = pool_alloc(T->device, 1, sizeof(int32_t) * ((size_t)ns * ns + 1), (void**)&P->h_mapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&P->d_mapped, P->h_mapped, 0);
    if (e == cudaSuccess && capacity > 0) e = pool_alloc_t(T->device, (size_t)capacity, &P->d_perm);
    if (e != cudaSuccess) {
        pool_free(P->d_counts);
        pool_free(P->h_mapped);
        pool_free(P->d_perm);
        delete P;
        return fail_cuda(e, "plan allocation");
    }
    *out = P;
    return 0;
}

int dcol_plan_create(const dcol_shape_table* T, const int32_t* d_idx1, const int32_t* d_idx2, int64_t B, void* stream_,
                     dcol_plan** out)
{
    if (!T || !out || B < 0 || (B > 0 && (!d_idx1 || !d_idx2))) return fail(DCOL_E_ARG, "dcol_plan_create: bad argument");
    if (B > 0x7fffffffLL) return fail(DCOL_E_ARG, "dcol_plan_create: at most 2^31-1 pairs per plan");
    DCOL_DEVICE(T->device);
    dcol_plan* P = nullptr;
    int rc = plan_alloc(T, B, &P);
    if (rc) return rc;
    rc = plan_build(P, d_idx1, d_idx2, B, (cudaStream_t)stream_);
    if (rc) {
        dcol_plan_destroy(P);
        return rc;
    }
    *out = P;
    return 0;
}

void dcol_plan_destroy(dcol_plan* P)
{
    if (!P) return;
    DeviceGuard guard_(P->device); /* not P->table->device: the table may already have been destroyed */
    /* a solve of this plan may still be in flight on the caller's stream: cudaFree would have waited for it, the pool
     * does not, so wait here (d_perm may be either of the two permutation buffers after a refine) */
    cudaDeviceSynchronize();
    pool_free(P->d_perm);
    pool_free(P->d_perm_alt);
    cudaFree(P->d_gstart);
    cudaFree(P->d_bins);
    pool_free(P->d_counts);
    cudaFree(P->d_state);
    pool_free(P->h_mapped);
    delete P;
}
int64_t dcol_plan_size(const dcol_plan* P) { return P ? P->B : 0; }
int32_t dcol_plan_n_groups(const dcol_plan* P) { return P ? (int32_t)P->groups.size() : 0; }
int32_t dcol_plan_n_launches(const dcol_plan* P) { return P ? P->n_launches : 0; }

static int solve_plan(const dcol_plan* P, const double* d_pose1, const double* d_pose2, double tol, int32_t max_iter,
                      uint32_t flags, double* d_alpha, double* d_contact, double* d_grad, int32_t* d_iters,
                      int32_t* d_status, int32_t n_dest, double* const* dest, int64_t record_offset, void* stream_,
                      double* d_jac = nullptr);

int dcol_proximity_batch_device(const dcol_plan* P, const double* d_pose1, const double* d_pose2, double tol,
                                int32_t max_iter, uint32_t flags, double* d_alpha, double* d_contact, double* d_grad,
                                int32_t* d_iters, int32_t* d_status, void* stream_)
{
    if (!P) return fail(DCOL_E_ARG, "dcol_proximity_batch_device: null plan");
    if (max_iter < 1 || max_iter > DCOL_MAX_ITER) return fail(DCOL_E_ARG, "max_iter must be in 1..50");
    if (flags & ~kSolveFlags) return fail(DCOL_E_ARG, "unknown flag");
    if ((flags & DCOL_WANT_GRAD1) && !(flags & DCOL_WANT_GRAD)) return fail(DCOL_E_ARG, "DCOL_WANT_GRAD1 needs DCOL_WANT_GRAD");
    if (P->B == 0) return 0;
    if (!d_pose1 || !d_pose2 || !d_alpha || !d_iters || !d_status || ((flags & DCOL_WANT_CONTACT) && !d_contact) ||
        ((flags & DCOL_WANT_GRAD) && !d_grad))
        return fail(DCOL_E_ARG, "dcol_proximity_batch_device: null buffer");
    return solve_plan(P, d_pose1, d_pose2, tol, max_iter, flags, d_alpha, d_contact, d_grad, d_iters, d_status, 0, nullptr, 0,
                      stream_);
}

int dcol_proximity_batch_jacobian(const dcol_plan* P, const double* d_pose1, const double* d_pose2, double tol,
                                  int32_t max_iter, uint32_t flags, double* d_alpha, double* d_contact, double* d_grad,
                                  double* d_jac, int32_t* d_iters, int32_t* d_status, void* stream_)
{
    if (!P) return fail(DCOL_E_ARG, "dcol_proximity_batch_jacobian: null plan");
    if (max_iter < 1 || max_iter > DCOL_MAX_ITER) return fail(DCOL_E_ARG, "max_iter must be in 1..50");
    if (flags & ~kSolveFlags) return fail(DCOL_E_ARG, "unknown flag");
    if ((flags & DCOL_WANT_GRAD1) && !(flags & DCOL_WANT_GRAD)) return fail(DCOL_E_ARG, "DCOL_WANT_GRAD1 needs DCOL_WANT_GRAD");
    if (P->B == 0) return 0;
    if (!d_pose1 || !d_pose2 || !d_alpha || !d_jac || !d_iters || !d_status || ((flags & DCOL_WANT_CONTACT) && !d_contact) ||
        ((flags & DCOL_WANT_GRAD) && !d_grad))
        return fail(DCOL_E_ARG, "dcol_proximity_batch_jacobian: null buffer");
    return solve_plan(P, d_pose1, d_pose2, tol, max_iter, flags, d_alpha, d_contact, d_grad, d_iters, d_status, 0, nullptr, 0,
                      stream_, d_jac);
}

int dcol_proximity_batch_records(const dcol_plan* P, const double* d_pose1, const double* d_pose2, double tol,
                                 int32_t max_iter, uint32_t flags, int32_t n_dest, double* const* dest,
                                 int64_t record_offset, double* d_contact, void* stream_)
{
    if (flags & ~(uint32_t)(DCOL_FIX_CASE4 | DCOL_DEST_MULTICAST | DCOL_ONE_PAIR_PER_THREAD | DCOL_LANE_REFILL)) return fail(DCOL_E_ARG, "unknown flag");
    if ((flags & DCOL_DEST_MULTICAST) && n_dest != 1) return fail(DCOL_E_ARG, "a multicast destination must be the only one");
    if (!P) return fail(DCOL_E_ARG, "dcol_proximity_batch_records: null plan");
    if (max_iter < 1 || max_iter > DCOL_MAX_ITER) return fail(DCOL_E_ARG, "max_iter must be in 1..50");
    if (n_dest < 1 || n_dest > DCOL_MAX_DEST || !dest || record_offset < 0) return fail(DCOL_E_ARG, "bad destination list");
    for (int d = 0; d < n_dest; ++d)
        if (!dest[d] || ((uintptr_t)dest[d] & 15)) return fail(DCOL_E_ARG, "record destinations must be 16-byte aligned");
    if (P->B == 0) return 0;
    if (!d_pose1 || !d_pose2) return fail(DCOL_E_ARG, "dcol_proximity_batch_records: null buffer");
    return solve_plan(P, d_pose1, d_pose2, tol, max_iter, flags | (d_contact ? DCOL_WANT_CONTACT : 0u), nullptr, d_contact, nullptr,
                      nullptr, nullptr, n_dest, dest, record_offset, stream_);
}

const int32_t* dcol_plan_perm(const dcol_plan* P) { return P ? P->d_perm : nullptr; }

int dcol_plan_refine(dcol_plan* P, const int32_t* d_iters, void* stream_)
{
    if (!P || !d_iters) return fail(DCOL_E_ARG, "dcol_plan_refine: bad argument");
    const int n_groups = (int)P->groups.size();
    if (P->B == 0 || n_groups == 0) return 0;
    if (n_groups > kRefineMaxGroups) return fail(DCOL_E_ARG, "dcol_plan_refine: at most 1024 groups");
    cudaStream_t stream = (cudaStream_t)stream_;
    DCOL_DEVICE(P->device);
    (void)cudaGetLastError();
    const int n_bins = n_groups * kRefineBins;
    if (!P->d_perm_alt) DCOL_CUDA(pool_alloc_t(P->device, (size_t)P->capacity, &P->d_perm_alt));
    if (P->refine_groups != n_groups) { /* the host entry point rebuilds its cached plans: group lists change */
        cudaFree(P->d_gstart);
        cudaFree(P->d_bins);
        P->d_gstart = P->d_bins = nullptr;
        DCOL_CUDA(cudaMalloc(&P->d_gstart, sizeof(int32_t) * (size_t)(n_groups + 1)));
        DCOL_CUDA(cudaMalloc(&P->d_bins, sizeof(int32_t) * 2 * (size_t)n_bins));
        P->refine_groups = n_groups;
    }
    std::vector<int32_t> gstart(n_groups + 1);
    for (int g = 0; g < n_groups; ++g) gstart[g] = (int32_t)P->groups[g].first;
    gstart[n_groups] = (int32_t)P->B;
    /* pageable source: the copy is staged before the call returns, so the vector may die */
    DCOL_CUDA(cudaMemcpyAsync(P->d_gstart, gstart.data(), sizeof(int32_t) * gstart.size(), cudaMemcpyHostToDevice, stream));
    DCOL_CUDA(cudaMemsetAsync(P->d_bins, 0, sizeof(int32_t) * (size_t)n_bins, stream));
    const unsigned blocks = (unsigned)((P->B + kPlanTile - 1) / kPlanTile);
    refine_pass<false><<<blocks, kPlanThreads, 0, stream>>>(P->d_perm, d_iters, P->d_gstart, n_groups, P->B, P->d_bins, nullptr);
    DCOL_CUDA(cudaGetLastError());
    refine_scan<<<1, 1024, 0, stream>>>(P->d_bins, n_bins, P->d_bins + n_bins);
    DCOL_CUDA(cudaGetLastError());
    refine_pass<true><<<blocks, kPlanThreads, 0, stream>>>(P->d_perm, d_iters, P->d_gstart, n_groups, P->B, P->d_bins + n_bins,
                                                           P->d_perm_alt);
    DCOL_CUDA(cudaGetLastError());
    std::swap(P->d_perm, P->d_perm_alt);
    return 0;
}

static int solve_plan(const dcol_plan* P, const double* d_pose1, const double* d_pose2, double tol, int32_t max_iter,
                      uint32_t flags, double* d_alpha, double* d_contact, double* d_grad, int32_t* d_iters,
                      int32_t* d_status, int32_t n_dest, double* const* dest, int64_t record_offset, void* stream_,
                      double* d_jac)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    dcol_shape_table* T = const_cast<dcol_shape_table*>(P->table);
    DCOL_DEVICE(T->device);
    (void)cudaGetLastError(); /* drop a stale error left by another library in this process */
    const int n_groups = (int)P->groups.size();
    static const int side_limit = getenv("DCOL_SIDE_STREAMS") ? std::max(1, std::min(atoi(getenv("DCOL_SIDE_STREAMS")), (int)dcol_shape_table::kSide)) : 8;
    const int n_side = n_groups > 1 ? std::min(n_groups, side_limit) : 0;
    std::lock_guard<std::mutex> lock(T->launch_mu);
    if (n_side) {
        if (!T->fork_ev) DCOL_CUDA(cudaEventCreateWithFlags(&T->fork_ev, cudaEventDisableTiming));
        for (int i = 0; i < n_side; ++i) {
            if (!T->side[i]) DCOL_CUDA(cudaStreamCreateWithFlags(&T->side[i], cudaStreamNonBlocking));
            if (!T->join_ev[i]) DCOL_CUDA(cudaEventCreateWithFlags(&T->join_ev[i], cudaEventDisableTiming));
        }
        DCOL_CUDA(cudaEventRecord(T->fork_ev, stream));
        for (int i = 0; i < n_side; ++i) DCOL_CUDA(cudaStreamWaitEvent(T->side[i], T->fork_ev, 0));
    }
    /* lane-refill path (DCOL_LANE_REFILL, or the library default): scratch records for the pairs of this plan */
    double* d_state = nullptr;
    int64_t state_stride = 0;
    if (!d_jac && !(flags & DCOL_ONE_PAIR_PER_THREAD) && ((flags & DCOL_LANE_REFILL) || refill_enabled())) {
        int words = 0;
        for (const Group& g : P->groups)
            if (g.supported || (flags & DCOL_FIX_CASE4)) words = std::max(words, state_words(T->cls[g.i1], T->cls[g.i2]));
        if (words > 0) {
            if (P->state_stride < words || P->state_pairs < P->B || P->state_groups < (int64_t)P->groups.size()) {
                cudaFree(P->d_state);
                P->d_state = nullptr;
                P->state_stride = P->state_pairs = P->state_groups = 0;
                /* every group's records start on a block boundary: 32 spare records per group */
                DCOL_CUDA(cudaMalloc(&P->d_state, sizeof(double) * (size_t)words * (size_t)(P->capacity_or_size() + 32 * (int64_t)P->groups.size())));
                P->state_stride = words;
                P->state_pairs = P->capacity_or_size();
                P->state_groups = (int64_t)P->groups.size();
            }
            d_state = P->d_state;
            state_stride = P->state_stride;
        }
    }
    /* largest groups first: the long grids start early, the short ones fill the tail */
    std::vector<int> order(n_groups);
    for (int i = 0; i < n_groups; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return P->groups[a].count > P->groups[b].count; });
    int rc = 0;
    /* DCOL_SOLVE_TRACE=1 (diagnostic, synchronises): start / end of every group's kernel on the device, on stderr */
    const bool trace = getenv("DCOL_SOLVE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    cudaEvent_t tev0 = nullptr;
    if (trace) {
        cudaEventCreate(&tev0);
        cudaEventRecord(tev0, stream);
    }
    for (int oi = 0; oi < n_groups && rc == 0; ++oi) {
        const Group& g = P->groups[order[oi]];
        cudaStream_t st = n_side ? T->side[oi % n_side] : stream;
        if (trace) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            tev.push_back(e);
        }
        BatchArgs a = { P->d_perm, g.first, g.count, d_pose1, d_pose2, tol, max_iter, flags,
                        d_alpha, d_contact, d_grad, d_iters, d_status, nullptr, n_dest, record_offset, {} };
        for (int d = 0; d < n_dest; ++d) a.dest[d] = dest[d];
        a.jac = d_jac;
        /* group order[oi]'s records: after all the pairs of the groups before it in plan order, plus 32 spare per group */
        a.state = d_state ? d_state + (g.first + 32 * (int64_t)order[oi]) * state_stride : nullptr;
        a.state_stride = state_stride;
        cudaError_t e;
        if (!g.supported && !(flags & DCOL_FIX_CASE4)) {
            fill_unsupported<<<(unsigned)((g.count + 255) / 256), 256, 0, st>>>(a);
            e = cudaGetLastError();
        } else {
            e = launch_group(T, g.i1, g.i2, a, st);
        }
        if (e != cudaSuccess) rc = fail_cuda(e, "pair_kernel launch");
        if (trace) {
            cudaEvent_t e2;
            cudaEventCreate(&e2);
            cudaEventRecord(e2, st);
            tev.push_back(e2);
        }
    }
    if (trace) {
        cudaDeviceSynchronize();
        fprintf(stderr, "solve trace: %d groups on %d side streams; per group: stream, pairs, classes, start us, end us\n", n_groups, n_side);
        for (int oi = 0; 2 * oi + 1 < (int)tev.size(); ++oi) {
            const Group& g = P->groups[order[oi]];
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tev0, tev[2 * oi]);
            cudaEventElapsedTime(&b, tev0, tev[2 * oi + 1]);
            fprintf(stderr, "  %2d  s%-2d %8lld  (%d,%d)  %9.1f %9.1f  dur %8.1f\n", oi, n_side ? oi % n_side : 0, (long long)g.count,
                    T->cls[g.i1], T->cls[g.i2], a * 1e3, b * 1e3, (b - a) * 1e3);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
        cudaEventDestroy(tev0);
    }
    for (int i = 0; i < n_side; ++i) { /* join even after a failed launch: nothing may outlive the call's stream order */
        cudaError_t e = cudaEventRecord(T->join_ev[i], T->side[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, T->join_ev[i], 0);
        if (e != cudaSuccess && rc == 0) rc = fail_cuda(e, "stream join");
    }
    return rc;
}

/* streams, events and the two sets of device scratch (chunk pairs each) of the host entry points */
static int host_pipeline_setup(dcol_shape_table* T, int64_t chunk, int n_slots = 2)
{
    /* The counting sort of chunk i+1 runs while the pair kernels of chunk i fill every CTA slot, and the host thread
     * waits for its histogram before it can launch anything else: on an equal-priority stream the sort's CTAs are only
     * scheduled when the solve's grid drains, which serialises the host's per-chunk launch work with the solve.  The
     * plan stream therefore gets the highest priority (its few CTAs take the next slots that free up). */
    int prio_least = 0, prio_greatest = 0;
    DCOL_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    if (getenv("DCOL_HOST_NOPRIO")) prio_greatest = prio_least; /* A/B switch of tools/diag_e2e.py */
    for (int i = 0; i < 4; ++i)
        if (!T->streams[i])
            DCOL_CUDA(cudaStreamCreateWithPriority(&T->streams[i], cudaStreamNonBlocking, i == 1 ? prio_greatest : prio_least));
    for (int i = 0; i < n_slots; ++i) {
        if (!T->run[i]) DCOL_CUDA(cudaStreamCreateWithPriority(&T->run[i], cudaStreamNonBlocking, prio_least));
        if (!T->ev_in[i]) DCOL_CUDA(cudaEventCreateWithFlags(&T->ev_in[i], cudaEventDisableTiming));
        if (!T->ev_plan[i]) DCOL_CUDA(cudaEventCreateWithFlags(&T->ev_plan[i], cudaEventDisableTiming));
        if (!T->ev_done[i]) DCOL_CUDA(cudaEventCreateWithFlags(&T->ev_done[i], cudaEventDisableTiming));
        if (!T->ev_out[i]) DCOL_CUDA(cudaEventCreateWithFlags(&T->ev_out[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < n_slots; ++i) {
        HostScratch& S = T->scratch[i];
        if (S.cap >= chunk) continue;
        S.release();
        dcol_plan_destroy(T->plans[i]);
        T->plans[i] = nullptr;
        if (i < 2) T->scene_key[i].clear();
        const int dv = T->device;
        cudaError_t e = pool_alloc_t(dv, chunk, &S.idx1);
        if (e == cudaSuccess) e = pool_alloc_t(dv, chunk, &S.idx2);
        if (e == cudaSuccess) e = pool_alloc_t(dv, chunk, &S.iters);
        if (e == cudaSuccess) e = pool_alloc_t(dv, chunk, &S.status);
        if (e == cudaSuccess) e = pool_alloc_t(dv, 6 * chunk, &S.pose1);
        if (e == cudaSuccess) e = pool_alloc_t(dv, 6 * chunk, &S.pose2);
        if (e == cudaSuccess) e = pool_alloc_t(dv, chunk, &S.alpha);
        if (e == cudaSuccess) e = pool_alloc_t(dv, 3 * chunk, &S.contact);
        if (e == cudaSuccess) e = pool_alloc_t(dv, 12 * chunk, &S.grad);
        if (e != cudaSuccess) { /* no half-allocated slot is left behind */
            S.release();
            return fail_cuda(e, "host scratch allocation");
        }
        S.cap = chunk;
        int rc0 = plan_alloc(T, chunk, &T->plans[i]);
        if (rc0) {
            S.release();
            return rc0;
        }
    }
    return 0;
}

/* Host buffers: the batch is cut into chunks that flow through three streams (copy in, plan + solve,
 * copy out) with two sets of device scratch, so PCIe traffic in both directions overlaps the solve. */
int dcol_proximity_batch_host(const dcol_shape_table* T_, const int32_t* idx1, const int32_t* idx2, const double* pose1,
                              const double* pose2, int64_t B, double tol, int32_t max_iter, uint32_t flags,
                              double* alpha, double* contact, double* grad, int32_t* iters, int32_t* status)
{
    dcol_shape_table* T = const_cast<dcol_shape_table*>(T_);
    if (!T || B < 0) return fail(DCOL_E_ARG, "dcol_proximity_batch_host: bad argument");
    if (max_iter < 1 || max_iter > DCOL_MAX_ITER) return fail(DCOL_E_ARG, "max_iter must be in 1..50");
    if (flags & ~kSolveFlags) return fail(DCOL_E_ARG, "unknown flag");
    if ((flags & DCOL_WANT_GRAD1) && !(flags & DCOL_WANT_GRAD)) return fail(DCOL_E_ARG, "DCOL_WANT_GRAD1 needs DCOL_WANT_GRAD");
    if (B == 0) return 0;
    if (!idx1 || !idx2 || !pose1 || !pose2 || !alpha || !iters || !status || ((flags & DCOL_WANT_CONTACT) && !contact) ||
        ((flags & DCOL_WANT_GRAD) && !grad))
        return fail(DCOL_E_ARG, "dcol_proximity_batch_host: null buffer");
    std::lock_guard<std::mutex> lock(T->mu);
    DCOL_DEVICE(T->device);
    const int64_t gw = (flags & DCOL_WANT_GRAD1) ? 6 : 12; /* doubles of gradient per pair */
    /* Chunk size, measured on B200 + PCIe 5 (tools/diag_e2e.py, tools/diag_host_trace.py; profiles/r02_host_pipeline.md):
     * the call takes about the copy-in of everything plus solve and copy-out of ONE chunk, so smaller chunks are better —
     * until the solves, whose 40 small grids cost a fixed ~0.55 ms per chunk on top of ~1.07 us per 1,000 pairs, take
     * longer in total than the copies (2^19 pairs: 16 x 1.19 ms against 18.8 ms; 2^18: 32 x 0.86 ms) */
    int64_t kChunk = 1 << 19;
    int n_slots = dcol_shape_table::kSlots;
    if (const char* env = getenv("DCOL_HOST_CHUNK")) kChunk = std::max<int64_t>(1024, atoll(env));
    if (const char* env = getenv("DCOL_HOST_SLOTS")) n_slots = std::max(2, std::min(atoi(env), (int)dcol_shape_table::kSlots));
    const int64_t chunk = std::min<int64_t>(B, kChunk);
    if (int rc0 = host_pipeline_setup(T, chunk, n_slots)) return rc0;
    T->scene_key[0].clear(); /* this call rebuilds the cached plans */
    T->scene_key[1].clear();
    /* Four streams: copy-in, plan (counting sort), solve, copy-out.  The only host wait per chunk is for
     * that chunk's own histogram, so chunk i+1 is copied in and planned while chunk i is being solved and
     * chunk i-1 is being copied out. */
    cudaStream_t s_in = T->streams[0], s_plan = T->streams[1], s_out = T->streams[3];
    const bool one_run_stream = getenv("DCOL_HOST_ONE_RUN_STREAM") != nullptr; /* A/B switches of tools/diag_e2e.py */
    const bool device_count = getenv("DCOL_HOST_DEVICE_COUNT") != nullptr;     /* histogram on the device (with a host wait per chunk) */
    int rc = 0;
    const int64_t n_chunks = (B + chunk - 1) / chunk;
    /* DCOL_HOST_TRACE=1: device timeline of every chunk's stages and the host's enqueue times on stderr (diagnostic) */
    const bool trace = getenv("DCOL_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<double> thost;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    const double host0 = now_ms();
    if (trace) mark(s_in);
    for (int64_t ci = 0; ci < n_chunks && rc == 0; ++ci) {
        const int slot = (int)(ci % n_slots);
        HostScratch& S = T->scratch[slot];
        dcol_plan* P = T->plans[slot];
        cudaStream_t s_run = one_run_stream ? T->streams[2] : T->run[slot];
        const int64_t k0 = ci * chunk, n = std::min(chunk, B - k0);
        if (ci >= n_slots) {
            DCOL_CUDA_BREAK(cudaEventSynchronize(T->ev_plan[slot])); /* the slot's scatter has read its cursors (long ago) */
            DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_in, T->ev_done[slot], 0));  /* inputs + perm consumed by the solve   */
            DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_plan, T->ev_done[slot], 0));
            DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_run, T->ev_out[slot], 0));  /* previous outputs have left the device */
        }
        if (trace) thost.push_back(now_ms() - host0);
        mark(s_in);
        DCOL_CUDA_BREAK(cudaMemcpyAsync(S.idx1, idx1 + k0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s_in));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(S.idx2, idx2 + k0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s_in));
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_in[slot], s_in));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(S.pose1, pose1 + 6 * k0, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s_in));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(S.pose2, pose2 + 6 * k0, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s_in));
        mark(s_in);
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_plan, T->ev_in[slot], 0));
        rc = device_count ? plan_build(P, S.idx1, S.idx2, n, s_plan)
                          : plan_build_host_counts(P, idx1 + k0, idx2 + k0, S.idx1, S.idx2, n, s_plan);
        if (rc) break;
        if (trace) thost.push_back(now_ms() - host0);
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_plan[slot], s_plan));
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_in[slot], s_in)); /* now also covers the poses */
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_run, T->ev_plan[slot], 0));
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_run, T->ev_in[slot], 0));
        mark(s_run);
        rc = dcol_proximity_batch_device(P, S.pose1, S.pose2, tol, max_iter, flags, S.alpha, S.contact, S.grad, S.iters,
                                         S.status, s_run);
        if (rc) break;
        mark(s_run);
        if (trace) thost.push_back(now_ms() - host0);
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_done[slot], s_run));
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_out, T->ev_done[slot], 0));
        mark(s_out);
        DCOL_CUDA_BREAK(cudaMemcpyAsync(alpha + k0, S.alpha, sizeof(double) * n, cudaMemcpyDeviceToHost, s_out));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(iters + k0, S.iters, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s_out));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(status + k0, S.status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s_out));
        if (flags & DCOL_WANT_CONTACT)
            DCOL_CUDA_BREAK(cudaMemcpyAsync(contact + 3 * k0, S.contact, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s_out));
        if (flags & DCOL_WANT_GRAD)
            DCOL_CUDA_BREAK(cudaMemcpyAsync(grad + gw * k0, S.grad, sizeof(double) * gw * n, cudaMemcpyDeviceToHost, s_out));
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_out[slot], s_out));
        mark(s_out);
    }
    cudaError_t e = cudaStreamSynchronize(s_out);
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_plan);
    cudaStreamSynchronize(T->streams[2]);
    for (int i = 0; i < n_slots; ++i) cudaStreamSynchronize(T->run[i]);
    if (trace) {
        /* events per chunk: in0 in1(pose copies enqueued) run0 run1 out0 out1, after the call's first mark */
        fprintf(stderr, "chunk  host:enq sync launched | dev: in0 in1 run0 run1 out0 out1 (ms since the first copy was enqueued)\n");
        for (size_t c = 0; 1 + 6 * (c + 1) <= tev.size() && 3 * (c + 1) <= thost.size(); ++c) {
            float v[6];
            for (int j = 0; j < 6; ++j) cudaEventElapsedTime(&v[j], tev[0], tev[1 + 6 * c + j]);
            fprintf(stderr, "%5zu  %7.3f %7.3f %7.3f | %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f\n", c, thost[3 * c], thost[3 * c + 1],
                    thost[3 * c + 2], v[0], v[1], v[2], v[3], v[4], v[5]);
        }
        fprintf(stderr, "host total %.3f ms\n", now_ms() - host0);
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    }
    if (rc == 0 && e != cudaSuccess) rc = fail_cuda(e, "dcol_proximity_batch_host");
    return rc;
}

/* Scene form of the host entry point (include/dcol.h): M victim poses x n_obs posed obstacles.  Only the M + n_obs
 * poses cross PCIe on the way in; the kernel scene_expand broadcasts them into the pair arrays on the device.  Same
 * chunked pipeline as dcol_proximity_batch_host; the plan of a full chunk is built once and reused (the shape pair of
 * pair (m, j) depends on j only). */
int dcol_proximity_scene_host(const dcol_shape_table* T_, int32_t victim_shape, const double* victim_pose, int64_t M,
                              const int32_t* obstacle_shape, const double* obstacle_pose, int32_t n_obs, double tol,
                              int32_t max_iter, uint32_t flags, double* alpha, double* grad1, int32_t* iters, int32_t* status)
{
    dcol_shape_table* T = const_cast<dcol_shape_table*>(T_);
    if (!T || M < 0 || n_obs < 0) return fail(DCOL_E_ARG, "dcol_proximity_scene_host: bad argument");
    if (max_iter < 1 || max_iter > DCOL_MAX_ITER) return fail(DCOL_E_ARG, "max_iter must be in 1..50");
    if (flags & ~(uint32_t)(DCOL_FIX_CASE4 | DCOL_ONE_PAIR_PER_THREAD | DCOL_LANE_REFILL)) return fail(DCOL_E_ARG, "unknown flag");
    if (M == 0 || n_obs == 0) return 0;
    if (!victim_pose || !obstacle_shape || !obstacle_pose || !alpha || !status)
        return fail(DCOL_E_ARG, "dcol_proximity_scene_host: null buffer");
    const int32_t ns = (int32_t)T->shapes.size();
    if (victim_shape < 0 || victim_shape >= ns) return fail(DCOL_E_INDEX, "shape index out of range");
    for (int32_t j = 0; j < n_obs; ++j)
        if (obstacle_shape[j] < 0 || obstacle_shape[j] >= ns) return fail(DCOL_E_INDEX, "shape index out of range");
    std::lock_guard<std::mutex> lock(T->mu);
    DCOL_DEVICE(T->device);
    int64_t kChunk = 1 << 19; /* measured (tools/diag_scene.py, config 5): 2^18 11.6 ms, 2^19 10.75, 2^20 11.15, 2^21 12.2 per 8.3 M pairs */
    if (const char* env = getenv("DCOL_HOST_CHUNK")) kChunk = std::max<int64_t>(1024, atoll(env));
    const int64_t Mc = std::max<int64_t>(1, std::min<int64_t>(M, kChunk / n_obs)); /* victim poses per chunk */
    if (Mc * n_obs > 0x7fffffffLL) return fail(DCOL_E_ARG, "dcol_proximity_scene_host: too many obstacles");
    if (int rc0 = host_pipeline_setup(T, Mc * n_obs)) return rc0;
    if (T->scene_vic_cap < Mc) {
        pool_free(T->scene_vic[0]);
        pool_free(T->scene_vic[1]);
        T->scene_vic[0] = T->scene_vic[1] = nullptr;
        T->scene_vic_cap = 0;
        DCOL_CUDA(pool_alloc_t(T->device, (size_t)(6 * Mc), &T->scene_vic[0]));
        DCOL_CUDA(pool_alloc_t(T->device, (size_t)(6 * Mc), &T->scene_vic[1]));
        T->scene_vic_cap = Mc;
    }
    if (T->scene_obs_cap < n_obs) {
        pool_free(T->scene_obs_pose);
        pool_free(T->scene_obs_shape);
        T->scene_obs_pose = nullptr;
        T->scene_obs_shape = nullptr;
        T->scene_obs_cap = 0;
        DCOL_CUDA(pool_alloc_t(T->device, (size_t)(6 * n_obs), &T->scene_obs_pose));
        DCOL_CUDA(pool_alloc_t(T->device, (size_t)n_obs, &T->scene_obs_shape));
        T->scene_obs_cap = n_obs;
    }
    cudaStream_t s_in = T->streams[0], s_run = T->streams[2], s_out = T->streams[3];
    const uint32_t solve_flags = flags | (grad1 ? (uint32_t)(DCOL_WANT_GRAD | DCOL_WANT_GRAD1) : 0u);
    int rc = 0;
    /* plans[0] serves the full chunks, plans[1] the last, shorter one; each is rebuilt only when its key changes */
    std::vector<int32_t> key_obs(obstacle_shape, obstacle_shape + n_obs);
    const int64_t n_chunks = (M + Mc - 1) / Mc;
    for (int64_t ci = 0; ci < n_chunks && rc == 0; ++ci) {
        const int slot = (int)(ci & 1);
        HostScratch& S = T->scratch[slot];
        const int64_t m0 = ci * Mc, mc = std::min(Mc, M - m0), n = mc * n_obs;
        if (ci == 0) {
            DCOL_CUDA_BREAK(cudaMemcpyAsync(T->scene_obs_pose, obstacle_pose, sizeof(double) * 6 * n_obs, cudaMemcpyHostToDevice, s_in));
            DCOL_CUDA_BREAK(cudaMemcpyAsync(T->scene_obs_shape, obstacle_shape, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, s_in));
        }
        if (ci >= 2) {
            DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_in, T->ev_done[slot], 0));  /* this slot's victim poses were consumed   */
            DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_run, T->ev_out[slot], 0));  /* this slot's outputs have left the device */
        }
        DCOL_CUDA_BREAK(cudaMemcpyAsync(T->scene_vic[slot], victim_pose + 6 * m0, sizeof(double) * 6 * mc, cudaMemcpyHostToDevice, s_in));
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_in[slot], s_in));
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_run, T->ev_in[slot], 0));
        const int which = (mc == Mc) ? 0 : 1;
        std::vector<int32_t> key;
        key.reserve(n_obs + 3);
        key.push_back(victim_shape);
        key.push_back((int32_t)mc);
        key.push_back(n_obs);
        key.insert(key.end(), key_obs.begin(), key_obs.end());
        const bool need_plan = T->scene_key[which] != key;
        scene_expand<<<(unsigned)((6 * n + 255) / 256), 256, 0, s_run>>>(victim_shape, T->scene_obs_shape, T->scene_vic[slot],
                                                                        T->scene_obs_pose, n, n_obs, need_plan ? S.idx1 : nullptr,
                                                                        S.idx2, S.pose1, S.pose2);
        DCOL_CUDA_BREAK(cudaGetLastError());
        dcol_plan* P = T->plans[which];
        if (need_plan) {
            T->scene_key[which].clear();
            rc = plan_build(P, S.idx1, S.idx2, n, s_run);
            if (rc) break;
            T->scene_key[which] = key;
        }
        rc = dcol_proximity_batch_device(P, S.pose1, S.pose2, tol, max_iter, solve_flags, S.alpha, nullptr, grad1 ? S.grad : nullptr,
                                         S.iters, S.status, s_run);
        if (rc) break;
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_done[slot], s_run));
        DCOL_CUDA_BREAK(cudaStreamWaitEvent(s_out, T->ev_done[slot], 0));
        const int64_t k0 = m0 * n_obs;
        DCOL_CUDA_BREAK(cudaMemcpyAsync(alpha + k0, S.alpha, sizeof(double) * n, cudaMemcpyDeviceToHost, s_out));
        DCOL_CUDA_BREAK(cudaMemcpyAsync(status + k0, S.status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s_out));
        if (iters) DCOL_CUDA_BREAK(cudaMemcpyAsync(iters + k0, S.iters, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s_out));
        if (grad1) DCOL_CUDA_BREAK(cudaMemcpyAsync(grad1 + 6 * k0, S.grad, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, s_out));
        DCOL_CUDA_BREAK(cudaEventRecord(T->ev_out[slot], s_out));
    }
    cudaError_t e = cudaStreamSynchronize(s_out);
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_run);
    if (rc == 0 && e != cudaSuccess) rc = fail_cuda(e, "dcol_proximity_scene_host");
    return rc;
}

/* Device buffers that other processes of the node can map (CUDA IPC): the record buffers of the fused
 * all-gather.  Plain cudaMalloc allocations, so the IPC handle refers to exactly this buffer. */
int dcol_device_alloc(int device, size_t bytes, void** out)
{
    if (!out) return fail(DCOL_E_ARG, "null argument");
    int rc = check_device(device);
    if (rc) return rc;
    DCOL_DEVICE(device);
    DCOL_CUDA(cudaMalloc(out, bytes ? bytes : 16));
    return 0;
}
void dcol_device_free(int device, void* p)
{
    if (!p) return;
    DeviceGuard guard_(device);
    cudaFree(p);
}
int dcol_ipc_export(int device, void* dev_ptr, void* handle64)
{
    if (!dev_ptr || !handle64) return fail(DCOL_E_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DCOL_DEVICE(device);
    cudaIpcMemHandle_t h;
    DCOL_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return 0;
}
/* Maps a peer process's buffer into this process (enables peer access between the two GPUs). */
int dcol_ipc_import(int device, const void* handle64, void** out)
{
    if (!handle64 || !out) return fail(DCOL_E_ARG, "null argument");
    DCOL_DEVICE(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DCOL_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
void dcol_ipc_close(int device, void* p)
{
    if (!p) return;
    DeviceGuard guard_(device);
    cudaIpcCloseMemHandle(p);
}

/* page-locked host memory, so that the host entry point's copies run asynchronously at PCIe rate */
int dcol_host_alloc(size_t bytes, void** out)
{
    if (!out) return fail(DCOL_E_ARG, "null argument");
    int rc = check_device(0);
    if (rc) return rc;
    DCOL_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
void dcol_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

void dcol_release_cached(void)
{
    BufPool& P = pool();
    std::multimap<BufPool::Key, void*> take;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        take.swap(P.parked);
        P.parked_bytes = 0;
    }
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (auto& kv : take) {
        if (cudaSetDevice(kv.first.device) != cudaSuccess) continue;
        if (kv.first.kind == 0) cudaFree(kv.second);
        else cudaFreeHost(kv.second);
    }
    if (have_cur) cudaSetDevice(cur);
    cudaGetLastError();
}

int dcol_debug_trace_pair(const dcol_shape_table* T, int32_t idx1, int32_t idx2, const double* pose1, const double* pose2,
                          double tol, double* alpha, double* x, double* s, double* z, int32_t* n, int32_t* m,
                          int32_t* iters, int32_t* status, double* mu_trace)
{
    if (!T || !pose1 || !pose2 || !alpha || !x || !s || !z || !n || !m || !iters || !status || !mu_trace)
        return fail(DCOL_E_ARG, "dcol_debug_trace_pair: null argument");
    const int32_t ns = (int32_t)T->shapes.size();
    if (idx1 < 0 || idx1 >= ns || idx2 < 0 || idx2 >= ns) return fail(DCOL_E_INDEX, "shape index out of range");
    DCOL_DEVICE(T->device);
    const double nan = __builtin_nan("");
    *alpha = nan; *n = 0; *m = 0; *iters = 0;
    for (int i = 0; i <= DCOL_MAX_ITER; ++i) mu_trace[i] = nan;
    if (!class_pair_supported(T->cls[idx1], T->cls[idx2])) {
        *status = DCOL_STATUS_UNSUPPORTED;
        return 0;
    }
    struct Dev {
        double pose[12], alpha;
        int32_t iters, status;
        TraceOut tr;
    };
    Dev* d = nullptr;
    DCOL_CUDA(cudaMalloc(&d, sizeof(Dev)));
    Dev* h = new Dev();
    memcpy(h->pose, pose1, 6 * sizeof(double));
    memcpy(h->pose + 6, pose2, 6 * sizeof(double));
    cudaError_t e = cudaMemcpy(d, h, sizeof(Dev), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        BatchArgs a = { nullptr, 0, 1, d->pose, d->pose + 6, tol, DCOL_MAX_ITER, 0u,
                        &d->alpha, nullptr, nullptr, &d->iters, &d->status, &d->tr, 0, 0, {} };
        e = launch_group(T, idx1, idx2, a, 0);
    }
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(Dev), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        delete h;
        return fail_cuda(e, "dcol_debug_trace_pair");
    }
    *alpha = h->alpha; *iters = h->iters; *status = h->status; *n = h->tr.n; *m = h->tr.m;
    memcpy(x, h->tr.x, sizeof(double) * 8);
    memcpy(s, h->tr.s, sizeof(double) * h->tr.m);
    memcpy(z, h->tr.z, sizeof(double) * h->tr.m);
    memcpy(mu_trace, h->tr.mu, sizeof(double) * (DCOL_MAX_ITER + 1));
    delete h;
    return 0;
}

int dcol_measure_fp64_peak(int device, double* flops_per_s)
{
    if (!flops_per_s) return fail(DCOL_E_ARG, "null argument");
    int rc = check_device(device);
    if (rc) return rc;
    DCOL_DEVICE(device);
    cudaDeviceProp prop;
    DCOL_CUDA(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    DCOL_CUDA(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    DCOL_CUDA(cudaEventCreate(&e0));
    DCOL_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, 0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
        cudaEventRecord(e1, 0);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(d);
            return fail_cuda(e, "fp64_peak_kernel");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *flops_per_s = best;
    return 0;
}

} /* extern "C" */
