/*
 * dcol.h — C ABI of the B200-native batched DCOL proximity solve + gradient.
 *
 * This is the drop-in boundary for ONE path of CogSP/DCOL-trajectory-optimization: the
 * primal-dual interior-point proximity solve and its pose gradient.  Every entry point
 * replaces a piece of the reference's Python call chain (paths relative to the reference
 * root):
 *
 *   dcol_shape / dcol_shape_table_*   primitive attribute bags + isinstance dispatch
 *                                     primitives/misc_primitive_constructor.py:4-88
 *                                     primitives/problem_matrices.py:255-364
 *   dcol_plan_*                       (new) grouping of a batch by type signature; the
 *                                     reference loops pair by pair
 *                                     (systems/cluttered_hallway_quadrotor.py:131-133,155)
 *   dcol_proximity_batch_{device,host}
 *                                     proximity_mrp            proximity/proximity.py:6-54
 *                                     proximity_gradient       proximity/proximity_gradient.py:91-138
 *                                     problem_matrices         primitives/problem_matrices.py:4-364
 *                                     combine_problem_matrices primitives/combine_problem_matrices.py:3-70
 *                                     solve_lp_pdip            proximity/pdip.py:373-470
 *                                     calc_NT_scalings & co    proximity/NT/NT_scaling.py:75-126,205-240,340-463
 *                                     obj_val_grad             proximity/proximity_gradient.py:50-88
 *
 * All arithmetic is IEEE-754 binary64.  Plain pointers and sizes only; no torch types.
 * Functions returning int return 0 on success, a negative DCOL_E_* code for argument
 * errors, or a positive cudaError_t; dcol_last_error() describes the last failure of the
 * calling thread.  Per-pair solver outcomes are NOT errors: they are reported in status[].
 */
#ifndef DCOL_H_
#define DCOL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCOL_ABI_VERSION 1

/* primitive kinds (order of SURVEY.md section 8) */
enum {
    DCOL_POLYTOPE = 0, /* {y : A y <= alpha b}, f faces            problem_matrices.py:181-209 */
    DCOL_CAPSULE  = 1, /* radius R, segment length L (body x axis) problem_matrices.py:4-44    */
    DCOL_CYLINDER = 2, /* radius R, length L                       problem_matrices.py:47-87   */
    DCOL_CONE     = 3, /* height H, half angle beta                problem_matrices.py:125-148 */
    DCOL_SPHERE   = 4, /* radius R                                 problem_matrices.py:151-178 */
    DCOL_POLYGON  = 5, /* planar {A y <= alpha b} padded by R      problem_matrices.py:90-120  */
    DCOL_ELLIPSOID = 6, /* EXTENSION (absent from the reference's code, Report.pdf sec. 3.1.5 eq. 27):
                           semi-axes (R, L, H) along the body axes; ||diag(1/R,1/L,1/H) Q'^T (x - r')|| <= alpha */
    DCOL_N_KINDS  = 7
};

#define DCOL_MAX_FACES 32 /* most half-spaces of one polytope / polygon */
#define DCOL_MAX_ITER  50 /* the reference's hard-coded loop cap, pdip.py:408 */

/* per-pair status: the reference signals these by exception class */
enum {
    DCOL_STATUS_OK          = 0,
    DCOL_STATUS_MAX_ITER    = 1, /* Exception("Maximum number of iterations reached, PDIP failed"), pdip.py:470 */
    DCOL_STATUS_NON_FINITE  = 2, /* ValueError("array must not contain infs or NaNs") from SciPy check_finite  */
    DCOL_STATUS_NOT_PD      = 3, /* numpy.linalg.LinAlgError from a Cholesky factorisation                      */
    DCOL_STATUS_UNSUPPORTED = 4  /* ValueError from np.vstack: both primitives carry extra variables,
                                    combine_problem_matrices.py:58-67                                          */
};

/* argument errors */
enum {
    DCOL_E_ARG     = -1, /* null pointer / negative size / bad flag  */
    DCOL_E_SHAPE   = -2, /* malformed shape record                   */
    DCOL_E_INDEX   = -3, /* shape index out of range                 */
    DCOL_E_NOGPU   = -4  /* no CUDA device (there is no CPU fallback) */
};

/* output selection */
enum {
    DCOL_WANT_CONTACT = 1u, /* x[0:3] of the solution, proximity.py:52                        */
    DCOL_WANT_GRAD    = 2u, /* d alpha / d [r1 p1 r2 p2], proximity_gradient.py:71-77 layout  */
    DCOL_DEST_MULTICAST = 8u, /* dcol_proximity_batch_records only: dest[0] is an NVLink multicast address (NVLS);
                                 every record leaves as multimem.st and the switch delivers it to all ranks' buffers */
    DCOL_FIX_CASE4    = 4u, /* EXTENSION: also solve the pairs in which both primitives carry extra variables
                               (capsule / cylinder / polygon squared), with the column layout
                               [x, alpha, extras1, extras2] that combine_problem_matrices.py:58-67 builds for the
                               second primitive but forgets to pad the first to; without this flag those pairs
                               report DCOL_STATUS_UNSUPPORTED, as the reference raises ValueError */
    DCOL_ONE_PAIR_PER_THREAD = 16u, /* force the kernels in which a thread owns one pair from start to finish ...      */
    DCOL_LANE_REFILL  = 64u, /* ... or the lane-refill kernels (a warp owns a chunk of pairs; a lane whose pair has
                               converged takes the next pre-initialised pair from a shared-memory pool).  Neither
                               flag: the library's default (environment DCOL_REFILL=0/1 overrides it).  Same per-pair
                               operations either way; results agree to rounding (different kernels, different
                               fused-multiply-add contraction), with identical iteration counts and status words */
    DCOL_WANT_GRAD1   = 32u /* with DCOL_WANT_GRAD: grad is [B][6], only d alpha / d [r1 p1] — the half every caller of
                               the reference consumes (systems/piano_mover.py:94,
                               cluttered_hallway_quadrotor.py:161-163 use g[0:6] and drop g[6:12]); halves the
                               device->host bytes of the host entry points */
};

/* One primitive SHAPE (no pose).  144 bytes, natural alignment. */
typedef struct dcol_shape {
    int32_t type;        /* DCOL_POLYTOPE .. DCOL_POLYGON                              */
    int32_t n_faces;     /* polytope / polygon: number of half-spaces, otherwise 0     */
    int32_t face_off;    /* first face of this shape in the packed A / b arrays        */
    int32_t reserved;
    double  R, L, H, beta;
    double  r_offset[3]; /* body-frame offset of the primitive's origin                */
    double  Q_offset[9]; /* body-frame rotation offset, row-major                      */
} dcol_shape;

typedef struct dcol_shape_table dcol_shape_table; /* device-resident copy of a shape list */
typedef struct dcol_plan dcol_plan;               /* a batch's pairs grouped by type signature */

const char* dcol_version(void);
const char* dcol_last_error(void);
int         dcol_device_count(void);

/* Upload n_shapes records and the packed half-space arrays A[n_faces][3] (polygons use the
 * first two columns), b[n_faces] to `device`.  Host pointers. */
int  dcol_shape_table_create(const dcol_shape* shapes, int32_t n_shapes, const double* A, const double* b,
                             int32_t n_faces, int device, dcol_shape_table** out);
void dcol_shape_table_destroy(dcol_shape_table* table);

/* Group B pairs by type signature (kind1, kind2, faces1, faces2) with a device counting sort.
 * idx1 / idx2 are DEVICE pointers to int32 shape indices.  The plan can be reused for any
 * number of solves over the same index arrays (ALTRO re-evaluates a fixed set of
 * victim x obstacle pairs with new poses every iteration).  Synchronises `stream` once. */
int     dcol_plan_create(const dcol_shape_table* table, const int32_t* d_idx1, const int32_t* d_idx2, int64_t B,
                         void* stream, dcol_plan** out);
void    dcol_plan_destroy(dcol_plan* plan);
int64_t dcol_plan_size(const dcol_plan* plan);
int32_t dcol_plan_n_groups(const dcol_plan* plan);
/* Kernel launches one solve over this plan enqueues. */
int32_t dcol_plan_n_launches(const dcol_plan* plan);

/* (new) Re-order the plan for the NEXT solve of the same pair list: inside every group, pairs are sorted by
 * descending d_iters[pair] (DEVICE pointer, pair order: the iters[] array a previous solve of this plan wrote).  A warp runs
 * until its slowest pair converges; a caller that re-solves a fixed pair list with slowly changing poses (every
 * AL-iLQR pass re-evaluates the same victim x obstacle constraints, ALTRO.py:276-314) gets warps of equal iteration
 * count this way.  Results are unaffected (same pairs, same outputs at the same indices); only the order in which
 * threads take pairs changes.  Enqueues on `stream` without synchronising; calls on one plan (solve, refine) must be
 * stream-ordered; dcol_plan_perm() changes.  At most 1024 groups. */
int dcol_plan_refine(dcol_plan* plan, const int32_t* d_iters, void* stream);

/* Solve every pair of the plan.  All pointers are DEVICE pointers; the call only enqueues work
 * on `stream` (a cudaStream_t, may be NULL for the default stream).
 *   pose1, pose2 : [B][6] rows (r, p)
 *   alpha        : [B]
 *   contact      : [B][3]   or NULL unless DCOL_WANT_CONTACT
 *   grad         : [B][12]  or NULL unless DCOL_WANT_GRAD ([B][6] with DCOL_WANT_GRAD1)
 *   iters,status : [B] int32 (PDIP iterations taken; DCOL_STATUS_*)
 * tol is the reference's pdip_tol (1e-6 at both call sites); max_iter must be in 1..DCOL_MAX_ITER
 * (the reference always runs with 50). */
int dcol_proximity_batch_device(const dcol_plan* plan, const double* d_pose1, const double* d_pose2, double tol,
                                int32_t max_iter, uint32_t flags, double* d_alpha, double* d_contact,
                                double* d_grad, int32_t* d_iters, int32_t* d_status, void* stream);

/* EXTENSION (Report.pdf section 2.3 eq. 4-6; absent from the reference's code,
 * which only differentiates the frozen-(x, z) Lagrangian, proximity/proximity_gradient.py:8-88):
 * the same solve, plus the SOLUTION JACOBIAN
 *   d_jac : [B][4][12]   rows: contact point x, y, z and alpha;  columns: [r1 p1 r2 p2]
 * obtained by differentiating the relaxed KKT system at the returned iterate,
 *   dx = -(G^T W^-2 G)^-1 (dG^T z + G^T W^-2 (dG x - dh)),
 * as an adjoint solve (four right-hand sides e_k) with the Cholesky factor of the reduced KKT matrix at the
 * final iterate, inside the solve kernel.  Row 3 agrees with d_grad up to O(mu) (envelope theorem); rows 0-2 are
 * what contact-rich simulation needs.  NaN where status != 0.  d_contact / d_grad may be NULL without their flag. */
int dcol_proximity_batch_jacobian(const dcol_plan* plan, const double* d_pose1, const double* d_pose2, double tol,
                                  int32_t max_iter, uint32_t flags, double* d_alpha, double* d_contact,
                                  double* d_grad, double* d_jac, int32_t* d_iters, int32_t* d_status, void* stream);

/* (new) SCENE form of the host entry point: M poses of ONE victim shape against n_obs posed obstacles, i.e. every
 * (victim pose, obstacle) pair — what the reference's systems evaluate knot by knot and obstacle by obstacle
 * (systems/cluttered_hallway_quadrotor.py:127-133 for 1 - alpha, :155-163 for the gradient, of which only d alpha /
 * d(victim pose) = g[0:6] is kept), for all knots of all rollout / line-search candidates in one call.
 * HOST pointers.  Only M*6 + n_obs*6 doubles go to the device (the poses are broadcast into pairs there) and
 * 8 + 4 [+ 4] [+ 48] bytes per pair come back, instead of 104 in / 112 out per pair through dcol_proximity_batch_host.
 *   victim_pose   : [M][6] rows (r, p)          obstacle_shape : [n_obs] indices into the table
 *   obstacle_pose : [n_obs][6]
 *   alpha, status : [M][n_obs]                  iters : [M][n_obs] or NULL
 *   grad1         : [M][n_obs][6] = d alpha / d [r1 p1], or NULL (then no gradient is computed)
 * flags: DCOL_FIX_CASE4 | DCOL_ONE_PAIR_PER_THREAD | DCOL_LANE_REFILL.  Calls on one table are serialised. */
int dcol_proximity_scene_host(const dcol_shape_table* table, int32_t victim_shape, const double* victim_pose, int64_t M,
                              const int32_t* obstacle_shape, const double* obstacle_pose, int32_t n_obs, double tol,
                              int32_t max_iter, uint32_t flags, double* alpha, double* grad1, int32_t* iters,
                              int32_t* status);

/* Record mode, for multi-GPU use: every pair's results go out as ONE 112-byte record
 *     { double alpha; double grad[12]; int32 iters; int32 status; }
 * written in PLAN order (record i belongs to pair dcol_plan_perm()[i]) at dest[d] + 14 * (record_offset + i)
 * doubles, for each of the n_dest (1..DCOL_MAX_DEST) destinations.  A destination may be memory of a peer
 * GPU mapped with dcol_ipc_import: the kernel's epilogue then stores the records over NVLink, i.e. the
 * all-gather of the results (BASELINE.json north_star) is fused into the solve.  Destinations must be
 * 16-byte aligned.  d_contact ([B][3], pair order) is optional.  dest is a HOST array of device pointers. */
#define DCOL_MAX_DEST 8
#define DCOL_RECORD_WORDS 14
int dcol_proximity_batch_records(const dcol_plan* plan, const double* d_pose1, const double* d_pose2, double tol,
                                 int32_t max_iter, uint32_t flags /* DCOL_FIX_CASE4 | DCOL_DEST_MULTICAST */, int32_t n_dest,
                                 double* const* dest, int64_t record_offset, double* d_contact, void* stream);
/* Device pointer to the plan's permutation: perm[i] = index (in the caller's arrays) of the i-th pair in
 * plan order; valid until dcol_plan_destroy. */
const int32_t* dcol_plan_perm(const dcol_plan* plan);

/* Device buffers other processes of the node can map through CUDA IPC (64-byte handles), used for the
 * record buffers of the fused all-gather: rank r allocates, exports, every other rank imports. */
int  dcol_device_alloc(int device, size_t bytes, void** out);
void dcol_device_free(int device, void* p);
int  dcol_ipc_export(int device, void* dev_ptr, void* handle64);
int  dcol_ipc_import(int device, const void* handle64, void** out);
void dcol_ipc_close(int device, void* p);

/* Same computation with HOST buffers: the batch flows in chunks through copy-in, plan + solve and
 * copy-out streams (device scratch cached in the table), then the call synchronises.  This is the
 * call a reference-side binding makes.  Calls on one table are serialised. */
int dcol_proximity_batch_host(const dcol_shape_table* table, const int32_t* idx1, const int32_t* idx2,
                              const double* pose1, const double* pose2, int64_t B, double tol, int32_t max_iter,
                              uint32_t flags, double* alpha, double* contact, double* grad, int32_t* iters,
                              int32_t* status);

/* Page-locked host memory for the buffers handed to dcol_proximity_batch_host (pageable memory
 * works too, but its copies are staged and do not overlap the solve). */
int  dcol_host_alloc(size_t bytes, void** out);
void dcol_host_free(void* p);

/* Tables and plans take their device scratch from a process-wide cache inside the library and return it there when
 * they are destroyed (cudaMalloc / cudaFree cost milliseconds; a caller that builds a table per solve would pay them every
 * time).  At most 3 GiB stay cached.  This call hands everything that is cached back to the driver; buffers of live tables
 * and plans are not touched.  (No counterpart in the reference, which has no device memory.) */
void dcol_release_cached(void);

/* Debug aid: solve ONE pair and also return the per-iteration mu = s'z/deg trace
 * (mu_trace[DCOL_MAX_ITER + 1], NaN padded) and the final (x[8], s[72], z[72]).  Host pointers. */
int dcol_debug_trace_pair(const dcol_shape_table* table, int32_t idx1, int32_t idx2, const double* pose1,
                          const double* pose2, double tol, double* alpha, double* x, double* s, double* z,
                          int32_t* n, int32_t* m, int32_t* iters, int32_t* status, double* mu_trace);

/* Measured FP64 FMA throughput of `device` in FLOP/s (dependent-chain-free DFMA loop, all SMs),
 * the roofline denominator MEASURED_PEAKS.json lacks. */
int dcol_measure_fp64_peak(int device, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DCOL_H_ */
