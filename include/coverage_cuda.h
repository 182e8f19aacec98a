/*
 * coverage_cuda.h -- C ABI of libcoverage_cuda, the B200 (sm_100a) batched coverage objective.
 *
 * The reference (Gabisanth/MaximumAreaCoverageOptimization.jl) is pure Julia and has NO FFI
 * boundary; its de-facto operator interface is what DirectSearch.jl accepts:
 *   SetObjective(p, f)            f(x::Vector{Float64})::Float64   src/TDM_STATIC_opt.jl:125
 *   AddExtremeConstraint(p, c)    c(x)::Bool                       src/TDM_STATIC_opt.jl:151-153
 *   createObjective(cells, N, r_max) -> f                          src/TDM_STATIC_opt.jl:82-100
 *   create_cons3(pre, FOV, d_lim)    -> c                          src/TDM_Constraints.jl:54-75
 * Each entry point below names the reference lines it replaces.  Julia binds them with `ccall`
 * (see INTEGRATION.md and maximumareacoverageoptimization.jl_b200/julia/CoverageCUDA.jl); the
 * tests bind them with Python ctypes.
 *
 * Conventions
 *  - plain C, no C++ types; sizes int64_t, reals double; every call returns COV_OK (0) or a
 *    negative cov_status, never throws; the message is at cov_last_error().
 *  - The library owns device buffers and streams.  Host pointers are caller-owned and borrowed
 *    only for the duration of the call; every call that takes host outputs is synchronous on
 *    return.  A handle is not thread-safe; distinct handles may be used from distinct threads
 *    (covers DirectSearch's SetMaxEvals threaded evaluation, src/TDM_STATIC_opt.jl:129).
 *  - Candidate layout: B x 3N doubles, candidate-major, each row [x_1..x_N, y_1..y_N, R_1..R_N]
 *    (src/AreaCoverageCalculation.jl:31,38-40).
 *  - Cell (i, j), i = 1..nx, j = 1..ny, has its centre at (i*dx - dx/2, j*dy - dy/2)
 *    (src/AreaCoverageCalculation.jl:16, src/DynamicArea.jl:65).  Per-cell arrays are indexed
 *    (i-1) + nx*(j-1), i.e. Julia's column-major grid[i, j].
 *  - There is no CPU fallback: without a CUDA device cov_create fails with COV_ERR_CUDA.
 */
#ifndef COVERAGE_CUDA_H
#define COVERAGE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COV_ABI_VERSION 3 /* 2: COV_OPT_PROGRESSIVE_INDEX, COV_KERNEL_ORDERED, cov_get_class_weights, cov_eval_batch_best;
                             3: cov_eval_batch_packed (COV_PACK_*) */

#if defined(__GNUC__)
#define COV_API __attribute__((visibility("default")))
#else
#define COV_API
#endif

typedef enum cov_status {
    COV_OK = 0,
    COV_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, N mismatch ...) */
    COV_ERR_STATE = -2,     /* grid or parameters not set yet */
    COV_ERR_CUDA = -3,      /* CUDA runtime error (message has the CUDA error string) */
    COV_ERR_OFF_LATTICE = -4, /* cov_set_points: a point is not a cell centre of the lattice */
    COV_ERR_LIMIT = -5,     /* a size exceeds what the kernels support (see cov_limits) */
    COV_ERR_NOMEM = -6
} cov_status;

/* Kernel selection for cov_set_option(COV_OPT_KERNEL). */
enum {
    COV_KERNEL_AUTO = 0,    /* row-span kernels (default): the small-swarm variant (N <= 8, framebuffer in shared
                               memory) for batches of a thousand or two candidates and more, else the general one */
    COV_KERNEL_SPAN = 1,    /* row-span kernels, small-swarm variant whenever it applies (from 128 candidates on) */
    COV_KERNEL_BRUTE = 2,   /* every cell against every disc, FP32 band + FP64 exact band cells */
    COV_KERNEL_EXACT = 3,   /* every cell against every disc in FP64 only (slow cross-check) */
    COV_KERNEL_SPAN_GENERAL = 4, /* force the general span kernel (one CTA per candidate, any N) */
    COV_KERNEL_ORDERED = 5  /* the reference's own loop over the point list in list order, FP64 only: the Float64 area is
                               the reference's whatever the weights. Chosen automatically (whatever this option says)
                               when cov_grid_info.area_exact == 0; selectable as a cross-check otherwise */
};

enum {
    COV_OPT_KERNEL = 1,         /* one of COV_KERNEL_* */
    COV_OPT_WARPS_PER_CTA = 2,  /* 0 = auto */
    COV_OPT_CTAS_PER_SM = 3,    /* 0 = auto */
    COV_OPT_BAND_ROWS = 4,      /* span kernel framebuffer band height in rows; 0 = auto */
    COV_OPT_FORCE_EXACT = 5,    /* 1: span/brute kernels skip the FP32 band and decide every edge in FP64 */
    COV_OPT_CHUNK = 6,          /* host-path pipeline chunk (candidates per H2D/launch/D2H slice); 0 = auto */
    COV_OPT_TRACE = 7,          /* 1: record a per-slice device timeline of every host-path call (cov_get_trace) */
    COV_OPT_PLANE_MODE = 9,     /* CTA kernel, how fire words are read: -1 auto (default), 0 through L2 only when
                                   the framebuffer atomic left new bits, 1 through L2 ahead of the atomics,
                                   2 staged in shared memory band by band with TMA bulk copies, 3 paint-then-sweep
                                   (spans only painted, then popc(framebuffer & staged plane) over the band), 4 the
                                   same with the plane read through L2 in the sweep (no staging: taller bands or a
                                   fourth CTA per SM; what auto picks unless the swarm is tiny for its grid) */
    COV_OPT_ZEROCOPY_OUT = 8,   /* 1 (default): the host path's kernels write their results straight into pinned host
                                   memory; 0: into device buffers, copied back slice by slice */
    COV_OPT_PROGRESSIVE_INDEX = 10 /* which progressive constraint the `progressive` output of cov_eval_batch_ex is:
                                   0 (default) cons1_progressive = sum over every UAV of max(R_i - r_max_i, 0.0)
                                   (src/TDM_Constraints.jl:182-195); k >= 1 the single-UAV forms, 1-based index as in
                                   the reference: 2 = cons2_progressive (:197-208), 3 = cons3_progressive (:210-221).
                                   k > N is rejected by the evaluation calls with COV_ERR_INVALID */
};

typedef struct cov_handle cov_handle;
typedef struct cov_multi cov_multi;

/* ---- lifecycle ------------------------------------------------------------------------- */
COV_API int cov_abi_version(void);
COV_API int cov_device_count(void);                       /* 0 when no CUDA device is usable */
COV_API int cov_create(int device, cov_handle **out);
COV_API void cov_destroy(cov_handle *h);
/* Message of the last failing call on this handle (h == NULL: last failing call of this thread
 * that had no handle, e.g. cov_create). Never NULL. */
COV_API const char *cov_last_error(const cov_handle *h);
COV_API int cov_set_option(cov_handle *h, int option, int64_t value);
COV_API int cov_get_option(const cov_handle *h, int option, int64_t *value);

/* ---- the cell store (replaces CellFunctions.Cells.points_of_interest as the kernel's input;
 *      src/CellFunctions.jl:5-16, producers src/AreaCoverageCalculation.jl:11-21,
 *      src/CellFunctions.jl:20-79) ----------------------------------------------------------- */

/* 1 bit per cell, every set cell is one list entry of weight `weight`.
 * bits: ny rows of ceil(nx/32) uint32 words; bit b of word w of row j-1 is cell i = 32*w + b + 1. */
COV_API int cov_set_grid_bits(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy,
                      const uint32_t *bits, double weight);

/* Per-cell multiplicity (how many list entries sit on the cell; the fire list repeats cells,
 * src/DynamicArea.jl:61-67) and optional per-cell weight class.
 * mult: nx*ny bytes; cls: nx*ny bytes or NULL (all class 0); class_weight: n_classes doubles. */
COV_API int cov_set_grid_cells(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy,
                       const uint8_t *mult, const uint8_t *cls, int64_t n_classes,
                       const double *class_weight);

/* The reference's own list: P entries [x, y, area, weight, covered] (P*5 doubles). Every point
 * must be a cell centre of the nx x ny lattice (bit-exactly i*dx - dx/2), else
 * COV_ERR_OFF_LATTICE; all entries on one cell must carry the same weight; distinct weights
 * become weight classes in order of first appearance. Entry 5 (the covered flag) is ignored as
 * in calculateArea (src/AreaCoverageCalculation.jl:69 is commented out). */
COV_API int cov_set_points(cov_handle *h, const double *pts5, int64_t P, int64_t nx, int64_t ny, double dx,
                   double dy);

/* createPOI(dx, dy, nx, ny) built on the device (src/AreaCoverageCalculation.jl:11-21). */
COV_API int cov_set_grid_full(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy);

/* Append list entries (update_POI, src/CellFunctions.jl:59-79): same rules as cov_set_points,
 * multiplicities add up. Limits of the store (also of cov_set_points): at most 4 distinct weights (weight classes),
 * one weight per cell, at most 255 entries per cell (COV_ERR_LIMIT / COV_ERR_INVALID otherwise). A failing append
 * leaves the store exactly as it was. */
COV_API int cov_add_points(cov_handle *h, const double *pts5, int64_t P);

typedef struct cov_grid_info {
    int64_t nx, ny;
    double dx, dy;
    int64_t n_entries;      /* list entries = sum of multiplicities */
    int64_t n_cells;        /* cells with multiplicity > 0 */
    int64_t n_planes;       /* bit planes the kernels sweep */
    int64_t n_classes;
    int32_t area_exact;     /* 1: weight*count is exactly representable for every reachable count,
                               so the returned Float64 area equals the reference's list-order sum */
    int32_t planes_in_smem; /* 1: the span kernel stages the planes in shared memory */
} cov_grid_info;
COV_API int cov_get_grid_info(const cov_handle *h, cov_grid_info *info);
/* The weight of every class as the device numbers them (class k of `class_count`, the order fixed when the
 * store was created; it does NOT follow later removals). Writes min(cap, n_classes) doubles. */
COV_API int cov_get_class_weights(const cov_handle *h, double *class_weight, int64_t cap);
/* Read the device-resident multiplicity plane back (nx*ny bytes). */
COV_API int cov_get_grid_cells(cov_handle *h, uint8_t *mult);

/* rmvCoveredPOI (src/CellFunctions.jl:81-108): delete every entry covered by the discs xyR
 * ([x;y;R], 3N doubles) from the device-resident cell store. removed (nullable) = entries deleted. */
COV_API int cov_remove_covered(cov_handle *h, const double *xyR, int64_t N, int64_t *removed);

/* ---- forest-fire cellular automaton on the device (replaces the offline generator
 *      src/DynamicArea.jl:17-86 and its xlsx hand-off :100-108 to CellFunctions.update_POI) ----------
 * state: nx*ny bytes indexed (i-1) + nx*(j-1) like grid[i, j]: 0 = EMPTY, 1 = TREE, 2 = FIRE.
 * cov_fire_init uploads it and re-creates the cell store on the same lattice (weight dx*dy); with
 * push_initial != 0 every burning cell becomes one list entry (the reference's initial_points).
 * cov_fire_step is one update_grid(): each interior TREE cell draws once per burning Moore
 * neighbour, `wind_speed*cos(wind_direction - atan(2-b, 2-a))*prob_spread > u`, u from
 * Philox4x32-10 keyed by seed with counter (cell, step, neighbour); every success pushes one list
 * entry (append != 0: straight into the device-resident store; multiplicities add up like the
 * duplicates of FirePoints.xlsx). n_pushed (nullable) = entries pushed by this step. */
COV_API int cov_fire_init(cov_handle *h, int64_t nx, int64_t ny, double dx, double dy, const uint8_t *state,
                          int32_t push_initial);
COV_API int cov_fire_step(cov_handle *h, double wind_speed, double wind_direction, double prob_spread,
                          uint64_t seed, int64_t step, int32_t append, int64_t *n_pushed);
COV_API int cov_fire_get_state(cov_handle *h, uint8_t *state);

/* ---- objective and constraint parameters (the closures' captured variables:
 *      createObjective(cells, N, r_max) src/TDM_STATIC_opt.jl:82, create_cons3(pre, FOV, d_lim)
 *      src/TDM_Constraints.jl:54, cons8's 15.0 :163, cons7 :142-154) --------------------------- */
COV_API int cov_set_params(cov_handle *h, int64_t N,
                   const double *r_max,        /* N */
                   double penalty_scale,       /* 1e5 in the reference */
                   const double *prev_xyR,     /* 3N [x;y;R] of the previous timestep, NULL: cons3 off */
                   const double *d_lim,        /* N (ignored when prev_xyR is NULL) */
                   double tan_half_fov,        /* tan(FOV/2) as the caller's libm computes it */
                   double sep_min,             /* cons8 threshold (15.0); <= 0: cons8 off */
                   int32_t use_cons7);         /* != 0: cons7 on */

/* ---- evaluation (replaces the per-trial-point loop of a MADS poll step, each iteration of which
 *      is cons_k(x) then AreaMaxObjective(x)) -------------------------------- */

/* Host buffers. obj[b] = -area + penalty_scale * sum|R_i - r_max_i| (src/TDM_STATIC_opt.jl:89-97);
 * count[b] (nullable) = covered list entries; feasible[b] (nullable) = 1 iff every enabled
 * extreme constraint holds. The objective is computed for infeasible candidates too; applying
 * the extreme barrier is the caller's choice (cov_argmin applies it). */
COV_API int cov_eval_batch(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                   uint8_t *feasible);

/* Extra per-candidate outputs (all nullable): class_count B x n_classes covered entries per
 * weight class; progressive B values of cons1_progressive (src/TDM_Constraints.jl:182-195). */
COV_API int cov_eval_batch_ex(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                      uint8_t *feasible, int64_t *class_count, double *progressive);

/* Device buffers, asynchronous on the handle's stream (no host copies, no synchronisation). */
COV_API int cov_eval_batch_device(cov_handle *h, const double *dX, int64_t B, double *d_obj,
                          int64_t *d_count, uint8_t *d_feasible);

/* The scalar closure: AreaMaxObjective(x) for one candidate (src/TDM_STATIC_opt.jl:83-98). */
COV_API int cov_eval_one(cov_handle *h, const double *x, double *obj);

/* Poll winner: smallest objective and its 0-based index. barrier != 0: infeasible candidates
 * count as +Inf (extreme barrier); if none is feasible best_idx = -1, best_obj = +Inf. */
COV_API int cov_argmin(cov_handle *h, const double *X, int64_t B, int32_t barrier, double *best_obj,
               int64_t *best_idx);

/* cov_eval_batch and the poll winner in one call: the per-candidate outputs as above (obj nullable here) plus
 * (best_obj, best_idx) reduced on the device -- the 16 bytes a rank contributes to the (min, index) exchange
 * when the candidates of one poll are sharded over several GPUs. */
COV_API int cov_eval_batch_best(cov_handle *h, const double *X, int64_t B, double *obj, int64_t *count,
                        uint8_t *feasible, int32_t barrier, double *best_obj, int64_t *best_idx);

/* The same evaluation for candidates that sit on a mesh, sent PACKED: the trial points of a MADS poll / search step
 * are x = q * granularity with integer q (granularity 1.0 on every variable in the reference,
 * src/TDM_STATIC_opt.jl:131-137), so 2 or 4 bytes per variable can cross PCIe instead of 8 -- the end-to-end rate
 * of cov_eval_batch is the PCIe rate. Q: B x 3N values in the candidate layout above, of type
 *   COV_PACK_I16  int16_t,  value = (double)q * granularity   (one correctly rounded FP64 multiply; exact
 *   COV_PACK_I32  int32_t,  value = (double)q * granularity    whenever granularity is a power of two)
 *   COV_PACK_F32  float,    value = (double)q                  (granularity ignored)
 * The device widens every slice into the Float64 matrix the kernels read, so the results are bit for bit those of
 * cov_eval_batch on the widened matrix. The caller vouches that its Float64 trial points ARE these values (integers
 * and FP32-representable numbers always are); nothing is rounded on the way in. obj / count / feasible as in
 * cov_eval_batch (obj may be NULL when a winner is asked for); best_obj and best_idx both NULL, or both given:
 * the poll winner as in cov_eval_batch_best. Pinned Q (cov_host_alloc) is DMA'd in place. */
enum { COV_PACK_F32 = 1, COV_PACK_I32 = 2, COV_PACK_I16 = 3 };
COV_API int cov_eval_batch_packed(cov_handle *h, const void *Q, int32_t pack, double granularity, int64_t B,
                          double *obj, int64_t *count, uint8_t *feasible, int32_t barrier, double *best_obj,
                          int64_t *best_idx);

/* A whole MADS solve in native code: the batch producer the reference lacks (DirectSearch.jl evaluates one
 * trial point per call). Settings of TDM_STATIC_opt.optimize (src/TDM_STATIC_opt.jl:118-222): start point x0
 * (3N), iteration limit n_iter (100 there), the same granularity on every variable (1.0 there), the enabled
 * constraints of cov_set_params as extreme barrier; x_out = the feasible incumbent if one was found, else x0.
 * Algorithm: granular-mesh MADS (Audet, Le Digabel & Tribes 2019), 2n Householder poll directions, complete
 * polling, one launch per poll set. stats (nullable): iterations, evaluations, batches, successes. */
COV_API int cov_mads_solve(cov_handle *h, const double *x0, int64_t n_iter, double granularity, uint64_t seed,
                           double *x_out, double *obj_out, int64_t *stats);

/* Covered-cell mask of ONE candidate: nx*ny bytes, 1 where the cell is inside some disc (whether
 * or not it holds entries). Lets the host replay an ordered Float64 sum over a weighted list. */
COV_API int cov_covered_mask(cov_handle *h, const double *x, uint8_t *mask);

/* ---- continuous variant: exact area of the union of the N discs of every candidate, by
 *      boundary integration in FP64 (Green's theorem; the reference ships only the pair primitives,
 *      src/Base_Functions.jl:230-355, and no driver, so this is an extension checked against closed forms and
 *      an independent restatement to 1e-9 relative, far inside north_star's 1e-5). No grid is involved.
 *      N <= 1024 (one warp per candidate up to 64 discs, one CTA per candidate above). Host buffers, synchronous. */
COV_API int cov_union_area_batch(cov_handle *h, const double *X, int64_t B, int64_t N, double *area);

/* ---- streams, memory, timing (so hosts without a CUDA binding can keep data resident) ------ */
COV_API int cov_sync(cov_handle *h);
COV_API void *cov_stream(cov_handle *h);                 /* the handle's cudaStream_t */
COV_API int cov_set_stream(cov_handle *h, void *stream); /* adopt a caller-owned cudaStream_t (NULL: back to own) */
COV_API int cov_host_alloc(cov_handle *h, int64_t bytes, void **out); /* pinned host memory */
COV_API int cov_host_free(cov_handle *h, void *p);
COV_API int cov_device_alloc(cov_handle *h, int64_t bytes, void **out);
COV_API int cov_device_free(cov_handle *h, void *p);
COV_API int cov_memcpy_h2d(cov_handle *h, void *dst, const void *src, int64_t bytes); /* async on the stream */
COV_API int cov_memcpy_d2h(cov_handle *h, void *dst, const void *src, int64_t bytes); /* async on the stream */
/* Kernel launches issued by this handle since creation (evidence for bench.py's gpu_launches). */
COV_API int64_t cov_launch_count(const cov_handle *h);
/* Timeline of the last host-path call recorded under COV_OPT_TRACE: per slice the device times (ms since
 * the call's first copy was queued) of: H2D done, kernel start, kernel end, D2H done. Returns the number of
 * values available; copies at most cap of them. */
COV_API int64_t cov_get_trace(const cov_handle *h, double *ms, int64_t cap);
/* Device time of the last cov_eval_batch* coverage-kernel launch(es) in milliseconds, measured
 * with CUDA events on the handle's stream (synchronises). */
COV_API int cov_last_kernel_ms(cov_handle *h, double *ms);
/* Shape of the last coverage-kernel launch of this handle (diagnostics; which kernel COV_KERNEL_AUTO and
 * COV_KERNEL_SPAN resolved to). COV_ERR_STATE before the first launch. */
typedef struct cov_launch_info {
    int32_t kernel;         /* cov_kernel: SPAN = small-swarm kernel, SPAN_GENERAL = CTA per candidate, BRUTE, EXACT */
    int32_t grid, block;    /* CTAs, threads per CTA */
    int32_t smem_bytes;     /* dynamic shared memory per CTA */
    int32_t band_rows;      /* framebuffer rows per band (ny: a single band) */
    int32_t planes_in_smem; /* 1: bit planes staged in shared memory */
    /* the template instantiation that ran (ABI 2; names the ncu profile bench.py's issue roofline reads) */
    int32_t multi;          /* 1: several bit planes / weight classes / multiplicities */
    int32_t chunk;          /* small-swarm kernel: candidates per work unit; CTA kernel: co-resident CTAs per SM the
                               instantiation is compiled for (3 or 4); else 0 */
    int32_t max_warps;      /* small-swarm kernel: warps per CTA the instantiation allows (20 or 24); else 0 */
    int32_t plane_mode;     /* CTA kernel: COV_OPT_PLANE_MODE it resolved to (0, 1, 2); else -1 */
} cov_launch_info;
COV_API int cov_last_launch(const cov_handle *h, cov_launch_info *out);
/* Running totals over every coverage-kernel launch of this handle: summed device time (CUDA
 * events on the launching stream) and number of launches (synchronises). Both nullable. */
COV_API int cov_kernel_time_total(cov_handle *h, double *ms, int64_t *launches);

/* Candidates ~ the benchmark distribution of BASELINE.md generated on the device:
 * x, y ~ U(0, lx/ly), h ~ U(h_min, h_max), R = h * tan_half_fov, Philox4x32-10 keyed by
 * (seed, candidate index, uav index). dX: B x 3N doubles on the device. */
COV_API int cov_generate_candidates(cov_handle *h, double *dX, int64_t B, int64_t N, uint64_t seed,
                            int64_t first_index, double lx, double ly, double h_min, double h_max,
                            double tan_half_fov);

/* ---- several GPUs of one box: candidates sharded contiguously, grid replicated ------------- */
COV_API int cov_multi_create(const int *devices, int n, cov_multi **out);
COV_API void cov_multi_destroy(cov_multi *m);
COV_API const char *cov_multi_last_error(const cov_multi *m);
COV_API int cov_multi_size(const cov_multi *m);
COV_API cov_handle *cov_multi_handle(cov_multi *m, int k); /* borrow shard k's handle (set grid/params on each) */
COV_API int cov_multi_eval_batch(cov_multi *m, const double *X, int64_t B, double *obj, int64_t *count,
                         uint8_t *feasible);
COV_API int cov_multi_argmin(cov_multi *m, const double *X, int64_t B, int32_t barrier, double *best_obj,
                     int64_t *best_idx);

/* ---- limits ---------------------------------------------------------------------------- */
typedef struct cov_limits {
    int64_t max_uavs;     /* N */
    int64_t max_nx, max_ny;
    int64_t max_planes;
    int64_t max_classes;
} cov_limits;
COV_API void cov_get_limits(cov_limits *out);

/* T(R): the smallest double t with sqrt(t) >= R, computed by the same closed form the kernels
 * use (exposed so the tests can check it against the definition). */
COV_API double cov_threshold(double R);

#ifdef __cplusplus
}
#endif
#endif /* COVERAGE_CUDA_H */
