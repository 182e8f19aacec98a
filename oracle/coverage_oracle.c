/*
 * coverage_oracle.c -- CPU restatement of the reference's coverage-objective hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file's shared object; libcoverage_cuda never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference (Gabisanth/MaximumAreaCoverageOptimization.jl) is pure Julia,
 * Julia is not installed here, and the reference's only test (test/runtests.jl:1-6) does not touch
 * this path, so there are no golden vectors from the reference itself.  This file restates the
 * cited Julia lines one for one (Float64, strict comparisons, same loop and summation order,
 * compiled with -ffp-contract=off so no FMA is formed) and is cross-checked against an
 * independently written NumPy restatement (oracle/coverage_oracle.py) and against the
 * known-answer vectors of SURVEY.md section 8c (tests/golden/).
 *
 * All `file:line` citations are relative to /root/reference/.
 *
 * Point-list layout (the reference's own): P entries of 5 doubles
 *   [x, y, area represented, weight ("importance"), covered flag]
 * (src/AreaCoverageCalculation.jl:16, src/DynamicArea.jl:65), flattened row-major to P*5 doubles.
 * Candidate layout: 3N doubles in SoA order [x_1..x_N, y_1..y_N, R_1..R_N]
 * (src/AreaCoverageCalculation.jl:31,38-40).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* src/AreaCoverageCalculation.jl:11-21  createPOI(dx, dy, x_length, y_length)
 * i outer (x), j inner (y); centre (i*dx - dx/2, j*dy - dy/2); area = weight = dx*dy; flag false.
 * `for i in 1:x_length` over a Float64 range visits 1.0, 2.0, ... <= x_length.
 * Returns the number of points written (out may be NULL to query the count). */
ORC_API int64_t orc_createPOI(double dx, double dy, double x_length, double y_length, double *out)
{
    int64_t n = 0;
    for (double i = 1.0; i <= x_length; i += 1.0) {
        for (double j = 1.0; j <= y_length; j += 1.0) {
            if (out) {
                double *p = out + 5 * n;
                p[0] = i * dx - dx / 2;
                p[1] = j * dy - dy / 2;
                p[2] = dx * dy;
                p[3] = dx * dy;
                p[4] = 0.0;
            }
            ++n;
        }
    }
    return n;
}

/* The predicate shared by calculateArea and rmvCoveredPOI
 * (src/AreaCoverageCalculation.jl:70,121; src/CellFunctions.jl:90):
 *   sqrt((px - cx)^2 + (py - cy)^2) < R        Float64, strict <, `^2` lowers to x*x, no FMA. */
static inline int orc_covered(double px, double py, double cx, double cy, double R)
{
    double ddx = px - cx;
    double ddy = py - cy;
    double s = ddx * ddx + ddy * ddy;
    return sqrt(s) < R;
}

/* src/AreaCoverageCalculation.jl:63-110  calculateArea(circles, points)
 * For each point in list order, the first disc that covers it adds points[p][4] (the weight) to a
 * Float64 running sum and breaks.  count_out = number of covered list entries; tests_out = number
 * of predicate evaluations actually executed (early break included). */
ORC_API double orc_calculateArea(const double *circles, int64_t N, const double *pts5, int64_t P,
                                 int64_t *count_out, int64_t *tests_out)
{
    double area_covered = 0.0;
    int64_t count = 0, tests = 0;
    for (int64_t p = 0; p < P; ++p) {
        const double *pt = pts5 + 5 * p;
        for (int64_t c = 0; c < N; ++c) {
            ++tests;
            if (orc_covered(pt[0], pt[1], circles[c], circles[N + c], circles[2 * N + c])) {
                area_covered += pt[3];
                ++count;
                break;
            }
        }
    }
    if (count_out) *count_out = count;
    if (tests_out) *tests_out = tests;
    return area_covered;
}

/* src/TDM_STATIC_opt.jl:82-100 (duplicate src/TDM_Constraints.jl:33-51)  AreaMaxObjective(x)
 * make_circles/make_MADS (src/AreaCoverageCalculation.jl:33-59) are a value-identity round trip.
 *   violation = sum_{i=1..N} abs(x[i+2N] - r_max[i])   sequential Float64
 *   return -area_covered + violation*1e5 */
ORC_API double orc_objective(const double *x, int64_t N, const double *r_max, const double *pts5,
                             int64_t P, int64_t *count_out)
{
    double area_covered = orc_calculateArea(x, N, pts5, P, count_out, NULL);
    double violation = 0.0;
    for (int64_t i = 0; i < N; ++i) violation += fabs(x[i + 2 * N] - r_max[i]);
    return -area_covered + violation * 1e5;
}

/* src/TDM_Constraints.jl:54-75  create_cons3(pre, FOV, d_lim) -> cons3(x)
 * z = R / tan(FOV/2); infeasible when sqrt(dx^2 + dy^2 + dz^2) > d_lim[i] (strict >).
 * `pre` is given in the same SoA [x;y;R] layout.  tan_half_fov is passed in by the caller so the
 * caller's own tan() decides the last ulp. Returns 1 = feasible. */
ORC_API int orc_cons3(const double *x, int64_t N, const double *pre, double tan_half_fov,
                      const double *d_lim)
{
    for (int64_t i = 0; i < N; ++i) {
        double x1 = pre[i], y1 = pre[N + i], z1 = pre[2 * N + i] / tan_half_fov;
        double x2 = x[i], y2 = x[N + i], z2 = x[2 * N + i] / tan_half_fov;
        double ax = x1 - x2, ay = y1 - y2, az = z1 - z2;
        if (sqrt(ax * ax + ay * ay + az * az) > d_lim[i]) return 0;
    }
    return 1;
}

/* src/TDM_Constraints.jl:142-154  cons7(x): if y < 200 then R must be <= 19*tan(FOV/2). */
ORC_API int orc_cons7(const double *x, int64_t N, double tan_half_fov)
{
    for (int64_t i = 0; i < N; ++i)
        if (x[i + N] < 200)
            if (x[i + 2 * N] > 19 * tan_half_fov) return 0;
    return 1;
}

/* src/TDM_Constraints.jl:157-172  cons8(x): every ordered pair i != j must have
 * sqrt((xi-xj)^2 + (yi-yj)^2) >= sep (reference hard-codes sep = 15.0; strict < rejects). */
ORC_API int orc_cons8(const double *x, int64_t N, double sep)
{
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < N; ++j)
            if (j != i) {
                double ax = x[i] - x[j], ay = x[i + N] - x[j + N];
                double hor_separation = sqrt(ax * ax + ay * ay);
                if (hor_separation < sep) return 0;
            }
    return 1;
}

/* Julia's max(v, 0.0) on Float64: NaN propagates (C's fmax would return 0.0 for a NaN), and
 * max(-0.0, 0.0) is +0.0. */
static inline double orc_julia_max0(double v) { return (v != v) ? v : (v > 0.0 ? v : 0.0); }

/* src/TDM_Constraints.jl:182-195  cons1_progressive(x) = sum max(R_i - r_max_i, 0.0).
 * (`violation = 0` starts as an Int and is promoted on the first add; 0 + v is exact.) */
ORC_API double orc_cons1_progressive(const double *x, int64_t N, const double *r_max)
{
    double violation = 0;
    for (int64_t i = 0; i < N; ++i) {
        double R_val = x[2 * N + i];
        violation += orc_julia_max0(R_val - r_max[i]);
    }
    return violation;
}

/* src/TDM_Constraints.jl:197-221  cons2_progressive / cons3_progressive: the same term for one
 * fixed UAV index (i = 2, i = 3 in the reference; 1-based `which` here). */
ORC_API double orc_consK_progressive(const double *x, int64_t N, const double *r_max, int64_t which)
{
    double violation = 0;
    double R_val = x[2 * N + (which - 1)];
    violation += orc_julia_max0(R_val - r_max[which - 1]);
    return violation;
}

/* src/CellFunctions.jl:81-108 (twin: src/AreaCoverageCalculation.jl:113-137)  rmvCoveredPOI
 * Deletes every list entry covered by any disc (same predicate), keeping list order.
 * In-place compaction; returns the new number of points. */
ORC_API int64_t orc_rmvCoveredPOI(const double *circles, int64_t N, double *pts5, int64_t P)
{
    int64_t keep = 0;
    for (int64_t p = 0; p < P; ++p) {
        const double *pt = pts5 + 5 * p;
        int del = 0;
        for (int64_t c = 0; c < N; ++c)
            if (orc_covered(pt[0], pt[1], circles[c], circles[N + c], circles[2 * N + c])) {
                del = 1;
                break;
            }
        if (!del) {
            if (keep != p) memmove(pts5 + 5 * keep, pt, 5 * sizeof(double));
            ++keep;
        }
    }
    return keep;
}

/* src/Base_Functions.jl:44-65  allocate_even_circles(r_centering_cir, N, r_uav, center_x, center_y)
 * out = [x;y;R].  (cos/sin come from the C library here, from Julia's libm in the reference; the
 * result is an INPUT of the hot path, so a last-ulp difference would move the input, not the
 * arithmetic under test.) */
ORC_API void orc_allocate_even_circles(double r_centering_cir, int64_t N, double r_uav,
                                       double center_x, double center_y, double *out)
{
    const double pi = 3.141592653589793;
    for (int64_t i = 1; i <= N; ++i) {
        double ref_angle = 2 * pi / (double)N * (double)(i - 1);
        out[i - 1] = r_centering_cir * cos(ref_angle) + center_x;
        out[N + i - 1] = r_centering_cir * sin(ref_angle) + center_y;
        out[2 * N + i - 1] = r_uav;
    }
}

/* Batched evaluation = the loop a MADS poll step runs over its trial points
 * (SURVEY.md 3.1): per candidate the extreme constraints, then the objective.  Unlike the
 * extreme barrier the objective is evaluated for every candidate, feasible or not, so the batch
 * outputs can be compared element for element.
 *   X        B x 3N, candidate-major
 *   pre      3N or NULL (cons3 off);  d_lim N
 *   sep_min  <= 0 -> cons8 off;  use_cons7 != 0 -> cons7 on
 *   obj B; count B (nullable); feasible B (nullable); prog B (nullable, cons1_progressive)
 * POSIX threads over candidates only (mirrors the reference's "threads on MADS evaluation" study,
 * src/MADS_runtime_comparison_Parallelisation.xlsx); never inside one candidate's sum.
 * threads <= 0: one per online core. */
typedef struct orc_batch_job {
    const double *X;
    int64_t B, N;
    const double *r_max, *pts5;
    int64_t P;
    const double *pre, *d_lim;
    double tan_half_fov, sep_min;
    int use_cons7;
    double *obj;
    int64_t *count;
    uint8_t *feasible;
    double *prog;
    int64_t next; /* work dispenser, 16 candidates at a time */
} orc_batch_job;

static void orc_eval_range(const orc_batch_job *J, int64_t b0, int64_t b1)
{
    const int64_t N = J->N;
    for (int64_t b = b0; b < b1; ++b) {
        const double *x = J->X + 3 * N * b;
        int64_t cnt = 0;
        J->obj[b] = orc_objective(x, N, J->r_max, J->pts5, J->P, &cnt);
        if (J->count) J->count[b] = cnt;
        if (J->feasible) {
            int ok = 1;
            if (J->pre) ok = ok && orc_cons3(x, N, J->pre, J->tan_half_fov, J->d_lim);
            if (J->sep_min > 0) ok = ok && orc_cons8(x, N, J->sep_min);
            if (J->use_cons7) ok = ok && orc_cons7(x, N, J->tan_half_fov);
            J->feasible[b] = (uint8_t)ok;
        }
        if (J->prog) J->prog[b] = orc_cons1_progressive(x, N, J->r_max);
    }
}

static void *orc_batch_worker(void *arg)
{
    orc_batch_job *J = (orc_batch_job *)arg;
    for (;;) {
        int64_t b0 = __atomic_fetch_add(&J->next, 16, __ATOMIC_RELAXED);
        if (b0 >= J->B) break;
        orc_eval_range(J, b0, b0 + 16 < J->B ? b0 + 16 : J->B);
    }
    return NULL;
}

ORC_API int orc_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

ORC_API void orc_eval_batch(const double *X, int64_t B, int64_t N, const double *r_max,
                            const double *pts5, int64_t P, const double *pre, const double *d_lim,
                            double tan_half_fov, double sep_min, int use_cons7, double *obj,
                            int64_t *count, uint8_t *feasible, double *prog, int threads)
{
    orc_batch_job J = {X, B, N, r_max, pts5, P, pre, d_lim, tan_half_fov, sep_min, use_cons7,
                       obj, count, feasible, prog, 0};
    if (threads <= 0) threads = orc_num_threads();
    if (threads > 1024) threads = 1024;
    if (threads == 1 || B <= 16) {
        orc_eval_range(&J, 0, B);
        return;
    }
    pthread_t tid[1024];
    int started = 0;
    for (int t = 0; t < threads - 1; ++t)
        if (pthread_create(&tid[started], NULL, orc_batch_worker, &J) == 0) ++started;
    orc_batch_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

/* Per-class covered-entry counts for a weighted list: class_of[p] in [0, n_classes).
 * Used to check the library's per-class integer counts on lists with non-uniform weights
 * (src/CellFunctions.jl:42-45,68-72 high-interest weights). */
ORC_API void orc_class_counts(const double *circles, int64_t N, const double *pts5, int64_t P,
                              const int32_t *class_of, int64_t n_classes, int64_t *counts)
{
    for (int64_t k = 0; k < n_classes; ++k) counts[k] = 0;
    for (int64_t p = 0; p < P; ++p) {
        const double *pt = pts5 + 5 * p;
        for (int64_t c = 0; c < N; ++c)
            if (orc_covered(pt[0], pt[1], circles[c], circles[N + c], circles[2 * N + c])) {
                counts[class_of ? class_of[p] : 0] += 1;
                break;
            }
    }
}

/* src/DynamicArea.jl:52-72  update_grid(grid): one step of the forest-fire cellular automaton.
 * grid is nx x ny, column-major like the Julia matrix (grid[i,j] at i-1 + nx*(j-1)), values
 * EMPTY=0, TREE=1, FIRE=2.  For each interior TREE cell with a burning Moore neighbour, for EACH
 * burning neighbour (window index (a,b), a,b in 1..3, iterated column-major as findall does):
 *   wind_speed*cos(wind_direction - atan(2-b, 2-a))*prob_spread > rand()  -> FIRE, push point.
 * The push sits inside the per-neighbour loop, so one cell can be pushed several times in a step.
 * The reference draws rand() from Julia's global RNG, which cannot be reproduced; here the
 * uniform for (step, i, j, a, b) is supplied by the caller through `u01` (a callback), so that a
 * counter-based generator gives the host and the device the same draws.
 * new_grid receives the next state; pts5 receives the pushed points (i*dx-dx/2, j*dy-dy/2,
 * dx*dy, dx*dy, false); returns the number pushed (pts5 may be NULL to count). */
typedef double (*orc_u01_fn)(void *ctx, int64_t step, int64_t i, int64_t j, int a, int b);

ORC_API int64_t orc_fire_update_grid(const uint8_t *grid, uint8_t *new_grid, int64_t nx, int64_t ny,
                                     double dx, double dy, double wind_speed, double wind_direction,
                                     double prob_spread, int64_t step, orc_u01_fn u01, void *ctx,
                                     double *pts5)
{
    int64_t n = 0;
    memcpy(new_grid, grid, (size_t)(nx * ny));
    /* Julia `for i in 2:nx-1, j in 2:ny-1` iterates j fastest. */
    for (int64_t i = 2; i <= nx - 1; ++i)
        for (int64_t j = 2; j <= ny - 1; ++j) {
            if (grid[(i - 1) + nx * (j - 1)] != 1) continue;
            /* findall over the 3x3 window is column-major: a (rows, i direction) fastest. */
            for (int b = 1; b <= 3; ++b)
                for (int a = 1; a <= 3; ++a) {
                    int64_t ii = i - 2 + a, jj = j - 2 + b;
                    if (grid[(ii - 1) + nx * (jj - 1)] != 2) continue;
                    double p = wind_speed * cos(wind_direction - atan2((double)(2 - b), (double)(2 - a))) *
                               prob_spread;
                    if (p > u01(ctx, step, i, j, a, b)) {
                        new_grid[(i - 1) + nx * (j - 1)] = 2;
                        if (pts5) {
                            double *q = pts5 + 5 * n;
                            q[0] = (double)i * dx - dx / 2;
                            q[1] = (double)j * dy - dy / 2;
                            q[2] = dx * dy;
                            q[3] = dx * dy;
                            q[4] = 0.0;
                        }
                        ++n;
                    }
                }
        }
    return n;
}

/* Counter-based uniform for the fire automaton: Philox4x32-10 (Salmon et al., SC'11) with counter
 * (cell index low 32, step, neighbour index k = (a-1) + 3(b-1), cell index high 32) and key = seed;
 * 53 bits -> [0, 1).  The device kernel (cov_grid_kernels.cu fire_step_kernel) draws the same. */
static void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                              uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct orc_philox_ctx {
    uint64_t seed;
    int64_t nx;
} orc_philox_ctx;

static double orc_philox_u01(void *vctx, int64_t step, int64_t i, int64_t j, int a, int b)
{
    const orc_philox_ctx *ctx = (const orc_philox_ctx *)vctx;
    const uint64_t cell = (uint64_t)((i - 1) + ctx->nx * (j - 1));
    uint32_t r[4];
    orc_philox4x32_10((uint32_t)cell, (uint32_t)step, (uint32_t)((a - 1) + 3 * (b - 1)), (uint32_t)(cell >> 32),
                      (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32), r);
    const uint64_t v = ((uint64_t)(r[0] >> 5) << 26) | (uint64_t)(r[1] >> 6);
    return (double)v * 1.1102230246251565e-16; /* 2^-53 */
}

/* One update_grid() step (src/DynamicArea.jl:52-72) with the Philox uniforms above. */
ORC_API int64_t orc_fire_step_philox(const uint8_t *grid, uint8_t *new_grid, int64_t nx, int64_t ny, double dx,
                                     double dy, double wind_speed, double wind_direction, double prob_spread,
                                     int64_t step, uint64_t seed, double *pts5)
{
    orc_philox_ctx ctx = {seed, nx};
    return orc_fire_update_grid(grid, new_grid, nx, ny, dx, dy, wind_speed, wind_direction, prob_spread, step,
                                orc_philox_u01, &ctx, pts5);
}

/* Threshold identity used by the CUDA kernels, stated here by its DEFINITION so tests can check
 * the device's closed form against it:  T(R) = the smallest double t (possibly +Inf) with
 * sqrt(t) >= R, so that for every double s >= 0:  sqrt(s) < R  <=>  s < T(R)
 * (correctly rounded sqrt is monotone).  Found by stepping from R*R with nextafter. */
ORC_API double orc_threshold_by_search(double R)
{
    if (!(R > 0)) return 0.0; /* sqrt(s) >= 0 >= R, NaN compares false: never covered */
    if (isinf(R)) return INFINITY;
    double t = R * R;
    if (isinf(t)) {
        t = 1.7976931348623157e308;
        if (sqrt(t) < R) return INFINITY;
    }
    while (t > 0 && sqrt(t) >= R) t = nextafter(t, -INFINITY);
    /* now sqrt(t) < R (or t == 0) */
    while (sqrt(t) < R) t = nextafter(t, INFINITY);
    return t;
}

