/* poll_demo.c -- libcoverage_cuda from plain C: the reference's default problem
 * (FullSimulation.jl:727-804: 5 UAVs, createPOI(5,5,100,100), r_max = 30*tan(50 deg)), the scalar
 * closure on the start configuration (KAT-3 of SURVEY.md 8c), then one poll set evaluated as a batch.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/poll_demo.c -o poll_demo \
 *       -Lmaximumareacoverageoptimization.jl_b200 -lcoverage_cuda -lm -Wl,-rpath,'$ORIGIN/maximumareacoverageoptimization.jl_b200'
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "coverage_cuda.h"

#define N 5
#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != COV_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cov_last_error(h));  \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(void)
{
    cov_handle *h = NULL;
    const double pi = 3.141592653589793, FOV = 100.0 / 180.0 * pi, t = tan(FOV / 2);
    double r_max[N], x0[3 * N], obj0 = 0;
    int i, b;
    if (cov_create(0, &h) != COV_OK) {
        fprintf(stderr, "cov_create: %s\n", cov_last_error(NULL));
        return 2;
    }
    CHECK(cov_set_grid_full(h, 100, 100, 5.0, 5.0));
    for (i = 0; i < N; ++i) r_max[i] = 30.0 * t;
    for (i = 0; i < N; ++i) { /* allocate_even_circles(15.0, 5, 10*tan(FOV/2), 250, 250) */
        const double ang = 2 * pi / N * i;
        x0[i] = 15.0 * cos(ang) + 250.0;
        x0[N + i] = 15.0 * sin(ang) + 250.0;
        x0[2 * N + i] = 10.0 * t;
    }
    CHECK(cov_set_params(h, N, r_max, 1e5, x0, (double[N]){10, 10, 10, 10, 10}, t, 0.0, 0));
    CHECK(cov_eval_one(h, x0, &obj0));
    printf("AreaMaxObjective(start) = %.17g (KAT-3: 11915685.925942099)\n", obj0);
    {
        enum { B = 30 };
        double X[B * 3 * N], obj[B], best = 0;
        int64_t count[B], idx = -1;
        uint8_t feas[B];
        srand(1);
        for (b = 0; b < B; ++b)
            for (i = 0; i < 3 * N; ++i) X[b * 3 * N + i] = floor(x0[i] + (rand() % 13) - 6 + 0.5);
        CHECK(cov_eval_batch(h, X, B, obj, count, feas));
        CHECK(cov_argmin(h, X, B, 1, &best, &idx));
        for (b = 0; b < 3; ++b)
            printf("candidate %d: obj %.6f count %lld feasible %d\n", b, obj[b], (long long)count[b], feas[b]);
        printf("poll winner with the extreme barrier: index %lld, objective %.6f\n", (long long)idx, best);
    }
    cov_destroy(h);
    return fabs(obj0 - 11915685.925942099) < 1e-6 ? 0 : 3;
}
