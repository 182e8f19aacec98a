// cov_span_cta.cu -- the general span kernel of libcoverage_cuda (sm_100a): any swarm size
// (N <= 1024) on any grid (<= 32768^2), one CTA per candidate.
//
// Same arithmetic as the small-swarm kernel (cov_span_common.cuh; reference
// src/AreaCoverageCalculation.jl:63-110, src/TDM_STATIC_opt.jl:82-100, src/TDM_Constraints.jl:54-195
// of /root/reference), organised for candidates that are a lot of work each (50 UAVs on 1024^2:
// ~4 400 (disc, row) spans; 200 UAVs on 4096^2: ~70 000 spans of ~10 words):
//   * the CTA's 16 warps share ONE candidate: its 3N doubles are staged in shared memory, the disc
//     records are built by all threads, the O(N^2) separation test is spread over the warps while one
//     lane forms the order-dependent penalty sum;
//   * the union framebuffer lives in shared memory as a BAND of grid rows (the whole grid when it
//     fits: 1024 x 33 words = 132 KB; 427-row bands for 4096^2);
//   * per band the (disc, row) items are flattened over all 512 threads (prefix sums of the clipped
//     row counts in shared memory, warp-uniform binary search + per-lane linear advance), two items
//     per thread and iteration, atomicOr into the band, popcount of the newly set bits against the
//     fire planes (read through L1/L2 with ld.global.nc: the planes are shared by every CTA).
// Persistent grid (one CTA per SM), candidates handed out by an atomic counter.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include "cov_device.cuh"
#include "cov_kernel_common.cuh"
#include "cov_span_common.cuh"
#include "cov_kernels.cuh"
#include "../../include/coverage_cuda.h"

namespace cov {

constexpr int kCtaThreads = 512;

struct CtaPlan {
    int stage_bytes, dp_bytes, prefix_bytes, scratch_bytes, fb_bytes, band_rows, total_bytes;
};
__host__ __device__ inline CtaPlan cta_plan(const GridDesc &g, int N, int budget, int band_rows_opt)
{
    CtaPlan p;
    p.stage_bytes = round_up(3 * N * 8, 16);
    p.dp_bytes = N * 32;
    p.prefix_bytes = round_up((N + 1) * 4, 16);
    p.scratch_bytes = 1024;
    const int fixed = p.stage_bytes + p.dp_bytes + p.prefix_bytes + p.scratch_bytes;
    const int row_bytes = g.stride * 4;
    int rows = (budget - fixed) / row_bytes;
    if (rows > g.ny) rows = g.ny;
    if (band_rows_opt > 0 && band_rows_opt < rows) rows = band_rows_opt;
    p.band_rows = rows;
    p.fb_bytes = rows > 0 ? round_up(rows * row_bytes, 16) : 0;
    p.total_bytes = fixed + p.fb_bytes;
    return p;
}

template <bool MULTI>
__global__ void __launch_bounds__(kCtaThreads, 1)
span_cta_kernel(const __grid_constant__ GridDesc g, const __grid_constant__ ObjParams o,
                const double *__restrict__ X, long long B, EvalOut out, unsigned long long *counter,
                int force_exact, int band_rows, int fb_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = kCtaThreads / 32;
    const int N = o.N;
    const int cstride = 3 * N;
    double *stage = reinterpret_cast<double *>(smem_raw);
    SDisc *dp = reinterpret_cast<SDisc *>(smem_raw + round_up(3 * N * 8, 16));
    uint32_t *prefix = reinterpret_cast<uint32_t *>(smem_raw + round_up(3 * N * 8, 16) + N * 32);
    unsigned char *scratch = smem_raw + round_up(3 * N * 8, 16) + N * 32 + round_up((N + 1) * 4, 16);
    uint32_t *fb = reinterpret_cast<uint32_t *>(scratch + 1024);
    // scratch: [0,8) next candidate; [8,16) violation; [16,24) progressive; [64, 64+16*4*4) per-warp counts
    unsigned long long *s_next = reinterpret_cast<unsigned long long *>(scratch);
    double *s_viol = reinterpret_cast<double *>(scratch + 8);
    double *s_prog = reinterpret_cast<double *>(scratch + 16);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(scratch + 64);       // [nwarps][kMaxClasses]
    uint32_t *s_carry = reinterpret_cast<uint32_t *>(scratch + 64 + 256); // scan carry

    for (int t = tid; t < fb_bytes / 16; t += kCtaThreads) reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);

    for (;;) {
        __syncthreads(); // previous candidate fully retired (stage/dp/prefix/scratch reusable, fb clean)
        if (tid == 0) *s_next = atomicAdd(counter, 1ull);
        __syncthreads();
        const long long cand = (long long)*s_next;
        if (cand >= B) break;
        const double *xr = X + cand * cstride;

        // ---- A. stage the candidate ----
        for (int t = tid; t < cstride; t += kCtaThreads) stage[t] = __ldg(xr + t);
        __syncthreads();

        // ---- B. disc records (all threads) ----
        for (int c = tid; c < N; c += kCtaThreads) {
            SDisc d;
            make_sdisc(g, stage[c], stage[N + c], stage[2 * N + c], d);
            d.flags |= 2u; // large swarms overlap as a rule: every disc goes through the framebuffer
            dp[c] = d;
        }

        // ---- C. penalty (one lane, the reference's order) and constraints (everyone) ----
        if (tid == 0) {
            double viol = 0.0, prog = 0.0;
            for (int i = 0; i < N; ++i) {
                const double diff = __dsub_rn(stage[2 * N + i], o.r_max[i]);
                viol = __dadd_rn(viol, fabs(diff));
            }
            if (out.progressive)
                for (int i = 0; i < N; ++i)
                    prog = __dadd_rn(prog, julia_max0(__dsub_rn(stage[2 * N + i], o.r_max[i])));
            *s_viol = viol;
            *s_prog = prog;
        }
        bool bad = false;
        if (o.use_cons3) {
            for (int i = tid; i < N; i += kCtaThreads) {
                const double ax = __dsub_rn(o.prev_x[i], stage[i]);
                const double ay = __dsub_rn(o.prev_y[i], stage[N + i]);
                const double az = __dsub_rn(o.prev_z[i], __ddiv_rn(stage[2 * N + i], o.tan_half_fov));
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
                bad |= (s >= o.cons3_G[i]);
            }
        }
        if (o.use_cons7) {
            for (int i = tid; i < N; i += kCtaThreads) bad |= (stage[N + i] < 200.0) && (stage[2 * N + i] > o.cons7_R);
        }
        if (o.use_cons8) {
            // unordered pairs decide the reference's ordered-pair loop ((xi-xj)^2 is symmetric)
            for (int i = warp; i < N - 1; i += nwarps) {
                const double xi = stage[i], yi = stage[N + i];
                for (int j2 = i + 1 + lane; j2 < N; j2 += 32) {
                    const double ax = __dsub_rn(xi, stage[j2]);
                    const double ay = __dsub_rn(yi, stage[N + j2]);
                    bad |= (__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)) < o.sep_T);
                }
            }
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0); // also publishes dp[] and s_viol

        uint32_t cnt[MULTI ? kMaxClasses : 1];
#pragma unroll
        for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) cnt[k] = 0;
        long long cls_total[kMaxClasses];
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) cls_total[k] = 0;

        // ---- D. bands of framebuffer rows ----
        for (int jb0 = 1; jb0 <= g.ny; jb0 += band_rows) {
            const int jb1 = min(g.ny, jb0 + band_rows - 1);
            // rows of every disc clipped to the band -> inclusive prefix sums in prefix[1..N]
            if (warp == 0) {
                uint32_t carry = 0;
                for (int c0 = 0; c0 < N; c0 += 32) {
                    const int c = c0 + lane;
                    uint32_t n = 0;
                    if (c < N) {
                        const uint32_t rows = dp[c].rows;
                        const int r0 = max((int)(rows & 0xffffu), jb0), r1 = min((int)(rows >> 16), jb1);
                        n = r1 >= r0 ? (uint32_t)(r1 - r0 + 1) : 0u;
                    }
                    uint32_t incl = n;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += v;
                    }
                    if (c < N) prefix[c + 1] = carry + incl;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) {
                    prefix[0] = 0;
                    *s_carry = carry;
                }
            }
            __syncthreads();
            const uint32_t total = *s_carry;
            if (total == 0) { // uniform across the CTA
                __syncthreads(); // everyone has read s_carry before warp 0 rewrites it
                continue;
            }

            // (disc, row) items over all threads, two per thread and iteration
            for (uint32_t tb = 0; tb < total; tb += 2 * kCtaThreads) {
                const uint32_t t0 = tb + tid, t1 = t0 + kCtaThreads;
                const bool has0 = t0 < total, has1 = t1 < total;
                int c0 = 0, c1 = 0;
                {
                    // warp-uniform binary search for the disc of the warp's first item, then a short
                    // per-lane advance (32 consecutive items span few discs)
                    const uint32_t w0 = min(tb + (uint32_t)(warp * 32), total - 1);
                    const uint32_t w1 = min(w0 + kCtaThreads, total - 1);
                    int lo = 0, hi = N - 1; // largest c with prefix[c] <= w0
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (prefix[mid] <= w0) lo = mid;
                        else hi = mid - 1;
                    }
                    c0 = lo;
                    hi = N - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (prefix[mid] <= w1) lo = mid;
                        else hi = mid - 1;
                    }
                    c1 = lo;
                    const uint32_t q0 = has0 ? t0 : w0, q1 = has1 ? t1 : w1;
                    while (c0 + 1 < N && prefix[c0 + 1] <= q0) ++c0;
                    while (c1 + 1 < N && prefix[c1 + 1] <= q1) ++c1;
                }
                const SDisc d0 = dp[c0], d1 = dp[c1];
                const uint32_t q0 = has0 ? t0 : min(tb + (uint32_t)(warp * 32), total - 1);
                const uint32_t q1 = has1 ? t1 : min(min(tb + (uint32_t)(warp * 32), total - 1) + kCtaThreads, total - 1);
                const int j0 = max((int)(d0.rows & 0xffffu), jb0) + (int)(q0 - prefix[c0]);
                const int j1 = max((int)(d1.rows & 0xffffu), jb0) + (int)(q1 - prefix[c1]);
                int lo0, hi0, lo1, hi1;
                int st0 = fast_span(g, d0, j0, force_exact, lo0, hi0);
                int st1 = fast_span(g, d1, j1, force_exact, lo1, hi1);
                if (!has0) st0 = kEmpty;
                if (!has1) st1 = kEmpty;
                if (st0 == kSlow) {
                    slow_item(g, xr, N, c0, j0, (d0.flags & 1u) || force_exact, lo0, hi0);
                    if (lo0 <= hi0) st0 = kSpan;
                    else { st0 = kEmpty; lo0 = hi0 = 1; }
                }
                if (st1 == kSlow) {
                    slow_item(g, xr, N, c1, j1, (d1.flags & 1u) || force_exact, lo1, hi1);
                    if (lo1 <= hi1) st1 = kSpan;
                    else { st1 = kEmpty; lo1 = hi1 = 1; }
                }
                paint_span<MULTI, false>(g, fb, g.planes, j0, lo0, hi0, st0 == kSpan, true, cnt, jb0);
                paint_span<MULTI, false>(g, fb, g.planes, j1, lo1, hi1, st1 == kSpan, true, cnt, jb0);
            }
            __syncthreads();
            // clear the band for the next band / candidate
            const int used = (jb1 - jb0 + 1) * g.stride;
            for (int t = tid; t < (used + 3) / 4; t += kCtaThreads)
                reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) { // keep 32-bit partials far from overflow
                cls_total[k] += cnt[k];
                cnt[k] = 0;
            }
            __syncthreads();
        }

        // ---- E. reduce the counts, assemble, write ----
#pragma unroll
        for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) {
            // per-thread totals fit 32 bits per band but not necessarily overall: reduce as 64-bit halves
            unsigned long long v = (unsigned long long)cls_total[k];
            for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) reinterpret_cast<unsigned long long *>(scratch + 512)[warp * kMaxClasses + k] = v;
        }
        __syncthreads();
        if (tid == 0) {
            long long tot[kMaxClasses];
            long long total_cnt = 0;
            for (int k = 0; k < kMaxClasses; ++k) {
                tot[k] = 0;
                if (k < (MULTI ? kMaxClasses : 1))
                    for (int w = 0; w < nwarps; ++w)
                        tot[k] += (long long)reinterpret_cast<unsigned long long *>(scratch + 512)[w * kMaxClasses + k];
                total_cnt += tot[k];
            }
            out.obj[cand] = assemble_objective(g, o, tot, *s_viol);
            if (out.count) out.count[cand] = total_cnt;
            if (out.feasible) out.feasible[cand] = (unsigned char)(any_bad ? 0 : 1);
            if (out.progressive) out.progressive[cand] = *s_prog;
            if (out.class_count)
                for (int k = 0; k < g.n_classes; ++k) out.class_count[cand * g.n_classes + k] = tot[k];
        }
    }
    (void)s_cnt;
}

cudaError_t launch_span_cta(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                            long long B, const EvalOut &out, unsigned long long *counter, cudaStream_t stream,
                            LaunchInfo *info)
{
    const CtaPlan p = cta_plan(g, o.N, cfg.max_smem_optin, cfg.band_rows);
    if (p.band_rows < 1) return cudaErrorInvalidConfiguration;
    const bool multi = !(g.n_planes == 1 && g.n_classes == 1 && g.plane_mult[0] == 1);
    const int grid = (int)std::min<long long>(B, (long long)cfg.num_sms);
    if (info) {
        info->grid = grid;
        info->block = kCtaThreads;
        info->smem_bytes = p.total_bytes;
        info->band_rows = p.band_rows;
        info->planes_in_smem = 0;
    }
    cudaError_t err;
    if (multi) {
        err = cudaFuncSetAttribute(span_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.total_bytes);
        if (err != cudaSuccess) return err;
        span_cta_kernel<true><<<grid, kCtaThreads, p.total_bytes, stream>>>(g, o, dX, B, out, counter, cfg.force_exact,
                                                                             p.band_rows, p.fb_bytes);
    } else {
        err = cudaFuncSetAttribute(span_cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.total_bytes);
        if (err != cudaSuccess) return err;
        span_cta_kernel<false><<<grid, kCtaThreads, p.total_bytes, stream>>>(g, o, dX, B, out, counter, cfg.force_exact,
                                                                              p.band_rows, p.fb_bytes);
    }
    return cudaGetLastError();
}

} // namespace cov
