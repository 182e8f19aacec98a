// cov_span_small.cu -- the small-swarm span kernel of libcoverage_cuda (sm_100a): N <= 8 UAVs and a
// grid whose union framebuffer fits a warp's share of shared memory (the reference's own
// configurations: N = 5 on 100 x 100, and BASELINE's 5 x 1M on 256 x 256).
//
// Same result as the general span kernel (cov_kernels.cu) -- the count of list entries inside the
// union of the discs, src/AreaCoverageCalculation.jl:63-110 of /root/reference, plus the objective
// and constraint outputs -- organised for short candidates:
//   phase 1  lane k of a warp owns candidate k of a 32- (or 16-) candidate chunk: it reads its
//            3N doubles, forms the penalty sums and constraint verdicts serially in FP64 (the
//            reference's own order) and writes one 32-byte record per disc to shared memory.
//            32 candidates are set up at once instead of one candidate on N of 32 lanes.
//   phase 2  per candidate, the (disc, row) pairs are FLATTENED over the lanes.  A lane computes
//            the covered columns [lo, hi] of its row in FP32, in coordinates relative to the
//            disc's nearest cell (so the FP32 error is ~2^-23 R instead of ~2^-21 * 500 m),
//            certifies both ends against an error band, and falls back to the exact FP64 walk
//            only when the band cannot decide.  The interval is OR-ed into the warp's
//            shared-memory framebuffer with atomicOr; the bits the lane was first to set are
//            AND-ed with the fire plane words and popcounted: a union count, whatever the order.
//   phase 3  lane k assembles candidate k's objective; 32 results leave as coalesced stores.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include "cov_device.cuh"
#include "cov_kernel_common.cuh"
#include "cov_span_common.cuh"
#include "cov_kernels.cuh"
#include "../../include/coverage_cuda.h"

namespace cov {

constexpr int kSmallMaxN = 8;

__host__ __device__ inline int small_warp_bytes(const GridDesc &g, int N, int chunk)
{
    const int fb = round_up(g.ny * g.stride * 4, 16);
    const int dp = chunk * N * 32;
    const int pre = chunk * 16; // 8 x u16 item prefixes per candidate
    return fb + dp + pre;
}

template <bool MULTI, int CHUNK>
__global__ void __launch_bounds__(512, 1)
span_small_kernel(const __grid_constant__ GridDesc g, const __grid_constant__ ObjParams o,
                  const double *__restrict__ X, long long B, EvalOut out, unsigned long long *counter,
                  int force_exact)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const int N = o.N;
    const int planes_bytes = g.n_planes * g.plane_words * 4;
    const int warp_bytes = small_warp_bytes(g, N, CHUNK);
    const int fb_bytes = round_up(g.ny * g.stride * 4, 16);
    uint32_t *planes_s = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned char *wbase = smem_raw + planes_bytes + (size_t)warp * warp_bytes;
    uint32_t *fb = reinterpret_cast<uint32_t *>(wbase);
    SDisc *dp = reinterpret_cast<SDisc *>(wbase + fb_bytes);
    uint4 *prefix = reinterpret_cast<uint4 *>(wbase + fb_bytes + CHUNK * N * 32);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + planes_bytes + (size_t)warps * warp_bytes);

    stage_planes(g, planes_s, bar, planes_bytes); // TMA bulk copy of the fire planes, once per CTA
    for (int t = lane; t < fb_bytes / 16; t += 32) reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);
    __syncwarp();

    const long long n_chunks = (B + CHUNK - 1) / CHUNK;
    const int cstride = 3 * N;
    ItemCtx ictx;
    ictx.g = &g;
    ictx.N = N;
    ictx.force_exact = force_exact;

    for (;;) {
        unsigned long long chunk = 0;
        if (lane == 0) chunk = atomicAdd(counter, 1ull);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if ((long long)chunk >= n_chunks) break;
        const long long base = (long long)chunk * CHUNK;
        const int in_chunk = (int)min((long long)CHUNK, B - base);

        // ---------------- phase 1: lane k sets up candidate k ----------------
        double my_viol = 0.0, my_prog = 0.0;
        int my_feas = 1;
        if ((int)lane < in_chunk) {
            const double *xr = X + (base + lane) * cstride;
            double px[kSmallMaxN], py[kSmallMaxN], pr[kSmallMaxN]; // compile-time indices only: registers
#pragma unroll
            for (int c = 0; c < kSmallMaxN; ++c) {
                px[c] = py[c] = pr[c] = 0.0;
                if (c < N) {
                    px[c] = __ldg(xr + c);
                    py[c] = __ldg(xr + N + c);
                    pr[c] = __ldg(xr + 2 * N + c);
                }
            }
            // penalty: sequential FP64 sum in index order (src/TDM_STATIC_opt.jl:89-93)
#pragma unroll
            for (int i = 0; i < kSmallMaxN; ++i)
                if (i < N) {
                    const double diff = __dsub_rn(pr[i], o.r_max[i]);
                    my_viol = __dadd_rn(my_viol, fabs(diff));
                    if (out.progressive) my_prog = __dadd_rn(my_prog, julia_max0(diff));
                }
            bool bad = false;
            if (o.use_cons3) {
#pragma unroll
                for (int i = 0; i < kSmallMaxN; ++i)
                    if (i < N) {
                        const double ax = __dsub_rn(o.prev_x[i], px[i]);
                        const double ay = __dsub_rn(o.prev_y[i], py[i]);
                        const double az = __dsub_rn(o.prev_z[i], __ddiv_rn(pr[i], o.tan_half_fov));
                        const double s =
                            __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
                        bad |= (s >= o.cons3_G[i]);
                    }
            }
            if (o.use_cons7) {
#pragma unroll
                for (int i = 0; i < kSmallMaxN; ++i)
                    if (i < N) bad |= (py[i] < 200.0) && (pr[i] > o.cons7_R);
            }
            if (o.use_cons8) {
#pragma unroll
                for (int i = 0; i < kSmallMaxN; ++i)
#pragma unroll
                    for (int j2 = i + 1; j2 < kSmallMaxN; ++j2)
                        if (j2 < N) {
                            const double ax = __dsub_rn(px[i], px[j2]);
                            const double ay = __dsub_rn(py[i], py[j2]);
                            bad |= (__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)) < o.sep_T);
                        }
            }
            my_feas = !bad;
            // discs that may share cells with another disc of the candidate (bounding boxes, two cells
            // of margin; NaN compares false -> "may share")
            uint32_t shared_mask = 0;
#pragma unroll
            for (int a = 0; a < kSmallMaxN; ++a)
#pragma unroll
                for (int b2 = a + 1; b2 < kSmallMaxN; ++b2)
                    if (b2 < N) {
                        const double sr = pr[a] + pr[b2];
                        const bool apart = (fabs(px[a] - px[b2]) >= sr + 2.0 * g.dx) ||
                                           (fabs(py[a] - py[b2]) >= sr + 2.0 * g.dy);
                        if (!apart) shared_mask |= (1u << a) | (1u << b2);
                    }
            uint32_t run = 0;
            uint32_t pre[kSmallMaxN];
#pragma unroll
            for (int c = 0; c < kSmallMaxN; ++c) {
                pre[c] = 0xffffu;
                if (c < N) {
                    SDisc d;
                    const int rows = make_sdisc(g, px[c], py[c], pr[c], d);
                    d.flags |= ((shared_mask >> c) & 1u) << 1;
                    dp[lane * N + c] = d;
                    run += (uint32_t)rows;
                    pre[c] = run < 0xfffeu ? run : 0xfffeu; // inclusive prefix, saturated (see phase 2)
                }
            }
            prefix[lane] = make_uint4(pre[0] | (pre[1] << 16), pre[2] | (pre[3] << 16), pre[4] | (pre[5] << 16),
                                      pre[6] | (pre[7] << 16));
        }
        __syncwarp();

        // ---------------- phase 2: one candidate at a time, (disc, row) items over the lanes ----------------
        long long my_cnt = 0;
        long long my_cls[kMaxClasses];
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) my_cls[k] = 0;

        for (int kc = 0; kc < in_chunk; ++kc) {
            const uint4 pq = prefix[kc];
            uint32_t pre[kSmallMaxN] = {pq.x & 0xffffu, pq.x >> 16, pq.y & 0xffffu, pq.y >> 16,
                                        pq.z & 0xffffu, pq.z >> 16, pq.w & 0xffffu, pq.w >> 16};
            // total number of items = the last real prefix
            uint32_t total = 0;
#pragma unroll
            for (int c = 0; c < kSmallMaxN; ++c)
                if (c < N) total = pre[c];
            const SDisc *cdp = dp + kc * N;
            ictx.xrow = X + (base + kc) * cstride;
            uint32_t cnt[MULTI ? kMaxClasses : 1];
#pragma unroll
            for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) cnt[k] = 0;

            bool any_shared = false;
            if (total >= 0xfffeu) {
                // >= 65534 (disc, row) items cannot happen on a framebuffer that fits shared memory, but
                // stay safe: disc by disc, everything through the framebuffer
                for (int c = 0; c < N; ++c) {
                    const SDisc d = cdp[c];
                    const int r0 = d.rows & 0xffffu, r1 = d.rows >> 16;
                    for (int j = r0 + (int)lane; j <= r1; j += 32) {
                        int lo, hi;
                        int st = fast_span(g, d, j, force_exact, lo, hi);
                        if (st == kSlow) {
                            slow_item(g, ictx.xrow, N, c, j, (d.flags & 1u) || force_exact, lo, hi);
                            if (lo <= hi) st = kSpan;
                            else { st = kEmpty; lo = hi = 1; }
                        }
                        paint_span<MULTI>(g, fb, planes_s, j, lo, hi, st == kSpan, true, cnt);
                    }
                }
                __syncwarp();
                for (int t = lane; t < fb_bytes / 16; t += 32)
                    reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);
            } else {
                // two items per lane and iteration: independent instruction streams hide the FP32 latencies
#pragma unroll 1
                for (uint32_t t0 = lane; t0 < total; t0 += 64) {
                    const uint32_t t1 = t0 + 32;
                    const bool has1 = t1 < total;
                    // disc index = number of inclusive prefixes <= t (non-decreasing; unused slots 0xffff)
                    int c0 = 0, c1 = 0;
                    uint32_t before0 = 0u, before1 = 0u;
#pragma unroll
                    for (int q = 0; q < kSmallMaxN - 1; ++q) {
                        if (t0 >= pre[q]) {
                            c0 = q + 1;
                            before0 = pre[q];
                        }
                        if (t1 >= pre[q]) {
                            c1 = q + 1;
                            before1 = pre[q];
                        }
                    }
                    if (!has1) {
                        c1 = c0;
                        before1 = before0 + 32; // any valid row of the same disc: the result is discarded
                    }
                    const SDisc d0 = cdp[c0], d1 = cdp[c1];
                    const int j0 = (int)(d0.rows & 0xffffu) + (int)(t0 - before0);
                    const int j1 = (int)(d1.rows & 0xffffu) + (int)(t1 - before1);
                    int lo0, hi0, lo1, hi1;
                    int st0 = fast_span(g, d0, j0, force_exact, lo0, hi0);
                    int st1 = fast_span(g, d1, j1, force_exact, lo1, hi1);
                    if (!has1) st1 = kEmpty;
                    if (st0 == kSlow) {
                        slow_item(g, ictx.xrow, N, c0, j0, (d0.flags & 1u) || force_exact, lo0, hi0);
                        if (lo0 <= hi0) st0 = kSpan;
                        else { st0 = kEmpty; lo0 = hi0 = 1; }
                    }
                    if (st1 == kSlow) {
                        slow_item(g, ictx.xrow, N, c1, j1, (d1.flags & 1u) || force_exact, lo1, hi1);
                        if (lo1 <= hi1) st1 = kSpan;
                        else { st1 = kEmpty; lo1 = hi1 = 1; }
                    }
                    const bool sh0 = (d0.flags & 2u) != 0, sh1 = (d1.flags & 2u) != 0;
                    any_shared |= sh0 | (sh1 && has1);
                    paint_span<MULTI>(g, fb, planes_s, j0, lo0, hi0, st0 == kSpan, sh0, cnt);
                    paint_span<MULTI>(g, fb, planes_s, j1, lo1, hi1, st1 == kSpan, sh1, cnt);
                }
                // clear what the shared discs painted: every word of their bounding boxes
                if (__any_sync(0xffffffffu, any_shared)) {
                    __syncwarp();
#pragma unroll 1
                    for (int c = 0; c < N; ++c) {
                        const SDisc d = cdp[c];
                        if (!(d.flags & 2u)) continue;
                        const int r0 = d.rows & 0xffffu, r1 = d.rows >> 16;
                        int wa = 0, wb = g.wpr - 1;
                        if (!(d.flags & 1u)) {
                            // columns within sqrt(Tf + delta) of the centre, one cell of margin
                            const float hw = ceilf(sqrtf(d.Tf + d.delta) * g.inv_dxf) + 2.0f;
                            const float nxf = int_to_float_small(g.nx);
                            const int ca = (int)fminf(fmaxf(d.icf - hw, 1.0f), nxf);
                            const int cb = (int)fminf(fmaxf(d.icf + hw, 1.0f), nxf);
                            wa = (ca - 1) >> 5;
                            wb = (cb - 1) >> 5;
                        }
                        for (int j = r0 + (int)lane; j <= r1; j += 32) {
                            uint32_t *frow = fb + (j - 1) * g.stride;
                            for (int w = wa; w <= wb; ++w) frow[w] = 0u;
                        }
                    }
                }
            }
            __syncwarp();

            long long total_cnt = 0;
#pragma unroll
            for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) {
                const long long v = (long long)__reduce_add_sync(0xffffffffu, cnt[k]);
                total_cnt += v;
                if ((int)lane == kc) my_cls[k] = v;
            }
            if ((int)lane == kc) my_cnt = total_cnt;
        }

        // ---------------- phase 3: lane k finishes candidate k; coalesced stores ----------------
        if ((int)lane < in_chunk) {
            const long long bidx = base + lane;
            out.obj[bidx] = assemble_objective(g, o, my_cls, my_viol);
            if (out.count) out.count[bidx] = my_cnt;
            if (out.feasible) out.feasible[bidx] = (unsigned char)my_feas;
            if (out.progressive) out.progressive[bidx] = my_prog;
            if (out.class_count)
                for (int k = 0; k < g.n_classes; ++k) out.class_count[bidx * g.n_classes + k] = my_cls[k];
        }
        __syncwarp(); // dp / prefix are rewritten by the next chunk
    }
}

// ------------------------------------------------------------------------------------------
// launcher: returns false when this kernel does not apply (the general span kernel takes over)
// ------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem_small(K kernel, int bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

bool span_small_applies(const GridDesc &g, int N, const LaunchCfg &cfg, int *warps_out, int *chunk_out)
{
    if (N > kSmallMaxN || g.stride > 255 || g.ny > 65535) return false;
    const int planes_bytes = g.n_planes * g.plane_words * 4;
    int best_w = 0, best_chunk = 0;
    for (int chunk : {32, 16}) {
        const int per = small_warp_bytes(g, N, chunk);
        int w = (cfg.max_smem_optin - planes_bytes - 16) / per;
        w = std::min(w, 16); // __launch_bounds__(512): 16 warps, <= 128 registers per thread
        if (cfg.warps_per_cta > 0) w = std::min(w, cfg.warps_per_cta);
        // prefer the 32-candidate chunk unless the 16-candidate one buys >= 25 % more warps
        if (w >= 4 && (best_w == 0 || w * 4 >= best_w * 5)) {
            best_w = w;
            best_chunk = chunk;
        }
    }
    if (best_w < 4) return false;
    if (warps_out) *warps_out = best_w;
    if (chunk_out) *chunk_out = best_chunk;
    return true;
}

cudaError_t launch_span_small(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                              long long B, const EvalOut &out, unsigned long long *counter, cudaStream_t stream,
                              LaunchInfo *info)
{
    int warps = 0, chunk = 0;
    if (!span_small_applies(g, o.N, cfg, &warps, &chunk)) return cudaErrorInvalidConfiguration;
    const bool multi = !(g.n_planes == 1 && g.n_classes == 1 && g.plane_mult[0] == 1);
    const int smem = g.n_planes * g.plane_words * 4 + warps * small_warp_bytes(g, o.N, chunk) + 16;
    const long long chunks = (B + chunk - 1) / chunk;
    const int grid = (int)std::min<long long>((chunks + warps - 1) / warps, (long long)cfg.num_sms);
    if (info) {
        info->grid = grid;
        info->block = warps * 32;
        info->smem_bytes = smem;
        info->band_rows = g.ny;
        info->planes_in_smem = 1;
    }
    cudaError_t err;
#define COV_LAUNCH_SMALL(M, C)                                                                             \
    do {                                                                                                   \
        err = set_smem_small(span_small_kernel<M, C>, smem);                                               \
        if (err != cudaSuccess) return err;                                                                \
        span_small_kernel<M, C><<<grid, warps * 32, smem, stream>>>(g, o, dX, B, out, counter,             \
                                                                    cfg.force_exact);                      \
    } while (0)
    if (multi) {
        if (chunk == 32) COV_LAUNCH_SMALL(true, 32);
        else COV_LAUNCH_SMALL(true, 16);
    } else {
        if (chunk == 32) COV_LAUNCH_SMALL(false, 32);
        else COV_LAUNCH_SMALL(false, 16);
    }
#undef COV_LAUNCH_SMALL
    return cudaGetLastError();
}

} // namespace cov
