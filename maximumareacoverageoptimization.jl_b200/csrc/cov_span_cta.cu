// cov_span_cta.cu -- the general span kernel of libcoverage_cuda (sm_100a): any swarm size
// (N <= 1024) on any grid (<= 32768^2), one CTA per candidate.
//
// Same arithmetic as the small-swarm kernel (cov_span_common.cuh; reference
// src/AreaCoverageCalculation.jl:63-110, src/TDM_STATIC_opt.jl:82-100, src/TDM_Constraints.jl:54-195
// of /root/reference), organised for candidates that are a lot of work each (50 UAVs on 1024^2:
// ~4 400 (disc, row) spans; 200 UAVs on 4096^2: ~70 000 spans of ~10 words):
//   * the CTA's 8 warps share ONE candidate: its 3N doubles are staged in shared memory, the disc
//     records are built by all threads, the O(N^2) separation test is spread over the warps while one
//     lane forms the order-dependent penalty sum;
//   * the union framebuffer lives in shared memory as a BAND of grid rows (268-row bands at 1024^2 and
//     96-row bands at 4096^2 when the band's fire words are staged beside it);
//   * per band the work is cut into units of 32 rows of one disc (a unit -> disc table; every thread claims
//     its discs' slots with one atomicAdd); warps take units in pairs from a shared-memory dispenser and every lane handles
//     one row of each unit: two independent spans per lane, atomicOr into the band, popcount of the
//     newly set bits against the fire plane, whose rows for the band are staged in shared memory by ONE
//     TMA bulk copy per band (cp.async.bulk + mbarrier, issued while the CTA builds the band's unit
//     table); small swarms and multi-plane stores read the fire words through L1/L2 instead.
// Persistent grid, up to three 256-thread CTAs per SM so that one candidate's serial phases
// (stage, setup, barriers) overlap another's span work; a CTA's first candidate is its block index,
// further ones come from an atomic counter (a poll set, one candidate per CTA, never touches it).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include "cov_device.cuh"
#include "cov_kernel_common.cuh"
#include "cov_span_common.cuh"
#include "cov_kernels.cuh"
#include "../../include/coverage_cuda.h"

namespace cov {

#ifndef COV_CTA_THREADS
#define COV_CTA_THREADS 256 // (overridable for occupancy experiments: make EXTRA="-DCOV_CTA_THREADS=384 -DCOV_CTAS_PER_SM=2")
#endif
#ifndef COV_CTAS_PER_SM
#define COV_CTAS_PER_SM 3
#endif
#ifndef COV_UNIT_ROWS
#define COV_UNIT_ROWS 16
#endif
// Rows of one disc that form a work unit: 16 -- each half of a warp works on its own unit (its own disc), so that a
// disc's rows inside a band are rounded up to 16 lanes instead of 32 (32: the whole warp on one unit, round 1's)
constexpr int kUnitRows = COV_UNIT_ROWS;
constexpr int kUnitShift = kUnitRows == 8 ? 3 : (kUnitRows == 16 ? 4 : 5);
static_assert(kUnitRows == 8 || kUnitRows == 16 || kUnitRows == 32, "work units are 8, 16 or 32 rows");
constexpr int kCtaThreads = COV_CTA_THREADS;
constexpr int kCtasPerSm = COV_CTAS_PER_SM;

struct CtaPlan {
    int fixed_bytes, tab_bytes, fb_bytes, plane_bytes, band_rows, total_bytes, ctas_per_sm;
};
// how the kernel reads the fire words
// kPlanesSweep: PAINT THEN SWEEP.  The spans are only painted -- the two edge words of a span with result-less
// atomicOr, the whole words in between with plain (paired, 64-bit) stores of all-ones -- and nothing is counted
// while painting; once the band is painted, all threads sweep it linearly with 128-bit loads:
// count += popc(framebuffer & fire plane), framebuffer = 0.  A covered cell is counted once however many discs
// cover it, the paint loop has no dependent atomic -> popc chain, and a word painted by k discs costs k cheap
// stores plus one sweep visit instead of k atomics with return values.  Framebuffer band and staged plane band
// share the layout of GridDesc::planes_q, so the sweep needs no row / column arithmetic at all.
// kPlanesSweepL2: the same with the plane read through L2 in the sweep instead of a staged copy -- half the
// shared memory per CTA, hence more co-resident CTAs or taller bands (an experiment, COV_OPT_PLANE_MODE 4).
enum { kPlanesLazy = 0, kPlanesEarly = 1, kPlanesStaged = 2, kPlanesSweep = 3, kPlanesSweepL2 = 4 };
// unit table: one 32-bit entry (disc << 16 | unit within the disc) per 32-row work unit of a band
__host__ __device__ inline int cta_tab_bytes(int N, int band_rows)
{
    return round_up(N * ((band_rows + kUnitRows - 1) / kUnitRows) * 4, 16);
}
__host__ __device__ inline int cta_fixed_bytes(int N)
{
    return round_up(3 * N * 8, 16) + N * 32 + round_up((N + 1) * 4, 16) + 1024;
}
// A band must hold at least this many rows before another co-resident CTA is worth its shared memory.  Measured on
// B200 (tools/band_threshold_exp.py; ms at 80 / 56 / 48 / 40): 5 UAVs on 4096^2 6.13 / 5.72 / 5.52 / 5.52, 20 on
// 4096^2 7.78 / 7.54 / 7.22 / 7.22, 1000 on 4096^2 4.97 / 3.62 / 3.62 / 3.62 (two CTAs instead of one), 50 on 8192^2
// 12.20 / 11.53 / 10.62 / 10.63, 200 on 8192^2 7.76 / 7.10 / 7.11 / 7.10; 200 on 4096^2 (C4) does not care whether it
// runs three CTAs with 108-row bands or four with 76 (5.86 / 5.87 / 5.86 / 5.85), nor 400 on 4096^2 (3 x 80 rows 3.91,
// 4 x 48 rows 3.87-3.90) -- but four CTAs execute 8 % more instructions for it (more bands: more tables, barriers and
// sweep set-up), so for swarms of 128 discs and more a FOURTH 256-thread CTA still has to leave 80-row bands (the
// per-band tables grow with the swarm; 50 UAVs on 8192^2 do gain from the fourth CTA: 3 x 68 rows 11.53, 4 x 48 10.62).
#ifndef COV_MIN_BAND_ROWS
#define COV_MIN_BAND_ROWS 48
#endif
#ifndef COV_MIN_BAND_ROWS_4TH
#define COV_MIN_BAND_ROWS_4TH 80
#endif
static CtaPlan cta_plan(const GridDesc &g, int N, int smem_per_sm, int band_rows_opt, int ctas_opt, bool staged,
                        bool sweep = false, int threads = kCtaThreads)
{
    CtaPlan p{};
    p.fixed_bytes = cta_fixed_bytes(N) + (staged ? 32 : 0); // + the band's plane rows and an mbarrier when staged
    const int fstride = sweep ? g.qstride : g.stride;
    const int row_bytes = fstride * 4 * (staged ? 2 : 1);
    // as many co-resident CTAs as possible (they overlap each other's serial phases), as long as a
    // band still holds a useful number of rows
    int best = 0;
    for (int ctas = (ctas_opt > 0 ? ctas_opt : kCtasPerSm); ctas >= 1; --ctas) {
        const int budget = smem_per_sm / ctas - 1024; // 1 KB per CTA is reserved by the system
        int rows = (int)((budget - p.fixed_bytes - 4 * N - 16) / (row_bytes + N / 8.0));
        if (rows > g.ny) rows = g.ny;
        const bool aligned = staged || sweep; // bands start on 16-byte boundaries of the plane
        if (aligned && rows < g.ny) rows &= ~3;
        while (rows > 1 && p.fixed_bytes + cta_tab_bytes(N, rows) + round_up(rows * row_bytes, 16) + 32 > budget) rows -= aligned ? 4 : 1;
        const int need = (threads >= 256 && ctas >= 4 && N >= 128) ? COV_MIN_BAND_ROWS_4TH : COV_MIN_BAND_ROWS;
        if (rows >= std::min(g.ny, need) || ctas == 1) {
            best = ctas;
            p.band_rows = rows;
            break;
        }
    }
    p.ctas_per_sm = best;
    if (band_rows_opt > 0 && band_rows_opt < p.band_rows) p.band_rows = (staged || sweep) ? std::max(4, band_rows_opt & ~3) : band_rows_opt;
    p.fb_bytes = p.band_rows > 0 ? round_up(p.band_rows * fstride * 4, 16) : 0;
    p.plane_bytes = staged && p.band_rows > 0 ? round_up(p.band_rows * fstride * 4, 16) + 16 : 0;
    p.tab_bytes = p.band_rows > 0 ? cta_tab_bytes(N, p.band_rows) : 0;
    p.total_bytes = p.fixed_bytes + p.tab_bytes + p.fb_bytes + p.plane_bytes;
    return p;
}

// Sweep mode: paint columns [lo, hi] of grid row j (1-based) into the band framebuffer (row 0 = grid row jb0),
// nothing else.  Edge words by atomicOr without a return value; whole words in between by plain stores of
// all-ones, in aligned pairs where possible (every writer of such a word writes the same value, and an atomicOr
// that meets the store on the same word leaves all-ones whichever comes first).
__device__ __forceinline__ void paint_only(uint32_t *fb, int qstride, int jb0, int j, int lo, int hi, bool valid)
{
    const int a = lo - 1, b = hi - 1;
    const int wa = a >> 5, wb = b >> 5;
    uint32_t *row = fb + (j - jb0) * qstride;
    const int sw = ((j - 1) >> 4) & 1;
    const uint32_t ma = 0xffffffffu << (a & 31), mb = 0xffffffffu >> (31 - (b & 31));
    COV_ASSERT(!valid || (lo >= 1 && lo <= hi && j >= jb0));
    if (!valid) return;
    if (wa == wb) {
        atomicOr(row + (wa ^ sw), ma & mb);
        return;
    }
    atomicOr(row + (wa ^ sw), ma);
    atomicOr(row + (wb ^ sw), mb);
    int s = wa + 1, e = wb; // whole words [s, e)
#ifdef COV_STRICT_ATOMICS
    // checking variant (make strict-atomics): whole words by atomicOr as well, so that no plain store ever meets an
    // atomic on the same word; the GPU suite must give identical results with both builds
    for (; s < e; ++s) atomicOr(row + (s ^ sw), 0xffffffffu);
    return;
#endif
    if (s < e && (s & 1)) {
        row[s ^ sw] = 0xffffffffu;
        ++s;
    }
    if (s < e && (e & 1)) {
        --e;
        row[e ^ sw] = 0xffffffffu;
    }
    // the first eight pairs straight-line and predicated (lanes of a warp have different word counts: a loop would
    // run for the longest and pay its branches for every lane); whatever is left, in a loop
    const uint2 ones = make_uint2(0xffffffffu, 0xffffffffu);
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (s + 2 * k < e) *reinterpret_cast<uint2 *>(row + s + 2 * k) = ones;
    if (__any_sync(__activemask(), s + 4 < e)) { // wide spans only (warp-wide vote: narrow ones skip the block)
#pragma unroll
        for (int k = 2; k < 8; ++k)
            if (s + 2 * k < e) *reinterpret_cast<uint2 *>(row + s + 2 * k) = ones;
#pragma unroll 1
        for (s += 16; s < e; s += 2) *reinterpret_cast<uint2 *>(row + s) = ones;
    }
}

// CTAS: co-resident CTAs per SM the instantiation is compiled for (3: up to 85 registers per thread; 4: 64)
// THREADS: 256, or 128 for swarms of up to 64 discs -- twice as many, smaller CTAs per SM: a candidate of a few
// thousand spans keeps four warps busy as well as eight, and every CTA-wide barrier waits for half as many warps
template <bool MULTI, int PLANES, int THREADS = kCtaThreads, int CTAS = kCtasPerSm>
__global__ void __launch_bounds__(THREADS, CTAS)
span_cta_kernel(const __grid_constant__ GridDesc g, const __grid_constant__ ObjParams o,
                const double *__restrict__ X, long long B, EvalOut out, unsigned long long *counter,
                int force_exact, int band_rows, int fb_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = THREADS / 32;
    const int N = o.N;
    const int cstride = 3 * N;
    double *stage = reinterpret_cast<double *>(smem_raw);
    SDisc *dp = reinterpret_cast<SDisc *>(smem_raw + round_up(3 * N * 8, 16));
    unsigned char *scratch = smem_raw + round_up(3 * N * 8, 16) + N * 32 + round_up((N + 1) * 4, 16);
    uint32_t *unit_tab = reinterpret_cast<uint32_t *>(scratch + 1024);
    uint32_t *fb = reinterpret_cast<uint32_t *>(scratch + 1024 + cta_tab_bytes(N, band_rows));
    // scratch: [0,8) next candidate; [8,16) violation; [16,24) progressive; [28,32) unit dispenser;
    //          [32,40) unit counters of even / odd bands; [512, 512 + nwarps*4*8) per-warp counts
    unsigned long long *s_next = reinterpret_cast<unsigned long long *>(scratch);
    double *s_viol = reinterpret_cast<double *>(scratch + 8);
    double *s_prog = reinterpret_cast<double *>(scratch + 16);
    uint32_t *s_disp = reinterpret_cast<uint32_t *>(scratch + 28);
    uint32_t *s_units = reinterpret_cast<uint32_t *>(scratch + 32); // unit counters of the even / odd bands
    unsigned long long *s_cnt = reinterpret_cast<unsigned long long *>(scratch + 512);

    // staged planes: the band's rows of the fire plane, brought in by one TMA bulk copy per band
    uint32_t *plane_s = fb + fb_bytes / 4;
    uint64_t *bar = reinterpret_cast<uint64_t *>(plane_s + fb_bytes / 4);
    uint32_t bar_phase = 0;
    constexpr bool kTma = PLANES == kPlanesStaged || PLANES == kPlanesSweep; // the band's plane rows come by TMA
    constexpr bool kSweep = PLANES == kPlanesSweep || PLANES == kPlanesSweepL2;
    const int fstride = kSweep ? g.qstride : g.stride;                        // words per framebuffer row
    if (kTma && tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = tid; t < fb_bytes / 16; t += THREADS) reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);

    // candidates: the first one is the CTA's own index, further ones come from the global dispenser (a poll
    // set, one candidate per CTA, never touches it: two global round trips less on a 10 us kernel)
    long long cand = blockIdx.x;
    for (bool first = true;; first = false) {
        __syncthreads(); // previous candidate fully retired (stage/dp/scratch reusable, fb clean)
        if (!first) {
            if ((long long)gridDim.x >= B) break;
            if (tid == 0) *s_next = (unsigned long long)gridDim.x + atomicAdd(counter, 1ull);
            __syncthreads();
            cand = (long long)*s_next;
        }
        if (cand >= B) break;
        const double *xr = X + cand * cstride;

        // ---- A. stage the candidate ----
        for (int t = tid; t < cstride; t += THREADS) stage[t] = __ldg(xr + t);
        __syncthreads();

        // ---- B. disc records (threads from the front) and, at the same time, the order-dependent
        //         penalty sums on the last thread ----
        for (int c = tid; c < N; c += THREADS) {
            SDisc d;
            make_sdisc(g, stage[c], stage[N + c], stage[2 * N + c], d);
            d.flags |= 2u; // large swarms overlap as a rule: every disc goes through the framebuffer
            dp[c] = d;
        }
        if (tid == 0) s_units[0] = s_units[1] = 0;
        if (tid == THREADS - 1) {
            double viol = 0.0, prog = 0.0;
            for (int i = 0; i < N; ++i) {
                const double diff = __dsub_rn(stage[2 * N + i], o.r_max[i]);
                viol = __dadd_rn(viol, fabs(diff));
                if (out.progressive && prog_takes(o, i)) prog = __dadd_rn(prog, julia_max0(diff));
            }
            *s_viol = viol;
            *s_prog = prog;
        }
        // ---- C. constraints (everyone; a conjunction, so order-free) ----
        bool bad = false;
        if (o.use_cons3) {
            for (int i = tid; i < N; i += THREADS) {
                const double ax = __dsub_rn(o.prev_x[i], stage[i]);
                const double ay = __dsub_rn(o.prev_y[i], stage[N + i]);
                const double az = __dsub_rn(o.prev_z[i], __ddiv_rn(stage[2 * N + i], o.tan_half_fov));
                const double s = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
                bad |= (s >= o.cons3_G[i]);
            }
        }
        if (o.use_cons7) {
            for (int i = tid; i < N; i += THREADS) bad |= (stage[N + i] < 200.0) && (stage[2 * N + i] > o.cons7_R);
        }
        if (o.use_cons8) {
            // unordered pairs decide the reference's ordered-pair loop ((xi-xj)^2 is symmetric)
            for (int i = nwarps - 1 - warp; i < N - 1; i += nwarps) { // back warps first: the front ones build discs
                const double xi = stage[i], yi = stage[N + i];
                for (int j2 = i + 1 + lane; j2 < N; j2 += 32) {
                    const double ax = __dsub_rn(xi, stage[j2]);
                    const double ay = __dsub_rn(yi, stage[N + j2]);
                    bad |= (__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)) < o.sep_T);
                }
            }
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0); // also publishes dp[] and the penalty sums

        uint32_t cnt[MULTI ? kMaxClasses : 1];
#pragma unroll
        for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) cnt[k] = 0;
        unsigned long long cls_total[MULTI ? kMaxClasses : 1];
#pragma unroll
        for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) cls_total[k] = 0;

        // ---- D. bands of framebuffer rows: two CTA barriers per band ----
        int band = 0;
        for (int jb0 = 1; jb0 <= g.ny; jb0 += band_rows, ++band) {
            const int jb1 = min(g.ny, jb0 + band_rows - 1);
            auto issue_plane_copy = [&]() {
                // rows jb0..jb1 of the plane are contiguous; (jb0 - 1) * stride is a multiple of 4 words
                const uint32_t words = (uint32_t)round_up((jb1 - jb0 + 1) * fstride, 4);
                const uint32_t *src = PLANES == kPlanesSweep ? g.planes_q : g.planes;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic reads of plane_s
                mbar_expect_tx(bar, words * 4);
                bulk_g2s(plane_s, src + (size_t)(jb0 - 1) * fstride, words * 4, bar);
            };
            // staged mode reads the plane while painting, and the previous band's painting is over (barrier B)
            if (PLANES == kPlanesStaged && tid == 32) issue_plane_copy();
            // work units = 32-row blocks of a disc's rows inside the band; unit_tab[u] = (disc, unit within the
            // disc).  Every thread claims the table slots of its discs with one shared-memory atomicAdd (the
            // order of the units is immaterial); the counters alternate between bands so that the next band's
            // counter can be zeroed while this band's is still being read.
            if (tid == 0) *s_disp = 0; // the dispenser of this band (its last user finished before barrier B)
            uint32_t *s_u = s_units + (band & 1);
            for (int cb = 0; cb < N; cb += THREADS) {
                const int c = cb + tid;
                if (c < N) {
                    const uint32_t rows = dp[c].rows;
                    const int r0 = max((int)(rows & 0xffffu), jb0), r1 = min((int)(rows >> 16), jb1);
                    const uint32_t n = r1 >= r0 ? (uint32_t)((r1 - r0 + kUnitRows) >> kUnitShift) : 0u;
                    if (n) {
                        const uint32_t first = atomicAdd(s_u, n);
                        for (uint32_t k = 0; k < n; ++k) unit_tab[first + k] = ((uint32_t)c << 16) | k;
                    }
                }
            }
            __syncthreads(); // (A) table complete; everybody is done with the previous band (sweep / clear included)
            const uint32_t units = *s_u;
            if (tid == 0) s_units[(band + 1) & 1] = 0;
            // sweep mode needs the plane only after the painting: the copy has the whole paint phase to land, and
            // issuing it after (A) saves the barrier that would protect the previous band's sweep reads
            if (PLANES == kPlanesSweep && tid == 32) issue_plane_copy();
            if (PLANES == kPlanesStaged) {
                mbar_wait(bar, bar_phase);
                bar_phase ^= 1u;
            }
            // paint_span indexes planes by grid row: shift the band buffer so that row jb0 lands on its row 0
            const uint32_t *planes_eff = PLANES == kPlanesStaged ? plane_s - (size_t)(jb0 - 1) * g.stride : g.planes;
            if (units != 0) {
                // warps take units in pairs from the CTA's dispenser: two independent spans per lane
                for (;;) {
                    // (dispensing single units towards the end of a band, to shorten the wait at the barrier, was
                    // measured and lost: a band has only 30-50 units, and two independent spans per lane are worth
                    // more than a finer tail -- C3 3.22 -> 3.90 ms)
                    // a warp takes 64 lane-slots of units per iteration: 2 units of 32 rows, or 4 of 16 (lanes 0-15
                    // and 16-31 then work on different units, hence possibly different discs)
                    constexpr uint32_t kPerSlot = 32 / kUnitRows; // units side by side in one 32-lane slot
                    uint32_t u0 = 0;
                    if (lane == 0) u0 = atomicAdd(s_disp, 2u * kPerSlot);
                    u0 = __shfl_sync(0xffffffffu, u0, 0);
                    if (u0 >= units) break;
                    const uint32_t half = (uint32_t)lane >> kUnitShift; // which unit of the slot this lane works on
                    const uint32_t ua = u0 + half, ub = u0 + kPerSlot + half;
                    const bool has0 = ua < units, has1 = ub < units;
                    // disc and position of a unit: one table entry each (one or two addresses per warp: broadcast)
                    const uint32_t e0 = unit_tab[min(ua, units - 1)], e1 = unit_tab[min(ub, units - 1)];
                    const int c0 = (int)(e0 >> 16), c1 = (int)(e1 >> 16);
                    COV_ASSERT(c0 >= 0 && c0 < N && c1 >= 0 && c1 < N);
                    const SDisc d0 = dp[c0], d1 = dp[c1];
                    const int ul = lane & (kUnitRows - 1);
                    const int j0 = max((int)(d0.rows & 0xffffu), jb0) + (int)((e0 & 0xffffu) << kUnitShift) + ul;
                    const int j1 = max((int)(d1.rows & 0xffffu), jb0) + (int)((e1 & 0xffffu) << kUnitShift) + ul;
                    const bool in0 = has0 && j0 <= min((int)(d0.rows >> 16), jb1);
                    const bool in1 = has1 && j1 <= min((int)(d1.rows >> 16), jb1);
                    const int jj0 = in0 ? j0 : jb0, jj1 = in1 ? j1 : jb0; // any row of the band: result discarded
                    COV_ASSERT(jj0 >= jb0 && jj0 <= jb1 && jj1 >= jb0 && jj1 <= jb1);
                    int lo0, hi0, lo1, hi1;
                    int st0 = fast_span(g, d0, jj0, force_exact, lo0, hi0);
                    int st1 = fast_span(g, d1, jj1, force_exact, lo1, hi1);
                    if (!in0) st0 = kEmpty;
                    if (!in1) st1 = kEmpty;
                    if (st0 == kSlow) {
                        int l2 = lo0, h2 = hi0; // temporaries: lo/hi stay in registers
                        slow_item(g, xr, N, c0, jj0, (d0.flags & 1u) || force_exact, l2, h2);
                        if (l2 <= h2) st0 = kSpan;
                        else { st0 = kEmpty; l2 = h2 = 1; }
                        lo0 = l2;
                        hi0 = h2;
                    }
                    if (st1 == kSlow) {
                        int l2 = lo1, h2 = hi1; // temporaries: lo/hi stay in registers
                        slow_item(g, xr, N, c1, jj1, (d1.flags & 1u) || force_exact, l2, h2);
                        if (l2 <= h2) st1 = kSpan;
                        else { st1 = kEmpty; l2 = h2 = 1; }
                        lo1 = l2;
                        hi1 = h2;
                    }
                    if (kSweep) {
                        paint_only(fb, fstride, jb0, jj0, lo0, hi0, st0 == kSpan);
                        paint_only(fb, fstride, jb0, jj1, lo1, hi1, st1 == kSpan);
                    } else {
                        paint_span<MULTI, PLANES == kPlanesStaged, true, PLANES == kPlanesEarly>(g, fb, planes_eff, jj0, lo0, hi0, st0 == kSpan, true, cnt, jb0);
                        paint_span<MULTI, PLANES == kPlanesStaged, true, PLANES == kPlanesEarly>(g, fb, planes_eff, jj1, lo1, hi1, st1 == kSpan, true, cnt, jb0);
                    }
                }
            }
            __syncthreads(); // (B) the band is painted
            if (kSweep) {
                if (PLANES == kPlanesSweep) {
                    mbar_wait(bar, bar_phase); // the band's plane rows (in flight since barrier A)
                    bar_phase ^= 1u;
                }
                if (units != 0) {
                    // count and clear in one linear pass: framebuffer and plane band have the same layout
                    const int used = (jb1 - jb0 + 1) * fstride;
                    uint4 *f4 = reinterpret_cast<uint4 *>(fb);
                    const uint4 *p4 = PLANES == kPlanesSweep
                                          ? reinterpret_cast<const uint4 *>(plane_s)
                                          : reinterpret_cast<const uint4 *>(g.planes_q + (size_t)(jb0 - 1) * fstride);
                    uint32_t c = 0;
                    for (int t = tid; t < (used + 3) / 4; t += THREADS) {
                        // (skipping unpainted quads -- no plane read, no clear -- was measured: -2 % for 13 UAVs on
                        // 1024^2, +3 % / +5 % on the C3 / C4 shapes: the branch costs more than it saves)
                        const uint4 f = f4[t], pl = PLANES == kPlanesSweep ? p4[t] : __ldg(p4 + t);
                        c += __popc(f.x & pl.x) + __popc(f.y & pl.y) + __popc(f.z & pl.z) + __popc(f.w & pl.w);
                        f4[t] = make_uint4(0, 0, 0, 0);
                    }
                    cnt[0] += c;
                }
                // (no barrier: the next band's table barrier (A) orders this sweep before the next copy and paint)
            } else if (units != 0) {
                // clear the band for the next band / candidate
                const int used = (jb1 - jb0 + 1) * g.stride;
                for (int t = tid; t < (used + 3) / 4; t += THREADS)
                    reinterpret_cast<uint4 *>(fb)[t] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) { // keep 32-bit partials far from overflow
                cls_total[k] += cnt[k];
                cnt[k] = 0;
            }
            // the next band's table writes wait for everyone at barrier (B); the clear is ordered against the next
            // band's painting by its barrier (A)
        }

        // ---- E. reduce the counts, assemble, write ----
#pragma unroll
        for (int k = 0; k < (MULTI ? kMaxClasses : 1); ++k) {
            unsigned long long v = cls_total[k];
            for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) s_cnt[warp * kMaxClasses + k] = v;
        }
        __syncthreads();
        if (tid == 0) {
            long long tot[kMaxClasses];
            long long total_cnt = 0;
            for (int k = 0; k < kMaxClasses; ++k) {
                tot[k] = 0;
                if (k < (MULTI ? kMaxClasses : 1))
                    for (int w = 0; w < nwarps; ++w) tot[k] += (long long)s_cnt[w * kMaxClasses + k];
                total_cnt += tot[k];
            }
            const double my_obj = assemble_objective(g, o, tot, *s_viol);
            out.obj[cand] = my_obj;
            store_mirrors(out, cand, my_obj, any_bad ? 0 : 1);
            if (out.count) out.count[cand] = total_cnt;
            if (out.feasible) out.feasible[cand] = (unsigned char)(any_bad ? 0 : 1);
            if (out.progressive) out.progressive[cand] = *s_prog;
            if (out.class_count)
                for (int k = 0; k < g.n_classes; ++k) out.class_count[cand * g.n_classes + k] = tot[k];
        }
    }
}

cudaError_t launch_span_cta(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                            long long B, const EvalOut &out, unsigned long long *counter, cudaStream_t stream,
                            LaunchInfo *info)
{
    const bool multi = !(g.n_planes == 1 && g.n_classes == 1 && g.plane_mult[0] == 1);
    // how the fire words are read: staged per band by TMA (single plane), else through L1/L2 -- requested
    // ahead of the framebuffer atomics for moderate swarms, only when the atomic left new bits for dense ones
    // (measured, B200: staging wins from a few dozen discs per candidate on -- C3 +6 %, C4 +27 % -- and loses
    // for a handful of discs on a big grid, where most of a staged band is never looked at)
    // Which way to count.  Paint-then-sweep with the plane through L2 (mode 4) is the default: no counting atomics,
    // no plane band in shared memory, hence a fourth co-resident CTA or taller bands.  Measured on B200 (ms; mode 4
    // vs the better of early / staged): 50 UAVs on 1024^2 2.95 / 3.45, 200 on 4096^2 5.99 / 8.99, 1000 on 4096^2
    // 4.93 / 8.48, 100 on 2048^2 2.77 / 3.44, 33 on 512^2 0.84 / 0.94, 20 on 256^2 0.90 / 1.10, 16 on 2048^2
    // 3.92 / 4.27, 5 on 4096^2 6.56 / 7.87 (many bands: occupancy decides), 5 on 100^2 (1000 candidates) 14 / 18 us.
    // It loses only where a small swarm leaves most of a grid untouched AND the grid fits two or three bands, so
    // that the sweep is the larger part of the work: 9 on 1024^2 4.64 / 3.90, 13 on 1024^2 5.00 / 4.68.  The radii
    // are not known at launch time, so swarm size and grid words decide.
    const long long grid_words = (long long)g.ny * g.qstride;
    const bool small_swarm_mid_grid = o.N < 16 && grid_words > 2000ll * o.N && grid_words <= 36000;
    int mode = multi ? kPlanesLazy
                     : (cfg.plane_mode >= 0 ? cfg.plane_mode
                                            : ((g.planes_q && !small_swarm_mid_grid) ? kPlanesSweepL2 : kPlanesEarly));
    if (multi || ((mode == kPlanesSweep || mode == kPlanesSweepL2) && !g.planes_q)) mode = kPlanesLazy;
    // the L2 sweep needs no plane band in shared memory: room for a fourth co-resident CTA (its own instantiation,
    // capped at 64 registers) when the bands stay tall enough
    // CTA size.  Measured on B200 (ms, 128 threads x 8 CTAs per SM vs 256 x 3 or 4): 9 / 13 UAVs on 1024^2 3.04 / 3.75
    // vs 3.91 / 4.69, 16 on 2048^2 3.30 vs 3.92, 20 on 256^2 0.65 vs 0.90, 33 on 512^2 0.78 vs 0.84, 50 on 1024^2 2.86
    // vs 2.95; 100 on 2048^2 2.82 vs 2.77, 200 on 4096^2 7.87 vs 5.98, 1000 on 4096^2 8.71 vs 4.93.
    bool small_cta = kCtaThreads == 256 && mode != kPlanesStaged && mode != kPlanesSweep &&
                     !(mode == kPlanesLazy && !multi) && (cfg.warps_per_cta > 0 ? cfg.warps_per_cta <= 4 : o.N <= 64);
    const bool staged = mode == kPlanesStaged || mode == kPlanesSweep;
    const bool sweep = mode == kPlanesSweep || mode == kPlanesSweepL2;
    CtaPlan p{};
    const int big_ctas_try = cfg.ctas_per_sm > 0 ? cfg.ctas_per_sm : (mode == kPlanesSweepL2 ? std::max(4, kCtasPerSm) : kCtasPerSm);
    if (small_cta) {
        p = cta_plan(g, o.N, cfg.max_smem_optin + 1024, cfg.band_rows, cfg.ctas_per_sm > 0 ? cfg.ctas_per_sm : 8, staged, sweep, 128);
        // on grids so wide that only a few bands' worth of CTAs fit, 128-thread CTAs would leave the SM short of warps:
        // they must bring at least as many warps per SM as the 256-thread plan (50 UAVs on 4096^2: 7 x 128 threads
        // 7.17 ms, 4 x 256 7.04 ms)
        if (cfg.warps_per_cta == 0) {
            const CtaPlan big = cta_plan(g, o.N, cfg.max_smem_optin + 1024, cfg.band_rows, big_ctas_try, staged, sweep);
            if (p.ctas_per_sm < 5 || p.ctas_per_sm * 4 < big.ctas_per_sm * 8) small_cta = false;
        }
    }
    const int threads = small_cta ? 128 : kCtaThreads;
    if (!small_cta) p = cta_plan(g, o.N, cfg.max_smem_optin + 1024, cfg.band_rows, big_ctas_try, staged, sweep);
    if ((mode == kPlanesStaged || mode == kPlanesSweep || mode == kPlanesSweepL2) && p.band_rows < 4) {
        mode = o.N <= 96 ? kPlanesEarly : kPlanesLazy;
        p = cta_plan(g, o.N, cfg.max_smem_optin + 1024, cfg.band_rows, cfg.ctas_per_sm, false);
    }
    if (p.band_rows < 1) return cudaErrorInvalidConfiguration;
    const int grid = (int)std::min<long long>(B, (long long)cfg.num_sms * p.ctas_per_sm);
    if (info) {
        info->grid = grid;
        info->block = threads;
        info->smem_bytes = p.total_bytes;
        info->band_rows = p.band_rows;
        info->planes_in_smem = mode == kPlanesStaged || mode == kPlanesSweep;
        info->kernel = COV_KERNEL_SPAN_GENERAL;
        info->multi = multi;
        info->chunk = small_cta ? 8 : ((mode == kPlanesSweepL2 && p.ctas_per_sm >= 4 && kCtasPerSm < 4) ? 4 : kCtasPerSm); // CTAs per SM compiled for
        info->max_warps = 0;
        info->plane_mode = mode;
    }
    cudaError_t err;
#define COV_LAUNCH_CTA_AS(M, E, T, C)                                                                             \
    do {                                                                                                          \
        err = cudaFuncSetAttribute(span_cta_kernel<M, E, T, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                   p.total_bytes);                                                                \
        if (err != cudaSuccess) return err;                                                                       \
        span_cta_kernel<M, E, T, C><<<grid, T, p.total_bytes, stream>>>(g, o, dX, B, out, counter,                \
                                                                        cfg.force_exact, p.band_rows, p.fb_bytes); \
    } while (0)
#define COV_LAUNCH_CTA(M, E) COV_LAUNCH_CTA_AS(M, E, kCtaThreads, kCtasPerSm)
    if (small_cta) {
        if (multi) COV_LAUNCH_CTA_AS(true, kPlanesLazy, 128, 8);
        else if (mode == kPlanesSweepL2) COV_LAUNCH_CTA_AS(false, kPlanesSweepL2, 128, 8);
        else COV_LAUNCH_CTA_AS(false, kPlanesEarly, 128, 8);
    } else if (multi) COV_LAUNCH_CTA(true, kPlanesLazy);
    else if (mode == kPlanesSweep) COV_LAUNCH_CTA(false, kPlanesSweep);
    else if (mode == kPlanesSweepL2 && p.ctas_per_sm >= 4 && kCtasPerSm < 4) COV_LAUNCH_CTA_AS(false, kPlanesSweepL2, kCtaThreads, 4);
    else if (mode == kPlanesSweepL2) COV_LAUNCH_CTA(false, kPlanesSweepL2);
    else if (mode == kPlanesStaged) COV_LAUNCH_CTA(false, kPlanesStaged);
    else if (mode == kPlanesEarly) COV_LAUNCH_CTA(false, kPlanesEarly);
    else COV_LAUNCH_CTA(false, kPlanesLazy);
#undef COV_LAUNCH_CTA
#undef COV_LAUNCH_CTA_AS
    return cudaGetLastError();
}

} // namespace cov
