// cov_span_common.cuh -- the (disc, row) span machinery shared by the span kernels
// (cov_span_small.cu: small swarms, one warp per candidate chunk; cov_span_cta.cu: any swarm, one
// CTA per candidate).  See DESIGN.md section 2 for the exactness argument.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "cov_device.cuh"
#include "cov_kernel_common.cuh"

// -DCOV_DEBUG_BOUNDS builds a checking variant of the library (device asserts on every framebuffer / plane
// index): the in-house substitute for compute-sanitizer, which is closed on the GPU pool.
#ifdef COV_DEBUG_BOUNDS
#include <cassert>
#define COV_ASSERT(c) assert(c)
#else
#define COV_ASSERT(c) ((void)0)
#endif

namespace cov {

struct __align__(16) SDisc {
    float fx, fy;     // centre minus the centre of cell (ic, jc)
    float Tf, delta;  // s_f < Tf - delta: certainly inside; s_f > Tf + delta: certainly outside
    float icf, jcf;   // ic, jc (exact integers held as floats)
    uint32_t rows;    // r0 | r1 << 16, 1-based inclusive; r0 > r1: no rows
    uint32_t flags;   // bit 0: irregular, every row is decided in FP64; bit 1: may share cells with
                      // another disc of the candidate (its rows go through the framebuffer)
};
static_assert(sizeof(SDisc) == 32, "SDisc must be 32 bytes");

__device__ __forceinline__ double cell_centre_any(int i, double d, double half_d)
{
    return __dsub_rn(__dmul_rn((double)i, d), half_d);
}

// One disc of one candidate -> its shared-memory record. Returns the number of rows.
__device__ __forceinline__ int make_sdisc(const GridDesc &g, double cx, double cy, double R, SDisc &d)
{
    const double T = threshold(R);
    const double big = 1.7976931348623157e308;
    bool live = (T > 0.0) && (fabs(cx) <= big) && (fabs(cy) <= big);
    int r0 = 1, r1 = 0;
    if (live) {
        if (isinf(R)) {
            r1 = g.ny;
        } else {
            // rows j with |py_j - cy| < R, i.e. a < j < b with a = (cy-R)/dy + 1/2, b = (cy+R)/dy + 1/2:
            // j from floor(a) + 1 to ceil(b) - 1, with a and b pushed outwards by `extra`, a bound on the
            // FP64 rounding of this expression and of the reference's own fl(py - cy) for far-away
            // centres, so that a row left out is certainly outside
            const double extra = (fabs(cy) + R) * 8.8817841970012523e-16 * g.inv_dy; // 2^-50
            double lo = floor((cy - R) * g.inv_dy + 0.5 - extra) + 1.0;
            double hi = ceil((cy + R) * g.inv_dy + 0.5 + extra) - 1.0;
            if (!(lo <= (double)g.ny) || !(hi >= 1.0)) {
                live = false;
            } else {
                r0 = (int)fmax(lo, 1.0);
                r1 = (int)fmin(hi, (double)g.ny);
            }
        }
    }
    if (!live) {
        r0 = 1;
        r1 = 0;
    }
    d.rows = (uint32_t)r0 | ((uint32_t)r1 << 16);
    // relative frame: nearest cell (ic, jc); everything FP32 sees is a small offset from it
    const double gx = cx * g.inv_dx + 0.5, gy = cy * g.inv_dy + 0.5;
    bool regular = live && fabs(gx) < 4194304.0 && fabs(gy) < 4194304.0 && T < 1e30;
    d.flags = 1u;
    d.fx = d.fy = d.Tf = d.delta = d.icf = d.jcf = 0.0f;
    if (regular) {
        const int ic = __double2int_rn(gx), jc = __double2int_rn(gy);
        const float fx = (float)__dsub_rn(cx, cell_centre_any(ic, g.dx, g.hdx));
        const float fy = (float)__dsub_rn(cy, cell_centre_any(jc, g.dy, g.hdy));
        const float Tf = (float)T;
        const float rt = sqrtf(Tf);
        // FP32 error of one relative coordinate (DESIGN.md "error band"):
        //   2^-23 (|offset| + |f|)  [fma rounding, dxf, f]  +  2^-48 M  [the FP64 roundings of the
        //   reference's own cell centres and of f], M = max(|cx|, |cy|, extent)
        const float M = fmaxf(fmaxf(fabsf((float)cx), fabsf((float)cy)), g.extent) * 1.0000002f;
        const float E = 1.1920929e-07f * (rt + fmaxf(fabsf(fx), fabsf(fy))) + 3.5527137e-15f * M;
        const float delta = 2.0f * (4.0f * rt * E + 2.0f * E * E + Tf * 4.76837158203125e-07f);
        if (Tf > 64.0f * E * E && Tf > delta) {
            d.fx = fx;
            d.fy = fy;
            d.Tf = Tf;
            d.delta = delta;
            d.icf = (float)ic;
            d.jcf = (float)jc;
            d.flags = 0u;
        }
    }
    return r1 - r0 + 1 > 0 ? r1 - r0 + 1 : 0;
}

struct ItemCtx {
    const GridDesc *g;
    const double *xrow;   // the candidate's 3N doubles in global memory (slow path only)
    int N;
    int force_exact;
};

// Exact FP64 span of row j of disc c (the slow path).
static __device__ __noinline__ void slow_span(const GridDesc &g, const double *xrow, int N, int c, int j, int lo_e,
                                              int hi_e, int &lo, int &hi)
{
    RowExact r;
    r.cx = xrow[c];
    const double cy = xrow[N + c];
    r.T = threshold(xrow[2 * N + c]);
    r.dx = g.dx;
    r.hdx = g.hdx;
    r.nx = g.nx;
    const double ddy = __dsub_rn(cell_centre(j, g.dy, g.hdy), cy);
    r.dy2 = __dmul_rn(ddy, ddy);
    if (!(fabs(r.cx) <= 1.7976931348623157e308) || !(r.T > 0.0)) {
        lo = 1;
        hi = 0;
        return;
    }
    exact_span(r, lo_e, hi_e, lo, hi);
}

// Straight-line FP32 computation + certification of the covered columns of row j for disc d.
// Returns kEmpty (certainly no cell), kSpan (certainly exactly [lo, hi]) or kSlow (FP64 decides;
// lo, hi then hold in-grid guesses for the exact walk).  No branches: two items interleave.
//
// Certification without evaluating a single cell.  In the disc's relative frame the cell at column offset u
// (an integer) has the radicand M(u) = (u dx - fx)^2 + dy2 in real arithmetic on the FP32 inputs; the FP32
// evaluation s_f32(u) the brute-force kernel uses differs from M(u) by at most 1.5 * 2^-23 s, and s_f32 from the
// reference's Float64 radicand by at most delta / 2 (delta carries a 2x slack), so with
// delta2 = 1.25 delta:   M(u) < Tf - delta2  =>  inside,   M(u) > Tf + delta2  =>  outside.
// Hence cell u is certainly inside when |u dx - fx| < W_in = sqrt(Tf - dy2 - delta2) and certainly outside when
// |u dx - fx| > W_out = sqrt(Tf - dy2 + delta2).  With W = sqrt(Tf - dy2) and Tf - dy2 >= 8 delta2:
//     |W_in/out - W| <= 0.52 delta2 / W.
// The kernel computes w = w2 * rsqrt(w2) (relative error < 2^-20.5 including the rounding of w2) and the two
// boundary positions xl = (fx - w) / dx, xr = (fx + w) / dx in cell units (relative error < 1.1 * 2^-22 each).
// Every true boundary position (for W_in and for W_out) therefore lies within
//     e = 0.75 delta / dx * rsqrt(w2)  +  2^-19.5 * max(|xl|, |xr|)
// of the computed one (0.75 delta = 0.6 delta2 >= 0.52 delta2 * (1 + 2^-19); 2^-19.5 >= 2^-20.5 + 1.1 * 2^-22
// with room for the rounding of e itself).  If no integer lies within e of xl nor of xr, then
// lo = ceil(xl) is certainly inside and lo - 1 certainly outside, hi = floor(xr) likewise, and when
// ceil(xl) > floor(xr) every column is certainly outside (each integer is left of xl - e or right of xr + e).
// The differences ceil(xl) - xl and xr - floor(xr) are exact in FP32.  Rows whose chord is too short for the
// linearisation (Tf - dy2 <= 10 delta) go to the FP64 walk unless the row is certainly outside the disc
// (dy2 > Tf + 10 delta); so do NaN / Inf / huge / tiny discs (flag bit 0) and every row under force_exact.
enum { kEmpty = 0, kSpan = 1, kSlow = 2 };
__device__ __forceinline__ int fast_span(const GridDesc &g, const SDisc &d, int j, int force_exact, int &lo, int &hi)
{
    const float v = int_to_float_small(j) - d.jcf;
    const float y = fmaf(v, g.dyf, -d.fy);
    const float dy2 = y * y;
    const float w2 = d.Tf - dy2;
    // one MUFU, flush-to-zero; w2 <= 0 gives Inf / NaN, which `deep` below keeps from being believed
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(w2));
    const float w = w2 * rs;
    const float xl = (d.fx - w) * g.inv_dxf, xr = (d.fx + w) * g.inv_dxf;
    const float ulo = ceilf(xl), uhi = floorf(xr);
    const float e = fmaf(d.delta * g.k_ca, rs, fmaxf(fabsf(xl), fabsf(xr)) * 1.3487e-06f); // 2^-19.5, rounded up
    const float a = ulo - xl, b = xr - uhi; // distances to the next integer inside the span, in [0, 1)
    const float w2min = 10.0f * d.delta;
    const bool deep = w2 > w2min;
    const bool sure = deep && (fminf(a, b) > e) && (fmaxf(a, b) < 1.0f - e);
    const float nxf = int_to_float_small(g.nx);
    const float lof = fmaxf(d.icf + ulo, 1.0f), hif = fminf(d.icf + uhi, nxf); // NaN -> the grid edge
    lo = (int)fminf(lof, nxf);
    hi = (int)fmaxf(hif, 1.0f);
    int st = sure ? ((lof <= hif) ? kSpan : kEmpty) : kSlow;
    st = (w2 < -w2min) ? kEmpty : st; // the whole row is certainly outside the disc
    const bool irregular = (d.flags & 1u) || force_exact;
    return irregular ? kSlow : st;
}

// The slow path of an item: FP64 exact walk from the guesses.
static __device__ __noinline__ void slow_item(const GridDesc &g, const double *xrow, int N, int c, int j, bool irregular,
                                              int &lo, int &hi)
{
    int lo_g = lo, hi_g = hi;
    if (irregular) {
        const double gx = xrow[c] * g.inv_dx + 0.5;
        lo_g = hi_g = (gx >= 1.0) ? ((gx <= (double)g.nx) ? (int)gx : g.nx) : 1;
    }
    slow_span(g, xrow, N, c, j, lo_g, hi_g, lo, hi);
}

// The same for candidate number `unit_base + k` of the batch: the address of its doubles is formed HERE, inside the
// out-of-line slow path, so that the callers' hot loops carry no 64-bit address arithmetic for it.
static __device__ __noinline__ void slow_item_of(const GridDesc &g, const double *X, long long unit_base, int k, int N,
                                                 int c, int j, bool irregular, int &lo, int &hi)
{
    slow_item(g, X + (unit_base + k) * (3ll * N), N, c, j, irregular, lo, hi);
}

// Count the list entries on columns [lo, hi] of row j.  shared = false: the disc shares no cell with
// any other disc of the candidate, so its cells are counted directly.  shared = true: the interval
// is OR-ed into the warp's framebuffer and only the bits this lane was first to set are counted.
template <bool MULTI, bool PLANES_SMEM = true>
__device__ __forceinline__ void count_word(const GridDesc &g, const uint32_t *pw, uint32_t nw, uint32_t *cnt)
{
    if (!MULTI) {
        if (PLANES_SMEM || nw) cnt[0] += __popc(nw & ld_plane<PLANES_SMEM>(pw));
    } else {
        if (!PLANES_SMEM && !nw) return;
        for (int l = 0; l < g.n_planes; ++l) {
            const uint32_t v = __popc(nw & ld_plane<PLANES_SMEM>(pw + (size_t)l * g.plane_words)) * g.plane_mult[l];
            const int kcls = g.plane_class[l];
#pragma unroll
            for (int k = 0; k < kMaxClasses; ++k) cnt[k] += (k == kcls) ? v : 0u;
        }
    }
}

// Framebuffer layout: row-major with `stride` words per row.  When the row length is a power of two >= 8
// the rows are stored without padding and word w of row r sits at column w ^ ((r >> 2) & 7): 32 lanes on 32
// consecutive rows at the same column still hit 32 different banks, and the buffer is 1/9 smaller than with
// an odd stride (which buys the small kernel a 16th warp per SM at 256 columns).
struct FbLayout {
    int stride;   // words per framebuffer row
    int swz;      // 7 when swizzled, else 0
    __device__ __forceinline__ int at(int row, int w) const { return row * stride + (w ^ ((row >> 2) & swz)); }
};
__host__ __device__ inline bool fb_can_swizzle(const GridDesc &g) { return g.wpr >= 8 && (g.wpr & (g.wpr - 1)) == 0; }

// fb_row0: the 1-based grid row held by framebuffer row 0 (banded framebuffers).
// valid = false paints nothing (lets several items share one straight-line instruction stream).
// The first word, the last word and (WIDE) one word in between are handled without branches; any
// further whole words in a loop.  EARLY_PLANES (planes in global memory, single plane): the fire
// words are requested BEFORE the framebuffer atomics, so the L2 round trip overlaps them; without it
// a fire word is only read when the atomic left new bits, which is the better trade when the discs
// overlap so heavily that most words are already covered (200 UAVs on 500 m x 500 m).
template <bool MULTI, bool PLANES_SMEM = true, bool WIDE = false, bool EARLY_PLANES = false>
__device__ __forceinline__ void paint_span(const GridDesc &g, uint32_t *fb, const uint32_t *planes, int j, int lo,
                                           int hi, bool valid, bool shared, uint32_t *cnt, int fb_row0 = 1,
                                           FbLayout fl = FbLayout{0, 0})
{
    if (fl.stride == 0) fl.stride = g.stride;
    const int a = lo - 1, b = hi - 1;
    const int wa = a >> 5, wb = b >> 5;
    const int frow_i = j - fb_row0;
    const uint32_t *prow = planes + (size_t)(j - 1) * g.stride;
    COV_ASSERT(!valid || (lo >= 1 && lo <= hi && hi <= g.nx && j >= fb_row0 && j <= g.ny));
    COV_ASSERT(wa >= 0 && wa < g.wpr && wb >= 0 && wb < g.wpr);
    const uint32_t ma = 0xffffffffu << (a & 31), mb = 0xffffffffu >> (31 - (b & 31));
    uint32_t m0 = (wa == wb) ? (ma & mb) : ma;
    uint32_t m1 = (wa == wb) ? 0u : mb;
    const int wm = min(wa + 1, wb);
    uint32_t m2 = (WIDE && wb > wa + 1) ? 0xffffffffu : 0u;
    if (!valid) m0 = m1 = m2 = 0u;
    constexpr bool EARLY = EARLY_PLANES && !MULTI && !PLANES_SMEM;
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    if (EARLY) {
        if (m0) p0 = __ldg(prow + wa);
        if (m1) p1 = __ldg(prow + wb);
        if (WIDE && m2) p2 = __ldg(prow + wm);
    }
    if (shared) {
        if (m0) m0 &= ~atomicOr(fb + fl.at(frow_i, wa), m0);
        if (m1) m1 &= ~atomicOr(fb + fl.at(frow_i, wb), m1);
        if (WIDE && m2) m2 &= ~atomicOr(fb + fl.at(frow_i, wm), m2);
    }
    if (EARLY) {
        cnt[0] += __popc(m0 & p0) + __popc(m1 & p1) + (WIDE ? __popc(m2 & p2) : 0);
    } else {
        count_word<MULTI, PLANES_SMEM>(g, prow + wa, m0, cnt);
        count_word<MULTI, PLANES_SMEM>(g, prow + wb, m1, cnt);
        if (WIDE) count_word<MULTI, PLANES_SMEM>(g, prow + wm, m2, cnt);
    }
    if (valid) {
        int w = wa + (WIDE ? 2 : 1);
        if (EARLY) {
            uint32_t pn = (w < wb) ? __ldg(prow + w) : 0u; // one word ahead of the atomics
            for (; w < wb; ++w) {
                const uint32_t pw = pn;
                if (w + 1 < wb) pn = __ldg(prow + w + 1);
                uint32_t m = 0xffffffffu;
                if (shared) m &= ~atomicOr(fb + fl.at(frow_i, w), m);
                cnt[0] += __popc(m & pw);
            }
        } else if (WIDE) {
            for (; w < wb; ++w) { // whole words in between (wide spans: the compiler may unroll)
                uint32_t m = 0xffffffffu;
                if (shared) m &= ~atomicOr(fb + fl.at(frow_i, w), m);
                count_word<MULTI, PLANES_SMEM>(g, prow + w, m, cnt);
            }
        } else {
#pragma unroll 1
            for (; w < wb; ++w) { // rare in the small-swarm kernel (spans of more than two words): keep the code small
                uint32_t m = 0xffffffffu;
                if (shared) m &= ~atomicOr(fb + fl.at(frow_i, w), m);
                count_word<MULTI, PLANES_SMEM>(g, prow + w, m, cnt);
            }
        }
    }
}

} // namespace cov
