// cov_kernels.cuh -- launch interface between the C ABI (cov_api.cu) and the kernels
// (cov_kernels.cu, cov_grid_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "cov_types.h"

namespace cov {

struct LaunchCfg {
    int kernel;          // COV_KERNEL_*
    int warps_per_cta;   // 0 = auto
    int ctas_per_sm;     // 0 = auto
    int band_rows;       // 0 = auto
    int force_exact;
    int num_sms;
    int max_smem_optin;  // bytes
    int plane_mode;      // CTA kernel: -1 auto, 0 lazy global, 1 early global, 2 staged per band (TMA), 3 sweep
    int ordered;         // 1: the store's area depends on the order of the Float64 additions -> ordered kernel
};

struct LaunchInfo {
    int grid, block, smem_bytes, band_rows, planes_in_smem;
    int kernel; // cov_kernel of the kernel that ran (SPAN = small-swarm, SPAN_GENERAL = CTA per candidate)
    // template instantiation that ran (names the ncu profile the issue roofline is computed from)
    int multi;      // 1: several planes / classes / multiplicities
    int chunk;      // small-swarm kernel: candidates per unit (CHUNK); else 0
    int max_warps;  // small-swarm kernel: MAXW of the instantiation (20 or 24); else 0
    int plane_mode; // CTA kernel: 0 lazy, 1 early, 2 staged (PLANES); else -1
};

// Coverage objective over B candidates (device pointers). counter: one ZEROED unsigned long long
// per launch (work dispenser; the caller hands out fresh slots of a ring it zeroes in bulk, so no
// memset sits between consecutive launches). Returns cudaError_t.
cudaError_t launch_eval(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                        long long B, const EvalOut &out, unsigned long long *counter,
                        cudaStream_t stream, LaunchInfo *info);

// The small-swarm span kernel (cov_span_small.cu): N <= 8 and a framebuffer that fits shared memory.
bool span_small_applies(const GridDesc &g, int N, const LaunchCfg &cfg, long long B, int *warps_out, int *chunk_out);
cudaError_t launch_span_small(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                              long long B, const EvalOut &out, unsigned long long *counter, cudaStream_t stream,
                              LaunchInfo *info);

// The general span kernel (cov_span_cta.cu): one CTA per candidate, banded framebuffer.
cudaError_t launch_span_cta(const GridDesc &g, const ObjParams &o, const LaunchCfg &cfg, const double *dX,
                            long long B, const EvalOut &out, unsigned long long *counter, cudaStream_t stream,
                            LaunchInfo *info);

// ---- cell-store kernels (cov_grid_kernels.cu) ----
// stats[0] = entries, stats[1] = cells, stats[2 + k] = OR of the multiplicities of class k
cudaError_t launch_grid_stats(const unsigned char *mult, const unsigned char *cls, long long ncell,
                              unsigned long long *stats, cudaStream_t s);
cudaError_t launch_pack_planes(const unsigned char *mult, const unsigned char *cls, const GridDesc &g,
                               uint32_t *planes, cudaStream_t s);
// plane 0 in the pair-aligned, swizzled layout of the CTA kernel's sweep mode (GridDesc::planes_q)
cudaError_t launch_requad_plane(const GridDesc &g, uint32_t *planes_q, cudaStream_t s);
cudaError_t launch_fill_full(unsigned char *mult, unsigned char *cls, long long ncell, cudaStream_t s);
cudaError_t launch_bits_to_cells(const uint32_t *bits, int nx, int ny, unsigned char *mult,
                                 unsigned char *cls, cudaStream_t s);
// xyT: 3N doubles on the device: cx[N], cy[N], T[N]. removed: one unsigned long long.
cudaError_t launch_remove_covered(unsigned char *mult, unsigned char *cls, const GridDesc &g, const double *xyT, int N,
                                  unsigned long long *removed, cudaStream_t s);
cudaError_t launch_normalize_cls(const unsigned char *mult, unsigned char *cls, long long ncell, cudaStream_t s);
cudaError_t launch_covered_mask(unsigned char *mask, const GridDesc &g, const double *xyT, int N,
                                cudaStream_t s);
cudaError_t launch_thresholds(const double *xyR, int N, double *xyT, cudaStream_t s);
cudaError_t launch_add_points(unsigned char *mult, unsigned char *cls, const int *cell_idx,
                              const unsigned char *cell_cls, long long P, int *overflow, cudaStream_t s);
cudaError_t launch_argmin(const double *obj, const unsigned char *feasible, long long B, int barrier,
                          double *scratch_obj, long long *scratch_idx, int scratch_n, cudaStream_t s);
cudaError_t launch_union_area(const double *dX, long long B, int N, double *d_area, cudaStream_t s);
cudaError_t launch_fire_step(const unsigned char *cur, unsigned char *nxt, unsigned char *mult, unsigned char *cls,
                             int nx, int ny, unsigned long long seed, unsigned int step, const double *p_dir,
                             int append, unsigned long long *pushed, int *overflow, cudaStream_t s);
cudaError_t launch_fire_seed(const unsigned char *state, unsigned char *mult, unsigned char *cls, long long ncell,
                             int push_initial, cudaStream_t s);
cudaError_t launch_generate(double *dX, long long B, int N, unsigned long long seed, long long first,
                            double lx, double ly, double h_min, double h_max, double tan_half_fov,
                            cudaStream_t s);
// n packed candidate values (pack: COV_PACK_F32 / _I32 / _I16 of coverage_cuda.h) -> doubles: value * g (FP32: value)
cudaError_t launch_unpack(const void *raw, int pack, double g, double *out, long long n, cudaStream_t s);

} // namespace cov
