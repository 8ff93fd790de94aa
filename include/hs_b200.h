/* hs_b200.h -- C ABI of the B200 (sm_100a) pyramidal Horn-Schunck solver.
 *
 * Second solver of 12334zq/optical-flow-1 behind the same library: it reuses
 * the TV-L1 path's pyramid, warp and up-sampling kernels and adds the SOR sweep.  Same conventions
 * as tvl1_b200.h (dense row-major planes, 0 / TVL1_ERR_* return codes, one tvl1_ctx per host thread
 * and GPU, no CPU fallback).  Citations are relative to the reference tree (/root/reference).
 *
 *   reference interface                                      entry point here
 *   ------------------------------------------------------------------------------------------
 *   horn_schunck_pyramidal      src/horn_schunck.h:35-48     hs_solve_f32 / _f64, hs_solve_batch_*
 *   horn_schunck_optical_flow   src/horn_schunck.h:15-26     hs_single_scale_f32 / _f64
 *   SOR loop of one warp step   src/horn_schunck_pyramidal.cpp:139-231   hs_sor_f32    (test hook)
 *   nscales rule of the CLI     src/horn_schunck_pyramidal_main.cpp:136-143   hs_clamp_nscales
 *
 * The C++ symbols with the reference's exact (mangled) signatures are exported by the same shared
 * object (csrc/tvl1flow_dropin.cpp) on top of this ABI.
 *
 * Semantics.  The reference sweeps the image in place (Gauss-Seidel / SOR, w = 1.9) in lexicographic
 * order, borders after the interior; it runs that sweep under an OpenMP parallel-for, so its own
 * result is only defined with one thread.  This implementation reproduces the ONE-thread order
 * exactly (every pixel reads the same mix of new and old neighbour values), in fp32.
 *
 * Limits of this build: every pyramid level needs nx >= 3 and ny >= 3.  Up to about HS_MAX_ROWS rows the
 * sweep of a frame pair keeps its ring of wave columns in the shared memory of one SM; taller levels keep
 * it in global memory (same schedule, same result, slower) and then need nx >= 24.
 */
#ifndef HS_B200_H
#define HS_B200_H

#include "tvl1_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HS_SOR_EXTRAPOLATION_PARAMETER 1.9   /* src/horn_schunck_pyramidal.cpp:21 */
#define HS_INPUT_PRESMOOTHING_SIGMA 0.8      /* src/horn_schunck_pyramidal.cpp:22 */
#define HS_MAX_ROWS 2560                     /* rows up to which the rings of k_hs_sor live in shared memory */

/* Solver parameters: the reference's own set (src/horn_schunck.h:35-48), same meaning and order. */
typedef struct hs_params {
    double alpha;    /* smoothing weight */
    int nscales;     /* number of pyramid levels (not clamped here; the CLI clamps, see hs_clamp_nscales) */
    double zfactor;  /* pyramid down-sampling factor, 0 < zfactor < 1 */
    int warps;       /* warps per level */
    double tol;      /* TOL: a warp step stops after the first sweep with sqrt(mean squared update) <= TOL */
    int maxiter;     /* ... or after maxiter sweeps */
} hs_params;

void hs_default_params(hs_params *p);   /* src/horn_schunck_pyramidal_main.cpp:25-30 */
/* "the smaller images of the pyramid don't have a size smaller than 16x16":
 * N = 1 + log(hypot(nx, ny) / 16) / log(1 / zfactor); nscales = min(nscales, (int) N)   (main.cpp:136-143) */
int hs_clamp_nscales(int nx, int ny, int nscales, double zfactor);

/* horn_schunck_pyramidal, HOST buffers.  iters_out / errs_out may be NULL; otherwise [nscales*warps]
 * per pair, coarsest level first: the numbers the reference prints as "Iterations %d (%g)"
 * (src/horn_schunck_pyramidal.cpp:233-235). */
int hs_solve_f32(tvl1_ctx *ctx, const float *I1, const float *I2, float *u, float *v, int nx, int ny,
                 const hs_params *prm, int *iters_out, double *errs_out);
int hs_solve_f64(tvl1_ctx *ctx, const double *I1, const double *I2, double *u, double *v, int nx, int ny,
                 const hs_params *prm, int *iters_out, double *errs_out);
/* npairs independent frame pairs of one shape, HOST buffers [npairs][ny][nx]. */
int hs_solve_batch_f32(tvl1_ctx *ctx, int npairs, const float *I1, const float *I2, float *u, float *v,
                       int nx, int ny, const hs_params *prm, int *iters_out, double *errs_out);
/* Same, DEVICE buffers (dense, 16-byte aligned); iters_out / errs_out are HOST pointers. */
int hs_solve_batch_dev_f32(tvl1_ctx *ctx, int npairs, const float *dI1, const float *dI2, float *du,
                           float *dv, int nx, int ny, const hs_params *prm, int *iters_out,
                           double *errs_out);

/* horn_schunck_optical_flow (one level, no normalisation / blur): u, v are in/out (the initial flow is
 * used, src/horn_schunck_pyramidal.cpp:123); nscales / zfactor of prm ignored; iters_out/errs_out [warps]. */
int hs_single_scale_f32(tvl1_ctx *ctx, const float *I1, const float *I2, float *u, float *v, int nx, int ny,
                        const hs_params *prm, int *iters_out, double *errs_out);
int hs_single_scale_f64(tvl1_ctx *ctx, const double *I1, const double *I2, double *u, double *v, int nx,
                        int ny, const hs_params *prm, int *iters_out, double *errs_out);

/* Test hook: the SOR loop of one warp step on a given system (HOST buffers).  I2wx, I2wy are the
 * warped gradients and rho_c = -(I1 - I2w + I2wx*u + I2wy*v) what the warp kernel stores
 * (src/horn_schunck_pyramidal.cpp:127-137 in terms of these: Au = -rho_c*I2wx, Du = I2wx^2 + alpha^2,
 * D = I2wx*I2wy).  u, v in/out.  prefetch = -1 picks the prefetch distance automatically, 0..3 forces
 * it (one-sweep kernel), -2 forces the rings into global memory (the path of levels with more than
 * ~HS_MAX_ROWS rows), -3 forces the pipelined kernel (the default wherever a level fits it), -4 the
 * experimental two-columns-per-step kernel (not yet validated on a GPU; never selected by default).
 * Outputs: sweeps done and the last sqrt(mean squared update). */
int hs_sor_f32(tvl1_ctx *ctx, const float *I2wx, const float *I2wy, const float *rho_c, float *u, float *v,
               int nx, int ny, double alpha, double tol, int maxiter, int prefetch, int *niter_out,
               double *err_out);

#ifdef __cplusplus
}
#endif
#endif /* HS_B200_H */
