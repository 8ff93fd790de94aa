/* occ_b200.h -- C ABI of the B200 (sm_100a) TV-L1 optical flow solver WITH OCCLUSION DETECTION.
 *
 * Third solver of 12334zq/optical-flow-1 behind the same shared object:
 * src/tvl1occflow.cpp, src/tvl1occflow_solvers.cpp, src/tvl1occflow_tv_rof_box.cpp and the median
 * filter of src/utils.cpp.  Citations are relative to the reference tree (/root/reference).
 *
 *   reference interface                                              entry point here
 *   ---------------------------------------------------------------------------------------------------
 *   Dual_TVL1_optic_flow_multiscale (7 planes) src/tvl1occflow.h:111-129   occ_solve_f64, occ_solve_batch_*
 *   Dual_TVL1_optic_flow (7 planes)            src/tvl1occflow.h:63-79     occ_single_scale_f64
 *   Scalar_ROF_BoxCellCentered     src/tvl1occflow_tv_rof_box.cpp:25-645   occ_rof_box_f64   (test hook)
 *   me_median_filtering (window 3)             src/utils.cpp:151-213       occ_median3_f64   (test hook)
 *   nscales rule of the CLI               src/tvl1occflow_main.cpp:191-196  occ_clamp_nscales
 *
 * The C++ symbols with the reference's exact (mangled) signatures are exported by the same shared object
 * (csrc/tvl1flow_dropin.cpp) on top of this ABI, so the reference's unmodified tvl1occflow_main.cpp links
 * against it (cli/Makefile).
 *
 * ARITHMETIC.  Unlike the TV-L1 path of tvl1_b200.h (fp32 by north_star's own statement), this solver
 * computes in IEEE fp64 -- the reference's shipped ofpix_t (src/of.h:4-10) -- with the reference's
 * association order and without fused multiply-adds (the translation unit is built with -fmad=false):
 * its box relaxation is a lexicographic Gauss-Seidel sweep and its occlusion map is cut at hard
 * thresholds, so "the same result" has no useful tolerance short of the same bits.  Flow, occlusion map
 * and iteration counts are BIT-IDENTICAL to the reference (tests/test_occ_gpu.py).  The sweep keeps
 * the reference's data dependences on a wavefront schedule (cell (i, j) at step 2i + j), it does not
 * re-order them.
 *
 * DEFINED BEHAVIOUR.  The reference keeps its dual variables in function-level statics and reads
 * eta1 / eta2 uninitialised (src/tvl1occflow_solvers.cpp:163-186, :241-264, the file's own #warning).
 * This library implements what a fresh process computes: p and eta start from zero at every pyramid
 * level (oracle/tvl1_oracle.c section (e), oracle/occ_ref_shim.cpp pin the compiled reference to the
 * same reading).  Unlike the reference the entry points here are reentrant.
 */
#ifndef OCC_B200_H
#define OCC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OCC_OK 0
#define OCC_ERR_CUDA 1          /* a CUDA runtime call failed */
#define OCC_ERR_SIGMA 2         /* "GaussianSmooth: sigma too large" (src/operators.cpp:520-522) */
#define OCC_ERR_ARG 3           /* bad argument */
#define OCC_ERR_NODEVICE 4      /* no usable CUDA device: there is no CPU fallback */

#define OCC_EXT_MAX_ITERATIONS 20    /* src/tvl1occflow_constants.h:25 */
#define OCC_OMEGA 1.25               /* :26 */
#define OCC_IS_ZERO 1E-10            /* :28 */
#define OCC_THR_CHI 0.75             /* :29 */
#define OCC_MAX_ITERATIONS_CHI 100   /* :30 */
#define OCC_PRESMOOTHING_SIGMA 0.8   /* :31 */
#define OCC_G_FACTOR 0.05            /* :35, with G_CHOICE 2 (:34) */
#define OCC_MAX_ITERATIONS_U 10      /* :38 */
#define OCC_TAU_ETA 0.15             /* :39 */
#define OCC_TAU_CHI 0.15             /* :40 */
#define OCC_MAX_LEVELS 16

typedef struct occ_ctx occ_ctx;

/* The reference's own parameter set (src/tvl1occflow.h:111-129), same meaning and order. */
typedef struct occ_params {
    double lambda;   /* weight of the data term */
    double alpha;    /* weight of chi |v|^2 */
    double beta;     /* weight of chi div(u) */
    double theta;    /* weight of (u - v)^2 */
    int nscales;     /* pyramid levels (not clamped here; the CLI clamps, see occ_clamp_nscales) */
    double zfactor;  /* pyramid down-sampling factor */
    int warps;       /* warps per level */
    double epsilon;  /* a warp step stops after the first outer iteration whose mean squared flow update is
                        <= epsilon (NOT epsilon^2: src/tvl1occflow.cpp:277), or after OCC_EXT_MAX_ITERATIONS */
} occ_params;

typedef struct occ_stats {
    unsigned long long kernel_launches;
    unsigned long long outer_iterations;      /* summed over triples, levels and warps */
    unsigned long long box_sweeps;            /* Gauss-Seidel sweeps launched (each serves every active plane) */
    unsigned long long box_cell_updates;      /* cells relaxed, all planes */
    unsigned long long chi_pixel_iterations;  /* pixels x primal-dual iterations of the occlusion map */
    unsigned long long host_syncs;
    double ms_total, ms_pyramid, ms_warp, ms_box, ms_chi, ms_other;   /* CUDA-event times when profiling is on */
} occ_stats;

void occ_default_params(occ_params *p);   /* src/tvl1occflow_constants.h:14-23; nscales = 100 like the CLI */
/* N = floor(log(min(nx, ny) / 16) / log(1 / zfactor)) + 1, nscales = min(nscales, N)  (main.cpp:191-196) */
int occ_clamp_nscales(int nx, int ny, int nscales, double zfactor);

int occ_create(int device, occ_ctx **out);
void occ_destroy(occ_ctx *ctx);
const char *occ_last_error(const occ_ctx *ctx);     /* ctx may be NULL: error of the last failed occ_create */
int occ_set_profiling(occ_ctx *ctx, int on);        /* CUDA events around the kernel groups (adds syncs) */
int occ_set_max_batch(occ_ctx *ctx, int triples);   /* triples advanced in lock-step (default 64; OCC_MAX_BATCH) */
int occ_get_stats(const occ_ctx *ctx, occ_stats *out);   /* of the last solve */
void *occ_get_stream(const occ_ctx *ctx);

/* Dual_TVL1_optic_flow_multiscale of src/tvl1occflow.cpp:335-482, HOST buffers, dense row-major [ny][nx].
 * I_1 = frame before I0, I1 = frame after, filtI0 = the image the weight g is taken from (NULL -> I0, as
 * the CLI does without a fourth image).  u1, u2, chi are outputs (their contents on entry are ignored,
 * :361-366); chi is 0 / 1 (thresholded at OCC_THR_CHI, :459).  iters_out / errs_out may be NULL; otherwise
 * [nscales*warps], coarsest level first: the numbers of the reference's "Warping: %d, Iterations: %d,
 * Error: %e" lines (:305-309). */
int occ_solve_f64(occ_ctx *ctx, const double *I_1, const double *I0, const double *I1, const double *filtI0,
                  double *u1, double *u2, double *chi, int nx, int ny, const occ_params *prm, int *iters_out,
                  double *errs_out);
/* ntriples independent frame triples of one shape, HOST buffers [ntriples][ny][nx]; iters_out / errs_out
 * [ntriples][nscales*warps].  Triples advance in lock-step; each stops on its own criterion. */
int occ_solve_batch_f64(occ_ctx *ctx, int ntriples, const double *I_1, const double *I0, const double *I1,
                        const double *filtI0, double *u1, double *u2, double *chi, int nx, int ny,
                        const occ_params *prm, int *iters_out, double *errs_out);
/* Same, DEVICE buffers; iters_out / errs_out are HOST pointers. */
int occ_solve_batch_dev_f64(occ_ctx *ctx, int ntriples, const double *dI_1, const double *dI0, const double *dI1,
                            const double *dfiltI0, double *du1, double *du2, double *dchi, int nx, int ny,
                            const occ_params *prm, int *iters_out, double *errs_out);

/* Dual_TVL1_optic_flow of src/tvl1occflow.cpp:144-330: one level, no smoothing; u1, u2, chi are IN/OUT (the
 * initial values are used, :211-224) and chi is NOT thresholded.  nscales / zfactor of prm ignored;
 * iters_out / errs_out [warps]. */
int occ_single_scale_f64(occ_ctx *ctx, const double *I_1, const double *I0, const double *I1,
                         const double *filtI0, double *u1, double *u2, double *chi, int nx, int ny,
                         const occ_params *prm, int *iters_out, double *errs_out);

/* Test hooks, HOST buffers.  occ_rof_box_f64: niter sweeps of Scalar_ROF_BoxCellCentered on (u, f, g) with
 * the dual values on the south (p1) and east (p2) side of every cell; u, p1, p2 in/out. */
int occ_rof_box_f64(occ_ctx *ctx, double *u, const double *f, double *p1, double *p2, const double *g,
                    double lambda, double omega, int nx, int ny, int niter);
int occ_median3_f64(occ_ctx *ctx, double *a, int nx, int ny);

#ifdef __cplusplus
}
#endif
#endif /* OCC_B200_H */
