/* tvl1_b200.h -- C ABI of the B200 (sm_100a) TV-L1 optical-flow solver.
 *
 * This is the drop-in boundary for the TV-L1 path of 12334zq/optical-flow-1.  Plain C, plain
 * pointers and sizes.  Citations are relative to the reference tree (/root/reference).
 *
 *   reference interface                                    entry point here
 *   ------------------------------------------------------------------------------------------
 *   Dual_TVL1_optic_flow_multiscale  src/tvl1flow.h:56-70   tvl1_solve_f32 / _f64, tvl1_solve_batch_*
 *   Dual_TVL1_optic_flow             src/tvl1flow.h:36-48   tvl1_single_scale_f32 / _f64
 *   (C99 float variant of both)      3rdparty/tvl1flow_3/tvl1flow_lib.c:47-60, :299-314
 *   image_normalization_2            src/utils.h:27         tvl1_normalize_f32            (test hook)
 *   gaussian                         src/operators.h:128    tvl1_gaussian_f32             (test hook)
 *   zoom_size / zoom_out / zoom_in   src/zoom.h:20,32,57    tvl1_zoom_size, tvl1_zoom_out_f32, tvl1_zoom_in_f32
 *   centered_gradient + 3x bicubic_interpolation_warp + rho_c/grad loop
 *                                    src/tvl1flow.cpp:84,94-109   tvl1_warp_f32           (test hook)
 *   loop body of the while           src/tvl1flow.cpp:114-181     tvl1_iterate_f32        (test hook)
 *
 * The C++ symbols with the reference's exact (mangled) signatures are exported by the same
 * shared object (csrc/tvl1flow_dropin.cpp) and are implemented on top of this ABI only.
 *
 * Conventions
 *  - images and flow fields are dense row-major planes, index p = i*nx + j, no padding
 *    (src/tvl1flow.cpp:61, SURVEY 8b "Data layout").  Batches are npairs such planes back to back.
 *  - every function returns 0 on success or a TVL1_ERR_* code; tvl1_last_error() gives text.
 *  - there is no CPU fallback: without a CUDA device tvl1_create() fails.
 *  - a tvl1_ctx is bound to one device and one host thread at a time; use one context per
 *    thread / per GPU for concurrent calls.
 */
#ifndef TVL1_B200_H
#define TVL1_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVL1_OK 0
#define TVL1_ERR_CUDA 1          /* a CUDA runtime call failed */
#define TVL1_ERR_SIGMA 2         /* "GaussianSmooth: sigma too large" (src/operators.cpp:520-522) */
#define TVL1_ERR_ARG 3           /* bad argument (null pointer, non-positive size, ...) */
#define TVL1_ERR_NODEVICE 4      /* no usable CUDA device */

#define TVL1_MAX_ITERATIONS 300      /* src/tvl1flow.cpp:22 */
#define TVL1_PRESMOOTHING_SIGMA 0.8  /* src/tvl1flow.cpp:23 */
#define TVL1_GRAD_IS_ZERO 1E-10      /* src/tvl1flow.cpp:24 */
#define TVL1_ZOOM_SIGMA_ZERO 0.6     /* src/zoom.cpp:15 */
#define TVL1_MAX_LEVELS 16           /* pyramid levels this build keeps statistics for */

typedef struct tvl1_ctx tvl1_ctx;

/* Solver parameters: the reference's own set (src/tvl1flow.h:56-70), same meaning and order. */
typedef struct tvl1_params {
    double tau;      /* time step */
    double lambda;   /* weight of the data term */
    double theta;    /* weight of (u - v)^2 */
    int nscales;     /* number of pyramid levels (not clamped here; the CLI clamps, tvl1flow_main.cpp:185-188) */
    double zfactor;  /* pyramid down-sampling factor, 0 < zfactor < 1 */
    int warps;       /* warps per level */
    double epsilon;  /* stopping threshold; the loop stops after the first iteration whose mean
                        squared update is <= epsilon^2, or after TVL1_MAX_ITERATIONS */
} tvl1_params;

/* Counters of the most recent solve on a context (all device work is stream-ordered). */
typedef struct tvl1_stats {
    unsigned long long kernel_launches;   /* kernels of this library launched */
    unsigned long long iterate_launches;  /* of which: fused iteration kernel */
    unsigned long long pixel_iterations;  /* sum over iteration launches of (pixels of every pair still iterating) */
    unsigned long long pixel_warps;       /* pixels processed by the warp+precompute kernel */
    double iterate_ms;                    /* device time inside iteration launches (CUDA events; 0 unless profiling on) */
    double warp_ms;                       /* device time inside warp launches (same) */
    double total_ms;                      /* device time of the whole solve (same) */
    double pyramid_ms;                    /* min/max + normalise + blur + zoom_out (same) */
    double zoom_in_ms;                    /* flow up-sampling between levels (same) */
    double export_ms;                     /* flow export to the caller's dense layout (same) */
    unsigned long long host_syncs;        /* stream synchronisations issued for loop control */
    /* per pyramid level (0 = finest), iteration kernel only */
    unsigned long long level_pixel_iterations[TVL1_MAX_LEVELS];
    unsigned long long level_iterate_launches[TVL1_MAX_LEVELS];
    double level_iterate_ms[TVL1_MAX_LEVELS];
    /* of level_iterate_ms: the level's first launch where it runs the first TWO iterations of every pair from zero
     * duals (k_iterate_t2); 0 where the first launch is one iteration */
    double level_first_block_ms[TVL1_MAX_LEVELS];
} tvl1_stats;

/* -- life cycle --------------------------------------------------------------------------- */
int tvl1_device_count(void);
int tvl1_create(int device, tvl1_ctx **out);
void tvl1_destroy(tvl1_ctx *ctx);
const char *tvl1_last_error(const tvl1_ctx *ctx);  /* ctx may be NULL: error of the last failed tvl1_create */
int tvl1_set_profiling(tvl1_ctx *ctx, int on);     /* bracket kernels with CUDA events (tvl1_stats *_ms) */
int tvl1_set_max_batch(tvl1_ctx *ctx, int pairs);  /* pairs advanced in lock-step per workspace (default 32) */
/* Batches larger than max_batch are cut into chunks that up to 8 lanes (sibling contexts on the same
 * GPU, one host thread each) process concurrently: copies of one chunk overlap kernels of another (experimental,
 * TVL1_HOST_PIPE=1 and pinned host buffers: one upload -> solve -> download pipeline per call, tvl1_plan_chunks).
 * host_lanes: host-buffer entry points (default 4); dev_lanes: device-buffer entry point (default 2). */
int tvl1_set_lanes(tvl1_ctx *ctx, int host_lanes, int dev_lanes);
/* How a host-buffer batch of `npairs` pairs in PINNED memory is cut into lock-step chunks by the call-wide
 * upload -> solve -> download pipeline (opt-in, TVL1_HOST_PIPE=1; no GPU needed for this query): ramped
 * sizes -- max_batch/8, /4, /2 (none below 8 pairs), full chunks, /2, /4, /8 -- so that the first kernels start and the last download
 * ends one short chunk away from the ends of the call (csrc/tvl1_solver.cu: ramp_schedule, solve_host_pipelined).
 * Writes up to `cap` chunk sizes to `sizes` and returns the number of chunks. */
int tvl1_plan_chunks(int npairs, int max_batch, int *sizes, int cap);
int tvl1_get_stats(const tvl1_ctx *ctx, tvl1_stats *out);
/* Bit s set: pyramid level s of the next solve may run two iterations per launch (k_iterate_t2) where the launch is
 * large enough to saturate HBM.  Follows the loop lengths of the context's previous solve (diagnostics; the flow does
 * not depend on it). */
unsigned int tvl1_get_blocked_levels(const tvl1_ctx *ctx);
void *tvl1_get_stream(const tvl1_ctx *ctx);        /* the cudaStream_t all work of this context is issued on */
void tvl1_default_params(tvl1_params *p);          /* tvl1flow_main.cpp:24-33 with nscales = 5 */

/* The C++ drop-in symbols (Dual_TVL1_optic_flow_multiscale / Dual_TVL1_optic_flow) keep one context
 * per calling host thread.  This selects the GPU for the calling thread (default: environment
 * variable TVL1_DEVICE, else 0): one thread per GPU shards the frame pairs of a video. */
void tvl1_dropin_set_device(int device);

/* -- the solver: Dual_TVL1_optic_flow_multiscale (src/tvl1flow.cpp:219-328) ---------------- */
/* HOST buffers.  iters_out / errs_out may be NULL; otherwise [nscales*warps] per pair, coarsest
 * level first, exactly the numbers the reference prints in verbose mode (tvl1flow.cpp:184-188). */
int tvl1_solve_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2,
                   int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
int tvl1_solve_f64(tvl1_ctx *ctx, const double *I0, const double *I1, double *u1, double *u2,
                   int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
/* npairs independent frame pairs of one shape, HOST buffers [npairs][ny][nx]. */
int tvl1_solve_batch_f32(tvl1_ctx *ctx, int npairs, const float *I0, const float *I1, float *u1,
                         float *u2, int nx, int ny, const tvl1_params *prm, int *iters_out,
                         double *errs_out);
int tvl1_solve_batch_f64(tvl1_ctx *ctx, int npairs, const double *I0, const double *I1, double *u1,
                         double *u2, int nx, int ny, const tvl1_params *prm, int *iters_out,
                         double *errs_out);
/* A frame sequence (video): nframes frames [nframes][ny][nx] in HOST memory give nframes-1 flows,
 * pair b = (frame b, frame b+1), u1/u2 [nframes-1][ny][nx].  Same results as tvl1_solve_batch_* on
 * the expanded pairs (each pair keeps its own joint normalisation, src/tvl1flow.cpp:255), but every
 * frame crosses PCIe once instead of twice.  This is the batch/video mode of the CLI shell
 * (src/tvl1flow_main.cpp:203-206 called in a loop over consecutive frames). */
int tvl1_solve_sequence_f32(tvl1_ctx *ctx, int nframes, const float *frames, float *u1, float *u2,
                            int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
int tvl1_solve_sequence_f64(tvl1_ctx *ctx, int nframes, const double *frames, double *u1, double *u2,
                            int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
/* 8-bit frames (what the reference's CLI reads from PGM / PNG files before it widens them to ofpix_t,
 * src/tvl1flow_main.cpp:175-176 via iio_read_image_float): the frames cross PCIe as bytes and are widened
 * on the device; flows in fp32.  Same bits as tvl1_solve_sequence_f32 on the widened frames. */
int tvl1_solve_sequence_u8(tvl1_ctx *ctx, int nframes, const unsigned char *frames, float *u1, float *u2,
                           int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
/* Same, DEVICE buffers (dense, 16-byte aligned).  A device-resident frame sequence is the call
 * below with dI1 = dI0 + nx*ny (inputs are read-only and may overlap).   Work is issued on the context's stream and
 * complete on return.  iters_out / errs_out are HOST pointers. */
int tvl1_solve_batch_dev_f32(tvl1_ctx *ctx, int npairs, const float *dI0, const float *dI1,
                             float *du1, float *du2, int nx, int ny, const tvl1_params *prm,
                             int *iters_out, double *errs_out);

/* -- row-band mode: ONE image pair split over several GPUs ----------------------- */
/* One context per rank/GPU.  Rank 0 makes an id (tvl1_band_unique_id), the caller distributes the
 * TVL1_NCCL_ID_BYTES bytes to every rank by whatever means it has (a process-group broadcast, MPI,
 * a file), every rank calls tvl1_band_init, then all ranks call tvl1_band_solve_* together with the
 * same full images and parameters.  Levels with at least min_split_rows rows are cut into row bands
 * whose 1-row halos (flow + dual variables) and error sum travel over NCCL every iteration; smaller
 * levels are solved redundantly on every rank.  Every rank receives the full flow.  (A negative
 * min_split_rows uses |min_split_rows| and takes the band code path even with a single rank.) */
#define TVL1_NCCL_ID_BYTES 128
int tvl1_band_unique_id(unsigned char *id_out /* [TVL1_NCCL_ID_BYTES] */);
int tvl1_band_init(tvl1_ctx *ctx, int rank, int world, const unsigned char *id /* [TVL1_NCCL_ID_BYTES] */);
int tvl1_band_solve_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2, int nx,
                        int ny, const tvl1_params *prm, int min_split_rows, int *iters_out,
                        double *errs_out);                                   /* HOST buffers */
int tvl1_band_solve_dev_f32(tvl1_ctx *ctx, const float *dI0, const float *dI1, float *du1, float *du2,
                            int nx, int ny, const tvl1_params *prm, int min_split_rows, int *iters_out,
                            double *errs_out);                               /* DEVICE buffers */
void tvl1_band_rows(int ny, int rank, int world, int *row_begin, int *row_end);  /* rows a rank owns */
/* How the per-iteration halo rows and error sums travel.  Default (1): inside the iteration kernel,
 * through peer memory mapped with CUDA IPC -- NVLink stores of the boundary rows plus an all-to-all
 * of the per-rank sums in mailboxes; no NCCL call and no extra kernel per iteration.  0: one NCCL
 * group (send/recv + all-reduce) per iteration (also the fallback when peer access is unavailable
 * or the box has more than 8 ranks).  tvl1_band_exchange_mode returns the mode in effect (-1: no
 * communicator yet). */
int tvl1_band_set_exchange(tvl1_ctx *ctx, int use_nccl_per_iteration);
int tvl1_band_exchange_mode(const tvl1_ctx *ctx);

/* -- one level: Dual_TVL1_optic_flow (src/tvl1flow.cpp:46-212) ----------------------------- */
/* u1,u2 are in/out (the initial flow is used, tvl1flow.cpp:94); nscales/zfactor of prm ignored;
 * iters_out/errs_out are [warps]. */
int tvl1_single_scale_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2,
                          int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);
int tvl1_single_scale_f64(tvl1_ctx *ctx, const double *I0, const double *I1, double *u1, double *u2,
                          int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out);

/* -- per-kernel hooks (HOST buffers; used by the parity tests) ----------------------------- */
void tvl1_zoom_size(int nx, int ny, int *nxx, int *nyy, double factor);
int tvl1_normalize_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *I0n, float *I1n,
                       int nx, int ny);
int tvl1_gaussian_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, double sigma);
int tvl1_zoom_out_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, double factor);
int tvl1_zoom_in_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, int nxx, int nyy,
                     double scale);
/* centered_gradient + the three warps + rho_c/grad, fused (tvl1flow.cpp:84,94-109) */
int tvl1_warp_f32(tvl1_ctx *ctx, const float *I0, const float *I1, const float *u1, const float *u2,
                  int nx, int ny, float *I1wx, float *I1wy, float *rho_c, float *grad);
/* exactly `iters` passes of the loop body (tvl1flow.cpp:114-181), no stopping test.
 * u1..p22 in/out; errs_out[iters] = mean squared update of each pass (may be NULL).  `grad` is
 * accepted for symmetry with the reference's data flow but the kernel recomputes it as
 * I1wx^2 + I1wy^2 (one plane less to read per iteration). */
int tvl1_iterate_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                     float *p22, const float *rho_c, const float *I1wx, const float *I1wy,
                     const float *grad, int nx, int ny, double tau, double lambda, double theta,
                     int iters, double *errs_out);

/* The same loop body through the cluster-resident kernel (whole while loop on chip): runs until the
 * stopping rule fires (epsilon < 0: never) or max_iter passes.  cluster = 0 picks the cluster size
 * automatically, otherwise 1/2/4/8/16 is forced (TVL1_ERR_ARG if the level does not fit).  grad is
 * recomputed from I1wx,I1wy on chip.  iters_out = passes done, errs_out[max_iter] = error of each. */
int tvl1_iterate_resident_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                              float *p22, const float *rho_c, const float *I1wx, const float *I1wy,
                              int nx, int ny, double tau, double lambda, double theta, double epsilon,
                              int max_iter, int cluster, int *iters_out, double *errs_out,
                              int *cluster_out);

/* The complete while loop of one warp step (src/tvl1flow.cpp:111-182) through the streaming kernels,
 * driven from the host: temporal_blocking = 0 uses k_iterate_t1 only (one iteration per launch),
 * 1 lets the device choose between k_iterate_t1 and the TMA-tiled k_iterate_tb (up to 4 iterations
 * per launch, exact replay when a block overshoots the stopping point), 2 additionally starts with a
 * full block.  epsilon < 0: run exactly max_iter iterations.  Outputs: iterations done, error of the
 * last one, kernel launches (while-loop turns) it took. */
int tvl1_iterate_loop_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                          float *p22, const float *rho_c, const float *I1wx, const float *I1wy, int nx,
                          int ny, double tau, double lambda, double theta, double epsilon, int max_iter,
                          int temporal_blocking, int *iters_out, double *err_out, int *launches_out);

/* -- bench hook: the fused iteration kernel alone on synthetic device-resident state -------- */
/* Runs `launches` iteration launches over `npairs` pairs of nx*ny (state and constants are
 * seeded pseudo-random, resident in HBM) and returns the CUDA-event time of those launches. */
int tvl1_bench_iterate(tvl1_ctx *ctx, int npairs, int nx, int ny, int launches, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* TVL1_B200_H */
