// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference objects (compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/).  It only includes the
// reference's own headers; no reference source is copied into this repository.
// Every entry point is the reference function of the same name with `ref_` prefixed:
//
//   Dual_TVL1_optic_flow_multiscale   src/tvl1flow.h:56-70
//   Dual_TVL1_optic_flow              src/tvl1flow.h:36-48
//   image_normalization_2             src/utils.h:27
//   gaussian                          src/operators.h:128
//   zoom_size / zoom_out / zoom_in    src/zoom.h:20,32,57
//   centered_gradient                 src/operators.h:93
//   bicubic_interpolation_warp        src/bicubic_interpolation.h:45
//   divergence / forward_gradient     src/operators.h:29,42
//   horn_schunck_pyramidal            src/horn_schunck.h:35-48      
//   horn_schunck_optical_flow         src/horn_schunck.h:15-26
//
// The pixel type is whatever `ofpix_t` the reference objects were built with
// (double as shipped; float for the cross-check build, see oracle/Makefile).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "of.h"
#include "tvl1flow.h"
#include "operators.h"
#include "bicubic_interpolation.h"
#include "zoom.h"
#include "utils.h"
#include "horn_schunck.h"

// Runs a callable with stderr redirected to a temp file and parses the reference's
// "Warping: %d, Iterations: %d, Error: %f" lines (src/tvl1flow.cpp:185-187), which are
// the only place iteration counts are observable.  Returns number of lines parsed.
template <class F>
static int capture_iters(F &&call, int *iters, double *errs, int cap)
{
    char path[] = "/tmp/tvl1_ref_stderr_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) { call(); return -1; }
    fflush(stderr);
    int saved = dup(2);
    dup2(fd, 2);
    call();
    fflush(stderr);
    dup2(saved, 2);
    close(saved);
    lseek(fd, 0, SEEK_SET);
    FILE *f = fdopen(fd, "r");
    int n = 0;
    char line[512];
    while (f && fgets(line, sizeof line, f)) {
        int w, it; double e;
        if (sscanf(line, "Warping: %d, Iterations: %d, Error: %lf", &w, &it, &e) == 3) {
            if (n < cap) { if (iters) iters[n] = it; if (errs) errs[n] = e; }
            n++;
        } else if (const char *h = strstr(line, "Iterations ")) {
            // Horn-Schunck: "Warping %d:Iterations %d (%g)" (src/horn_schunck_pyramidal.cpp:118-120,233-235)
            if (sscanf(h, "Iterations %d (%lf)", &it, &e) == 2) {
                if (n < cap) { if (iters) iters[n] = it; if (errs) errs[n] = e; }
                n++;
            }
        }
    }
    if (f) fclose(f); else close(fd);
    unlink(path);
    return n;
}

extern "C" {

int ref_sizeof_pix(void) { return (int) sizeof(ofpix_t); }

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

void ref_multiscale(ofpix_t *I0, ofpix_t *I1, ofpix_t *u1, ofpix_t *u2, int nx, int ny,
                    double tau, double lambda, double theta, int nscales, double zfactor,
                    int warps, double epsilon, int verbose)
{
    Dual_TVL1_optic_flow_multiscale(I0, I1, u1, u2, nx, ny, tau, lambda, theta, nscales,
                                    zfactor, warps, epsilon, verbose != 0);
}

void ref_single_scale(ofpix_t *I0, ofpix_t *I1, ofpix_t *u1, ofpix_t *u2, int nx, int ny,
                      double tau, double lambda, double theta, int warps, double epsilon,
                      int verbose)
{
    Dual_TVL1_optic_flow(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon,
                         verbose != 0);
}

// iters/errs are filled in the order the reference prints them: coarsest scale first,
// warps 0..warps-1 inside each scale.  cap = nscales*warps.
int ref_multiscale_iters(ofpix_t *I0, ofpix_t *I1, ofpix_t *u1, ofpix_t *u2, int nx, int ny,
                         double tau, double lambda, double theta, int nscales, double zfactor,
                         int warps, double epsilon, int *iters, double *errs, int cap)
{
    return capture_iters([&] {
        Dual_TVL1_optic_flow_multiscale(I0, I1, u1, u2, nx, ny, tau, lambda, theta, nscales,
                                        zfactor, warps, epsilon, true);
    }, iters, errs, cap);
}

int ref_single_scale_iters(ofpix_t *I0, ofpix_t *I1, ofpix_t *u1, ofpix_t *u2, int nx, int ny,
                           double tau, double lambda, double theta, int warps, double epsilon,
                           int *iters, double *errs, int cap)
{
    return capture_iters([&] {
        Dual_TVL1_optic_flow(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, true);
    }, iters, errs, cap);
}

void ref_normalize(const ofpix_t *I0, const ofpix_t *I1, ofpix_t *I0n, ofpix_t *I1n, int size)
{
    image_normalization_2(I0, I1, I0n, I1n, size);
}

// returns 0, or 1 if the reference threw ("GaussianSmooth: sigma too large")
int ref_gaussian(ofpix_t *I, int nx, int ny, double sigma)
{
    try { gaussian(I, nx, ny, sigma); } catch (...) { return 1; }
    return 0;
}

void ref_zoom_size(int nx, int ny, int *nxx, int *nyy, double factor)
{
    zoom_size(nx, ny, nxx, nyy, factor);
}

int ref_zoom_out(const ofpix_t *I, ofpix_t *Iout, int nx, int ny, double factor)
{
    try { zoom_out(I, Iout, nx, ny, factor); } catch (...) { return 1; }
    return 0;
}

void ref_zoom_in(const ofpix_t *I, ofpix_t *Iout, int nx, int ny, int nxx, int nyy)
{
    zoom_in(I, Iout, nx, ny, nxx, nyy);
}

void ref_centered_gradient(const ofpix_t *I, ofpix_t *dx, ofpix_t *dy, int nx, int ny)
{
    centered_gradient(I, dx, dy, nx, ny, 1);
}

void ref_warp(const ofpix_t *I, const ofpix_t *u, const ofpix_t *v, ofpix_t *out, int nx, int ny,
              int border_out)
{
    bicubic_interpolation_warp(I, u, v, out, nx, ny, border_out != 0);
}

void ref_divergence(const ofpix_t *v1, const ofpix_t *v2, ofpix_t *div, int nx, int ny)
{
    divergence(v1, v2, div, nx, ny);
}

void ref_forward_gradient(const ofpix_t *f, ofpix_t *fx, ofpix_t *fy, int nx, int ny)
{
    forward_gradient(f, fx, fy, nx, ny);
}

// ---- pyramidal Horn-Schunck (one thread: the reference's SOR sweep is only well defined then) ----
int ref_hs_multiscale_iters(const ofpix_t *I1, const ofpix_t *I2, ofpix_t *u, ofpix_t *v, int nx, int ny,
                            double alpha, int nscales, double zfactor, int warps, double TOL, int maxiter,
                            int *iters, double *errs, int cap)
{
    return capture_iters([&] {
        horn_schunck_pyramidal(I1, I2, u, v, nx, ny, alpha, nscales, zfactor, warps, TOL, maxiter, true);
    }, iters, errs, cap);
}

int ref_hs_single_scale_iters(const ofpix_t *I1, const ofpix_t *I2, ofpix_t *u, ofpix_t *v, int nx, int ny,
                              double alpha, int warps, double TOL, int maxiter, int *iters, double *errs,
                              int cap)
{
    return capture_iters([&] {
        horn_schunck_optical_flow(I1, I2, u, v, nx, ny, alpha, warps, TOL, maxiter, true);
    }, iters, errs, cap);
}

void ref_hs_multiscale(const ofpix_t *I1, const ofpix_t *I2, ofpix_t *u, ofpix_t *v, int nx, int ny,
                       double alpha, int nscales, double zfactor, int warps, double TOL, int maxiter)
{
    horn_schunck_pyramidal(I1, I2, u, v, nx, ny, alpha, nscales, zfactor, warps, TOL, maxiter, false);
}

} // extern "C"
