// Drop-in replacements for the reference's TV-L1 library entry points, with the reference's exact
// C++ signatures (and therefore mangled names; the reference cannot use extern "C" because
// src/tvl1occflow.h:63-79,111-129 overloads the same names):
//
//   void Dual_TVL1_optic_flow_multiscale(ofpix_t*, ofpix_t*, ofpix_t*, ofpix_t*, const int, const int,
//        const double, const double, const double, const int, const double, const int, const double,
//        const bool)                                                  src/tvl1flow.h:56-70
//   void Dual_TVL1_optic_flow(ofpix_t*, ofpix_t*, ofpix_t*, ofpix_t*, const int, const int,
//        const double, const double, const double, const int, const double, const bool)
//                                                                     src/tvl1flow.h:36-48
//
// ofpix_t is double as shipped (src/of.h:4-10): _Z31Dual_TVL1_optic_flow_multiscalePdS_S_S_iidddididb
// and _Z20Dual_TVL1_optic_flowPdS_S_S_iidddidb.  The float overloads below cover a reference built
// with OFPIX_DOUBLE commented out.  tvl1flow_main.cpp links against this object unchanged.
//
// The upstream IPOL C99 library the reference derives from has the same two functions with C linkage,
// float pixels and float scalars (3rdparty/tvl1flow_3/tvl1flow_lib.c:45-59 and :299-314); C callers of
// that library bind the unmangled names, exported at the end of this file.  (C++ allows one C-linkage
// member in an overload set, and its parameter list differs from the float overloads above it.)
//
// Host C++ only: everything GPU goes through the C ABI of include/tvl1_b200.h.
// Behaviour kept from the reference: void return, failures surface as exceptions
// (std::runtime_error("GaussianSmooth: sigma too large") from src/operators.cpp:520-522), verbose mode
// prints "Scale %d: %dx%d" (src/tvl1flow.cpp:285) and "Warping: %d, Iterations: %d, Error: %f"
// (:185-187) on stderr; I0/I1 are not modified; u1/u2 are fully overwritten (multiscale) or
// used as the initial flow (single level).
#include "../../include/tvl1_b200.h"
#include "../../include/hs_b200.h"
#include "../../include/occ_b200.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct ThreadCtx {
    tvl1_ctx *ctx = nullptr;
    int device = -1;           // device the context was made for
    ~ThreadCtx() { tvl1_destroy(ctx); }
};

thread_local int t_device = -1;    // tvl1_dropin_set_device(); -1: TVL1_DEVICE or 0

// One context per calling host thread: the reference is re-entrant, and batch sharding drives one
// host thread per GPU.  The device is the one the thread chose with tvl1_dropin_set_device(), else
// the process-wide TVL1_DEVICE, else 0.
tvl1_ctx *thread_ctx()
{
    static thread_local ThreadCtx tc;
    int dev = t_device;
    if (dev < 0) {
        dev = 0;
        if (const char *e = std::getenv("TVL1_DEVICE")) dev = std::atoi(e);
    }
    if (tc.ctx && tc.device != dev) { tvl1_destroy(tc.ctx); tc.ctx = nullptr; }
    if (!tc.ctx) {
        if (tvl1_create(dev, &tc.ctx) != TVL1_OK)
            throw std::runtime_error(std::string("tvl1_b200: ") + tvl1_last_error(nullptr));
        tc.device = dev;
    }
    return tc.ctx;
}

[[noreturn]] void raise(tvl1_ctx *ctx, int rc)
{
    if (rc == TVL1_ERR_SIGMA) throw std::runtime_error("GaussianSmooth: sigma too large");
    throw std::runtime_error(std::string("tvl1_b200: ") + tvl1_last_error(ctx));
}

void print_level(int warps, const int *iters, const double *errs)
{
    for (int w = 0; w < warps; w++)
        fprintf(stderr, "Warping: %d, Iterations: %d, Error: %f\n", w, iters[w], errs[w]);
}

template <typename T>
void multiscale(T *I0, T *I1, T *u1, T *u2, int nx, int ny, double tau, double lambda, double theta,
                int nscales, double zfactor, int warps, double epsilon, bool verbose)
{
    tvl1_ctx *ctx = thread_ctx();
    tvl1_params p{ tau, lambda, theta, nscales, zfactor, warps, epsilon };
    std::vector<int> iters((size_t) nscales * warps);
    std::vector<double> errs((size_t) nscales * warps);
    int rc;
    if (sizeof(T) == 8)
        rc = tvl1_solve_f64(ctx, (const double *) I0, (const double *) I1, (double *) u1, (double *) u2,
                            nx, ny, &p, iters.data(), errs.data());
    else
        rc = tvl1_solve_f32(ctx, (const float *) I0, (const float *) I1, (float *) u1, (float *) u2,
                            nx, ny, &p, iters.data(), errs.data());
    if (rc != TVL1_OK) raise(ctx, rc);
    if (verbose) {
        std::vector<int> sx(nscales), sy(nscales);
        sx[0] = nx; sy[0] = ny;
        for (int s = 1; s < nscales; s++) tvl1_zoom_size(sx[s - 1], sy[s - 1], &sx[s], &sy[s], zfactor);
        for (int s = nscales - 1; s >= 0; s--) {
            fprintf(stderr, "Scale %d: %dx%d\n", s, sx[s], sy[s]);
            const size_t k = (size_t) (nscales - 1 - s) * warps;
            print_level(warps, iters.data() + k, errs.data() + k);
        }
    }
}

template <typename T>
void single_scale(T *I0, T *I1, T *u1, T *u2, int nx, int ny, double tau, double lambda, double theta,
                  int warps, double epsilon, bool verbose)
{
    tvl1_ctx *ctx = thread_ctx();
    tvl1_params p{ tau, lambda, theta, 1, 0.5, warps, epsilon };
    std::vector<int> iters(warps);
    std::vector<double> errs(warps);
    int rc;
    if (sizeof(T) == 8)
        rc = tvl1_single_scale_f64(ctx, (const double *) I0, (const double *) I1, (double *) u1,
                                   (double *) u2, nx, ny, &p, iters.data(), errs.data());
    else
        rc = tvl1_single_scale_f32(ctx, (const float *) I0, (const float *) I1, (float *) u1,
                                   (float *) u2, nx, ny, &p, iters.data(), errs.data());
    if (rc != TVL1_OK) raise(ctx, rc);
    if (verbose) print_level(warps, iters.data(), errs.data());
}

// ---- pyramidal Horn-Schunck (src/horn_schunck.h:15-48) -------------------------------------------
// Verbose output as the reference prints it: the parameter banners (src/horn_schunck_pyramidal.cpp:
// 92-96, :274-278), "Scale: %d %dx%d" (:328) and, per warp, "Warping %d:" + "Iterations %d (%g)"
// (:118-120, :233-235).
void hs_print_level(int nx, int ny, double alpha, int warps, double TOL, int maxiter, const int *iters,
                    const double *errs)
{
    fprintf(stderr, "Single-scale Horn-Schunck of a %dx%d image\n\ta=%g nw=%d eps=%g mi=%d v=%d\n", nx, ny,
            alpha, warps, TOL, maxiter, 1);
    for (int w = 0; w < warps; w++) fprintf(stderr, "Warping %d:Iterations %d (%g)\n", w, iters[w], errs[w]);
}

template <typename T>
void hs_pyramidal(const T *I1, const T *I2, T *u, T *v, int nx, int ny, double alpha, int nscales, double zfactor,
                  int warps, double TOL, int maxiter, bool verbose)
{
    tvl1_ctx *ctx = thread_ctx();
    hs_params p{ alpha, nscales, zfactor, warps, TOL, maxiter };
    std::vector<int> iters((size_t) nscales * warps);
    std::vector<double> errs((size_t) nscales * warps);
    if (verbose)
        fprintf(stderr, "Multiscale Horn-Schunck of a %dx%d pair\n\ta=%g ns=%d zf=%g nw=%d eps=%g mi=%d\n", nx, ny,
                alpha, nscales, zfactor, warps, TOL, maxiter);
    int rc;
    if (sizeof(T) == 8)
        rc = hs_solve_f64(ctx, (const double *) I1, (const double *) I2, (double *) u, (double *) v, nx, ny, &p,
                          iters.data(), errs.data());
    else
        rc = hs_solve_f32(ctx, (const float *) I1, (const float *) I2, (float *) u, (float *) v, nx, ny, &p,
                          iters.data(), errs.data());
    if (rc != TVL1_OK) raise(ctx, rc);
    if (verbose) {
        std::vector<int> sx(nscales), sy(nscales);
        sx[0] = nx; sy[0] = ny;
        for (int s = 1; s < nscales; s++) tvl1_zoom_size(sx[s - 1], sy[s - 1], &sx[s], &sy[s], zfactor);
        for (int s = nscales - 1; s >= 0; s--) {
            fprintf(stderr, "Scale: %d %dx%d\n", s, sx[s], sy[s]);
            const size_t k = (size_t) (nscales - 1 - s) * warps;
            hs_print_level(sx[s], sy[s], alpha, warps, TOL, maxiter, iters.data() + k, errs.data() + k);
        }
    }
}

template <typename T>
void hs_one_level(const T *I1, const T *I2, T *u, T *v, int nx, int ny, double alpha, int warps, double TOL,
                  int maxiter, bool verbose)
{
    tvl1_ctx *ctx = thread_ctx();
    hs_params p{ alpha, 1, 0.5, warps, TOL, maxiter };
    std::vector<int> iters(warps);
    std::vector<double> errs(warps);
    int rc;
    if (sizeof(T) == 8)
        rc = hs_single_scale_f64(ctx, (const double *) I1, (const double *) I2, (double *) u, (double *) v, nx, ny,
                                 &p, iters.data(), errs.data());
    else
        rc = hs_single_scale_f32(ctx, (const float *) I1, (const float *) I2, (float *) u, (float *) v, nx, ny, &p,
                                 iters.data(), errs.data());
    if (rc != TVL1_OK) raise(ctx, rc);
    if (verbose) hs_print_level(nx, ny, alpha, warps, TOL, maxiter, iters.data(), errs.data());
}

} // namespace

// Selects the GPU for the calling thread's drop-in calls (declared in include/tvl1_b200.h).
extern "C" void tvl1_dropin_set_device(int device) { t_device = device; }

// ---- ofpix_t = double (reference as shipped) ----------------------------------------------------
void Dual_TVL1_optic_flow_multiscale(double *I0, double *I1, double *u1, double *u2, const int nxx,
                                     const int nyy, const double tau, const double lambda,
                                     const double theta, const int nscales, const double zfactor,
                                     const int warps, const double epsilon, const bool verbose)
{
    multiscale<double>(I0, I1, u1, u2, nxx, nyy, tau, lambda, theta, nscales, zfactor, warps, epsilon, verbose);
}

void Dual_TVL1_optic_flow(double *I0, double *I1, double *u1, double *u2, const int nx, const int ny,
                          const double tau, const double lambda, const double theta, const int warps,
                          const double epsilon, const bool verbose)
{
    single_scale<double>(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, verbose);
}

// ---- ofpix_t = float (src/of.h with OFPIX_DOUBLE commented out) ---------------------------------
void Dual_TVL1_optic_flow_multiscale(float *I0, float *I1, float *u1, float *u2, const int nxx,
                                     const int nyy, const double tau, const double lambda,
                                     const double theta, const int nscales, const double zfactor,
                                     const int warps, const double epsilon, const bool verbose)
{
    multiscale<float>(I0, I1, u1, u2, nxx, nyy, tau, lambda, theta, nscales, zfactor, warps, epsilon, verbose);
}

void Dual_TVL1_optic_flow(float *I0, float *I1, float *u1, float *u2, const int nx, const int ny,
                          const double tau, const double lambda, const double theta, const int warps,
                          const double epsilon, const bool verbose)
{
    single_scale<float>(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, verbose);
}

// ---- upstream C99 library (3rdparty/tvl1flow_3/tvl1flow_lib.c): C linkage, float everywhere ------
extern "C" void Dual_TVL1_optic_flow_multiscale(float *I0, float *I1, float *u1, float *u2,
                                                const int nxx, const int nyy, const float tau,
                                                const float lambda, const float theta,
                                                const int nscales, const float zfactor,
                                                const int warps, const float epsilon,
                                                const bool verbose)
{
    multiscale<float>(I0, I1, u1, u2, nxx, nyy, tau, lambda, theta, nscales, zfactor, warps, epsilon, verbose);
}

extern "C" void Dual_TVL1_optic_flow(float *I0, float *I1, float *u1, float *u2, const int nx,
                                     const int ny, const float tau, const float lambda,
                                     const float theta, const int warps, const float epsilon,
                                     const bool verbose)
{
    single_scale<float>(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, verbose);
}

// ---- pyramidal Horn-Schunck: src/horn_schunck.h:15-48, ofpix_t = double (as shipped) and float ------
// _Z22horn_schunck_pyramidalPKdS0_PdS1_iidididib, _Z25horn_schunck_optical_flowPKdS0_PdS1_iididib
void horn_schunck_pyramidal(const double *I1, const double *I2, double *u, double *v, const int nx, const int ny,
                            const double alpha, const int nscales, const double zfactor, const int warps,
                            const double TOL, const int maxiter, const bool verbose)
{
    hs_pyramidal<double>(I1, I2, u, v, nx, ny, alpha, nscales, zfactor, warps, TOL, maxiter, verbose);
}

void horn_schunck_optical_flow(const double *I1, const double *I2, double *u, double *v, const int nx,
                               const int ny, const double alpha, const int warps, const double TOL,
                               const int maxiter, const bool verbose)
{
    hs_one_level<double>(I1, I2, u, v, nx, ny, alpha, warps, TOL, maxiter, verbose);
}

void horn_schunck_pyramidal(const float *I1, const float *I2, float *u, float *v, const int nx, const int ny,
                            const double alpha, const int nscales, const double zfactor, const int warps,
                            const double TOL, const int maxiter, const bool verbose)
{
    hs_pyramidal<float>(I1, I2, u, v, nx, ny, alpha, nscales, zfactor, warps, TOL, maxiter, verbose);
}

void horn_schunck_optical_flow(const float *I1, const float *I2, float *u, float *v, const int nx, const int ny,
                               const double alpha, const int warps, const double TOL, const int maxiter,
                               const bool verbose)
{
    hs_one_level<float>(I1, I2, u, v, nx, ny, alpha, warps, TOL, maxiter, verbose);
}

// ---- TV-L1 with occlusions: src/tvl1occflow.h:63-79 and :111-129, ofpix_t = double (as shipped) -------
// _Z31Dual_TVL1_optic_flow_multiscalePdS_S_S_S_S_S_iiddddididb, _Z20Dual_TVL1_optic_flowPdS_S_S_S_S_S_iidddd idb:
// the seven-plane overloads of the TV-L1 names above.  Verbose mode prints what the reference prints:
// "verbose" on stdout once per level (src/tvl1occflow.cpp:192-194) and "Warping: %d, Iterations: %d,
// Error: %e" on stderr (:292-296).  The occlusion solver computes in fp64 (include/occ_b200.h).
namespace {

struct OccThreadCtx {
    occ_ctx *ctx = nullptr;
    int device = -1;
    ~OccThreadCtx() { occ_destroy(ctx); }
};

occ_ctx *occ_thread_ctx()
{
    static thread_local OccThreadCtx tc;
    int dev = t_device;
    if (dev < 0) {
        dev = 0;
        if (const char *e = std::getenv("TVL1_DEVICE")) dev = std::atoi(e);
    }
    if (tc.ctx && tc.device != dev) { occ_destroy(tc.ctx); tc.ctx = nullptr; }
    if (!tc.ctx) {
        if (occ_create(dev, &tc.ctx) != OCC_OK)
            throw std::runtime_error(std::string("occ_b200: ") + occ_last_error(nullptr));
        tc.device = dev;
    }
    return tc.ctx;
}

[[noreturn]] void occ_raise(occ_ctx *ctx, int rc)
{
    if (rc == OCC_ERR_SIGMA) throw std::runtime_error("GaussianSmooth: sigma too large");
    throw std::runtime_error(std::string("occ_b200: ") + occ_last_error(ctx));
}

void occ_print_level(int warps, const int *iters, const double *errs)
{
    printf("verbose\n");
    fflush(stdout);
    for (int w = 0; w < warps; w++)
        fprintf(stderr, "Warping: %d, Iterations: %d, Error: %e\n", w, iters[w], errs[w]);
}

} // namespace

void Dual_TVL1_optic_flow_multiscale(double *I_1, double *I0, double *I1, double *filtI0, double *u1, double *u2,
                                     double *chi, const int nxx, const int nyy, const double lambda,
                                     const double alpha, const double beta, const double theta, const int nscales,
                                     const double zfactor, const int warps, const double epsilon, const bool verbose)
{
    occ_ctx *ctx = occ_thread_ctx();
    occ_params p{ lambda, alpha, beta, theta, nscales, zfactor, warps, epsilon };
    std::vector<int> iters((size_t) nscales * warps);
    std::vector<double> errs((size_t) nscales * warps);
    const int rc = occ_solve_f64(ctx, I_1, I0, I1, filtI0, u1, u2, chi, nxx, nyy, &p, iters.data(), errs.data());
    if (rc != OCC_OK) occ_raise(ctx, rc);
    if (verbose)
        for (int k = 0; k < nscales; k++) occ_print_level(warps, iters.data() + (size_t) k * warps, errs.data() + (size_t) k * warps);
}

void Dual_TVL1_optic_flow(double *I_1, double *I0, double *I1, double *filtI0, double *u1, double *u2, double *chi,
                          const int nx, const int ny, const double lambda, const double alpha, const double beta,
                          const double theta, const int warps, const double epsilon, const bool verbose)
{
    occ_ctx *ctx = occ_thread_ctx();
    occ_params p{ lambda, alpha, beta, theta, 1, 0.5, warps, epsilon };
    std::vector<int> iters(warps);
    std::vector<double> errs(warps);
    const int rc = occ_single_scale_f64(ctx, I_1, I0, I1, filtI0, u1, u2, chi, nx, ny, &p, iters.data(), errs.data());
    if (rc != OCC_OK) occ_raise(ctx, rc);
    if (verbose) occ_print_level(warps, iters.data(), errs.data());
}
