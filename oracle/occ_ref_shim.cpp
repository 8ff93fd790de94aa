// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference objects of the TV-L1 + occlusions solver
// (src/tvl1occflow.cpp, src/tvl1occflow_solvers.cpp, src/tvl1occflow_tv_rof_box.cpp and the shared
// operators / bicubic / zoom / utils TUs), compiled in place by oracle/Makefile into
// oracle/_ref/libocc_ref_f64.so.  Nothing of the reference is copied.
//
// WHY A SEPARATE LIBRARY.  tvl1occflow.h overloads the names of tvl1flow.h, and -- more important --
// Solver_wrt_chi reads eta1/eta2 from freshly new[]-ed memory without initialising it
// (src/tvl1occflow_solvers.cpp:246-262, the file's own #warning), so the reference's output depends on
// what the allocator hands back.  This shim pins the one defined reading of that code, the one a fresh
// process produces for images large enough to be served from zero pages: it replaces the global
// operator new[] of THIS shared object with a zero-filling one.  The reference sources are untouched;
// only their allocations become deterministic.  (The statics of Solver_wrt_u / Solver_wrt_chi are
// re-created whenever the image width changes, i.e. at every pyramid level, and then start from zero.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <fcntl.h>
#include <unistd.h>

#ifdef _OPENMP
#include <omp.h>
#endif

// (made local to this library by oracle/occ_ref.map: calls from the reference objects bind here)
void *operator new[](std::size_t n)
{
    void *p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }

#include "of.h"
#include "tvl1occflow.h"
#include "tvl1occflow_solvers.h"
#include "tvl1occflow_tv_rof_box.h"
#include "tvl1occflow_constants.h"
#include "utils.h"

// "Warping: %d, Iterations: %d, Error: %e" lines of src/tvl1occflow.cpp:305-309 -> iteration counts
template <class F>
static int capture_iters(F &&call, int *iters, double *errs, int cap)
{
    char path[] = "/tmp/occ_ref_stderr_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) { call(); return -1; }
    fflush(stderr);
    fflush(stdout);
    int saved = dup(2), saved1 = dup(1);
    dup2(fd, 2);
    int devnull = open("/dev/null", 1);
    if (devnull >= 0) dup2(devnull, 1);           // the reference prints "verbose" on stdout (:201-203)
    call();
    fflush(stderr);
    fflush(stdout);
    dup2(saved, 2);
    dup2(saved1, 1);
    close(saved);
    close(saved1);
    if (devnull >= 0) close(devnull);
    lseek(fd, 0, SEEK_SET);
    FILE *f = fdopen(fd, "r");
    int n = 0;
    char line[512];
    while (f && fgets(line, sizeof line, f)) {
        int w, it; double e;
        if (sscanf(line, "Warping: %d, Iterations: %d, Error: %lf", &w, &it, &e) == 3) {
            if (n < cap) { if (iters) iters[n] = it; if (errs) errs[n] = e; }
            n++;
        }
    }
    if (f) fclose(f); else close(fd);
    unlink(path);
    return n;
}

extern "C" {

int occ_ref_sizeof_pix(void) { return (int) sizeof(ofpix_t); }

// 1 if new[] inside this library is the zero-filling one (dirty a block, free it, take it again)
int occ_ref_new_is_zeroing(void)
{
    for (int rep = 0; rep < 4; rep++) {
        const int n = 1000 + 37 * rep;
        ofpix_t *a = new ofpix_t[n];
        for (int i = 0; i < n; i++) if (a[i] != 0) { delete [] a; return 0; }
        for (int i = 0; i < n; i++) a[i] = 7;
        delete [] a;
    }
    return 1;
}

void occ_ref_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void) n;
#endif
}

// Dual_TVL1_optic_flow_multiscale, src/tvl1occflow.h:111-129.  iters/errs: [nscales*warps], coarsest first.
int occ_ref_multiscale(const ofpix_t *I_1, const ofpix_t *I0, const ofpix_t *I1, const ofpix_t *filtI0, ofpix_t *u1,
                       ofpix_t *u2, ofpix_t *chi, int nx, int ny, double lambda, double alpha, double beta,
                       double theta, int nscales, double zfactor, int warps, double epsilon, int *iters, double *errs)
{
    return capture_iters([&] {
        Dual_TVL1_optic_flow_multiscale(const_cast<ofpix_t *>(I_1), const_cast<ofpix_t *>(I0), const_cast<ofpix_t *>(I1),
                                        const_cast<ofpix_t *>(filtI0), u1, u2, chi, nx, ny, lambda, alpha, beta, theta,
                                        nscales, zfactor, warps, epsilon, true);
    }, iters, errs, nscales * warps);
}

// Scalar_ROF_BoxCellCentered, src/tvl1occflow_tv_rof_box.h
void occ_ref_rof_box(ofpix_t *u, const ofpix_t *f, ofpix_t *p1, ofpix_t *p2, const ofpix_t *g, double lambda,
                     double omega, int nx, int ny, int niter)
{
    Scalar_ROF_BoxCellCentered(u, f, p1, p2, g, lambda, omega, nx, ny, niter);
}

void occ_ref_median3(ofpix_t *in, int nx, int ny) { me_median_filtering(in, nx, ny, 3); }

void occ_ref_solver_v(ofpix_t *u1, ofpix_t *u2, ofpix_t *v1, ofpix_t *v2, ofpix_t *chi, const ofpix_t *I1wx,
                      const ofpix_t *I1wy, const ofpix_t *I_1wx, const ofpix_t *I_1wy, const ofpix_t *rho1_c,
                      const ofpix_t *rho3_c, ofpix_t *Vfwd_1, ofpix_t *Vfwd_2, ofpix_t *Vbck_1, ofpix_t *Vbck_2,
                      const ofpix_t *grad1, const ofpix_t *grad3, double alpha, double theta, double lambda, int nx, int ny)
{
    Solver_wrt_v(u1, u2, v1, v2, chi, I1wx, I1wy, I_1wx, I_1wy, rho1_c, rho3_c, Vfwd_1, Vfwd_2, Vbck_1, Vbck_2, grad1,
                 grad3, alpha, theta, lambda, nx, ny);
}

} // extern "C"
