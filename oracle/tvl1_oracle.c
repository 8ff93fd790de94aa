/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
 *
 * CPU restatement (plain C99) of the TV-L1 hot path of 12334zq/optical-flow-1, used as the
 * parity oracle for the CUDA implementation.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load the library built from this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 *   (1) against the golden vectors in tests/golden/ that were produced by the unmodified
 *       reference objects (tests/golden/make_golden.py), and
 *   (2) when oracle/_ref/ exists (this container, and the GPU box via the shipped .so), directly
 *       against the reference functions, bit-for-bit in the double build.
 *
 * Storage type PIX is `double` (reference as shipped, src/of.h:4-10) or `float`
 * (-DORC_FLOAT; the reference's own float build keeps every temporary in double, and so
 * does this file: only loads/stores are PIX).
 *
 * All citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_FLOAT
typedef float PIX;
/* the reference is C++: hypot(float, float) resolves to the float overload (src/tvl1flow.cpp:172) */
#define ORC_HYPOT(a, b) hypotf((a), (b))
#define ORC_SQRT_PIX(a) sqrtf(a)
#else
typedef double PIX;
#define ORC_HYPOT(a, b) hypot((a), (b))
#define ORC_SQRT_PIX(a) sqrt(a)
#endif

#define ORC_MAX_ITERATIONS 300      /* src/tvl1flow.cpp:22 */
#define ORC_PRESMOOTHING_SIGMA 0.8  /* src/tvl1flow.cpp:23 */
#define ORC_GRAD_IS_ZERO 1E-10      /* src/tvl1flow.cpp:24 */
#define ORC_ZOOM_SIGMA_ZERO 0.6     /* src/zoom.cpp:15 */
#define ORC_GAUSS_WINDOW 5          /* src/operators.h:120 */

int orc_sizeof_pix(void) { return (int) sizeof(PIX); }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * (a) pyramid
 * ---------------------------------------------------------------------------------------- */

/* src/utils.cpp:509-525 (getminmax) + :283-326 (image_normalization_2): joint min/max of both
 * images; if max-min > 0 map affinely onto [0,255] with "255*(I-min)/den", else copy. */
void orc_normalize(const PIX *I0, const PIX *I1, PIX *I0n, PIX *I1n, int size)
{
    PIX mn = I0[0], mx = I0[0];
    for (int i = 1; i < size; i++) {
        if (I0[i] < mn) mn = I0[i];
        if (I0[i] > mx) mx = I0[i];
    }
    PIX mn2 = I1[0], mx2 = I1[0];
    for (int i = 1; i < size; i++) {
        if (I1[i] < mn2) mn2 = I1[i];
        if (I1[i] > mx2) mx2 = I1[i];
    }
    if (mx2 > mx) mx = mx2;
    if (mn2 < mn) mn = mn2;
    const PIX den = mx - mn;
    if (den > 0) {
        #pragma omp parallel for
        for (int i = 0; i < size; i++) {
            I0n[i] = 255.0 * (I0[i] - mn) / den;
            I1n[i] = 255.0 * (I1[i] - mn) / den;
        }
    } else {
        for (int i = 0; i < size; i++) { I0n[i] = I0[i]; I1n[i] = I1[i]; }
    }
}

/* 1-D kernel of src/operators.cpp:525-539: B[i] = exp(-i*i/(2 s^2)) / (s*sqrt(2*3.1415926)),
 * normalised by 2*sum(B) - B[0].  radius+1 taps, radius+1 = (int)(5*sigma)+1 (:515).
 * Returns the number of taps (size); B must hold at least that many doubles. */
int orc_gaussian_taps(double sigma, double *B, int cap)
{
    const double den = 2 * sigma * sigma;
    const int size = (int) (ORC_GAUSS_WINDOW * sigma) + 1;
    if (size > cap) return -size;
    for (int i = 0; i < size; i++)
        B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += B[i];
    norm *= 2;
    norm -= B[0];
    for (int i = 0; i < size; i++) B[i] /= norm;
    return size;
}

/* One pass of src/operators.cpp:541-578 (rows) / :581-619 (columns) over a strided line.
 * The reference pads a scratch line with `size` samples on each side (REFLECTING case,
 * :557-562): left pad x=-k -> I[k] (edge sample not repeated), right pad x=n-1+k -> I[n-k]
 * (edge sample repeated once: x=n -> I[n-1]).  Sum order as in :573-576. */
static void orc_gauss_line(PIX *line, int stride, int n, const double *B, int size, PIX *R)
{
    const int bd = n + size;
    for (int i = size; i < bd; i++) R[i] = line[(i - size) * stride];
    for (int i = 0, j = bd; i < size; i++, j++) {
        R[i] = line[(size - i) * stride];
        R[j] = line[(n - i - 1) * stride];
    }
    for (int i = size; i < bd; i++) {
        double sum = B[0] * R[i];
        for (int j = 1; j < size; j++) sum += B[j] * (R[i - j] + R[i + j]);
        line[(i - size) * stride] = sum;
    }
}

/* src/operators.cpp:506-624 with the defaults of src/operators.h:120-134 (reflecting boundary,
 * window 5).  In place; rows first, then columns.  Returns 1 where the reference throws
 * "GaussianSmooth: sigma too large" (:520-522, tests the width only). */
int orc_gaussian(PIX *I, int xdim, int ydim, double sigma)
{
    double B[64];
    const int size = orc_gaussian_taps(sigma, B, 64);
    if (size < 0) return 2;
    if (size > xdim) return 1;
    PIX *R = (PIX *) malloc(sizeof(PIX) * (size_t) (2 * size + (xdim > ydim ? xdim : ydim)));
    for (int k = 0; k < ydim; k++) orc_gauss_line(I + (size_t) k * xdim, 1, xdim, B, size, R);
    for (int k = 0; k < xdim; k++) orc_gauss_line(I + k, xdim, ydim, B, size, R);
    free(R);
    return 0;
}

/* src/zoom.cpp:22-34 */
void orc_zoom_size(int nx, int ny, int *nxx, int *nyy, double factor)
{
    *nxx = (int) (nx * factor + 0.5);
    *nyy = (int) (ny * factor + 0.5);
}

/* src/bicubic_interpolation.cpp:108-123: Keys / Catmull-Rom cubic through v[0..3] at offset x
 * from v[1]. */
static double orc_cubic(const double v[4], double x)
{
    return v[1] + 0.5 * x * (v[2] - v[0]
                             + x * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3]
                                    + x * (3.0 * (v[1] - v[2]) + v[3] - v[0])));
}

/* src/bicubic_interpolation.cpp:24-39 */
static int orc_neumann(int x, int nx, int *out)
{
    if (x < 0) { x = 0; *out = 1; }
    else if (x >= nx) { x = nx - 1; *out = 1; }
    return x;
}

/* src/bicubic_interpolation.cpp:153-245 with BOUNDARY_CONDITION 0 (:14).  Base index is the
 * C truncation (int)uu; the step direction follows the sign of the coordinate; the "minus"
 * row uses sx rather than sy (:173, upstream quirk, kept).  Interpolation is along y inside
 * each of the four x-columns first, then along x (:137-144, :236-240). */
double orc_bicubic_at(const PIX *in, double uu, double vv, int nx, int ny, int border_out)
{
    const int sx = (uu < 0) ? -1 : 1;
    const int sy = (vv < 0) ? -1 : 1;
    int out = 0;
    const int x   = orc_neumann((int) uu, nx, &out);
    const int y   = orc_neumann((int) vv, ny, &out);
    const int mx  = orc_neumann((int) uu - sx, nx, &out);
    const int my  = orc_neumann((int) vv - sx, ny, &out);
    const int dx  = orc_neumann((int) uu + sx, nx, &out);
    const int dy  = orc_neumann((int) vv + sy, ny, &out);
    const int ddx = orc_neumann((int) uu + 2 * sx, nx, &out);
    const int ddy = orc_neumann((int) vv + 2 * sy, ny, &out);
    if (out && border_out) return 0.0;

    const int xs[4] = { mx, x, dx, ddx };
    const int ys[4] = { my, y, dy, ddy };
    double col[4];
    for (int a = 0; a < 4; a++) {
        double p[4];
        for (int b = 0; b < 4; b++) p[b] = in[xs[a] + nx * ys[b]];
        col[a] = orc_cubic(p, vv - y);
    }
    return orc_cubic(col, uu - x);
}

/* src/zoom.cpp:41-78: copy, blur with sigma = 0.6*sqrt(1/f^2 - 1), then sample the blurred
 * image at (j/f, i/f) with clamped neighbours (border_out = false). */
int orc_zoom_out(const PIX *I, PIX *Iout, int nx, int ny, double factor)
{
    PIX *Is = (PIX *) malloc(sizeof(PIX) * (size_t) nx * ny);
    memcpy(Is, I, sizeof(PIX) * (size_t) nx * ny);
    int nxx, nyy;
    orc_zoom_size(nx, ny, &nxx, &nyy, factor);
    const double sigma = ORC_ZOOM_SIGMA_ZERO * sqrt(1.0 / (factor * factor) - 1.0);
    const int rc = orc_gaussian(Is, nx, ny, sigma);
    if (rc) { free(Is); return rc; }
    #pragma omp parallel for
    for (int i1 = 0; i1 < nyy; i1++)
        for (int j1 = 0; j1 < nxx; j1++) {
            const double i2 = i1 / factor;
            const double j2 = j1 / factor;
            Iout[i1 * nxx + j1] = orc_bicubic_at(Is, j2, i2, nx, ny, 0);
        }
    free(Is);
    return 0;
}

/* src/zoom.cpp:132-155: sample at (j/(nxx/nx), i/(nyy/ny)), clamped neighbours. */
void orc_zoom_in(const PIX *I, PIX *Iout, int nx, int ny, int nxx, int nyy)
{
    const double factorx = ((double) nxx / nx);
    const double factory = ((double) nyy / ny);
    #pragma omp parallel for
    for (int i1 = 0; i1 < nyy; i1++)
        for (int j1 = 0; j1 < nxx; j1++) {
            const double i2 = i1 / factory;
            const double j2 = j1 / factorx;
            Iout[i1 * nxx + j1] = orc_bicubic_at(I, j2, i2, nx, ny, 0);
        }
}

/* ------------------------------------------------------------------------------------------
 * (b) warp + precompute
 * ---------------------------------------------------------------------------------------- */

static int orc_clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

/* src/operators.cpp:335-406 with nz = 1.  Body, edge (:360-384) and corner (:388-404) cases all
 * reduce to half the difference of the index-clamped neighbours. */
void orc_centered_gradient(const PIX *in, PIX *dx, PIX *dy, int nx, int ny)
{
    #pragma omp parallel for
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int k = i * nx + j;
            dx[k] = 0.5 * (in[i * nx + orc_clampi(j + 1, nx)] - in[i * nx + orc_clampi(j - 1, nx)]);
            dy[k] = 0.5 * (in[orc_clampi(i + 1, ny) * nx + j] - in[orc_clampi(i - 1, ny) * nx + j]);
        }
}

/* src/bicubic_interpolation.cpp:352-374 */
void orc_warp(const PIX *in, const PIX *u, const PIX *v, PIX *out, int nx, int ny, int border_out)
{
    #pragma omp parallel for
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int p = i * nx + j;
            const double uu = j + u[p];
            const double vv = i + v[p];
            out[p] = orc_bicubic_at(in, uu, vv, nx, ny, border_out);
        }
}

/* src/tvl1flow.cpp:94-109: the three warps followed by grad = I1wx^2 + I1wy^2 and
 * rho_c = I1w - I1wx*u1 - I1wy*u2 - I0. */
void orc_warp_precompute(const PIX *I0, const PIX *I1, const PIX *I1x, const PIX *I1y,
                         const PIX *u1, const PIX *u2, PIX *I1w, PIX *I1wx, PIX *I1wy,
                         PIX *rho_c, PIX *grad, int nx, int ny)
{
    const int size = nx * ny;
    orc_warp(I1, u1, u2, I1w, nx, ny, 1);
    orc_warp(I1x, u1, u2, I1wx, nx, ny, 1);
    orc_warp(I1y, u1, u2, I1wy, nx, ny, 1);
    #pragma omp parallel for
    for (int i = 0; i < size; i++) {
        const double Ix2 = I1wx[i] * I1wx[i];
        const double Iy2 = I1wy[i] * I1wy[i];
        grad[i] = (Ix2 + Iy2);
        rho_c[i] = (I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i]);
    }
}

/* ------------------------------------------------------------------------------------------
 * (c) primal-dual iteration
 * ---------------------------------------------------------------------------------------- */

/* src/operators.cpp:35-78: backward differences; v1[-1] and v2[-1] read as 0, and the
 * "+v1" term is dropped on the last column, the "+v2" term on the last row (:58-77). */
void orc_divergence(const PIX *v1, const PIX *v2, PIX *div, int nx, int ny)
{
    #pragma omp parallel for
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int p = i * nx + j;
            double d;
            /* association order follows the reference's per-case expressions */
            if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
                const double v1x = v1[p] - v1[p - 1];
                const double v2y = v2[p] - v2[p - nx];
                d = v1x + v2y;
            } else {
                /* terms in the order the reference writes them: v1[p], -v1[p-1], v2[p],
                 * -v2[p-nx]; evaluated in the storage type like the reference's expressions */
                PIX a = 0;
                int first = 1;
                if (j < nx - 1) { a = v1[p]; first = 0; }
                if (j > 0) { a = first ? -v1[p - 1] : a - v1[p - 1]; first = 0; }
                if (i < ny - 1) { a = first ? v2[p] : a + v2[p]; first = 0; }
                if (i > 0) { a = first ? -v2[p - nx] : a - v2[p - nx]; first = 0; }
                d = a;
            }
            div[p] = d;
        }
}

/* src/operators.cpp:86-125: forward differences, 0 on the last column (fx) / last row (fy). */
void orc_forward_gradient(const PIX *f, PIX *fx, PIX *fy, int nx, int ny)
{
    #pragma omp parallel for
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int p = i * nx + j;
            fx[p] = (j < nx - 1) ? f[p + 1] - f[p] : 0;
            fy[p] = (i < ny - 1) ? f[p + nx] - f[p] : 0;
        }
}

typedef struct {
    PIX *v1, *v2, *div_p1, *div_p2, *u1x, *u1y, *u2x, *u2y;
} orc_iter_ws;

static void orc_ws_alloc(orc_iter_ws *w, size_t size)
{
    PIX **f = (PIX **) w;
    for (int k = 0; k < 8; k++) f[k] = (PIX *) malloc(sizeof(PIX) * size);
}

static void orc_ws_free(orc_iter_ws *w)
{
    PIX **f = (PIX **) w;
    for (int k = 0; k < 8; k++) free(f[k]);
}

/* One pass of the loop body src/tvl1flow.cpp:114-181; returns the mean squared update (:162). */
static double orc_iterate_once(PIX *u1, PIX *u2, PIX *p11, PIX *p12, PIX *p21, PIX *p22,
                               const PIX *rho_c, const PIX *I1wx, const PIX *I1wy,
                               const PIX *grad, int nx, int ny, double tau, double lambda,
                               double theta, orc_iter_ws *w)
{
    const int size = nx * ny;
    const double l_t = lambda * theta;

    /* thresholding step TH, :117-143 */
    #pragma omp parallel for
    for (int i = 0; i < size; i++) {
        const double rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
        double d1, d2;
        if (rho < -l_t * grad[i]) {
            d1 = l_t * I1wx[i];
            d2 = l_t * I1wy[i];
        } else if (rho > l_t * grad[i]) {
            d1 = -l_t * I1wx[i];
            d2 = -l_t * I1wy[i];
        } else if (grad[i] < ORC_GRAD_IS_ZERO) {
            d1 = d2 = 0;
        } else {
            const double fi = -rho / grad[i];
            d1 = fi * I1wx[i];
            d2 = fi * I1wy[i];
        }
        w->v1[i] = u1[i] + d1;
        w->v2[i] = u2[i] + d2;
    }

    /* :146-147 */
    orc_divergence(p11, p12, w->div_p1, nx, ny);
    orc_divergence(p21, p22, w->div_p2, nx, ny);

    /* primal update and error, :150-162 */
    double error = 0.0;
    #pragma omp parallel for reduction(+:error)
    for (int i = 0; i < size; i++) {
        const double u1k = u1[i];
        const double u2k = u2[i];
        u1[i] = w->v1[i] + theta * w->div_p1[i];
        u2[i] = w->v2[i] + theta * w->div_p2[i];
        error += (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
    }
    error /= size;

    /* :165-166 */
    orc_forward_gradient(u1, w->u1x, w->u1y, nx, ny);
    orc_forward_gradient(u2, w->u2x, w->u2y, nx, ny);

    /* dual update, :169-181 */
    #pragma omp parallel for
    for (int i = 0; i < size; i++) {
        const double taut = tau / theta;
        const double g1 = ORC_HYPOT(w->u1x[i], w->u1y[i]);
        const double g2 = ORC_HYPOT(w->u2x[i], w->u2y[i]);
        const double ng1 = 1.0 + taut * g1;
        const double ng2 = 1.0 + taut * g2;
        p11[i] = (p11[i] + taut * w->u1x[i]) / ng1;
        p12[i] = (p12[i] + taut * w->u1y[i]) / ng1;
        p21[i] = (p21[i] + taut * w->u2x[i]) / ng2;
        p22[i] = (p22[i] + taut * w->u2y[i]) / ng2;
    }
    return error;
}

/* Exactly `iters` passes of the loop body, no stopping test; errs[k] = mean squared update of
 * pass k (may be NULL).  Test hook for the fused CUDA iteration kernel. */
void orc_iterate(PIX *u1, PIX *u2, PIX *p11, PIX *p12, PIX *p21, PIX *p22, const PIX *rho_c,
                 const PIX *I1wx, const PIX *I1wy, const PIX *grad, int nx, int ny, double tau,
                 double lambda, double theta, int iters, double *errs)
{
    orc_iter_ws w;
    orc_ws_alloc(&w, (size_t) nx * ny);
    for (int k = 0; k < iters; k++) {
        const double e = orc_iterate_once(u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, grad,
                                          nx, ny, tau, lambda, theta, &w);
        if (errs) errs[k] = e;
    }
    orc_ws_free(&w);
}

/* src/tvl1flow.cpp:46-212.  u1,u2 are in/out.  iters/errs (each `warps` long, may be NULL)
 * receive what the reference prints in verbose mode (:184-188). */
void orc_single_scale(const PIX *I0, const PIX *I1, PIX *u1, PIX *u2, int nx, int ny, double tau,
                      double lambda, double theta, int warps, double epsilon, int *iters,
                      double *errs)
{
    const size_t size = (size_t) nx * ny;
    PIX *buf = (PIX *) malloc(sizeof(PIX) * size * 11);
    PIX *I1x = buf, *I1y = buf + size, *I1w = buf + 2 * size, *I1wx = buf + 3 * size,
        *I1wy = buf + 4 * size, *rho_c = buf + 5 * size, *grad = buf + 6 * size,
        *p11 = buf + 7 * size, *p12 = buf + 8 * size, *p21 = buf + 9 * size,
        *p22 = buf + 10 * size;
    orc_iter_ws w;
    orc_ws_alloc(&w, size);

    orc_centered_gradient(I1, I1x, I1y, nx, ny);              /* :84 */
    for (size_t i = 0; i < size; i++) p11[i] = p12[i] = p21[i] = p22[i] = 0.0;   /* :87-90 */

    for (int warpings = 0; warpings < warps; warpings++) {    /* :92 */
        orc_warp_precompute(I0, I1, I1x, I1y, u1, u2, I1w, I1wx, I1wy, rho_c, grad, nx, ny);
        int n = 0;
        double error = INFINITY;
        while (error > epsilon * epsilon && n < ORC_MAX_ITERATIONS) {   /* :113 */
            n++;
            error = orc_iterate_once(u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, grad, nx, ny,
                                     tau, lambda, theta, &w);
        }
        if (iters) iters[warpings] = n;
        if (errs) errs[warpings] = error;
    }
    orc_ws_free(&w);
    free(buf);
}

/* src/tvl1flow.cpp:219-328.  iters/errs are [nscales*warps], filled coarsest scale first
 * (the order of the reference's verbose output).  Returns 0, or 1 where the reference would
 * throw from gaussian(). */
int orc_multiscale(const PIX *I0, const PIX *I1, PIX *u1, PIX *u2, int nxx, int nyy, double tau,
                   double lambda, double theta, int nscales, double zfactor, int warps,
                   double epsilon, int *iters, double *errs)
{
    const int size = nxx * nyy;
    PIX **I0s = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **I1s = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **u1s = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **u2s = (PIX **) calloc(nscales, sizeof(PIX *));
    int *nx = (int *) calloc(nscales, sizeof(int));
    int *ny = (int *) calloc(nscales, sizeof(int));
    int rc = 0, built = 1;

    I0s[0] = (PIX *) malloc(sizeof(PIX) * size);
    I1s[0] = (PIX *) malloc(sizeof(PIX) * size);
    u1s[0] = u1; u2s[0] = u2; nx[0] = nxx; ny[0] = nyy;

    orc_normalize(I0, I1, I0s[0], I1s[0], size);                                  /* :255 */
    rc |= orc_gaussian(I0s[0], nx[0], ny[0], ORC_PRESMOOTHING_SIGMA);             /* :258 */
    rc |= orc_gaussian(I1s[0], nx[0], ny[0], ORC_PRESMOOTHING_SIGMA);             /* :259 */

    for (int s = 1; s < nscales && !rc; s++) {                                    /* :262-275 */
        orc_zoom_size(nx[s - 1], ny[s - 1], &nx[s], &ny[s], zfactor);
        const size_t sizes = (size_t) nx[s] * ny[s];
        I0s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        I1s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        u1s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        u2s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        built = s + 1;
        rc |= orc_zoom_out(I0s[s - 1], I0s[s], nx[s - 1], ny[s - 1], zfactor);
        rc |= orc_zoom_out(I1s[s - 1], I1s[s], nx[s - 1], ny[s - 1], zfactor);
    }

    if (!rc) {
        for (int i = 0; i < nx[nscales - 1] * ny[nscales - 1]; i++)              /* :278-280 */
            u1s[nscales - 1][i] = u2s[nscales - 1][i] = 0.0;

        for (int s = nscales - 1; s >= 0; s--) {                                  /* :283-310 */
            const int k = nscales - 1 - s;
            orc_single_scale(I0s[s], I1s[s], u1s[s], u2s[s], nx[s], ny[s], tau, lambda, theta,
                             warps, epsilon, iters ? iters + k * warps : 0,
                             errs ? errs + k * warps : 0);
            if (!s) break;
            orc_zoom_in(u1s[s], u1s[s - 1], nx[s], ny[s], nx[s - 1], ny[s - 1]);
            orc_zoom_in(u2s[s], u2s[s - 1], nx[s], ny[s], nx[s - 1], ny[s - 1]);
            for (int i = 0; i < nx[s - 1] * ny[s - 1]; i++) {
                u1s[s - 1][i] *= 1.0 / zfactor;
                u2s[s - 1][i] *= 1.0 / zfactor;
            }
        }
    }

    for (int i = 1; i < built; i++) { free(I0s[i]); free(I1s[i]); free(u1s[i]); free(u2s[i]); }
    free(I0s[0]); free(I1s[0]);
    free(I0s); free(I1s); free(u1s); free(u2s); free(nx); free(ny);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * (d) pyramidal Horn-Schunck: src/horn_schunck_pyramidal.cpp
 *
 * It shares the pyramid (a) and the warp (b) with TV-L1; only the inner SOR sweep is new.
 * The reference runs the interior sweep under `#pragma omp parallel for` although it updates
 * u, v in place (src/horn_schunck_pyramidal.cpp:148-158): with more than one thread the result
 * depends on scheduling.  This restatement is the ONE-thread semantics (lexicographic
 * Gauss-Seidel order), which is what tests pin against the reference run with one thread.
 * ---------------------------------------------------------------------------------------- */
#define ORC_HS_SOR_W 1.9                 /* src/horn_schunck_pyramidal.cpp:21 */
#define ORC_HS_PRESMOOTHING_SIGMA 0.8    /* src/horn_schunck_pyramidal.cpp:22 */

/* src/horn_schunck_pyramidal.cpp:31-71.  d[0..3] are the diagonal neighbours in the order the call
 * site passes them (p1..p4), a[0..3] the axial ones (p5..p8). */
static double orc_hs_sor_px(const PIX *Au, const PIX *Av, const PIX *Du, const PIX *Dv, const PIX *D,
                            PIX *u, PIX *v, double al, int p, const int d[4], const int a[4])
{
    const double w = ORC_HS_SOR_W;
    const double ula = 1. / 12. * (u[d[0]] + u[d[1]] + u[d[2]] + u[d[3]]) +
                       1. / 6. * (u[a[0]] + u[a[1]] + u[a[2]] + u[a[3]]);
    const double vla = 1. / 12. * (v[d[0]] + v[d[1]] + v[d[2]] + v[d[3]]) +
                       1. / 6. * (v[a[0]] + v[a[1]] + v[a[2]] + v[a[3]]);
    const PIX uk = u[p];
    const PIX vk = v[p];
    u[p] = (1.0 - w) * uk + w * (Au[p] - D[p] * v[p] + al * ula) / Du[p];
    v[p] = (1.0 - w) * vk + w * (Av[p] - D[p] * u[p] + al * vla) / Dv[p];
    return (u[p] - uk) * (u[p] - uk) + (v[p] - vk) * (v[p] - vk);
}

/* Neighbour indices of pixel (i, j): every call site of :147-228 passes the index-clamped
 * 8-neighbourhood (up-left, up-right, down-left, down-right; up, left, down, right) -- except the
 * bottom-right corner (:223-228), whose four diagonal arguments come in the order
 * (left, self, up-left, up); the floating-point sum follows the argument order, so it is kept. */
static double orc_hs_update(const PIX *Au, const PIX *Av, const PIX *Du, const PIX *Dv, const PIX *D,
                            PIX *u, PIX *v, double al, int i, int j, int nx, int ny)
{
    const int im = orc_clampi(i - 1, ny), ip = orc_clampi(i + 1, ny);
    const int jm = orc_clampi(j - 1, nx), jp = orc_clampi(j + 1, nx);
    int d[4] = { im * nx + jm, im * nx + jp, ip * nx + jm, ip * nx + jp };
    const int a[4] = { im * nx + j, i * nx + jm, ip * nx + j, i * nx + jp };
    if (i == ny - 1 && j == nx - 1) {
        const int k = i * nx + j;
        d[0] = k - 1; d[1] = k; d[2] = k - nx - 1; d[3] = k - nx;
    }
    return orc_hs_sor_px(Au, Av, Du, Dv, D, u, v, al, i * nx + j, d, a);
}

/* One SOR sweep, src/horn_schunck_pyramidal.cpp:144-230: interior in row-major order, then first /
 * last row interleaved per column, then first / last column interleaved per row, then the corners
 * (UL, UR, BL, BR).  Returns sqrt(sum / size). */
double orc_hs_sor_sweep(const PIX *Au, const PIX *Av, const PIX *Du, const PIX *Dv, const PIX *D,
                        PIX *u, PIX *v, double alpha2, int nx, int ny)
{
    double error = 0;
    for (int i = 1; i < ny - 1; i++)
        for (int j = 1; j < nx - 1; j++)
            error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, i, j, nx, ny);
    for (int j = 1; j < nx - 1; j++) {
        error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, 0, j, nx, ny);
        error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, ny - 1, j, nx, ny);
    }
    for (int i = 1; i < ny - 1; i++) {
        error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, i, 0, nx, ny);
        error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, i, nx - 1, nx, ny);
    }
    error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, 0, 0, nx, ny);
    error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, 0, nx - 1, nx, ny);
    error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, ny - 1, 0, nx, ny);
    error += orc_hs_update(Au, Av, Du, Dv, D, u, v, alpha2, ny - 1, nx - 1, nx, ny);
    return sqrt(error / (nx * ny));
}

/* The constant parts of the linear system, src/horn_schunck_pyramidal.cpp:127-137. */
void orc_hs_system(const PIX *I1, const PIX *I2w, const PIX *I2wx, const PIX *I2wy, const PIX *u,
                   const PIX *v, PIX *Au, PIX *Av, PIX *Du, PIX *Dv, PIX *D, double alpha2, int size)
{
    for (int i = 0; i < size; i++) {
        const double I2wl = I2wx[i] * u[i] + I2wy[i] * v[i];
        const double dif = I1[i] - I2w[i] + I2wl;
        Au[i] = dif * I2wx[i];
        Av[i] = dif * I2wy[i];
        Du[i] = I2wx[i] * I2wx[i] + alpha2;
        Dv[i] = I2wy[i] * I2wy[i] + alpha2;
        D[i] = I2wx[i] * I2wy[i];
    }
}

/* The SOR loop of one warp step (:139-231) on a given system; u, v in/out.  Returns the number of
 * sweeps; *err_out = the last sqrt(sum/size). */
int orc_hs_sor(const PIX *Au, const PIX *Av, const PIX *Du, const PIX *Dv, const PIX *D, PIX *u, PIX *v,
               double alpha2, int nx, int ny, double TOL, int maxiter, double *err_out)
{
    int niter = 0;
    double error = 1000;
    while (error > TOL && niter < maxiter) {
        niter++;
        error = orc_hs_sor_sweep(Au, Av, Du, Dv, D, u, v, alpha2, nx, ny);
    }
    if (err_out) *err_out = error;
    return niter;
}

/* horn_schunck_optical_flow, src/horn_schunck_pyramidal.cpp:78-249.  iters/errs: `warps` entries
 * (what the reference prints as "Iterations %d (%g)", :233-235), may be NULL. */
void orc_hs_single_scale(const PIX *I1, const PIX *I2, PIX *u, PIX *v, int nx, int ny, double alpha,
                         int warps, double TOL, int maxiter, int *iters, double *errs)
{
    const int size = nx * ny;
    const double alpha2 = alpha * alpha;
    PIX *buf = (PIX *) malloc(sizeof(PIX) * (size_t) size * 10);
    PIX *I2x = buf, *I2y = buf + size, *I2w = buf + 2 * (size_t) size, *I2wx = buf + 3 * (size_t) size,
        *I2wy = buf + 4 * (size_t) size, *Au = buf + 5 * (size_t) size, *Av = buf + 6 * (size_t) size,
        *Du = buf + 7 * (size_t) size, *Dv = buf + 8 * (size_t) size, *D = buf + 9 * (size_t) size;
    orc_centered_gradient(I2, I2x, I2y, nx, ny);                       /* :114 */
    for (int n = 0; n < warps; n++) {                                  /* :117 */
        orc_warp(I2, u, v, I2w, nx, ny, 1);                            /* :123-125 */
        orc_warp(I2x, u, v, I2wx, nx, ny, 1);
        orc_warp(I2y, u, v, I2wy, nx, ny, 1);
        orc_hs_system(I1, I2w, I2wx, I2wy, u, v, Au, Av, Du, Dv, D, alpha2, size);
        double error;
        const int niter = orc_hs_sor(Au, Av, Du, Dv, D, u, v, alpha2, nx, ny, TOL, maxiter, &error);
        if (iters) iters[n] = niter;
        if (errs) errs[n] = error;
    }
    free(buf);
}

/* horn_schunck_pyramidal, src/horn_schunck_pyramidal.cpp:258-370.  iters/errs are [nscales*warps],
 * coarsest scale first.  Returns 0, or 1 where the reference would throw from gaussian(). */
int orc_hs_multiscale(const PIX *I1, const PIX *I2, PIX *u, PIX *v, int nx0, int ny0, double alpha,
                      int nscales, double zfactor, int warps, double TOL, int maxiter, int *iters,
                      double *errs)
{
    const int size = nx0 * ny0;
    PIX **I1s = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **I2s = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **us = (PIX **) calloc(nscales, sizeof(PIX *));
    PIX **vs = (PIX **) calloc(nscales, sizeof(PIX *));
    int *nx = (int *) calloc(nscales, sizeof(int));
    int *ny = (int *) calloc(nscales, sizeof(int));
    int rc = 0, built = 1;

    I1s[0] = (PIX *) malloc(sizeof(PIX) * size);
    I2s[0] = (PIX *) malloc(sizeof(PIX) * size);
    orc_normalize(I1, I2, I1s[0], I2s[0], size);                                  /* :293 */
    rc |= orc_gaussian(I1s[0], nx0, ny0, ORC_HS_PRESMOOTHING_SIGMA);              /* :296 */
    rc |= orc_gaussian(I2s[0], nx0, ny0, ORC_HS_PRESMOOTHING_SIGMA);              /* :297 */
    us[0] = u; vs[0] = v; nx[0] = nx0; ny[0] = ny0;

    for (int s = 1; s < nscales && !rc; s++) {                                    /* :305-317 */
        orc_zoom_size(nx[s - 1], ny[s - 1], &nx[s], &ny[s], zfactor);
        const size_t sizes = (size_t) nx[s] * ny[s];
        I1s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        I2s[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        us[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        vs[s] = (PIX *) malloc(sizeof(PIX) * sizes);
        built = s + 1;
        rc |= orc_zoom_out(I1s[s - 1], I1s[s], nx[s - 1], ny[s - 1], zfactor);
        rc |= orc_zoom_out(I2s[s - 1], I2s[s], nx[s - 1], ny[s - 1], zfactor);
    }

    if (!rc) {
        for (int i = 0; i < nx[nscales - 1] * ny[nscales - 1]; i++)              /* :320-323 */
            us[nscales - 1][i] = vs[nscales - 1][i] = 0;
        for (int s = nscales - 1; s >= 0; s--) {                                  /* :326-353 */
            const int k = nscales - 1 - s;
            orc_hs_single_scale(I1s[s], I2s[s], us[s], vs[s], nx[s], ny[s], alpha, warps, TOL, maxiter,
                                iters ? iters + k * warps : 0, errs ? errs + k * warps : 0);
            if (!s) break;
            orc_zoom_in(us[s], us[s - 1], nx[s], ny[s], nx[s - 1], ny[s - 1]);
            orc_zoom_in(vs[s], vs[s - 1], nx[s], ny[s], nx[s - 1], ny[s - 1]);
            for (int i = 0; i < nx[s - 1] * ny[s - 1]; i++) {
                us[s - 1][i] *= 1.0 / zfactor;
                vs[s - 1][i] *= 1.0 / zfactor;
            }
        }
    }

    for (int i = 1; i < built; i++) { free(I1s[i]); free(I2s[i]); free(us[i]); free(vs[i]); }
    free(I1s[0]); free(I2s[0]);
    free(I1s); free(I2s); free(us); free(vs); free(nx); free(ny);
    return rc;
}

/* ==========================================================================================
 * (e) TV-L1 with occlusion detection: src/tvl1occflow.cpp,
 *     src/tvl1occflow_solvers.cpp, src/tvl1occflow_tv_rof_box.cpp, me_median_filtering of src/utils.cpp
 *
 * DEFINED BEHAVIOUR.  The reference keeps the dual variables of Solver_wrt_u (p11..p22) and of
 * Solver_wrt_chi (eta1, eta2) in function-level statics that are re-created whenever the image width
 * changes (src/tvl1occflow_solvers.cpp:163-186, :241-253); eta1/eta2 are then READ UNINITIALISED
 * (:262, the file's own #warning).  What a fresh process computes de facto -- large new[] blocks come
 * from zero pages -- is: p and eta start from zero at every pyramid level.  That is the behaviour
 * restated here, and the compiled reference is pinned to it by building it with a zero-filling
 * operator new[] (oracle/occ_ref_shim.cpp; the reference sources stay untouched).
 * Consecutive levels always differ in width, so "re-created when nx changes" == "per level".
 * ========================================================================================== */
#define ORC_OCC_EXT_MAX_ITERATIONS 20    /* src/tvl1occflow_constants.h:25 */
#define ORC_OCC_OMEGA 1.25               /* :26 */
#define ORC_OCC_IS_ZERO 1E-10            /* :28 */
#define ORC_OCC_THR_CHI 0.75             /* :29 */
#define ORC_OCC_MAX_ITERATIONS_CHI 100   /* :30 */
#define ORC_OCC_PRESMOOTHING_SIGMA 0.8   /* :31 */
#define ORC_OCC_G_FACTOR 0.05            /* :35 (G_CHOICE 2) */
#define ORC_OCC_MAX_ITERATIONS_U 10      /* :37 */
#define ORC_OCC_TAU_ETA 0.15             /* :38 */
#define ORC_OCC_TAU_CHI 0.15             /* :39 */

/* me_median_filtering, src/utils.cpp:151-213, window 3: median of the 3x3 neighbourhood with the
 * symmetric boundary x<0 -> -x-1, x>=n -> 2n-x-1. */
void orc_median3(PIX *in, int nx, int ny)
{
    PIX *out = (PIX *) malloc(sizeof(PIX) * (size_t) nx * ny);
    #pragma omp parallel for
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            PIX v[9];
            int n = 0;
            for (int yy = y - 1; yy <= y + 1; yy++)
                for (int xx = x - 1; xx <= x + 1; xx++) {
                    int x0 = xx, y0 = yy;
                    if (x0 < 0) x0 = -x0 - 1;
                    if (x0 >= nx) x0 = 2 * nx - x0 - 1;
                    if (y0 < 0) y0 = -y0 - 1;
                    if (y0 >= ny) y0 = 2 * ny - y0 - 1;
                    v[n++] = in[y0 * nx + x0];
                }
            for (int a = 1; a < 9; a++) {            /* insertion sort: the median is order-independent */
                const PIX t = v[a];
                int b = a - 1;
                while (b >= 0 && v[b] > t) { v[b + 1] = v[b]; b--; }
                v[b + 1] = t;
            }
            out[y * nx + x] = v[4];
        }
    memcpy(in, out, sizeof(PIX) * (size_t) nx * ny);
    free(out);
}

/* choosed_g with G_CHOICE 2, src/tvl1occflow.cpp:100-136: g = 1 / (1 + G_FACTOR |grad I|). */
void orc_occ_g(const PIX *I, PIX *g, int nx, int ny)
{
    const int size = nx * ny;
    PIX *Ix = (PIX *) malloc(sizeof(PIX) * size), *Iy = (PIX *) malloc(sizeof(PIX) * size);
    orc_centered_gradient(I, Ix, Iy, nx, ny);
    for (int i = 0; i < size; i++) {
        /* C++ overload resolution in the reference: sqrt of an ofpix_t expression is computed in ofpix_t */
        const double gggrad = ORC_SQRT_PIX(Ix[i] * Ix[i] + Iy[i] * Iy[i]);
        const double aux = 1. + ORC_OCC_G_FACTOR * gggrad;
        g[i] = 1. / aux;
    }
    free(Ix); free(Iy);
}

/* Solver_wrt_v, src/tvl1occflow_solvers.cpp:54-148. */
void orc_occ_solver_v(const PIX *u1, const PIX *u2, PIX *v1, PIX *v2, const PIX *chi, const PIX *I1wx,
                      const PIX *I1wy, const PIX *I_1wx, const PIX *I_1wy, const PIX *rho1_c,
                      const PIX *rho3_c, PIX *Vfwd_1, PIX *Vfwd_2, PIX *Vbck_1, PIX *Vbck_2, const PIX *grad1,
                      const PIX *grad3, double alpha, double theta, double lambda, int size)
{
    const double l_t = lambda * theta;
    const double _1pat = 1. + alpha * theta;
    const double at_d_1pat = alpha * theta / _1pat;
    const double lt_d_1pat = 2. * lambda * theta / _1pat;
    for (int i = 0; i < size; i++) {
        double d1 = 0, d2 = 0;
        const double rho1 = rho1_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
        if (rho1 < -l_t * grad1[i]) { d1 = l_t * I1wx[i]; d2 = l_t * I1wy[i]; }
        else if (rho1 > l_t * grad1[i]) { d1 = -l_t * I1wx[i]; d2 = -l_t * I1wy[i]; }
        else if (grad1[i] < ORC_OCC_IS_ZERO) { d1 = d2 = 0; }
        else { d1 = -rho1 * I1wx[i] / grad1[i]; d2 = -rho1 * I1wy[i] / grad1[i]; }
        Vfwd_1[i] = u1[i] + d1;
        Vfwd_2[i] = u2[i] + d2;

        const double rho3 = rho3_c[i] - (I_1wx[i] * u1[i] + I_1wy[i] * u2[i]);
        const double A = rho3 + at_d_1pat * (I_1wx[i] * u1[i] + I_1wy[i] * u2[i]);
        if (A < -lt_d_1pat * grad3[i]) {
            d1 = -lt_d_1pat * I_1wx[i]; d2 = -lt_d_1pat * I_1wy[i];
            Vbck_1[i] = (u1[i] / _1pat) + d1; Vbck_2[i] = (u2[i] / _1pat) + d2;
        } else if (A > lt_d_1pat * grad3[i]) {
            d1 = lt_d_1pat * I_1wx[i]; d2 = lt_d_1pat * I_1wy[i];
            Vbck_1[i] = (u1[i] / _1pat) + d1; Vbck_2[i] = (u2[i] / _1pat) + d2;
        } else {
            if (grad3[i] < ORC_OCC_IS_ZERO) { d1 = d2 = 0; }
            else { d1 = rho3 * I_1wx[i] / grad3[i]; d2 = rho3 * I_1wy[i] / grad3[i]; }
            Vbck_1[i] = u1[i] + d1; Vbck_2[i] = u2[i] + d2;
        }
        if (chi[i] < ORC_OCC_THR_CHI) { v1[i] = Vfwd_1[i]; v2[i] = Vfwd_2[i]; }
        else { v1[i] = Vbck_1[i]; v2[i] = Vbck_2[i]; }
    }
}

/* Scalar_ROF_BoxCellCentered, src/tvl1occflow_tv_rof_box.cpp:25-645, restated on compact arrays.
 * The reference works on a (2nx+1) x (2ny+1) staggered grid whose only live unknowns are the dual values
 * on the SOUTH and EAST side of every cell: pS[i][j] (its initialP1) and pE[i][j] (its initialP2); a
 * cell's NORTH side is the south side of the cell above, its WEST side the east side of the cell to the
 * left, and sides on the image border are never updated (north / west: 0; south of the last row / east
 * of the last column: whatever pS / pE hold there).  One sweep = alfa = |grad u| / (lambda g) per cell
 * (:186-202), then a Gauss-Seidel pass over the cells in row-major order, each cell re-solving the
 * 2x2 / 3x3 / 4x4 system of its own sides with relaxation omega (corner :206-247 / :304-331 / :452-481
 * / :527-554, edge :250-301 / :334-383 / :486-524 / :484-..., interior :385-482), then
 * u = lambda f + lambda (pS - pN + pE - pW) (:557-585).  Every expression keeps the reference's
 * association order (double build bit-exact). */
#define OCC_PS(i, j) ((i) < 0 ? (PIX) 0 : pS[(i) * nx + (j)])
#define OCC_PE(i, j) ((j) < 0 ? (PIX) 0 : pE[(i) * nx + (j)])
void orc_occ_rof_box(PIX *u, const PIX *f, PIX *pS, PIX *pE, const PIX *g, double lambda, double omega,
                     int nx, int ny, int nIter)
{
    const int size = nx * ny;
    PIX *ux = (PIX *) malloc(sizeof(PIX) * size), *uy = (PIX *) malloc(sizeof(PIX) * size);
    PIX *al = (PIX *) malloc(sizeof(PIX) * size);
    for (int iter = 1; iter <= nIter; iter++) {
        orc_forward_gradient(u, ux, uy, nx, ny);
        /* the reference's OWN hypot, src/tvl1occflow_tv_rof_box.cpp:15-20: sqrt(x*x + y*y) in double, not libm's */
        for (int k = 0; k < size; k++) {
            const double x = ux[k], y = uy[k];
            al[k] = sqrt(x * x + y * y) / (lambda * g[k]);
        }
        for (int i = 0; i < ny; i++)
            for (int j = 0; j < nx; j++) {
                const int hasW = j > 0, hasN = i > 0, hasS = i < ny - 1, hasE = j < nx - 1;
                const int c = i * nx + j;
                /* beta of each side present (0 otherwise): -2 - alfa of the cell that owns the side.  Operand
                 * types as in the reference: beta[], stgGrid_P and stgGrid_F are ofpix_t, the free terms double */
                const PIX b0 = hasW ? -2 - al[c - 1] : 0, b1 = hasN ? -2 - al[c - nx] : 0;
                const PIX b2 = hasS ? -2 - al[c] : 0, b3 = hasE ? -2 - al[c] : 0;
                /* side values of the neighbours (staggered names of the reference in comments) */
                double W = 0, N = 0, S = 0, E = 0;
                const int n_edge = !hasN && hasW && hasE;       /* north side, not a corner: F comes first */
                if (hasW) {
                    const PIX jm3 = OCC_PE(i, j - 2), ip1_jm2 = pS[c - 1], im1_jm2 = OCC_PS(i - 1, j - 1);
                    const PIX Fw = f[c] - f[c - 1];
                    W = n_edge ? -Fw - jm3 + ip1_jm2 - im1_jm2 : -jm3 + ip1_jm2 - im1_jm2 - Fw;
                }
                if (hasN) {
                    const PIX im3 = OCC_PS(i - 2, j), im2_jp1 = pE[c - nx], im2_jm1 = OCC_PE(i - 1, j - 1);
                    const PIX Fn = f[c] - f[c - nx];
                    N = -im3 + im2_jp1 - im2_jm1 - Fn;
                }
                if (hasS) {
                    const PIX ip3 = pS[c + nx], ip2_jp1 = pE[c + nx], ip2_jm1 = OCC_PE(i + 1, j - 1);
                    const PIX Fs = f[c + nx] - f[c];
                    S = n_edge ? -Fs - ip3 - ip2_jp1 + ip2_jm1 : -ip3 - ip2_jp1 + ip2_jm1 - Fs;
                }
                if (hasE) {
                    const PIX jp3 = pE[c + 1], ip1_jp2 = pS[c + 1], im1_jp2 = OCC_PS(i - 1, j + 1);
                    const PIX Fe = f[c + 1] - f[c];
                    E = n_edge ? -Fe - jp3 - ip1_jp2 + im1_jp2 : -jp3 - ip1_jp2 + im1_jp2 - Fe;
                }
                PIX *qW = hasW ? &pE[c - 1] : 0, *qN = hasN ? &pS[c - nx] : 0, *qS = &pS[c], *qE = &pE[c];
                double den;
#define OCC_RELAX(q, num) (*(q) = (1 - omega) * *(q) + omega * (num) / den)
                if (hasW && hasN && hasS && hasE) {                              /* :385-482, Gauss elimination */
                    const double a = 1 / b0;
                    const double b = -(b0 + 1) / (b0 * b1 - 1);
                    const double alf = 1 + a;
                    const double gam = -a + b * alf;
                    const double x = N + a * W;
                    const double y = -a * W + b * x;
                    const double cc = (1 - gam) / (b2 + gam);
                    *qE = (1 - omega) * *qE + omega * (E + y + cc * (S + y)) / (b3 + gam + cc * (gam - 1));
                    *qS = (1 - omega) * *qS + omega * (S + y + *qE * (1 - gam)) / (b2 + gam);
                    *qN = (1 - omega) * *qN + omega * (x - alf * (*qE + *qS)) / (b1 - a);
                    *qW = (1 - omega) * *qW + omega * (W + *qN - *qS - *qE) / (b0);
                } else if (!hasN && !hasW) {                                      /* NW corner :206-247 */
                    den = b2 * b3 - 1;
                    OCC_RELAX(qS, S * b3 + E);
                    OCC_RELAX(qE, E * b2 + S);
                } else if (!hasN && !hasE) {                                      /* NE corner :304-331 */
                    den = b0 * b2 - 1;
                    OCC_RELAX(qW, W * b2 - S);
                    OCC_RELAX(qS, S * b0 - W);
                } else if (!hasN) {                                               /* north side :250-301 */
                    den = b0 * b2 * b3 - b0 - b2 - b3 - 2;
                    OCC_RELAX(qW, W * b2 * b3 - E * b2 - S * b3 - W - E - S);
                    OCC_RELAX(qS, S * b0 * b3 - W * b3 + E * b0 - W + E - S);
                    OCC_RELAX(qE, E * b0 * b2 - W * b2 + S * b0 - W - E + S);
                } else if (!hasS && !hasW) {                                      /* SW corner */
                    den = b3 * b1 - 1;
                    OCC_RELAX(qN, b3 * N - E);
                    OCC_RELAX(qE, b1 * E - N);
                } else if (!hasS && !hasE) {                                      /* SE corner */
                    den = b0 * b1 - 1;
                    OCC_RELAX(qW, W * b1 + N);
                    OCC_RELAX(qN, N * b0 + W);
                } else if (!hasS) {                                               /* south side */
                    den = b0 * b1 * b3 - b0 - b1 - b3 - 2;
                    OCC_RELAX(qW, W * b1 * b3 - E + N - E * b1 - W + N * b3);
                    OCC_RELAX(qN, N * b0 * b3 + W - E - N - E * b0 + W * b3);
                    OCC_RELAX(qE, E * b0 * b1 - N - W - W * b1 - N * b0 - E);
                } else if (!hasW) {                                               /* west side :334-383 */
                    den = b1 * b2 * b3 - (b1 + b2 + b3) - 2;
                    OCC_RELAX(qN, b2 * b3 * N - E * b2 - S * b3 - N - S - E);
                    OCC_RELAX(qS, b1 * b3 * S + E * b1 - N * b3 - N - S + E);
                    OCC_RELAX(qE, b1 * b2 * E - N * b2 + S * b1 - N + S - E);
                } else {                                                          /* east side */
                    den = (b0 * b1 * b2) + (-b0 - b1 - b2 - 2);
                    OCC_RELAX(qW, W * b1 * b2 - S + N - S * b1 - W + N * b2);
                    OCC_RELAX(qN, N * b0 * b2 + W - S - N - S * b0 + W * b2);
                    OCC_RELAX(qS, S * b0 * b1 - N - W - W * b1 - N * b0 - S);
                }
#undef OCC_RELAX
            }
        for (int i = 0; i < ny; i++)
            for (int j = 0; j < nx; j++) {
                const int c = i * nx + j;
                u[c] = lambda * f[c] + lambda * (pS[c] - OCC_PS(i - 1, j) + pE[c] - OCC_PE(i, j - 1));
            }
    }
    free(ux); free(uy); free(al);
}
#undef OCC_PS
#undef OCC_PE

/* Solver_wrt_u, src/tvl1occflow_solvers.cpp:150-213 (p = the level's persistent dual variables). */
void orc_occ_solver_u(PIX *u1, PIX *u2, const PIX *v1, const PIX *v2, const PIX *chi, const PIX *g,
                      PIX *p11, PIX *p12, PIX *p21, PIX *p22, double theta, double beta, int nx, int ny)
{
    const int size = nx * ny;
    PIX *chix = (PIX *) malloc(sizeof(PIX) * size), *chiy = (PIX *) malloc(sizeof(PIX) * size);
    PIX *f1 = (PIX *) malloc(sizeof(PIX) * size), *f2 = (PIX *) malloc(sizeof(PIX) * size);
    orc_forward_gradient(chi, chix, chiy, nx, ny);
    for (int i = 0; i < size; i++) {
        f1[i] = v1[i] / theta + beta * chix[i];
        f2[i] = v2[i] / theta + beta * chiy[i];
        u1[i] = v1[i] + theta * beta * chix[i];
        u2[i] = v2[i] + theta * beta * chiy[i];
    }
    orc_occ_rof_box(u1, f1, p11, p12, g, theta, ORC_OCC_OMEGA, nx, ny, ORC_OCC_MAX_ITERATIONS_U);
    orc_occ_rof_box(u2, f2, p21, p22, g, theta, ORC_OCC_OMEGA, nx, ny, ORC_OCC_MAX_ITERATIONS_U);
    free(chix); free(chiy); free(f1); free(f2);
}

/* Solver_wrt_chi, src/tvl1occflow_solvers.cpp:215-338 (eta = the level's persistent dual variable). */
void orc_occ_solver_chi(const PIX *u1, const PIX *u2, PIX *chi, const PIX *I1wx, const PIX *I1wy,
                        const PIX *I_1wx, const PIX *I_1wy, const PIX *rho1_c, const PIX *rho3_c,
                        const PIX *Vfwd_1, const PIX *Vfwd_2, const PIX *Vbck_1, const PIX *Vbck_2, const PIX *g,
                        PIX *eta1, PIX *eta2, double lambda, double theta, double alpha, double beta,
                        double tau_chi, double tau_eta, int nx, int ny, int niter)
{
    const int size = nx * ny;
    PIX *chix = (PIX *) malloc(sizeof(PIX) * size), *chiy = (PIX *) malloc(sizeof(PIX) * size);
    PIX *geta1 = (PIX *) malloc(sizeof(PIX) * size), *geta2 = (PIX *) malloc(sizeof(PIX) * size);
    PIX *div_eta = (PIX *) malloc(sizeof(PIX) * size), *div_u = (PIX *) malloc(sizeof(PIX) * size);
    for (int n_chi = 0; n_chi < niter; n_chi++) {
        orc_forward_gradient(chi, chix, chiy, nx, ny);
        for (int i = 0; i < size; i++) {
            eta1[i] = eta1[i] + tau_eta * g[i] * chix[i];
            eta2[i] = eta2[i] + tau_eta * g[i] * chiy[i];
        }
        for (int j = 0; j < size; j++) {                                  /* project, :33-52 */
            const double norm2 = eta1[j] * eta1[j] + eta2[j] * eta2[j];
            if (norm2 < ORC_OCC_IS_ZERO) { eta1[j] = 0.0; eta2[j] = 0.0; }
            else { const double norm = sqrt(norm2); eta1[j] = eta1[j] / norm; eta2[j] = eta2[j] / norm; }
        }
        for (int i = 0; i < size; i++) { geta1[i] = g[i] * eta1[i]; geta2[i] = g[i] * eta2[i]; }
        orc_divergence(geta1, geta2, div_eta, nx, ny);
        orc_divergence(u1, u2, div_u, nx, ny);
        for (int i = 0; i < size; i++) {
            const double rho1 = rho1_c[i] + (I1wx[i] * Vfwd_1[i] + I1wy[i] * Vfwd_2[i]);
            const double abs_rho1 = (rho1 < 0.) ? -rho1 : rho1;
            const double rho3 = rho3_c[i] - (I_1wx[i] * Vbck_1[i] + I_1wy[i] * Vbck_2[i]);
            const double abs_rho3 = (rho3 < 0.) ? -rho3 : rho3;
            double F, G;
            if (chi[i] < 0.5) {
                F = -lambda * abs_rho1;
                G = -(0.5 / theta) * ((Vfwd_1[i] - u1[i]) * (Vfwd_1[i] - u1[i]) + (Vfwd_2[i] - u2[i]) * (Vfwd_2[i] - u2[i]));
            } else {
                F = lambda * abs_rho3;
                G = (0.5 / theta) * ((Vbck_1[i] - u1[i]) * (Vbck_1[i] - u1[i]) + (Vbck_2[i] - u2[i]) * (Vbck_2[i] - u2[i]))
                    + alpha * theta * (Vbck_1[i] * Vbck_1[i] + Vbck_2[i] * Vbck_2[i]);
            }
            chi[i] = chi[i] + tau_chi * (div_eta[i] - F - G - beta * div_u[i]);
            if (chi[i] > 1.) chi[i] = 1.;
            else if (chi[i] < 0.) chi[i] = 0.;
        }
    }
    free(chix); free(chiy); free(geta1); free(geta2); free(div_eta); free(div_u);
}

/* Dual_TVL1_optic_flow of src/tvl1occflow.cpp:144-330: one level.  iters/errs: [warps]. */
void orc_occ_single_scale(const PIX *I_1, const PIX *I0, const PIX *I1, const PIX *filtI0, PIX *u1, PIX *u2,
                          PIX *chi, int nx, int ny, double lambda, double alpha, double beta, double theta,
                          int warps, double epsilon, int *iters, double *errs)
{
    const int size = nx * ny;
    PIX *buf = (PIX *) calloc((size_t) 30 * size, sizeof(PIX));
    PIX *I1x = buf, *I1y = buf + size, *I1w = buf + 2 * size, *I1wx = buf + 3 * size, *I1wy = buf + 4 * size;
    PIX *I_1x = buf + 5 * size, *I_1y = buf + 6 * size, *I_1w = buf + 7 * size, *I_1wx = buf + 8 * size;
    PIX *I_1wy = buf + 9 * size, *rho1_c = buf + 10 * size, *rho3_c = buf + 11 * size, *v1 = buf + 12 * size;
    PIX *v2 = buf + 13 * size, *v11 = buf + 14 * size, *v12 = buf + 15 * size, *v31 = buf + 16 * size;
    PIX *v32 = buf + 17 * size, *gp1 = buf + 18 * size, *gp2 = buf + 19 * size, *grad1 = buf + 20 * size;
    PIX *grad3 = buf + 21 * size, *g = buf + 22 * size, *u1prev = buf + 23 * size, *u2prev = buf + 24 * size;
    PIX *p11 = buf + 25 * size, *p12 = buf + 26 * size, *p21 = buf + 27 * size, *p22 = buf + 28 * size;
    PIX *eta = (PIX *) calloc((size_t) 2 * size, sizeof(PIX));          /* zero at every level, see the header */

    orc_occ_g(filtI0, g, nx, ny);                                        /* :205 */
    orc_centered_gradient(I1, I1x, I1y, nx, ny);                         /* :207-208 */
    orc_centered_gradient(I_1, I_1x, I_1y, nx, ny);
    for (int i = 0; i < size; i++) { u1prev[i] = u1[i]; u2prev[i] = u2[i]; }   /* :211-224 (the rest is zero) */

    for (int w = 0; w < warps; w++) {
        orc_warp(I1, u1, u2, I1w, nx, ny, 0);                            /* :232-234; border_out defaults to false here */
        orc_warp(I1x, u1, u2, I1wx, nx, ny, 0);
        orc_warp(I1y, u1, u2, I1wy, nx, ny, 0);
        for (int i = 0; i < size; i++) { gp1[i] = -u1[i]; gp2[i] = -u2[i]; }    /* :237-241 */
        orc_warp(I_1, gp1, gp2, I_1w, nx, ny, 0);                        /* :243-245 */
        orc_warp(I_1x, gp1, gp2, I_1wx, nx, ny, 0);
        orc_warp(I_1y, gp1, gp2, I_1wy, nx, ny, 0);
        for (int i = 0; i < size; i++) {                                 /* :250-270 */
            double Ix2 = I1wx[i] * I1wx[i], Iy2 = I1wy[i] * I1wy[i];
            grad1[i] = (Ix2 + Iy2);
            Ix2 = I_1wx[i] * I_1wx[i]; Iy2 = I_1wy[i] * I_1wy[i];
            grad3[i] = (Ix2 + Iy2);
            rho1_c[i] = (I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i]);
            rho3_c[i] = (I_1w[i] + I_1wx[i] * u1[i] + I_1wy[i] * u2[i] - I0[i]);
        }
        int n = 0;
        double error = INFINITY;
        while (error > epsilon && n < ORC_OCC_EXT_MAX_ITERATIONS) {      /* :277: epsilon, not epsilon^2 */
            n++;
            orc_occ_solver_v(u1, u2, v1, v2, chi, I1wx, I1wy, I_1wx, I_1wy, rho1_c, rho3_c, v11, v12, v31, v32,
                             grad1, grad3, alpha, theta, lambda, size);
            orc_occ_solver_u(u1, u2, v1, v2, chi, g, p11, p12, p21, p22, theta, beta, nx, ny);
            orc_median3(u1, nx, ny);                                     /* :292-293 */
            orc_median3(u2, nx, ny);
            orc_occ_solver_chi(u1, u2, chi, I1wx, I1wy, I_1wx, I_1wy, rho1_c, rho3_c, v11, v12, v31, v32, g,
                               eta, eta + size, lambda, theta, alpha, beta, ORC_OCC_TAU_CHI, ORC_OCC_TAU_ETA,
                               nx, ny, ORC_OCC_MAX_ITERATIONS_CHI);
            error = 0.0;                                                 /* L2error, :70-88 */
            for (int i = 0; i < size; i++) {
                error += (u1[i] - u1prev[i]) * (u1[i] - u1prev[i]) + (u2[i] - u2prev[i]) * (u2[i] - u2prev[i]);
                u1prev[i] = u1[i];
                u2prev[i] = u2[i];
            }
            error /= size;
        }
        if (iters) iters[w] = n;
        if (errs) errs[w] = error;
    }
    free(buf); free(eta);
}

/* Dual_TVL1_optic_flow_multiscale of src/tvl1occflow.cpp:335-482.  The result of image_normalization_4
 * (:379-380) is overwritten by the raw images two lines later (:383-388): the pyramid is built from the
 * UN-normalised inputs, so it is here.  chi is thresholded at THR_CHI at the end (:459).
 * iters/errs: [nscales][warps], coarsest level first. */
int orc_occ_multiscale(const PIX *I_1, const PIX *I0, const PIX *I1, const PIX *filtI0, PIX *u1, PIX *u2,
                       PIX *chi, int nxx, int nyy, double lambda, double alpha, double beta, double theta,
                       int nscales, double zfactor, int warps, double epsilon, int *iters, double *errs)
{
    const int size = nxx * nyy;
    PIX **im[4], **us[3];
    int *nx = (int *) calloc(nscales, sizeof(int)), *ny = (int *) calloc(nscales, sizeof(int));
    int rc = 0;
    const PIX *src[4] = { I_1, I0, I1, filtI0 };
    for (int k = 0; k < 4; k++) im[k] = (PIX **) calloc(nscales, sizeof(PIX *));
    for (int k = 0; k < 3; k++) us[k] = (PIX **) calloc(nscales, sizeof(PIX *));
    us[0][0] = u1; us[1][0] = u2; us[2][0] = chi;
    nx[0] = nxx; ny[0] = nyy;
    for (int k = 0; k < 4; k++) {
        im[k][0] = (PIX *) malloc(sizeof(PIX) * size);
        memcpy(im[k][0], src[k], sizeof(PIX) * size);
    }
    for (int i = 0; i < size; i++) u1[i] = u2[i] = chi[i] = 0.0;
    for (int k = 0; k < 4 && !rc; k++) rc |= orc_gaussian(im[k][0], nxx, nyy, ORC_OCC_PRESMOOTHING_SIGMA);
    int built = 1;
    for (int s = 1; s < nscales && !rc; s++) {
        orc_zoom_size(nx[s - 1], ny[s - 1], &nx[s], &ny[s], zfactor);
        const size_t sizes = (size_t) nx[s] * ny[s];
        for (int k = 0; k < 4; k++) im[k][s] = (PIX *) malloc(sizeof(PIX) * sizes);
        for (int k = 0; k < 3; k++) us[k][s] = (PIX *) calloc(sizes, sizeof(PIX));
        built = s + 1;
        for (int k = 0; k < 4 && !rc; k++) rc |= orc_zoom_out(im[k][s - 1], im[k][s], nx[s - 1], ny[s - 1], zfactor);
    }
    for (int s = nscales - 1; s >= 0 && !rc; s--) {
        const int k = nscales - 1 - s;
        orc_occ_single_scale(im[0][s], im[1][s], im[2][s], im[3][s], us[0][s], us[1][s], us[2][s], nx[s], ny[s],
                             lambda, alpha, beta, theta, warps, epsilon, iters ? iters + k * warps : 0,
                             errs ? errs + k * warps : 0);
        if (s) {
            for (int c = 0; c < 3; c++) orc_zoom_in(us[c][s], us[c][s - 1], nx[s], ny[s], nx[s - 1], ny[s - 1]);
            for (int i = 0; i < nx[s - 1] * ny[s - 1]; i++) {
                us[0][s - 1][i] *= (double) 1.0 / zfactor;
                us[1][s - 1][i] *= (double) 1.0 / zfactor;
            }
        } else {
            for (int i = 0; i < size; i++) chi[i] = (chi[i] > ORC_OCC_THR_CHI);
        }
    }
    for (int s = 0; s < built; s++) {
        for (int k = 0; k < 4; k++) free(im[k][s]);
        if (s) for (int k = 0; k < 3; k++) free(us[k][s]);
    }
    for (int k = 0; k < 4; k++) free(im[k]);
    for (int k = 0; k < 3; k++) free(us[k]);
    free(nx); free(ny);
    return rc;
}
