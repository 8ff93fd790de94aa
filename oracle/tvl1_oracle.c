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
#else
typedef double PIX;
#define ORC_HYPOT(a, b) hypot((a), (b))
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
