// CUDA kernels of the TV-L1 + occlusions solver (include/occ_b200.h).
//
// IEEE fp64 throughout, reference association order, no fused multiply-adds: this header is only
// included by occ_solver.cu, which is compiled with -fmad=false (the reference is built for baseline
// x86-64, which has no FMA; fp64 division and square root are correctly rounded on both sides).  Every
// kernel below restates one loop nest of the reference (cited at the kernel) as one thread per pixel;
// the only loop whose ORDER matters, the Gauss-Seidel pass of the box relaxation, keeps its data
// dependences on a wavefront (k_occ_rof_gs).
//
// Batch layout: a "plane" is a dense [ny][nx] array of double; a batch of B triples keeps plane b of a
// field at field + b * N (N = nx * ny); fields that come in groups ([k][B][N]) say so.  ctl[b].active
// gates every kernel of the outer loop: a triple that met its stopping rule costs nothing afterwards.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace occ {

struct TripleCtl {
    int active;        // still inside the while of src/tvl1occflow.cpp:277
    int n;             // outer iterations of the current warp step
    double error;      // last L2error
};

struct GaussTaps {
    int size;          // taps = radius + 1
    double B[40];
};

constexpr int kErrThreads = 256;

// ---------------------------------------------------------------------------------------------------
// pyramid: gaussian (src/operators.cpp:506-624), zoom_out / zoom_in (src/zoom.cpp:41-78, :132-155)
// ---------------------------------------------------------------------------------------------------

// reflecting pad of src/operators.cpp:557-562: x = -k -> I[k]; x = n-1+k -> I[n-k]
__device__ __forceinline__ int gauss_reflect(int x, int n)
{
    if (x < 0) x = -x;
    else if (x >= n) x = 2 * n - 1 - x;
    return x < 0 ? 0 : (x >= n ? n - 1 : x);     // only for lines shorter than the window (undefined upstream)
}

// One pass along x (DIR 0) or y (DIR 1); sum order of :573-576.
template <int DIR>
__global__ void __launch_bounds__(256) k_occ_gauss_pass(const double *__restrict__ in, double *__restrict__ out,
                                                    int nx, int ny, GaussTaps t)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny;
    const double *src = in + blockIdx.z * N;
    const int n = DIR ? ny : nx, c = DIR ? i : j;
    auto at = [&](int x) { x = gauss_reflect(x, n); return DIR ? src[(size_t) x * nx + j] : src[(size_t) i * nx + x]; };
    double sum = t.B[0] * at(c);
    for (int k = 1; k < t.size; k++) sum += t.B[k] * (at(c - k) + at(c + k));
    out[blockIdx.z * N + (size_t) i * nx + j] = sum;
}

// src/bicubic_interpolation.cpp:108-123
__device__ __forceinline__ double cubic(double v0, double v1, double v2, double v3, double x)
{
    return v1 + 0.5 * x * (v2 - v0 + x * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + x * (3.0 * (v1 - v2) + v3 - v0)));
}

// Index part of src/bicubic_interpolation.cpp:153-245 (BOUNDARY_CONDITION 0 = neumann, :24-39), shared by
// every plane sampled at the same position.  The "minus" row uses sx (:173, upstream quirk, kept).
struct BicubicPos {
    int xs[4], ys[4];
    double fx, fy;
    bool out;
};

__device__ __forceinline__ int neumann(int x, int n, bool &out)
{
    if (x < 0) { x = 0; out = true; }
    else if (x >= n) { x = n - 1; out = true; }
    return x;
}

__device__ __forceinline__ BicubicPos bicubic_pos(double uu, double vv, int nx, int ny)
{
    BicubicPos p;
    const int sx = (uu < 0) ? -1 : 1, sy = (vv < 0) ? -1 : 1;
    p.out = false;
    const int x = neumann((int) uu, nx, p.out), y = neumann((int) vv, ny, p.out);
    p.xs[0] = neumann((int) uu - sx, nx, p.out);
    p.ys[0] = neumann((int) vv - sx, ny, p.out);
    p.xs[1] = x;
    p.ys[1] = y;
    p.xs[2] = neumann((int) uu + sx, nx, p.out);
    p.ys[2] = neumann((int) vv + sy, ny, p.out);
    p.xs[3] = neumann((int) uu + 2 * sx, nx, p.out);
    p.ys[3] = neumann((int) vv + 2 * sy, ny, p.out);
    p.fx = uu - x;
    p.fy = vv - y;
    return p;
}

// along y inside each of the four x-columns first, then along x (:137-144, :236-240); border_out = false
__device__ __forceinline__ double bicubic_eval(const double *__restrict__ in, const BicubicPos &p, int nx)
{
    double col[4];
#pragma unroll
    for (int a = 0; a < 4; a++)
        col[a] = cubic(in[p.xs[a] + (size_t) nx * p.ys[0]], in[p.xs[a] + (size_t) nx * p.ys[1]],
                       in[p.xs[a] + (size_t) nx * p.ys[2]], in[p.xs[a] + (size_t) nx * p.ys[3]], p.fy);
    return cubic(col[0], col[1], col[2], col[3], p.fx);
}

// zoom_out's sampling loop (src/zoom.cpp:67-75) on the blurred copy, and zoom_in (:143-153) with the
// "*= 1 / zfactor" of src/tvl1occflow.cpp:447-451 applied to the stored value (mul = 0: none).
__global__ void __launch_bounds__(256) k_occ_resample(const double *__restrict__ in, double *__restrict__ out, int nx,
                                                  int ny, int nxx, int nyy, double fx, double fy, double mul)
{
    const int j1 = blockIdx.x * 32 + threadIdx.x, i1 = blockIdx.y * 8 + threadIdx.y;
    if (j1 >= nxx || i1 >= nyy) return;
    const double i2 = i1 / fy, j2 = j1 / fx;
    const BicubicPos p = bicubic_pos(j2, i2, nx, ny);
    double v = bicubic_eval(in + blockIdx.z * (size_t) nx * ny, p, nx);
    if (mul != 0.0) v = v * mul;
    out[blockIdx.z * (size_t) nxx * nyy + (size_t) i1 * nxx + j1] = v;
}

// ---------------------------------------------------------------------------------------------------
// per level: g, image gradients; per warp step: the six warps and the constants of the outer loop
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

// centered_gradient (src/operators.cpp:335-406) of I1 and I_1, and choosed_g with G_CHOICE 2
// (src/tvl1occflow.cpp:100-136): g = 1 / (1 + G_FACTOR |grad filtI0|).
__global__ void __launch_bounds__(256) k_occ_level_setup(const double *__restrict__ I1, const double *__restrict__ Im1,
                                                     const double *__restrict__ filt, double *__restrict__ I1x,
                                                     double *__restrict__ I1y, double *__restrict__ Im1x,
                                                     double *__restrict__ Im1y, double *__restrict__ g, int nx, int ny,
                                                     double g_factor)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t o = blockIdx.z * (size_t) nx * ny;
    const int k = i * nx + j;
    const int l = i * nx + clampi(j - 1, nx), r = i * nx + clampi(j + 1, nx);
    const int u = clampi(i - 1, ny) * nx + j, d = clampi(i + 1, ny) * nx + j;
    I1x[o + k] = 0.5 * (I1[o + r] - I1[o + l]);
    I1y[o + k] = 0.5 * (I1[o + d] - I1[o + u]);
    Im1x[o + k] = 0.5 * (Im1[o + r] - Im1[o + l]);
    Im1y[o + k] = 0.5 * (Im1[o + d] - Im1[o + u]);
    const double gx = 0.5 * (filt[o + r] - filt[o + l]), gy = 0.5 * (filt[o + d] - filt[o + u]);
    const double gggrad = sqrt(gx * gx + gy * gy);
    const double aux = 1. + g_factor * gggrad;
    g[o + k] = 1. / aux;
}

// src/tvl1occflow.cpp:232-270: I1, I1x, I1y sampled at x + u, I_1, I_1x, I_1y at x - u (border_out = false),
// grad1 / grad3, rho1_c / rho3_c.  I1w and I_1w are consumed here.
__global__ void __launch_bounds__(256) k_occ_warp(const double *__restrict__ I0, const double *__restrict__ I1,
                                              const double *__restrict__ I1x, const double *__restrict__ I1y,
                                              const double *__restrict__ Im1, const double *__restrict__ Im1x,
                                              const double *__restrict__ Im1y, const double *__restrict__ u1,
                                              const double *__restrict__ u2, double *__restrict__ I1wx,
                                              double *__restrict__ I1wy, double *__restrict__ Im1wx,
                                              double *__restrict__ Im1wy, double *__restrict__ rho1_c,
                                              double *__restrict__ rho3_c, double *__restrict__ grad1,
                                              double *__restrict__ grad3, int nx, int ny)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t o = blockIdx.z * (size_t) nx * ny;
    const size_t p = o + (size_t) i * nx + j;
    const double a = u1[p], b = u2[p];
    const BicubicPos f = bicubic_pos(j + a, i + b, nx, ny);
    const double w = bicubic_eval(I1 + o, f, nx), wx = bicubic_eval(I1x + o, f, nx), wy = bicubic_eval(I1y + o, f, nx);
    const BicubicPos q = bicubic_pos(j + (-a), i + (-b), nx, ny);
    const double m = bicubic_eval(Im1 + o, q, nx), mx = bicubic_eval(Im1x + o, q, nx), my = bicubic_eval(Im1y + o, q, nx);
    I1wx[p] = wx;
    I1wy[p] = wy;
    Im1wx[p] = mx;
    Im1wy[p] = my;
    double Ix2 = wx * wx, Iy2 = wy * wy;
    grad1[p] = Ix2 + Iy2;
    Ix2 = mx * mx;
    Iy2 = my * my;
    grad3[p] = Ix2 + Iy2;
    rho1_c[p] = w - wx * a - wy * b - I0[p];
    rho3_c[p] = m + mx * a + my * b - I0[p];
}

// ---------------------------------------------------------------------------------------------------
// outer loop, step 1: Solver_wrt_v (src/tvl1occflow_solvers.cpp:54-148), and the prologue of
// Solver_wrt_u (:188-203): f = v / theta + beta grad(chi), u = v + theta beta grad(chi).
// U and F are [2][B][N] (component, triple).
// ---------------------------------------------------------------------------------------------------
struct VParams {
    double l_t, one_pat, at_d_1pat, lt_d_1pat, theta, beta, is_zero, thr_chi;
};

__global__ void __launch_bounds__(256) k_occ_solver_v(const TripleCtl *__restrict__ ctl, double *__restrict__ U,
                                                  double *__restrict__ F, const double *__restrict__ chi,
                                                  const double *__restrict__ I1wx, const double *__restrict__ I1wy,
                                                  const double *__restrict__ Im1wx, const double *__restrict__ Im1wy,
                                                  const double *__restrict__ rho1_c, const double *__restrict__ rho3_c,
                                                  const double *__restrict__ grad1, const double *__restrict__ grad3,
                                                  double *__restrict__ Vfwd, double *__restrict__ Vbck, int nx, int ny,
                                                  int B, VParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t p = b * N + (size_t) i * nx + j;
    const double a = U[p], c = U[BN + p];
    const double wx = I1wx[p], wy = I1wy[p], mx = Im1wx[p], my = Im1wy[p], g1 = grad1[p], g3 = grad3[p];
    double d1, d2;
    const double rho1 = rho1_c[p] + (wx * a + wy * c);
    if (rho1 < -P.l_t * g1) { d1 = P.l_t * wx; d2 = P.l_t * wy; }
    else if (rho1 > P.l_t * g1) { d1 = -P.l_t * wx; d2 = -P.l_t * wy; }
    else if (g1 < P.is_zero) { d1 = d2 = 0; }
    else { d1 = -rho1 * wx / g1; d2 = -rho1 * wy / g1; }
    const double vf1 = a + d1, vf2 = c + d2;
    Vfwd[p] = vf1;
    Vfwd[BN + p] = vf2;

    double vb1, vb2;
    const double rho3 = rho3_c[p] - (mx * a + my * c);
    const double A = rho3 + P.at_d_1pat * (mx * a + my * c);
    if (A < -P.lt_d_1pat * g3) {
        d1 = -P.lt_d_1pat * mx; d2 = -P.lt_d_1pat * my;
        vb1 = (a / P.one_pat) + d1; vb2 = (c / P.one_pat) + d2;
    } else if (A > P.lt_d_1pat * g3) {
        d1 = P.lt_d_1pat * mx; d2 = P.lt_d_1pat * my;
        vb1 = (a / P.one_pat) + d1; vb2 = (c / P.one_pat) + d2;
    } else {
        if (g3 < P.is_zero) { d1 = d2 = 0; }
        else { d1 = rho3 * mx / g3; d2 = rho3 * my / g3; }
        vb1 = a + d1; vb2 = c + d2;
    }
    Vbck[p] = vb1;
    Vbck[BN + p] = vb2;
    const double x = chi[p];
    const double v1 = (x < P.thr_chi) ? vf1 : vb1, v2 = (x < P.thr_chi) ? vf2 : vb2;
    // forward_gradient of chi (src/operators.cpp:86-125)
    const double chix = (j < nx - 1) ? chi[p + 1] - x : 0, chiy = (i < ny - 1) ? chi[p + nx] - x : 0;
    F[p] = v1 / P.theta + P.beta * chix;
    F[BN + p] = v2 / P.theta + P.beta * chiy;
    U[p] = v1 + P.theta * P.beta * chix;
    U[BN + p] = v2 + P.theta * P.beta * chiy;
}

// ---------------------------------------------------------------------------------------------------
// outer loop, step 2: Scalar_ROF_BoxCellCentered (src/tvl1occflow_tv_rof_box.cpp:25-645) on the 2B
// planes (component k of triple b = problem q = k * B + b): U, F, AL are [2][B][N]; P is [4][B][N] =
// p11, p12, p21, p22, the dual values on the SOUTH (2k) and EAST (2k + 1) side of every cell.
// ---------------------------------------------------------------------------------------------------

// alfa = |grad u| / (lambda g), :186-202, with the reference's own hypot (:15-20).  g is [B][N].
__global__ void __launch_bounds__(256) k_occ_rof_alfa(const TripleCtl *__restrict__ ctl, const double *__restrict__ U,
                                                  const double *__restrict__ g, double *__restrict__ AL, int nx,
                                                  int ny, int B, double lambda)
{
    const int q = blockIdx.z, b = q % B;
    if (ctl && !ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny;
    const size_t c = (size_t) i * nx + j;
    const double *u = U + q * N;
    const double x = (j < nx - 1) ? u[c + 1] - u[c] : 0, y = (i < ny - 1) ? u[c + nx] - u[c] : 0;
    AL[q * N + c] = sqrt(x * x + y * y) / (lambda * g[b * N + c]);
}

// ---- wave layout ------------------------------------------------------------------------------------
// The Gauss-Seidel pass below runs cell (i, j) at step t = 2i + j, thread = row.  In row-major planes the
// cells of one step lie nx - 2 elements apart: every load of a warp touches 32 different sectors, and the
// pass is bound by L1 sector throughput (ncu: ~1.5 us per step, prefetching changed nothing).  In the WAVE
// layout   W[((j + 2i) mod nx) * ny + i]   the cells of a step are contiguous in i, and so are all their
// neighbours ((i+di, j+dj) lives in wave column t + dj + 2di, row i + di): every access of a step is
// coalesced.  j -> (j + 2i) mod nx is a permutation for fixed i, so a wave plane has exactly nx * ny
// elements.  The dual variables stay in this layout for a whole pyramid level; f, alfa and the
// coefficients are brought into it by 32 x 32 tile transposes.
__device__ __forceinline__ int wmod(int x, int m)
{
    x %= m;
    return x < 0 ? x + m : x;
}

// Tile transposes between row-major planes and the wave layout, planes [z][N]; a CTA moves a 32 x 32 tile of
// (wave column c, row i): the row-major side is coalesced along j (= c - 2i mod nx, consecutive in c), the
// wave side along i.
template <bool TO_WAVE>
__global__ void __launch_bounds__(256) k_occ_wave_transpose(const TripleCtl *__restrict__ ctl, const double *__restrict__ in,
                                                            double *__restrict__ out, int nx, int ny, int B)
{
    __shared__ double tile[32][33];
    const int z = blockIdx.z;
    if (ctl && !ctl[z % B].active) return;
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int c0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const size_t N = (size_t) nx * ny;
    const double *src = in + z * N;
    double *dst = out + z * N;
    if (TO_WAVE) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = i0 + ty + 8 * r, c = c0 + tx;
            if (i < ny && c < nx) tile[ty + 8 * r][tx] = src[(size_t) i * nx + wmod(c - 2 * i, nx)];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int c = c0 + ty + 8 * r, i = i0 + tx;
            if (i < ny && c < nx) dst[(size_t) c * ny + i] = tile[tx][ty + 8 * r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int c = c0 + ty + 8 * r, i = i0 + tx;
            if (i < ny && c < nx) tile[tx][ty + 8 * r] = src[(size_t) c * ny + i];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = i0 + ty + 8 * r, c = c0 + tx;
            if (i < ny && c < nx) dst[(size_t) i * nx + wmod(c - 2 * i, nx)] = tile[ty + 8 * r][tx];
        }
    }
}

// The 4x4 system of an interior cell (:385-482) is eliminated with coefficients that depend on alfa only,
// not on the sides: a, b, alf, gam, c and the three denominators.  They hold three of the seven divisions of
// a cell and sit on no dependence chain, so they are formed here, one thread per cell, and the serial
// Gauss-Seidel pass only reads them: K is [2B][kRofK][N] in the wave layout = a, b, alf, gam, c,
// b3 + gam + c (gam - 1), b2 + gam, b1 - a, b0 (same expressions, same bits).  Border cells keep working
// from alfa.  ALW = alfa in the wave layout; thread (i, c): i along x, so every access is coalesced.
constexpr int kRofK = 9;

__global__ void __launch_bounds__(256) k_occ_rof_coef(const TripleCtl *__restrict__ ctl, const double *__restrict__ ALW,
                                                      double *__restrict__ K, int nx, int ny, int B)
{
    const int q = blockIdx.z, bb = q % B;
    if (ctl && !ctl[bb].active) return;
    const int i = blockIdx.x * 32 + threadIdx.x, c = blockIdx.y * 8 + threadIdx.y;
    if (i < 1 || i >= ny - 1 || c >= nx) return;
    const int j = wmod(c - 2 * i, nx);
    if (j < 1 || j >= nx - 1) return;
    const size_t N = (size_t) nx * ny;
    const double *al = ALW + q * N;
    const double alc = al[(size_t) c * ny + i];
    const double b0 = -2 - al[(size_t) wmod(c - 1, nx) * ny + i], b1 = -2 - al[(size_t) wmod(c - 2, nx) * ny + i - 1];
    const double b2 = -2 - alc, b3 = -2 - alc;
    const double a = 1 / b0;
    const double b = -(b0 + 1) / (b0 * b1 - 1);
    const double alf = 1 + a;
    const double gam = -a + b * alf;
    const double cc = (1 - gam) / (b2 + gam);
    double *k = K + q * N * kRofK + (size_t) c * ny + i;
    k[0] = a;
    k[N] = b;
    k[2 * N] = alf;
    k[3 * N] = gam;
    k[4 * N] = cc;
    k[5 * N] = b3 + gam + cc * (gam - 1);
    k[6 * N] = b2 + gam;
    k[7 * N] = b1 - a;
    k[8 * N] = b0;
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// How a cell of the Gauss-Seidel pass reaches the sides, f, alfa and the coefficients of a neighbour
// (di rows down, dj columns right): row-major planes, or the wave layout with the step's column bases
// cb[d + 4] = ((t + d) mod nx) * ny, d = -4 .. 2 (the same for every row of the step).
struct RofRowMajor {
    double *pS, *pE;
    const double *f, *al;
    int nx, c;
    __device__ __forceinline__ double &ps(int di, int dj) const { return pS[c + di * nx + dj]; }
    __device__ __forceinline__ double &pe(int di, int dj) const { return pE[c + di * nx + dj]; }
    __device__ __forceinline__ double fv(int di, int dj) const { return f[c + di * nx + dj]; }
    __device__ __forceinline__ double alv(int di, int dj) const { return al[c + di * nx + dj]; }
    __device__ __forceinline__ double kv(int) const { return 0.0; }
};

struct RofWave {
    double *pS, *pE;
    const double *f, *al, *K;
    size_t N;
    int cb[7];
    int i;
    __device__ __forceinline__ double &ps(int di, int dj) const { return pS[cb[dj + 2 * di + 4] + i + di]; }
    __device__ __forceinline__ double &pe(int di, int dj) const { return pE[cb[dj + 2 * di + 4] + i + di]; }
    __device__ __forceinline__ double fv(int di, int dj) const { return f[cb[dj + 2 * di + 4] + i + di]; }
    __device__ __forceinline__ double alv(int di, int dj) const { return al[cb[dj + 2 * di + 4] + i + di]; }
    __device__ __forceinline__ double kv(int m) const { return K[m * N + cb[4] + i]; }
};

// One cell of the Gauss-Seidel pass: re-solves the 2x2 / 3x3 / 4x4 system of the cell's own sides with
// relaxation omega (corner :206-247 / :304-331 / ..., edge :250-301 / :334-383 / ..., interior :385-482).
// The sides are read and written through ordinary (coherent) accesses: other threads of the CTA wrote the
// values this cell needs at earlier wavefront steps.  USE_K: interior coefficients from k_occ_rof_coef.
template <bool USE_K, class Acc>
__device__ __forceinline__ void rof_cell(int i, int j, int nx, int ny, const Acc &A, double omega)
{
    const bool hasW = j > 0, hasN = i > 0, hasS = i < ny - 1, hasE = j < nx - 1;
    const bool interior = hasW && hasN && hasS && hasE;
    double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    if (!(USE_K && interior)) {
        b0 = hasW ? -2 - A.alv(0, -1) : 0; b1 = hasN ? -2 - A.alv(-1, 0) : 0;
        b2 = hasS ? -2 - A.alv(0, 0) : 0; b3 = hasE ? -2 - A.alv(0, 0) : 0;
    }
    double W = 0, N = 0, S = 0, E = 0;
    const bool n_edge = !hasN && hasW && hasE;
    const double fc = A.fv(0, 0);
    if (hasW) {
        const double jm3 = j >= 2 ? A.pe(0, -2) : 0.0, ip1_jm2 = A.ps(0, -1), im1_jm2 = hasN ? A.ps(-1, -1) : 0.0;
        const double Fw = fc - A.fv(0, -1);
        W = n_edge ? -Fw - jm3 + ip1_jm2 - im1_jm2 : -jm3 + ip1_jm2 - im1_jm2 - Fw;
    }
    if (hasN) {
        const double im3 = i >= 2 ? A.ps(-2, 0) : 0.0, im2_jp1 = A.pe(-1, 0), im2_jm1 = hasW ? A.pe(-1, -1) : 0.0;
        const double Fn = fc - A.fv(-1, 0);
        N = -im3 + im2_jp1 - im2_jm1 - Fn;
    }
    if (hasS) {
        const double ip3 = A.ps(1, 0), ip2_jp1 = A.pe(1, 0), ip2_jm1 = hasW ? A.pe(1, -1) : 0.0;
        const double Fs = A.fv(1, 0) - fc;
        S = n_edge ? -Fs - ip3 - ip2_jp1 + ip2_jm1 : -ip3 - ip2_jp1 + ip2_jm1 - Fs;
    }
    if (hasE) {
        const double jp3 = A.pe(0, 1), ip1_jp2 = A.ps(0, 1), im1_jp2 = hasN ? A.ps(-1, 1) : 0.0;
        const double Fe = A.fv(0, 1) - fc;
        E = n_edge ? -Fe - jp3 - ip1_jp2 + im1_jp2 : -jp3 - ip1_jp2 + im1_jp2 - Fe;
    }
    double den;
#define RELAX_(q, num) ((q) = (1 - omega) * (q) + omega * (num) / den)
    if (interior) {
        double a, b, alf, gam, cc, dE, dS, dN, dW;
        if (USE_K) {
            a = A.kv(0); b = A.kv(1); alf = A.kv(2); gam = A.kv(3); cc = A.kv(4);
            dE = A.kv(5); dS = A.kv(6); dN = A.kv(7); dW = A.kv(8);
        } else {
            a = 1 / b0;
            b = -(b0 + 1) / (b0 * b1 - 1);
            alf = 1 + a;
            gam = -a + b * alf;
            cc = (1 - gam) / (b2 + gam);
            dE = b3 + gam + cc * (gam - 1);
            dS = b2 + gam;
            dN = b1 - a;
            dW = b0;
        }
        const double x = N + a * W;
        const double y = -a * W + b * x;
        const double e = (1 - omega) * A.pe(0, 0) + omega * (E + y + cc * (S + y)) / dE;
        A.pe(0, 0) = e;
        const double s = (1 - omega) * A.ps(0, 0) + omega * (S + y + e * (1 - gam)) / dS;
        A.ps(0, 0) = s;
        const double n = (1 - omega) * A.ps(-1, 0) + omega * (x - alf * (e + s)) / dN;
        A.ps(-1, 0) = n;
        A.pe(0, -1) = (1 - omega) * A.pe(0, -1) + omega * (W + n - s - e) / dW;
    } else if (!hasN && !hasW) {
        den = b2 * b3 - 1;
        RELAX_(A.ps(0, 0), S * b3 + E);
        RELAX_(A.pe(0, 0), E * b2 + S);
    } else if (!hasN && !hasE) {
        den = b0 * b2 - 1;
        RELAX_(A.pe(0, -1), W * b2 - S);
        RELAX_(A.ps(0, 0), S * b0 - W);
    } else if (!hasN) {
        den = b0 * b2 * b3 - b0 - b2 - b3 - 2;
        RELAX_(A.pe(0, -1), W * b2 * b3 - E * b2 - S * b3 - W - E - S);
        RELAX_(A.ps(0, 0), S * b0 * b3 - W * b3 + E * b0 - W + E - S);
        RELAX_(A.pe(0, 0), E * b0 * b2 - W * b2 + S * b0 - W - E + S);
    } else if (!hasS && !hasW) {
        den = b3 * b1 - 1;
        RELAX_(A.ps(-1, 0), b3 * N - E);
        RELAX_(A.pe(0, 0), b1 * E - N);
    } else if (!hasS && !hasE) {
        den = b0 * b1 - 1;
        RELAX_(A.pe(0, -1), W * b1 + N);
        RELAX_(A.ps(-1, 0), N * b0 + W);
    } else if (!hasS) {
        den = b0 * b1 * b3 - b0 - b1 - b3 - 2;
        RELAX_(A.pe(0, -1), W * b1 * b3 - E + N - E * b1 - W + N * b3);
        RELAX_(A.ps(-1, 0), N * b0 * b3 + W - E - N - E * b0 + W * b3);
        RELAX_(A.pe(0, 0), E * b0 * b1 - N - W - W * b1 - N * b0 - E);
    } else if (!hasW) {
        den = b1 * b2 * b3 - (b1 + b2 + b3) - 2;
        RELAX_(A.ps(-1, 0), b2 * b3 * N - E * b2 - S * b3 - N - S - E);
        RELAX_(A.ps(0, 0), b1 * b3 * S + E * b1 - N * b3 - N - S + E);
        RELAX_(A.pe(0, 0), b1 * b2 * E - N * b2 + S * b1 - N + S - E);
    } else {
        den = (b0 * b1 * b2) + (-b0 - b1 - b2 - 2);
        RELAX_(A.pe(0, -1), W * b1 * b2 - S + N - S * b1 - W + N * b2);
        RELAX_(A.ps(-1, 0), N * b0 * b2 + W - S - N - S * b0 + W * b2);
        RELAX_(A.ps(0, 0), S * b0 * b1 - N - W - W * b1 - N * b0 - S);
    }
#undef RELAX_
}

// The Gauss-Seidel pass of one sweep, one CTA per plane.  The reference visits the cells in row-major
// order; cell (i, j) reads sides written by (i, j-1) and (i-1, j+1) (and cells before them) and must see
// the OLD sides of (i, j+1), (i+1, j-1) and later cells.  Step t = 2i + j keeps exactly that: both
// predecessors run at t - 1, both successors at t + 1, and the cells of one step -- (i, j) and
// (i-1, j+2), ... -- touch disjoint sides.  Thread r owns rows r, r + blockDim.x, ...; one barrier per step.
// Row-major planes (the A/B reference of the wave kernel below, OCC_GS_WAVE=0).
__global__ void __launch_bounds__(1024) k_occ_rof_gs(const TripleCtl *__restrict__ ctl, double *P,
                                                 const double *__restrict__ F, const double *__restrict__ AL,
                                                 int nx, int ny, int B, double omega)
{
    const int q = blockIdx.x, k = q / B, b = q % B;
    if (ctl && !ctl[b].active) return;
    const size_t N = (size_t) nx * ny;
    RofRowMajor A;
    A.pS = P + ((size_t) (2 * k) * B + b) * N;
    A.pE = P + ((size_t) (2 * k + 1) * B + b) * N;
    A.f = F + q * N;
    A.al = AL + q * N;
    A.nx = nx;
    const int steps = 2 * (ny - 1) + nx;
    for (int t = 0; t < steps; t++) {
        for (int i = threadIdx.x; i < ny; i += blockDim.x) {
            const int j = t - 2 * i;
            if (j < 0) break;
            if (j < nx) {
                A.c = i * nx + j;
                rof_cell<false>(i, j, nx, ny, A, omega);
            }
        }
        __syncthreads();
    }
}

// The same pass on the wave layout (the default): P, FW, ALW, K are wave planes.  At step t every active
// row works in wave column t mod nx, so the seven column bases a cell needs are the same for the whole CTA.
// Rows whose row number is a multiple of four pull the sectors of the column kPf steps ahead into L2 (the
// step is a chain of dependent operations: it must not wait for DRAM).
template <bool USE_K>
__global__ void __launch_bounds__(1024) k_occ_rof_gs_wave(const TripleCtl *__restrict__ ctl, double *P,
                                                      const double *__restrict__ FW, const double *__restrict__ ALW,
                                                      const double *__restrict__ Kc, int nx, int ny, int B, double omega)
{
    const int q = blockIdx.x, k = q / B, b = q % B;
    if (ctl && !ctl[b].active) return;
    constexpr int kPf = 6;
    const size_t N = (size_t) nx * ny;
    RofWave A;
    A.pS = P + ((size_t) (2 * k) * B + b) * N;
    A.pE = P + ((size_t) (2 * k + 1) * B + b) * N;
    A.f = FW + q * N;
    A.al = ALW + q * N;
    A.K = Kc + q * N * kRofK;
    A.N = N;
    const int steps = 2 * (ny - 1) + nx;
    int w = 0;                                      // t mod nx
    for (int t = 0; t < steps; t++) {
#pragma unroll
        for (int d = -4; d <= 2; d++) {
            int c = w + d;
            c = c < 0 ? c + nx : (c >= nx ? c - nx : c);
            if (nx < 4) c = wmod(w + d, nx);
            A.cb[d + 4] = c * ny;
        }
        int cpf = w + kPf;
        cpf = cpf >= nx ? wmod(cpf, nx) : cpf;
        for (int i = threadIdx.x; i < ny; i += blockDim.x) {
            const int j = t - 2 * i;
            if (j < 0) break;
            if (j < nx) {
                if ((i & 3) == 0 && j + kPf < nx) {
                    const size_t o = (size_t) cpf * ny + i;
                    prefetch_l2(A.pS + o);
                    prefetch_l2(A.pE + o);
                    prefetch_l2(A.f + o);
                    if (USE_K) {
#pragma unroll
                        for (int m = 0; m < kRofK; m++) prefetch_l2(A.K + m * N + o);
                    } else {
                        prefetch_l2(A.al + o);
                    }
                }
                A.i = i;
                rof_cell<USE_K>(i, j, nx, ny, A, omega);
            }
        }
        __syncthreads();
        w = w + 1 == nx ? 0 : w + 1;
    }
}

// u = lambda f + lambda (pS - pN + pE - pW), :557-585
__global__ void __launch_bounds__(256) k_occ_rof_u(const TripleCtl *__restrict__ ctl, double *__restrict__ U,
                                               const double *__restrict__ F, const double *__restrict__ P, int nx,
                                               int ny, int B, double lambda)
{
    const int q = blockIdx.z, k = q / B, b = q % B;
    if (ctl && !ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny;
    const size_t c = (size_t) i * nx + j;
    const double *pS = P + ((size_t) (2 * k) * B + b) * N, *pE = P + ((size_t) (2 * k + 1) * B + b) * N;
    const double pn = i > 0 ? pS[c - nx] : 0.0, pw = j > 0 ? pE[c - 1] : 0.0;
    U[q * N + c] = lambda * F[q * N + c] + lambda * (pS[c] - pn + pE[c] - pw);
}

// ---------------------------------------------------------------------------------------------------
// outer loop: me_median_filtering, window 3 (src/utils.cpp:151-213), symmetric boundary.  in / out [2][B][N].
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_occ_median3(const TripleCtl *__restrict__ ctl, const double *__restrict__ in,
                                                 double *__restrict__ out, int nx, int ny, int B)
{
    const int q = blockIdx.z, b = q % B;
    if (ctl && !ctl[b].active) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const size_t N = (size_t) nx * ny;
    const double *src = in + q * N;
    double v[9];
    int n = 0;
#pragma unroll
    for (int yy = -1; yy <= 1; yy++)
#pragma unroll
        for (int xx = -1; xx <= 1; xx++) {
            int x0 = x + xx, y0 = y + yy;
            if (x0 < 0) x0 = -x0 - 1;
            if (x0 >= nx) x0 = 2 * nx - x0 - 1;
            if (y0 < 0) y0 = -y0 - 1;
            if (y0 >= ny) y0 = 2 * ny - y0 - 1;
            v[n++] = src[(size_t) y0 * nx + x0];
        }
    // the reference's insertion sort (:196-206) as adjacent exchanges on strict ">": stable, so even the
    // sign of a zero among equal values comes out as upstream
#pragma unroll
    for (int a = 1; a < 9; a++)
#pragma unroll
        for (int k = a - 1; k >= 0; k--) {
            const double lo = v[k], hi = v[k + 1];
            const bool sw = lo > hi;
            v[k] = sw ? hi : lo;
            v[k + 1] = sw ? lo : hi;
        }
    out[q * N + (size_t) y * nx + x] = v[4];
}

// ---------------------------------------------------------------------------------------------------
// outer loop, step 3: Solver_wrt_chi (src/tvl1occflow_solvers.cpp:215-338).  Everything that does not
// change during its 100 iterations is formed once (k_occ_chi_setup): F and G of both branches of :300-318
// and beta div(u); an iteration is the dual step + projection (k_occ_chi_eta, :262-268 and :33-52) and the
// primal step (k_occ_chi_update, :270-325).  C is [5][B][N] = F0, G0, F1, G1, beta div u; ETA is [2][B][N].
// ---------------------------------------------------------------------------------------------------

// divergence (src/operators.cpp:35-78) with its per-case association order; a(x) / b(x) give v1 / v2
template <class A1, class A2>
__device__ __forceinline__ double divergence_at(int i, int j, int nx, int ny, A1 v1, A2 v2)
{
    const int p = i * nx + j;
    if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
        const double v1x = v1(p) - v1(p - 1);
        const double v2y = v2(p) - v2(p - nx);
        return v1x + v2y;
    }
    double a = 0;
    bool first = true;
    if (j < nx - 1) { a = v1(p); first = false; }
    if (j > 0) { a = first ? -v1(p - 1) : a - v1(p - 1); first = false; }
    if (i < ny - 1) { a = first ? v2(p) : a + v2(p); first = false; }
    if (i > 0) { a = first ? -v2(p - nx) : a - v2(p - nx); first = false; }
    return a;
}

struct ChiParams {
    double lambda, half_over_theta, alpha_theta, beta, tau_chi, tau_eta, is_zero;
};

__global__ void __launch_bounds__(256) k_occ_chi_setup(const TripleCtl *__restrict__ ctl, const double *__restrict__ U,
                                                   const double *__restrict__ I1wx, const double *__restrict__ I1wy,
                                                   const double *__restrict__ Im1wx, const double *__restrict__ Im1wy,
                                                   const double *__restrict__ rho1_c, const double *__restrict__ rho3_c,
                                                   const double *__restrict__ Vfwd, const double *__restrict__ Vbck,
                                                   double *__restrict__ C, int nx, int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t p = b * N + (size_t) i * nx + j;
    const double *u1 = U + b * N, *u2 = U + BN + b * N;
    const double a = U[p], c = U[BN + p];
    const double vf1 = Vfwd[p], vf2 = Vfwd[BN + p], vb1 = Vbck[p], vb2 = Vbck[BN + p];
    const double rho1 = rho1_c[p] + (I1wx[p] * vf1 + I1wy[p] * vf2);
    const double abs_rho1 = (rho1 < 0.) ? -rho1 : rho1;
    const double rho3 = rho3_c[p] - (Im1wx[p] * vb1 + Im1wy[p] * vb2);
    const double abs_rho3 = (rho3 < 0.) ? -rho3 : rho3;
    C[p] = -P.lambda * abs_rho1;
    C[BN + p] = -P.half_over_theta * ((vf1 - a) * (vf1 - a) + (vf2 - c) * (vf2 - c));
    C[2 * BN + p] = P.lambda * abs_rho3;
    C[3 * BN + p] = P.half_over_theta * ((vb1 - a) * (vb1 - a) + (vb2 - c) * (vb2 - c))
                    + P.alpha_theta * (vb1 * vb1 + vb2 * vb2);
    const double div_u = divergence_at(i, j, nx, ny, [&](int x) { return u1[x]; }, [&](int x) { return u2[x]; });
    C[4 * BN + p] = P.beta * div_u;
}

__global__ void __launch_bounds__(256) k_occ_chi_eta(const TripleCtl *__restrict__ ctl, const double *__restrict__ chi,
                                                 const double *__restrict__ g, double *__restrict__ ETA, int nx,
                                                 int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t p = b * N + (size_t) i * nx + j;
    const double x = chi[p], gg = g[p];
    const double chix = (j < nx - 1) ? chi[p + 1] - x : 0, chiy = (i < ny - 1) ? chi[p + nx] - x : 0;
    double e1 = ETA[p] + P.tau_eta * gg * chix;
    double e2 = ETA[BN + p] + P.tau_eta * gg * chiy;
    const double norm2 = e1 * e1 + e2 * e2;
    if (norm2 < P.is_zero) { e1 = 0.0; e2 = 0.0; }
    else { const double norm = sqrt(norm2); e1 = e1 / norm; e2 = e2 / norm; }
    ETA[p] = e1;
    ETA[BN + p] = e2;
}

__global__ void __launch_bounds__(256) k_occ_chi_update(const TripleCtl *__restrict__ ctl, double *__restrict__ chi,
                                                    const double *__restrict__ g, const double *__restrict__ ETA,
                                                    const double *__restrict__ C, int nx, int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t p = b * N + (size_t) i * nx + j;
    const double *gb = g + b * N, *e1 = ETA + b * N, *e2 = ETA + BN + b * N;
    const double div_eta = divergence_at(i, j, nx, ny, [&](int x) { return gb[x] * e1[x]; },
                                         [&](int x) { return gb[x] * e2[x]; });
    double x = chi[p];
    const int br = (x < 0.5) ? 0 : 2;
    const double Fv = C[br * BN + p], Gv = C[(br + 1) * BN + p];
    x = x + P.tau_chi * (div_eta - Fv - Gv - C[4 * BN + p]);
    if (x > 1.) x = 1.;
    else if (x < 0.) x = 0.;
    chi[p] = x;
}

// Both steps of one iteration in one pass (the default; the two kernels above are the A/B reference and
// give the same bits): a CTA owns a 32 x 16 tile, projects the dual variable on the tile plus the column
// to its left and the row above it (the values the divergence at its own pixels reads), keeps g * eta in
// shared memory and updates chi.  chi and eta are read from one buffer and written to another (a
// neighbouring tile still needs the old values): 100 iterations end in the buffer they started from.
constexpr int kChiTW = 32, kChiTH = 16;

__global__ void __launch_bounds__(256) k_occ_chi_fused(const TripleCtl *__restrict__ ctl, const double *__restrict__ chi_in,
                                                       double *__restrict__ chi_out, const double *__restrict__ g,
                                                       const double *__restrict__ eta_in, double *__restrict__ eta_out,
                                                       const double *__restrict__ C, int nx, int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    __shared__ double s1[kChiTH + 1][kChiTW + 1], s2[kChiTH + 1][kChiTW + 1];     // g * eta1, g * eta2 at (y0-1.., x0-1..)
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t o = b * N;
    const int x0 = blockIdx.x * kChiTW, y0 = blockIdx.y * kChiTH;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < (kChiTH + 1) * (kChiTW + 1); k += 256) {
        const int ly = k / (kChiTW + 1), lx = k - ly * (kChiTW + 1);
        const int i = y0 - 1 + ly, j = x0 - 1 + lx;
        if (i < 0 || j < 0 || i >= ny || j >= nx) continue;
        const size_t p = o + (size_t) i * nx + j;
        const double x = chi_in[p], gg = g[p];
        const double chix = (j < nx - 1) ? chi_in[p + 1] - x : 0, chiy = (i < ny - 1) ? chi_in[p + nx] - x : 0;
        double e1 = eta_in[p] + P.tau_eta * gg * chix;
        double e2 = eta_in[BN + p] + P.tau_eta * gg * chiy;
        const double norm2 = e1 * e1 + e2 * e2;
        if (norm2 < P.is_zero) { e1 = 0.0; e2 = 0.0; }
        else { const double norm = sqrt(norm2); e1 = e1 / norm; e2 = e2 / norm; }
        if (ly && lx) { eta_out[p] = e1; eta_out[BN + p] = e2; }
        s1[ly][lx] = gg * e1;
        s2[ly][lx] = gg * e2;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kChiTH / 8; r++) {
        const int ly = threadIdx.y + 8 * r + 1, lx = threadIdx.x + 1;
        const int i = y0 + ly - 1, j = x0 + lx - 1;
        if (i >= ny || j >= nx) continue;
        const size_t p = o + (size_t) i * nx + j;
        // divergence (src/operators.cpp:35-78) of (g eta1, g eta2), per-case association order
        double div_eta;
        if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
            const double v1x = s1[ly][lx] - s1[ly][lx - 1];
            const double v2y = s2[ly][lx] - s2[ly - 1][lx];
            div_eta = v1x + v2y;
        } else {
            double a = 0;
            bool first = true;
            if (j < nx - 1) { a = s1[ly][lx]; first = false; }
            if (j > 0) { a = first ? -s1[ly][lx - 1] : a - s1[ly][lx - 1]; first = false; }
            if (i < ny - 1) { a = first ? s2[ly][lx] : a + s2[ly][lx]; first = false; }
            if (i > 0) { a = first ? -s2[ly - 1][lx] : a - s2[ly - 1][lx]; first = false; }
            div_eta = a;
        }
        double x = chi_in[p];
        const int br = (x < 0.5) ? 0 : 2;
        const double Fv = C[br * BN + p], Gv = C[(br + 1) * BN + p];
        x = x + P.tau_chi * (div_eta - Fv - Gv - C[4 * BN + p]);
        if (x > 1.) x = 1.;
        else if (x < 0.) x = 0.;
        chi_out[p] = x;
    }
}

// The same iteration as a row march (opt-in, OCC_CHI_MARCH=1): a warp owns 31 columns x kChiMR rows.  Lane l holds
// column x0 + l - 1: lanes 1..31 own their pixel, lane 0 only evaluates the dual variable of the column to the
// left, so the backward x-difference of the divergence is a shuffle; the warp walks down its strip carrying
// chi of the next row and g * eta2 of the row above in registers, and starts one row above the strip for the
// backward y-difference.  No shared memory, no barrier, no index division: k_occ_chi_fused spent two thirds
// of its instructions on 64-bit index arithmetic and control (ncu source page: IMAD 16 %, DFMA+DMUL+DADD 18 %).
// Same expressions, same bits.  chi / eta are read from one buffer and written to the other, as above.
// Measured: 229.5 ms against 219.8 for the tiled kernel per 148 triples (and 302.7 with L2 prefetches of the
// rows ahead): the instructions were not the limiter -- with ten fp64 planes per pixel-iteration the loop
// sits on HBM -- so the tiled kernel stays the default.
constexpr int kChiMR = 16;

__global__ void __launch_bounds__(128) k_occ_chi_march(const TripleCtl *__restrict__ ctl, const double *__restrict__ chi_in,
                                                       double *__restrict__ chi_out, const double *__restrict__ g,
                                                       const double *__restrict__ eta_in, double *__restrict__ eta_out,
                                                       const double *__restrict__ C, int nx, int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    const int lane = threadIdx.x;
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * kChiMR;
    if (y0 >= ny) return;                                        // warp-uniform
    const int y1 = min(y0 + kChiMR, ny);
    const int j = blockIdx.x * 31 + lane - 1;
    const bool col = j >= 0 && j < nx;                           // this lane has a pixel at all
    const bool own = lane >= 1 && col;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const double *chi_b = chi_in + b * N, *g_b = g + b * N, *e1_b = eta_in + b * N, *e2_b = eta_in + BN + b * N;
    const double *c_b = C + b * N;
    double *chi_o = chi_out + b * N, *e1_o = eta_out + b * N, *e2_o = eta_out + BN + b * N;
    const int jc = col ? j : 0;                                  // a valid address for idle lanes
    const bool has_right = col && j < nx - 1;
    const int i_first = y0 > 0 ? y0 - 1 : 0;
    int c = i_first * nx + jc;                                   // N < 2^30 (checked by the host)
    double x = col ? chi_b[c] : 0.0;                             // chi of the current row
    double up = 0.0;                                             // g * eta2 of the row above
    for (int i = i_first; i < y1; i++, c += nx) {
        const bool below = i < ny - 1;
        const double xn = (col && below) ? chi_b[c + nx] : 0.0;  // chi of the next row (its x on the next turn)
        double xr = __shfl_down_sync(0xffffffffu, x, 1);         // chi of the column to the right
        if (lane == 31 && has_right) xr = chi_b[c + 1];
        double ge1 = 0.0, ge2 = 0.0;
        if (col) {
            const double gg = g_b[c];
            const double chix = has_right ? xr - x : 0, chiy = below ? xn - x : 0;
            double e1 = e1_b[c] + P.tau_eta * gg * chix;
            double e2 = e2_b[c] + P.tau_eta * gg * chiy;
            const double norm2 = e1 * e1 + e2 * e2;
            if (norm2 < P.is_zero) { e1 = 0.0; e2 = 0.0; }
            else { const double norm = sqrt(norm2); e1 = e1 / norm; e2 = e2 / norm; }
            if (own && i >= y0) { e1_o[c] = e1; e2_o[c] = e2; }
            ge1 = gg * e1;
            ge2 = gg * e2;
        }
        const double left = __shfl_up_sync(0xffffffffu, ge1, 1);
        if (own && i >= y0) {
            // divergence (src/operators.cpp:35-78) of (g eta1, g eta2), per-case association order
            double div_eta;
            if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
                const double v1x = ge1 - left;
                const double v2y = ge2 - up;
                div_eta = v1x + v2y;
            } else {
                double a = 0;
                bool first = true;
                if (j < nx - 1) { a = ge1; first = false; }
                if (j > 0) { a = first ? -left : a - left; first = false; }
                if (i < ny - 1) { a = first ? ge2 : a + ge2; first = false; }
                if (i > 0) { a = first ? -up : a - up; first = false; }
                div_eta = a;
            }
            const double *cc = c_b + c + (x < 0.5 ? (size_t) 0 : 2 * BN);
            double v = x + P.tau_chi * (div_eta - cc[0] - cc[BN] - c_b[4 * BN + c]);
            if (v > 1.) v = 1.;
            else if (v < 0.) v = 0.;
            chi_o[c] = v;
        }
        up = ge2;
        x = xn;
    }
}

// Temporal blocking of the same iteration (opt-in, OCC_CHI_TB=1: measured SLOWER than k_occ_chi_fused, 253 vs
// 221 ms per 148 triples -- with a square root and two divisions per pixel-iteration in fp64 the loop is bound
// by the fp64 pipe, and the halo recompute costs more than the saved HBM traffic).  Solver_wrt_chi runs a FIXED number of iterations
// (MAX_ITERATIONS_CHI, no stopping rule), chi(p) after one iteration depends on the values within one pixel
// of p, and nothing else changes meanwhile: a CTA loads its 32 x 32 tile plus a halo of kChiT pixels (the
// three evolving planes into shared memory, the six constants of each of a thread's pixels into registers),
// runs kChiT complete iterations on chip -- the values of the outermost ring go stale by one pixel per
// iteration and never reach the tile -- and stores the tile.  HBM traffic per pixel-iteration falls from
// ~90 B to ~25 B; the arithmetic (one square root and two divisions per pixel-iteration, fp64) is done
// (32 + 2 kChiT)^2 / 32^2 / ... ~ 1.4 times.  Same formulas, same bits as the kernels above.
constexpr int kChiT = 5;                                     // divides MAX_ITERATIONS_CHI, even number of launches
constexpr int kChiBW = 32, kChiBH = 32;
constexpr int kChiRW = kChiBW + 2 * kChiT, kChiRH = kChiBH + 2 * kChiT;
constexpr int kChiTbThreads = 512;
constexpr int kChiSlots = (kChiRW * kChiRH + kChiTbThreads - 1) / kChiTbThreads;
constexpr size_t kChiTbSmem = 5 * sizeof(double) * kChiRW * kChiRH;

__global__ void __launch_bounds__(kChiTbThreads) k_occ_chi_tb(const TripleCtl *__restrict__ ctl,
                                                              const double *__restrict__ chi_in, double *__restrict__ chi_out,
                                                              const double *__restrict__ g, const double *__restrict__ eta_in,
                                                              double *__restrict__ eta_out, const double *__restrict__ C,
                                                              int nx, int ny, int B, ChiParams P)
{
    const int b = blockIdx.z;
    if (!ctl[b].active) return;
    extern __shared__ double s_chi_tb[];
    double *s_chi = s_chi_tb, *s_e1 = s_chi + kChiRW * kChiRH, *s_e2 = s_e1 + kChiRW * kChiRH,
           *s_g1 = s_e2 + kChiRW * kChiRH, *s_g2 = s_g1 + kChiRW * kChiRH;
    const size_t N = (size_t) nx * ny, BN = (size_t) B * N;
    const size_t o = b * N;
    const int rx0 = blockIdx.x * kChiBW - kChiT, ry0 = blockIdx.y * kChiBH - kChiT;
    const int tid = threadIdx.x;
    // this thread's pixels of the region (slot s: region index tid + s * threads)
    int lpos[kChiSlots];                 // ly * kChiRW + lx, or -1 outside the region / image
    double cg[kChiSlots], cF0[kChiSlots], cG0[kChiSlots], cF1[kChiSlots], cG1[kChiSlots], cbd[kChiSlots];
#pragma unroll
    for (int s = 0; s < kChiSlots; s++) {
        const int k = tid + s * kChiTbThreads;
        const int ly = k / kChiRW, lx = k - ly * kChiRW;
        const int i = ry0 + ly, j = rx0 + lx;
        lpos[s] = -1;
        cg[s] = cF0[s] = cG0[s] = cF1[s] = cG1[s] = cbd[s] = 0.0;
        if (k < kChiRW * kChiRH && i >= 0 && j >= 0 && i < ny && j < nx) {
            const size_t p = o + (size_t) i * nx + j;
            lpos[s] = k;
            s_chi[k] = chi_in[p];
            s_e1[k] = eta_in[p];
            s_e2[k] = eta_in[BN + p];
            cg[s] = g[p];
            cF0[s] = C[p];
            cG0[s] = C[BN + p];
            cF1[s] = C[2 * BN + p];
            cG1[s] = C[3 * BN + p];
            cbd[s] = C[4 * BN + p];
        }
    }
    __syncthreads();
    for (int it = 0; it < kChiT; it++) {
        // dual step + projection (:262-268, :33-52) at every pixel of the region
#pragma unroll
        for (int s = 0; s < kChiSlots; s++) {
            const int k = lpos[s];
            if (k < 0) continue;
            const int ly = k / kChiRW, lx = k - ly * kChiRW;
            const int i = ry0 + ly, j = rx0 + lx;
            const double x = s_chi[k], gg = cg[s];
            // forward differences: 0 on the last column / row of the IMAGE; at the edge of the region the
            // neighbour is missing and the value is one of those that go stale
            const double chix = (j < nx - 1 && lx + 1 < kChiRW) ? s_chi[k + 1] - x : 0;
            const double chiy = (i < ny - 1 && ly + 1 < kChiRH) ? s_chi[k + kChiRW] - x : 0;
            double e1 = s_e1[k] + P.tau_eta * gg * chix;
            double e2 = s_e2[k] + P.tau_eta * gg * chiy;
            const double norm2 = e1 * e1 + e2 * e2;
            if (norm2 < P.is_zero) { e1 = 0.0; e2 = 0.0; }
            else { const double norm = sqrt(norm2); e1 = e1 / norm; e2 = e2 / norm; }
            s_e1[k] = e1;
            s_e2[k] = e2;
            s_g1[k] = gg * e1;
            s_g2[k] = gg * e2;
        }
        __syncthreads();
        // primal step (:270-325)
#pragma unroll
        for (int s = 0; s < kChiSlots; s++) {
            const int k = lpos[s];
            if (k < 0) continue;
            const int ly = k / kChiRW, lx = k - ly * kChiRW;
            const int i = ry0 + ly, j = rx0 + lx;
            const double left = lx > 0 ? s_g1[k - 1] : 0.0, up = ly > 0 ? s_g2[k - kChiRW] : 0.0;
            double div_eta;
            if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
                const double v1x = s_g1[k] - left;
                const double v2y = s_g2[k] - up;
                div_eta = v1x + v2y;
            } else {
                double a = 0;
                bool first = true;
                if (j < nx - 1) { a = s_g1[k]; first = false; }
                if (j > 0) { a = first ? -left : a - left; first = false; }
                if (i < ny - 1) { a = first ? s_g2[k] : a + s_g2[k]; first = false; }
                if (i > 0) { a = first ? -up : a - up; first = false; }
                div_eta = a;
            }
            double x = s_chi[k];
            const bool lo = x < 0.5;
            const double Fv = lo ? cF0[s] : cF1[s], Gv = lo ? cG0[s] : cG1[s];
            x = x + P.tau_chi * (div_eta - Fv - Gv - cbd[s]);
            if (x > 1.) x = 1.;
            else if (x < 0.) x = 0.;
            s_chi[k] = x;
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < kChiSlots; s++) {
        const int k = lpos[s];
        if (k < 0) continue;
        const int ly = k / kChiRW, lx = k - ly * kChiRW;
        if (ly < kChiT || ly >= kChiT + kChiBH || lx < kChiT || lx >= kChiT + kChiBW) continue;
        const size_t p = o + (size_t) (ry0 + ly) * nx + rx0 + lx;
        chi_out[p] = s_chi[k];
        eta_out[p] = s_e1[k];
        eta_out[BN + p] = s_e2[k];
    }
}

// ---------------------------------------------------------------------------------------------------
// outer loop: L2error (src/tvl1occflow.cpp:70-88) and the while test of :277.  Fixed-order fp64 sums:
// per-CTA partials, then one thread per triple adds them in index order (the reference adds the pixels
// in index order; the sums agree to rounding and only gate the loop).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kErrThreads) k_occ_error_partial(const TripleCtl *__restrict__ ctl,
                                                               const double *__restrict__ U, double *__restrict__ Uprev,
                                                               double *__restrict__ partials, int N, int B, int parts)
{
    const int b = blockIdx.y;
    if (!ctl[b].active) return;
    const size_t BN = (size_t) B * N;
    const int per = (N + parts - 1) / parts;
    const int lo = blockIdx.x * per, hi = min(N, lo + per);
    double s = 0;
    for (int i = lo + threadIdx.x; i < hi; i += kErrThreads) {
        const size_t p = (size_t) b * N + i;
        const double a = U[p], c = U[BN + p];
        const double da = a - Uprev[p], dc = c - Uprev[BN + p];
        s += da * da + dc * dc;
        Uprev[p] = a;
        Uprev[BN + p] = c;
    }
    __shared__ double sh[kErrThreads];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = kErrThreads / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[(size_t) b * parts + blockIdx.x] = sh[0];
}

__global__ void k_occ_error_decide(TripleCtl *ctl, const double *__restrict__ partials, int N, int B, int parts,
                               double epsilon, int max_outer, int *n_active)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !ctl[b].active) return;
    double s = 0;
    for (int k = 0; k < parts; k++) s += partials[(size_t) b * parts + k];
    const double error = s / N;
    const int n = ctl[b].n + 1;
    ctl[b].n = n;
    ctl[b].error = error;
    if (!(error > epsilon && n < max_outer)) ctl[b].active = 0;
    else atomicAdd(n_active, 1);
}

// start of a warp step: n = 0, error = INFINITY (:275-276); end: counts out (:305-309)
__global__ void k_occ_ctl_begin(TripleCtl *ctl, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ctl[b].active = 1;
    ctl[b].n = 0;
    ctl[b].error = __longlong_as_double(0x7ff0000000000000ll);
}

__global__ void k_occ_ctl_end(const TripleCtl *__restrict__ ctl, int B, int *iters, double *errs, int slot, int stride)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    iters[(size_t) b * stride + slot] = ctl[b].n;
    errs[(size_t) b * stride + slot] = ctl[b].error;
}

// chi = (chi > THR_CHI), src/tvl1occflow.cpp:459 and :46-57
__global__ void __launch_bounds__(256) k_occ_threshold(double *__restrict__ chi, size_t count, double thr)
{
    const size_t p = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (p < count) chi[p] = (chi[p] > thr) ? 1.0 : 0.0;
}

} // namespace occ
