// Per-thread step functions of the Horn-Schunck SOR kernel (k_hs_sor, hs_kernels.cuh).
//
// The reference's SOR sweep (src/horn_schunck_pyramidal.cpp:144-230) is a lexicographic Gauss-Seidel
// pass: pixel (i, j) reads the NEW values of its up-left, up, up-right and left neighbours and the OLD
// values of the others, then the borders are swept (first / last row, first / last column, corners),
// each reading what the sweeps before it left behind.  An identical result needs the same data
// dependences, not the same loop: this file defines a *wavefront schedule* that honours every one of
// them, so that all rows of an image advance together.
//
//   time step t, row i  ->  interior / last-row pixel (i, j = t - 2i)        when 1 <= j <= nx-2
//                           first-column pixel (i, 0)                        when j == 4
//                           last-column pixel  (i, nx-1)                     when j == nx+1
//   row 0               ->  first-row pixel (0, t - 4)
//   after the last step ->  the four corners, in the reference's order
//
// Every value a pixel reads was written at a strictly earlier step when the reference reads the new
// value, and is overwritten at a strictly later step when the reference reads the old one (the
// derivation is in DESIGN.md section 10; tests/test_hs_schedule.py replays this very file on the CPU in
// adversarial thread orders and copy-landing times and compares it bit for bit with the sequential
// sweep).
//
// Storage: the five planes the sweep touches are kept in a *wave layout*,
//   element (i, j)  at  W[((j + 2i) mod nx) * ny + i],
// so that the pixels of one time step are contiguous in i: every access of a step is coalesced.
// u and v are interleaved (one 8-byte element per pixel), so are I2wx and I2wy; rho_c is a plane of its
// own: 28 bytes per pixel and sweep, moved by three asynchronous copies, one 8-byte store and nine
// 8-byte shared-memory loads.
// A ring of S = 8 + P wave columns of u and v lives in shared memory (columns t-3 .. t+4+P at step t;
// column a is fetched with cp.async at step a-4-P and is complete P steps later, one step before its
// first use), and a ring of P + 2 coefficient columns.  Border pixels are rare and go through global memory.
//
// This header is compiled twice: by nvcc into the kernel, and by g++ into the schedule emulator of
// the tests (HS_SOR_EMULATE).  Arithmetic uses explicitly rounded fp32 operations in both, so the two
// agree bit for bit.
#pragma once

#if defined(__CUDACC__) && !defined(HS_SOR_EMULATE)
#define HS_FN __device__ __forceinline__
#define HS_FN_OUTLINE __device__ __noinline__
#define HS_HD __host__ __device__ __forceinline__
#define hs_fma(a, b, c) __fmaf_rn((a), (b), (c))
#define hs_mul(a, b) __fmul_rn((a), (b))
#define hs_add(a, b) __fadd_rn((a), (b))
#define hs_sub(a, b) __fsub_rn((a), (b))
#define hs_div(a, b) __fdiv_rn((a), (b))
namespace hs { typedef float2 F2; }
#else
#include <math.h>
namespace hs { struct F2 { float x, y; }; }
#define HS_FN static inline
#define HS_FN_OUTLINE static
#define HS_HD static inline
#define hs_fma(a, b, c) fmaf((a), (b), (c))
#define hs_mul(a, b) ((float) ((float) (a) * (float) (b)))
#define hs_add(a, b) ((float) ((float) (a) + (float) (b)))
#define hs_sub(a, b) ((float) ((float) (a) - (float) (b)))
#define hs_div(a, b) ((float) ((float) (a) / (float) (b)))
#endif

namespace hs {

constexpr float kSorW = 1.9f;                 // src/horn_schunck_pyramidal.cpp:21
constexpr float kOneMinusW = (float) (1.0 - 1.9);
constexpr float kTwelfth = (float) (1.0 / 12.0);
constexpr float kSixth = (float) (1.0 / 6.0);
constexpr int kRingBase = 8;                  // wave columns t-3 .. t+4 are live at step t
constexpr int kMaxPrefetch = 3;

struct SorView {
    F2 *wuv;                                  // flow (u, v), wave layout, updated in place
    const F2 *wxy;                            // (I2wx, I2wy), wave layout
    const float *wrho;                        // rho_c = -(I1 - I2w + I2wx u + I2wy v), wave layout
    int nx, ny;
    float alpha2;
    F2 *ring_uv;                              // [S][rp]
    F2 *cxy;                                  // [CD][rp]
    float *crho;                              // [CD][rp]
    int S, CD, rp, P;                         // S = 8 + P, CD = P + 2, rp >= ny
};

// The modular counters of a time step.  Every thread keeps its own copy in registers: make_step does
// the divisions once per sweep, advance() goes from step t to t + 1 with increments and wrap-arounds.
struct Step {
    int t;
    int r0;                                   // (t - 3) mod S: ring slot of wave column t - 3
    int c_r;                                  // t mod CD: coefficient ring slot of this step
    int wr_m;                                 // t mod nx: wave column the new values of this step go to
    int ld_m;                                 // (t + 4 + P) mod nx: wave column fetched into the ring at this step
    int cf_m;                                 // (t + 1 + P) mod nx: wave column of the coefficients fetched
};

HS_FN int pmod(int x, int m) { const int r = x % m; return r < 0 ? r + m : r; }
HS_FN int wrap(int x, int m) { return x >= m ? x - m : x; }      // for 0 <= x < 2m
HS_FN int wave_index(int i, int j, int nx, int ny) { return ((j + 2 * i) % nx) * ny + i; }

HS_FN Step make_step(const SorView &V, int t)
{
    Step s;
    s.t = t;
    s.r0 = pmod(t - 3, V.S);
    s.c_r = pmod(t, V.CD);
    s.wr_m = pmod(t, V.nx);
    s.ld_m = pmod(t + 4 + V.P, V.nx);
    s.cf_m = pmod(t + 1 + V.P, V.nx);
    return s;
}

HS_FN void advance(const SorView &V, Step &s)
{
    s.t++;
    s.r0 = wrap(s.r0 + 1, V.S);
    s.c_r = wrap(s.c_r + 1, V.CD);
    s.wr_m = wrap(s.wr_m + 1, V.nx);
    s.ld_m = wrap(s.ld_m + 1, V.nx);
    s.cf_m = wrap(s.cf_m + 1, V.nx);
}

// One SOR update, src/horn_schunck_pyramidal.cpp:31-71.  d* = diagonal neighbours in the order the
// call site lists them (p1..p4), a* = axial ones (p5..p8).  The system's constant parts (:127-137)
// are formed from the three stored planes: Au = dif*Ix, Du = Ix^2 + alpha^2, D = Ix*Iy, dif = -rho.
HS_FN float sor_px(float ix, float iy, float rho, float alpha2, float ud0, float ud1, float ud2, float ud3,
                   float ua0, float ua1, float ua2, float ua3, float vd0, float vd1, float vd2, float vd3,
                   float va0, float va1, float va2, float va3, float uk, float vk, float *un_out,
                   float *vn_out)
{
    const float ula = hs_fma(kSixth, hs_add(hs_add(hs_add(ua0, ua1), ua2), ua3),
                             hs_mul(kTwelfth, hs_add(hs_add(hs_add(ud0, ud1), ud2), ud3)));
    const float vla = hs_fma(kSixth, hs_add(hs_add(hs_add(va0, va1), va2), va3),
                             hs_mul(kTwelfth, hs_add(hs_add(hs_add(vd0, vd1), vd2), vd3)));
    const float dif = -rho;
    const float Au = hs_mul(dif, ix), Av = hs_mul(dif, iy);
    const float Du = hs_fma(ix, ix, alpha2), Dv = hs_fma(iy, iy, alpha2);
    const float D = hs_mul(ix, iy);
    const float un = hs_fma(kSorW, hs_div(hs_fma(alpha2, ula, hs_fma(-D, vk, Au)), Du), hs_mul(kOneMinusW, uk));
    const float vn = hs_fma(kSorW, hs_div(hs_fma(alpha2, vla, hs_fma(-D, un, Av)), Dv), hs_mul(kOneMinusW, vk));
    *un_out = un;
    *vn_out = vn;
    const float du = hs_sub(un, uk), dv = hs_sub(vn, vk);
    return hs_fma(du, du, hs_mul(dv, dv));
}

// Border pixel (i, j) through global memory: index-clamped 8-neighbourhood, which is what every
// border call site of :160-228 passes -- except the bottom-right corner (:223-228), whose diagonal
// arguments come in the order (left, self, up-left, up); floating-point sums follow that order.
// (Arguments by value: the caller's SorView stays in registers.)
HS_FN_OUTLINE float update_global_px(F2 *wuv, const F2 *wxy, const float *wrho, int nx, int ny, float alpha2,
                                     int i, int j)
{
    const int im = i > 0 ? i - 1 : 0, ip = i < ny - 1 ? i + 1 : ny - 1;
    const int jm = j > 0 ? j - 1 : 0, jp = j < nx - 1 ? j + 1 : nx - 1;
    int d0 = wave_index(im, jm, nx, ny), d1 = wave_index(im, jp, nx, ny);
    int d2 = wave_index(ip, jm, nx, ny), d3 = wave_index(ip, jp, nx, ny);
    const int a0 = wave_index(im, j, nx, ny), a1 = wave_index(i, jm, nx, ny);
    const int a2 = wave_index(ip, j, nx, ny), a3 = wave_index(i, jp, nx, ny);
    const int p = wave_index(i, j, nx, ny);
    if (i == ny - 1 && j == nx - 1) {
        d0 = a1; d1 = p; d2 = wave_index(im, jm, nx, ny); d3 = a0;
    }
    const F2 D0 = wuv[d0], D1 = wuv[d1], D2 = wuv[d2], D3 = wuv[d3];
    const F2 A0 = wuv[a0], A1 = wuv[a1], A2 = wuv[a2], A3 = wuv[a3];
    const F2 c = wuv[p], g = wxy[p];
    F2 n;
    const float e = sor_px(g.x, g.y, wrho[p], alpha2, D0.x, D1.x, D2.x, D3.x, A0.x, A1.x, A2.x, A3.x,
                           D0.y, D1.y, D2.y, D3.y, A0.y, A1.y, A2.y, A3.y, c.x, c.y, &n.x, &n.y);
    wuv[p] = n;
    return e;
}

HS_FN float update_global(const SorView &V, int i, int j)
{
    return update_global_px(V.wuv, V.wxy, V.wrho, V.nx, V.ny, V.alpha2, i, j);
}

// Fetches of row i at one step.  `Cp::cp8(dst, src)` / `Cp::cp4(dst, src)` are 8- and 4-byte asynchronous
// global -> shared copies (cp.async in the kernel, queued copies in the emulator).
template <class Cp>
HS_FN void issue_row(const SorView &V, const Step &s, int i, Cp &cp)
{
    {
        // wave column a = t + 4 + P goes to the ring slot of column t - 4 (S = 8 + P), whose last reader
        // was step t - 1
        const int j = s.t + 4 + V.P - 2 * i;
        if (j >= 0 && j <= V.nx - 1) {
            cp.cp8(V.ring_uv + wrap(s.r0 + V.S - 1, V.S) * V.rp + i, V.wuv + s.ld_m * V.ny + i);
        }
    }
    if (i >= 1) {
        // coefficients of time t + 1 + P go to the slot of time t - 1 (CD = P + 2)
        const int j = s.t + 1 + V.P - 2 * i;
        if (j >= 1 && j <= V.nx - 2) {
            const int slot = wrap(s.c_r + V.CD - 1, V.CD) * V.rp + i, col = s.cf_m * V.ny + i;
            cp.cp8(V.cxy + slot, V.wxy + col);
            cp.cp4(V.crho + slot, V.wrho + col);
        }
    }
}

// All updates of row i at step s.t (>= 3); returns their contribution to the squared-update sum.
HS_FN double compute_row(const SorView &V, const Step &s, int i)
{
    const int nx = V.nx, ny = V.ny;
    const int j = s.t - 2 * i;
    double e = 0.0;
    if (i >= 1 && j >= 1 && j <= nx - 2) {
        // interior row, or the last row (its lower neighbours clamp onto the row itself)
        const F2 *ring = V.ring_uv + i;
        int so[7];                                // ring offsets of wave columns t-3 .. t+3
        for (int k = 0; k < 7; k++) so[k] = wrap(s.r0 + k, V.S) * V.rp;
        const F2 ul = ring[so[0] - 1], up = ring[so[1] - 1], ur = ring[so[2] - 1];
        const F2 l = ring[so[2]], c = ring[so[3]], r = ring[so[4]];
        F2 dl = l, d = c, dr = r;
        if (i < ny - 1) { dl = ring[so[4] + 1]; d = ring[so[5] + 1]; dr = ring[so[6] + 1]; }
        const int cslot = s.c_r * V.rp + i;
        const F2 g = V.cxy[cslot];
        F2 n;
        e += (double) sor_px(g.x, g.y, V.crho[cslot], V.alpha2, ul.x, ur.x, dl.x, dr.x, up.x, l.x, d.x, r.x,
                             ul.y, ur.y, dl.y, dr.y, up.y, l.y, d.y, r.y, c.x, c.y, &n.x, &n.y);
        V.ring_uv[so[3] + i] = n;
        V.wuv[s.wr_m * V.ny + i] = n;
    }
    if (i >= 1 && i <= ny - 2) {
        if (j == 4) e += (double) update_global(V, i, 0);
        if (j == nx + 1) e += (double) update_global(V, i, nx - 1);
    } else if (i == 0) {
        const int jj = s.t - 4;
        if (jj >= 1 && jj <= nx - 2) e += (double) update_global(V, 0, jj);
    }
    return e;
}

// After the last step: the corners in the reference's order (:198-228).
HS_FN double corners(const SorView &V)
{
    double e = 0.0;
    e += (double) update_global(V, 0, 0);
    e += (double) update_global(V, 0, V.nx - 1);
    e += (double) update_global(V, V.ny - 1, 0);
    e += (double) update_global(V, V.ny - 1, V.nx - 1);
    return e;
}

HS_FN int first_step(const SorView &V) { return -4 - V.P; }
HS_FN int last_step(const SorView &V) { return 2 * V.ny + V.nx - 3; }

} // namespace hs
