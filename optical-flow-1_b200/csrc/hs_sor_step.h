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
// Storage: the five planes the sweep touches (u, v, I2wx, I2wy, rho_c) are kept in a *wave layout*,
//   element (i, j)  at  W[((j + 2i) mod nx) * ny + i],
// so that the pixels of one time step are contiguous in i: every access of a step is coalesced.
// A ring of S = 8 + P wave columns of u and v lives in shared memory (columns t-3 .. t+4+P at step t;
// column a is fetched with cp.async at step a-4-P, P steps before its first use), and a ring of P + 2
// coefficient columns.  Border pixels are rare and go through global memory.
//
// This header is compiled twice: by nvcc into the kernel, and by g++ into the schedule emulator of
// the tests (HS_SOR_EMULATE).  Arithmetic uses explicitly rounded fp32 operations in both, so the two
// agree bit for bit.
#pragma once

#if defined(__CUDACC__) && !defined(HS_SOR_EMULATE)
#define HS_FN __device__ __forceinline__
#define HS_FN_OUTLINE __device__ __noinline__
#define hs_fma(a, b, c) __fmaf_rn((a), (b), (c))
#define hs_mul(a, b) __fmul_rn((a), (b))
#define hs_add(a, b) __fadd_rn((a), (b))
#define hs_sub(a, b) __fsub_rn((a), (b))
#define hs_div(a, b) __fdiv_rn((a), (b))
#else
#include <math.h>
#define HS_FN static inline
#define HS_FN_OUTLINE static
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
    float *wu, *wv;                           // flow, wave layout, updated in place
    const float *wix, *wiy, *wrho;            // I2wx, I2wy and rho_c = -(I1 - I2w + I2wx u + I2wy v), wave layout
    int nx, ny;
    float alpha2;
    float *ring_u, *ring_v;                   // [S][rp]
    float *cix, *ciy, *crho;                  // [CD][rp]
    int S, CD, rp, P;                         // S = 8 + P, CD = P + 2, rp >= ny
};

// What one time step needs of the modular arithmetic, computed once per thread and step: make_step
// does the divisions, advance() goes from step t to t + 1 with increments and wrap-arounds only.
struct Step {
    int t;
    int ld_a, ld_m, ld_r;                     // ring fetch of wave column a = t + 4 + P: a mod nx, a mod S
    int cf_t, cf_m, cf_r;                     // coefficient fetch for time t + 1 + P: mod nx, mod CD
    int r0;                                   // (t - 3) mod S: ring slot of wave column t - 3
    int wr_m;                                 // t mod nx: wave column the new values of this step go to
    int c_r;                                  // t mod CD: coefficient ring slot of this step
    int slot[7];                              // ring offsets (floats) of wave columns t-3 .. t+3
    int ld_col, ld_slot, cf_col, cf_slot, wr_col, cslot;   // the same as float offsets
};

HS_FN int pmod(int x, int m) { const int r = x % m; return r < 0 ? r + m : r; }
HS_FN int wave_index(int i, int j, int nx, int ny) { return ((j + 2 * i) % nx) * ny + i; }

HS_FN void finish_step(const SorView &V, Step &s)
{
    for (int k = 0; k < 7; k++) {
        int r = s.r0 + k;
        if (r >= V.S) r -= V.S;
        s.slot[k] = r * V.rp;
    }
    s.ld_col = s.ld_m * V.ny;
    s.ld_slot = s.ld_r * V.rp;
    s.cf_col = s.cf_m * V.ny;
    s.cf_slot = s.cf_r * V.rp;
    s.wr_col = s.wr_m * V.ny;
    s.cslot = s.c_r * V.rp;
}

HS_FN Step make_step(const SorView &V, int t)
{
    Step s;
    s.t = t;
    s.ld_a = t + 4 + V.P;
    s.ld_m = pmod(s.ld_a, V.nx);
    s.ld_r = pmod(s.ld_a, V.S);
    s.cf_t = t + 1 + V.P;
    s.cf_m = pmod(s.cf_t, V.nx);
    s.cf_r = pmod(s.cf_t, V.CD);
    s.r0 = pmod(t - 3, V.S);
    s.wr_m = pmod(t, V.nx);
    s.c_r = pmod(t, V.CD);
    finish_step(V, s);
    return s;
}

HS_FN void advance(const SorView &V, Step &s)
{
    s.t++;
    s.ld_a++;
    s.cf_t++;
    if (++s.ld_m == V.nx) s.ld_m = 0;
    if (++s.ld_r == V.S) s.ld_r = 0;
    if (++s.cf_m == V.nx) s.cf_m = 0;
    if (++s.cf_r == V.CD) s.cf_r = 0;
    if (++s.r0 == V.S) s.r0 = 0;
    if (++s.wr_m == V.nx) s.wr_m = 0;
    if (++s.c_r == V.CD) s.c_r = 0;
    finish_step(V, s);
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
HS_FN_OUTLINE float update_global(const SorView &V, int i, int j)
{
    const int nx = V.nx, ny = V.ny;
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
    float un, vn;
    const float e = sor_px(V.wix[p], V.wiy[p], V.wrho[p], V.alpha2, V.wu[d0], V.wu[d1], V.wu[d2], V.wu[d3],
                           V.wu[a0], V.wu[a1], V.wu[a2], V.wu[a3], V.wv[d0], V.wv[d1], V.wv[d2], V.wv[d3],
                           V.wv[a0], V.wv[a1], V.wv[a2], V.wv[a3], V.wu[p], V.wv[p], &un, &vn);
    V.wu[p] = un;
    V.wv[p] = vn;
    return e;
}

// Fetches of row i at one step.  `Cp::cp4(dst, src)` is a 4-byte asynchronous global -> shared copy
// (cp.async in the kernel, a queued copy in the emulator).
template <class Cp>
HS_FN void issue_row(const SorView &V, const Step &s, int i, Cp &cp)
{
    {
        const int j = s.ld_a - 2 * i;
        if (j >= 0 && j <= V.nx - 1) {
            cp.cp4(V.ring_u + s.ld_slot + i, V.wu + s.ld_col + i);
            cp.cp4(V.ring_v + s.ld_slot + i, V.wv + s.ld_col + i);
        }
    }
    if (i >= 1) {
        const int j = s.cf_t - 2 * i;
        if (j >= 1 && j <= V.nx - 2) {
            cp.cp4(V.cix + s.cf_slot + i, V.wix + s.cf_col + i);
            cp.cp4(V.ciy + s.cf_slot + i, V.wiy + s.cf_col + i);
            cp.cp4(V.crho + s.cf_slot + i, V.wrho + s.cf_col + i);
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
        const float *ru = V.ring_u, *rv = V.ring_v;
        const float u_ul = ru[s.slot[0] + i - 1], u_up = ru[s.slot[1] + i - 1], u_ur = ru[s.slot[2] + i - 1];
        const float v_ul = rv[s.slot[0] + i - 1], v_up = rv[s.slot[1] + i - 1], v_ur = rv[s.slot[2] + i - 1];
        const float u_l = ru[s.slot[2] + i], u_c = ru[s.slot[3] + i], u_r = ru[s.slot[4] + i];
        const float v_l = rv[s.slot[2] + i], v_c = rv[s.slot[3] + i], v_r = rv[s.slot[4] + i];
        float u_dl = u_l, u_d = u_c, u_dr = u_r, v_dl = v_l, v_d = v_c, v_dr = v_r;
        if (i < ny - 1) {
            u_dl = ru[s.slot[4] + i + 1]; u_d = ru[s.slot[5] + i + 1]; u_dr = ru[s.slot[6] + i + 1];
            v_dl = rv[s.slot[4] + i + 1]; v_d = rv[s.slot[5] + i + 1]; v_dr = rv[s.slot[6] + i + 1];
        }
        float un, vn;
        e += (double) sor_px(V.cix[s.cslot + i], V.ciy[s.cslot + i], V.crho[s.cslot + i], V.alpha2,
                             u_ul, u_ur, u_dl, u_dr, u_up, u_l, u_d, u_r,
                             v_ul, v_ur, v_dl, v_dr, v_up, v_l, v_d, v_r, u_c, v_c, &un, &vn);
        V.ring_u[s.slot[3] + i] = un;
        V.ring_v[s.slot[3] + i] = vn;
        V.wu[s.wr_col + i] = un;
        V.wv[s.wr_col + i] = vn;
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
