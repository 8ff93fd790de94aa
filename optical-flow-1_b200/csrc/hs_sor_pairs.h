// Per-thread step functions of the next Horn-Schunck SOR kernel: pipelined sweeps (hs_sor_pipe.h) with TWO
// columns per thread-step.  NOT YET A KERNEL: this header is verified by the CPU replay of
// tests/test_hs_schedule.py (rings, asynchronous copies, adversarial orders) and waits for its CUDA
// wrapper and a GPU run ("Next for this kernel").
//
// Row i processes the column pair c = (2c, 2c+1) of sweep n at global time
//     T = n*L + 2*i + c,      L = max(cl + 5, 16),   cl = (nx-2)/2 the pair holding the last interior column,
// left pixel first.  The row skew stays two steps: (i, 2c+1) needs the new (i-1, 2c+2), which row i-1 wrote
// one step earlier.  Borders:
//     first row pair c at n*L + c + 4; UL corner at n*L + 6, UR corner at n*L + cl + 6          (thread of row 0)
//     first column (i, 0) at n*L + 2i + 3, last column (i, nx-1) at n*L + 2i + cl + 3            (thread of row i)
//     last row at its natural time; BL corner at n*L + 2*ny, BR corner at n*L + 2*ny + cl        (thread ny-1)
// Stopping rule, snapshots and replay as in hs_sor_pipe.h; sweep n is complete at
// t_done(n) = n*L + max(cl + 6, 2*ny + cl).
//
// Storage: pair wave layout, pair (i, c) at W[((c + 2i) mod L) * ny + i]:
//     wuv  16-byte elements (u, v of column 2c; u, v of column 2c+1)
//     wxy  16-byte elements (I2wx, I2wy of both columns),  wrho 8-byte elements (rho_c of both columns)
// still 28 bytes per pixel and sweep.  Per pair: 9 + 2 shared-memory loads of 16 / 8 bytes instead of
// 2 x (9 + 2) of 8 / 4 bytes, one 16-byte store to the ring and one to HBM, three asynchronous copies, one
// barrier, one set of index arithmetic.
#pragma once
#include "hs_sor_pipe.h"

#if defined(__CUDACC__) && !defined(HS_SOR_EMULATE)
namespace hs { typedef float4 F4; }
#else
namespace hs { struct F4 { float x, y, z, w; }; }
#endif

namespace hs {

struct PairsView {
    F4 *wuv;                                  // flow pairs, pair wave layout (period L), updated in place
    const F4 *wxy;                            // gradient pairs
    const F2 *wrho;                           // rho_c pairs
    F4 *snap0, *snap1;                        // snapshot planes, same layout
    double *part;                             // [D][rp] per-row squared-update sums of the sweeps in flight
    double *esum;                             // [rp] running sum of each row's current sweep (shared memory)
    int nx, ny, L, K, D;
    int cl, cn;                               // (nx-2)/2: last pair with an interior column; (nx+1)/2: pairs per row
    float alpha2;
    F4 *ring_uv;                              // [S][rp]
    F4 *cxy;                                  // [CD][rp]
    F2 *crho;                                 // [CD][rp]
    int S, CD, rp, P;
    int limit, account;
};

HS_HD int pairs_cl(int nx) { return (nx - 2) / 2; }
HS_HD int pairs_cn(int nx) { return (nx + 1) / 2; }
HS_HD int pairs_period(int nx) { const int l = pairs_cl(nx) + 5; return l > 16 ? l : 16; }
HS_HD int pairs_t_done(int n, int L, int nx, int ny)
{
    const int cl = pairs_cl(nx), a = cl + 6, b = 2 * ny + cl;
    return n * L + (a > b ? a : b);
}
HS_HD int pairs_snapshot_period(int want, int L, int nx, int ny) { const int m = (2 * ny + nx) / L + 2; return want > m ? want : m; }
HS_HD int pairs_error_depth(int L, int nx, int ny) { return (2 * ny + nx) / L + 3; }

// index of pixel (i, j) when a pair plane is viewed as an array of single-pixel elements
HS_FN int pairs_px_index(int i, int j, int L, int ny) { return 2 * ((((j >> 1) + 2 * i) % L) * ny + i) + (j & 1); }

HS_FN F4 *pairs_snap_of(const PairsView &V, int n)
{
    if (!V.account || (n + 1) % V.K != 0) return 0;
    return (((n + 1) / V.K) & 1) ? V.snap1 : V.snap0;
}

// Border pixel (i, j) of sweep n through global memory; the planes viewed as single-pixel arrays.
HS_FN_OUTLINE float pairs_update_global_px(F2 *wuv, const F2 *wxy, const float *wrho, F2 *snap, int nx, int ny, int L,
                                           float alpha2, int i, int j)
{
    const int im = i > 0 ? i - 1 : 0, ip = i < ny - 1 ? i + 1 : ny - 1;
    const int jm = j > 0 ? j - 1 : 0, jp = j < nx - 1 ? j + 1 : nx - 1;
    int d0 = pairs_px_index(im, jm, L, ny), d1 = pairs_px_index(im, jp, L, ny);
    int d2 = pairs_px_index(ip, jm, L, ny), d3 = pairs_px_index(ip, jp, L, ny);
    const int a0 = pairs_px_index(im, j, L, ny), a1 = pairs_px_index(i, jm, L, ny);
    const int a2 = pairs_px_index(ip, j, L, ny), a3 = pairs_px_index(i, jp, L, ny);
    const int p = pairs_px_index(i, j, L, ny);
    if (i == ny - 1 && j == nx - 1) {
        d0 = a1; d1 = p; d2 = pairs_px_index(im, jm, L, ny); d3 = a0;
    }
    const F2 D0 = wuv[d0], D1 = wuv[d1], D2 = wuv[d2], D3 = wuv[d3];
    const F2 A0 = wuv[a0], A1 = wuv[a1], A2 = wuv[a2], A3 = wuv[a3];
    const F2 c = wuv[p], g = wxy[p];
    F2 n;
    const float e = sor_px(g.x, g.y, wrho[p], alpha2, D0.x, D1.x, D2.x, D3.x, A0.x, A1.x, A2.x, A3.x,
                           D0.y, D1.y, D2.y, D3.y, A0.y, A1.y, A2.y, A3.y, c.x, c.y, &n.x, &n.y);
    wuv[p] = n;
    if (snap) snap[p] = n;
    return e;
}

HS_FN void pairs_border(const PairsView &V, int i, int j, int n)
{
    const float e = pairs_update_global_px((F2 *) V.wuv, (const F2 *) V.wxy, (const float *) V.wrho,
                                           (F2 *) pairs_snap_of(V, n), V.nx, V.ny, V.L, V.alpha2, i, j);
    if (V.account) V.esum[i] += (double) e;
}

HS_FN void pairs_deposit(const PairsView &V, int i, int n)
{
    if (V.account) {
        V.part[(n % V.D) * V.rp + i] = V.esum[i];
        V.esum[i] = 0.0;
    }
}

// pos: position (n, c) of row i at time s.T (x = T - 2i = n*L + c).  Cp: cp16 / cp8 asynchronous copies.
template <class Cp>
HS_FN void pairs_issue_row(const PairsView &V, const PipeStep &s, int i, RowPos pos, Cp &cp)
{
    {
        const RowPos f = pipe_pos_add(pos, 4 + V.P, V.L);
        if (f.n >= 0 && f.j <= V.cn - 1)
            cp.cp16(V.ring_uv + wrap(s.r0 + V.S - 1, V.S) * V.rp + i, V.wuv + s.ld_m * V.ny + i);
    }
    if (i >= 1) {
        const RowPos f = pipe_pos_add(pos, 1 + V.P, V.L);
        if (f.n >= 0 && f.j <= V.cl) {
            const int slot = wrap(s.c_r + V.CD - 1, V.CD) * V.rp + i, col = s.cf_m * V.ny + i;
            cp.cp16(V.cxy + slot, V.wxy + col);
            cp.cp8(V.crho + slot, V.wrho + col);
        }
    }
}

HS_FN void pairs_compute_row(const PairsView &V, const PipeStep &s, int i, RowPos pos)
{
    const int nx = V.nx, ny = V.ny;
    if (i == 0) {
        // first row pair c at n*L + c + 4 (through global memory), then the upper corners
        const RowPos q = pipe_pos_sub(pos, 0, 4, V.L);
        const int n = q.n, c = q.j;
        if (n < 0 || n >= V.limit) return;
        if (c <= V.cl) {
            if (2 * c >= 1 && 2 * c <= nx - 2) pairs_border(V, 0, 2 * c, n);
            if (2 * c + 1 <= nx - 2) pairs_border(V, 0, 2 * c + 1, n);
        }
        if (c == 2) pairs_border(V, 0, 0, n);
        if (c == V.cl + 2) { pairs_border(V, 0, nx - 1, n); pairs_deposit(V, 0, n); }
        return;
    }
    const int n = pos.n, c = pos.j;
    if (n < 0 || n >= V.limit) return;
    if (c <= V.cl) {
        const F4 *ring = V.ring_uv + i;
        int so[7];
        for (int k = 0; k < 7; k++) so[k] = wrap(s.r0 + k, V.S) * V.rp;
        const F4 A = ring[so[0] - 1], B = ring[so[1] - 1], Cc = ring[so[2] - 1];      // row above: new
        const F4 Lf = ring[so[2]], S = ring[so[3]], R = ring[so[4]];                  // this row: new | old | old
        F4 D1 = S, D2 = S, D3 = S;                                                     // row below: old
        const bool last = i == ny - 1;
        if (!last) { D1 = ring[so[4] + 1]; D2 = ring[so[5] + 1]; D3 = ring[so[6] + 1]; }
        const int cslot = s.c_r * V.rp + i;
        const F4 g = V.cxy[cslot];
        const F2 rh = V.crho[cslot];
        F4 nw = S;
        double e = 0.0;
        const int j0 = 2 * c;
        if (j0 >= 1) {          // (j0 <= nx-2 holds for every c <= cl)
            // last row: its lower neighbours clamp onto the row itself (left: new, self and right: old)
            const float dlx = last ? Lf.z : D1.z, dly = last ? Lf.w : D1.w;
            const float dx = last ? S.x : D2.x, dy = last ? S.y : D2.y;
            const float drx = last ? S.z : D2.z, dry = last ? S.w : D2.w;
            e += (double) sor_px(g.x, g.y, rh.x, V.alpha2, A.z, B.z, dlx, drx, B.x, Lf.z, dx, S.z,
                                 A.w, B.w, dly, dry, B.y, Lf.w, dy, S.w, S.x, S.y, &nw.x, &nw.y);
        }
        if (j0 + 1 <= nx - 2) {
            const float lx = j0 >= 1 ? nw.x : S.x, ly = j0 >= 1 ? nw.y : S.y;       // column 0 is a border pixel: old value
            const float dlx = last ? lx : D2.x, dly = last ? ly : D2.y;
            const float dx = last ? S.z : D2.z, dy = last ? S.w : D2.w;
            const float drx = last ? R.x : D3.x, dry = last ? R.y : D3.y;
            e += (double) sor_px(g.z, g.w, rh.y, V.alpha2, B.x, Cc.x, dlx, drx, B.z, lx, dx, R.x,
                                 B.y, Cc.y, dly, dry, B.w, ly, dy, R.y, S.z, S.w, &nw.z, &nw.w);
        }
        V.ring_uv[so[3] + i] = nw;
        const int p = s.wr_m * ny + i;
        V.wuv[p] = nw;
        if (V.account) {
            V.esum[i] += e;
            F4 *snap = pairs_snap_of(V, n);
            if (snap) snap[p] = nw;
        }
    }
    if (i <= ny - 2) {
        if (c == 3) pairs_border(V, i, 0, n);
        if (c == V.cl + 3) { pairs_border(V, i, nx - 1, n); pairs_deposit(V, i, n); }
    } else {
        if (c == 2) pairs_border(V, i, 0, n);
        if (c == V.cl + 2) { pairs_border(V, i, nx - 1, n); pairs_deposit(V, i, n); }
    }
}

} // namespace hs
