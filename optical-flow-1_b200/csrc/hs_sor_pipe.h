// Per-thread step functions of the PIPELINED Horn-Schunck SOR kernel (k_hs_sor_pipe, hs_kernels.cuh).
//
// hs_sor_step.h runs one sweep per time loop: row i is busy for nx of the 2*ny + nx steps.  Here row i
// processes pixel (i, j) of sweep n at global time
//     T = n*L + 2*i + j,      L = max(nx + 2, 20),
// so consecutive sweeps overlap and every row works all the time.  The data dependences of the
// sequential sweep still hold -- new neighbours were written 1..3 steps earlier, old ones L-3..L-1 steps
// earlier and are overwritten 1..3 steps later -- so u, v stay one in-place array.  Borders:
//     first row (0, j) at n*L + j + 4, UL corner at n*L + 8, UR corner at n*L + nx + 4      (thread of row 0)
//     first column (i, 0) at n*L + 2i + 4, last column (i, nx-1) at n*L + 2i + nx + 1      (thread of row i)
//     last row at its natural time, BL corner at n*L + 2*ny + 1, BR corner at n*L + 2*ny + nx - 2.
//
// Stopping rule (src/horn_schunck_pyramidal.cpp:143): the error of sweep n is complete at
// t_done(n) = n*L + max(8, 2*ny + nx - 2), when the upper rows are already one or two sweeps ahead.  The
// values of every K-th sweep are therefore also written to one of two snapshot planes; on a stop after
// sweep n the kernel restores the snapshot of sweep m = K*floor(n/K) and replays n - m sweeps with a known
// count (limit), without error accounting.  A stop by maxiter needs no restore: no row starts a sweep
// beyond the limit.  tests/test_hs_pipeline_proto.py proves the ordering on plain arrays,
// tests/test_hs_schedule.py replays THIS file (rings, asynchronous copies) against the sequential loop.
//
// Wave layout with period L:  element (i, j) at W[((j + 2i) mod L) * ny + i]  -- at time T every row
// touches wave column T mod L.  The shared-memory rings are those of hs_sor_step.h, keyed by T.
#pragma once
#include "hs_sor_step.h"

namespace hs {

struct PipeView {
    F2 *wuv;                                  // flow, wave layout (period L), updated in place
    const F2 *wxy;                            // (I2wx, I2wy)
    const float *wrho;                        // rho_c
    F2 *snap0, *snap1;                        // snapshot planes, same layout
    double *part;                             // [D][rp]: per-row squared-update sums of the sweeps in flight
    double *esum;                             // [rp]: running sum of each row's current sweep (shared memory)
    int nx, ny, L, K, D;
    float alpha2;
    F2 *ring_uv;                              // [S][rp]
    F2 *cxy;                                  // [CD][rp]
    float *crho;                              // [CD][rp]
    int S, CD, rp, P;
    int limit;                                // sweeps 0 .. limit-1 may be started
    int account;                              // 1: speculative phase (errors + snapshots), 0: replay
};

struct PipeStep {
    int T;
    int r0;                                   // (T - 3) mod S
    int c_r;                                  // T mod CD
    int wr_m;                                 // T mod L
    int ld_m;                                 // (T + 4 + P) mod L
    int cf_m;                                 // (T + 1 + P) mod L
};

HS_HD int pipe_period(int nx) { return nx + 2 > 20 ? nx + 2 : 20; }
HS_FN int pipe_wave_index(int i, int j, int L, int ny) { return ((j + 2 * i) % L) * ny + i; }
// completion time of sweep n (the BR corner; on 3-row images the UL corner)
HS_HD int pipe_t_done(int n, int L, int nx, int ny)
{
    const int tail = 2 * ny + nx - 2;
    return n * L + (tail > 8 ? tail : 8);
}
// sweeps the snapshots are apart / error rows kept: rows run (2ny + nx)/L sweeps ahead of the decision
HS_HD int pipe_snapshot_period(int want, int L, int nx, int ny) { const int m = (2 * ny + nx) / L + 2; return want > m ? want : m; }
HS_HD int pipe_error_depth(int L, int nx, int ny) { return (2 * ny + nx) / L + 3; }

// Where row i stands at time T: x = T - 2i = n*L + j (floor division; n < 0: the row has not started).
// A thread keeps the position of its first row and derives the others and the fetch positions by adding /
// subtracting: no division in the time loop.
struct RowPos { int n, j; };
HS_FN RowPos pipe_pos(int x, int L)
{
    RowPos p;
    p.n = x >= 0 ? x / L : -((-x + L - 1) / L);
    p.j = x - p.n * L;
    return p;
}
HS_FN RowPos pipe_pos_add(RowPos p, int d, int L)              // 0 <= d < L
{
    p.j += d;
    if (p.j >= L) { p.j -= L; p.n++; }
    return p;
}
HS_FN RowPos pipe_pos_sub(RowPos p, int dn, int dj, int L)     // minus dn*L + dj, 0 <= dj < L
{
    p.j -= dj;
    p.n -= dn;
    if (p.j < 0) { p.j += L; p.n--; }
    return p;
}

HS_FN PipeStep pipe_make_step(const PipeView &V, int T)
{
    PipeStep s;
    s.T = T;
    s.r0 = pmod(T - 3, V.S);
    s.c_r = pmod(T, V.CD);
    s.wr_m = pmod(T, V.L);
    s.ld_m = pmod(T + 4 + V.P, V.L);
    s.cf_m = pmod(T + 1 + V.P, V.L);
    return s;
}

HS_FN void pipe_advance(const PipeView &V, PipeStep &s)
{
    s.T++;
    s.r0 = wrap(s.r0 + 1, V.S);
    s.c_r = wrap(s.c_r + 1, V.CD);
    s.wr_m = wrap(s.wr_m + 1, V.L);
    s.ld_m = wrap(s.ld_m + 1, V.L);
    s.cf_m = wrap(s.cf_m + 1, V.L);
}

// Border pixel (i, j) of sweep n through global memory (see update_global_px of hs_sor_step.h).
HS_FN_OUTLINE float pipe_update_global_px(F2 *wuv, const F2 *wxy, const float *wrho, F2 *snap, int nx, int ny, int L,
                                          float alpha2, int i, int j)
{
    const int im = i > 0 ? i - 1 : 0, ip = i < ny - 1 ? i + 1 : ny - 1;
    const int jm = j > 0 ? j - 1 : 0, jp = j < nx - 1 ? j + 1 : nx - 1;
    int d0 = pipe_wave_index(im, jm, L, ny), d1 = pipe_wave_index(im, jp, L, ny);
    int d2 = pipe_wave_index(ip, jm, L, ny), d3 = pipe_wave_index(ip, jp, L, ny);
    const int a0 = pipe_wave_index(im, j, L, ny), a1 = pipe_wave_index(i, jm, L, ny);
    const int a2 = pipe_wave_index(ip, j, L, ny), a3 = pipe_wave_index(i, jp, L, ny);
    const int p = pipe_wave_index(i, j, L, ny);
    if (i == ny - 1 && j == nx - 1) {
        d0 = a1; d1 = p; d2 = pipe_wave_index(im, jm, L, ny); d3 = a0;
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

// snapshot plane the values of sweep n (0-based) go to, or null
HS_FN F2 *pipe_snap_of(const PipeView &V, int n)
{
    if (!V.account || (n + 1) % V.K != 0) return 0;
    return (((n + 1) / V.K) & 1) ? V.snap1 : V.snap0;
}

HS_FN void pipe_border(const PipeView &V, int i, int j, int n)
{
    const float e = pipe_update_global_px(V.wuv, V.wxy, V.wrho, pipe_snap_of(V, n), V.nx, V.ny, V.L, V.alpha2, i, j);
    if (V.account) V.esum[i] += (double) e;
}

// the row finished sweep n: hand its error sum to the decision
HS_FN void pipe_deposit(const PipeView &V, int i, int n)
{
    if (V.account) {
        V.part[(n % V.D) * V.rp + i] = V.esum[i];
        V.esum[i] = 0.0;
    }
}

// pos: position of row i at time s.T
template <class Cp>
HS_FN void pipe_issue_row(const PipeView &V, const PipeStep &s, int i, RowPos pos, Cp &cp)
{
    {
        const RowPos f = pipe_pos_add(pos, 4 + V.P, V.L);       // where row i stands at time T + 4 + P
        if (f.n >= 0 && f.j <= V.nx - 1)
            cp.cp8(V.ring_uv + wrap(s.r0 + V.S - 1, V.S) * V.rp + i, V.wuv + s.ld_m * V.ny + i);
    }
    if (i >= 1) {
        const RowPos f = pipe_pos_add(pos, 1 + V.P, V.L);
        if (f.n >= 0 && f.j >= 1 && f.j <= V.nx - 2) {
            const int slot = wrap(s.c_r + V.CD - 1, V.CD) * V.rp + i, col = s.cf_m * V.ny + i;
            cp.cp8(V.cxy + slot, V.wxy + col);
            cp.cp4(V.crho + slot, V.wrho + col);
        }
    }
}

HS_FN void pipe_compute_row(const PipeView &V, const PipeStep &s, int i, RowPos pos)
{
    const int nx = V.nx, ny = V.ny;
    if (i == 0) {
        // first row (0, j) at n*L + j + 4, then the upper corners
        const RowPos q = pipe_pos_sub(pos, 0, 4, V.L);
        const int n = q.n, j = q.j;
        if (n < 0 || n >= V.limit) return;
        if (j >= 1 && j <= nx - 2) pipe_border(V, 0, j, n);
        if (j == 4) pipe_border(V, 0, 0, n);
        if (j == nx) pipe_border(V, 0, nx - 1, n);
        if (j == (nx > 4 ? nx : 4)) pipe_deposit(V, 0, n);      // after both corners (nx = 3: UR comes first)
        return;
    }
    const int n = pos.n, j = pos.j;
    if (n < 0 || n >= V.limit) return;
    if (j >= 1 && j <= nx - 2) {
        // interior row, or the last row (its lower neighbours clamp onto the row itself)
        const F2 *ring = V.ring_uv + i;
        int so[7];
        for (int k = 0; k < 7; k++) so[k] = wrap(s.r0 + k, V.S) * V.rp;
        const F2 ul = ring[so[0] - 1], up = ring[so[1] - 1], ur = ring[so[2] - 1];
        const F2 l = ring[so[2]], c = ring[so[3]], r = ring[so[4]];
        F2 dl = l, d = c, dr = r;
        if (i < ny - 1) { dl = ring[so[4] + 1]; d = ring[so[5] + 1]; dr = ring[so[6] + 1]; }
        const int cslot = s.c_r * V.rp + i;
        const F2 g = V.cxy[cslot];
        F2 nw;
        const float e = sor_px(g.x, g.y, V.crho[cslot], V.alpha2, ul.x, ur.x, dl.x, dr.x, up.x, l.x, d.x, r.x,
                               ul.y, ur.y, dl.y, dr.y, up.y, l.y, d.y, r.y, c.x, c.y, &nw.x, &nw.y);
        V.ring_uv[so[3] + i] = nw;
        const int p = s.wr_m * ny + i;
        V.wuv[p] = nw;
        if (V.account) {
            V.esum[i] += (double) e;
            F2 *snap = pipe_snap_of(V, n);
            if (snap) snap[p] = nw;
        }
    }
    if (i <= ny - 2) {
        if (j == 4) pipe_border(V, i, 0, n);
        if (j == nx + 1) { pipe_border(V, i, nx - 1, n); pipe_deposit(V, i, n); }
    } else {
        if (j == 3) pipe_border(V, i, 0, n);
        if (j == nx) { pipe_border(V, i, nx - 1, n); pipe_deposit(V, i, n); }
    }
}

} // namespace hs
