// TEST INFRASTRUCTURE / DESIGN PROTOTYPE ("Next for this kernel").
//
// CPU model of the *pipelined* Horn-Schunck SOR schedule that is to replace the one-sweep-at-a-time
// time loop of k_hs_sor: row i processes pixel (i, j) of sweep n at global time
//     T = n*L + 2*i + j,        L = max(nx + 2, 8)   (instead of a period of 2*ny + nx),
// so consecutive sweeps overlap and every row is busy all the time.  Borders ride along:
//     first row   (0, j)        at n*L + j + 4            (thread of row 0)
//     last row    (ny-1, j)     at n*L + 2*(ny-1) + j     (natural)
//     first col   (i, 0)        at n*L + 2*i + 4          (thread of row i)
//     last col    (i, nx-1)     at n*L + 2*i + nx + 1     (thread of row i)
//     corners     UL n*L + 8, UR n*L + nx + 4 (thread 0); BL n*L + 2*ny + 1, BR n*L + 2*ny + nx - 2 (thread ny-1)
// u and v stay ONE in-place array: every read still sees the version the sequential sweep sees (new
// neighbours were written 1..3 steps earlier, old ones L-3..L-1 steps earlier and are overwritten 1..3
// steps later).
//
// The stopping rule `while (error > TOL && niter < maxiter)` needs the error of a complete sweep, which
// is known 2*ny + nx steps after the top rows finished it -- they are a sweep or two ahead by then.
// Exactness is kept by snapshots: every pixel value of the sweeps K, 2K, 3K, ... is also written to one of
// two snapshot buffers (alternating); on a stop after sweep n the state is restored from the snapshot
// of sweep m = K*floor(n/K) (the initial state for m = 0) and n - m sweeps are replayed with a known
// count, which pipelines without speculation.  A stop by maxiter needs nothing: no row ever starts a
// sweep beyond maxiter.
//
// This file proves the schedule against the sequential sweep (hs_emu_seq_sor of hs_schedule_emu.cpp,
// same fp32 arithmetic via hs_sor_step.h) under adversarial thread orders; it accesses plain row-major
// arrays -- the wave layout and the shared-memory rings of the kernel are orthogonal to the ordering
// question and are covered by hs_schedule_emu.cpp.
#define HS_SOR_EMULATE 1
#include "../../optical-flow-1_b200/csrc/hs_sor_step.h"

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t) (s >> 33); }
};

struct Pipe {
    int nx, ny, L, K, D, maxiter;      // D: sweeps whose error partials are kept (rows run up to (2ny+nx)/L sweeps ahead)
    float alpha2;
    const float *ix, *iy, *rho;
    float *u, *v;
    std::vector<float> snap_u[2], snap_v[2];
    std::vector<double> part;          // [D][ny + 1] squared-update partial sums per (sweep mod D, row)
    bool account;                      // speculative phase: keep errors and snapshots
    int limit;                         // sweeps 0 .. limit-1 may be started

    // one update of pixel (i, j) belonging to sweep n (0-based), reference neighbour order
    void update(int i, int j, int n, int part_row)
    {
        const int im = i > 0 ? i - 1 : 0, ip = i < ny - 1 ? i + 1 : ny - 1;
        const int jm = j > 0 ? j - 1 : 0, jp = j < nx - 1 ? j + 1 : nx - 1;
        int d0 = im * nx + jm, d1 = im * nx + jp, d2 = ip * nx + jm, d3 = ip * nx + jp;
        const int a0 = im * nx + j, a1 = i * nx + jm, a2 = ip * nx + j, a3 = i * nx + jp, p = i * nx + j;
        if (i == ny - 1 && j == nx - 1) { d0 = p - 1; d1 = p; d2 = p - nx - 1; d3 = p - nx; }
        float un, vn;
        const float e = hs::sor_px(ix[p], iy[p], rho[p], alpha2, u[d0], u[d1], u[d2], u[d3], u[a0], u[a1], u[a2],
                                   u[a3], v[d0], v[d1], v[d2], v[d3], v[a0], v[a1], v[a2], v[a3], u[p], v[p], &un,
                                   &vn);
        u[p] = un;
        v[p] = vn;
        if (account) {
            part[(size_t) (n % D) * (ny + 1) + part_row] += (double) e;
            if ((n + 1) % K == 0) {                     // sweep n+1 (1-based) is a snapshot sweep
                const int q = ((n + 1) / K) & 1;
                snap_u[q][p] = un;
                snap_v[q][p] = vn;
            }
        }
    }

    // everything the thread of row i does at global time T
    void row_work(int T, int i)
    {
        if (i >= 1 && i <= ny - 2) {
            const int tau = T - 2 * i - 1;
            if (tau < 0) return;
            const int n = tau / L, j = tau % L + 1;
            if (n >= limit) return;
            if (j == 1) part_reset(n, i);
            if (j <= nx - 2) update(i, j, n, i);
            if (j == 4) update(i, 0, n, i);
            if (j == nx + 1) update(i, nx - 1, n, i);
        } else if (i == 0) {
            // first row at n*L + j + 4, UL corner at n*L + 8, UR corner at n*L + nx + 4
            const int tau = T - 5;
            if (tau < 0) return;
            const int n = tau / L, j = tau % L + 1;
            if (n >= limit) return;
            if (j == 1) part_reset(n, 0);
            if (j <= nx - 2) update(0, j, n, 0);
            if (j == 4) update(0, 0, n, 0);
            if (j == nx) update(0, nx - 1, n, 0);
        } else {
            // last row (natural time), BL corner at n*L + 2*ny + 1, BR corner at n*L + 2*ny + nx - 2
            const int tau = T - 2 * i - 1;
            if (tau < 0) return;
            const int n = tau / L, j = tau % L + 1;
            if (n >= limit) return;
            if (j == 1) part_reset(n, i);
            if (j <= nx - 2) update(i, j, n, i);
            if (j == 3) update(i, 0, n, i);
            if (j == nx) update(i, nx - 1, n, i);
        }
    }
    void part_reset(int n, int row) { if (account) part[(size_t) (n % D) * (ny + 1) + row] = 0.0; }
    // last update of sweep n: the BR corner, or (3-row images) the UL corner at n*L + 8
    int t_done(int n) const { return n * L + std::max(8, 2 * ny + nx - 2); }
};

void thread_order(std::vector<int> &ord, int mode, Rng &rng)
{
    const int n = (int) ord.size();
    for (int k = 0; k < n; k++) ord[k] = k;
    if (mode == 1) std::reverse(ord.begin(), ord.end());
    if (mode == 2) for (int k = n - 1; k > 0; k--) std::swap(ord[k], ord[rng.next() % (k + 1)]);
}

} // namespace

extern "C" {

// Returns the number of sweeps (as the reference's loop would), -1 for unsupported sizes.  *replayed =
// sweeps re-run after a restore, *speculated = sweeps that had been started beyond the stopping point.
int hs_emu_pipe_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                    float alpha2, double tol, int maxiter, int K, int nthreads, int order, unsigned seed,
                    double *err_out, int *replayed, int *speculated)
{
    if (nx < 3 || ny < 3 || maxiter < 1 || nthreads < 1) return -1;
    Pipe P;
    P.nx = nx; P.ny = ny; P.L = std::max(nx + 2, 8); P.maxiter = maxiter; P.alpha2 = alpha2;
    // a snapshot buffer must not be overwritten (2K sweeps later) before every decision that may need it
    P.K = std::max(K, (2 * ny + nx) / P.L + 2);
    P.ix = ix; P.iy = iy; P.rho = rho; P.u = u; P.v = v;
    const size_t n = (size_t) nx * ny;
    for (int q = 0; q < 2; q++) { P.snap_u[q].assign(u, u + n); P.snap_v[q].assign(v, v + n); }
    P.D = (2 * ny + nx) / P.L + 3;
    P.part.assign((size_t) P.D * (ny + 1), 0.0);
    P.account = true;
    P.limit = maxiter;

    Rng rng{ seed * 2654435761ull + 99 };
    std::vector<int> ord(nthreads);
    auto step = [&](int T) {
        thread_order(ord, order, rng);
        for (int tid : ord)
            for (int i = tid; i < ny; i += nthreads) P.row_work(T, i);
    };

    int decided = 0, niter = 0;
    double error = 1000;
    for (int T = 0;; T++) {
        step(T);                                            // one barrier per step
        if (T == P.t_done(decided)) {
            double e = 0;
            for (int r = 0; r <= ny; r++) e += P.part[(size_t) (decided % P.D) * (ny + 1) + r];   // fixed order
            error = sqrt(e / (nx * ny));
            niter = ++decided;
            if (!(error > tol && niter < maxiter)) break;
        }
    }
    int rep = 0, spec = 0;
    if (niter < maxiter) {
        // rows above the bottom ran ahead: started sweeps niter, niter+1, ...
        spec = (P.t_done(niter - 1) - 3) / P.L + 1 - niter;
        const int m = (niter / P.K) * P.K, q = (niter / P.K) & 1;
        memcpy(u, P.snap_u[q].data(), n * sizeof(float));
        memcpy(v, P.snap_v[q].data(), n * sizeof(float));
        rep = niter - m;
        if (rep > 0) {
            P.account = false;
            P.limit = rep;
            for (int T = 0; T <= P.t_done(rep - 1); T++) step(T);
        }
    }
    if (err_out) *err_out = error;
    if (replayed) *replayed = rep;
    if (speculated) *speculated = spec;
    return niter;
}

} // extern "C"
