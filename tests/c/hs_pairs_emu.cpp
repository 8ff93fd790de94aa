// TEST INFRASTRUCTURE / DESIGN PROTOTYPE ("Next for this kernel").
//
// CPU model of the next Horn-Schunck SOR schedule: TWO columns per thread-step on top of the pipelined
// sweeps of hs_pipeline_emu.cpp / hs_sor_pipe.h.  Row i processes the column pair c = (2c, 2c+1) of sweep n
// at global time
//     T = n*L + 2*i + c,      L = max(cl + 5, 12),  cl = (nx-2)/2 the pair holding the last interior column,
// left pixel first.  Per pixel this halves the barriers, the index arithmetic and the loop overhead, and
// the ring becomes 16-byte elements (9 LDS.128 per pair instead of 18 LDS.64).  The row skew stays 2 steps
// (= 4 columns): (i, 2c+1) needs the new (i-1, 2c+2), which row i-1 wrote one step earlier.  Borders:
//     first row pair c at n*L + c + 4; UL corner at n*L + 6, UR corner at n*L + cl + 6          (thread of row 0)
//     first column (i, 0) at n*L + 2i + 3, last column (i, nx-1) at n*L + 2i + cl + 3            (thread of row i)
//     last row at its natural time; BL corner at n*L + 2*ny, BR corner at n*L + 2*ny + cl        (thread ny-1)
// Stopping rule: snapshots + replay exactly as in hs_pipeline_emu.cpp.  Proven here against the
// sequential sweep on plain arrays (tests/test_hs_pipeline_proto.py); not yet a kernel.
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

    int cl() const { return (nx - 2) / 2; }

    void pair(int i, int c, int n, int part_row)
    {
        const int j0 = 2 * c, j1 = 2 * c + 1;
        if (j0 >= 1 && j0 <= nx - 2) update(i, j0, n, part_row);
        if (j1 >= 1 && j1 <= nx - 2) update(i, j1, n, part_row);
    }

    // everything the thread of row i does at global time T
    void row_work(int T, int i)
    {
        if (i >= 1 && i <= ny - 2) {
            const int x = T - 2 * i;
            if (x < 0) return;
            const int n = x / L, c = x % L;
            if (n >= limit) return;
            if (c == 0) part_reset(n, i);
            if (c <= cl()) pair(i, c, n, i);
            if (c == 3) update(i, 0, n, i);
            if (c == cl() + 3) update(i, nx - 1, n, i);
        } else if (i == 0) {
            const int x = T - 4;
            if (x < 0) return;
            const int n = x / L, c = x % L;
            if (n >= limit) return;
            if (c == 0) part_reset(n, 0);
            if (c <= cl()) pair(0, c, n, 0);
            if (c == 2) update(0, 0, n, 0);
            if (c == cl() + 2) update(0, nx - 1, n, 0);
        } else {
            const int x = T - 2 * i;
            if (x < 0) return;
            const int n = x / L, c = x % L;
            if (n >= limit) return;
            if (c == 0) part_reset(n, i);
            if (c <= cl()) pair(i, c, n, i);
            if (c == 2) update(i, 0, n, i);
            if (c == cl() + 2) update(i, nx - 1, n, i);
        }
    }
    void part_reset(int n, int row) { if (account) part[(size_t) (n % D) * (ny + 1) + row] = 0.0; }
    // last update of sweep n: the BR corner at n*L + 2*(ny-1) + cl + 2, or (3-row images) the upper corners
    int t_done(int n) const { return n * L + std::max(cl() + 6, 2 * ny + cl()); }
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
int hs_emu_pairs_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                    float alpha2, double tol, int maxiter, int K, int nthreads, int order, unsigned seed,
                    double *err_out, int *replayed, int *speculated)
{
    if (nx < 3 || ny < 3 || maxiter < 1 || nthreads < 1) return -1;
    Pipe P;
    P.nx = nx; P.ny = ny; P.L = std::max((nx - 2) / 2 + 5, 12); P.maxiter = maxiter; P.alpha2 = alpha2;
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
        spec = (P.t_done(niter - 1) - 2) / P.L + 1 - niter;
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
