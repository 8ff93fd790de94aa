// TEST INFRASTRUCTURE.  CPU replay of the wavefront schedule of the Horn-Schunck SOR kernel.
//
// Includes the product's own per-thread step functions (optical-flow-1_b200/csrc/hs_sor_step.h, the
// same text nvcc compiles into k_hs_sor) and drives them the way the kernel does -- a barrier per time
// step, `nthreads` threads that each own rows tid, tid + nthreads, ..., asynchronous copies that land
// at any time between their issue and the wait that covers them -- but lets the test choose the
// adversary: the order in which the threads of a step run, whether fetches or updates of a step go
// first, and when an asynchronous copy reads its source and writes its destination.
// hs_emu_seq_sor is the sequential sweep (src/horn_schunck_pyramidal.cpp:144-230, the order of
// oracle/tvl1_oracle.c) in the same fp32 arithmetic: the schedule is correct iff both agree bit for
// bit under every adversary.
#define HS_SOR_EMULATE 1
#include "../../optical-flow-1_b200/csrc/hs_sor_step.h"
#include "../../optical-flow-1_b200/csrc/hs_sor_pipe.h"
#include "../../optical-flow-1_b200/csrc/hs_sor_pairs.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Pending { float *dst; const float *src; float value[4]; int n; int group; };

struct EmuCp {
    std::vector<Pending> q;
    int group = 0;
    int land = 0;          // 0: read + write at issue; 1: read at issue, write at completion; 2: both at completion
    void copy(float *dst, const float *src, int n)
    {
        if (land == 0) { for (int k = 0; k < n; k++) dst[k] = src[k]; return; }
        Pending p{ dst, src, { 0.f, 0.f, 0.f, 0.f }, n, group };
        for (int k = 0; k < n; k++) p.value[k] = src[k];
        q.push_back(p);
    }
    void cp4(float *dst, const float *src) { copy(dst, src, 1); }
    void cp8(hs::F2 *dst, const hs::F2 *src) { copy(&dst->x, &src->x, 2); }
    void cp16(hs::F4 *dst, const hs::F4 *src) { copy(&dst->x, &src->x, 4); }
    void commit() { group++; }
    // cp.async.wait_group n: at most the n most recently committed groups stay pending
    void wait(int n)
    {
        size_t keep = 0;
        for (size_t k = 0; k < q.size(); k++) {
            if (q[k].group < group - n)
                for (int c = 0; c < q[k].n; c++) q[k].dst[c] = (land == 1) ? q[k].value[c] : q[k].src[c];
            else q[keep++] = q[k];
        }
        q.resize(keep);
    }
};

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t) (s >> 33); }
};

void thread_order(std::vector<int> &ord, int mode, Rng &rng)
{
    const int n = (int) ord.size();
    for (int k = 0; k < n; k++) ord[k] = k;
    if (mode == 1) std::reverse(ord.begin(), ord.end());
    if (mode == 2) for (int k = n - 1; k > 0; k--) std::swap(ord[k], ord[rng.next() % (k + 1)]);
}

// sequential update of pixel (i, j) on row-major planes, clamped neighbours (+ the BR corner's order)
double seq_update(const float *ix, const float *iy, const float *rho, float *u, float *v, float alpha2,
                  int i, int j, int nx, int ny)
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
    return (double) e;
}

} // namespace

extern "C" {

// The reference's loop (:139-231) in fp32 with the kernel's per-pixel arithmetic, strictly sequential.
// u, v: row-major, in/out.  Returns the number of sweeps.
int hs_emu_seq_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                   float alpha2, double tol, int maxiter, double *err_out)
{
    int niter = 0;
    double error = 1000;
    while (error > tol && niter < maxiter) {
        niter++;
        double e = 0;
        for (int i = 1; i < ny - 1; i++)
            for (int j = 1; j < nx - 1; j++) e += seq_update(ix, iy, rho, u, v, alpha2, i, j, nx, ny);
        for (int j = 1; j < nx - 1; j++) {
            e += seq_update(ix, iy, rho, u, v, alpha2, 0, j, nx, ny);
            e += seq_update(ix, iy, rho, u, v, alpha2, ny - 1, j, nx, ny);
        }
        for (int i = 1; i < ny - 1; i++) {
            e += seq_update(ix, iy, rho, u, v, alpha2, i, 0, nx, ny);
            e += seq_update(ix, iy, rho, u, v, alpha2, i, nx - 1, nx, ny);
        }
        e += seq_update(ix, iy, rho, u, v, alpha2, 0, 0, nx, ny);
        e += seq_update(ix, iy, rho, u, v, alpha2, 0, nx - 1, nx, ny);
        e += seq_update(ix, iy, rho, u, v, alpha2, ny - 1, 0, nx, ny);
        e += seq_update(ix, iy, rho, u, v, alpha2, ny - 1, nx - 1, nx, ny);
        error = sqrt(e / (nx * ny));
    }
    if (err_out) *err_out = error;
    return niter;
}

// The kernel's schedule.  order: 0 ascending thread ids, 1 descending, 2 shuffled per phase.
// phase: 0 all fetches of a step then all updates, 1 per thread fetch + update, 2 all updates then
// all fetches.  land: see EmuCp.  Returns the number of sweeps, or -1 for unsupported sizes.
int hs_emu_wave_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                    float alpha2, double tol, int maxiter, int P, int nthreads, int order, int phase,
                    int land, unsigned seed, double *err_out)
{
    if (nx < 3 || ny < 3 || P < 0 || P > hs::kMaxPrefetch || nthreads < 1) return -1;
    const size_t n = (size_t) nx * ny;
    std::vector<hs::F2> wuv(n), wxy(n);
    std::vector<float> wrho(n);
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int w = hs::wave_index(i, j, nx, ny), p = i * nx + j;
            wuv[w].x = u[p]; wuv[w].y = v[p];
            wxy[w].x = ix[p]; wxy[w].y = iy[p];
            wrho[w] = rho[p];
        }

    hs::SorView V;
    V.wuv = wuv.data(); V.wxy = wxy.data(); V.wrho = wrho.data();
    V.nx = nx; V.ny = ny; V.alpha2 = alpha2;
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = ny + 3;
    // poison the rings: a read of a slot that was never fetched shows up as NaN in the result
    std::vector<hs::F2> ring((size_t) (V.S + V.CD) * V.rp, hs::F2{ nanf(""), nanf("") });
    std::vector<float> ring_rho((size_t) V.CD * V.rp, nanf(""));
    V.ring_uv = ring.data();
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = ring_rho.data();

    EmuCp cp;
    cp.land = land;
    Rng rng{ seed * 2654435761ull + 12345 };
    std::vector<int> ord(nthreads);
    int niter = 0;
    double error = 1000;
    while (error > tol && niter < maxiter) {
        niter++;
        std::vector<double> esum(nthreads, 0.0);
        hs::Step s = hs::make_step(V, hs::first_step(V));
        for (int t = hs::first_step(V); t <= hs::last_step(V); t++, hs::advance(V, s)) {
            cp.wait(P);                                     // cp.async.wait_group P; __syncthreads()
            const hs::Step chk = hs::make_step(V, t);       // advance() must agree with the closed form
            if (memcmp(&s, &chk, sizeof s) != 0) return -2;
            auto fetch = [&](int tid) { for (int i = tid; i < ny; i += nthreads) hs::issue_row(V, s, i, cp); };
            auto update = [&](int tid) {
                if (t >= 3) for (int i = tid; i < ny; i += nthreads) esum[tid] += hs::compute_row(V, s, i);
            };
            if (phase == 0) {
                thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
                thread_order(ord, order, rng); for (int tid : ord) update(tid);
            } else if (phase == 1) {
                thread_order(ord, order, rng); for (int tid : ord) { fetch(tid); update(tid); }
            } else {
                thread_order(ord, order, rng); for (int tid : ord) update(tid);
                thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
            }
            cp.commit();                                    // one group per step (every thread commits)
        }
        cp.wait(0);                                         // cp.async.wait_group 0; __syncthreads()
        double e = hs::corners(V);                          // thread 0
        for (int tid = 0; tid < nthreads; tid++) e += esum[tid];
        error = sqrt(e / (nx * ny));
    }
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int w = hs::wave_index(i, j, nx, ny), p = i * nx + j;
            u[p] = wuv[w].x;
            v[p] = wuv[w].y;
        }
    if (err_out) *err_out = error;
    return niter;
}

// The PIPELINED schedule (hs_sor_pipe.h) driven like k_hs_sor_pipe: speculative phase with error
// accounting and snapshots, decision one barrier after a sweep completes, restore + replay on a stop by
// TOL.  Same adversary as hs_emu_wave_sor.  *replayed: sweeps re-run after the restore.
int hs_emu_pipe_wave_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                         float alpha2, double tol, int maxiter, int K, int P, int nthreads, int order, int phase,
                         int land, unsigned seed, double *err_out, int *replayed)
{
    if (nx < 3 || ny < 3 || P < 0 || P > hs::kMaxPrefetch || nthreads < 1 || maxiter < 1) return -1;
    const int L = hs::pipe_period(nx);
    const size_t n = (size_t) L * ny;
    std::vector<hs::F2> wuv(n, hs::F2{ nanf(""), nanf("") }), wxy(n, hs::F2{ nanf(""), nanf("") });
    std::vector<float> wrho(n, nanf(""));
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int w = hs::pipe_wave_index(i, j, L, ny), p = i * nx + j;
            wuv[w].x = u[p]; wuv[w].y = v[p];
            wxy[w].x = ix[p]; wxy[w].y = iy[p];
            wrho[w] = rho[p];
        }
    std::vector<hs::F2> snap0(wuv), snap1(wuv);

    hs::PipeView V;
    V.wuv = wuv.data(); V.wxy = wxy.data(); V.wrho = wrho.data();
    V.snap0 = snap0.data(); V.snap1 = snap1.data();
    V.nx = nx; V.ny = ny; V.L = L; V.alpha2 = alpha2;
    V.K = hs::pipe_snapshot_period(K, L, nx, ny);
    V.D = hs::pipe_error_depth(L, nx, ny);
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = ny + 3;
    std::vector<double> part((size_t) V.D * V.rp, nan("")), esum(V.rp, 0.0);
    V.part = part.data(); V.esum = esum.data();
    std::vector<hs::F2> ring((size_t) (V.S + V.CD) * V.rp, hs::F2{ nanf(""), nanf("") });
    std::vector<float> ring_rho((size_t) V.CD * V.rp, nanf(""));
    V.ring_uv = ring.data();
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = ring_rho.data();

    EmuCp cp;
    cp.land = land;
    Rng rng{ seed * 2654435761ull + 4711 };
    std::vector<int> ord(nthreads);
    const int T_first = -4 - P;
    // per-thread row positions, advanced like the kernel does (no division in the time loop)
    std::vector<hs::RowPos> base(nthreads);
    const int step_dn = (2 * nthreads) / L, step_dj = (2 * nthreads) % L;
    auto reset_positions = [&]() { for (int tid = 0; tid < nthreads; tid++) base[tid] = hs::pipe_pos(T_first - 2 * tid, L); };
    auto step = [&](const hs::PipeStep &s) {
        auto fetch = [&](int tid) {
            hs::RowPos p = base[tid];
            for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L)) {
                const hs::RowPos chk = hs::pipe_pos(s.T - 2 * i, L);
                if (chk.n != p.n || chk.j != p.j) abort();
                hs::pipe_issue_row(V, s, i, p, cp);
            }
        };
        auto update = [&](int tid) {
            hs::RowPos p = base[tid];
            if (s.T >= 1)
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pipe_compute_row(V, s, i, p);
        };
        if (phase == 0) {
            thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
            thread_order(ord, order, rng); for (int tid : ord) update(tid);
        } else if (phase == 1) {
            thread_order(ord, order, rng); for (int tid : ord) { fetch(tid); update(tid); }
        } else {
            thread_order(ord, order, rng); for (int tid : ord) update(tid);
            thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
        }
        cp.commit();
        for (int tid = 0; tid < nthreads; tid++) base[tid] = hs::pipe_pos_add(base[tid], 1, L);
    };

    // speculative phase
    V.limit = maxiter;
    V.account = 1;
    reset_positions();
    int decided = 0, niter = 0;
    double error = 1000;
    hs::PipeStep s = hs::pipe_make_step(V, T_first);
    for (int T = T_first;; T++, hs::pipe_advance(V, s)) {
        cp.wait(P);                                         // cp.async.wait_group P; __syncthreads()
        const hs::PipeStep chk = hs::pipe_make_step(V, T);
        if (memcmp(&s, &chk, sizeof s) != 0) return -2;
        if (T == hs::pipe_t_done(decided, L, nx, ny) + 1) {
            double e = 0;
            for (int r = 0; r < ny; r++) e += part[(size_t) (decided % V.D) * V.rp + r];   // fixed order
            error = sqrt(e / (nx * ny));
            niter = ++decided;
            if (!(error > tol && niter < maxiter)) break;   // (an extra barrier in the kernel)
        }
        step(s);
    }
    cp.wait(0);
    int rep = 0;
    if (niter < maxiter) {
        const int m = (niter / V.K) * V.K;
        const std::vector<hs::F2> &src = ((niter / V.K) & 1) ? snap1 : snap0;
        std::copy(src.begin(), src.end(), wuv.begin());
        rep = niter - m;
        if (rep > 0) {
            V.limit = rep;
            V.account = 0;
            std::fill(ring.begin(), ring.end(), hs::F2{ nanf(""), nanf("") });
            reset_positions();
            hs::PipeStep r = hs::pipe_make_step(V, T_first);
            for (int T = T_first; T <= hs::pipe_t_done(rep - 1, L, nx, ny); T++, hs::pipe_advance(V, r)) {
                cp.wait(P);
                step(r);
            }
            cp.wait(0);
        }
    }
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int w = hs::pipe_wave_index(i, j, L, ny), p = i * nx + j;
            u[p] = wuv[w].x;
            v[p] = wuv[w].y;
        }
    if (err_out) *err_out = error;
    if (replayed) *replayed = rep;
    return niter;
}

// The two-columns-per-step schedule (hs_sor_pairs.h), driven exactly like hs_emu_pipe_wave_sor.
int hs_emu_pairs_wave_sor(float *u, float *v, const float *ix, const float *iy, const float *rho, int nx, int ny,
                          float alpha2, double tol, int maxiter, int K, int P, int nthreads, int order, int phase,
                          int land, unsigned seed, double *err_out, int *replayed)
{
    if (nx < 3 || ny < 3 || P < 0 || P > hs::kMaxPrefetch || nthreads < 1 || maxiter < 1) return -1;
    const int L = hs::pairs_period(nx);
    const size_t n = (size_t) L * ny;
    const float qn = nanf("");
    std::vector<hs::F4> wuv(n, hs::F4{ qn, qn, qn, qn }), wxy(n, hs::F4{ qn, qn, qn, qn });
    std::vector<hs::F2> wrho(n, hs::F2{ qn, qn });
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const int w = hs::pairs_px_index(i, j, L, ny), p = i * nx + j;
            ((hs::F2 *) wuv.data())[w] = hs::F2{ u[p], v[p] };
            ((hs::F2 *) wxy.data())[w] = hs::F2{ ix[p], iy[p] };
            ((float *) wrho.data())[w] = rho[p];
        }
    std::vector<hs::F4> snap0(wuv), snap1(wuv);

    hs::PairsView V;
    V.wuv = wuv.data(); V.wxy = wxy.data(); V.wrho = wrho.data();
    V.snap0 = snap0.data(); V.snap1 = snap1.data();
    V.nx = nx; V.ny = ny; V.L = L; V.alpha2 = alpha2;
    V.cl = hs::pairs_cl(nx); V.cn = hs::pairs_cn(nx);
    V.K = hs::pairs_snapshot_period(K, L, nx, ny);
    V.D = hs::pairs_error_depth(L, nx, ny);
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = ny + 3;
    std::vector<double> part((size_t) V.D * V.rp, nan("")), esum(V.rp, 0.0);
    V.part = part.data(); V.esum = esum.data();
    std::vector<hs::F4> ring((size_t) (V.S + V.CD) * V.rp, hs::F4{ qn, qn, qn, qn });
    std::vector<hs::F2> ring_rho((size_t) V.CD * V.rp, hs::F2{ qn, qn });
    V.ring_uv = ring.data();
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = ring_rho.data();
    // PipeStep / pipe_make_step only use S, CD, L, P of the view
    hs::PipeView W;
    W.S = V.S; W.CD = V.CD; W.L = V.L; W.P = V.P;

    EmuCp cp;
    cp.land = land;
    Rng rng{ seed * 2654435761ull + 1234567 };
    std::vector<int> ord(nthreads);
    const int T_first = -4 - P;
    std::vector<hs::RowPos> base(nthreads);
    const int step_dn = (2 * nthreads) / L, step_dj = (2 * nthreads) % L;
    auto reset_positions = [&]() { for (int tid = 0; tid < nthreads; tid++) base[tid] = hs::pipe_pos(T_first - 2 * tid, L); };
    auto step = [&](const hs::PipeStep &s) {
        auto fetch = [&](int tid) {
            hs::RowPos p = base[tid];
            for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                hs::pairs_issue_row(V, s, i, p, cp);
        };
        auto update = [&](int tid) {
            hs::RowPos p = base[tid];
            if (s.T >= 0)
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pairs_compute_row(V, s, i, p);
        };
        if (phase == 0) {
            thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
            thread_order(ord, order, rng); for (int tid : ord) update(tid);
        } else if (phase == 1) {
            thread_order(ord, order, rng); for (int tid : ord) { fetch(tid); update(tid); }
        } else {
            thread_order(ord, order, rng); for (int tid : ord) update(tid);
            thread_order(ord, order, rng); for (int tid : ord) fetch(tid);
        }
        cp.commit();
        for (int tid = 0; tid < nthreads; tid++) base[tid] = hs::pipe_pos_add(base[tid], 1, L);
    };

    V.limit = maxiter;
    V.account = 1;
    reset_positions();
    int decided = 0, niter = 0;
    double error = 1000;
    hs::PipeStep s = hs::pipe_make_step(W, T_first);
    for (int T = T_first;; T++, hs::pipe_advance(W, s)) {
        cp.wait(P);
        if (T == hs::pairs_t_done(decided, L, nx, ny) + 1) {
            double e = 0;
            for (int r = 0; r < ny; r++) e += part[(size_t) (decided % V.D) * V.rp + r];
            error = sqrt(e / (nx * ny));
            niter = ++decided;
            if (!(error > tol && niter < maxiter)) break;
        }
        step(s);
    }
    cp.wait(0);
    int rep = 0;
    if (niter < maxiter) {
        const int m = (niter / V.K) * V.K;
        const std::vector<hs::F4> &src = ((niter / V.K) & 1) ? snap1 : snap0;
        std::copy(src.begin(), src.end(), wuv.begin());
        rep = niter - m;
        if (rep > 0) {
            V.limit = rep;
            V.account = 0;
            std::fill(ring.begin(), ring.end(), hs::F4{ qn, qn, qn, qn });
            reset_positions();
            hs::PipeStep r = hs::pipe_make_step(W, T_first);
            for (int T = T_first; T <= hs::pairs_t_done(rep - 1, L, nx, ny); T++, hs::pipe_advance(W, r)) {
                cp.wait(P);
                step(r);
            }
            cp.wait(0);
        }
    }
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            const hs::F2 uv = ((const hs::F2 *) wuv.data())[hs::pairs_px_index(i, j, L, ny)];
            u[i * nx + j] = uv.x;
            v[i * nx + j] = uv.y;
        }
    if (err_out) *err_out = error;
    if (replayed) *replayed = rep;
    return niter;
}

} // extern "C"
