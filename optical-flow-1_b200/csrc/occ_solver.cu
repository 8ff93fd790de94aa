// Host side of the C ABI declared in include/occ_b200.h: the coarse-to-fine driver
// (src/tvl1occflow.cpp:335-482) and the per-level driver (src/tvl1occflow.cpp:144-330) of the
// reference's TV-L1 solver with occlusion detection, for batches of independent frame triples that
// advance in lock-step.  Built with -fmad=false (see occ_kernels.cuh).  The outer loop of a warp step
// is driven from the host: one 4-byte read-back per outer iteration (an outer iteration is 20 box
// sweeps and 100 primal-dual iterations of the occlusion map; the round trip is noise next to it).
#include "../../include/occ_b200.h"
#include "occ_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace occ;

namespace {

thread_local std::string g_occ_create_error;

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct Level { int nx, ny; size_t N; };

enum Group { G_PYRAMID = 0, G_WARP, G_BOX, G_CHI, G_OTHER, G_COUNT };

struct Workspace {
    int nx = 0, ny = 0, nscales = 0, B = 0;
    double zfactor = 0;
    bool filt_alias = false;
    std::vector<Level> lv;
    double *pool = nullptr;
    size_t pool_doubles = 0;
    // per level
    std::vector<double *> im[4];      // I_1, I0, I1, filtI0: [B][N_s]
    std::vector<double *> U, chi;     // U: [2][B][N_s] (u1 planes, then u2 planes); chi: [B][N_s]
    // sized for the finest level
    double *Ig = nullptr;             // [4][B][N]  I1x, I1y, I_1x, I_1y
    double *Wc = nullptr;             // [8][B][N]  I1wx, I1wy, I_1wx, I_1wy, rho1_c, rho3_c, grad1, grad3
    double *g = nullptr, *F = nullptr, *AL = nullptr, *P = nullptr, *ETA = nullptr, *Vfwd = nullptr,
           *Vbck = nullptr, *C = nullptr, *Uprev = nullptr, *tmpU = nullptr;
    // wave-layout side of the box relaxation: coefficients [2B][9][N], f and alfa [2B][N], and a row-major
    // copy of the duals for the u update [4B][N] (P itself is kept in the wave layout, see rof_box)
    double *K = nullptr, *FW = nullptr, *ALW = nullptr, *Pn = nullptr;
    TripleCtl *ctl = nullptr;
    double *partials = nullptr;
    int parts = 0;
    int *n_active = nullptr;
    int *stat_iters = nullptr;
    double *stat_errs = nullptr;
    int stat_stride = 0;
};

struct Span { cudaEvent_t a, b; int group; };

} // namespace

struct occ_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool profiling = false;
    int max_batch = 64;
    bool gs_wave = true;             // OCC_GS_WAVE=0: Gauss-Seidel pass on row-major planes (A/B)
    bool gs_coef = false;            // OCC_GS_COEF=1: interior coefficients from a parallel kernel (k_occ_rof_coef): three
                                     // of seven divisions leave the serial pass, but their 72 B per cell cost more (A/B)
    bool chi_tb = false;             // OCC_CHI_TB=1: five occlusion-map iterations per launch on chip (k_occ_chi_tb; A/B:
                                     // bit-identical, but the loop is fp64-bound, not HBM-bound: 253 vs 221 ms)
    bool chi_march = false;          // OCC_CHI_MARCH=1: row-marching form of the occlusion-map iteration (A/B: 4 % slower)
    bool chi_fused = true;           // OCC_CHI_FUSED=0: the two-kernel form of the occlusion-map iteration (A/B)
    int sm_count = 148;
    occ_stats stats{};
    Workspace ws;
    int *h_n_active = nullptr;       // pinned
    double *stage = nullptr;         // device staging of the host-buffer entry points: [7][chunk][N]
    size_t stage_doubles = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<Span> spans;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                     __FILE__, __LINE__);                                                          \
            ctx->err = buf_;                                                                       \
            return OCC_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define CKL() CK(cudaGetLastError()); ctx->stats.kernel_launches++

#define TRY(expr)                                                                                  \
    do {                                                                                           \
        int rc_ = (expr);                                                                          \
        if (rc_ != OCC_OK) return rc_;                                                             \
    } while (0)

int fail_arg(occ_ctx *ctx, const char *msg)
{
    ctx->err = msg;
    return OCC_ERR_ARG;
}

// 1-D kernel of src/operators.cpp:525-539 (host libm, like the reference)
int make_taps(occ_ctx *ctx, double sigma, int width, GaussTaps &t)
{
    const double den = 2 * sigma * sigma;
    const int size = (int) (5 * sigma) + 1;      // window 5, src/operators.h:120
    if (size > (int) (sizeof t.B / sizeof t.B[0])) return fail_arg(ctx, "gaussian window larger than this build supports");
    if (size > width) {
        ctx->err = "GaussianSmooth: sigma too large";
        return OCC_ERR_SIGMA;
    }
    t.size = size;
    for (int i = 0; i < size; i++) t.B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += t.B[i];
    norm *= 2;
    norm -= t.B[0];
    for (int i = 0; i < size; i++) t.B[i] /= norm;
    return OCC_OK;
}

void free_workspace(Workspace &w)
{
    cudaFree(w.pool);
    cudaFree(w.ctl);
    cudaFree(w.partials);
    cudaFree(w.n_active);
    cudaFree(w.stat_iters);
    cudaFree(w.stat_errs);
    w = Workspace{};
}

int ensure_workspace(occ_ctx *ctx, int nx, int ny, int nscales, double zfactor, int B, bool filt_alias, int stat_stride)
{
    Workspace &w = ctx->ws;
    if (w.pool && w.nx == nx && w.ny == ny && w.nscales == nscales && w.zfactor == zfactor && w.B == B &&
        w.filt_alias == filt_alias && w.stat_stride >= stat_stride)
        return OCC_OK;
    free_workspace(w);
    w.nx = nx; w.ny = ny; w.nscales = nscales; w.zfactor = zfactor; w.B = B; w.filt_alias = filt_alias;
    w.stat_stride = stat_stride;
    w.lv.resize(nscales);
    w.lv[0] = Level{ nx, ny, (size_t) nx * ny };
    for (int s = 1; s < nscales; s++) {
        // zoom_size, src/zoom.cpp:22-34
        const int nxx = (int) (w.lv[s - 1].nx * zfactor + 0.5), nyy = (int) (w.lv[s - 1].ny * zfactor + 0.5);
        if (nxx < 1 || nyy < 1) return fail_arg(ctx, "pyramid level of zero size");
        w.lv[s] = Level{ nxx, nyy, (size_t) nxx * nyy };
    }
    size_t SN = 0;
    for (auto &l : w.lv) SN += l.N;
    const size_t N0 = w.lv[0].N, BN0 = (size_t) B * N0;
    const int n_im = filt_alias ? 3 : 4;
    const size_t total = (size_t) B * SN * (n_im + 3) + BN0 * (4 + 8 + 1 + 2 + 2 + 4 + 2 + 2 + 2 + 5 + 2 + 2 + 2 * kRofK + 2 + 2 + 4);
    CK(cudaMalloc(&w.pool, total * sizeof(double)));
    w.pool_doubles = total;
    double *p = w.pool;
    auto take = [&](size_t n) { double *r = p; p += n; return r; };
    for (int k = 0; k < 4; k++) w.im[k].assign(nscales, nullptr);
    w.U.assign(nscales, nullptr);
    w.chi.assign(nscales, nullptr);
    for (int s = 0; s < nscales; s++) {
        for (int k = 0; k < n_im; k++) w.im[k][s] = take((size_t) B * w.lv[s].N);
        if (filt_alias) w.im[3][s] = w.im[1][s];
        w.U[s] = take(2 * (size_t) B * w.lv[s].N);
        w.chi[s] = take((size_t) B * w.lv[s].N);
    }
    w.Ig = take(4 * BN0);
    w.Wc = take(8 * BN0);
    w.g = take(BN0);
    w.F = take(2 * BN0);
    w.AL = take(2 * BN0);
    w.P = take(4 * BN0);
    w.ETA = take(2 * BN0);
    w.Vfwd = take(2 * BN0);
    w.Vbck = take(2 * BN0);
    w.C = take(5 * BN0);
    w.Uprev = take(2 * BN0);
    w.tmpU = take(2 * BN0);
    w.K = take(2 * kRofK * BN0);
    w.FW = take(2 * BN0);
    w.ALW = take(2 * BN0);
    w.Pn = take(4 * BN0);
    w.parts = std::max(1, std::min(64, (int) (N0 / 4096)));
    CK(cudaMalloc(&w.ctl, sizeof(TripleCtl) * B));
    CK(cudaMalloc(&w.partials, sizeof(double) * B * w.parts));
    CK(cudaMalloc(&w.n_active, sizeof(int)));
    CK(cudaMalloc(&w.stat_iters, sizeof(int) * B * stat_stride));
    CK(cudaMalloc(&w.stat_errs, sizeof(double) * B * stat_stride));
    return OCC_OK;
}

// ---- profiling spans (CUDA events, resolved after the solve: no extra synchronisation) -------------
cudaEvent_t take_event(occ_ctx *ctx)
{
    cudaEvent_t e = nullptr;
    if (!ctx->ev_pool.empty()) { e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

struct Scope {
    occ_ctx *ctx;
    Span s{};
    bool on;
    Scope(occ_ctx *c, int group) : ctx(c), on(c->profiling)
    {
        if (!on) return;
        s.group = group;
        s.a = take_event(ctx);
        s.b = take_event(ctx);
        cudaEventRecord(s.a, ctx->stream);
    }
    ~Scope()
    {
        if (!on) return;
        cudaEventRecord(s.b, ctx->stream);
        ctx->spans.push_back(s);
    }
};

void resolve_spans(occ_ctx *ctx)
{
    double ms[G_COUNT] = {};
    for (auto &s : ctx->spans) {
        float t = 0;
        if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms[s.group] += t;
        ctx->ev_pool.push_back(s.a);
        ctx->ev_pool.push_back(s.b);
    }
    ctx->spans.clear();
    ctx->stats.ms_pyramid += ms[G_PYRAMID];
    ctx->stats.ms_warp += ms[G_WARP];
    ctx->stats.ms_box += ms[G_BOX];
    ctx->stats.ms_chi += ms[G_CHI];
    ctx->stats.ms_other += ms[G_OTHER];
}

inline dim3 grid2d(int nx, int ny, int z) { return dim3(ceil_div(nx, 32), ceil_div(ny, 8), z); }
const dim3 kBlock2d(32, 8);

// gaussian (src/operators.cpp:506-624): rows into tmp, columns into out; z planes of nx x ny
int gaussian(occ_ctx *ctx, const double *in, double *tmp, double *out, int nx, int ny, int z, double sigma)
{
    GaussTaps t;
    TRY(make_taps(ctx, sigma, nx, t));
    k_occ_gauss_pass<0><<<grid2d(nx, ny, z), kBlock2d, 0, ctx->stream>>>(in, tmp, nx, ny, t);
    CKL();
    k_occ_gauss_pass<1><<<grid2d(nx, ny, z), kBlock2d, 0, ctx->stream>>>(tmp, out, nx, ny, t);
    CKL();
    return OCC_OK;
}

// src/tvl1occflow.cpp:383-423: presmoothing of the four finest images, then zoom_out level by level
// (src/zoom.cpp:41-78).  The result of image_normalization_4 (:379-380) is overwritten by the raw images
// (:383-388) upstream, so the pyramid is built from the un-normalised inputs.
int build_pyramid(occ_ctx *ctx, const double *const src[4])
{
    Workspace &w = ctx->ws;
    Scope sc(ctx, G_PYRAMID);
    const int B = w.B, n_im = w.filt_alias ? 3 : 4;
    double *tmpA = w.AL, *tmpB = w.tmpU;      // [B][N0] each at least
    const double sigma_z = 0.6 * sqrt(1.0 / (w.zfactor * w.zfactor) - 1.0);   // ZOOM_SIGMA_ZERO, src/zoom.cpp:15,56
    for (int k = 0; k < n_im; k++) {
        TRY(gaussian(ctx, src[k], tmpA, w.im[k][0], w.lv[0].nx, w.lv[0].ny, B, OCC_PRESMOOTHING_SIGMA));
        for (int s = 1; s < w.nscales; s++) {
            const Level &a = w.lv[s - 1], &b = w.lv[s];
            TRY(gaussian(ctx, w.im[k][s - 1], tmpA, tmpB, a.nx, a.ny, B, sigma_z));
            k_occ_resample<<<grid2d(b.nx, b.ny, B), kBlock2d, 0, ctx->stream>>>(tmpB, w.im[k][s], a.nx, a.ny, b.nx, b.ny,
                                                                          w.zfactor, w.zfactor, 0.0);
            CKL();
        }
    }
    return OCC_OK;
}

int rof_threads(int ny) { return std::max(32, std::min(1024, ceil_div(ny, 32) * 32)); }

struct RofBufs {
    double *U, *P, *AL;            // row-major u [planes][N]; duals [2 planes][N] (wave layout if `wave`); alfa scratch
    const double *F, *g;
    double *K, *FW, *ALW, *Pn;     // wave mode only
};

// Scalar_ROF_BoxCellCentered on `planes` problems (2B inside the solver, 1 for the hook).  Wave mode (the
// default): the duals live in the wave layout; per call f goes there once, per sweep alfa does, the
// coefficient kernel and the Gauss-Seidel pass work there, and the u update reads a row-major copy.
int rof_box(occ_ctx *ctx, const TripleCtl *ctl, const RofBufs &R, int nx, int ny, int B, int planes, double lambda,
            double omega, int niter)
{
    cudaStream_t st = ctx->stream;
    const dim3 gT(ceil_div(nx, 32), ceil_div(ny, 32), planes), gT2(gT.x, gT.y, 2 * planes);
    if (ctx->gs_wave && niter > 0) {
        k_occ_wave_transpose<true><<<gT, kBlock2d, 0, st>>>(ctl, R.F, R.FW, nx, ny, B);
        CKL();
    }
    for (int it = 0; it < niter; it++) {
        k_occ_rof_alfa<<<grid2d(nx, ny, planes), kBlock2d, 0, st>>>(ctl, R.U, R.g, R.AL, nx, ny, B, lambda);
        CKL();
        if (ctx->gs_wave) {
            k_occ_wave_transpose<true><<<gT, kBlock2d, 0, st>>>(ctl, R.AL, R.ALW, nx, ny, B);
            CKL();
            if (ctx->gs_coef) {
                k_occ_rof_coef<<<dim3(ceil_div(ny, 32), ceil_div(nx, 8), planes), kBlock2d, 0, st>>>(ctl, R.ALW, R.K, nx, ny, B);
                CKL();
                k_occ_rof_gs_wave<true><<<planes, rof_threads(ny), 0, st>>>(ctl, R.P, R.FW, R.ALW, R.K, nx, ny, B, omega);
            } else {
                k_occ_rof_gs_wave<false><<<planes, rof_threads(ny), 0, st>>>(ctl, R.P, R.FW, R.ALW, R.K, nx, ny, B, omega);
            }
            CKL();
            k_occ_wave_transpose<false><<<gT2, kBlock2d, 0, st>>>(ctl, R.P, R.Pn, nx, ny, B);
            CKL();
            k_occ_rof_u<<<grid2d(nx, ny, planes), kBlock2d, 0, st>>>(ctl, R.U, R.F, R.Pn, nx, ny, B, lambda);
            CKL();
        } else {
            k_occ_rof_gs<<<planes, rof_threads(ny), 0, st>>>(ctl, R.P, R.F, R.AL, nx, ny, B, omega);
            CKL();
            k_occ_rof_u<<<grid2d(nx, ny, planes), kBlock2d, 0, st>>>(ctl, R.U, R.F, R.P, nx, ny, B, lambda);
            CKL();
        }
        ctx->stats.box_sweeps++;
    }
    return OCC_OK;
}

__global__ void __launch_bounds__(256) k_occ_copy_active(const TripleCtl *__restrict__ ctl, const double *__restrict__ in,
                                                     double *__restrict__ out, size_t N, int B)
{
    const int q = blockIdx.y, b = q % B;
    if (!ctl[b].active) return;
    const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (i < N) out[q * N + i] = in[q * N + i];
}

// Dual_TVL1_optic_flow of src/tvl1occflow.cpp:144-330 on level s of the workspace
int run_level(occ_ctx *ctx, int s, const occ_params &prm, int stat_base)
{
    Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    const int nx = l.nx, ny = l.ny, B = w.B;
    const size_t N = l.N, BN = (size_t) B * N;
    cudaStream_t st = ctx->stream;
    double *U = w.U[s], *chi = w.chi[s];
    double *I1x = w.Ig, *I1y = w.Ig + BN, *Im1x = w.Ig + 2 * BN, *Im1y = w.Ig + 3 * BN;
    double *I1wx = w.Wc, *I1wy = w.Wc + BN, *Im1wx = w.Wc + 2 * BN, *Im1wy = w.Wc + 3 * BN, *rho1 = w.Wc + 4 * BN,
           *rho3 = w.Wc + 5 * BN, *grad1 = w.Wc + 6 * BN, *grad3 = w.Wc + 7 * BN;
    const dim3 gB = grid2d(nx, ny, B), g2B = grid2d(nx, ny, 2 * B);
    const int ctl_blocks = ceil_div(B, 128);

    VParams vp;
    vp.l_t = prm.lambda * prm.theta;
    vp.one_pat = 1. + prm.alpha * prm.theta;
    vp.at_d_1pat = prm.alpha * prm.theta / vp.one_pat;
    vp.lt_d_1pat = 2. * prm.lambda * prm.theta / vp.one_pat;
    vp.theta = prm.theta;
    vp.beta = prm.beta;
    vp.is_zero = OCC_IS_ZERO;
    vp.thr_chi = OCC_THR_CHI;
    ChiParams cp;
    cp.lambda = prm.lambda;
    cp.half_over_theta = 0.5 / prm.theta;
    cp.alpha_theta = prm.alpha * prm.theta;
    cp.beta = prm.beta;
    cp.tau_chi = OCC_TAU_CHI;
    cp.tau_eta = OCC_TAU_ETA;
    cp.is_zero = OCC_IS_ZERO;

    {
        Scope sc(ctx, G_OTHER);
        // the dual variables of Solver_wrt_u / Solver_wrt_chi start from zero at every level (occ_b200.h)
        CK(cudaMemsetAsync(w.P, 0, 4 * BN * sizeof(double), st));
        CK(cudaMemsetAsync(w.ETA, 0, 2 * BN * sizeof(double), st));
        k_occ_level_setup<<<gB, kBlock2d, 0, st>>>(w.im[2][s], w.im[0][s], w.im[3][s], I1x, I1y, Im1x, Im1y, w.g, nx, ny,
                                               OCC_G_FACTOR);
        CKL();
        CK(cudaMemcpyAsync(w.Uprev, U, 2 * BN * sizeof(double), cudaMemcpyDeviceToDevice, st));   // :211-224
    }
    for (int wp = 0; wp < prm.warps; wp++) {
        {
            Scope sc(ctx, G_WARP);
            k_occ_warp<<<gB, kBlock2d, 0, st>>>(w.im[1][s], w.im[2][s], I1x, I1y, w.im[0][s], Im1x, Im1y, U, U + BN, I1wx,
                                            I1wy, Im1wx, Im1wy, rho1, rho3, grad1, grad3, nx, ny);
            CKL();
            k_occ_ctl_begin<<<ctl_blocks, 128, 0, st>>>(w.ctl, B);
            CKL();
        }
        int active_now = B;
        for (int n = 0; n < OCC_EXT_MAX_ITERATIONS; n++) {
            {
                Scope sc(ctx, G_OTHER);
                k_occ_solver_v<<<gB, kBlock2d, 0, st>>>(w.ctl, U, w.F, chi, I1wx, I1wy, Im1wx, Im1wy, rho1, rho3, grad1,
                                                    grad3, w.Vfwd, w.Vbck, nx, ny, B, vp);
                CKL();
            }
            {
                Scope sc(ctx, G_BOX);
                const RofBufs R = { U, w.P, w.AL, w.F, w.g, w.K, w.FW, w.ALW, w.Pn };
                TRY(rof_box(ctx, w.ctl, R, nx, ny, B, 2 * B, prm.theta, OCC_OMEGA, OCC_MAX_ITERATIONS_U));
            }
            {
                Scope sc(ctx, G_OTHER);
                k_occ_median3<<<g2B, kBlock2d, 0, st>>>(w.ctl, U, w.tmpU, nx, ny, B);
                CKL();
                k_occ_copy_active<<<dim3((unsigned) ((N + 255) / 256), 2 * B), 256, 0, st>>>(w.ctl, w.tmpU, U, N, B);
                CKL();
            }
            {
                Scope sc(ctx, G_CHI);
                k_occ_chi_setup<<<gB, kBlock2d, 0, st>>>(w.ctl, U, I1wx, I1wy, Im1wx, Im1wy, rho1, rho3, w.Vfwd, w.Vbck, w.C,
                                                     nx, ny, B, cp);
                CKL();
                if (ctx->chi_tb) {
                    // kChiT iterations per launch on chip, ping-pong between (chi, ETA) and (tmpU, AL), both free here
                    static_assert(OCC_MAX_ITERATIONS_CHI % (2 * kChiT) == 0, "an even number of launches ends where it began");
                    const dim3 gT(ceil_div(nx, kChiBW), ceil_div(ny, kChiBH), B);
                    for (int k = 0; k < OCC_MAX_ITERATIONS_CHI; k += 2 * kChiT) {
                        k_occ_chi_tb<<<gT, kChiTbThreads, kChiTbSmem, st>>>(w.ctl, chi, w.tmpU, w.g, w.ETA, w.AL, w.C, nx, ny, B, cp);
                        CKL();
                        k_occ_chi_tb<<<gT, kChiTbThreads, kChiTbSmem, st>>>(w.ctl, w.tmpU, chi, w.g, w.AL, w.ETA, w.C, nx, ny, B, cp);
                        CKL();
                    }
                } else if (ctx->chi_march) {
                    static_assert(OCC_MAX_ITERATIONS_CHI % 2 == 0, "the occlusion-map loop ping-pongs");
                    const dim3 gM(ceil_div(nx, 31), ceil_div(ny, 4 * kChiMR), B);
                    for (int k = 0; k < OCC_MAX_ITERATIONS_CHI; k += 2) {
                        k_occ_chi_march<<<gM, dim3(32, 4), 0, st>>>(w.ctl, chi, w.tmpU, w.g, w.ETA, w.AL, w.C, nx, ny, B, cp);
                        CKL();
                        k_occ_chi_march<<<gM, dim3(32, 4), 0, st>>>(w.ctl, w.tmpU, chi, w.g, w.AL, w.ETA, w.C, nx, ny, B, cp);
                        CKL();
                    }
                } else if (ctx->chi_fused) {
                    // ping-pong between (chi, ETA) and (tmpU, AL), both free here; an even count ends where it began
                    static_assert(OCC_MAX_ITERATIONS_CHI % 2 == 0, "the fused occlusion-map loop ping-pongs");
                    const dim3 gF(ceil_div(nx, kChiTW), ceil_div(ny, kChiTH), B);
                    for (int k = 0; k < OCC_MAX_ITERATIONS_CHI; k += 2) {
                        k_occ_chi_fused<<<gF, kBlock2d, 0, st>>>(w.ctl, chi, w.tmpU, w.g, w.ETA, w.AL, w.C, nx, ny, B, cp);
                        CKL();
                        k_occ_chi_fused<<<gF, kBlock2d, 0, st>>>(w.ctl, w.tmpU, chi, w.g, w.AL, w.ETA, w.C, nx, ny, B, cp);
                        CKL();
                    }
                } else {
                    for (int k = 0; k < OCC_MAX_ITERATIONS_CHI; k++) {
                        k_occ_chi_eta<<<gB, kBlock2d, 0, st>>>(w.ctl, chi, w.g, w.ETA, nx, ny, B, cp);
                        CKL();
                        k_occ_chi_update<<<gB, kBlock2d, 0, st>>>(w.ctl, chi, w.g, w.ETA, w.C, nx, ny, B, cp);
                        CKL();
                    }
                }
            }
            {
                Scope sc(ctx, G_OTHER);
                k_occ_error_partial<<<dim3(w.parts, B), kErrThreads, 0, st>>>(w.ctl, U, w.Uprev, w.partials, (int) N, B,
                                                                          w.parts);
                CKL();
                CK(cudaMemsetAsync(w.n_active, 0, sizeof(int), st));
                k_occ_error_decide<<<ctl_blocks, 128, 0, st>>>(w.ctl, w.partials, (int) N, B, w.parts, prm.epsilon,
                                                           OCC_EXT_MAX_ITERATIONS, w.n_active);
                CKL();
                CK(cudaMemcpyAsync(ctx->h_n_active, w.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            }
            CK(cudaStreamSynchronize(st));
            ctx->stats.host_syncs++;
            ctx->stats.box_cell_updates += (unsigned long long) active_now * 2 * N * OCC_MAX_ITERATIONS_U;
            ctx->stats.chi_pixel_iterations += (unsigned long long) active_now * N * OCC_MAX_ITERATIONS_CHI;
            active_now = *ctx->h_n_active;
            if (active_now == 0) break;
        }
        k_occ_ctl_end<<<ctl_blocks, 128, 0, st>>>(w.ctl, B, w.stat_iters, w.stat_errs, stat_base + wp, w.stat_stride);
        CKL();
    }
    return OCC_OK;
}

int fetch_stats(occ_ctx *ctx, int B, int nstat, int *iters_out, double *errs_out)
{
    Workspace &w = ctx->ws;
    std::vector<int> it((size_t) B * w.stat_stride);
    std::vector<double> er((size_t) B * w.stat_stride);
    CK(cudaMemcpyAsync(it.data(), w.stat_iters, it.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(er.data(), w.stat_errs, er.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.host_syncs++;
    for (int b = 0; b < B; b++)
        for (int k = 0; k < nstat; k++) {
            const int n = it[(size_t) b * w.stat_stride + k];
            ctx->stats.outer_iterations += n;
            if (iters_out) iters_out[(size_t) b * nstat + k] = n;
            if (errs_out) errs_out[(size_t) b * nstat + k] = er[(size_t) b * w.stat_stride + k];
        }
    return OCC_OK;
}

int check_params(occ_ctx *ctx, int nx, int ny, const occ_params *prm, bool multiscale)
{
    if (!prm) return fail_arg(ctx, "null parameters");
    if (nx < 2 || ny < 2) return fail_arg(ctx, "image smaller than 2x2");
    if ((size_t) nx * ny > (size_t) 1 << 30) return fail_arg(ctx, "image too large");
    if (prm->warps < 1 || prm->warps > 1024) return fail_arg(ctx, "warps out of range");
    if (multiscale && (prm->nscales < 1 || prm->nscales > OCC_MAX_LEVELS)) return fail_arg(ctx, "nscales out of range");
    if (multiscale && prm->nscales > 1 && !(prm->zfactor > 0 && prm->zfactor < 1)) return fail_arg(ctx, "zfactor out of range");
    return OCC_OK;
}

// B triples, device buffers [B][N]
int run_multiscale(occ_ctx *ctx, int B, const double *dIm1, const double *dI0, const double *dI1, const double *dfilt,
                   double *du1, double *du2, double *dchi, int nx, int ny, const occ_params &prm, int *iters_out,
                   double *errs_out)
{
    const bool alias = (dfilt == nullptr || dfilt == dI0);
    const int nstat = prm.nscales * prm.warps;
    TRY(ensure_workspace(ctx, nx, ny, prm.nscales, prm.nscales > 1 ? prm.zfactor : 0.5, B, alias, nstat));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->profiling) {
        t0 = take_event(ctx);
        t1 = take_event(ctx);
        cudaEventRecord(t0, st);
    }
    const double *src[4] = { dIm1, dI0, dI1, dfilt };
    TRY(build_pyramid(ctx, src));
    const int top = prm.nscales - 1;
    const size_t BNt = (size_t) B * w.lv[top].N;
    CK(cudaMemsetAsync(w.U[top], 0, 2 * BNt * sizeof(double), st));      // :361-366 and the new[] of :414-416
    CK(cudaMemsetAsync(w.chi[top], 0, BNt * sizeof(double), st));
    for (int s = top; s >= 0; s--) {
        TRY(run_level(ctx, s, prm, (top - s) * prm.warps));
        if (s) {
            // zoom_in of u1, u2, chi and the flow rescaling, :436-451
            Scope sc(ctx, G_OTHER);
            const Level &a = w.lv[s], &b = w.lv[s - 1];
            const double fx = (double) b.nx / a.nx, fy = (double) b.ny / a.ny;
            k_occ_resample<<<grid2d(b.nx, b.ny, 2 * B), kBlock2d, 0, st>>>(w.U[s], w.U[s - 1], a.nx, a.ny, b.nx, b.ny, fx, fy,
                                                                     (double) 1.0 / prm.zfactor);
            CKL();
            k_occ_resample<<<grid2d(b.nx, b.ny, B), kBlock2d, 0, st>>>(w.chi[s], w.chi[s - 1], a.nx, a.ny, b.nx, b.ny, fx, fy,
                                                                 0.0);
            CKL();
        }
    }
    {
        Scope sc(ctx, G_OTHER);
        const size_t BN = (size_t) B * w.lv[0].N;
        k_occ_threshold<<<(unsigned) ((BN + 255) / 256), 256, 0, st>>>(w.chi[0], BN, OCC_THR_CHI);
        CKL();
        CK(cudaMemcpyAsync(du1, w.U[0], BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(du2, w.U[0] + BN, BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dchi, w.chi[0], BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    if (ctx->profiling) cudaEventRecord(t1, st);
    TRY(fetch_stats(ctx, B, nstat, iters_out, errs_out));
    if (ctx->profiling) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        ctx->stats.ms_total += ms;
        ctx->ev_pool.push_back(t0);
        ctx->ev_pool.push_back(t1);
        resolve_spans(ctx);
    }
    return OCC_OK;
}

int run_single_scale(occ_ctx *ctx, int B, const double *dIm1, const double *dI0, const double *dI1, const double *dfilt,
                     double *du1, double *du2, double *dchi, int nx, int ny, const occ_params &prm, int *iters_out,
                     double *errs_out)
{
    const bool alias = (dfilt == nullptr || dfilt == dI0);
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, B, alias, prm.warps));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t BN = (size_t) B * w.lv[0].N;
    const double *src[4] = { dIm1, dI0, dI1, alias ? dI0 : dfilt };
    for (int k = 0; k < (alias ? 3 : 4); k++)
        CK(cudaMemcpyAsync(w.im[k][0], src[k], BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(w.U[0], du1, BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(w.U[0] + BN, du2, BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(w.chi[0], dchi, BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    TRY(run_level(ctx, 0, prm, 0));
    CK(cudaMemcpyAsync(du1, w.U[0], BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(du2, w.U[0] + BN, BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dchi, w.chi[0], BN * sizeof(double), cudaMemcpyDeviceToDevice, st));
    TRY(fetch_stats(ctx, B, prm.warps, iters_out, errs_out));
    if (ctx->profiling) resolve_spans(ctx);
    return OCC_OK;
}

int ensure_stage(occ_ctx *ctx, size_t doubles)
{
    if (ctx->stage_doubles >= doubles) return OCC_OK;
    cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_doubles = 0;
    CK(cudaMalloc(&ctx->stage, doubles * sizeof(double)));
    ctx->stage_doubles = doubles;
    return OCC_OK;
}

// host-buffer driver shared by the multiscale and the single-scale entry points
int solve_host(occ_ctx *ctx, int ntriples, const double *Im1, const double *I0, const double *I1, const double *filt,
               double *u1, double *u2, double *chi, int nx, int ny, const occ_params *prm, int *iters_out,
               double *errs_out, bool multiscale)
{
    if (!ctx) return OCC_ERR_ARG;
    if (!Im1 || !I0 || !I1 || !u1 || !u2 || !chi) return fail_arg(ctx, "null image or flow pointer");
    if (ntriples < 1) return fail_arg(ctx, "ntriples must be positive");
    TRY(check_params(ctx, nx, ny, prm, multiscale));
    CK(cudaSetDevice(ctx->device));
    ctx->stats = occ_stats{};
    const size_t N = (size_t) nx * ny;
    const int nstat = (multiscale ? prm->nscales : 1) * prm->warps;
    const bool alias = (filt == nullptr || filt == I0);
    const int chunk = std::min(ntriples, ctx->max_batch);
    TRY(ensure_stage(ctx, 7 * (size_t) chunk * N));
    cudaStream_t st = ctx->stream;
    for (int first = 0; first < ntriples; first += chunk) {
        const int B = std::min(chunk, ntriples - first);
        const size_t BN = (size_t) B * N, o = (size_t) first * N, cN = (size_t) chunk * N;
        double *d = ctx->stage;
        CK(cudaMemcpyAsync(d, Im1 + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d + cN, I0 + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d + 2 * cN, I1 + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
        if (!alias) CK(cudaMemcpyAsync(d + 3 * cN, filt + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
        if (!multiscale) {
            CK(cudaMemcpyAsync(d + 4 * cN, u1 + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d + 5 * cN, u2 + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d + 6 * cN, chi + o, BN * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        int *it = iters_out ? iters_out + (size_t) first * nstat : nullptr;
        double *er = errs_out ? errs_out + (size_t) first * nstat : nullptr;
        if (multiscale)
            TRY(run_multiscale(ctx, B, d, d + cN, d + 2 * cN, alias ? nullptr : d + 3 * cN, d + 4 * cN, d + 5 * cN,
                               d + 6 * cN, nx, ny, *prm, it, er));
        else
            TRY(run_single_scale(ctx, B, d, d + cN, d + 2 * cN, alias ? nullptr : d + 3 * cN, d + 4 * cN, d + 5 * cN,
                                 d + 6 * cN, nx, ny, *prm, it, er));
        CK(cudaMemcpyAsync(u1 + o, d + 4 * cN, BN * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(u2 + o, d + 5 * cN, BN * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(chi + o, d + 6 * cN, BN * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats.host_syncs++;
    }
    return OCC_OK;
}

} // namespace

extern "C" {

void occ_default_params(occ_params *p)
{
    if (!p) return;
    p->lambda = 0.15;
    p->alpha = 0.01;
    p->beta = 0.15;
    p->theta = 0.3;
    p->nscales = 100;
    p->zfactor = 0.5;
    p->warps = 2;
    p->epsilon = 0.01;
}

int occ_clamp_nscales(int nx, int ny, int nscales, double zfactor)
{
    const int N = (int) floor(log((float) std::min(nx, ny) / 16.0) / log(1. / zfactor)) + 1;
    return N < nscales ? N : nscales;
}

int occ_create(int device, occ_ctx **out)
{
    if (!out) return OCC_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        cudaGetLastError();
        g_occ_create_error = "no CUDA device: this library has no CPU fallback";
        return OCC_ERR_NODEVICE;
    }
    if (device < 0 || device >= count) {
        g_occ_create_error = "device index out of range";
        return OCC_ERR_ARG;
    }
    occ_ctx *ctx = new occ_ctx;
    ctx->device = device;
    auto fail = [&](const char *what, cudaError_t e) {
        g_occ_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        if (ctx->h_n_active) cudaFreeHost(ctx->h_n_active);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return OCC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaMallocHost(&ctx->h_n_active, sizeof(int))) != cudaSuccess) return fail("cudaMallocHost", e);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char *s = getenv("OCC_MAX_BATCH")) ctx->max_batch = std::max(1, atoi(s));
    if (const char *s = getenv("OCC_CHI_FUSED")) ctx->chi_fused = s[0] != '0';
    if (const char *s = getenv("OCC_CHI_TB")) ctx->chi_tb = s[0] == '1';
    if (const char *s = getenv("OCC_CHI_MARCH")) ctx->chi_march = s[0] == '1';
    if (cudaFuncSetAttribute(k_occ_chi_tb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kChiTbSmem) != cudaSuccess) {
        cudaGetLastError();
        ctx->chi_tb = false;
    }
    if (const char *s = getenv("OCC_GS_WAVE")) ctx->gs_wave = s[0] != '0';
    if (const char *s = getenv("OCC_GS_COEF")) ctx->gs_coef = s[0] != '0';
    *out = ctx;
    return OCC_OK;
}

void occ_destroy(occ_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_workspace(ctx->ws);
    cudaFree(ctx->stage);
    cudaFreeHost(ctx->h_n_active);
    for (auto &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *occ_last_error(const occ_ctx *ctx) { return ctx ? ctx->err.c_str() : g_occ_create_error.c_str(); }

int occ_set_profiling(occ_ctx *ctx, int on)
{
    if (!ctx) return OCC_ERR_ARG;
    ctx->profiling = on != 0;
    return OCC_OK;
}

int occ_set_max_batch(occ_ctx *ctx, int triples)
{
    if (!ctx || triples < 1) return OCC_ERR_ARG;
    ctx->max_batch = triples;
    return OCC_OK;
}

int occ_get_stats(const occ_ctx *ctx, occ_stats *out)
{
    if (!ctx || !out) return OCC_ERR_ARG;
    *out = ctx->stats;
    return OCC_OK;
}

void *occ_get_stream(const occ_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

int occ_solve_f64(occ_ctx *ctx, const double *I_1, const double *I0, const double *I1, const double *filtI0,
                  double *u1, double *u2, double *chi, int nx, int ny, const occ_params *prm, int *iters_out,
                  double *errs_out)
{
    return solve_host(ctx, 1, I_1, I0, I1, filtI0, u1, u2, chi, nx, ny, prm, iters_out, errs_out, true);
}

int occ_solve_batch_f64(occ_ctx *ctx, int ntriples, const double *I_1, const double *I0, const double *I1,
                        const double *filtI0, double *u1, double *u2, double *chi, int nx, int ny,
                        const occ_params *prm, int *iters_out, double *errs_out)
{
    return solve_host(ctx, ntriples, I_1, I0, I1, filtI0, u1, u2, chi, nx, ny, prm, iters_out, errs_out, true);
}

int occ_solve_batch_dev_f64(occ_ctx *ctx, int ntriples, const double *dI_1, const double *dI0, const double *dI1,
                            const double *dfiltI0, double *du1, double *du2, double *dchi, int nx, int ny,
                            const occ_params *prm, int *iters_out, double *errs_out)
{
    if (!ctx) return OCC_ERR_ARG;
    if (!dI_1 || !dI0 || !dI1 || !du1 || !du2 || !dchi) return fail_arg(ctx, "null image or flow pointer");
    if (ntriples < 1) return fail_arg(ctx, "ntriples must be positive");
    TRY(check_params(ctx, nx, ny, prm, true));
    CK(cudaSetDevice(ctx->device));
    ctx->stats = occ_stats{};
    const size_t N = (size_t) nx * ny;
    const int nstat = prm->nscales * prm->warps;
    for (int first = 0; first < ntriples; first += ctx->max_batch) {
        const int B = std::min(ctx->max_batch, ntriples - first);
        const size_t o = (size_t) first * N;
        TRY(run_multiscale(ctx, B, dI_1 + o, dI0 + o, dI1 + o, dfiltI0 ? dfiltI0 + o : nullptr, du1 + o, du2 + o,
                           dchi + o, nx, ny, *prm, iters_out ? iters_out + (size_t) first * nstat : nullptr,
                           errs_out ? errs_out + (size_t) first * nstat : nullptr));
    }
    return OCC_OK;
}

int occ_single_scale_f64(occ_ctx *ctx, const double *I_1, const double *I0, const double *I1,
                         const double *filtI0, double *u1, double *u2, double *chi, int nx, int ny,
                         const occ_params *prm, int *iters_out, double *errs_out)
{
    return solve_host(ctx, 1, I_1, I0, I1, filtI0, u1, u2, chi, nx, ny, prm, iters_out, errs_out, false);
}

int occ_rof_box_f64(occ_ctx *ctx, double *u, const double *f, double *p1, double *p2, const double *g,
                    double lambda, double omega, int nx, int ny, int niter)
{
    if (!ctx) return OCC_ERR_ARG;
    if (!u || !f || !p1 || !p2 || !g) return fail_arg(ctx, "null pointer");
    if (nx < 2 || ny < 2 || niter < 0) return fail_arg(ctx, "bad size");
    CK(cudaSetDevice(ctx->device));
    ctx->stats = occ_stats{};
    const size_t N = (size_t) nx * ny;
    TRY(ensure_stage(ctx, (12 + kRofK) * N));
    cudaStream_t st = ctx->stream;
    double *dU = ctx->stage, *dF = dU + N, *dP = dF + N, *dG = dP + 2 * N, *dAL = dG + N, *dK = dAL + N,
           *dFW = dK + kRofK * N, *dALW = dFW + N, *dPn = dALW + N, *dPin = dPn + 2 * N;
    CK(cudaMemcpyAsync(dU, u, N * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dF, f, N * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dG, g, N * sizeof(double), cudaMemcpyHostToDevice, st));
    const dim3 gT2(ceil_div(nx, 32), ceil_div(ny, 32), 2);
    double *dP0 = ctx->gs_wave ? dPin : dP;            // the caller's duals are row-major
    CK(cudaMemcpyAsync(dP0, p1, N * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dP0 + N, p2, N * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ctx->gs_wave) {
        k_occ_wave_transpose<true><<<gT2, kBlock2d, 0, st>>>(nullptr, dPin, dP, nx, ny, 1);
        CKL();
    }
    const RofBufs R = { dU, dP, dAL, dF, dG, dK, dFW, dALW, dPn };
    TRY(rof_box(ctx, nullptr, R, nx, ny, 1, 1, lambda, omega, niter));
    if (ctx->gs_wave) {
        k_occ_wave_transpose<false><<<gT2, kBlock2d, 0, st>>>(nullptr, dP, dPin, nx, ny, 1);
        CKL();
    }
    CK(cudaMemcpyAsync(u, dU, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p1, dP0, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p2, dP0 + N, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OCC_OK;
}

int occ_median3_f64(occ_ctx *ctx, double *a, int nx, int ny)
{
    if (!ctx) return OCC_ERR_ARG;
    if (!a || nx < 1 || ny < 1) return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    ctx->stats = occ_stats{};
    const size_t N = (size_t) nx * ny;
    TRY(ensure_stage(ctx, 2 * N));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->stage, a, N * sizeof(double), cudaMemcpyHostToDevice, st));
    k_occ_median3<<<grid2d(nx, ny, 1), kBlock2d, 0, st>>>(nullptr, ctx->stage, ctx->stage + N, nx, ny, 1);
    CKL();
    CK(cudaMemcpyAsync(a, ctx->stage + N, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OCC_OK;
}

} // extern "C"
