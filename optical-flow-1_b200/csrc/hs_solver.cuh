// Host side of include/hs_b200.h: the coarse-to-fine driver (src/horn_schunck_pyramidal.cpp:258-370)
// and the per-level driver (:78-249) of the pyramidal Horn-Schunck method on top of the TV-L1 path's
// workspace, pyramid, warp and up-sampling kernels.  Included by tvl1_solver.cu (one translation unit,
// the kernels are shared).
#pragma once

namespace {

// Prefetch distance of k_hs_sor for a level.  Measured on 148 / 296 pairs of 1920x1080
// (profiles/r1h_hs2_*.json): the distance itself does not matter (P = 0, 1, 3 within 3 %: a time step of
// a 1080-row image takes longer than an HBM round trip), but two CTAs per SM do (+47 % throughput:
// one pair's barrier stalls are filled by the other's work).  So: the smallest rings first.
constexpr size_t kHsSmemOneCta = kHsSmemLimit - 4096;              // one CTA per SM
constexpr size_t kHsSmemTwoCtas = (228 * 1024) / 2 - 1024 - 512;    // two CTAs per SM (1 KB reserved + static each)

int hs_pick_prefetch(int ny, int want)
{
    const int rp = round_up(ny, 32);
    if (want >= 0 && want <= hs::kMaxPrefetch) return hs_ring_bytes(want, rp) <= kHsSmemOneCta ? want : -1;
    if (hs_ring_bytes(1, rp) <= kHsSmemTwoCtas) return 1;
    if (hs_ring_bytes(0, rp) <= kHsSmemTwoCtas) return 0;
    if (hs_ring_bytes(1, rp) <= kHsSmemOneCta) return 1;
    if (hs_ring_bytes(0, rp) <= kHsSmemOneCta) return 0;
    return -1;
}

// Levels whose rows do not fit the shared-memory rings keep the rings in the pair's wave region in
// global memory (k_hs_sor<0, true>): needs 22 floats per (padded) row behind the 5n floats of the planes.
bool hs_global_ring_fits(const Workspace &w, const Level &l)
{
    const size_t n = (size_t) l.nx * l.ny;
    return hs_global_ring_offset(n) + hs_global_ring_floats(round_up(l.ny, 32)) <= 6 * w.plane0;
}

int hs_check_level(tvl1_ctx *ctx, const Level &l)
{
    if (l.nx < 3 || l.ny < 3) return fail_arg(ctx, "Horn-Schunck: every pyramid level needs nx >= 3 and ny >= 3");
    if (hs_pick_prefetch(l.ny, -1) < 0 && !hs_global_ring_fits(ctx->ws, l))
        return fail_arg(ctx, "Horn-Schunck: image too tall and narrow (more than HS_MAX_ROWS rows need nx >= 24)");
    return TVL1_OK;
}

template <int P, bool G>
int hs_launch_sor_p(tvl1_ctx *ctx, const HsSorParams &A, int B, int threads, size_t smem)
{
    static bool attr_done[64] = { false };
    if (!G && !attr_done[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(k_hs_sor<P, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHsSmemOneCta));
        attr_done[ctx->device & 63] = true;
    }
    k_hs_sor<P, G><<<B, threads, G ? 0 : smem, ctx->stream>>>(A);
    CKL(ctx);
    return TVL1_OK;
}

template <int P>
int hs_launch_pipe_p(tvl1_ctx *ctx, const HsSorParams &A, int B, int threads, size_t smem)
{
    static bool attr_done[64] = { false };
    if (!attr_done[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(k_hs_sor_pipe<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHsSmemOneCta));
        attr_done[ctx->device & 63] = true;
    }
    k_hs_sor_pipe<P><<<B, threads, smem, ctx->stream>>>(A);
    CKL(ctx);
    return TVL1_OK;
}

// hs_sor_f32(prefetch = ...): tests force a kernel variant
constexpr int kHsForceGlobalRing = -2;     // rings in global memory
constexpr int kHsForcePipelined = -3;      // pipelined sweeps (k_hs_sor_pipe)
constexpr int kHsForcePairs = -4;          // pipelined sweeps, two columns per thread-step (k_hs_sor_pairs, experimental)

enum HsKernel { HS_ONE_SWEEP = 0, HS_PIPELINED = 1, HS_PAIRS = 2 };

// ---- pipelined sweeps (k_hs_sor_pipe, hs_sor_pipe.h) ------------------------------------------------
// Shared memory of a CTA: the rings plus one double per (padded) row.
size_t hs_pipe_smem(int P, int rp) { return hs_ring_bytes(P, rp) + sizeof(double) * rp; }

int hs_pipe_prefetch(int ny, int want)
{
    const int rp = round_up(ny, 32);
    if (want >= 0 && want <= hs::kMaxPrefetch) return hs_pipe_smem(want, rp) <= kHsSmemOneCta ? want : -1;
    if (hs_pipe_smem(1, rp) <= kHsSmemTwoCtas) return 1;
    if (hs_pipe_smem(0, rp) <= kHsSmemTwoCtas) return 0;
    if (hs_pipe_smem(1, rp) <= kHsSmemOneCta) return 1;
    if (hs_pipe_smem(0, rp) <= kHsSmemOneCta) return 0;
    return -1;
}

// Can level l run pipelined?  Its wave planes have period L = nx + 2 (5 L ny floats) and must fit the
// pair's region of the idle ping-pong set; the rings must fit one SM.
bool hs_pipe_fits(const Workspace &w, const Level &l)
{
    const int L = hs::pipe_period(l.nx);
    return l.nx >= 17 && (size_t) 5 * L * l.ny <= 6 * w.plane0 && hs_pipe_prefetch(l.ny, -1) >= 0;
}

// ---- two columns per thread-step (k_hs_sor_pairs; bit-equal to the sequential sweep on a B200, tests/test_hs_gpu.py; selected with HS_PAIRS=1) ----
int hs_pairs_prefetch(int ny, int want)
{
    const int rp = round_up(ny, 32);
    if (want >= 0 && want <= hs::kMaxPrefetch) return hs_pairs_smem(want, rp) <= kHsSmemOneCta ? want : -1;
    if (hs_pairs_smem(1, rp) <= kHsSmemOneCta) return 1;
    if (hs_pairs_smem(0, rp) <= kHsSmemOneCta) return 0;
    return -1;
}

bool hs_pairs_fits(const Workspace &w, const Level &l)
{
    const int L = hs::pairs_period(l.nx);
    return l.nx >= 32 && (size_t) 10 * L * l.ny <= 6 * w.plane0 && hs_pairs_prefetch(l.ny, -1) >= 0;
}

// The two-column kernel serves every level it fits (measured on 148 x 1080p, 5 levels x 3 warps x 30
// sweeps: 880 -> 752 ms against the pipelined one, profiles/r2c_hs_*.log); HS_PAIRS=0 switches it off.
bool hs_pairs_enabled()
{
    const char *e = std::getenv("HS_PAIRS");
    return !e || std::atoi(e) != 0;
}

template <int P>
int hs_launch_pairs_p(tvl1_ctx *ctx, const HsSorParams &A, int B, int threads, size_t smem)
{
    static bool attr_done[64] = { false };
    if (!attr_done[ctx->device & 63]) {
        CK(cudaFuncSetAttribute(k_hs_sor_pairs<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHsSmemOneCta));
        attr_done[ctx->device & 63] = true;
    }
    k_hs_sor_pairs<P><<<B, threads, smem, ctx->stream>>>(A);
    CKL(ctx);
    return TVL1_OK;
}

// The pipelined kernel serves every level it fits; HS_PIPELINE=0 keeps the one-sweep kernel everywhere
// (A/B measurements -- both give the bits of the sequential sweep).
bool hs_pipe_enabled()
{
    const char *e = std::getenv("HS_PIPELINE");
    return !e || std::atoi(e) != 0;
}

int hs_threads(const Level &l)
{
    // rows per thread as even as possible: ceil(ny / ceil(ny / 1024)) threads, whole warps
    const int rows_per_thread = ceil_div(l.ny, kHsMaxThreads);
    return std::min(kHsMaxThreads, round_up(ceil_div(l.ny, rows_per_thread), 32));
}

// CTAs of a kernel variant one SM holds (registers, shared memory, threads), 0 if unknown.
template <class K>
int hs_ctas_per_sm(K kernel, int threads, size_t smem)
{
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kHsSmemOneCta) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

int hs_ctas_per_sm_one(int P, int threads, size_t smem)
{
    switch (P) {
    case 0: return hs_ctas_per_sm(k_hs_sor<0, false>, threads, smem);
    case 1: return hs_ctas_per_sm(k_hs_sor<1, false>, threads, smem);
    case 2: return hs_ctas_per_sm(k_hs_sor<2, false>, threads, smem);
    default: return hs_ctas_per_sm(k_hs_sor<3, false>, threads, smem);
    }
}

int hs_ctas_per_sm_pipe(int P, int threads, size_t smem)
{
    switch (P) {
    case 0: return hs_ctas_per_sm(k_hs_sor_pipe<0>, threads, smem);
    case 1: return hs_ctas_per_sm(k_hs_sor_pipe<1>, threads, smem);
    case 2: return hs_ctas_per_sm(k_hs_sor_pipe<2>, threads, smem);
    default: return hs_ctas_per_sm(k_hs_sor_pipe<3>, threads, smem);
    }
}

// Which SOR kernel serves level l for a batch of B pairs.  Measured on 1920x1080:
// the pipelined kernel is 1.2x faster CTA for CTA, but it needs 64 registers per thread, so at 544
// threads only ONE of its CTAs fits an SM where TWO of the one-sweep kernel do -- and two resident pairs
// per SM are worth 1.47x.  So: pipelined, unless the batch needs more CTAs than the pipelined kernel can
// keep resident while the one-sweep kernel could hold more per SM.  (Both give the same bits.)
bool hs_level_pipelined(tvl1_ctx *ctx, const Level &l, int B, int prefetch);

int hs_ctas_per_sm_pairs(int P, int threads, size_t smem)
{
    switch (P) {
    case 0: return hs_ctas_per_sm(k_hs_sor_pairs<0>, threads, smem);
    case 1: return hs_ctas_per_sm(k_hs_sor_pairs<1>, threads, smem);
    case 2: return hs_ctas_per_sm(k_hs_sor_pairs<2>, threads, smem);
    default: return hs_ctas_per_sm(k_hs_sor_pairs<3>, threads, smem);
    }
}

HsKernel hs_level_kernel(tvl1_ctx *ctx, const Level &l, int B, int prefetch)
{
    if (prefetch == kHsForcePairs) return HS_PAIRS;
    if (prefetch == -1 && hs_pairs_enabled() && hs_pairs_fits(ctx->ws, l)) {
        // same occupancy rule as for the pipelined kernel: two columns per step, unless the batch needs
        // more CTAs than it can keep resident while the one-sweep kernel could hold more per SM
        const int P_one = hs_pick_prefetch(l.ny, -1);
        if (P_one < 0) return HS_PAIRS;
        const int threads = hs_threads(l), rp = round_up(l.ny, 32);
        const int P_pairs = hs_pairs_prefetch(l.ny, -1);
        const int nb_pairs = hs_ctas_per_sm_pairs(P_pairs, threads, hs_pairs_smem(P_pairs, rp));
        if (nb_pairs < 1 || B <= nb_pairs * ctx->sm_count) return HS_PAIRS;
        if (hs_ctas_per_sm_one(P_one, threads, hs_ring_bytes(P_one, rp)) <= nb_pairs) return HS_PAIRS;
    }
    return hs_level_pipelined(ctx, l, B, prefetch) ? HS_PIPELINED : HS_ONE_SWEEP;
}

bool hs_level_pipelined(tvl1_ctx *ctx, const Level &l, int B, int prefetch)
{
    if (prefetch == kHsForcePipelined) return true;
    if (prefetch != -1) return false;
    if (!hs_pipe_enabled() || !hs_pipe_fits(ctx->ws, l)) return false;
    const int P_one = hs_pick_prefetch(l.ny, -1);
    if (P_one < 0) return true;                       // the one-sweep kernel would need its rings in global memory
    const int threads = hs_threads(l), rp = round_up(l.ny, 32);
    const int P_pipe = hs_pipe_prefetch(l.ny, -1);
    const int nb_pipe = hs_ctas_per_sm_pipe(P_pipe, threads, hs_pipe_smem(P_pipe, rp));
    if (nb_pipe < 1 || B <= nb_pipe * ctx->sm_count) return true;
    const int nb_one = hs_ctas_per_sm_one(P_one, threads, hs_ring_bytes(P_one, rp));
    return nb_one <= nb_pipe;
}

// Snapshot planes (2 x float2 x L ny per pair) and error sums (D x rp doubles per pair) for the largest level.
int hs_ensure_pipe_buffers(tvl1_ctx *ctx)
{
    Workspace &w = ctx->ws;
    size_t snap = 0, part = 0;
    for (const Level &l : w.lv) {
        const int L = hs::pipe_period(l.nx);
        const int Lp = hs::pairs_period(l.nx);
        snap = std::max(snap, std::max((size_t) 4 * L * l.ny, (size_t) 8 * Lp * l.ny));
        part = std::max(part, (size_t) std::max(hs::pipe_error_depth(L, l.nx, l.ny),
                                                hs::pairs_error_depth(Lp, l.nx, l.ny)) * round_up(l.ny, 32));
    }
    if (w.hs_snap && w.hs_snap_stride >= snap && w.hs_part_stride >= part) return TVL1_OK;
    cudaFree(w.hs_snap); cudaFree(w.hs_part);
    w.hs_snap = nullptr; w.hs_part = nullptr;
    CK(cudaMalloc(&w.hs_snap, sizeof(float) * snap * w.B));
    CK(cudaMalloc(&w.hs_part, sizeof(double) * part * w.B));
    w.hs_snap_stride = snap; w.hs_part_stride = part;
    return TVL1_OK;
}

// The SOR loop of one warp step for every pair of the batch: one launch, one CTA per pair.
int hs_launch_sor(tvl1_ctx *ctx, int s, int B, const hs_params &prm, int stat_slot, HsKernel kern, int prefetch = -1)
{
    Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    int want = prefetch;
    if (want < 0) {
        want = -1;
        if (const char *e = std::getenv("HS_PREFETCH")) want = std::atoi(e);       // measurements (profiles/run_hs.py)
    }
    HsSorParams A = {};
    A.state = w.state; A.plane0 = w.plane0; A.set_stride = w.set_stride;
    A.ctl = w.ctl;
    A.nx = l.nx; A.ny = l.ny; A.rp = round_up(l.ny, 32);
    A.alpha2 = (float) (prm.alpha * prm.alpha);                             // :99
    A.tol = prm.tol; A.max_iter = prm.maxiter;
    A.stat_iters = w.stat_iters; A.stat_errs = w.stat_errs;
    A.stat_stride = w.stat_stride; A.stat_slot = stat_slot;
    A.px_iters = w.counters; A.level = std::min(s, TVL1_MAX_LEVELS - 1);
    const int threads = hs_threads(l);
    const bool pipe = kern == HS_PIPELINED;
    if (kern == HS_PAIRS) {
        if (!hs_pairs_fits(w, l)) return fail_arg(ctx, "Horn-Schunck: level does not fit the two-column kernel");
        const int P = hs_pairs_prefetch(l.ny, want);
        if (P < 0) return fail_arg(ctx, "Horn-Schunck: prefetch distance does not fit shared memory");
        TRY(hs_ensure_pipe_buffers(ctx));
        A.L = hs::pairs_period(l.nx);
        int K = 8;
        if (const char *e = std::getenv("HS_SNAP_K")) K = std::max(1, std::atoi(e));
        A.K = hs::pairs_snapshot_period(K, A.L, l.nx, l.ny);
        A.D = hs::pairs_error_depth(A.L, l.nx, l.ny);
        A.snap = w.hs_snap; A.snap_stride = w.hs_snap_stride;
        A.part = w.hs_part; A.part_stride = w.hs_part_stride;
        const size_t smem = hs_pairs_smem(P, A.rp);
        switch (P) {
        case 0: TRY(hs_launch_pairs_p<0>(ctx, A, B, threads, smem)); break;
        case 1: TRY(hs_launch_pairs_p<1>(ctx, A, B, threads, smem)); break;
        case 2: TRY(hs_launch_pairs_p<2>(ctx, A, B, threads, smem)); break;
        default: TRY(hs_launch_pairs_p<3>(ctx, A, B, threads, smem)); break;
        }
        ctx->stats.iterate_launches++;
        return TVL1_OK;
    }
    if (pipe) {
        if (!hs_pipe_fits(w, l)) return fail_arg(ctx, "Horn-Schunck: level does not fit the pipelined kernel");
        const int P = hs_pipe_prefetch(l.ny, want);
        if (P < 0) return fail_arg(ctx, "Horn-Schunck: prefetch distance does not fit shared memory");
        TRY(hs_ensure_pipe_buffers(ctx));
        A.L = hs::pipe_period(l.nx);
        int K = 8;
        if (const char *e = std::getenv("HS_SNAP_K")) K = std::max(1, std::atoi(e));
        A.K = hs::pipe_snapshot_period(K, A.L, l.nx, l.ny);
        A.D = hs::pipe_error_depth(A.L, l.nx, l.ny);
        A.snap = w.hs_snap; A.snap_stride = w.hs_snap_stride;
        A.part = w.hs_part; A.part_stride = w.hs_part_stride;
        const size_t smem = hs_pipe_smem(P, A.rp);
        switch (P) {
        case 0: TRY(hs_launch_pipe_p<0>(ctx, A, B, threads, smem)); break;
        case 1: TRY(hs_launch_pipe_p<1>(ctx, A, B, threads, smem)); break;
        case 2: TRY(hs_launch_pipe_p<2>(ctx, A, B, threads, smem)); break;
        default: TRY(hs_launch_pipe_p<3>(ctx, A, B, threads, smem)); break;
        }
        ctx->stats.iterate_launches++;
        return TVL1_OK;
    }
    int P = prefetch == kHsForceGlobalRing ? -1 : hs_pick_prefetch(l.ny, want);
    const bool global_ring = P < 0;
    if (global_ring) {
        if (want >= 0) return fail_arg(ctx, "Horn-Schunck: prefetch distance does not fit shared memory");
        if (!hs_global_ring_fits(w, l)) return fail_arg(ctx, "Horn-Schunck: no room for the rings in global memory");
        P = 0;
    }
    const size_t smem = hs_ring_bytes(P, A.rp);
    if (global_ring) TRY((hs_launch_sor_p<0, true>(ctx, A, B, threads, 0)));
    else switch (P) {
    case 0: TRY((hs_launch_sor_p<0, false>(ctx, A, B, threads, smem))); break;
    case 1: TRY((hs_launch_sor_p<1, false>(ctx, A, B, threads, smem))); break;
    case 2: TRY((hs_launch_sor_p<2, false>(ctx, A, B, threads, smem))); break;
    default: TRY((hs_launch_sor_p<3, false>(ctx, A, B, threads, smem))); break;
    }
    ctx->stats.iterate_launches++;
    return TVL1_OK;
}

// wave planes with period L (and the initial snapshot) for k_hs_sor_pipe, pair planes for k_hs_sor_pairs,
// else period nx
int hs_launch_to_wave(tvl1_ctx *ctx, int s, int B, HsKernel kern)
{
    const bool pipe = kern == HS_PIPELINED;
    if (kern != HS_ONE_SWEEP) TRY(hs_ensure_pipe_buffers(ctx));
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    if (kern == HS_PAIRS) {
        const int L = hs::pairs_period(l.nx);
        dim3 g(ceil_div(L, 32), ceil_div(l.ny, 32), B);
        k_hs_to_wave_pairs<<<g, dim3(32, 8), 0, ctx->stream>>>(w.state, w.consts, w.plane0, w.field_stride,
                                                               w.set_stride, w.ctl, l, L, w.hs_snap, w.hs_snap_stride);
        CKL(ctx);
        return TVL1_OK;
    }
    const int mod = pipe ? hs::pipe_period(l.nx) : l.nx;
    dim3 g(ceil_div(mod, 32), ceil_div(l.ny, 32), B);
    k_hs_to_wave<<<g, dim3(32, 8), 0, ctx->stream>>>(w.state, w.consts, w.plane0, w.field_stride, w.set_stride,
                                                     w.ctl, l, mod, pipe ? w.hs_snap : nullptr, w.hs_snap_stride);
    CKL(ctx);
    return TVL1_OK;
}

int hs_launch_from_wave(tvl1_ctx *ctx, int s, int B, HsKernel kern)
{
    const bool pipe = kern == HS_PIPELINED;
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    if (kern == HS_PAIRS) {
        const int L = hs::pairs_period(l.nx);
        dim3 g(ceil_div(L, 32), ceil_div(l.ny, 32), B);
        k_hs_from_wave_pairs<<<g, dim3(32, 8), 0, ctx->stream>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                                                 l, L);
        CKL(ctx);
        return TVL1_OK;
    }
    const int mod = pipe ? hs::pipe_period(l.nx) : l.nx;
    dim3 g(ceil_div(mod, 32), ceil_div(l.ny, 32), B);
    k_hs_from_wave<<<g, dim3(32, 8), 0, ctx->stream>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, l,
                                                       mod);
    CKL(ctx);
    return TVL1_OK;
}

// One pyramid level: horn_schunck_optical_flow, src/horn_schunck_pyramidal.cpp:78-249, for every pair.
int hs_run_level(tvl1_ctx *ctx, int s, int B, const hs_params &prm, int stat_base)
{
    for (int wi = 0; wi < prm.warps; wi++) {                                // :117
        {
            Span sp(ctx, 1);
            TRY(launch_warp(ctx, s, B));                                    // :114, :123-125, dif of :130
        }
        Span sp(ctx, 0, std::min(s, TVL1_MAX_LEVELS - 1));
        const HsKernel kern = hs_level_kernel(ctx, ctx->ws.lv[s], B, -1);
        TRY(hs_launch_to_wave(ctx, s, B, kern));
        TRY(hs_launch_sor(ctx, s, B, prm, stat_base + wi, kern));           // :127-137 (on the fly), :139-231
        TRY(hs_launch_from_wave(ctx, s, B, kern));
    }
    return TVL1_OK;
}

tvl1_params hs_as_tvl1(const hs_params &prm, bool multiscale)
{
    // the pyramid builder and the argument checks only look at nscales, zfactor and warps
    tvl1_params t{ 0.25, 0.15, 0.3, multiscale ? prm.nscales : 1, multiscale ? prm.zfactor : 0.5, prm.warps, prm.tol };
    return t;
}

int hs_check_params(tvl1_ctx *ctx, const hs_params *prm)
{
    if (!ctx) return TVL1_ERR_ARG;
    if (!prm) return fail_arg(ctx, "null pointer argument");
    if (!(prm->alpha > 0.0)) return fail_arg(ctx, "alpha must be positive");
    if (prm->maxiter < 1) return fail_arg(ctx, "maxiter must be >= 1");
    return TVL1_OK;
}

// horn_schunck_pyramidal for B <= max_batch pairs, device-resident dense inputs / outputs.
int run_hs_multiscale(tvl1_ctx *ctx, int B, const float *dI1, const float *dI2, float *du, float *dv, int nx,
                      int ny, const hs_params &prm, int *iters_out, double *errs_out)
{
    const int ns = prm.nscales;
    const int nstat = ns * prm.warps;
    const tvl1_params tp = hs_as_tvl1(prm, true);
    TRY(ensure_workspace(ctx, nx, ny, ns, prm.zfactor, B, nstat));
    Workspace &w = ctx->ws;
    for (int s = 0; s < ns; s++) TRY(hs_check_level(ctx, w.lv[s]));
    cudaStream_t st = ctx->stream;

    Span total(ctx, 2);
    TRY(build_pyramid(ctx, B, dI1, dI2, nx, ny, tp));                        // :293-317
    TRY(launch_zero(ctx, ns - 1, B, F_U1, 2));                              // :320-323
    for (int s = ns - 1; s >= 0; s--) {                                     // :326-353
        TRY(hs_run_level(ctx, s, B, prm, (ns - 1 - s) * prm.warps));
        if (!s) break;
        const Level &c = w.lv[s], &f = w.lv[s - 1];
        Span zs(ctx, 4);
        dim3 g(ceil_div(f.nx, kZiTW), ceil_div(f.ny, kZiTH), B);
        k_zoom_in_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, c, f,
                                                  (double) f.nx / c.nx, (double) f.ny / c.ny,
                                                  (float) (1.0 / prm.zfactor), 0, f.ny);   // :345-352
        CKL(ctx);
        k_flip_cur<<<ceil_div(B, 128), 128, 0, st>>>(w.ctl, B);
        CKL(ctx);
    }
    {
        Span ex(ctx, 5);
        dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), B);
        k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, w.lv[0],
                                                 du, dv);
        CKL(ctx);
    }
    total.end();
    TRY(fetch_stats(ctx, B, nstat, iters_out, errs_out));
    return TVL1_OK;
}

// horn_schunck_optical_flow (one level, no normalisation / blur), device-resident dense buffers.
int run_hs_single_scale(tvl1_ctx *ctx, int B, const float *dI1, const float *dI2, float *du, float *dv, int nx,
                        int ny, const hs_params &prm, int *iters_out, double *errs_out)
{
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, B, prm.warps));
    Workspace &w = ctx->ws;
    TRY(hs_check_level(ctx, w.lv[0]));
    cudaStream_t st = ctx->stream;
    Span total(ctx, 2);
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    k_init_ctl<<<ceil_div(B, 128), 128, 0, st>>>(w.ctl, w.mm, B);
    CKL(ctx);
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), B);
    k_pack<<<g, dim3(32, 8), 0, st>>>(dI1, w.I0(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_pack<<<g, dim3(32, 8), 0, st>>>(dI2, w.I1(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_import_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, w.lv[0], du, dv);
    CKL(ctx);
    TRY(hs_run_level(ctx, 0, B, prm, 0));
    k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, w.lv[0], du, dv);
    CKL(ctx);
    total.end();
    TRY(fetch_stats(ctx, B, prm.warps, iters_out, errs_out));
    return TVL1_OK;
}

template <typename T>
int hs_solve_host(tvl1_ctx *ctx, int npairs, const T *I1, const T *I2, T *u, T *v, int nx, int ny,
                  const hs_params *prm, int *iters_out, double *errs_out, bool multiscale)
{
    TRY(hs_check_params(ctx, prm));
    const tvl1_params tp = hs_as_tvl1(*prm, multiscale);
    ctx->hs_mode = true;
    ctx->hs = *prm;
    const int rc = solve_host<T>(ctx, npairs, I1, I2, u, v, nx, ny, &tp, iters_out, errs_out, multiscale);
    ctx->hs_mode = false;
    return rc;
}

} // namespace

extern "C" {

void hs_default_params(hs_params *p)
{
    if (!p) return;
    p->alpha = 7; p->nscales = 10; p->zfactor = 0.5; p->warps = 10; p->tol = 0.0001; p->maxiter = 150;
}

int hs_clamp_nscales(int nx, int ny, int nscales, double zfactor)
{
    const double N = 1 + std::log(std::hypot((double) nx, (double) ny) / 16) / std::log(1 / zfactor);
    if (N < nscales) nscales = (int) N;
    return nscales;
}

int hs_solve_f32(tvl1_ctx *ctx, const float *I1, const float *I2, float *u, float *v, int nx, int ny,
                 const hs_params *prm, int *iters_out, double *errs_out)
{
    return hs_solve_host<float>(ctx, 1, I1, I2, u, v, nx, ny, prm, iters_out, errs_out, true);
}

int hs_solve_f64(tvl1_ctx *ctx, const double *I1, const double *I2, double *u, double *v, int nx, int ny,
                 const hs_params *prm, int *iters_out, double *errs_out)
{
    return hs_solve_host<double>(ctx, 1, I1, I2, u, v, nx, ny, prm, iters_out, errs_out, true);
}

int hs_solve_batch_f32(tvl1_ctx *ctx, int npairs, const float *I1, const float *I2, float *u, float *v, int nx,
                       int ny, const hs_params *prm, int *iters_out, double *errs_out)
{
    return hs_solve_host<float>(ctx, npairs, I1, I2, u, v, nx, ny, prm, iters_out, errs_out, true);
}

int hs_solve_batch_dev_f32(tvl1_ctx *ctx, int npairs, const float *dI1, const float *dI2, float *du, float *dv,
                           int nx, int ny, const hs_params *prm, int *iters_out, double *errs_out)
{
    TRY(hs_check_params(ctx, prm));
    const tvl1_params tp = hs_as_tvl1(*prm, true);
    TRY(check_common(ctx, dI1, dI2, du, dv, nx, ny, &tp, true));
    if (npairs < 1) return fail_arg(ctx, "npairs must be >= 1");
    reset_stats(ctx);
    const hs_params hp = *prm;
    const size_t n = (size_t) nx * ny;
    const int nstat = hp.nscales * hp.warps;
    const int Bmax = std::min(npairs, ctx->max_batch);
    return run_lanes(ctx, ceil_div(npairs, Bmax), ctx->dev_lanes, [&](tvl1_ctx *c, int k) -> int {
        const int first = k * Bmax, B = std::min(Bmax, npairs - first);
        const size_t off = (size_t) first * n;
        return run_hs_multiscale(c, B, dI1 + off, dI2 + off, du + off, dv + off, nx, ny, hp,
                                 iters_out ? iters_out + (size_t) first * nstat : nullptr,
                                 errs_out ? errs_out + (size_t) first * nstat : nullptr);
    });
}

int hs_single_scale_f32(tvl1_ctx *ctx, const float *I1, const float *I2, float *u, float *v, int nx, int ny,
                        const hs_params *prm, int *iters_out, double *errs_out)
{
    return hs_solve_host<float>(ctx, 1, I1, I2, u, v, nx, ny, prm, iters_out, errs_out, false);
}

int hs_single_scale_f64(tvl1_ctx *ctx, const double *I1, const double *I2, double *u, double *v, int nx, int ny,
                        const hs_params *prm, int *iters_out, double *errs_out)
{
    return hs_solve_host<double>(ctx, 1, I1, I2, u, v, nx, ny, prm, iters_out, errs_out, false);
}

int hs_sor_f32(tvl1_ctx *ctx, const float *I2wx, const float *I2wy, const float *rho_c, float *u, float *v,
               int nx, int ny, double alpha, double tol, int maxiter, int prefetch, int *niter_out,
               double *err_out)
{
    if (!ctx) return TVL1_ERR_ARG;
    if (!I2wx || !I2wy || !rho_c || !u || !v) return fail_arg(ctx, "null pointer argument");
    if (!(alpha > 0.0) || maxiter < 1) return fail_arg(ctx, "alpha must be positive and maxiter >= 1");
    if (prefetch > hs::kMaxPrefetch || prefetch < kHsForcePairs) return fail_arg(ctx, "prefetch must be -4 .. 3");
    CK(cudaSetDevice(ctx->device));
    reset_stats(ctx);
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, 1, 1));
    Workspace &w = ctx->ws;
    TRY(hs_check_level(ctx, w.lv[0]));
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t) nx * ny;
    Dev d(ctx->stream);
    float *buf = d.alloc(5 * n);
    if (!buf) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    const float *src[5] = { u, v, I2wx, I2wy, rho_c };
    for (int k = 0; k < 5; k++) CK(cudaMemcpyAsync(buf + k * n, src[k], n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    k_init_ctl<<<1, 32, 0, st>>>(w.ctl, w.mm, 1);
    CKL(ctx);
    const dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
    // flow -> live set; I2wx, I2wy, rho_c -> the constant planes k_warp would have written
    k_import_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, w.lv[0], buf,
                                             buf + n);
    CKL(ctx);
    const int fields[3] = { C_IX, C_IY, C_RHO };
    for (int k = 0; k < 3; k++) {
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf + (2 + k) * n, w.consts + (size_t) fields[k] * w.field_stride, nx, ny,
                                          w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    hs_params prm{ alpha, 1, 0.5, 1, tol, maxiter };
    const HsKernel kern = hs_level_kernel(ctx, w.lv[0], 1, prefetch);
    if (kern == HS_PIPELINED && !hs_pipe_fits(w, w.lv[0]))
        return fail_arg(ctx, "Horn-Schunck: level does not fit the pipelined kernel");
    if (kern == HS_PAIRS && !hs_pairs_fits(w, w.lv[0]))
        return fail_arg(ctx, "Horn-Schunck: level does not fit the two-column kernel");
    TRY(hs_launch_to_wave(ctx, 0, 1, kern));
    TRY(hs_launch_sor(ctx, 0, 1, prm, 0, kern, prefetch));
    TRY(hs_launch_from_wave(ctx, 0, 1, kern));
    k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl, w.lv[0], buf,
                                             buf + n);
    CKL(ctx);
    CK(cudaMemcpyAsync(u, buf, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(v, buf + n, n * 4, cudaMemcpyDeviceToHost, st));
    int it = 0;
    double er = 0;
    TRY(fetch_stats(ctx, 1, 1, &it, &er));
    if (niter_out) *niter_out = it;
    if (err_out) *err_out = er;
    return TVL1_OK;
}

} // extern "C"
