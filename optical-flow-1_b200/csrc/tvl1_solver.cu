// Host side of the C ABI declared in include/tvl1_b200.h: device memory, the coarse-to-fine
// driver (src/tvl1flow.cpp:219-328) and the per-level driver (src/tvl1flow.cpp:46-212) of the
// reference, re-designed for a GPU: batches of independent frame pairs advance in lock-step, one
// launch per primal-dual iteration for the whole batch, with the stopping rule evaluated on the
// device per pair.  The host only decides how many launches to enqueue before it looks at the
// "pairs still iterating" counter again.
#include "../../include/tvl1_b200.h"
#include "../../include/hs_b200.h"
#include "tvl1_kernels.cuh"
#include "hs_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

using namespace tvl1;

namespace {

thread_local std::string g_create_error;

constexpr int kIterR = 8;      // rows per warp strip in k_iterate_t1
constexpr int kIterWY = 4;     // warps per CTA

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

struct Workspace {
    int nx = 0, ny = 0, nscales = 0, B = 0;
    double zfactor = 0;
    std::vector<Level> lv;
    std::vector<size_t> pyr_off;     // float offset of level s: [I0 x B][I1 x B] planes of pitch*ny
    size_t plane0 = 0, field_stride = 0, set_stride = 0;
    float *pyr = nullptr, *state = nullptr, *consts = nullptr, *tmp = nullptr;
    PairCtl *ctl = nullptr;
    unsigned int *mm = nullptr;
    double *partials = nullptr;
    double *tb_partials = nullptr;            // [B][tb_parts][kTbT] (temporally blocked kernel)
    int tb_parts = 0;
    LoopCtl *loop = nullptr;
    int *stat_iters = nullptr;
    double *stat_errs = nullptr;
    unsigned long long *counters = nullptr;   // [level] pixel-iterations, [16 + level] iteration launches
    int parts_per_pair = 0;
    int stat_stride = 0;
    size_t bytes = 0;
    // cluster-resident iteration kernel: cluster size per level (0 = streaming kernel) and band height
    std::vector<int> res_cluster, res_rows;
    int resident_key = -1;
    int row_pad = 1;
    // pipelined Horn-Schunck kernel (hs_solver.cuh): snapshot planes and per-row error sums, allocated on first use
    float *hs_snap = nullptr;
    double *hs_part = nullptr;
    size_t hs_snap_stride = 0, hs_part_stride = 0;

    size_t plane(int s) const { return (size_t) lv[s].pitch * lv[s].ny; }
    float *I0(int s) const { return pyr + pyr_off[s]; }
    float *I1(int s) const { return pyr + pyr_off[s] + (size_t) B * plane(s); }
};

struct EventPair { cudaEvent_t a, b; int kind, level; };   // kind 0 iterate, 1 warp, 2 total, 3 pyramid, 4 zoom_in, 5 export

// The coarse-to-fine part of one solve, captured once per (workspace, parameters) as a CUDA graph
// whose primal-dual loops are conditional WHILE nodes: the device ends each loop itself
// (cudaGraphSetConditional from the iteration kernel), so a solve needs no host round trip.
struct SolveGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    tvl1_params prm{};
    bool multiscale = false, profiling = false;
    long long variant = 0;
    unsigned long long static_launches = 0;   // kernel nodes outside the while bodies
    unsigned long long pixel_warps = 0;
    std::vector<EventPair> events;            // external event-record nodes (profiling)
};

} // namespace

struct tvl1_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool profiling = false;
    int max_batch = 32;
    tvl1_stats stats{};
    Workspace ws;
    static constexpr int kAltWs = 4;
    Workspace ws_alt[kAltWs];                // same image shape, other batch sizes (ramped / ragged chunks of a host-buffer batch)
    int alt_evict = 0;                       // round-robin victim when every alternate is in use
    LoopCtl *h_loop = nullptr;               // pinned
    cudaEvent_t sync_event = nullptr;        // blocking-sync event: lane threads sleep instead of spinning
    bool blocking_wait = false;              // set while several lanes share the GPU
    cudaStream_t body_stream = nullptr;      // capture stream for while-node bodies
    bool use_graph = true;                   // TVL1_NO_GRAPH=1 selects the host-driven loop
    bool use_resident = true;                // TVL1_NO_RESIDENT=1 keeps every level on the streaming kernel
    bool use_tb = true;                      // TVL1_NO_TB=1: never use the temporally blocked kernel
    // ... which pays on the levels whose loops are long enough for two-iteration blocks (measured on the 256-pair
    // batch at default epsilon: 4.6 iterations per warp step on level 1, 32.6 -> 29.5 ms; 1.9 on level 0, 51.2 -> 54.1 ms).
    // Per level, from the iteration counts of the context's previous solve, with hysteresis; until there is a previous
    // solve: every level but the finest.  The flow does not depend on the choice (same bits from every kernel).
    unsigned int t2_levels = ~1u;            // bit s: level s may use it
    bool t2_adapt = true;                    // TVL1_T2_ADAPT=0: keep the initial mask (TVL1_T2_LEVELS=<mask>)
    bool t2_when_shared = true;              // TVL1_T2_SHARED=0: not while lanes share the GPU
    bool t2_stage_when_shared = true;        // TVL1_T2_STAGE_SHARED=0: direct loads while lanes share the GPU
    bool t2_stage = true;                    // TVL1_T2_STAGE=0: k_iterate_t2 loads its rows straight into registers (A/B)
    bool t2_first = true;                    // TVL1_T2_FIRST=0: a streamed level's first launch is one iteration (k_iterate_t1)
    int use_t2 = 1;                          // TVL1_T2=0: never use the two-iterations-per-launch marching kernel; 2: wherever
                                             // the shared-memory kernel is not used, however small the launch (tests)
    bool zero_in_first = true;               // TVL1_ZERO_PASS=1: zero the duals of a streamed level with a pass of their own (A/B)
    bool gauss_shfl = true;                  // TVL1_GAUSS_SHFL=0: the marching blur that loads its own row inputs (A/B)
    bool warp_tma = true;                    // TVL1_WARP_TMA=0: stage the warp kernel's box with cp.async only
    int slot_ctas = 32768;                   // CTAs a full iteration launch should have at least (TVL1_SLOT_CTAS)
    int tail_pairs = 16;                     // lock-step batches: once this few pairs still iterate, the loop goes on
                                             // with narrow launches of tail_slot_ctas CTAs (TVL1_TAIL_PAIRS, 0 = off)
    int tail_pairs_shared = 0;               // ... the same for a chunk solved while other lanes share the GPU (TVL1_TAIL_PAIRS_SHARED)
    int tail_slot_ctas = 2048;               // (TVL1_TAIL_SLOT_CTAS)
    bool tail_tb = true;                     // temporal blocking in the tail even where the full batch runs without (TVL1_TAIL_TB)
    long long tb_max_pixels = 192ll << 20;   // ... which serves lock-step batches up to this many pixels per level
    int force_cluster = 0;                   // tests: force this cluster size where it fits
    bool shared_gpu = false;                 // set while several lanes run chunks concurrently on this GPU
    bool tb_when_shared = false;             // TVL1_TB_SHARED=1: temporal blocking also then (A/B)
    bool capturing = false;
    SolveGraph sg;
    SolveGraph sg_alt[kAltWs];               // solve graphs of ws_alt[]
    SolveGraph *cap = nullptr;               // graph being captured (profiling events attach to it)
    SolveGraph level_sg[TVL1_MAX_LEVELS];    // row-band mode: one graph per pyramid level
    // row-band mode (one image over several GPUs)
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int band_rank = 0, band_world = 1;
    double *d_band_sum = nullptr;
    double *d_agree = nullptr;
    // ... over peer memory (CUDA IPC + NVLink): halos and error sums move inside the iteration kernel
    bool p2p_ready = false;
    bool band_use_nccl_per_iteration = false;   // TVL1_BAND_NCCL=1: the NCCL send/recv variant
    BandMailbox *my_box = nullptr;
    BandMailbox *boxes[kMaxRanks] = {};
    float *peer_state[kMaxRanks] = {};             // every rank's state buffer ([band_rank] = our own)
    unsigned int *d_gather_ticket = nullptr;       // arrival counter of k_band_allgather
    const float *p2p_state_key = nullptr;
    unsigned char *d_handles = nullptr;            // [world][64] scratch for the handle all-gather
    static constexpr int kMaxLanes = 8;
    tvl1_ctx *sib[kMaxLanes - 1] = {};   // sibling contexts (extra lanes, same GPU)
    int host_lanes = 4;                      // lanes used by the host-buffer batch entry points
    int short_div = 2;                       // host-buffer batches: first / last chunks of max_batch / short_div pairs
                                             // (TVL1_SHORT_DIV, 0 or 1 = all chunks equal)
    int dev_lanes = 2;                       // lanes used by the device-buffer batch entry point
    bool is_sibling = false;
    tvl1_ctx *band_ctx = nullptr;            // private context of the row-band mode
    std::vector<cudaEvent_t> ev_pool;
    std::vector<EventPair> ev_used;
    // staging for the host-buffer entry points
    void *stage_in[2] = { nullptr, nullptr };
    void *stage_out[2] = { nullptr, nullptr };
    float *stage_f32[4] = { nullptr, nullptr, nullptr, nullptr };
    size_t stage_bytes = 0, stage_f32_bytes = 0;
    // ... pinned host buffers: a call-wide pipeline (solve_host_pipelined).  Chunks are uploaded in order on ONE copy
    // stream into a ring of device slots, solved by the lanes, downloaded in order on a second copy stream.
    struct HostSlot {
        void *in[2] = { nullptr, nullptr };        // uploaded inputs (T), I0 / I1 stacks or one stack of frames
        void *out[2] = { nullptr, nullptr };       // results (T) waiting for their download
        cudaEvent_t up = nullptr, done = nullptr, down = nullptr;  // upload complete / solved (results in `out`) / download complete
        bool down_recorded = false;
    };
    static constexpr int kMaxSlots = kMaxLanes + 3;
    HostSlot slots[kMaxSlots];
    size_t slot_in_bytes = 0, slot_out_bytes = 0;
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    bool host_pipe = false;                        // TVL1_HOST_PIPE=1: the call-wide pipeline for pinned buffers (faster, but see DESIGN 3.6:
                                                   // an intermittent launch failure was seen with it and is not understood yet)
    int ramp_min_chunk = 8;                        // smallest ramp step of its chunk schedule (TVL1_MIN_CHUNK; kRampMinChunk)
    int pipe_lanes = 3;                            // lanes of the pipeline (they only solve: three measured best, TVL1_PIPE_LANES)
    std::vector<int> chunk_override;               // TVL1_CHUNKS=8,16,...: explicit chunk sizes (experiments)
    void *pipe_buf[2] = { nullptr, nullptr };      // pinned staging ring for pageable host buffers
    cudaEvent_t pipe_ev[2] = { nullptr, nullptr };
    int sm_count = 148;
    // Horn-Schunck entry points (include/hs_b200.h) reuse the host-buffer drivers below: while hs_mode
    // is set, a chunk is solved by run_hs_multiscale / run_hs_single_scale with these parameters
    bool hs_mode = false;
    hs_params hs{};
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
            return TVL1_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define CKL(ctx_) CK(cudaGetLastError()); (ctx_)->stats.kernel_launches++

#define TRY(expr)                                                                                  \
    do {                                                                                           \
        int rc_ = (expr);                                                                          \
        if (rc_ != TVL1_OK) return rc_;                                                            \
    } while (0)

// Wait for the context's stream without spinning a host core (several lanes per GPU and several
// ranks per box share the host CPUs).  Loop-control round trips keep using cudaStreamSynchronize.
cudaError_t sleep_until_done(tvl1_ctx *ctx)
{
    if (!ctx->blocking_wait) return cudaStreamSynchronize(ctx->stream);   // single lane: lowest latency
    cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->sync_event);
}

int fail_arg(tvl1_ctx *ctx, const char *msg)
{
    if (ctx) ctx->err = msg;
    return TVL1_ERR_ARG;
}

// 1-D Gaussian kernel of src/operators.cpp:515-539 (fp64 on the host, stored as fp32 taps)
int make_taps(double sigma, GaussTaps &t)
{
    const double den = 2 * sigma * sigma;
    const int size = (int) (5 * sigma) + 1;
    if (size > kMaxTaps || size < 1) return -1;
    double B[kMaxTaps];
    for (int i = 0; i < size; i++) B[i] = 1 / (sigma * std::sqrt(2.0 * 3.1415926)) * std::exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += B[i];
    norm *= 2;
    norm -= B[0];
    t.size = size;
    for (int i = 0; i < kMaxTaps; i++) t.w[i] = i < size ? (float) (B[i] / norm) : 0.f;
    return size;
}

void free_graph(SolveGraph &g, std::vector<cudaEvent_t> &pool)
{
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    for (auto &p : g.events) { pool.push_back(p.a); pool.push_back(p.b); }
    g = SolveGraph();
}

void free_workspace(Workspace &w)
{
    cudaFree(w.pyr); cudaFree(w.state); cudaFree(w.consts); cudaFree(w.tmp); cudaFree(w.ctl);
    cudaFree(w.mm); cudaFree(w.partials); cudaFree(w.loop); cudaFree(w.tb_partials); cudaFree(w.stat_iters);
    cudaFree(w.stat_errs); cudaFree(w.counters); cudaFree(w.hs_snap); cudaFree(w.hs_part);
    w = Workspace();
}

int iterate_parts(const Level &l)
{
    return ceil_div(l.nx, 124) * ceil_div(l.ny, 4 * kIterWY);     // smallest strip height launch_iterate uses
}

// Smallest cluster (1,2,4,8,16 CTAs) whose row bands fit one SM each: <= 2048 float4 groups per CTA
// (512 threads x 4) and the six state planes of the band in <= 227 KB of shared memory.  Returns 0
// when the level has to stream through HBM instead.
int pick_cluster(tvl1_ctx *ctx, const Level &l, int B, int *rows_out)
{
    if (!ctx->use_resident) return 0;
    static bool attr_done[64] = { false };
    int best = 0;
    for (int C = 1; C <= kResMaxCluster; C *= 2) {
        if (ctx->force_cluster && C != ctx->force_cluster) continue;
        // a small batch cannot fill the GPU with minimal clusters: keep growing the cluster (shorter
        // bands, shorter iterations) while all clusters of the batch still run concurrently -- but
        // only while a CTA still has real work: below ~2k pixels the two cluster barriers per
        // iteration cost more than the band saves (a 1-CTA "cluster" only needs __syncthreads)
        if (best && !ctx->force_cluster &&
            (B * C > ctx->sm_count || (long long) *rows_out * l.pitch <= 2048)) break;
        const int RB = ceil_div(l.ny, C);
        if ((C - 1) * RB >= l.ny) continue;                        // every CTA needs at least one row
        if (RB * (l.pitch / 4) > kResThreads * kResQuads) continue;
        const size_t smem = resident_smem_bytes(l.pitch, RB);
        if (smem > kResSmemLimit) continue;
        if (!attr_done[ctx->device & 63]) {
            if (cudaFuncSetAttribute(k_iterate_resident, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int) kResSmemLimit) != cudaSuccess ||
                cudaFuncSetAttribute(k_iterate_resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                cudaGetLastError();
                return 0;
            }
            attr_done[ctx->device & 63] = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, 1, 1);
        cfg.blockDim = dim3(kResThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, k_iterate_resident, &cfg) != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            continue;
        }
        *rows_out = RB;
        best = C;
        if (ctx->force_cluster) break;
    }
    return best;
}

// what the cluster choice of a workspace depends on besides the level sizes and the batch size
int resident_key_of(const tvl1_ctx *ctx)
{
    return ctx->use_resident ? 1 + ctx->force_cluster : 0;
}

bool workspace_matches(const tvl1_ctx *ctx, const Workspace &w, int nx, int ny, int nscales, double zfactor, int B,
                       int stat_stride, int row_pad)
{
    return w.nx == nx && w.ny == ny && w.nscales == nscales && w.zfactor == zfactor && w.B == B &&
           w.stat_stride >= stat_stride && w.resident_key == resident_key_of(ctx) &&
           w.row_pad == row_pad;
}

int ensure_workspace(tvl1_ctx *ctx, int nx, int ny, int nscales, double zfactor, int B, int stat_stride,
                     int row_pad = 1)
{
    Workspace &w = ctx->ws;
    if (workspace_matches(ctx, w, nx, ny, nscales, zfactor, B, stat_stride, row_pad)) return TVL1_OK;
    // A host-buffer batch is cut into chunks of a few different sizes (short first and last chunks, full
    // ones in between, a ragged remainder: chunk_schedule), call after call: the displaced workspaces and
    // their solve graphs are kept as alternates and swapped back in, so none is rebuilt.  Only the batch size
    // may differ between the workspaces a context holds; another image shape drops the alternates.
    if (!ctx->nccl_comm) {
        for (int i = 0; i < tvl1_ctx::kAltWs; i++)
            if (workspace_matches(ctx, ctx->ws_alt[i], nx, ny, nscales, zfactor, B, stat_stride, row_pad)) {
                std::swap(ctx->ws, ctx->ws_alt[i]);
                std::swap(ctx->sg, ctx->sg_alt[i]);
                return TVL1_OK;
            }
        const bool same_shape = w.state && w.nx == nx && w.ny == ny && w.nscales == nscales && w.zfactor == zfactor &&
                                w.row_pad == row_pad && w.B != B;
        if (!same_shape) {
            for (int i = 0; i < tvl1_ctx::kAltWs; i++) {
                free_graph(ctx->sg_alt[i], ctx->ev_pool);
                free_workspace(ctx->ws_alt[i]);
            }
        } else {
            int slot = -1;
            for (int i = 0; i < tvl1_ctx::kAltWs && slot < 0; i++)
                if (!ctx->ws_alt[i].state) slot = i;
            if (slot < 0) {
                slot = ctx->alt_evict;
                ctx->alt_evict = (ctx->alt_evict + 1) % tvl1_ctx::kAltWs;
                free_graph(ctx->sg_alt[slot], ctx->ev_pool);
                free_workspace(ctx->ws_alt[slot]);
            }
            std::swap(ctx->ws, ctx->ws_alt[slot]);
            std::swap(ctx->sg, ctx->sg_alt[slot]);
        }
    }
    free_graph(ctx->sg, ctx->ev_pool);
    for (auto &g : ctx->level_sg) free_graph(g, ctx->ev_pool);
    free_workspace(w);
    w.nx = nx; w.ny = ny; w.nscales = nscales; w.zfactor = zfactor; w.B = B;
    w.stat_stride = stat_stride;
    w.resident_key = resident_key_of(ctx);
    w.row_pad = row_pad;
    w.lv.resize(nscales);
    w.pyr_off.resize(nscales);
    w.res_cluster.assign(nscales, 0);
    w.res_rows.assign(nscales, 0);
    int cx = nx, cy = ny;
    size_t off = 0;
    int parts = 0;
    for (int s = 0; s < nscales; s++) {
        if (s > 0) {
            int nxx, nyy;
            tvl1_zoom_size(cx, cy, &nxx, &nyy, zfactor);
            cx = nxx; cy = nyy;
        }
        if (cx < 1 || cy < 1) return fail_arg(ctx, "pyramid level has zero size (nscales too large)");
        w.lv[s] = Level{ cx, cy, round_up(cx, 4) };
        w.pyr_off[s] = off;
        off += 2 * (size_t) B * w.plane(s);
        parts = std::max(parts, iterate_parts(w.lv[s]));
        w.tb_parts = std::max(w.tb_parts, ceil_div(cx, kTbW) * ceil_div(cy, kTbH));
        w.tb_parts = std::max(w.tb_parts, ceil_div(cx, kT2W) * ceil_div(cy, 16 * kIterWY));      // strips of k_iterate_t2
        w.res_cluster[s] = pick_cluster(ctx, w.lv[s], B, &w.res_rows[s]);
    }
    // row-band mode gathers equal-sized bands in place: round the rows of a plane up to a multiple of the rank count
    w.plane0 = (size_t) w.lv[0].pitch * round_up(w.lv[0].ny, row_pad);
    w.field_stride = (size_t) B * w.plane0;
    w.set_stride = (size_t) F_COUNT * w.field_stride;
    w.parts_per_pair = parts;
    const size_t fl = sizeof(float);
    CK(cudaMalloc(&w.pyr, off * fl));
    CK(cudaMalloc(&w.state, 2 * w.set_stride * fl));
    CK(cudaMalloc(&w.consts, (size_t) C_COUNT * w.field_stride * fl));
    if (zfactor != 0.5 && nscales > 1) CK(cudaMalloc(&w.tmp, 2 * w.field_stride * fl));
    CK(cudaMalloc(&w.ctl, sizeof(PairCtl) * B));
    CK(cudaMalloc(&w.mm, sizeof(unsigned int) * 2 * B));
    CK(cudaMalloc(&w.partials, sizeof(double) * (size_t) B * parts));
    CK(cudaMalloc(&w.tb_partials, sizeof(double) * (size_t) B * w.tb_parts * kTbT));
    CK(cudaMalloc(&w.loop, sizeof(LoopCtl)));
    CK(cudaMalloc(&w.stat_iters, sizeof(int) * (size_t) B * stat_stride));
    CK(cudaMalloc(&w.stat_errs, sizeof(double) * (size_t) B * stat_stride));
    CK(cudaMalloc(&w.counters, sizeof(unsigned long long) * kCounterWords));
    w.bytes = (off + 2 * w.set_stride + C_COUNT * w.field_stride) * fl;
    // padding columns are never consumed, but keep them finite
    CK(cudaMemsetAsync(w.state, 0, 2 * w.set_stride * fl, ctx->stream));
    CK(cudaMemsetAsync(w.consts, 0, (size_t) C_COUNT * w.field_stride * fl, ctx->stream));
    CK(cudaMemsetAsync(w.pyr, 0, off * fl, ctx->stream));
    return TVL1_OK;
}

// ---- profiling events ---------------------------------------------------------------------------
cudaEvent_t take_event(tvl1_ctx *ctx)
{
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct Span {
    tvl1_ctx *ctx; int idx = -1;
    Span(tvl1_ctx *c, int kind, int level = 0) : ctx(c)
    {
        if (!c->profiling) return;
        EventPair p{ take_event(c), take_event(c), kind, level };
        if (c->capturing) {
            cudaEventRecordWithFlags(p.a, c->stream, cudaEventRecordExternal);
            c->cap->events.push_back(p);
            idx = (int) c->cap->events.size() - 1;
        } else {
            cudaEventRecord(p.a, c->stream);
            c->ev_used.push_back(p);
            idx = (int) c->ev_used.size() - 1;
        }
    }
    void end()
    {
        if (idx < 0) return;
        if (ctx->capturing) cudaEventRecordWithFlags(ctx->cap->events[idx].b, ctx->stream, cudaEventRecordExternal);
        else cudaEventRecord(ctx->ev_used[idx].b, ctx->stream);
        idx = -1;
    }
    ~Span() { end(); }
};

void add_span_time(tvl1_ctx *ctx, const EventPair &p)
{
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) != cudaSuccess || cudaEventElapsedTime(&ms, p.a, p.b) != cudaSuccess) return;
    if (p.kind == 0) {
        ctx->stats.iterate_ms += ms;
        ctx->stats.level_iterate_ms[std::min(p.level, TVL1_MAX_LEVELS - 1)] += ms;
    } else if (p.kind == 1) ctx->stats.warp_ms += ms;
    else if (p.kind == 2) ctx->stats.total_ms += ms;
    else if (p.kind == 3) ctx->stats.pyramid_ms += ms;
    else if (p.kind == 4) ctx->stats.zoom_in_ms += ms;
    else if (p.kind == 6) ctx->stats.level_first_block_ms[std::min(p.level, TVL1_MAX_LEVELS - 1)] += ms;   // (inside a kind-0 span)
    else ctx->stats.export_ms += ms;
}

void resolve_events(tvl1_ctx *ctx)
{
    for (auto &p : ctx->ev_used) {
        add_span_time(ctx, p);
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->ev_used.clear();
}

// ---- launch helpers -----------------------------------------------------------------------------
int launch_gauss(tvl1_ctx *ctx, int D, const float *in, int in_pitch, size_t in_stride, float *out,
                 int out_pitch, size_t out_stride, int nx, int ny, int onx, int ony,
                 const GaussTaps &taps, const unsigned int *mm, int B, int nimg)
{
    cudaStream_t st = ctx->stream;
    const int r = taps.size - 1;
#define TVL1_GAUSS(D_, R_) k_gauss<D_, R_><<<g, dim3(32, 8), 0, st>>>(in, in_pitch, in_stride, out, out_pitch, \
                                                                      out_stride, nx, ny, onx, ony, taps, mm, B)
#define TVL1_GAUSS_MARCH(D_, R_) k_gauss_march<D_, R_><<<gm, dim3(32, 4), 0, st>>>(in, in_pitch, in_stride, out, \
                                                           out_pitch, out_stride, nx, ny, onx, ony, taps, mm, B)
#define TVL1_GAUSS_SHFL(D_, R_) k_gauss_shfl<D_, R_><<<gs, dim3(32, 4), 0, st>>>(in, in_pitch, in_stride, out, \
                                                           out_pitch, out_stride, nx, ny, onx, ony, taps, mm, B)
    const dim3 gm(ceil_div(onx, 128), ceil_div(ony, 128), nimg);   // marching kernel: 128 columns x 4 strips of 32 rows
    const dim3 gs(ceil_div(onx, 120), ceil_div(ony, 128), nimg);   // ... with shuffled row inputs: 120 columns per warp
    const bool shfl = ctx->gauss_shfl;
    if (D == 1) {
        dim3 g(ceil_div(onx, 64), ceil_div(ony, 32), nimg);
        if (r == 4) { if (shfl) TVL1_GAUSS_SHFL(1, 4); else TVL1_GAUSS_MARCH(1, 4); }
        else if (r == 5) TVL1_GAUSS_MARCH(1, 5);
        else TVL1_GAUSS(1, 0);
    } else {
        dim3 g(ceil_div(onx, 32), ceil_div(ony, 16), nimg);
        if (r == 5) { if (shfl) TVL1_GAUSS_SHFL(2, 5); else TVL1_GAUSS_MARCH(2, 5); }
        else TVL1_GAUSS(2, 0);
    }
#undef TVL1_GAUSS
#undef TVL1_GAUSS_MARCH
#undef TVL1_GAUSS_SHFL
    CKL(ctx);
    return TVL1_OK;
}

IterParams iter_params(const tvl1_ctx *ctx, const Level &lv, const tvl1_params &prm, int stat_slot,
                       int max_iter, int level = 0)
{
    const Workspace &w = ctx->ws;
    IterParams P = {};
    P.state = w.state; P.consts = w.consts; P.ctl = w.ctl; P.partials = w.partials;
    P.loop = w.loop; P.cond = 0; P.use_cond = 0;
    P.batch = w.B; P.tb_partials = w.tb_partials; P.tb_parts = w.tb_parts; P.tb = 0; P.tb_max = kTbT;
    P.row_begin = 0; P.row_end = lv.ny; P.band_sum = nullptr;
    P.stat_iters = w.stat_iters; P.stat_errs = w.stat_errs;
    P.px_iters = w.counters;
    P.level = std::min(level, TVL1_MAX_LEVELS - 1);
    P.plane0 = w.plane0; P.field_stride = w.field_stride; P.set_stride = w.set_stride;
    P.lv = lv; P.parts_per_pair = w.parts_per_pair;
    P.stat_stride = w.stat_stride; P.stat_slot = stat_slot; P.max_iter = max_iter;
    P.l_t = (float) (prm.lambda * prm.theta);                   // src/tvl1flow.cpp:62
    P.theta = (float) prm.theta;
    P.taut = (float) (prm.tau / prm.theta);                     // src/tvl1flow.cpp:171
    P.eps2 = prm.epsilon * prm.epsilon;                         // src/tvl1flow.cpp:113
    return P;
}

int launch_iterate_tb(tvl1_ctx *ctx, const IterParams &P, int B, bool tail);
bool tb_usable(tvl1_ctx *ctx, const Level &l, int B, bool tail = false);

// Two iterations per launch in registers (k_iterate_t2): for the launches that saturate HBM, i.e. where the
// shared-memory kernel is not used -- big lock-step batches, and chunks solved while other lanes share the GPU.
// Needs enough strips of 120 x 64 pixels to fill the GPU; whole images only (no row bands).
bool t2_usable(const tvl1_ctx *ctx, const Level &l, int B, bool peers, int level)
{
    if (!ctx->use_t2 || peers) return false;
    if (ctx->shared_gpu && !ctx->t2_when_shared) return false;
    if (ctx->use_t2 == 2) return true;
    if (!((ctx->t2_levels >> std::min(level, 31)) & 1u)) return false;
    const long long strips = (long long) ceil_div(l.nx, kT2W) * ceil_div(l.ny, 16 * kIterWY) * B;
    return l.nx >= kT2W && l.ny >= 16 && strips >= 4ll * ctx->sm_count;
}

// A streamed level starts from zero duals (src/tvl1flow.cpp:87-90), and its first warp step practically never stops
// after one iteration: where the launch is big enough for k_iterate_t2, the level's first LAUNCH runs its first TWO
// iterations for every pair, taking the duals as zero (no zeroing pass, no loads) -- 47 B per pixel instead of 44 + 60.
// A pair that does stop after one iteration has the block rejected and its first iteration replayed, again from
// zero duals (PairCtl::pzero).
bool t2_first_usable(const tvl1_ctx *ctx, const IterParams &P, int B)
{
    if (!ctx->t2_first || ctx->use_t2 == 0 || P.peers.enabled || P.max_iter < 2) return false;
    if (ctx->shared_gpu && !ctx->t2_when_shared) return false;
    if (ctx->use_t2 == 2) return true;
    const long long strips = (long long) ceil_div(P.lv.nx, kT2W) * ceil_div(P.lv.ny, 16 * kIterWY) * B;
    return P.lv.nx >= kT2W && P.lv.ny >= 16 && strips >= 4ll * ctx->sm_count;
}

// mode: 0 one iteration per launch, 1 blocks of up to kTbT through k_iterate_tb, 2 blocks of two through k_iterate_t2
void set_blocking(IterParams &P, int mode)
{
    P.tb = mode;
    P.tb_max = mode == 2 ? kT2T : kTbT;
}

// grid.z of the iteration kernels: pair slots (see for_each_pair_of_slot).  Enough slots that a launch
// with every pair active still has a few thousand CTAs, few enough that a launch with hardly any
// active pair does not spend its time starting CTAs that exit at once.
int pair_slots(const tvl1_ctx *ctx, int tiles, int B, bool tail = false, bool rounds = false)
{
    // at least ceil(B / 32) slots: one ballot of the kernel covers a slot's pairs (the temporally blocked
    // kernel goes round again instead: any number of slots)
    if (tail) return std::min(B, std::max(rounds ? 1 : ceil_div(B, 32), ceil_div(ctx->tail_slot_ctas, std::max(tiles, 1))));
    return std::min(B, std::max(std::max(32, ceil_div(B, 32)), ceil_div(ctx->slot_ctas, std::max(tiles, 1))));
}

int launch_iterate_t2(tvl1_ctx *ctx, const IterParams &P, int B, bool tail)
{
    const int rows = P.row_end - P.row_begin;
    const int tiles_x = ceil_div(P.lv.nx, kT2W);
    const int Bw = tail ? std::max(1, (ctx->shared_gpu ? ctx->tail_pairs_shared : ctx->tail_pairs) / 8) : B;
    // rows staged through shared memory (cp.async ring) need more than the default 48 KB per CTA
    static bool attr_done[64] = { false };
    bool stage = ctx->t2_stage && !(ctx->shared_gpu && !ctx->t2_stage_when_shared);
    if (stage && !attr_done[ctx->device & 63]) {
        const int bytes = (int) t2_smem_bytes(kIterWY);
        if (cudaFuncSetAttribute(k_iterate_t2<32, kIterWY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
            cudaFuncSetAttribute(k_iterate_t2<16, kIterWY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
            cudaGetLastError();
            ctx->t2_stage = stage = false;
        } else {
            attr_done[ctx->device & 63] = true;
        }
    }
    const size_t smem = stage ? t2_smem_bytes(kIterWY) : 0;
    // tall strips: the two halo rows above and the one below are loaded per strip
    if ((long long) tiles_x * ceil_div(rows, 32 * kIterWY) * Bw >= 4ll * ctx->sm_count) {
        dim3 g(tiles_x, ceil_div(rows, 32 * kIterWY), 1);
        g.z = pair_slots(ctx, g.x * g.y, B, tail);
        if (stage) k_iterate_t2<32, kIterWY, true><<<g, 32 * kIterWY, smem, ctx->stream>>>(P);
        else k_iterate_t2<32, kIterWY, false><<<g, 32 * kIterWY, 0, ctx->stream>>>(P);
    } else {
        dim3 g(tiles_x, ceil_div(rows, 16 * kIterWY), 1);
        g.z = pair_slots(ctx, g.x * g.y, B, tail);
        if (stage) k_iterate_t2<16, kIterWY, true><<<g, 32 * kIterWY, smem, ctx->stream>>>(P);
        else k_iterate_t2<16, kIterWY, false><<<g, 32 * kIterWY, 0, ctx->stream>>>(P);
    }
    CK(cudaGetLastError());
    return TVL1_OK;
}

// `tail`: narrow launch for the late iterations of a lock-step batch, when only a few pairs are left
int launch_iterate(tvl1_ctx *ctx, const IterParams &P, int B, bool tail = false)
{
    // Strip height per warp: 16 rows (fewest CTAs, least halo traffic) when that still gives every SM
    // a few CTAs, else 8, else 4 -- a single mid-size image is latency-bound on how many rows a warp
    // walks, not on bytes.
    const int rows = P.row_end - P.row_begin;
    const int tiles_x = ceil_div(P.lv.nx, 124);
    const long long want = 4ll * ctx->sm_count;
    // a tail launch serves the few pairs still iterating (often one or two): size the strips for an eighth of the switch-over count
    const int Bw = tail ? std::max(1, (ctx->shared_gpu ? ctx->tail_pairs_shared : ctx->tail_pairs) / 8) : B;
    auto ctas = [&](int R) { return (long long) tiles_x * ceil_div(rows, R * kIterWY) * Bw; };
    if (ctas(16) >= want) {
        dim3 g(tiles_x, ceil_div(rows, 16 * kIterWY), 1);
        g.z = pair_slots(ctx, g.x * g.y, B, tail);
        k_iterate_t1<16, kIterWY><<<g, 32 * kIterWY, 0, ctx->stream>>>(P);
    } else if (ctas(8) >= want) {
        dim3 g(tiles_x, ceil_div(rows, 8 * kIterWY), 1);
        g.z = pair_slots(ctx, g.x * g.y, B, tail);
        k_iterate_t1<8, kIterWY><<<g, 32 * kIterWY, 0, ctx->stream>>>(P);
    } else {
        dim3 g(tiles_x, ceil_div(rows, 4 * kIterWY), 1);
        g.z = pair_slots(ctx, g.x * g.y, B, tail);
        k_iterate_t1<4, kIterWY><<<g, 32 * kIterWY, 0, ctx->stream>>>(P);
    }
    CK(cudaGetLastError());      // launches of this kernel are counted on the device (fetch_stats)
    // pairs whose next block has more than one iteration
    if (P.tb == 2) TRY(launch_iterate_t2(ctx, P, B, tail));
    else if (P.tb) TRY(launch_iterate_tb(ctx, P, B, tail));
    return TVL1_OK;
}


// ---- TMA descriptors for the temporally blocked kernel ------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) p;
        else
            cudaGetLastError();
    }
    return fn;
}

// planes[z][y][x] with row pitch `pitch` and plane stride `plane` (floats); box bw x bh x 1, zero fill
bool make_plane_map(CUtensorMap *m, const float *base, int nx, int ny, int nplanes, int pitch, size_t plane,
                    int bw = kTbBW, int bh = kTbBH)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = { (cuuint64_t) nx, (cuuint64_t) ny, (cuuint64_t) nplanes };
    const cuuint64_t strides[2] = { (cuuint64_t) pitch * sizeof(float), (cuuint64_t) plane * sizeof(float) };
    const cuuint32_t box[3] = { (cuuint32_t) bw, (cuuint32_t) bh, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// B = pairs that a launch is expected to serve (the lock-step batch, or the tail of one)
bool tb_usable(tvl1_ctx *ctx, const Level &l, int B, bool tail)
{
    static bool attr_done[64] = { false };      // function attributes are per device
    bool &attr = attr_done[ctx->device & 63];
    if (!ctx->use_tb || !tensor_map_encoder()) return false;
    // The blocked kernel trades HBM traffic for instruction issue and shared-memory wavefronts: it wins
    // while the launch leaves the GPU under-used (single images, small batches), and loses to the
    // streaming kernel once the SMs are saturated anyway -- by a big lock-step batch, or by the other
    // lanes of a chunked batch (measured, profiles/r2g_chunks.txt: 16 pairs x 4 lanes 178.9 -> 165.7 ms,
    // 64 x 4 lanes 163.7 -> 157.0 ms per 256 x 1080p without it).
    // (the narrow launches of a tail are the exception: what they cost the other lanes is small, what they save the
    // chunk's slow pairs is most of their time)
    if (ctx->shared_gpu && !ctx->tb_when_shared && !tail) return false;
    if ((long long) B * l.nx * l.ny > ctx->tb_max_pixels) return false;
    if (l.nx < kTbBW || l.ny < kTbBH) return false;
    if (!attr) {
        if (cudaFuncSetAttribute(k_iterate_tb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTbSmemBytes) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr = true;
    }
    return true;
}

int launch_iterate_tb(tvl1_ctx *ctx, const IterParams &P, int B, bool tail)
{
    const Workspace &w = ctx->ws;
    TbMaps maps;
    bool ok = true;
    for (int set = 0; set < 2; set++)
        ok = ok && make_plane_map(&maps.state[set], w.state + (size_t) set * w.set_stride, P.lv.nx, P.lv.ny,
                                  F_COUNT * B, P.lv.pitch, w.plane0);
    ok = ok && make_plane_map(&maps.consts, w.consts, P.lv.nx, P.lv.ny, C_COUNT * B, P.lv.pitch, w.plane0);
    if (!ok) { ctx->err = "cuTensorMapEncodeTiled failed"; return TVL1_ERR_CUDA; }
    dim3 g(ceil_div(P.lv.nx, kTbW), ceil_div(P.row_end - P.row_begin, kTbH), 1);
    g.z = pair_slots(ctx, g.x * g.y, B, tail, true);
    k_iterate_tb<<<g, kTbThreads, kTbSmemBytes, ctx->stream>>>(maps, P);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// The whole while loop of one warp step for every pair, on chip (one cluster per pair).
int launch_resident(tvl1_ctx *ctx, int s, int B, const tvl1_params &prm, int stat_slot, int max_iter,
                    double eps2, double *err_trace)
{
    const Workspace &w = ctx->ws;
    ResParams P;
    P.state = w.state; P.consts = w.consts; P.ctl = w.ctl;
    P.stat_iters = w.stat_iters; P.stat_errs = w.stat_errs; P.counters = w.counters;
    P.err_trace = err_trace;
    P.plane0 = w.plane0; P.field_stride = w.field_stride; P.set_stride = w.set_stride;
    P.lv = w.lv[s]; P.rows_per_cta = w.res_rows[s];
    P.stat_stride = w.stat_stride; P.stat_slot = stat_slot; P.max_iter = max_iter;
    P.level = std::min(s, TVL1_MAX_LEVELS - 1);
    P.l_t = (float) (prm.lambda * prm.theta);
    P.theta = (float) prm.theta;
    P.taut = (float) (prm.tau / prm.theta);
    P.eps2 = eps2;
    const int C = w.res_cluster[s];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, B);
    cfg.blockDim = dim3(kResThreads);
    cfg.dynamicSmemBytes = resident_smem_bytes(P.lv.pitch, P.rows_per_cta);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_iterate_resident, P));
    return TVL1_OK;          // counted on the device like the streaming kernel
}

int launch_warp(tvl1_ctx *ctx, int s, int B, int write_grad = 0, int row_begin = 0, int row_end = -1)
{
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    if (row_end < 0) row_end = l.ny;
    dim3 g(ceil_div(l.nx, kWarpTW), ceil_div(row_end - row_begin, kWarpTH), B);
    // interior tiles pull their box of I1 with one TMA copy: a 3-D tensor map (nx, ny, pair) of the level
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    const int use_tma = ctx->warp_tma && l.nx >= kWarpBW && l.ny >= kWarpBH &&
                        make_plane_map(&map, w.I1(s), l.nx, l.ny, B, l.pitch, w.plane(s), kWarpBW, kWarpBH) ? 1 : 0;
    k_warp<<<g, dim3(32, 8), 0, ctx->stream>>>(map, use_tma, w.I0(s), w.I1(s), w.plane(s), w.state, w.plane0,
                                               w.field_stride, w.set_stride, w.ctl, w.consts, l, write_grad,
                                               row_begin, row_end);
    CKL(ctx);
    ctx->stats.pixel_warps += (unsigned long long) B * l.nx * (row_end - row_begin);
    return TVL1_OK;
}

int launch_zero(tvl1_ctx *ctx, int s, int B, int first_field, int nfields)
{
    const Workspace &w = ctx->ws;
    const size_t n4 = w.plane(s) / 4;
    dim3 g((unsigned) std::min<size_t>((n4 + 255) / 256, 1024), nfields, B);
    k_zero_fields<<<g, 256, 0, ctx->stream>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                              n4, first_field);
    CKL(ctx);
    return TVL1_OK;
}

// The while loop of src/tvl1flow.cpp:113 for the whole batch, as a conditional WHILE node of the
// solve graph: the body is one launch of the fused iteration kernel; the kernel clears the
// condition when the last pair stops.  (default launch value 1 => at least one iteration)
int add_while_node(tvl1_ctx *ctx, const IterParams &P, int B, cudaGraphConditionalHandle h, bool tail)
{
    cudaStreamCaptureStatus status;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    CK(cudaStreamGetCaptureInfo_v2(ctx->stream, &status, nullptr, &g, &deps, &ndeps));
    cudaGraphNodeParams np = { cudaGraphNodeTypeConditional };
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, deps, ndeps, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(ctx->body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    std::swap(ctx->stream, ctx->body_stream);
    const int rc = launch_iterate(ctx, P, B, tail);
    std::swap(ctx->stream, ctx->body_stream);
    cudaGraph_t ended = nullptr;
    CK(cudaStreamEndCapture(ctx->body_stream, &ended));
    TRY(rc);
    CK(cudaStreamUpdateCaptureDependencies(ctx->stream, &node, 1, cudaStreamSetCaptureDependencies));
    return TVL1_OK;
}

// first_zero: this is the first warp step of a level whose duals were NOT zeroed by a pass of their own: the
// loop's first iteration (which always runs: error starts at infinity) is launched explicitly, with the duals
// taken as zero instead of read, before the while nodes -- 32 B per pixel and level less HBM traffic.
// First launch of a streamed level (duals read as zero): two iterations for every pair through k_iterate_t2 where
// that kernel fits (t2_first_usable), else one iteration through k_iterate_t1.
int launch_first_zero(tvl1_ctx *ctx, const IterParams &P, int B)
{
    IterParams P0 = P;
    P0.p_zero = 1;
    if (t2_first_usable(ctx, P, B)) {
        Span sp(ctx, 6, P.level);
        set_blocking(P0, 2);
        P0.take_all = 1;
        return launch_iterate_t2(ctx, P0, B, false);
    }
    P0.tb = 0;                          // every pair's first block is one iteration: the streaming kernel serves it
    return launch_iterate(ctx, P0, B);
}

int add_while_loop(tvl1_ctx *ctx, IterParams P, int B, bool first_zero)
{
    cudaStreamCaptureStatus status;
    cudaGraph_t g = nullptr;
    CK(cudaStreamGetCaptureInfo_v2(ctx->stream, &status, nullptr, &g, nullptr, nullptr));
    cudaGraphConditionalHandle h_all, h_bulk = 0;
    CK(cudaGraphConditionalHandleCreate(&h_all, g, 1, cudaGraphCondAssignDefault));
    P.cond = h_all;
    P.use_cond = 1;
    P.cond_bulk = 0;
    P.bulk_min = -1;
    // A big lock-step batch launches until its slowest pair has converged, and most of those launches
    // find only a few pairs still active.  Two while nodes in sequence: wide launches (full-size grid)
    // while more than tail_pairs pairs iterate, then narrow ones (a tenth of the CTAs: a launch with
    // little work costs what it does, not what it takes to start and retire 32k empty CTAs).  The
    // active count only falls within a warp step, so the second loop never has to hand back.
    // A chunk solved while other lanes share the GPU has a tail too: the one or two slow pairs that keep its lock-step
    // loop turning.  They do not cost the GPU much (the other lanes fill it) but they decide when the chunk -- and, if it
    // is a late one, the whole host-buffer call -- ends: the same hand-over, at tail_pairs_shared pairs.
    const int tail_pairs = ctx->shared_gpu ? ctx->tail_pairs_shared : ctx->tail_pairs;
    const bool two_phase = tail_pairs > 0 && B > 2 * tail_pairs && !P.peers.enabled;
    if (two_phase) {
        CK(cudaGraphConditionalHandleCreate(&h_bulk, g, 1, cudaGraphCondAssignDefault));
        P.cond_bulk = h_bulk;
        P.bulk_min = tail_pairs;
    }
    if (first_zero) TRY(launch_first_zero(ctx, P, B));
    if (two_phase) {
        TRY(add_while_node(ctx, P, B, h_bulk, false));
        // the pairs of the tail are the slow ones (tens of iterations where the batch needs two): worth
        // temporal blocking even where the full batch is not (a wide launch of both kernels costs more
        // than blocking saves; a narrow one does not)
        if (P.tb != 1 && ctx->tail_tb && tb_usable(ctx, P.lv, tail_pairs, true)) set_blocking(P, 1);
        return add_while_node(ctx, P, B, h_all, true);
    }
    return add_while_node(ctx, P, B, h_all, false);
}

// Host-driven variant of the same loop (TVL1_NO_GRAPH=1, and the per-kernel hooks): enqueue a
// chunk of iteration launches, then read back how many pairs still iterate.  Launches for pairs
// that already stopped exit at once, so over-shooting costs microseconds while every look costs
// a stream synchronisation; the first chunk is the count the previous warp step needed.
int run_iterations(tvl1_ctx *ctx, const IterParams &P, int B, int &chunk_hint, bool first_zero = false)
{
    if (ctx->capturing) {
        Span sp(ctx, 0, P.level);
        return add_while_loop(ctx, P, B, first_zero);
    }
    int launched = 0;
    if (first_zero) {
        Span sp(ctx, 0, P.level);
        TRY(launch_first_zero(ctx, P, B));
        launched = 2;
    }
    int chunk = std::max(1, std::min(chunk_hint, P.max_iter));
    while (launched < P.max_iter) {
        const int k = std::min(chunk, P.max_iter - launched);
        {
            Span sp(ctx, 0, P.level);
            for (int i = 0; i < k; i++) TRY(launch_iterate(ctx, P, B));
        }
        launched += k;
        CK(cudaMemcpyAsync(ctx->h_loop, ctx->ws.loop, sizeof(LoopCtl), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->stats.host_syncs++;
        if (ctx->h_loop->active_pairs == 0) break;
        chunk = std::max(2, chunk_hint / 4);
    }
    chunk_hint = std::max(1, ctx->h_loop->max_n);
    return TVL1_OK;
}

// One pyramid level: src/tvl1flow.cpp:46-212 for every pair of the batch.
int run_level(tvl1_ctx *ctx, int s, int B, const tvl1_params &prm, int stat_base, int &chunk_hint)
{
    const Workspace &w = ctx->ws;
    // p = 0 (:87-90): a zeroing pass for the levels that live on chip; the streamed levels take the duals
    // as zero in their first iteration instead (run_iterations, first_zero)
    const bool zero_in_first = ctx->zero_in_first && w.res_cluster[s] == 0;
    if (!zero_in_first) TRY(launch_zero(ctx, s, B, F_P11, 4));
    for (int wi = 0; wi < prm.warps; wi++) {                                // :92
        {
            Span sp(ctx, 1);
            TRY(launch_warp(ctx, s, B));                                    // :84, :94-109
        }
        if (w.res_cluster[s] > 0) {                                         // :111-182 on chip
            Span sp(ctx, 0, std::min(s, TVL1_MAX_LEVELS - 1));
            TRY(launch_resident(ctx, s, B, prm, stat_base + wi, kMaxIterations,
                                prm.epsilon * prm.epsilon, nullptr));
            continue;
        }
        k_begin_warp<<<ceil_div(B, 128), 128, 0, ctx->stream>>>(w.ctl, w.loop, B);   // :111-112
        CKL(ctx);
        IterParams P = iter_params(ctx, w.lv[s], prm, stat_base + wi, kMaxIterations, s);
        // blocks of two in registers where the launch saturates HBM and the level's loops are long enough, else blocks of up
        // to four in shared memory where the launch is small enough for that to pay, else one iteration per launch
        set_blocking(P, t2_usable(ctx, w.lv[s], B, false, s) ? 2 : tb_usable(ctx, w.lv[s], B) ? 1 : 0);
        TRY(run_iterations(ctx, P, B, chunk_hint, zero_in_first && wi == 0));   // :113-182
    }
    return TVL1_OK;
}

int check_sigma(tvl1_ctx *ctx, double sigma, int width, GaussTaps &taps)
{
    const int size = make_taps(sigma, taps);
    if (size < 0) return fail_arg(ctx, "Gaussian window larger than this build supports (zfactor too small)");
    if (size > width) {                                                     // src/operators.cpp:520-522
        ctx->err = "GaussianSmooth: sigma too large";
        return TVL1_ERR_SIGMA;
    }
    return TVL1_OK;
}

void reset_stats(tvl1_ctx *ctx) { ctx->stats = tvl1_stats{}; }

int fetch_stats(tvl1_ctx *ctx, int B, int nstat, int *iters_out, double *errs_out)
{
    const Workspace &w = ctx->ws;
    unsigned long long c[kCounterWords] = { 0 };
    CK(cudaMemcpyAsync(c, w.counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
    if (iters_out)
        CK(cudaMemcpy2DAsync(iters_out, sizeof(int) * nstat, w.stat_iters, sizeof(int) * w.stat_stride,
                             sizeof(int) * nstat, B, cudaMemcpyDeviceToHost, ctx->stream));
    if (errs_out)
        CK(cudaMemcpy2DAsync(errs_out, sizeof(double) * nstat, w.stat_errs, sizeof(double) * w.stat_stride,
                             sizeof(double) * nstat, B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(sleep_until_done(ctx));
    // (not from the chunks of a host-buffer batch: the loop lengths of 8 or 16 pairs swing with a single slow pair, and
    // every change of the mask re-captures the lane's solve graph while the other lanes are running)
    if (ctx->t2_adapt && ctx->use_t2 == 1 && nstat > 0 && !ctx->hs_mode && !ctx->shared_gpu) {
        // iterations per loop (pair and warp step) on each streamed level of this solve -> which levels the next one
        // blocks (t2_usable).  Loops are counted up to kLoopClip iterations: one slow pair that runs to the cap must not
        // switch the kernel for its whole chunk and back (every switch re-captures the solve graph).
        const int nlev = (int) w.lv.size();
        const int warps = std::max(1, nstat / std::max(1, nlev));
        for (int l = 0; l < nlev && l < 16; l++) {
            if (w.res_cluster[l] != 0) continue;
            const double mean = (double) c[2 * TVL1_MAX_LEVELS + l] / ((double) B * warps);
            if (mean >= 3.5) ctx->t2_levels |= 1u << l;
            else if (mean < 2.5) ctx->t2_levels &= ~(1u << l);
        }
    }
    for (int l = 0; l < TVL1_MAX_LEVELS; l++) {
        ctx->stats.pixel_iterations += c[l];
        ctx->stats.level_pixel_iterations[l] += c[l];
        ctx->stats.iterate_launches += c[TVL1_MAX_LEVELS + l];
        ctx->stats.level_iterate_launches[l] += c[TVL1_MAX_LEVELS + l];
        ctx->stats.kernel_launches += c[TVL1_MAX_LEVELS + l];
    }
    return TVL1_OK;
}

// src/tvl1flow.cpp:278-310 (multiscale) or the single level of Dual_TVL1_optic_flow.
int enqueue_coarse_to_fine(tvl1_ctx *ctx, int B, const tvl1_params &prm, bool multiscale)
{
    const Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const int ns = multiscale ? prm.nscales : 1;
    if (multiscale) TRY(launch_zero(ctx, ns - 1, B, F_U1, 2));
    int chunk_hint = 16;
    for (int s = ns - 1; s >= 0; s--) {
        TRY(run_level(ctx, s, B, prm, (ns - 1 - s) * prm.warps, chunk_hint));
        if (!s) break;
        const Level &c = w.lv[s], &f = w.lv[s - 1];
        Span zs(ctx, 4);
        dim3 g(ceil_div(f.nx, kZiTW), ceil_div(f.ny, kZiTH), B);
        k_zoom_in_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                                  c, f, (double) f.nx / c.nx, (double) f.ny / c.ny,
                                                  (float) (1.0 / prm.zfactor), 0, f.ny);
        CKL(ctx);
        k_flip_cur<<<ceil_div(B, 128), 128, 0, st>>>(w.ctl, B);
        CKL(ctx);
        chunk_hint = std::max(chunk_hint, 4);
    }
    return TVL1_OK;
}

bool same_params(const tvl1_params &a, const tvl1_params &b)
{
    return a.tau == b.tau && a.lambda == b.lambda && a.theta == b.theta && a.nscales == b.nscales &&
           a.zfactor == b.zfactor && a.warps == b.warps && a.epsilon == b.epsilon;
}

// Capture `enqueue` into `sg` (once per key) and replay it.  `variant` distinguishes the graphs a
// context keeps (whole coarse-to-fine solve; one level of the row-band mode, ...).
template <class Fn>
int replay_graph(tvl1_ctx *ctx, SolveGraph &sg, const tvl1_params &prm, bool multiscale, long long variant,
                 Fn &&enqueue)
{
    if (!sg.exec || !same_params(sg.prm, prm) || sg.multiscale != multiscale || sg.profiling != ctx->profiling ||
        sg.variant != variant) {
        free_graph(sg, ctx->ev_pool);
        sg.prm = prm; sg.multiscale = multiscale; sg.profiling = ctx->profiling; sg.variant = variant;
        const tvl1_stats keep = ctx->stats;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        ctx->capturing = true;
        ctx->cap = &sg;
        const int rc = enqueue();
        ctx->capturing = false;
        ctx->cap = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &sg.graph);
        sg.static_launches = ctx->stats.kernel_launches - keep.kernel_launches;
        sg.pixel_warps = ctx->stats.pixel_warps - keep.pixel_warps;
        ctx->stats = keep;
        if (rc != TVL1_OK) { free_graph(sg, ctx->ev_pool); return rc; }
        CK(e);
        CK(cudaGraphInstantiate(&sg.exec, sg.graph, 0));
    }
    CK(cudaGraphLaunch(sg.exec, ctx->stream));
    ctx->stats.kernel_launches += sg.static_launches;
    ctx->stats.pixel_warps += sg.pixel_warps;
    if (!sg.events.empty()) {
        CK(sleep_until_done(ctx));
        for (const auto &p : sg.events) add_span_time(ctx, p);
    }
    return TVL1_OK;
}

int run_coarse_to_fine(tvl1_ctx *ctx, int B, const tvl1_params &prm, bool multiscale)
{
    if (!ctx->use_graph) return enqueue_coarse_to_fine(ctx, B, prm, multiscale);
    // (the kernel choice depends on whether other lanes share the GPU: tb_usable)
    // (... and on the levels the two-iteration kernel serves, which follow the previous solve: t2_adapt)
    long long t2_key = 0;
    for (int s = 0; s < (int) ctx->ws.lv.size() && s < 16; s++)
        if (ctx->use_t2 == 1 && ctx->ws.res_cluster[s] == 0 && ((ctx->t2_levels >> s) & 1u)) t2_key |= 2ll << s;
    return replay_graph(ctx, ctx->sg, prm, multiscale, (ctx->shared_gpu ? 1 : 0) | t2_key,
                        [&]() { return enqueue_coarse_to_fine(ctx, B, prm, multiscale); });
}

// Resets the per-solve device state and builds both image pyramids (src/tvl1flow.cpp:255-275).
int build_pyramid(tvl1_ctx *ctx, int B, const float *dI0, const float *dI1, int nx, int ny, const tvl1_params &prm)
{
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const int ns = prm.nscales;
    // validate the blur windows first: the reference throws before producing anything
    GaussTaps pre, zoom;
    TRY(check_sigma(ctx, TVL1_PRESMOOTHING_SIGMA, nx, pre));
    const double zsigma = TVL1_ZOOM_SIGMA_ZERO * std::sqrt(1.0 / (prm.zfactor * prm.zfactor) - 1.0);
    for (int s = 1; s < ns; s++) TRY(check_sigma(ctx, zsigma, w.lv[s - 1].nx, zoom));

    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    k_init_ctl<<<ceil_div(B, 128), 128, 0, st>>>(w.ctl, w.mm, B);
    CKL(ctx);

    // normalisation + pre-smoothing, src/tvl1flow.cpp:255-259
    const size_t n = (size_t) nx * ny;
    Span pyr(ctx, 3);
    {
        dim3 g((unsigned) std::min<size_t>((n + 2047) / 2048, 256), B);
        k_minmax<<<g, 256, 0, st>>>(dI0, dI1, n, w.mm);
        CKL(ctx);
    }
    TRY(launch_gauss(ctx, 1, dI0, nx, n, w.I0(0), w.lv[0].pitch, w.plane(0), nx, ny, nx, ny, pre, w.mm, B, B));
    TRY(launch_gauss(ctx, 1, dI1, nx, n, w.I1(0), w.lv[0].pitch, w.plane(0), nx, ny, nx, ny, pre, w.mm, B, B));

    // pyramid, src/tvl1flow.cpp:262-275 (zoom_out, src/zoom.cpp:41-78)
    for (int s = 1; s < ns; s++) {
        const Level &a = w.lv[s - 1], &c = w.lv[s];
        if (prm.zfactor == 0.5) {
            TRY(launch_gauss(ctx, 2, w.I0(s - 1), a.pitch, w.plane(s - 1), w.I0(s), c.pitch, w.plane(s),
                             a.nx, a.ny, c.nx, c.ny, zoom, nullptr, B, 2 * B));
        } else {
            TRY(launch_gauss(ctx, 1, w.I0(s - 1), a.pitch, w.plane(s - 1), w.tmp, a.pitch, w.plane(s - 1),
                             a.nx, a.ny, a.nx, a.ny, zoom, nullptr, B, 2 * B));
            dim3 g(ceil_div(c.nx, 32), ceil_div(c.ny, 8), 2 * B);
            k_resample<<<g, dim3(32, 8), 0, st>>>(w.tmp, a.pitch, w.plane(s - 1), a.nx, a.ny, w.I0(s),
                                                  c.pitch, w.plane(s), c.nx, c.ny, prm.zfactor,
                                                  prm.zfactor, 1.0f);
            CKL(ctx);
        }
    }
    return TVL1_OK;
}

// Dual_TVL1_optic_flow_multiscale for B <= max_batch pairs, device-resident dense inputs/outputs.
int run_multiscale(tvl1_ctx *ctx, int B, const float *dI0, const float *dI1, float *du1, float *du2,
                   int nx, int ny, const tvl1_params &prm, int *iters_out, double *errs_out)
{
    const int ns = prm.nscales;
    const int nstat = ns * prm.warps;
    TRY(ensure_workspace(ctx, nx, ny, ns, prm.zfactor, B, nstat));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;

    Span total(ctx, 2);
    TRY(build_pyramid(ctx, B, dI0, dI1, nx, ny, prm));
    TRY(run_coarse_to_fine(ctx, B, prm, true));
    {
        Span ex(ctx, 5);
        dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), B);
        k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                                 w.lv[0], du1, du2);
        CKL(ctx);
    }
    total.end();
    TRY(fetch_stats(ctx, B, nstat, iters_out, errs_out));
    return TVL1_OK;
}

// Dual_TVL1_optic_flow (one level, no normalisation / blur), device-resident dense buffers.
int run_single_scale(tvl1_ctx *ctx, int B, const float *dI0, const float *dI1, float *du1, float *du2,
                     int nx, int ny, const tvl1_params &prm, int *iters_out, double *errs_out)
{
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, B, prm.warps));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    Span total(ctx, 2);
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    k_init_ctl<<<ceil_div(B, 128), 128, 0, st>>>(w.ctl, w.mm, B);
    CKL(ctx);
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), B);
    k_pack<<<g, dim3(32, 8), 0, st>>>(dI0, w.I0(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_pack<<<g, dim3(32, 8), 0, st>>>(dI1, w.I1(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_import_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                             w.lv[0], du1, du2);
    CKL(ctx);
    TRY(run_coarse_to_fine(ctx, B, prm, false));
    k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                             w.lv[0], du1, du2);
    CKL(ctx);
    total.end();
    TRY(fetch_stats(ctx, B, prm.warps, iters_out, errs_out));
    return TVL1_OK;
}

int ensure_stage(tvl1_ctx *ctx, size_t bytes_each, bool need_f32)
{
    if (ctx->stage_bytes < bytes_each) {
        for (int i = 0; i < 2; i++) {
            cudaFree(ctx->stage_in[i]); ctx->stage_in[i] = nullptr;
            cudaFree(ctx->stage_out[i]); ctx->stage_out[i] = nullptr;
        }
        ctx->stage_bytes = 0;
        for (int i = 0; i < 2; i++) {
            CK(cudaMalloc(&ctx->stage_in[i], bytes_each));
            CK(cudaMalloc(&ctx->stage_out[i], bytes_each));
        }
        ctx->stage_bytes = bytes_each;
    }
    if (need_f32 && ctx->stage_f32_bytes < bytes_each / 2) {
        for (int i = 0; i < 4; i++) { cudaFree(ctx->stage_f32[i]); ctx->stage_f32[i] = nullptr; }
        ctx->stage_f32_bytes = 0;
        for (int i = 0; i < 4; i++) CK(cudaMalloc(&ctx->stage_f32[i], bytes_each / 2));
        ctx->stage_f32_bytes = bytes_each / 2;
    }
    return TVL1_OK;
}

int check_common(tvl1_ctx *ctx, const void *a, const void *b, const void *c, const void *d, int nx,
                 int ny, const tvl1_params *prm, bool multiscale)
{
    if (!ctx) return TVL1_ERR_ARG;
    if (!a || !b || !c || !d || !prm) return fail_arg(ctx, "null pointer argument");
    if (nx < 1 || ny < 1) return fail_arg(ctx, "image size must be positive");
    if (prm->warps < 1) return fail_arg(ctx, "warps must be >= 1");
    if (multiscale && prm->nscales < 1) return fail_arg(ctx, "nscales must be >= 1");
    if (multiscale && !(prm->zfactor > 0.0 && prm->zfactor < 1.0) && prm->nscales > 1)
        return fail_arg(ctx, "zfactor must be in (0,1)");
    if (!(prm->theta != 0.0)) return fail_arg(ctx, "theta must be non-zero");
    CK(cudaSetDevice(ctx->device));
    return TVL1_OK;
}

// ---- pageable host memory <-> device through pinned staging ---------------------------------------
// cudaMemcpyAsync from/to pageable memory is staged by the driver on one thread (6-12 GB/s here).  The
// reference's callers hand us plain new[]/malloc buffers, fp64 at that, so the drop-in call was
// dominated by those copies.  Instead: a few host threads copy (fp32) or convert (fp64 <-> fp32, halving
// the PCIe bytes) slices into a double-buffered pinned ring while the previous slice is on the wire.
constexpr size_t kPipeSlot = 8u << 20;      // bytes per pinned slot
constexpr int kPipeThreads = 4;

// The ring pays off when it halves the bytes (fp64 callers) or for very large planes; small fp32
// transfers stay on the driver's own staged copy (thread start-up would cost more than it saves).
template <typename T>
bool use_ring(size_t count)
{
    return sizeof(T) == 8 ? count >= (1u << 18) : count * sizeof(float) >= (24u << 20);
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

template <class Fn>
void parallel_slices(size_t n, Fn &&fn)
{
    const int nt = n >= (1u << 18) ? kPipeThreads : 1;
    if (nt == 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (int t = 1; t < nt; t++) {
        const size_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (b > a) th.emplace_back([&fn, a, b] { fn(a, b); });
    }
    fn(0, std::min(n, per));
    for (auto &t : th) t.join();
}

int ensure_pipe(tvl1_ctx *ctx)
{
    for (int i = 0; i < 2; i++) {
        if (!ctx->pipe_buf[i]) CK(cudaMallocHost(&ctx->pipe_buf[i], kPipeSlot));
        if (!ctx->pipe_ev[i]) CK(cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming));
    }
    return TVL1_OK;
}

// host (pageable, T) -> device fp32
template <typename T>
int upload_pageable(tvl1_ctx *ctx, float *dst, const T *src, size_t count)
{
    TRY(ensure_pipe(ctx));
    const size_t per = kPipeSlot / sizeof(float);
    int slot = 0;
    for (size_t off = 0; off < count; off += per, slot ^= 1) {
        const size_t n = std::min(per, count - off);
        CK(cudaEventSynchronize(ctx->pipe_ev[slot]));         // the copy that last used this slot is done
        float *pin = (float *) ctx->pipe_buf[slot];
        const T *s = src + off;
        parallel_slices(n, [pin, s](size_t a, size_t b) {
            if (sizeof(T) == sizeof(float)) memcpy(pin + a, s + a, (b - a) * sizeof(float));
            else for (size_t i = a; i < b; i++) pin[i] = (float) s[i];
        });
        CK(cudaMemcpyAsync(dst + off, pin, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->pipe_ev[slot], ctx->stream));
    }
    return TVL1_OK;
}

// device fp32 -> host (pageable, T); returns when the data is in `dst`
template <typename T>
int download_pageable(tvl1_ctx *ctx, T *dst, const float *src, size_t count)
{
    TRY(ensure_pipe(ctx));
    const size_t per = kPipeSlot / sizeof(float);
    size_t prev_off = 0, prev_n = 0;
    int slot = 0;
    auto drain = [&](int sl, size_t off, size_t n) -> int {
        CK(cudaEventSynchronize(ctx->pipe_ev[sl]));
        const float *pin = (const float *) ctx->pipe_buf[sl];
        T *d = dst + off;
        parallel_slices(n, [pin, d](size_t a, size_t b) {
            if (sizeof(T) == sizeof(float)) memcpy(d + a, pin + a, (b - a) * sizeof(float));
            else for (size_t i = a; i < b; i++) d[i] = (T) pin[i];
        });
        return TVL1_OK;
    };
    for (size_t off = 0; off < count; off += per, slot ^= 1) {
        const size_t n = std::min(per, count - off);
        CK(cudaMemcpyAsync(ctx->pipe_buf[slot], src + off, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->pipe_ev[slot], ctx->stream));
        if (prev_n) TRY(drain(slot ^ 1, prev_off, prev_n));   // overlaps the copy just issued
        prev_off = off; prev_n = n;
    }
    if (prev_n) TRY(drain(slot ^ 1, prev_off, prev_n));
    return TVL1_OK;
}

int run_hs_multiscale(tvl1_ctx *ctx, int B, const float *dI1, const float *dI2, float *du, float *dv, int nx,
                      int ny, const hs_params &prm, int *iters_out, double *errs_out);
int run_hs_single_scale(tvl1_ctx *ctx, int B, const float *dI1, const float *dI2, float *du, float *dv, int nx,
                        int ny, const hs_params &prm, int *iters_out, double *errs_out);

// One chunk of <= max_batch pairs through one lane (context): H2D, solve, D2H, all on the lane's
// stream.  I1 == nullptr selects the frame-sequence form: I0 holds consecutive frames, pair b is
// (frame b, frame b+1), and the chunk's B+1 frames cross PCIe once.
template <typename T>
int solve_chunk(tvl1_ctx *ctx, int first, int B, const T *I0, const T *I1, T *u1, T *u2, int nx, int ny,
                const tvl1_params *prm, int *iters_out, double *errs_out, bool multiscale, int nstat)
{
    const bool sequence = I1 == nullptr;
    const bool f64 = sizeof(T) == 8;
    const size_t n = (size_t) nx * ny;
    cudaStream_t st = ctx->stream;
    const size_t cnt = (size_t) B * n, off = (size_t) first * n;
    const unsigned g = (unsigned) std::min<size_t>((cnt + 255) / 256, 4096);
    // fp32 working buffers on the device
    float *d0, *d1, *o0, *o1;
    if (f64) { d0 = ctx->stage_f32[0]; d1 = ctx->stage_f32[1]; o0 = ctx->stage_f32[2]; o1 = ctx->stage_f32[3]; }
    else { d0 = (float *) ctx->stage_in[0]; d1 = (float *) ctx->stage_in[1];
           o0 = (float *) ctx->stage_out[0]; o1 = (float *) ctx->stage_out[1]; }
    // host -> device: pinned buffers go straight over PCIe (fp64 is narrowed on the device), pageable
    // buffers through the pinned staging ring (fp64 is narrowed on the host)
    auto put = [&](const T *h, void *dev_T, float *dev_f32, size_t count) -> int {
        if (use_ring<T>(count) && is_pageable(h)) return upload_pageable<T>(ctx, dev_f32, h, count);
        CK(cudaMemcpyAsync(dev_T, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
        if (f64) {
            k_f64_to_f32<<<g, 256, 0, st>>>((const double *) dev_T, dev_f32, count);
            CKL(ctx);
        }
        return TVL1_OK;
    };
    if (sequence) {
        TRY(put(I0 + off, ctx->stage_in[0], d0, cnt + n));
        d1 = d0 + n;
    } else {
        TRY(put(I0 + off, ctx->stage_in[0], d0, cnt));
        TRY(put(I1 + off, ctx->stage_in[1], d1, cnt));
    }
    if (!multiscale) {   // u1,u2 are in/out: the initial flow is used (src/tvl1flow.cpp:94)
        TRY(put(u1 + off, ctx->stage_out[0], o0, cnt));
        TRY(put(u2 + off, ctx->stage_out[1], o1, cnt));
    }
    int *it = iters_out ? iters_out + (size_t) first * nstat : nullptr;
    double *er = errs_out ? errs_out + (size_t) first * nstat : nullptr;
    if (ctx->hs_mode) {
        if (multiscale) TRY(run_hs_multiscale(ctx, B, d0, d1, o0, o1, nx, ny, ctx->hs, it, er));
        else TRY(run_hs_single_scale(ctx, B, d0, d1, o0, o1, nx, ny, ctx->hs, it, er));
    } else if (multiscale) TRY(run_multiscale(ctx, B, d0, d1, o0, o1, nx, ny, *prm, it, er));
    else TRY(run_single_scale(ctx, B, d0, d1, o0, o1, nx, ny, *prm, it, er));
    auto get = [&](T *h, void *dev_T, const float *dev_f32) -> int {
        if (use_ring<T>(cnt) && is_pageable(h)) return download_pageable<T>(ctx, h, dev_f32, cnt);
        if (f64) {
            k_f32_to_f64<<<g, 256, 0, st>>>(dev_f32, (double *) dev_T, cnt);
            CKL(ctx);
        }
        CK(cudaMemcpyAsync(h, f64 ? dev_T : (void *) dev_f32, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
        return TVL1_OK;
    };
    TRY(get(u1 + off, ctx->stage_out[0], o0));
    TRY(get(u2 + off, ctx->stage_out[1], o1));
    CK(sleep_until_done(ctx));
    return TVL1_OK;
}

void add_stats(tvl1_stats &a, const tvl1_stats &b)
{
    a.kernel_launches += b.kernel_launches; a.iterate_launches += b.iterate_launches;
    a.pixel_iterations += b.pixel_iterations; a.pixel_warps += b.pixel_warps;
    a.iterate_ms += b.iterate_ms; a.warp_ms += b.warp_ms; a.total_ms += b.total_ms;
    a.pyramid_ms += b.pyramid_ms; a.zoom_in_ms += b.zoom_in_ms; a.export_ms += b.export_ms;
    a.host_syncs += b.host_syncs;
    for (int l = 0; l < TVL1_MAX_LEVELS; l++) {
        a.level_pixel_iterations[l] += b.level_pixel_iterations[l];
        a.level_iterate_launches[l] += b.level_iterate_launches[l];
        a.level_iterate_ms[l] += b.level_iterate_ms[l];
        a.level_first_block_ms[l] += b.level_first_block_ms[l];
    }
}

// Runs `nchunks` chunks of work over up to four lanes.  Lane 0 is `ctx` on the calling thread; the
// others are private sibling contexts on the same GPU (own stream, workspace, solve graph), each
// driven by its own host thread.  Lanes take chunks round-robin and run concurrently on the GPU: the
// copies of one chunk overlap the kernels of the other, and the sparse tail launches of one lane
// (few pairs still iterating) are filled by the other lane's work.
// TVL1_PIPE_TRACE=1: one stderr line per chunk of a host-buffer batch (lane, pairs, host times in ms since the call began)
struct ChunkTrace {
    bool on = std::getenv("TVL1_PIPE_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    void line(const char *what, const tvl1_ctx *lane, int k, int B, double a, double b, double c) const
    {
        if (on) fprintf(stderr, "[tvl1 %s] chunk %2d pairs %3d lane %p: start %7.2f solved %7.2f done %7.2f\n", what, k, B,
                        (const void *) lane, a, b, c);
    }
};

// Lane 0 is `ctx`; the others are created on first use and take over the caller's settings.
int setup_lanes(tvl1_ctx *ctx, int nchunks, int want_lanes, tvl1_ctx **lanes, int *nlanes_out)
{
    lanes[0] = ctx;
    int nlanes = 1;
    if (!ctx->is_sibling)
        nlanes = std::max(1, std::min(std::min(want_lanes, nchunks), (int) tvl1_ctx::kMaxLanes));
    for (int l = 1; l < nlanes; l++) {
        tvl1_ctx *&sb = ctx->sib[l - 1];
        if (!sb) {
            if (tvl1_create(ctx->device, &sb) != TVL1_OK) {
                ctx->err = std::string("extra lane: ") + tvl1_last_error(nullptr);
                return TVL1_ERR_CUDA;
            }
            sb->is_sibling = true;
        }
        sb->max_batch = ctx->max_batch;
        sb->profiling = ctx->profiling;
        sb->use_graph = ctx->use_graph;
        sb->use_resident = ctx->use_resident;
        sb->hs_mode = ctx->hs_mode;
        sb->hs = ctx->hs;
        reset_stats(sb);
        lanes[l] = sb;
    }
    for (int l = 0; l < nlanes; l++) lanes[l]->blocking_wait = lanes[l]->shared_gpu = nlanes > 1;
    *nlanes_out = nlanes;
    return TVL1_OK;
}

// Runs `work(lane index)` on every lane (lane 0 on the calling thread), then folds the lanes' statistics and
// the first error into `ctx`.
template <class Fn>
int join_lanes(tvl1_ctx *ctx, tvl1_ctx **lanes, int nlanes, int *rcs, Fn &&work)
{
    std::vector<std::thread> threads;
    for (int l = 1; l < nlanes; l++) threads.emplace_back(work, l);
    work(0);
    for (auto &t : threads) t.join();
    for (int l = 1; l < nlanes; l++) {
        add_stats(ctx->stats, lanes[l]->stats);
        if (rcs[l] != TVL1_OK && rcs[0] == TVL1_OK) { ctx->err = lanes[l]->err; rcs[0] = rcs[l]; }
    }
    return rcs[0];
}

template <class Fn>
int run_lanes(tvl1_ctx *ctx, int nchunks, int want_lanes, Fn &&chunk_fn)
{
    tvl1_ctx *lanes[tvl1_ctx::kMaxLanes];
    int nlanes = 1;
    TRY(setup_lanes(ctx, nchunks, want_lanes, lanes, &nlanes));
    int rcs[tvl1_ctx::kMaxLanes] = {};          // TVL1_OK == 0
    std::atomic<int> next{0};
    return join_lanes(ctx, lanes, nlanes, rcs, [&](int l) {
        tvl1_ctx *c = lanes[l];
        cudaSetDevice(c->device);
        // chunks are taken from a shared queue in order (their sizes may differ, see chunk_schedule)
        for (int k = next++; k < nchunks && rcs[l] == TVL1_OK; k = next++) rcs[l] = chunk_fn(c, k);
        resolve_events(c);
    });
}

// Cuts a host-buffer batch into lock-step chunks.  The pipeline of a call cannot start computing before
// its first chunk has arrived and cannot finish before its last chunk has been downloaded, and towards
// the end the lanes run out of work at different times; so when there are enough chunks the first and
// the last ones (one per lane) are SHORT (Bmax / short_div pairs), the ones in between full.  Only two
// chunk sizes occur (plus at most one ragged remainder), which is what a lane keeps workspaces for.
void chunk_schedule(const tvl1_ctx *ctx, int npairs, int Bmax, int lanes, std::vector<std::pair<int, int>> &out)
{
    const int div = ctx->short_div;
    const int small = div > 1 ? Bmax / div : 0;
    const int nlanes = std::max(1, std::min(lanes, (int) tvl1_ctx::kMaxLanes));
    int first = 0;
    auto push = [&](int b) { out.emplace_back(first, b); first += b; };
    // (npairs % small != 0 would add a third, ragged size: a lane keeps workspaces for two)
    if (small >= 1 && Bmax % small == 0 && npairs % small == 0 && npairs >= 2 * nlanes * small + nlanes * Bmax) {
        for (int l = 0; l < nlanes; l++) push(small);
        int rest = npairs - first - nlanes * small;
        while (rest >= Bmax) { push(Bmax); rest -= Bmax; }
        while (rest >= small) { push(small); rest -= small; }
        if (rest > 0) push(rest);
        for (int l = 0; l < nlanes; l++) push(small);
    } else {
        while (first < npairs) push(std::min(Bmax, npairs - first));
    }
}

// Chunk sizes for the pipelined host-buffer call (solve_host_pipelined).  The call cannot start computing before
// its first chunk has arrived and cannot return before its last chunk has been downloaded, while chunks solve
// at the full rate only when they are large: so the sizes RAMP -- Bmax/8, Bmax/4, Bmax/2 at the start, full
// chunks in the middle, Bmax/2, Bmax/4, Bmax/8 at the end (uploads run ahead of the kernels: the copy engines
// move a pair faster than the SMs solve it).  At most four sizes plus one ragged remainder occur, which is what
// a lane keeps workspaces and solve graphs for (tvl1_ctx::kAltWs).
// `min_chunk`: no ramp step below this many pairs.  Very small lock-step chunks make very short while-loop bodies, several
// of them concurrently on several lanes: the one configuration in which an intermittent launch failure was seen.
constexpr int kRampMinChunk = 8;
void ramp_schedule(int npairs, int Bmax, const std::vector<int> &override_sizes, std::vector<std::pair<int, int>> &out,
                   int min_chunk = kRampMinChunk)
{
    int first = 0;
    auto push = [&](int b) { if (b > 0) { out.emplace_back(first, b); first += b; } };
    if (!override_sizes.empty()) {
        for (int b : override_sizes) push(std::min(std::min(std::max(b, 1), Bmax), npairs - first));
        while (first < npairs) push(std::min(Bmax, npairs - first));
        return;
    }
    std::vector<int> ramp;                          // ascending, distinct, below Bmax
    for (int d = 8; d >= 2; d /= 2)
        if (Bmax / d >= std::max(1, min_chunk) && (ramp.empty() || ramp.back() != Bmax / d)) ramp.push_back(Bmax / d);
    auto sum = [&]() { int t = 0; for (int b : ramp) t += b; return t; };
    while (!ramp.empty() && npairs < 2 * sum() + Bmax) ramp.erase(ramp.begin());      // small batches: shorter ramps
    for (int b : ramp) push(b);
    int middle = npairs - first - sum();
    while (middle >= Bmax) { push(Bmax); middle -= Bmax; }
    // what is left of the middle joins the descending ramp (largest first), a ragged remainder goes last
    std::vector<int> tail(ramp.rbegin(), ramp.rend());
    for (int i = (int) ramp.size() - 1; i >= 0; i--)
        if (middle >= ramp[i]) { tail.push_back(ramp[i]); middle -= ramp[i]; }
    std::sort(tail.begin(), tail.end(), [](int a, int b) { return a > b; });
    for (int b : tail) push(b);
    push(middle);
}

int ensure_slots(tvl1_ctx *ctx, int nslots, size_t in_bytes, size_t out_bytes)
{
    if (!ctx->up_stream) CK(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    if (!ctx->down_stream) CK(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    const bool grow = ctx->slot_in_bytes < in_bytes || ctx->slot_out_bytes < out_bytes;
    if (grow) {
        for (auto &sl : ctx->slots)
            for (int i = 0; i < 2; i++) {
                cudaFree(sl.in[i]); sl.in[i] = nullptr;
                cudaFree(sl.out[i]); sl.out[i] = nullptr;
            }
        ctx->slot_in_bytes = std::max(ctx->slot_in_bytes, in_bytes);
        ctx->slot_out_bytes = std::max(ctx->slot_out_bytes, out_bytes);
    }
    for (int k = 0; k < nslots; k++) {
        tvl1_ctx::HostSlot &sl = ctx->slots[k];
        for (int i = 0; i < 2; i++) {
            if (!sl.in[i]) CK(cudaMalloc(&sl.in[i], ctx->slot_in_bytes));
            if (!sl.out[i]) CK(cudaMalloc(&sl.out[i], ctx->slot_out_bytes));
        }
        if (!sl.up) CK(cudaEventCreateWithFlags(&sl.up, cudaEventDisableTiming));
        if (!sl.done) CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        if (!sl.down) CK(cudaEventCreateWithFlags(&sl.down, cudaEventBlockingSync | cudaEventDisableTiming));
        sl.down_recorded = false;
    }
    return TVL1_OK;
}

int ensure_stage_f32(tvl1_ctx *ctx, size_t bytes_each)
{
    if (ctx->stage_f32_bytes < bytes_each) {
        for (int i = 0; i < 4; i++) { cudaFree(ctx->stage_f32[i]); ctx->stage_f32[i] = nullptr; }
        ctx->stage_f32_bytes = 0;
        for (int i = 0; i < 4; i++) CK(cudaMalloc(&ctx->stage_f32[i], bytes_each));
        ctx->stage_f32_bytes = bytes_each;
    }
    return TVL1_OK;
}

// Pinned (or registered) host buffers: the whole call is ONE three-stage pipeline.
//   upload    the chunks cross PCIe in order on one copy stream, each into a slot of a small ring of device
//             buffers, up to `nslots` chunks ahead of the kernels (in order on one stream: the first chunk is not
//             slowed by the second, which is what happens when every lane copies on a stream of its own);
//   solve     the lanes (host thread + stream + workspace each, as in run_lanes) take the chunks in order; a lane
//             waits for the chunk's upload event, solves it in place in the slot and is never idle for a copy;
//   download  results leave in order on a second copy stream, behind the lane's completion event.
// With the ramped chunk sizes of ramp_schedule the GPU starts after the upload of an eighth of a full chunk and
// the call ends one such download after the last kernel.
// TI: element type of the host images (float, double, or unsigned char for 8-bit frame sequences: bytes over PCIe,
// widened on the device); TO: element type of the host flows (float or double).
template <typename TI, typename TO>
int solve_host_pipelined(tvl1_ctx *ctx, const std::vector<std::pair<int, int>> &chunks, const TI *I0, const TI *I1,
                         TO *u1, TO *u2, int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    const bool sequence = I1 == nullptr;
    constexpr bool in_f32 = std::is_same<TI, float>::value, in_f64 = std::is_same<TI, double>::value;
    constexpr bool out_f64 = std::is_same<TO, double>::value;
    const size_t n = (size_t) nx * ny;
    const int nchunks = (int) chunks.size();
    const int nstat = prm->nscales * prm->warps;
    int Bmax = 0;
    for (const auto &c : chunks) Bmax = std::max(Bmax, c.second);
    tvl1_ctx *lanes[tvl1_ctx::kMaxLanes];
    int nlanes = 1;
    TRY(setup_lanes(ctx, nchunks, std::min(ctx->host_lanes, ctx->pipe_lanes), lanes, &nlanes));
    const int nslots = std::min(nchunks, std::min(nlanes + 3, (int) tvl1_ctx::kMaxSlots));
    const size_t in_frames = (size_t) Bmax + (sequence ? 1 : 0);
    TRY(ensure_slots(ctx, nslots, in_frames * n * sizeof(TI), (size_t) Bmax * n * sizeof(TO)));
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> slot_of(nchunks, -1), free_slots;
    for (int k = nslots - 1; k >= 0; k--) free_slots.push_back(k);
    std::vector<char> slot_used(nslots, 0);
    int uploaded = 0, claimed = 0;
    bool failed = false;
    // (mu held) issue uploads, in chunk order, while a slot is free.  A slot goes back on the free list as soon as its
    // chunk has been solved, whichever chunk that is: a chunk that takes long (one slow pair keeps its lock-step
    // batch iterating) must not hold up the uploads behind it.
    auto pump = [&]() -> cudaError_t {
        while (uploaded < nchunks && !free_slots.empty()) {
            const int first = chunks[uploaded].first, B = chunks[uploaded].second;
            const int si = free_slots.back();
            free_slots.pop_back();
            slot_of[uploaded] = si;
            tvl1_ctx::HostSlot &sl = ctx->slots[si];
            const size_t off = (size_t) first * n, cnt = (size_t) B * n;
            // (the slot's previous chunk has been solved: its lane said so under `mu`; the wait orders the copy
            // engine behind that lane's stream whatever the solver's own host waits are)
            cudaError_t e = slot_used[si] ? cudaStreamWaitEvent(ctx->up_stream, sl.done, 0) : cudaSuccess;
            slot_used[si] = 1;
            if (e != cudaSuccess) return e;
            if (sequence)
                e = cudaMemcpyAsync(sl.in[0], I0 + off, (cnt + n) * sizeof(TI), cudaMemcpyHostToDevice, ctx->up_stream);
            else {
                e = cudaMemcpyAsync(sl.in[0], I0 + off, cnt * sizeof(TI), cudaMemcpyHostToDevice, ctx->up_stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(sl.in[1], I1 + off, cnt * sizeof(TI), cudaMemcpyHostToDevice, ctx->up_stream);
            }
            if (e == cudaSuccess) e = cudaEventRecord(sl.up, ctx->up_stream);
            if (e != cudaSuccess) return e;
            uploaded++;
        }
        return cudaSuccess;
    };
    {
        std::lock_guard<std::mutex> lk(mu);
        CK(pump());
    }
    int rcs[tvl1_ctx::kMaxLanes] = {};
    const ChunkTrace trace;
    auto chunk = [&](tvl1_ctx *c, int k) -> int {
        const double t_start = trace.ms();
        tvl1_ctx *ctx_root = ctx;
        tvl1_ctx *ctx = c;   // for CK / TRY
        const int first = chunks[k].first, B = chunks[k].second;
        const size_t cnt = (size_t) B * n;
        const int si = slot_of[k];
        tvl1_ctx::HostSlot &sl = ctx_root->slots[si];
        cudaStream_t st = c->stream;
        if (!in_f32 || out_f64) TRY(ensure_stage_f32(c, in_frames * n * sizeof(float)));
        CK(cudaStreamWaitEvent(st, sl.up, 0));
        if (sl.down_recorded) CK(cudaStreamWaitEvent(st, sl.down, 0));    // the slot's previous results have left
        float *d0, *d1, *o0, *o1;
        const unsigned g = (unsigned) std::min<size_t>((cnt + 255) / 256, 4096);
        if (in_f32) {
            d0 = (float *) sl.in[0]; d1 = sequence ? d0 + n : (float *) sl.in[1];
        } else {                         // narrowed (fp64) or widened (8-bit) on the device, into the lane's own buffers
            d0 = c->stage_f32[0]; d1 = sequence ? d0 + n : c->stage_f32[1];
            const size_t cnt0 = sequence ? cnt + n : cnt;
            if (in_f64) k_f64_to_f32<<<g, 256, 0, st>>>((const double *) sl.in[0], d0, cnt0);
            else k_u8_to_f32<<<(unsigned) std::min<size_t>((cnt0 / 4 + 255) / 256 + 1, 4096), 256, 0, st>>>((const unsigned char *) sl.in[0], d0, cnt0);
            CKL(ctx);
            if (!sequence) {
                if (in_f64) k_f64_to_f32<<<g, 256, 0, st>>>((const double *) sl.in[1], d1, cnt);
                else k_u8_to_f32<<<(unsigned) std::min<size_t>((cnt / 4 + 255) / 256 + 1, 4096), 256, 0, st>>>((const unsigned char *) sl.in[1], d1, cnt);
                CKL(ctx);
            }
        }
        if (out_f64) { o0 = c->stage_f32[2]; o1 = c->stage_f32[3]; }
        else { o0 = (float *) sl.out[0]; o1 = (float *) sl.out[1]; }
        int *it = iters_out ? iters_out + (size_t) first * nstat : nullptr;
        double *er = errs_out ? errs_out + (size_t) first * nstat : nullptr;
        if (c->hs_mode) TRY(run_hs_multiscale(c, B, d0, d1, o0, o1, nx, ny, c->hs, it, er));
        else TRY(run_multiscale(c, B, d0, d1, o0, o1, nx, ny, *prm, it, er));
        if (out_f64) {
            k_f32_to_f64<<<g, 256, 0, st>>>(o0, (double *) sl.out[0], cnt);
            CKL(ctx);
            k_f32_to_f64<<<g, 256, 0, st>>>(o1, (double *) sl.out[1], cnt);
            CKL(ctx);
        }
        CK(cudaEventRecord(sl.done, st));
        const double t_solved = trace.ms();
        std::lock_guard<std::mutex> lk(mu);
        const size_t off = (size_t) first * n;
        CK(cudaStreamWaitEvent(ctx_root->down_stream, sl.done, 0));
        CK(cudaMemcpyAsync(u1 + off, sl.out[0], cnt * sizeof(TO), cudaMemcpyDeviceToHost, ctx_root->down_stream));
        CK(cudaMemcpyAsync(u2 + off, sl.out[1], cnt * sizeof(TO), cudaMemcpyDeviceToHost, ctx_root->down_stream));
        CK(cudaEventRecord(sl.down, ctx_root->down_stream));
        sl.down_recorded = true;
        free_slots.push_back(si);
        CK(pump());
        trace.line("pipe", c, k, B, t_start, t_solved, trace.ms());
        return TVL1_OK;
    };
    const int rc = join_lanes(ctx, lanes, nlanes, rcs, [&](int l) {
        tvl1_ctx *c = lanes[l];
        cudaSetDevice(c->device);
        for (;;) {
            int k;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (failed || claimed >= nchunks) break;
                k = claimed++;
                cv.wait(lk, [&] { return failed || uploaded > k; });
                if (failed) break;
            }
            rcs[l] = chunk(c, k);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (rcs[l] != TVL1_OK) failed = true;
            }
            cv.notify_all();
            if (rcs[l] != TVL1_OK) break;
        }
        resolve_events(c);
    });
    // the last downloads
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < nslots; k++)
        if (ctx->slots[k].down_recorded) {
            const cudaError_t ek = cudaEventSynchronize(ctx->slots[k].down);
            if (e == cudaSuccess) e = ek;
        }
    if (rc != TVL1_OK) { cudaStreamSynchronize(ctx->up_stream); return rc; }
    CK(e);
    return TVL1_OK;
}

// Host-buffer driver shared by the f32/f64, multiscale/single-scale entry points.  A batch larger
// than max_batch is cut into chunks; two lanes (this context and a private sibling on the same
// GPU, each with its own stream, workspace and host thread) take alternate chunks, so the H2D/D2H
// copies of one chunk overlap the kernels of the other.
template <typename T>
int solve_host(tvl1_ctx *ctx, int npairs, const T *I0, const T *I1, T *u1, T *u2, int nx, int ny,
               const tvl1_params *prm, int *iters_out, double *errs_out, bool multiscale)
{
    TRY(check_common(ctx, I0, I1 ? I1 : I0, u1, u2, nx, ny, prm, multiscale));
    if (npairs < 1) return fail_arg(ctx, "npairs must be >= 1");
    reset_stats(ctx);
    const bool f64 = sizeof(T) == 8;
    const size_t n = (size_t) nx * ny;
    const int Bmax = std::min(npairs, ctx->max_batch);
    const int nstat = (multiscale ? prm->nscales : 1) * prm->warps;
    const size_t stage_frames = (size_t) Bmax + (I1 ? 0 : 1);     // frame sequence: B+1 frames per chunk
    std::vector<std::pair<int, int>> chunks;                      // (first pair, pairs)
    // pinned (or registered) buffers of a batch of several chunks: one call-wide pipeline with ramped chunk sizes
    if (multiscale && ctx->host_pipe && !ctx->hs_mode && !ctx->is_sibling && npairs > Bmax && !is_pageable(I0) && (!I1 || !is_pageable(I1)) &&
        !is_pageable(u1) && !is_pageable(u2)) {
        ramp_schedule(npairs, Bmax, ctx->chunk_override, chunks, ctx->ramp_min_chunk);
        return solve_host_pipelined<T, T>(ctx, chunks, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out);
    }
    chunk_schedule(ctx, npairs, Bmax, ctx->host_lanes, chunks);
    const ChunkTrace trace;
    return run_lanes(ctx, (int) chunks.size(), ctx->host_lanes, [&](tvl1_ctx *c, int k) -> int {
        tvl1_ctx *ctx = c;   // for CK / TRY
        const double t_start = trace.ms();
        TRY(ensure_stage(ctx, stage_frames * n * sizeof(T), f64));
        const int rc = solve_chunk<T>(ctx, chunks[k].first, chunks[k].second, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out,
                                      multiscale, nstat);
        trace.line("lane", c, k, chunks[k].second, t_start, trace.ms(), trace.ms());
        return rc;
    });
}


// 8-bit frame sequence (video): frames [nframes][ny][nx] of unsigned char in HOST memory -> nframes - 1 flows
// in fp32.  A quarter of the upload of the fp32 form; the conversion is exact, so the flows are the bits
// tvl1_solve_sequence_f32 gives on the same values.
int solve_chunk_u8(tvl1_ctx *ctx, int first, int B, const unsigned char *frames, float *u1, float *u2, int nx, int ny,
                   const tvl1_params *prm, int *iters_out, double *errs_out, int nstat)
{
    const size_t n = (size_t) nx * ny;
    cudaStream_t st = ctx->stream;
    const size_t in_cnt = (size_t) (B + 1) * n, out_cnt = (size_t) B * n, off = (size_t) first * n;
    unsigned char *d8 = (unsigned char *) ctx->stage_in[1];
    float *d0 = (float *) ctx->stage_in[0], *o0 = (float *) ctx->stage_out[0], *o1 = (float *) ctx->stage_out[1];
    CK(cudaMemcpyAsync(d8, frames + off, in_cnt, cudaMemcpyHostToDevice, st));
    k_u8_to_f32<<<(unsigned) std::min<size_t>((in_cnt / 4 + 255) / 256 + 1, 4096), 256, 0, st>>>(d8, d0, in_cnt);
    CKL(ctx);
    TRY(run_multiscale(ctx, B, d0, d0 + n, o0, o1, nx, ny, *prm, iters_out ? iters_out + (size_t) first * nstat : nullptr,
                       errs_out ? errs_out + (size_t) first * nstat : nullptr));
    CK(cudaMemcpyAsync(u1 + off, o0, out_cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(u2 + off, o1, out_cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(sleep_until_done(ctx));
    return TVL1_OK;
}

int solve_sequence_u8(tvl1_ctx *ctx, int nframes, const unsigned char *frames, float *u1, float *u2, int nx, int ny,
                      const tvl1_params *prm, int *iters_out, double *errs_out)
{
    TRY(check_common(ctx, frames, frames, u1, u2, nx, ny, prm, true));
    if (nframes < 2) return fail_arg(ctx, "a frame sequence needs at least 2 frames");
    reset_stats(ctx);
    const int npairs = nframes - 1;
    const size_t n = (size_t) nx * ny;
    const int Bmax = std::min(npairs, ctx->max_batch);
    const int nstat = prm->nscales * prm->warps;
    std::vector<std::pair<int, int>> chunks;
    if (ctx->host_pipe && !ctx->is_sibling && npairs > Bmax && !is_pageable(frames) && !is_pageable(u1) && !is_pageable(u2)) {
        ramp_schedule(npairs, Bmax, ctx->chunk_override, chunks, ctx->ramp_min_chunk);
        return solve_host_pipelined<unsigned char, float>(ctx, chunks, frames, nullptr, u1, u2, nx, ny, prm, iters_out, errs_out);
    }
    chunk_schedule(ctx, npairs, Bmax, ctx->host_lanes, chunks);
    return run_lanes(ctx, (int) chunks.size(), ctx->host_lanes, [&](tvl1_ctx *c, int k) -> int {
        tvl1_ctx *ctx = c;   // for CK / TRY
        TRY(ensure_stage(ctx, (size_t) (Bmax + 1) * n * sizeof(float), false));
        return solve_chunk_u8(ctx, chunks[k].first, chunks[k].second, frames, u1, u2, nx, ny, prm, iters_out, errs_out, nstat);
    });
}

// =================================================================================================
// Row-band mode: ONE image pair split over the GPUs of a box (BASELINE configs[3..4]).
// Every rank holds the full images and builds the full pyramids (cheap); coarse levels are solved
// redundantly and identically by every rank; a level with at least `min_split_rows` rows is cut into
// contiguous row bands.  Per primal-dual iteration of a split level a rank
//   1. runs the fused iteration kernel on its own rows (halo rows are ordinary rows of its
//      full-size planes),
//   2. all-reduces the sum of squared updates (one double) and exchanges the 1-row halos with its
//      band neighbours over NCCL send/recv (NVLink): row r0 of {u1,u2,p11,p12,p21,p22} goes up,
//      row r1-1 of {p12,p22} goes down -- both ping-pong sets, so the exchange does not depend on
//      the device-side parity,
//   3. applies the stopping rule of src/tvl1flow.cpp:113 to the reduced error (k_band_decide), the
//      same number on every rank.
// After a split level the flow bands are all-gathered in place so that every rank can up-sample
// (zoom_in) its band of the next level, and the caller of every rank receives the full flow.
// NCCL is bound at run time (dlopen) so that the library has no link-time dependency on it.
// =================================================================================================
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void *lib = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

const char *load_nccl()
{
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (g_nccl.lib) return nullptr;
    // The NCCL the process already uses comes first (a host program or torch.distributed has usually
    // loaded one; a second copy of another version under the same soname would break whoever binds
    // later), then an explicit path (TVL1_NCCL_LIB), then the system library.  Never RTLD_GLOBAL: our
    // symbols are looked up with dlsym and nobody else should resolve against what we happened to load.
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) if (const char *path = std::getenv("TVL1_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return "libnccl.so.2 not found";
#define TVL1_SYM(field, name)                                                        \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));         \
    if (!g_nccl.field) return "missing NCCL symbol " name
    TVL1_SYM(GetUniqueId, "ncclGetUniqueId");
    TVL1_SYM(CommInitRank, "ncclCommInitRank");
    TVL1_SYM(CommDestroy, "ncclCommDestroy");
    TVL1_SYM(Send, "ncclSend");
    TVL1_SYM(Recv, "ncclRecv");
    TVL1_SYM(AllReduce, "ncclAllReduce");
    TVL1_SYM(AllGather, "ncclAllGather");
    TVL1_SYM(GroupStart, "ncclGroupStart");
    TVL1_SYM(GroupEnd, "ncclGroupEnd");
    TVL1_SYM(GetErrorString, "ncclGetErrorString");
#undef TVL1_SYM
    g_nccl.lib = h;
    return nullptr;
}

#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            ctx->err = std::string(#call " failed: ") + g_nccl.GetErrorString(r_);                 \
            return TVL1_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

void band_rows(int ny, int rank, int world, int *r0, int *r1, int *rows_per)
{
    *rows_per = ceil_div(ny, world);
    *r0 = std::min(ny, rank * *rows_per);
    *r1 = std::min(ny, *r0 + *rows_per);
}


// Sum of a few host doubles over the ranks of the band communicator (collective; used to take
// identical decisions on every rank).
int agree_sum(tvl1_ctx *ctx, double *vals, int n)
{
    if (n > 8) return fail_arg(ctx, "agree_sum: too many values");
    if (!ctx->d_agree) CK(cudaMalloc(&ctx->d_agree, 8 * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_agree, vals, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NK(g_nccl.AllReduce(ctx->d_agree, ctx->d_agree, n, ncclDouble, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->stream));
    CK(cudaMemcpyAsync(vals, ctx->d_agree, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TVL1_OK;
}

// ---- peer-memory plumbing: CUDA IPC handles travel through one NCCL all-gather ------------------
int exchange_ipc_handles(tvl1_ctx *ctx, void *dev_ptr, std::vector<cudaIpcMemHandle_t> &all)
{
    const int G = ctx->band_world;
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, dev_ptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    CK(cudaMemcpyAsync(ctx->d_handles + 64 * ctx->band_rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    NK(g_nccl.AllGather(ctx->d_handles + 64 * ctx->band_rank, ctx->d_handles, 64, ncclChar,
                        (ncclComm_t) ctx->nccl_comm, ctx->stream));
    all.resize(G);
    CK(cudaMemcpyAsync(all.data(), ctx->d_handles, 64 * G, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TVL1_OK;
}

void p2p_release_state(tvl1_ctx *ctx)
{
    for (auto &g : ctx->level_sg) free_graph(g, ctx->ev_pool);   // peer pointers are baked into them
    for (int r = 0; r < kMaxRanks; r++) {
        if (ctx->peer_state[r] && r != ctx->band_rank) cudaIpcCloseMemHandle(ctx->peer_state[r]);
        ctx->peer_state[r] = nullptr;
    }
    ctx->p2p_state_key = nullptr;
}

int p2p_setup_mailboxes(tvl1_ctx *ctx)
{
    const int G = ctx->band_world, me = ctx->band_rank;
    ctx->p2p_ready = false;
    if (G > kMaxRanks) return TVL1_OK;                         // NCCL variant only
    if (!ctx->d_handles) CK(cudaMalloc(&ctx->d_handles, 64 * kMaxRanks));
    if (!ctx->my_box) CK(cudaMalloc(&ctx->my_box, 2 << 20));   // its own allocation: exported whole
    CK(cudaMemsetAsync(ctx->my_box, 0, sizeof(BandMailbox), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->d_gather_ticket) {
        CK(cudaMalloc(&ctx->d_gather_ticket, sizeof(unsigned int)));
        CK(cudaMemsetAsync(ctx->d_gather_ticket, 0, sizeof(unsigned int), ctx->stream));
    }
    for (int r = 0; r < kMaxRanks; r++) {                      // mappings of an earlier tvl1_band_init
        if (ctx->boxes[r] && ctx->boxes[r] != ctx->my_box) cudaIpcCloseMemHandle(ctx->boxes[r]);
        ctx->boxes[r] = nullptr;
    }
    std::vector<cudaIpcMemHandle_t> all;
    TRY(exchange_ipc_handles(ctx, ctx->my_box, all));
    bool ok = true;
    for (int r = 0; r < G && ok; r++) {
        if (r == me) { ctx->boxes[r] = ctx->my_box; continue; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;                                        // no peer access from this rank
            break;
        }
        ctx->boxes[r] = (BandMailbox *) p;
    }
    // The ranks must agree: one rank on the mailboxes while another waits in NCCL would hang the job.
    // Peer memory is used only if EVERY rank mapped every mailbox; the NCCL-per-iteration switch is
    // taken if ANY rank asks for it.
    double flags[2] = { ok ? 0.0 : 1.0, ctx->band_use_nccl_per_iteration ? 1.0 : 0.0 };
    TRY(agree_sum(ctx, flags, 2));
    ctx->band_use_nccl_per_iteration = flags[1] > 0.0;
    if (flags[0] > 0.0) {                                      // somebody failed: everybody stays on NCCL
        for (int r = 0; r < kMaxRanks; r++) {
            if (ctx->boxes[r] && ctx->boxes[r] != ctx->my_box) cudaIpcCloseMemHandle(ctx->boxes[r]);
            ctx->boxes[r] = nullptr;
        }
        return TVL1_OK;
    }
    ctx->p2p_ready = true;
    return TVL1_OK;
}

// (re)bind the neighbours' state buffers after the workspace was (re)allocated -- collective
int p2p_bind_state(tvl1_ctx *ctx)
{
    if (!ctx->p2p_ready || ctx->p2p_state_key != nullptr) return TVL1_OK;   // released <=> must (re)bind
    std::vector<cudaIpcMemHandle_t> all;
    TRY(exchange_ipc_handles(ctx, ctx->ws.state, all));
    for (int r = 0; r < ctx->band_world; r++) {     // neighbours for the halos, everybody for the all-gather
        if (r == ctx->band_rank) { ctx->peer_state[r] = ctx->ws.state; continue; }
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_state[r] = (float *) p;
    }
    ctx->p2p_state_key = ctx->ws.state;
    return TVL1_OK;
}

BandPeers band_peers(const tvl1_ctx *ctx, const Level &l, int r0, int r1)
{
    BandPeers pp = {};
    pp.enabled = 1; pp.rank = ctx->band_rank; pp.world = ctx->band_world;
    pp.halo = kTbT;
    pp.up_state = (r0 > 0 && ctx->band_rank > 0) ? ctx->peer_state[ctx->band_rank - 1] : nullptr;
    pp.dn_state = (r1 < l.ny && ctx->band_rank + 1 < ctx->band_world) ? ctx->peer_state[ctx->band_rank + 1] : nullptr;
    for (int r = 0; r < ctx->band_world; r++) pp.box[r] = ctx->boxes[r];
    return pp;
}

// One split level with the exchange fused into the iteration kernels (no NCCL, no extra kernels, no
// host round trip).  A band keeps kTbT rows beyond each of its edges current -- pushed by the
// neighbours after every accepted iteration / block of iterations -- so the temporally blocked kernel
// runs inside the bands exactly as on one GPU (T-row halos every T iterations), and the
// constants of those rows are computed locally (the warp is per-pixel work on replicated pyramids).
// The level ends with the all-gather of the flow through peer memory.
int band_level_p2p(tvl1_ctx *ctx, int s, const tvl1_params &prm, int stat_base, int &hint)
{
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    int r0, r1, rows_per;
    band_rows(l.ny, ctx->band_rank, ctx->band_world, &r0, &r1, &rows_per);
    const BandPeers pp = band_peers(ctx, l, r0, r1);
    const int h0 = std::max(r0 - pp.halo, 0), h1 = std::min(r1 + pp.halo, l.ny);
    TRY(launch_zero(ctx, s, 1, F_P11, 4));
    for (int wi = 0; wi < prm.warps; wi++) {
        {
            Span sp(ctx, 1);
            TRY(launch_warp(ctx, s, 1, 0, h0, h1));                     // own rows + both halos
        }
        k_begin_warp<<<1, 32, 0, ctx->stream>>>(w.ctl, w.loop, 1);
        CKL(ctx);
        IterParams P = iter_params(ctx, l, prm, stat_base + wi, kMaxIterations, s);
        P.row_begin = r0; P.row_end = r1;
        P.peers = pp;
        P.tb = (r1 - r0 >= kTbBH && tb_usable(ctx, l, 1)) ? 1 : 0;
        TRY(run_iterations(ctx, P, 1, hint));
    }
    GatherParams G = {};
    for (int r = 0; r < ctx->band_world; r++) G.state[r] = ctx->peer_state[r];
    G.peers = pp; G.ctl = w.ctl; G.ticket = ctx->d_gather_ticket;
    G.set_stride = w.set_stride; G.field_stride = w.field_stride;
    G.pitch = l.pitch; G.row_begin = r0; G.row_end = r1;
    const size_t n4 = (size_t) (r1 - r0) * l.pitch / 4;
    const int blocks = (int) std::min<size_t>(std::max<size_t>((2 * n4 + 1023) / 1024, 1), (size_t) 4 * ctx->sm_count);
    k_band_allgather<<<blocks, 256, 0, ctx->stream>>>(G);
    CKL(ctx);
    return TVL1_OK;
}

// all-reduce of the error sum + halo exchange of both ping-pong sets, one NCCL group
int band_exchange(tvl1_ctx *ctx, int s, bool with_sum)
{
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    ncclComm_t comm = (ncclComm_t) ctx->nccl_comm;
    cudaStream_t st = ctx->stream;
    int r0, r1, rows_per;
    band_rows(l.ny, ctx->band_rank, ctx->band_world, &r0, &r1, &rows_per);
    const int up = ctx->band_rank - 1, dn = ctx->band_rank + 1;
    const bool has_up = up >= 0 && r0 > 0 && r0 < l.ny, has_dn = dn < ctx->band_world && r1 < l.ny && r1 > r0;
    const size_t cnt = l.pitch;
    auto row = [&](int set, int f, int y) { return w.state + (size_t) set * w.set_stride + (size_t) f * w.field_stride + (size_t) y * l.pitch; };
    NK(g_nccl.GroupStart());
    if (with_sum) NK(g_nccl.AllReduce(ctx->d_band_sum, ctx->d_band_sum, 1, ncclDouble, ncclSum, comm, st));
    for (int set = 0; set < 2; set++) {
        if (has_dn) {   // link (me, me+1): my last row of p12,p22 goes down; its first row of everything comes up
            NK(g_nccl.Send(row(set, F_P12, r1 - 1), cnt, ncclFloat, dn, comm, st));
            NK(g_nccl.Send(row(set, F_P22, r1 - 1), cnt, ncclFloat, dn, comm, st));
            for (int f = 0; f < F_COUNT; f++) NK(g_nccl.Recv(row(set, f, r1), cnt, ncclFloat, dn, comm, st));
        }
        if (has_up) {   // link (me-1, me)
            NK(g_nccl.Recv(row(set, F_P12, r0 - 1), cnt, ncclFloat, up, comm, st));
            NK(g_nccl.Recv(row(set, F_P22, r0 - 1), cnt, ncclFloat, up, comm, st));
            for (int f = 0; f < F_COUNT; f++) NK(g_nccl.Send(row(set, f, r0), cnt, ncclFloat, up, comm, st));
        }
    }
    NK(g_nccl.GroupEnd());
    return TVL1_OK;
}

int band_read_ctl(tvl1_ctx *ctx, PairCtl *out)
{
    CK(cudaMemcpyAsync(out, ctx->ws.ctl, sizeof(PairCtl), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.host_syncs++;
    return TVL1_OK;
}

// one split level: src/tvl1flow.cpp:46-212 on this rank's rows
int band_level(tvl1_ctx *ctx, int s, const tvl1_params &prm, int stat_base, int &hint)
{
    const Workspace &w = ctx->ws;
    const Level &l = w.lv[s];
    cudaStream_t st = ctx->stream;
    int r0, r1, rows_per;
    band_rows(l.ny, ctx->band_rank, ctx->band_world, &r0, &r1, &rows_per);
    TRY(launch_zero(ctx, s, 1, F_P11, 4));
    for (int wi = 0; wi < prm.warps; wi++) {
        if (r1 > r0) {
            Span sp(ctx, 1);
            TRY(launch_warp(ctx, s, 1, 0, r0, std::min(r1 + 1, l.ny)));   // + the halo row below
        }
        k_begin_warp<<<1, 32, 0, st>>>(w.ctl, w.loop, 1);
        CKL(ctx);
        IterParams P = iter_params(ctx, l, prm, stat_base + wi, kMaxIterations, s);
        P.row_begin = r0; P.row_end = r1; P.band_sum = ctx->d_band_sum;
        int launched = 0, chunk = std::max(1, std::min(hint, P.max_iter));
        while (launched < P.max_iter) {
            const int k = std::min(chunk, P.max_iter - launched);
            {
                Span sp(ctx, 0, P.level);
                for (int i = 0; i < k; i++) {
                    CK(cudaMemsetAsync(ctx->d_band_sum, 0, sizeof(double), st));
                    if (r1 > r0) TRY(launch_iterate(ctx, P, 1));
                    TRY(band_exchange(ctx, s, true));
                    k_band_decide<<<1, 32, 0, st>>>(w.ctl, w.loop, ctx->d_band_sum, (double) l.nx * (double) l.ny,
                                                    P.eps2, P.max_iter, w.stat_iters, w.stat_errs, P.stat_slot,
                                                    w.counters + P.level,
                                                    (unsigned long long) l.nx * (unsigned long long) (r1 - r0));
                    CKL(ctx);
                }
            }
            launched += k;
            CK(cudaMemcpyAsync(ctx->h_loop, w.loop, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ctx->stats.host_syncs++;
            if (ctx->h_loop->active_pairs == 0) break;
            chunk = std::max(1, hint / 8);
        }
        hint = std::max(1, ctx->h_loop->max_n);
    }
    return TVL1_OK;
}

int run_band(tvl1_ctx *ctx, const float *dI0, const float *dI1, float *du1, float *du2, int nx, int ny,
             const tvl1_params &prm, int min_split_rows, int *iters_out, double *errs_out)
{
    const int ns = prm.nscales, nstat = ns * prm.warps, G = ctx->band_world;
    cudaStream_t st = ctx->stream;
    ncclComm_t comm = (ncclComm_t) ctx->nccl_comm;
    // This context's workspace is used by band solves only (see band_context), and those are
    // collective: every rank takes the same (re)allocation decision here.  Before anybody frees a
    // buffer that neighbours have mapped, all ranks drop their mappings and meet at a barrier.
    if (!workspace_matches(ctx, ctx->ws, nx, ny, ns, prm.zfactor, 1, nstat, G)) {
        p2p_release_state(ctx);
        NK(g_nccl.AllReduce(ctx->d_band_sum, ctx->d_band_sum, 1, ncclDouble, ncclSum, comm, st));
        CK(cudaStreamSynchronize(st));
    }
    TRY(ensure_workspace(ctx, nx, ny, ns, prm.zfactor, 1, nstat, G));
    Workspace &w = ctx->ws;
    const bool p2p = ctx->p2p_ready && !ctx->band_use_nccl_per_iteration;
    if (ctx->p2p_ready) TRY(p2p_bind_state(ctx));
    // a negative threshold also splits on a single rank (one band = the whole level): the band code
    // path without neighbours, used by the single-GPU tests
    const bool force = min_split_rows < 0;
    const int halo = p2p ? kTbT : 1;
    min_split_rows = std::max(std::abs(min_split_rows), 2 * halo * G);
    // every band, the last (shortest) one included, must hold at least `halo` rows
    auto is_split = [&](int s) {
        const int ny_s = w.lv[s].ny;
        return (G > 1 || force) && ny_s >= min_split_rows && ny_s - (G - 1) * ceil_div(ny_s, G) >= halo;
    };
    Span total(ctx, 2);
    if (p2p && G > 1) {
        // nobody pushes halos into a rank that is still reading the previous solve's result
        BandPeers pp = {};
        pp.enabled = 1; pp.rank = ctx->band_rank; pp.world = G;
        for (int r = 0; r < G; r++) pp.box[r] = ctx->boxes[r];
        k_band_barrier<<<1, 32, 0, st>>>(pp);
        CKL(ctx);
    }
    TRY(build_pyramid(ctx, 1, dI0, dI1, nx, ny, prm));
    TRY(launch_zero(ctx, ns - 1, 1, F_U1, 2));
    int hint = 16;
    for (int s = ns - 1; s >= 0; s--) {
        const int stat_base = (ns - 1 - s) * prm.warps;
        const int ls = std::min(s, TVL1_MAX_LEVELS - 1);
        if (is_split(s) && !p2p) {
            TRY(band_level(ctx, s, prm, stat_base, hint));         // NCCL calls per iteration: host-driven
        } else if (ctx->use_graph && s < TVL1_MAX_LEVELS) {
            // one graph per level (loops are conditional WHILE nodes; every rank replays the same
            // number of iterations because every rank takes the same stop decision)
            const long long variant = 1 + (is_split(s) ? 1 : 0) + 2ll * G + 64ll * ctx->band_rank + 4096ll * min_split_rows;
            if (is_split(s))
                TRY(replay_graph(ctx, ctx->level_sg[ls], prm, true, variant,
                                 [&]() { int h = 16; return band_level_p2p(ctx, s, prm, stat_base, h); }));
            else
                TRY(replay_graph(ctx, ctx->level_sg[ls], prm, true, variant,
                                 [&]() { int h = 16; return run_level(ctx, s, 1, prm, stat_base, h); }));
        } else if (is_split(s)) {
            TRY(band_level_p2p(ctx, s, prm, stat_base, hint));
        } else {
            TRY(run_level(ctx, s, 1, prm, stat_base, hint));       // replicated: identical on every rank
        }
        if (is_split(s) && !p2p) {
            // every rank gets the whole flow of this level (in place: band r sits at rows r*rows_per)
            // -- the peer-memory path has done this on the device (k_band_allgather)
            PairCtl c;
            TRY(band_read_ctl(ctx, &c));
            int r0, r1, rows_per;
            band_rows(w.lv[s].ny, ctx->band_rank, G, &r0, &r1, &rows_per);
            const size_t cnt = (size_t) rows_per * w.lv[s].pitch;
            NK(g_nccl.GroupStart());
            for (int f = F_U1; f <= F_U2; f++) {
                float *plane = w.state + (size_t) c.cur * w.set_stride + (size_t) f * w.field_stride;
                NK(g_nccl.AllGather(plane + (size_t) ctx->band_rank * cnt, plane, cnt, ncclFloat, comm, st));
            }
            NK(g_nccl.GroupEnd());
        }
        if (!s) break;
        const Level &c = w.lv[s], &f = w.lv[s - 1];
        int z0 = 0, z1 = f.ny;
        if (is_split(s - 1)) {       // own rows plus both halo rows, straight from the full coarse flow
            int r0, r1, rows_per;
            band_rows(f.ny, ctx->band_rank, G, &r0, &r1, &rows_per);
            z0 = std::max(r0 - halo, 0);
            z1 = std::min(r1 + halo, f.ny);
        }
        if (z1 > z0) {
            Span zs(ctx, 4);
            dim3 g(ceil_div(f.nx, kZiTW), ceil_div(z1 - z0, kZiTH), 1);
            k_zoom_in_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                                      c, f, (double) f.nx / c.nx, (double) f.ny / c.ny,
                                                      (float) (1.0 / prm.zfactor), z0, z1);
            CKL(ctx);
        }
        k_flip_cur<<<1, 32, 0, st>>>(w.ctl, 1);
        CKL(ctx);
    }
    {
        Span ex(ctx, 5);
        dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
        k_export_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                                 w.lv[0], du1, du2);
        CKL(ctx);
    }
    total.end();
    TRY(fetch_stats(ctx, 1, nstat, iters_out, errs_out));
    if (p2p) {
        int flag = 0;
        CK(cudaMemcpy(&flag, &ctx->my_box->timed_out, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) { ctx->err = "row-band exchange timed out waiting for a peer rank"; return TVL1_ERR_CUDA; }
    }
    return TVL1_OK;
}

// ---- RAII device scratch for the per-kernel hooks -----------------------------------------------
struct Dev {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Dev(cudaStream_t s) : st(s) {}
    ~Dev() { cudaStreamSynchronize(st); for (void *p : ptrs) cudaFree(p); }
    float *alloc(size_t floats)
    {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(floats, 4) * sizeof(float)) != cudaSuccess) return nullptr;
        // zero on the context's own (non-blocking) stream: a legacy-stream memset would not be ordered
        // against the copies and kernels that follow
        cudaMemsetAsync(p, 0, std::max<size_t>(floats, 4) * sizeof(float), st);
        ptrs.push_back(p);
        return (float *) p;
    }
};

} // namespace

#include "hs_solver.cuh"

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

static int band_init_impl(tvl1_ctx *ctx, int rank, int world, const unsigned char *id_bytes);
static int band_solve_host_impl(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2, int nx,
                                int ny, const tvl1_params *prm, int min_split_rows, int *iters_out, double *errs_out);
static int band_solve_dev_impl(tvl1_ctx *ctx, const float *dI0, const float *dI1, float *du1, float *du2,
                               int nx, int ny, const tvl1_params *prm, int min_split_rows, int *iters_out,
                               double *errs_out);

int tvl1_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int tvl1_create(int device, tvl1_ctx **out)
{
    if (!out) return TVL1_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (this library has no CPU fallback)";
        return TVL1_ERR_NODEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return TVL1_ERR_ARG; }
    tvl1_ctx *ctx = new tvl1_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->body_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaMallocHost(&ctx->h_loop, sizeof(LoopCtl))) != cudaSuccess) {
        g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        if (ctx->h_loop) cudaFreeHost(ctx->h_loop);
        if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
        if (ctx->body_stream) cudaStreamDestroy(ctx->body_stream);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TVL1_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char *ng = std::getenv("TVL1_NO_GRAPH")) ctx->use_graph = !(ng[0] == '1');
    if (const char *nr = std::getenv("TVL1_NO_RESIDENT")) ctx->use_resident = !(nr[0] == '1');
    if (const char *nt = std::getenv("TVL1_NO_TB")) ctx->use_tb = !(nt[0] == '1');
    if (const char *ml = std::getenv("TVL1_T2_LEVELS")) ctx->t2_levels = (unsigned int) std::strtoul(ml, nullptr, 0);
    if (const char *ts2 = std::getenv("TVL1_T2_STAGE")) ctx->t2_stage = !(ts2[0] == '0');
    if (const char *ts3 = std::getenv("TVL1_T2_SHARED")) ctx->t2_when_shared = !(ts3[0] == '0');
    if (const char *ts4 = std::getenv("TVL1_T2_STAGE_SHARED")) ctx->t2_stage_when_shared = !(ts4[0] == '0');
    if (const char *tf = std::getenv("TVL1_T2_FIRST")) ctx->t2_first = !(tf[0] == '0');
    if (const char *ta = std::getenv("TVL1_T2_ADAPT")) ctx->t2_adapt = !(ta[0] == '0');
    if (const char *t2 = std::getenv("TVL1_T2")) ctx->use_t2 = std::max(0, std::min(2, std::atoi(t2)));
    if (const char *zp = std::getenv("TVL1_ZERO_PASS")) ctx->zero_in_first = !(zp[0] == '1');
    if (const char *gs = std::getenv("TVL1_GAUSS_SHFL")) ctx->gauss_shfl = !(gs[0] == '0');
    if (const char *wt = std::getenv("TVL1_WARP_TMA")) ctx->warp_tma = !(wt[0] == '0');
    if (const char *sc = std::getenv("TVL1_SLOT_CTAS")) ctx->slot_ctas = std::max(1, std::atoi(sc));
    if (const char *tp = std::getenv("TVL1_TAIL_PAIRS")) ctx->tail_pairs = std::max(0, std::atoi(tp));
    if (const char *tp = std::getenv("TVL1_TAIL_PAIRS_SHARED")) ctx->tail_pairs_shared = std::max(0, std::atoi(tp));
    if (const char *tc = std::getenv("TVL1_TAIL_SLOT_CTAS")) ctx->tail_slot_ctas = std::max(1, std::atoi(tc));
    if (const char *tt = std::getenv("TVL1_TAIL_TB")) ctx->tail_tb = tt[0] == '1';
    if (const char *ts = std::getenv("TVL1_TB_SHARED")) ctx->tb_when_shared = ts[0] == '1';
    if (const char *sd = std::getenv("TVL1_SHORT_DIV")) ctx->short_div = std::max(0, std::atoi(sd));
    if (const char *hp = std::getenv("TVL1_HOST_PIPE")) ctx->host_pipe = hp[0] == '1';
    if (const char *mc = std::getenv("TVL1_MIN_CHUNK")) ctx->ramp_min_chunk = std::max(1, std::atoi(mc));
    if (const char *pl = std::getenv("TVL1_PIPE_LANES")) ctx->pipe_lanes = std::max(1, std::min((int) tvl1_ctx::kMaxLanes, std::atoi(pl)));
    if (const char *cs = std::getenv("TVL1_CHUNKS"))
        for (const char *q = cs; *q;) {
            char *end = nullptr;
            const long v = std::strtol(q, &end, 10);
            if (end == q) break;
            ctx->chunk_override.push_back((int) std::max(1l, v));
            q = *end == ',' ? end + 1 : end;
        }
    if (const char *mp = std::getenv("TVL1_TB_MAX_MPIX")) ctx->tb_max_pixels = std::max(0ll, std::atoll(mp)) << 20;
    *out = ctx;
    return TVL1_OK;
}

void tvl1_destroy(tvl1_ctx *ctx)
{
    if (!ctx) return;
    for (auto &sb : ctx->sib) { if (sb) tvl1_destroy(sb); sb = nullptr; }
    if (ctx->band_ctx) { tvl1_destroy(ctx->band_ctx); ctx->band_ctx = nullptr; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    p2p_release_state(ctx);
    for (auto &g : ctx->level_sg) free_graph(g, ctx->ev_pool);
    for (int r = 0; r < kMaxRanks; r++)
        if (ctx->boxes[r] && ctx->boxes[r] != ctx->my_box) cudaIpcCloseMemHandle(ctx->boxes[r]);
    cudaFree(ctx->my_box);
    cudaFree(ctx->d_handles);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t) ctx->nccl_comm);
    cudaFree(ctx->d_band_sum);
    cudaFree(ctx->d_agree);
    cudaFree(ctx->d_gather_ticket);
    free_graph(ctx->sg, ctx->ev_pool);
    free_workspace(ctx->ws);
    for (int i = 0; i < tvl1_ctx::kAltWs; i++) {
        free_graph(ctx->sg_alt[i], ctx->ev_pool);
        free_workspace(ctx->ws_alt[i]);
    }
    for (int i = 0; i < 2; i++) { cudaFree(ctx->stage_in[i]); cudaFree(ctx->stage_out[i]); }
    for (int i = 0; i < 4; i++) cudaFree(ctx->stage_f32[i]);
    for (auto &sl : ctx->slots) {
        for (int i = 0; i < 2; i++) { cudaFree(sl.in[i]); cudaFree(sl.out[i]); }
        if (sl.up) cudaEventDestroy(sl.up);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.down) cudaEventDestroy(sl.down);
    }
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
    resolve_events(ctx);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (ctx->pipe_buf[i]) cudaFreeHost(ctx->pipe_buf[i]);
        if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    }
    if (ctx->h_loop) cudaFreeHost(ctx->h_loop);
    if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
    if (ctx->body_stream) cudaStreamDestroy(ctx->body_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *tvl1_last_error(const tvl1_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int tvl1_set_profiling(tvl1_ctx *ctx, int on)
{
    if (!ctx) return TVL1_ERR_ARG;
    ctx->profiling = on != 0;
    return TVL1_OK;
}

int tvl1_set_lanes(tvl1_ctx *ctx, int host_lanes, int dev_lanes)
{
    if (!ctx || host_lanes < 1 || dev_lanes < 1 || host_lanes > tvl1_ctx::kMaxLanes || dev_lanes > tvl1_ctx::kMaxLanes)
        return TVL1_ERR_ARG;
    ctx->host_lanes = host_lanes;
    ctx->pipe_lanes = host_lanes;            // an explicit choice also applies to the pinned-buffer pipeline (default there: 3)
    ctx->dev_lanes = dev_lanes;
    return TVL1_OK;
}

int tvl1_set_max_batch(tvl1_ctx *ctx, int pairs)
{
    if (!ctx || pairs < 1) return TVL1_ERR_ARG;
    ctx->max_batch = pairs;
    return TVL1_OK;
}

int tvl1_plan_chunks(int npairs, int max_batch, int *sizes, int cap)
{
    if (npairs < 1 || max_batch < 1) return 0;
    std::vector<std::pair<int, int>> chunks;
    ramp_schedule(npairs, std::min(npairs, max_batch), {}, chunks);
    for (int k = 0; k < (int) chunks.size() && k < cap && sizes; k++) sizes[k] = chunks[k].second;
    return (int) chunks.size();
}

unsigned int tvl1_get_blocked_levels(const tvl1_ctx *ctx)
{
    return ctx && ctx->use_t2 ? (ctx->use_t2 == 2 ? ~0u : ctx->t2_levels) : 0u;
}

void *tvl1_get_stream(const tvl1_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

int tvl1_get_stats(const tvl1_ctx *ctx, tvl1_stats *out)
{
    if (!ctx || !out) return TVL1_ERR_ARG;
    *out = ctx->stats;
    return TVL1_OK;
}

void tvl1_default_params(tvl1_params *p)
{
    if (!p) return;
    p->tau = 0.25; p->lambda = 0.15; p->theta = 0.3; p->nscales = 5; p->zfactor = 0.5;
    p->warps = 5; p->epsilon = 0.01;
}

void tvl1_zoom_size(int nx, int ny, int *nxx, int *nyy, double factor)
{
    // src/zoom.cpp:22-34
    *nxx = (int) (nx * factor + 0.5);
    *nyy = (int) (ny * factor + 0.5);
}

int tvl1_solve_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2, int nx,
                   int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    return solve_host<float>(ctx, 1, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_f64(tvl1_ctx *ctx, const double *I0, const double *I1, double *u1, double *u2, int nx,
                   int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    return solve_host<double>(ctx, 1, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_batch_f32(tvl1_ctx *ctx, int npairs, const float *I0, const float *I1, float *u1,
                         float *u2, int nx, int ny, const tvl1_params *prm, int *iters_out,
                         double *errs_out)
{
    return solve_host<float>(ctx, npairs, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_batch_f64(tvl1_ctx *ctx, int npairs, const double *I0, const double *I1, double *u1,
                         double *u2, int nx, int ny, const tvl1_params *prm, int *iters_out,
                         double *errs_out)
{
    return solve_host<double>(ctx, npairs, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_sequence_f32(tvl1_ctx *ctx, int nframes, const float *frames, float *u1, float *u2,
                            int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    if (ctx && nframes < 2) return fail_arg(ctx, "a frame sequence needs at least 2 frames");
    return solve_host<float>(ctx, nframes - 1, frames, nullptr, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_sequence_u8(tvl1_ctx *ctx, int nframes, const unsigned char *frames, float *u1, float *u2,
                           int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    if (!ctx) return TVL1_ERR_ARG;
    return solve_sequence_u8(ctx, nframes, frames, u1, u2, nx, ny, prm, iters_out, errs_out);
}

int tvl1_solve_sequence_f64(tvl1_ctx *ctx, int nframes, const double *frames, double *u1, double *u2,
                            int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    if (ctx && nframes < 2) return fail_arg(ctx, "a frame sequence needs at least 2 frames");
    return solve_host<double>(ctx, nframes - 1, frames, nullptr, u1, u2, nx, ny, prm, iters_out, errs_out, true);
}

int tvl1_solve_batch_dev_f32(tvl1_ctx *ctx, int npairs, const float *dI0, const float *dI1,
                             float *du1, float *du2, int nx, int ny, const tvl1_params *prm,
                             int *iters_out, double *errs_out)
{
    TRY(check_common(ctx, dI0, dI1, du1, du2, nx, ny, prm, true));
    if (npairs < 1) return fail_arg(ctx, "npairs must be >= 1");
    reset_stats(ctx);
    const size_t n = (size_t) nx * ny;
    const int nstat = prm->nscales * prm->warps;
    const int Bmax = std::min(npairs, ctx->max_batch);
    return run_lanes(ctx, ceil_div(npairs, Bmax), ctx->dev_lanes, [&](tvl1_ctx *c, int k) -> int {
        const int first = k * Bmax, B = std::min(Bmax, npairs - first);
        const size_t off = (size_t) first * n;
        return run_multiscale(c, B, dI0 + off, dI1 + off, du1 + off, du2 + off, nx, ny, *prm,
                              iters_out ? iters_out + (size_t) first * nstat : nullptr,
                              errs_out ? errs_out + (size_t) first * nstat : nullptr);
    });
}

int tvl1_single_scale_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2,
                          int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    return solve_host<float>(ctx, 1, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, false);
}

int tvl1_single_scale_f64(tvl1_ctx *ctx, const double *I0, const double *I1, double *u1, double *u2,
                          int nx, int ny, const tvl1_params *prm, int *iters_out, double *errs_out)
{
    return solve_host<double>(ctx, 1, I0, I1, u1, u2, nx, ny, prm, iters_out, errs_out, false);
}


// ---- row-band mode ------------------------------------------------------------------------------

int tvl1_band_unique_id(unsigned char *id_out)
{
    if (!id_out) return TVL1_ERR_ARG;
    if (const char *e = load_nccl()) { g_create_error = e; return TVL1_ERR_CUDA; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return TVL1_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) <= TVL1_NCCL_ID_BYTES, "id buffer too small");
    memset(id_out, 0, TVL1_NCCL_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return TVL1_OK;
}

// Band solves run on a private sibling context: its workspace is (re)allocated only by collective
// calls, so the ranks' decisions to re-export / re-map buffers can never diverge because one rank did
// some single-GPU work in between.
static tvl1_ctx *band_context(tvl1_ctx *ctx)
{
    if (!ctx) return nullptr;
    if (!ctx->band_ctx) {
        if (tvl1_create(ctx->device, &ctx->band_ctx) != TVL1_OK) {
            ctx->err = std::string("band context: ") + tvl1_last_error(nullptr);
            return nullptr;
        }
        ctx->band_ctx->is_sibling = true;
    }
    ctx->band_ctx->profiling = ctx->profiling;
    ctx->band_ctx->use_resident = ctx->use_resident;
    return ctx->band_ctx;
}

static int band_return(tvl1_ctx *ctx, tvl1_ctx *bc, int rc)
{
    ctx->stats = bc->stats;
    if (rc != TVL1_OK) ctx->err = bc->err;
    return rc;
}

int tvl1_band_init(tvl1_ctx *outer, int rank, int world, const unsigned char *id_bytes)
{
    if (!outer || !id_bytes || world < 1 || rank < 0 || rank >= world) return fail_arg(outer, "bad argument");
    tvl1_ctx *bc = band_context(outer);
    if (!bc) return TVL1_ERR_CUDA;
    return band_return(outer, bc, band_init_impl(bc, rank, world, id_bytes));
}

static int band_init_impl(tvl1_ctx *ctx, int rank, int world, const unsigned char *id_bytes)
{
    if (const char *e = load_nccl()) { ctx->err = e; return TVL1_ERR_CUDA; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) { g_nccl.CommDestroy((ncclComm_t) ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm;
    NK(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->band_rank = rank;
    ctx->band_world = world;
    if (!ctx->d_band_sum) CK(cudaMalloc(&ctx->d_band_sum, sizeof(double)));
    if (const char *e = std::getenv("TVL1_BAND_NCCL")) ctx->band_use_nccl_per_iteration = (e[0] == '1');
    p2p_release_state(ctx);
    TRY(p2p_setup_mailboxes(ctx));
    return TVL1_OK;
}

int tvl1_band_solve_f32(tvl1_ctx *outer, const float *I0, const float *I1, float *u1, float *u2, int nx,
                        int ny, const tvl1_params *prm, int min_split_rows, int *iters_out, double *errs_out)
{
    tvl1_ctx *bc = band_context(outer);
    if (!bc) return outer ? TVL1_ERR_CUDA : TVL1_ERR_ARG;
    return band_return(outer, bc, band_solve_host_impl(bc, I0, I1, u1, u2, nx, ny, prm, min_split_rows, iters_out, errs_out));
}

static int band_solve_host_impl(tvl1_ctx *ctx, const float *I0, const float *I1, float *u1, float *u2, int nx,
                                int ny, const tvl1_params *prm, int min_split_rows, int *iters_out, double *errs_out)
{
    TRY(check_common(ctx, I0, I1, u1, u2, nx, ny, prm, true));
    if (!ctx->nccl_comm) return fail_arg(ctx, "tvl1_band_init has not been called on this context");
    reset_stats(ctx);
    const size_t n = (size_t) nx * ny;
    TRY(ensure_stage(ctx, n * sizeof(float), false));
    cudaStream_t st = ctx->stream;
    const float *in[2] = { I0, I1 };
    for (int k = 0; k < 2; k++) {
        if (use_ring<float>(n) && is_pageable(in[k])) TRY(upload_pageable<float>(ctx, (float *) ctx->stage_in[k], in[k], n));
        else CK(cudaMemcpyAsync(ctx->stage_in[k], in[k], n * 4, cudaMemcpyHostToDevice, st));
    }
    TRY(run_band(ctx, (const float *) ctx->stage_in[0], (const float *) ctx->stage_in[1],
                 (float *) ctx->stage_out[0], (float *) ctx->stage_out[1], nx, ny, *prm, min_split_rows,
                 iters_out, errs_out));
    float *outp[2] = { u1, u2 };
    for (int k = 0; k < 2; k++) {
        if (use_ring<float>(n) && is_pageable(outp[k])) TRY(download_pageable<float>(ctx, outp[k], (const float *) ctx->stage_out[k], n));
        else CK(cudaMemcpyAsync(outp[k], ctx->stage_out[k], n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    resolve_events(ctx);
    return TVL1_OK;
}

int tvl1_band_solve_dev_f32(tvl1_ctx *outer, const float *dI0, const float *dI1, float *du1, float *du2,
                            int nx, int ny, const tvl1_params *prm, int min_split_rows, int *iters_out,
                            double *errs_out)
{
    tvl1_ctx *bc = band_context(outer);
    if (!bc) return outer ? TVL1_ERR_CUDA : TVL1_ERR_ARG;
    return band_return(outer, bc, band_solve_dev_impl(bc, dI0, dI1, du1, du2, nx, ny, prm, min_split_rows, iters_out, errs_out));
}

static int band_solve_dev_impl(tvl1_ctx *ctx, const float *dI0, const float *dI1, float *du1, float *du2,
                               int nx, int ny, const tvl1_params *prm, int min_split_rows, int *iters_out,
                               double *errs_out)
{
    TRY(check_common(ctx, dI0, dI1, du1, du2, nx, ny, prm, true));
    if (!ctx->nccl_comm) return fail_arg(ctx, "tvl1_band_init has not been called on this context");
    reset_stats(ctx);
    TRY(run_band(ctx, dI0, dI1, du1, du2, nx, ny, *prm, min_split_rows, iters_out, errs_out));
    resolve_events(ctx);
    return TVL1_OK;
}

int tvl1_band_set_exchange(tvl1_ctx *outer, int use_nccl_per_iteration)
{
    tvl1_ctx *ctx = band_context(outer);
    if (!ctx) return TVL1_ERR_ARG;
    if (!ctx->nccl_comm) return fail_arg(outer, "tvl1_band_init has not been called on this context");
    // collective: NCCL per iteration is used if any rank asks for it (ranks in different modes would hang)
    CK(cudaSetDevice(ctx->device));
    double want = use_nccl_per_iteration != 0 ? 1.0 : 0.0;
    const int rc = agree_sum(ctx, &want, 1);
    if (rc != TVL1_OK) { outer->err = ctx->err; return rc; }
    ctx->band_use_nccl_per_iteration = want > 0.0;
    return TVL1_OK;
}

int tvl1_band_exchange_mode(const tvl1_ctx *outer)
{
    const tvl1_ctx *ctx = outer ? outer->band_ctx : nullptr;
    if (!ctx || !ctx->nccl_comm) return -1;
    return (ctx->p2p_ready && !ctx->band_use_nccl_per_iteration) ? 1 : 0;
}

void tvl1_band_rows(int ny, int rank, int world, int *row_begin, int *row_end)
{
    int rp;
    band_rows(ny, rank, world, row_begin, row_end, &rp);
}

// ---- hooks --------------------------------------------------------------------------------------

int tvl1_normalize_f32(tvl1_ctx *ctx, const float *I0, const float *I1, float *I0n, float *I1n,
                       int nx, int ny)
{
    if (!ctx || !I0 || !I1 || !I0n || !I1n || nx < 1 || ny < 1) return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    Dev d(ctx->stream);
    const size_t n = (size_t) nx * ny;
    float *a = d.alloc(n), *b = d.alloc(n), *oa = d.alloc(n), *ob = d.alloc(n);
    unsigned int *mm = (unsigned int *) d.alloc(4);
    PairCtl *ctl = (PairCtl *) d.alloc(sizeof(PairCtl));
    if (!a || !b || !oa || !ob || !mm || !ctl) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(a, I0, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b, I1, n * 4, cudaMemcpyHostToDevice, st));
    k_init_ctl<<<1, 32, 0, st>>>(ctl, mm, 1);
    CKL(ctx);
    k_minmax<<<dim3(64, 1), 256, 0, st>>>(a, b, n, mm);
    CKL(ctx);
    GaussTaps id{};
    id.size = 1;
    id.w[0] = 1.0f;   // identity blur: the normalisation is fused into the blur kernel's load
    TRY(launch_gauss(ctx, 1, a, nx, n, oa, nx, n, nx, ny, nx, ny, id, mm, 1, 1));
    TRY(launch_gauss(ctx, 1, b, nx, n, ob, nx, n, nx, ny, nx, ny, id, mm, 1, 1));
    CK(cudaMemcpyAsync(I0n, oa, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(I1n, ob, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_gaussian_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, double sigma)
{
    if (!ctx || !I || !out || nx < 1 || ny < 1) return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    GaussTaps taps;
    TRY(check_sigma(ctx, sigma, nx, taps));
    Dev d(ctx->stream);
    const size_t n = (size_t) nx * ny;
    float *a = d.alloc(n), *o = d.alloc(n);
    if (!a || !o) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(a, I, n * 4, cudaMemcpyHostToDevice, st));
    TRY(launch_gauss(ctx, 1, a, nx, n, o, nx, n, nx, ny, nx, ny, taps, nullptr, 1, 1));
    CK(cudaMemcpyAsync(out, o, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_zoom_out_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, double factor)
{
    if (!ctx || !I || !out || nx < 1 || ny < 1 || !(factor > 0.0 && factor < 1.0))
        return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    int nxx, nyy;
    tvl1_zoom_size(nx, ny, &nxx, &nyy, factor);
    if (nxx < 1 || nyy < 1) return fail_arg(ctx, "zoomed image is empty");
    GaussTaps taps;
    TRY(check_sigma(ctx, TVL1_ZOOM_SIGMA_ZERO * std::sqrt(1.0 / (factor * factor) - 1.0), nx, taps));
    Dev d(ctx->stream);
    const size_t n = (size_t) nx * ny, m = (size_t) nxx * nyy;
    float *a = d.alloc(n), *t = d.alloc(n), *o = d.alloc(m);
    if (!a || !t || !o) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(a, I, n * 4, cudaMemcpyHostToDevice, st));
    if (factor == 0.5) {
        TRY(launch_gauss(ctx, 2, a, nx, n, o, nxx, m, nx, ny, nxx, nyy, taps, nullptr, 1, 1));
    } else {
        TRY(launch_gauss(ctx, 1, a, nx, n, t, nx, n, nx, ny, nx, ny, taps, nullptr, 1, 1));
        dim3 g(ceil_div(nxx, 32), ceil_div(nyy, 8), 1);
        k_resample<<<g, dim3(32, 8), 0, st>>>(t, nx, n, nx, ny, o, nxx, m, nxx, nyy, factor, factor, 1.0f);
        CKL(ctx);
    }
    CK(cudaMemcpyAsync(out, o, m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_zoom_in_f32(tvl1_ctx *ctx, const float *I, float *out, int nx, int ny, int nxx, int nyy,
                     double scale)
{
    if (!ctx || !I || !out || nx < 1 || ny < 1 || nxx < 1 || nyy < 1) return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    Dev d(ctx->stream);
    const size_t n = (size_t) nx * ny, m = (size_t) nxx * nyy;
    float *a = d.alloc(n), *o = d.alloc(m);
    if (!a || !o) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(a, I, n * 4, cudaMemcpyHostToDevice, st));
    dim3 g(ceil_div(nxx, 32), ceil_div(nyy, 8), 1);
    k_resample<<<g, dim3(32, 8), 0, st>>>(a, nx, n, nx, ny, o, nxx, m, nxx, nyy, (double) nxx / nx,
                                          (double) nyy / ny, (float) scale);
    CKL(ctx);
    CK(cudaMemcpyAsync(out, o, m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_warp_f32(tvl1_ctx *ctx, const float *I0, const float *I1, const float *u1, const float *u2,
                  int nx, int ny, float *I1wx, float *I1wy, float *rho_c, float *grad)
{
    if (!ctx || !I0 || !I1 || !u1 || !u2 || !I1wx || !I1wy || !rho_c || !grad || nx < 1 || ny < 1)
        return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, 1, 1));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t) nx * ny;
    Dev d(ctx->stream);
    float *buf = d.alloc(4 * n);
    if (!buf) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    const float *src[4] = { I0, I1, u1, u2 };
    for (int k = 0; k < 4; k++) CK(cudaMemcpyAsync(buf + k * n, src[k], n * 4, cudaMemcpyHostToDevice, st));
    k_init_ctl<<<1, 32, 0, st>>>(w.ctl, w.mm, 1);
    CKL(ctx);
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
    k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.I0(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_pack<<<g, dim3(32, 8), 0, st>>>(buf + n, w.I1(0), nx, ny, w.lv[0].pitch, w.plane(0));
    CKL(ctx);
    k_import_flow<<<g, dim3(32, 8), 0, st>>>(w.state, w.plane0, w.field_stride, w.set_stride, w.ctl,
                                             w.lv[0], buf + 2 * n, buf + 3 * n);
    CKL(ctx);
    TRY(launch_warp(ctx, 0, 1, 1));
    float *outs[4] = { I1wx, I1wy, rho_c, grad };
    const int which[4] = { C_IX, C_IY, C_RHO, C_GRAD };
    for (int k = 0; k < 4; k++) {
        k_unpack<<<g, dim3(32, 8), 0, st>>>(w.consts + (size_t) which[k] * w.field_stride, buf + k * n,
                                            nx, ny, w.lv[0].pitch, w.plane0);
        CKL(ctx);
        CK(cudaMemcpyAsync(outs[k], buf + k * n, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_iterate_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                     float *p22, const float *rho_c, const float *I1wx, const float *I1wy,
                     const float *grad, int nx, int ny, double tau, double lambda, double theta,
                     int iters, double *errs_out)
{
    if (!ctx || !u1 || !u2 || !p11 || !p12 || !p21 || !p22 || !rho_c || !I1wx || !I1wy || !grad ||
        nx < 1 || ny < 1 || iters < 0)
        return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, 1, 1));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t) nx * ny;
    Dev d(ctx->stream);
    float *buf = d.alloc(n);
    if (!buf) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    k_init_ctl<<<1, 32, 0, st>>>(w.ctl, w.mm, 1);
    CKL(ctx);
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
    float *st_host[6] = { u1, u2, p11, p12, p21, p22 };
    for (int f = 0; f < 6; f++) {
        CK(cudaMemcpyAsync(buf, st_host[f], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.state + (size_t) f * w.field_stride, nx, ny,
                                          w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    const float *c_host[4] = { I1wx, I1wy, rho_c, grad };   // order of enum Const
    for (int c = 0; c < 4; c++) {
        CK(cudaMemcpyAsync(buf, c_host[c], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.consts + (size_t) c * w.field_stride, nx, ny,
                                          w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    tvl1_params prm;
    tvl1_default_params(&prm);
    prm.tau = tau; prm.lambda = lambda; prm.theta = theta; prm.epsilon = 0.0;
    int cur = 0;
    for (int k = 0; k < iters; k++) {
        k_begin_warp<<<1, 32, 0, st>>>(w.ctl, w.loop, 1);
        CKL(ctx);
        const IterParams P = iter_params(ctx, w.lv[0], prm, 0, 1);   // max_iter 1: record and stop
        TRY(launch_iterate(ctx, P, 1));
        cur ^= 1;
        if (errs_out) CK(cudaMemcpyAsync(errs_out + k, w.stat_errs, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    for (int f = 0; f < 6; f++) {
        k_unpack<<<g, dim3(32, 8), 0, st>>>(w.state + (size_t) cur * w.set_stride + (size_t) f * w.field_stride,
                                            buf, nx, ny, w.lv[0].pitch, w.plane0);
        CKL(ctx);
        CK(cudaMemcpyAsync(st_host[f], buf, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_iterate_resident_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                              float *p22, const float *rho_c, const float *I1wx, const float *I1wy,
                              int nx, int ny, double tau, double lambda, double theta, double epsilon,
                              int max_iter, int cluster, int *iters_out, double *errs_out, int *cluster_out)
{
    if (!ctx || !u1 || !u2 || !p11 || !p12 || !p21 || !p22 || !rho_c || !I1wx || !I1wy || nx < 1 ||
        ny < 1 || max_iter < 1 || max_iter > 100000)
        return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    const int saved_force = ctx->force_cluster;
    const bool saved_res = ctx->use_resident;
    ctx->force_cluster = cluster;
    ctx->use_resident = true;
    const int rc0 = ensure_workspace(ctx, nx, ny, 1, 0.5, 1, 1);
    ctx->force_cluster = saved_force;
    ctx->use_resident = saved_res;
    TRY(rc0);
    Workspace &w = ctx->ws;
    if (cluster_out) *cluster_out = w.res_cluster[0];
    if (w.res_cluster[0] == 0) return fail_arg(ctx, "level does not fit the cluster-resident kernel with this cluster size");
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t) nx * ny;
    Dev d(ctx->stream);
    float *buf = d.alloc(n);
    double *trace = (double *) d.alloc(2 * (size_t) max_iter);
    if (!buf || !trace) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    k_init_ctl<<<1, 32, 0, st>>>(w.ctl, w.mm, 1);
    CKL(ctx);
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
    float *st_host[6] = { u1, u2, p11, p12, p21, p22 };
    for (int f = 0; f < 6; f++) {
        CK(cudaMemcpyAsync(buf, st_host[f], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.state + (size_t) f * w.field_stride, nx, ny,
                                          w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    const float *c_host[3] = { I1wx, I1wy, rho_c };   // order of enum Const; grad is recomputed on chip
    for (int c = 0; c < 3; c++) {
        CK(cudaMemcpyAsync(buf, c_host[c], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.consts + (size_t) c * w.field_stride, nx, ny,
                                          w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    tvl1_params prm;
    tvl1_default_params(&prm);
    prm.tau = tau; prm.lambda = lambda; prm.theta = theta;
    const double eps2 = epsilon < 0 ? -1.0 : epsilon * epsilon;      // negative: never stop early
    TRY(launch_resident(ctx, 0, 1, prm, 0, max_iter, eps2, trace));
    int it = 0;
    CK(cudaMemcpyAsync(&it, w.stat_iters, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (iters_out) *iters_out = it;
    if (errs_out) CK(cudaMemcpyAsync(errs_out, trace, sizeof(double) * it, cudaMemcpyDeviceToHost, st));
    for (int f = 0; f < 6; f++) {      // the kernel wrote set 1 (cur was 0)
        k_unpack<<<g, dim3(32, 8), 0, st>>>(w.state + w.set_stride + (size_t) f * w.field_stride, buf, nx, ny,
                                            w.lv[0].pitch, w.plane0);
        CKL(ctx);
        CK(cudaMemcpyAsync(st_host[f], buf, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

int tvl1_iterate_loop_f32(tvl1_ctx *ctx, float *u1, float *u2, float *p11, float *p12, float *p21,
                          float *p22, const float *rho_c, const float *I1wx, const float *I1wy, int nx,
                          int ny, double tau, double lambda, double theta, double epsilon, int max_iter,
                          int temporal_blocking, int *iters_out, double *err_out, int *launches_out)
{
    if (!ctx || !u1 || !u2 || !p11 || !p12 || !p21 || !p22 || !rho_c || !I1wx || !I1wy || nx < 1 ||
        ny < 1 || max_iter < 1)
        return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    const bool saved_res = ctx->use_resident;
    ctx->use_resident = false;                               // this hook is about the streaming kernels
    const int rc0 = ensure_workspace(ctx, nx, ny, 1, 0.5, 1, 1);
    ctx->use_resident = saved_res;
    TRY(rc0);
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t) nx * ny;
    Dev d(ctx->stream);
    float *buf = d.alloc(n);
    if (!buf) { ctx->err = "cudaMalloc failed"; return TVL1_ERR_CUDA; }
    k_init_ctl<<<1, 32, 0, st>>>(w.ctl, w.mm, 1);
    CKL(ctx);
    CK(cudaMemsetAsync(w.counters, 0, sizeof(unsigned long long) * kCounterWords, st));
    dim3 g(ceil_div(nx, 32), ceil_div(ny, 8), 1);
    float *st_host[6] = { u1, u2, p11, p12, p21, p22 };
    for (int f = 0; f < 6; f++) {
        CK(cudaMemcpyAsync(buf, st_host[f], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.state + (size_t) f * w.field_stride, nx, ny, w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    const float *c_host[3] = { I1wx, I1wy, rho_c };
    for (int c = 0; c < 3; c++) {
        CK(cudaMemcpyAsync(buf, c_host[c], n * 4, cudaMemcpyHostToDevice, st));
        k_pack<<<g, dim3(32, 8), 0, st>>>(buf, w.consts + (size_t) c * w.field_stride, nx, ny, w.lv[0].pitch, w.plane0);
        CKL(ctx);
    }
    tvl1_params prm;
    tvl1_default_params(&prm);
    prm.tau = tau; prm.lambda = lambda; prm.theta = theta; prm.epsilon = epsilon < 0 ? 0.0 : epsilon;
    k_begin_warp<<<1, 32, 0, st>>>(w.ctl, w.loop, 1);
    CKL(ctx);
    IterParams P = iter_params(ctx, w.lv[0], prm, 0, max_iter);
    if (epsilon < 0) P.eps2 = -1.0;                          // never stop before max_iter
    if (temporal_blocking) {
        const bool t2 = temporal_blocking >= 3;               // 3 / 4: the two-iteration marching kernel
        if (!t2 && !tb_usable(ctx, w.lv[0], 1)) return fail_arg(ctx, "temporally blocked kernel not usable for this size");
        set_blocking(P, t2 ? 2 : 1);
        if (temporal_blocking == 2 || temporal_blocking == 4) {   // force full blocks from the first launch on
            PairCtl c;
            CK(cudaMemcpyAsync(&c, w.ctl, sizeof c, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            c.nsteps = std::min(P.tb_max, max_iter);
            CK(cudaMemcpyAsync(w.ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
        }
    }
    int launches = 0;
    for (;;) {
        TRY(launch_iterate(ctx, P, 1));
        launches++;
        CK(cudaMemcpyAsync(ctx->h_loop, w.loop, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (ctx->h_loop->active_pairs == 0 || launches > 4 * max_iter + 8) break;
    }
    PairCtl c;
    CK(cudaMemcpyAsync(&c, w.ctl, sizeof c, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (iters_out) *iters_out = c.n;
    if (err_out) *err_out = c.err;
    if (launches_out) *launches_out = launches;
    for (int f = 0; f < 6; f++) {
        k_unpack<<<g, dim3(32, 8), 0, st>>>(w.state + (size_t) c.cur * w.set_stride + (size_t) f * w.field_stride,
                                            buf, nx, ny, w.lv[0].pitch, w.plane0);
        CKL(ctx);
        CK(cudaMemcpyAsync(st_host[f], buf, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return c.active ? fail_arg(ctx, "iteration loop did not terminate") : TVL1_OK;
}

int tvl1_bench_iterate(tvl1_ctx *ctx, int npairs, int nx, int ny, int launches, double *ms_out)
{
    if (!ctx || npairs < 1 || nx < 1 || ny < 1 || launches < 1 || !ms_out) return fail_arg(ctx, "bad argument");
    CK(cudaSetDevice(ctx->device));
    TRY(ensure_workspace(ctx, nx, ny, 1, 0.5, npairs, 1));
    Workspace &w = ctx->ws;
    cudaStream_t st = ctx->stream;
    k_fill_bench<<<ctx->sm_count * 8, 256, 0, st>>>(w.state, w.consts, w.field_stride, w.set_stride);
    CKL(ctx);
    k_init_ctl<<<ceil_div(npairs, 128), 128, 0, st>>>(w.ctl, w.mm, npairs);
    CKL(ctx);
    k_begin_warp<<<ceil_div(npairs, 128), 128, 0, st>>>(w.ctl, w.loop, npairs);
    CKL(ctx);
    tvl1_params prm;
    tvl1_default_params(&prm);
    prm.epsilon = 0.0;
    IterParams P = iter_params(ctx, w.lv[0], prm, 0, 1 << 30);
    P.eps2 = -1.0;   // never stop: every launch does the full work
    // TVL1_BENCH_TB=1 / 2: blocks through k_iterate_tb / k_iterate_t2 (the device switches to full blocks after two
    // single iterations: a launch then advances every pair by 4 / 2 iterations)
    if (const char *bt = std::getenv("TVL1_BENCH_TB")) {
        const int mode = std::atoi(bt);
        if (mode == 1 && tb_usable(ctx, w.lv[0], 1)) set_blocking(P, 1);
        if (mode == 2) set_blocking(P, 2);
    }
    for (int i = 0; i < 3; i++) TRY(launch_iterate(ctx, P, npairs));
    cudaEvent_t a = take_event(ctx), b = take_event(ctx);
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < launches; i++) TRY(launch_iterate(ctx, P, npairs));
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    ctx->ev_pool.push_back(a);
    ctx->ev_pool.push_back(b);
    *ms_out = ms;
    return TVL1_OK;
}

} // extern "C"
