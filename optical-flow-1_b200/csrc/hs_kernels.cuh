// Device kernels of the pyramidal Horn-Schunck path (src/horn_schunck_pyramidal.cpp).
// The pyramid, the warp (k_warp gives I2wx, I2wy and rho_c = -(I1 - I2w + I2wx u + I2wy v), which is
// all the linear system of :127-137 is made of) and the flow up-sampling are the TV-L1 kernels of
// tvl1_kernels.cuh; what is new is the SOR sweep.
//
//   k_hs_to_wave    row-major planes (u, v, I2wx, I2wy, rho_c) -> wave layout (hs_sor_step.h)
//   k_hs_sor<P>     the whole `while (error > TOL && niter < maxiter)` loop of one warp step,
//                   one CTA per frame pair, lexicographic Gauss-Seidel order kept exactly by a
//                   wavefront schedule; no host round trip, no second launch
//   k_hs_from_wave  wave layout (u, v) -> row-major planes
#pragma once
#include "hs_sor_step.h"
#include "hs_sor_pipe.h"
#include "hs_sor_pairs.h"
#include "tvl1_kernels.cuh"

#include <type_traits>

namespace tvl1 {

constexpr int kHsMaxThreads = 1024;
constexpr size_t kHsSmemLimit = 227 * 1024;

// Bytes of shared memory the rings of k_hs_sor<P> take for `rp` padded rows.
__host__ __device__ inline size_t hs_ring_bytes(int P, int rp)
{
    return (size_t) (2 * (hs::kRingBase + P) + 3 * (P + 2)) * rp * sizeof(float);
}

struct HsSorParams {
    float *state;                 // state[set][field][b][plane0]; the wave planes live in the idle set
    size_t plane0, set_stride;
    const PairCtl *ctl;
    int nx, ny, rp;
    float alpha2;
    double tol;
    int max_iter;
    int *stat_iters;
    double *stat_errs;
    int stat_stride, stat_slot;
    unsigned long long *px_iters; // [level] pixel-iterations
    int level;
    // pipelined kernel only (k_hs_sor_pipe): wave period, snapshot period, error depth, snapshot planes
    // (2 x 2 L ny floats per pair) and per-row error sums (D x rp doubles per pair)
    int L, K, D;
    float *snap;
    size_t snap_stride;
    double *part;
    size_t part_stride;
};

// Where the wave planes of pair b live: the idle ping-pong set of the state buffer (6 B plane0 floats),
// re-partitioned per pair as [b][6 * plane0]:  (u, v) interleaved [2n] | (I2wx, I2wy) interleaved [2n] |
// rho_c [n],  n = nx * ny <= plane0.
__device__ __forceinline__ float *hs_wave_base(float *state, size_t set_stride, size_t plane0, int cur, int b)
{
    return state + (size_t) (cur ^ 1) * set_stride + (size_t) b * 6 * plane0;
}

// Tile transposes between the row-major pitched planes and the wave layout
//   W[((j + 2i) mod M) * ny + i] = plane[i * pitch + j],   M = nx (k_hs_sor) or L (k_hs_sor_pipe).
// A CTA moves a 32 x 32 tile of (wave column c, row i): row-major side coalesced along j (= c - 2i,
// consecutive in c), wave side coalesced along i.  snap (pipelined kernel): the initial flow is also
// the snapshot of "sweep 0".
__global__ void __launch_bounds__(256)
k_hs_to_wave(float *__restrict__ state, const float *__restrict__ consts, size_t plane0, size_t field_stride,
             size_t set_stride, const PairCtl *__restrict__ ctl, Level lv, int mod, float *__restrict__ snap,
             size_t snap_stride)
{
    __shared__ float tile[5][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int b = blockIdx.z;
    const int cur = ctl[b].cur;
    const int nx = lv.nx, ny = lv.ny, pitch = lv.pitch;
    const int c0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const float *src[5];
    src[0] = state + (size_t) cur * set_stride + (size_t) F_U1 * field_stride + (size_t) b * plane0;
    src[1] = state + (size_t) cur * set_stride + (size_t) F_U2 * field_stride + (size_t) b * plane0;
    src[2] = consts + (size_t) C_IX * field_stride + (size_t) b * plane0;
    src[3] = consts + (size_t) C_IY * field_stride + (size_t) b * plane0;
    src[4] = consts + (size_t) C_RHO * field_stride + (size_t) b * plane0;
    const size_t n = (size_t) mod * ny;
    float *base = hs_wave_base(state, set_stride, plane0, cur, b);
    float2 *wuv = reinterpret_cast<float2 *>(base);
    float2 *wxy = reinterpret_cast<float2 *>(base + 2 * n);
    float *wrho = base + 4 * n;
    float2 *snap0 = snap ? reinterpret_cast<float2 *>(snap + (size_t) b * snap_stride) : nullptr;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = i0 + ty + 8 * r, c = c0 + tx;
        if (i < ny && c < mod) {
            const int j = hs::pmod(c - 2 * i, mod);
            if (j < nx) {
                const int o = i * pitch + j;
#pragma unroll
                for (int k = 0; k < 5; k++) tile[k][ty + 8 * r][tx] = src[k][o];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int c = c0 + ty + 8 * r, i = i0 + tx;
        if (i < ny && c < mod && hs::pmod(c - 2 * i, mod) < nx) {
            const size_t o = (size_t) c * ny + i;
            const float2 uv = make_float2(tile[0][tx][ty + 8 * r], tile[1][tx][ty + 8 * r]);
            wuv[o] = uv;
            if (snap0) snap0[o] = uv;
            wxy[o] = make_float2(tile[2][tx][ty + 8 * r], tile[3][tx][ty + 8 * r]);
            wrho[o] = tile[4][tx][ty + 8 * r];
        }
    }
}

__global__ void __launch_bounds__(256)
k_hs_from_wave(float *__restrict__ state, size_t plane0, size_t field_stride, size_t set_stride,
               const PairCtl *__restrict__ ctl, Level lv, int mod)
{
    __shared__ float tile[2][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int b = blockIdx.z;
    const int cur = ctl[b].cur;
    const int nx = lv.nx, ny = lv.ny, pitch = lv.pitch;
    const int c0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const float2 *wuv = reinterpret_cast<const float2 *>(hs_wave_base(state, set_stride, plane0, cur, b));
    float *dst = state + (size_t) cur * set_stride + (size_t) b * plane0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int c = c0 + ty + 8 * r, i = i0 + tx;
        if (i < ny && c < mod && hs::pmod(c - 2 * i, mod) < nx) {
            const float2 uv = wuv[(size_t) c * ny + i];
            tile[0][ty + 8 * r][tx] = uv.x;
            tile[1][ty + 8 * r][tx] = uv.y;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = i0 + ty + 8 * r, c = c0 + tx;
        if (i < ny && c < mod) {
            const int j = hs::pmod(c - 2 * i, mod);
            if (j < nx) {
                const int o = i * pitch + j;
                dst[(size_t) F_U1 * field_stride + o] = tile[0][tx][ty + 8 * r];
                dst[(size_t) F_U2 * field_stride + o] = tile[1][tx][ty + 8 * r];
            }
        }
    }
}

__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned int) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

struct HsCpAsync {
    __device__ __forceinline__ void cp4(float *dst, const float *src) { cp_async4(dst, src); }
    __device__ __forceinline__ void cp8(float2 *dst, const float2 *src) { cp_async8(dst, src); }
};
// Rings in global memory (images with more rows than one SM's shared memory can ring-buffer): plain
// copies, ordered by the per-step barrier like everything else the threads of a CTA exchange.
struct HsCpSync {
    __device__ __forceinline__ void cp4(float *dst, const float *src) { *dst = *src; }
    __device__ __forceinline__ void cp8(float2 *dst, const float2 *src) { *dst = *src; }
};

// Floats of ring storage a pair needs when the rings live in global memory (P = 0), and where they
// start inside the pair's wave region (after the 5n floats of the wave planes, 16-byte aligned).
__host__ __device__ inline size_t hs_global_ring_floats(int rp) { return (size_t) (2 * hs::kRingBase + 3 * 2) * rp; }
__host__ __device__ inline size_t hs_global_ring_offset(size_t n) { return (5 * n + 3) / 4 * 4; }

template <int N>
__device__ __forceinline__ void hs_cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// One CTA per frame pair runs the complete SOR loop of a warp step (src/horn_schunck_pyramidal.cpp:
// 139-231).  Thread tid owns rows tid, tid + blockDim.x, ...; a sweep is the time-step loop of
// hs_sor_step.h with one __syncthreads per step; wave columns of u, v and of the coefficients are
// fetched P steps ahead with cp.async into shared-memory rings.  The squared-update sum is reduced
// in fp64 in a fixed order, so the stopping decision does not depend on scheduling.
// GLOBAL_RING (with P = 0): the rings live in the pair's wave region in global memory instead -- the
// same schedule for images whose rows do not fit the shared-memory rings (ny > HS_MAX_ROWS).
template <int P, bool GLOBAL_RING>
__global__ void __launch_bounds__(kHsMaxThreads)
k_hs_sor(HsSorParams A)
{
    extern __shared__ __align__(16) float hs_smem[];
    __shared__ double s_red[kHsMaxThreads / 32];
    __shared__ double s_err;
    __shared__ int s_go;

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int b = blockIdx.x;
    const int nx = A.nx, ny = A.ny;
    float *wave = hs_wave_base(A.state, A.set_stride, A.plane0, A.ctl[b].cur, b);
    const size_t n = (size_t) nx * ny;

    hs::SorView V;
    V.wuv = reinterpret_cast<float2 *>(wave);
    V.wxy = reinterpret_cast<const float2 *>(wave + 2 * n);
    V.wrho = wave + 4 * n;
    V.nx = nx; V.ny = ny; V.alpha2 = A.alpha2;
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = A.rp;
    V.ring_uv = reinterpret_cast<float2 *>(GLOBAL_RING ? wave + hs_global_ring_offset(n) : hs_smem);
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = reinterpret_cast<float *>(V.cxy + (size_t) V.CD * V.rp);

    typename std::conditional<GLOBAL_RING, HsCpSync, HsCpAsync>::type cp;
    const int t_first = hs::first_step(V), t_last = hs::last_step(V);
    int niter = 0;
    while (true) {
        niter++;
        double e = 0.0;
        hs::Step s = hs::make_step(V, t_first);
        for (int t = t_first; t <= t_last; t++, hs::advance(V, s)) {
            hs_cp_async_wait<P>();
            __syncthreads();
            for (int i = tid; i < ny; i += nthreads) hs::issue_row(V, s, i, cp);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (t >= 3)
                for (int i = tid; i < ny; i += nthreads) e += hs::compute_row(V, s, i);
        }
        hs_cp_async_wait<0>();
        __syncthreads();
        if (tid == 0) e += hs::corners(V);
        e = warp_sum(e);
        if ((tid & 31) == 0) s_red[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < (nthreads + 31) / 32; w++) sum += s_red[w];
            const double error = sqrt(sum / (double) (nx * ny));        // :230
            s_err = error;
            s_go = (error > A.tol && niter < A.max_iter) ? 1 : 0;       // :143
        }
        __syncthreads();
        if (!s_go) break;
    }
    if (tid == 0) {
        A.stat_iters[(size_t) b * A.stat_stride + A.stat_slot] = niter;
        A.stat_errs[(size_t) b * A.stat_stride + A.stat_slot] = s_err;
        atomicAdd(A.px_iters + A.level, (unsigned long long) niter * (unsigned long long) (nx * ny));
    }
}

// The pipelined form of k_hs_sor (hs_sor_pipe.h): sweep n of row i at time n*L + 2i + j, so that all rows
// work all the time (a row of k_hs_sor works nx of 2*ny + nx steps).  Speculative phase with per-row error
// sums (decided one barrier after a sweep completes, by warp 0 in a fixed order) and snapshots of every
// K-th sweep; on a stop by TOL the snapshot is restored and the remaining sweeps are replayed with a
// known count.  Dynamic shared memory: the rings of k_hs_sor<P> followed by rp doubles (running error
// sum of each row's current sweep).
template <int P>
__global__ void __launch_bounds__(kHsMaxThreads)
k_hs_sor_pipe(HsSorParams A)
{
    extern __shared__ __align__(16) float hs_smem[];
    __shared__ double s_err;
    __shared__ int s_go;

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int b = blockIdx.x;
    const int nx = A.nx, ny = A.ny, L = A.L;
    float *wave = hs_wave_base(A.state, A.set_stride, A.plane0, A.ctl[b].cur, b);
    const size_t n = (size_t) L * ny;

    hs::PipeView V;
    V.wuv = reinterpret_cast<float2 *>(wave);
    V.wxy = reinterpret_cast<const float2 *>(wave + 2 * n);
    V.wrho = wave + 4 * n;
    V.snap0 = reinterpret_cast<float2 *>(A.snap + (size_t) b * A.snap_stride);
    V.snap1 = V.snap0 + n;
    V.part = A.part + (size_t) b * A.part_stride;
    V.nx = nx; V.ny = ny; V.L = L; V.K = A.K; V.D = A.D; V.alpha2 = A.alpha2;
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = A.rp;
    V.ring_uv = reinterpret_cast<float2 *>(hs_smem);
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = reinterpret_cast<float *>(V.cxy + (size_t) V.CD * V.rp);
    V.esum = reinterpret_cast<double *>(V.crho + (size_t) V.CD * V.rp);
    V.limit = A.max_iter;
    V.account = 1;

    for (int i = tid; i < V.rp; i += nthreads) V.esum[i] = 0.0;
    __syncthreads();

    HsCpAsync cp;
    const int T_first = -4 - P;
    // row positions without divisions: this thread's first row, the others 2*nthreads further back each
    const int step_dn = (2 * nthreads) / L, step_dj = (2 * nthreads) % L;
    int niter = 0;
    {   // speculative phase
        int decided = 0, t_dec = hs::pipe_t_done(0, L, nx, ny) + 1;
        hs::PipeStep s = hs::pipe_make_step(V, T_first);
        hs::RowPos base = hs::pipe_pos(T_first - 2 * tid, L);
        for (int T = T_first;; T++, hs::pipe_advance(V, s)) {
            hs_cp_async_wait<P>();
            __syncthreads();
            if (T == t_dec) {
                if (tid < 32) {
                    double e = 0.0;
                    const double *row = V.part + (size_t) (decided % V.D) * V.rp;
                    for (int r = tid; r < ny; r += 32) e += row[r];
                    e = warp_sum(e);
                    if (tid == 0) {
                        const double error = sqrt(e / (double) (nx * ny));                 // :230
                        s_err = error;
                        s_go = (error > A.tol && decided + 1 < A.max_iter) ? 1 : 0;        // :143
                    }
                }
                __syncthreads();
                niter = ++decided;
                if (!s_go) break;
                t_dec += L;
            }
            hs::RowPos p = base;
            for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                hs::pipe_issue_row(V, s, i, p, cp);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (T >= 1) {
                p = base;
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pipe_compute_row(V, s, i, p);
            }
            base = hs::pipe_pos_add(base, 1, L);
        }
        hs_cp_async_wait<0>();
        __syncthreads();
    }
    if (niter < A.max_iter) {
        // stopped by TOL: the upper rows ran ahead.  Restore the snapshot at or before sweep niter, replay.
        const int m = (niter / V.K) * V.K;
        const float2 *src = ((niter / V.K) & 1) ? V.snap1 : V.snap0;
        for (size_t k = tid; k < n; k += nthreads) V.wuv[k] = src[k];
        __syncthreads();
        const int rep = niter - m;
        if (rep > 0) {
            V.limit = rep;
            V.account = 0;
            const int T_last = hs::pipe_t_done(rep - 1, L, nx, ny);
            hs::PipeStep s = hs::pipe_make_step(V, T_first);
            hs::RowPos base = hs::pipe_pos(T_first - 2 * tid, L);
            for (int T = T_first; T <= T_last; T++, hs::pipe_advance(V, s)) {
                hs_cp_async_wait<P>();
                __syncthreads();
                hs::RowPos p = base;
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pipe_issue_row(V, s, i, p, cp);
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (T >= 1) {
                    p = base;
                    for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                        hs::pipe_compute_row(V, s, i, p);
                }
                base = hs::pipe_pos_add(base, 1, L);
            }
            hs_cp_async_wait<0>();
            __syncthreads();
        }
    }
    if (tid == 0) {
        A.stat_iters[(size_t) b * A.stat_stride + A.stat_slot] = niter;
        A.stat_errs[(size_t) b * A.stat_stride + A.stat_slot] = s_err;
        atomicAdd(A.px_iters + A.level, (unsigned long long) niter * (unsigned long long) (nx * ny));
    }
}

// ------------------------------------------------------------------------------------------------
// Two columns per thread-step (hs_sor_pairs.h).  EXPERIMENTAL: the step functions are verified by the CPU
// replay, these wrappers have not run on a GPU yet; selected only by HS_PAIRS=1 / hook code -4.
// ------------------------------------------------------------------------------------------------

// Row-major planes -> pair wave layout: pair (i, c) = columns 2c, 2c+1 at W[((c + 2i) mod L) * ny + i];
// wuv / wxy 16-byte elements, wrho 8-byte elements; snap: initial snapshot of the flow.
__global__ void __launch_bounds__(256)
k_hs_to_wave_pairs(float *__restrict__ state, const float *__restrict__ consts, size_t plane0, size_t field_stride,
                   size_t set_stride, const PairCtl *__restrict__ ctl, Level lv, int L, float *__restrict__ snap,
                   size_t snap_stride)
{
    __shared__ float2 tile[5][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int b = blockIdx.z;
    const int cur = ctl[b].cur;
    const int nx = lv.nx, ny = lv.ny, pitch = lv.pitch, cn = hs::pairs_cn(nx);
    const int w0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const float *src[5];
    src[0] = state + (size_t) cur * set_stride + (size_t) F_U1 * field_stride + (size_t) b * plane0;
    src[1] = state + (size_t) cur * set_stride + (size_t) F_U2 * field_stride + (size_t) b * plane0;
    src[2] = consts + (size_t) C_IX * field_stride + (size_t) b * plane0;
    src[3] = consts + (size_t) C_IY * field_stride + (size_t) b * plane0;
    src[4] = consts + (size_t) C_RHO * field_stride + (size_t) b * plane0;
    const size_t n = (size_t) L * ny;
    float *base = hs_wave_base(state, set_stride, plane0, cur, b);
    float4 *wuv = reinterpret_cast<float4 *>(base);
    float4 *wxy = reinterpret_cast<float4 *>(base + 4 * n);
    float2 *wrho = reinterpret_cast<float2 *>(base + 8 * n);
    float4 *snap0 = snap ? reinterpret_cast<float4 *>(snap + (size_t) b * snap_stride) : nullptr;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = i0 + ty + 8 * r, w = w0 + tx;
        if (i < ny && w < L) {
            const int c = hs::pmod(w - 2 * i, L);
            if (c < cn) {
                const int o = i * pitch + 2 * c;            // even, pitch is a multiple of 4: 8-byte aligned
                const bool two = 2 * c + 1 < nx;
#pragma unroll
                for (int k = 0; k < 5; k++)
                    tile[k][ty + 8 * r][tx] = two ? *reinterpret_cast<const float2 *>(src[k] + o)
                                                  : make_float2(src[k][o], 0.f);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int w = w0 + ty + 8 * r, i = i0 + tx;
        if (i < ny && w < L && hs::pmod(w - 2 * i, L) < cn) {
            const size_t o = (size_t) w * ny + i;
            const float2 u = tile[0][tx][ty + 8 * r], v = tile[1][tx][ty + 8 * r];
            const float2 gx = tile[2][tx][ty + 8 * r], gy = tile[3][tx][ty + 8 * r];
            const float4 uv = make_float4(u.x, v.x, u.y, v.y);
            wuv[o] = uv;
            if (snap0) snap0[o] = uv;
            wxy[o] = make_float4(gx.x, gy.x, gx.y, gy.y);
            wrho[o] = tile[4][tx][ty + 8 * r];
        }
    }
}

__global__ void __launch_bounds__(256)
k_hs_from_wave_pairs(float *__restrict__ state, size_t plane0, size_t field_stride, size_t set_stride,
                     const PairCtl *__restrict__ ctl, Level lv, int L)
{
    __shared__ float4 tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int b = blockIdx.z;
    const int cur = ctl[b].cur;
    const int nx = lv.nx, ny = lv.ny, pitch = lv.pitch, cn = hs::pairs_cn(nx);
    const int w0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const float4 *wuv = reinterpret_cast<const float4 *>(hs_wave_base(state, set_stride, plane0, cur, b));
    float *dst = state + (size_t) cur * set_stride + (size_t) b * plane0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int w = w0 + ty + 8 * r, i = i0 + tx;
        if (i < ny && w < L && hs::pmod(w - 2 * i, L) < cn) tile[ty + 8 * r][tx] = wuv[(size_t) w * ny + i];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int i = i0 + ty + 8 * r, w = w0 + tx;
        if (i < ny && w < L) {
            const int c = hs::pmod(w - 2 * i, L);
            if (c < cn) {
                const float4 uv = tile[tx][ty + 8 * r];
                const int o = i * pitch + 2 * c;
                dst[(size_t) F_U1 * field_stride + o] = uv.x;
                dst[(size_t) F_U2 * field_stride + o] = uv.y;
                if (2 * c + 1 < nx) {
                    dst[(size_t) F_U1 * field_stride + o + 1] = uv.z;
                    dst[(size_t) F_U2 * field_stride + o + 1] = uv.w;
                }
            }
        }
    }
}

struct HsCpAsyncPairs {
    __device__ __forceinline__ void cp8(float2 *dst, const float2 *src) { cp_async8(dst, src); }
    __device__ __forceinline__ void cp16(float4 *dst, const float4 *src)
    {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned int) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
    }
};

// Bytes of shared memory of k_hs_sor_pairs<P>: rings of 16-byte pairs + coefficient rings + one double per row.
__host__ __device__ inline size_t hs_pairs_smem(int P, int rp)
{
    return (size_t) (16 * (hs::kRingBase + P) + 24 * (P + 2) + 8) * rp;
}

template <int P>
__global__ void __launch_bounds__(kHsMaxThreads)
k_hs_sor_pairs(HsSorParams A)
{
    extern __shared__ __align__(16) float hs_smem[];
    __shared__ double s_err;
    __shared__ int s_go;

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int b = blockIdx.x;
    const int nx = A.nx, ny = A.ny, L = A.L;
    float *wave = hs_wave_base(A.state, A.set_stride, A.plane0, A.ctl[b].cur, b);
    const size_t n = (size_t) L * ny;

    hs::PairsView V;
    V.wuv = reinterpret_cast<float4 *>(wave);
    V.wxy = reinterpret_cast<const float4 *>(wave + 4 * n);
    V.wrho = reinterpret_cast<const float2 *>(wave + 8 * n);
    V.snap0 = reinterpret_cast<float4 *>(A.snap + (size_t) b * A.snap_stride);
    V.snap1 = V.snap0 + n;
    V.part = A.part + (size_t) b * A.part_stride;
    V.nx = nx; V.ny = ny; V.L = L; V.K = A.K; V.D = A.D; V.alpha2 = A.alpha2;
    V.cl = hs::pairs_cl(nx); V.cn = hs::pairs_cn(nx);
    V.P = P; V.S = hs::kRingBase + P; V.CD = P + 2; V.rp = A.rp;
    V.ring_uv = reinterpret_cast<float4 *>(hs_smem);
    V.cxy = V.ring_uv + (size_t) V.S * V.rp;
    V.crho = reinterpret_cast<float2 *>(V.cxy + (size_t) V.CD * V.rp);
    V.esum = reinterpret_cast<double *>(V.crho + (size_t) V.CD * V.rp);
    V.limit = A.max_iter;
    V.account = 1;
    hs::PipeView W;                      // pipe_make_step / pipe_advance read S, CD, L, P only
    W.S = V.S; W.CD = V.CD; W.L = L; W.P = P;

    for (int i = tid; i < V.rp; i += nthreads) V.esum[i] = 0.0;
    __syncthreads();

    HsCpAsyncPairs cp;
    const int T_first = -4 - P;
    const int step_dn = (2 * nthreads) / L, step_dj = (2 * nthreads) % L;
    int niter = 0;
    {   // speculative phase
        int decided = 0, t_dec = hs::pairs_t_done(0, L, nx, ny) + 1;
        hs::PipeStep s = hs::pipe_make_step(W, T_first);
        hs::RowPos base = hs::pipe_pos(T_first - 2 * tid, L);
        for (int T = T_first;; T++, hs::pipe_advance(W, s)) {
            hs_cp_async_wait<P>();
            __syncthreads();
            if (T == t_dec) {
                if (tid < 32) {
                    double e = 0.0;
                    const double *row = V.part + (size_t) (decided % V.D) * V.rp;
                    for (int r = tid; r < ny; r += 32) e += row[r];
                    e = warp_sum(e);
                    if (tid == 0) {
                        const double error = sqrt(e / (double) (nx * ny));
                        s_err = error;
                        s_go = (error > A.tol && decided + 1 < A.max_iter) ? 1 : 0;
                    }
                }
                __syncthreads();
                niter = ++decided;
                if (!s_go) break;
                t_dec += L;
            }
            hs::RowPos p = base;
            for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                hs::pairs_issue_row(V, s, i, p, cp);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (T >= 0) {
                p = base;
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pairs_compute_row(V, s, i, p);
            }
            base = hs::pipe_pos_add(base, 1, L);
        }
        hs_cp_async_wait<0>();
        __syncthreads();
    }
    if (niter < A.max_iter) {
        const int m = (niter / V.K) * V.K;
        const float4 *src = ((niter / V.K) & 1) ? V.snap1 : V.snap0;
        for (size_t k = tid; k < n; k += nthreads) V.wuv[k] = src[k];
        __syncthreads();
        const int rep = niter - m;
        if (rep > 0) {
            V.limit = rep;
            V.account = 0;
            const int T_last = hs::pairs_t_done(rep - 1, L, nx, ny);
            hs::PipeStep s = hs::pipe_make_step(W, T_first);
            hs::RowPos base = hs::pipe_pos(T_first - 2 * tid, L);
            for (int T = T_first; T <= T_last; T++, hs::pipe_advance(W, s)) {
                hs_cp_async_wait<P>();
                __syncthreads();
                hs::RowPos p = base;
                for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                    hs::pairs_issue_row(V, s, i, p, cp);
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (T >= 0) {
                    p = base;
                    for (int i = tid; i < ny; i += nthreads, p = hs::pipe_pos_sub(p, step_dn, step_dj, L))
                        hs::pairs_compute_row(V, s, i, p);
                }
                base = hs::pipe_pos_add(base, 1, L);
            }
            hs_cp_async_wait<0>();
            __syncthreads();
        }
    }
    if (tid == 0) {
        A.stat_iters[(size_t) b * A.stat_stride + A.stat_slot] = niter;
        A.stat_errs[(size_t) b * A.stat_stride + A.stat_slot] = s_err;
        atomicAdd(A.px_iters + A.level, (unsigned long long) niter * (unsigned long long) (nx * ny));
    }
}

} // namespace tvl1
