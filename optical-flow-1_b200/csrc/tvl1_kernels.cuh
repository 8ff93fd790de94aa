// Device kernels of the B200 TV-L1 solver (sm_100a).  fp32 storage and arithmetic, fp64 only for
// the stopping-test reduction and for pyramid sample coordinates.
//
// Data layout in HBM:
//   * every image / field is a row-major plane with a row pitch that is a multiple of 4 floats, so
//     every row start is 16-byte aligned (float4 / TMA friendly).  Columns [nx, pitch) are padding:
//     never read into a valid result, may hold garbage.
//   * a batch of B pairs stores plane b at  base + b * plane0   (plane0 = pitch*ny of the finest
//     level, fixed for all levels so buffers are reused across the pyramid).
//   * flow + dual state: state[set][field][b][plane0], set in {0,1} (ping-pong), field in
//     {u1,u2,p11,p12,p21,p22}.  PairCtl[b].cur names the live set of pair b.
//   * per-warp constants: consts[field][b][plane0], field in {I1wx, I1wy, rho_c, grad}.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <limits.h>
#include <stdint.h>

namespace tvl1 {

constexpr int kMaxIterations = 300;       // src/tvl1flow.cpp:22
constexpr float kGradIsZero = 1e-10f;     // src/tvl1flow.cpp:24
constexpr int kStatLevels = 16;           // == TVL1_MAX_LEVELS
constexpr int kCounterWords = 3 * kStatLevels;   // per level: pixel-iterations, iteration launches, sum over the loops of a
                                          // solve of min(iterations, kLoopClip) (streamed levels: decide_block)
constexpr int kLoopClip = 16;
constexpr int kTbT = 4;                   // iterations per launch of the temporally blocked kernel
constexpr int kMaxTaps = 16;              // (int)(5*sigma)+1 <= 16  <=>  sigma < 3.2 (zfactor > 0.19)

enum Field { F_U1 = 0, F_U2, F_P11, F_P12, F_P21, F_P22, F_COUNT };
enum Const { C_IX = 0, C_IY, C_RHO, C_GRAD, C_COUNT };

struct Level {
    int nx, ny, pitch;
};

// Per-pair loop control, owned by the device (src/tvl1flow.cpp:111-113 lives here).
struct PairCtl {
    int cur;              // live state set
    int active;           // 1 while the current warp step still iterates
    int n;                // iterations done in the current warp step
    unsigned int arrive;  // CTA arrival counter (last-block election)
    double err;           // mean squared update of the last iteration
    // temporal blocking (k_iterate_tb): iterations the next launch runs for this pair, and whether
    // that launch is the exact replay of a block that overshot the stopping point
    int nsteps;
    int replay;
    // a level's first block ran two iterations from zero duals (k_iterate_t2 with p_zero) and overshot the stopping
    // point: the replay of its first iteration takes the duals as zero again (the start state's were never written)
    int pzero;
    int pad;
};

// Whole-batch loop state of the current warp step.
struct LoopCtl {
    int active_pairs;     // pairs still iterating
    int max_n;            // largest iteration count among the pairs that already stopped
};

struct GaussTaps {
    int size;             // taps on one side incl. centre: (int)(5*sigma)+1, src/operators.cpp:515
    float w[kMaxTaps];
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void st4(float *p, float4 v)
{
    *reinterpret_cast<float4 *>(p) = v;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi)
{
    return min(max(v, lo), hi);
}
__device__ __forceinline__ unsigned int smem_u32(const void *p)
{
    return (unsigned int) __cvta_generic_to_shared(p);
}
// one box of a 3-D tensor map -> shared memory, completion on an mbarrier (SASS: UTMALDG.3D)
__device__ __forceinline__ void tma_load_3d(float *dst, const CUtensorMap *map, int x, int y, int z,
                                            unsigned long long *mbar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z),
                    "r"(smem_u32(mbar)) : "memory");
}
// order-preserving float <-> uint map for atomicMin/atomicMax
__device__ __forceinline__ unsigned int f2ord(float f)
{
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// (a) pyramid
// ------------------------------------------------------------------------------------------------

// Joint min / max of both images of each pair: getminmax x2, src/utils.cpp:509-525 and :293-308.
// in0/in1: dense [B][n] floats.  mm[b] = {ord(min), ord(max)}, pre-set to {~0u, 0u}.
__global__ void k_minmax(const float *__restrict__ in0, const float *__restrict__ in1, size_t n,
                         unsigned int *__restrict__ mm)
{
    const int b = blockIdx.y;
    const float *a = in0 + (size_t) b * n;
    const float *c = in1 + (size_t) b * n;
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t) gridDim.x * blockDim.x) {
        const float x = __ldg(a + i), y = __ldg(c + i);
        lo = fminf(lo, fminf(x, y));
        hi = fmaxf(hi, fmaxf(x, y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[32], shi[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { slo[w] = lo; shi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        lo = lane < nw ? slo[lane] : INFINITY;
        hi = lane < nw ? shi[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(&mm[2 * b], f2ord(lo));
            atomicMax(&mm[2 * b + 1], f2ord(hi));
        }
    }
}

// One pixel of image_normalization_2: 255*(I-min)/den evaluated left to right as the reference does
// (src/utils.cpp:315-316), every step IEEE-rounded and never contracted.
__device__ __forceinline__ float normalize_px(float v, float mn, float den)
{
    return __fdiv_rn(__fmul_rn(255.0f, __fsub_rn(v, mn)), den);
}

// Separable Gaussian with the reference's boundary rule, optionally fused with the [0,255]
// normalisation on load and with a decimation by D on store.
//   gaussian            src/operators.cpp:506-624  (rows then columns, reflecting boundary:
//                       x<0 -> -x, x>=n -> 2n-1-x; sum order B0*c + sum_j Bj*(l_j + r_j))
//   image_normalization src/utils.cpp:310-318      ("255*(I-min)/den" when den > 0)
//   zoom_out, f = 1/D   src/zoom.cpp:41-78         (sample points j/f are integers, where the cubic
//                       returns the centre tap exactly -> blur followed by [D*i, D*j] decimation)
// One CTA produces a TW x TH output tile from a (D*TW+2r) x (D*TH+2r) input tile in shared memory.
// z = image index in [0, nimg); pair (for min/max) = z % B.
// RT = compile-time radius (taps.size - 1) for the two windows the default parameters use
// (sigma 0.8 -> 4, zoom sigma at zfactor 0.5 -> 5): weights live in registers and the tap loops
// unroll; RT = 0 is the generic run-time radius.
template <int D, int RT>
__global__ void __launch_bounds__(256)
k_gauss(const float *__restrict__ in, int in_pitch, size_t in_stride, float *__restrict__ out,
        int out_pitch, size_t out_stride, int nx, int ny, int onx, int ony, const GaussTaps taps,
        const unsigned int *__restrict__ mm, int B)
{
    constexpr int TW = 64 / D, TH = 32 / D;        // output tile
    constexpr int IW = 64, IH = 32;                // input footprint without halo
    constexpr int RMAX = RT > 0 ? RT : kMaxTaps - 1;
    constexpr int NW = RT > 0 ? RT + 1 : kMaxTaps;
    __shared__ float s_in[(IH + 2 * RMAX)][IW + 2 * RMAX + 1];
    __shared__ float s_row[(IH + 2 * RMAX)][TW + 1];

    const int tx = threadIdx.x, ty = threadIdx.y;  // block (32, 8)
    const int r = RT > 0 ? RT : taps.size - 1;
    const int z = blockIdx.z;
    const float *src = in + (size_t) z * in_stride;
    float *dst = out + (size_t) z * out_stride;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int ix0 = ox0 * D - r, iy0 = oy0 * D - r;
    const int iw = IW + 2 * r - (D - 1), ih = IH + 2 * r - (D - 1);

    float w[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = taps.w[i];

    // image_normalization_2 fused on load, in the reference's own operation order
    // 255*(I-min)/den (src/utils.cpp:315-316): subtract, multiply, IEEE-rounded divide
    float mn = 0.f, den = 1.f;
    bool norm = false;
    if (mm) {
        mn = ord2f(mm[2 * (z % B)]);
        den = ord2f(mm[2 * (z % B) + 1]) - mn;
        norm = den > 0.f;
    }
    // tiles that do not touch the border skip the reflection arithmetic
    const bool interior = ix0 >= 0 && iy0 >= 0 && ix0 + iw <= nx && iy0 + ih <= ny;

    for (int ly = ty; ly < ih; ly += 8) {
        int gy = iy0 + ly;
        if (!interior) {
            gy = gy < 0 ? -gy : (gy >= ny ? 2 * ny - 1 - gy : gy);
            gy = clampi(gy, 0, ny - 1);   // only reachable for tiles hanging over the image edge
        }
        const float *row = src + (size_t) gy * in_pitch;
        for (int lx = tx; lx < iw; lx += 32) {
            int gx = ix0 + lx;
            if (!interior) {
                gx = gx < 0 ? -gx : (gx >= nx ? 2 * nx - 1 - gx : gx);
                gx = clampi(gx, 0, nx - 1);
            }
            float v = __ldg(row + gx);
            if (norm) v = normalize_px(v, mn, den);
            s_in[ly][lx] = v;
        }
    }
    __syncthreads();

    // row pass: every input row of the tile, output columns only
    for (int ly = ty; ly < ih; ly += 8) {
#pragma unroll
        for (int ox = tx; ox < TW; ox += 32) {
            const float *c = &s_in[ly][ox * D + r];
            float sum = w[0] * c[0];
            if (RT > 0) {
#pragma unroll
                for (int j = 1; j <= RT; j++) sum += w[j] * (c[-j] + c[j]);
            } else {
                for (int j = 1; j <= r; j++) sum += taps.w[j] * (c[-j] + c[j]);
            }
            s_row[ly][ox] = sum;
        }
    }
    __syncthreads();

    // column pass on output rows
#pragma unroll
    for (int oy = ty; oy < TH; oy += 8) {
        const int gy = oy0 + oy;
        if (gy >= ony) break;
        const int c = oy * D + r;
#pragma unroll
        for (int ox = tx; ox < TW; ox += 32) {
            const int gx = ox0 + ox;
            if (gx >= onx) break;
            float sum = w[0] * s_row[c][ox];
            if (RT > 0) {
#pragma unroll
                for (int j = 1; j <= RT; j++) sum += w[j] * (s_row[c - j][ox] + s_row[c + j][ox]);
            } else {
                for (int j = 1; j <= r; j++) sum += taps.w[j] * (s_row[c - j][ox] + s_row[c + j][ox]);
            }
            dst[(size_t) gy * out_pitch + gx] = sum;
        }
    }
}

// Register-marching variant of the separable blur for the windows the default parameters use
// (R = 4: sigma 0.8; R = 5: zoom sigma at zfactor 0.5).  A thread owns 4 output columns and walks down
// a strip of output rows: per input row it loads the 3D+2R+1 values its row pass needs (aligned float4
// loads away from the left/right border), forms the 4 row-pass results, and keeps the last 2R+1 of
// them in registers for the column pass -- no shared memory, every input element is read once per
// strip (+2R halo rows).  Same boundary rule and summation order as k_gauss.
template <int D, int R>
__global__ void __launch_bounds__(128)
k_gauss_march(const float *__restrict__ in, int in_pitch, size_t in_stride, float *__restrict__ out,
              int out_pitch, size_t out_stride, int nx, int ny, int onx, int ony, const GaussTaps taps,
              const unsigned int *__restrict__ mm, int B)
{
    constexpr int S = 32;                         // output rows per strip
    constexpr int W = 2 * R + 1;                  // window of row-pass results
    constexpr int NIN = 3 * D + 2 * R + 1;        // input columns feeding 4 outputs
    constexpr int OFF = (4 - (R % 4)) % 4;        // ix0 = 4*D*k - R  ->  aligned start is OFF columns earlier
    constexpr int NV4 = (OFF + NIN + 3) / 4;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int z = blockIdx.z;
    const int ox0 = (blockIdx.x * 32 + lane) * 4;
    const int oy0 = (blockIdx.y * 4 + warp) * S;
    if (oy0 >= ony || ox0 >= onx) return;
    const int oy1 = min(oy0 + S, ony);
    const float *src = in + (size_t) z * in_stride;
    float *dst = out + (size_t) z * out_stride;
    const int ix0 = ox0 * D - R;
    const int ia = ix0 - OFF;
    const bool fast = ia >= 0 && ia + 4 * NV4 <= nx && (in_pitch & 3) == 0 &&
                      ((reinterpret_cast<size_t>(src) & 15) == 0);

    float w[R + 1];
#pragma unroll
    for (int i = 0; i <= R; i++) w[i] = taps.w[i];
    float mn = 0.f, den = 1.f;
    bool norm = false;
    if (mm) {
        mn = ord2f(mm[2 * (z % B)]);
        den = ord2f(mm[2 * (z % B) + 1]) - mn;
        norm = den > 0.f;
    }

    float win[W][4];
#pragma unroll
    for (int u = 0; u < W; u++) win[u][0] = win[u][1] = win[u][2] = win[u][3] = 0.f;

    const int iy_first = oy0 * D - R, iy_last = (oy1 - 1) * D + R;
    for (int base = iy_first; base <= iy_last; base += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const int iy = base + u;
            if (iy > iy_last) break;
            int gy = iy < 0 ? -iy : (iy >= ny ? 2 * ny - 1 - iy : iy);
            gy = clampi(gy, 0, ny - 1);
            const float *row = src + (size_t) gy * in_pitch;
            float v[4 * NV4];
            if (fast) {
#pragma unroll
                for (int q = 0; q < NV4; q++) {
                    const float4 t = ldg4(row + ia + 4 * q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < NIN; k++) {
                    int gx = ix0 + k;
                    gx = gx < 0 ? -gx : (gx >= nx ? 2 * nx - 1 - gx : gx);
                    v[OFF + k] = __ldg(row + clampi(gx, 0, nx - 1));
                }
            }
            if (norm) {
#pragma unroll
                for (int k = 0; k < NIN; k++) v[OFF + k] = normalize_px(v[OFF + k], mn, den);
            }
            // row pass at the four output columns (window slot u)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = OFF + R + k * D;
                float sum = w[0] * v[c];
#pragma unroll
                for (int j = 1; j <= R; j++) sum += w[j] * (v[c - j] + v[c + j]);
                win[u][k] = sum;
            }
            // the window now ends at input row iy: it is centred on row iy - R
            const int cy = iy - R;
            if (cy >= oy0 * D && (cy % D) == 0) {
                const int oy = cy / D;
                if (oy < oy1) {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float sum = w[0] * win[(u + W - R) % W][k];
#pragma unroll
                        for (int j = 1; j <= R; j++)
                            sum += w[j] * (win[(u + W - R - j) % W][k] + win[(u + W - R + j) % W][k]);
                        o[k] = sum;
                    }
                    float *orow = dst + (size_t) oy * out_pitch + ox0;
                    if (ox0 + 3 < onx && (out_pitch & 3) == 0 && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
                        st4(orow, make_float4(o[0], o[1], o[2], o[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (ox0 + k < onx) orow[k] = o[k];
                    }
                }
            }
        }
    }
}

// Same marching blur with the row-pass inputs shared through warp shuffles (round 2).  A lane loads only
// ITS OWN 4*D input columns of a row (D aligned float4 loads), normalises them once, and takes the R
// columns to its left and the R-D+1 to its right from its lane neighbours: one load and -- on the
// normalising pass -- one IEEE division per input element instead of three (k_gauss_march reads and
// normalises the 3D+2R+1 columns behind its four outputs itself, so neighbouring lanes repeat each
// other's work).  Lanes 0 and 31 of a warp only load (their outputs belong to the neighbouring warps): a
// warp produces 30 x 4 output columns.  Same boundary rule, same summation order, same bits.
// Requires R <= 4*D and R-D+1 <= 4*D (the default windows: D=1,R=4 and D=2,R=5).
template <int D, int R>
__global__ void __launch_bounds__(128)
k_gauss_shfl(const float *__restrict__ in, int in_pitch, size_t in_stride, float *__restrict__ out,
             int out_pitch, size_t out_stride, int nx, int ny, int onx, int ony, const GaussTaps taps,
             const unsigned int *__restrict__ mm, int B)
{
    constexpr int S = 32;                         // output rows per strip
    constexpr int W = 2 * R + 1;                  // window of row-pass results
    constexpr int OWN = 4 * D;                    // input columns a lane loads
    constexpr int NR = R - D + 1;                 // columns taken from the right neighbour
    static_assert(R <= OWN && NR <= OWN && NR >= 0, "window does not fit the neighbours' blocks");
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int z = blockIdx.z;
    const int grp = blockIdx.x * 30 + lane - 1;   // column group of this lane (outputs 4*grp .. 4*grp+3)
    const int ox0 = grp * 4;
    const int oy0 = (blockIdx.y * 4 + warp) * S;
    if (oy0 >= ony || blockIdx.x * 120 >= onx) return;          // warp-uniform
    const int oy1 = min(oy0 + S, ony);
    const float *src = in + (size_t) z * in_stride;
    float *dst = out + (size_t) z * out_stride;
    const int ix0 = grp * OWN;                    // first own input column (may be < 0 or >= nx: reflected)
    const bool fast = ix0 >= 0 && ix0 + OWN <= nx && (in_pitch & 3) == 0 &&
                      ((reinterpret_cast<size_t>(src) & 15) == 0);
    const bool writer = lane >= 1 && lane <= 30 && ox0 < onx;

    float w[R + 1];
#pragma unroll
    for (int i = 0; i <= R; i++) w[i] = taps.w[i];
    float mn = 0.f, den = 1.f;
    bool norm = false;
    if (mm) {
        mn = ord2f(mm[2 * (z % B)]);
        den = ord2f(mm[2 * (z % B) + 1]) - mn;
        norm = den > 0.f;
    }

    float win[W][4];
#pragma unroll
    for (int u = 0; u < W; u++) win[u][0] = win[u][1] = win[u][2] = win[u][3] = 0.f;

    const int iy_first = oy0 * D - R, iy_last = (oy1 - 1) * D + R;
    // A warp walks down its strip row by row, so without help it has ONE row of loads in flight: the own
    // columns of the row kGaussPf rows further down are pulled into L2 ahead (no registers: a register ring
    // of prefetched rows was measured slower, 137 registers).
    constexpr int kGaussPf = 8;
    auto row_of = [&](int iy) {
        int gy = iy < 0 ? -iy : (iy >= ny ? 2 * ny - 1 - iy : iy);
        return src + (size_t) clampi(gy, 0, ny - 1) * in_pitch;
    };
    for (int base = iy_first; base <= iy_last; base += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const int iy = base + u;
            if (iy > iy_last) break;
            const float *row = row_of(iy);
            if (fast && iy + kGaussPf <= iy_last)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(row_of(iy + kGaussPf) + ix0));
            // v[R .. R+OWN) = own columns, v[0 .. R) from the left neighbour, v[R+OWN .. R+OWN+NR) from the right
            float v[R + OWN + NR];
            if (fast) {
#pragma unroll
                for (int q = 0; q < D; q++) {
                    const float4 t = ldg4(row + ix0 + 4 * q);
                    v[R + 4 * q] = t.x; v[R + 4 * q + 1] = t.y; v[R + 4 * q + 2] = t.z; v[R + 4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < OWN; k++) {
                    int gx = ix0 + k;
                    gx = gx < 0 ? -gx : (gx >= nx ? 2 * nx - 1 - gx : gx);
                    v[R + k] = __ldg(row + clampi(gx, 0, nx - 1));
                }
            }
            if (norm) {
#pragma unroll
                for (int k = 0; k < OWN; k++) v[R + k] = normalize_px(v[R + k], mn, den);
            }
#pragma unroll
            for (int m = 0; m < R; m++) v[m] = __shfl_up_sync(0xffffffffu, v[R + OWN - R + m], 1);
#pragma unroll
            for (int m = 0; m < NR; m++) v[R + OWN + m] = __shfl_down_sync(0xffffffffu, v[R + m], 1);
            // row pass at the four output columns (window slot u)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = R + k * D;
                float sum = w[0] * v[c];
#pragma unroll
                for (int j = 1; j <= R; j++) sum += w[j] * (v[c - j] + v[c + j]);
                win[u][k] = sum;
            }
            // the window now ends at input row iy: it is centred on row iy - R
            const int cy = iy - R;
            if (cy >= oy0 * D && (cy % D) == 0) {
                const int oy = cy / D;
                if (oy < oy1 && writer) {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float sum = w[0] * win[(u + W - R) % W][k];
#pragma unroll
                        for (int j = 1; j <= R; j++)
                            sum += w[j] * (win[(u + W - R - j) % W][k] + win[(u + W - R + j) % W][k]);
                        o[k] = sum;
                    }
                    float *orow = dst + (size_t) oy * out_pitch + ox0;
                    if (ox0 + 3 < onx && (out_pitch & 3) == 0 && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
                        st4(orow, make_float4(o[0], o[1], o[2], o[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (ox0 + k < onx) orow[k] = o[k];
                    }
                }
            }
        }
    }
}

// Keys cubic through v0..v3 at offset t from v1: src/bicubic_interpolation.cpp:108-123.
__device__ __forceinline__ float cubic_cell(float v0, float v1, float v2, float v3, float t)
{
    return v1 + 0.5f * t * (v2 - v0 + t * (2.0f * v0 - 5.0f * v1 + 4.0f * v2 - v3
                                           + t * (3.0f * (v1 - v2) + v3 - v0)));
}

// Catmull-Rom / Keys(a=-1/2) weights of the four taps at fraction t; algebraically the cubic of
// src/bicubic_interpolation.cpp:108-123 written as a dot product so that one set of weights
// serves I1, dI1/dx and dI1/dy.
__device__ __forceinline__ void keys_weights(float t, float w[4])
{
    const float t2 = t * t, t3 = t2 * t;
    w[0] = 0.5f * (-t + 2.0f * t2 - t3);
    w[1] = 0.5f * (2.0f - 5.0f * t2 + 3.0f * t3);
    w[2] = 0.5f * (t + 4.0f * t2 - 3.0f * t3);
    w[3] = 0.5f * (t3 - t2);
}

// bicubic_interpolation_at with border_out = false and non-negative coordinates
// (src/bicubic_interpolation.cpp:153-245): neighbours x-1,x,x+1,x+2 index-clamped (Neumann),
// fractions relative to the (unclamped, since 0 <= uu < nx) base index; y first inside each
// x column, then x.
__device__ __forceinline__ float bicubic_clamped(const float *__restrict__ img, int pitch, int nx,
                                                 int ny, double uu, double vv)
{
    const int x = clampi((int) uu, 0, nx - 1), y = clampi((int) vv, 0, ny - 1);
    const float tx = (float) (uu - x), ty = (float) (vv - y);
    const int xs[4] = { max(x - 1, 0), x, min(x + 1, nx - 1), min(x + 2, nx - 1) };
    const int ys[4] = { max(y - 1, 0), y, min(y + 1, ny - 1), min(y + 2, ny - 1) };
    float col[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        col[a] = cubic_cell(__ldg(img + (size_t) ys[0] * pitch + xs[a]),
                            __ldg(img + (size_t) ys[1] * pitch + xs[a]),
                            __ldg(img + (size_t) ys[2] * pitch + xs[a]),
                            __ldg(img + (size_t) ys[3] * pitch + xs[a]), ty);
    }
    return cubic_cell(col[0], col[1], col[2], col[3], tx);
}

// Bicubic resampling of nimg planes: out(i1, j1) = scale * in(j1 / fx, i1 / fy).
//   zoom_out (general factor)  src/zoom.cpp:67-75      fx = fy = factor, scale = 1
//   zoom_in + flow rescale     src/zoom.cpp:132-155, src/tvl1flow.cpp:302-309
//                              fx = nxx/nx, fy = nyy/ny, scale = 1/zfactor
// Sample coordinates are formed in fp64 exactly as the reference does (a division), so base
// index and fraction agree bit for bit.
__global__ void k_resample(const float *__restrict__ in, int in_pitch, size_t in_stride, int nx,
                           int ny, float *__restrict__ out, int out_pitch, size_t out_stride,
                           int onx, int ony, double fx, double fy, float scale)
{
    const int j1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y * blockDim.y + threadIdx.y;
    if (j1 >= onx || i1 >= ony) return;
    const float *src = in + (size_t) blockIdx.z * in_stride;
    const double j2 = j1 / fx, i2 = i1 / fy;
    const float g = bicubic_clamped(src, in_pitch, nx, ny, j2, i2);
    out[(size_t) blockIdx.z * out_stride + (size_t) i1 * out_pitch + j1] = g * scale;
}

// zoom_in of both flow components of every pair between the ping-pong state sets:
// reads set cur (coarse level), writes set cur^1 (fine level), then k_flip_cur flips.
// One CTA makes a 64x16 tile of the fine level; the coarse footprint of the tile (index-clamped,
// so the gather below needs no boundary logic) is staged in shared memory for both components.
constexpr int kZiTW = 64, kZiTH = 16, kZiCW = kZiTW + 8, kZiCH = kZiTH + 8;

__global__ void __launch_bounds__(256)
k_zoom_in_flow(float *__restrict__ state, size_t plane0, size_t field_stride, size_t set_stride,
               const PairCtl *__restrict__ ctl, Level coarse, Level fine, double fx, double fy,
               float scale, int row_begin, int row_end)
{
    __shared__ float s_c[2][kZiCH][kZiCW];
    __shared__ float s_h[2][kZiCH][kZiTW];          // rows of the footprint interpolated to the fine columns
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int b = blockIdx.z;
    const int cur = ctl[b].cur;
    const float *src = state + (size_t) cur * set_stride + (size_t) b * plane0;
    float *dst = state + (size_t) (cur ^ 1) * set_stride + (size_t) b * plane0;
    const int X0 = blockIdx.x * kZiTW, Y0 = row_begin + blockIdx.y * kZiTH;
    const int X1 = min(X0 + kZiTW, fine.nx) - 1, Y1 = min(Y0 + kZiTH, row_end) - 1;
    // sample coordinate j1 / f in fp64 as the reference forms it; for f == 2 (every level pair of even size
    // at zfactor 0.5) the product with 0.5 is the same number without the division
    const bool hx = fx == 2.0, hy = fy == 2.0;
    auto pos_x = [&](int j1) { return hx ? j1 * 0.5 : j1 / fx; };
    auto pos_y = [&](int i1) { return hy ? i1 * 0.5 : i1 / fy; };
    // coarse footprint: taps x-1 .. x+2 around x = (int)(j1 / fx), monotone in j1
    const int xlo = (int) pos_x(X0) - 1, xhi = (int) pos_x(X1) + 2;
    const int ylo = (int) pos_y(Y0) - 1, yhi = (int) pos_y(Y1) + 2;
    const int cw = xhi - xlo + 1, ch = yhi - ylo + 1;
    const bool staged = (cw <= kZiCW) && (ch <= kZiCH);   // always true when fx, fy >= 1
    if (staged) {
        for (int ly = ty; ly < ch; ly += 8) {
            const size_t ro = (size_t) clampi(ylo + ly, 0, coarse.ny - 1) * coarse.pitch;
            for (int lx = tx; lx < cw; lx += 32) {
                const size_t o = ro + clampi(xlo + lx, 0, coarse.nx - 1);
                s_c[0][ly][lx] = __ldg(src + o);
                s_c[1][ly][lx] = __ldg(src + field_stride + o);
            }
        }
    }
    __syncthreads();
    if (!staged) {
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int j1 = X0 + tx + 32 * (q & 1), i1 = Y0 + ty + 8 * (q >> 1);
            if (j1 >= fine.nx || i1 >= row_end) continue;
            const double j2 = j1 / fx, i2 = i1 / fy;
            const size_t o = (size_t) i1 * fine.pitch + j1;
            dst[o] = bicubic_clamped(src, coarse.pitch, coarse.nx, coarse.ny, j2, i2) * scale;
            dst[field_stride + o] =
                bicubic_clamped(src + field_stride, coarse.pitch, coarse.nx, coarse.ny, j2, i2) * scale;
        }
        return;
    }
    // The sample grid is separable: a thread's four pixels share two column positions and two row
    // positions; base indices and fractions come from the reference's fp64 coordinate, the cubic is
    // applied as Keys weights (same polynomial as cubic_cell, one set per column / row).  Round 2: the
    // horizontal interpolation of a footprint row is formed ONCE per fine column (s_h) instead of once per
    // fine pixel and tap row -- the same products and sums in the same order, 2.4x fewer shared-memory loads.
    int cx[2], cy[2];
    float wx[2][4], wy[2][4];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double j2 = pos_x(X0 + tx + 32 * k), i2 = pos_y(Y0 + ty + 8 * k);
        const int x = clampi((int) j2, 0, coarse.nx - 1), y = clampi((int) i2, 0, coarse.ny - 1);
        keys_weights((float) (j2 - x), wx[k]);
        keys_weights((float) (i2 - y), wy[k]);
        cx[k] = min(x - 1 - xlo, kZiCW - 4);         // (columns beyond the image: never stored)
        cy[k] = min(y - 1 - ylo, kZiCH - 4);
    }
    for (int ly = ty; ly < ch; ly += 8) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
#pragma unroll
            for (int comp = 0; comp < 2; comp++) {
                const float *row = &s_c[comp][ly][cx[k]];
                s_h[comp][ly][tx + 32 * k] =
                    fmaf(wx[k][3], row[3], fmaf(wx[k][2], row[2], fmaf(wx[k][1], row[1], wx[k][0] * row[0])));
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int kx = q & 1, ky = q >> 1;
        const int j1 = X0 + tx + 32 * kx, i1 = Y0 + ty + 8 * ky;
        if (j1 >= fine.nx || i1 >= row_end) continue;
        const int o = i1 * fine.pitch + j1;
#pragma unroll
        for (int comp = 0; comp < 2; comp++) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < 4; r++) acc = fmaf(wy[ky][r], s_h[comp][cy[ky] + r][tx + 32 * kx], acc);
            dst[(size_t) comp * field_stride + o] = acc * scale;
        }
    }
}

__global__ void k_flip_cur(PairCtl *ctl, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) ctl[b].cur ^= 1;
}

// Zeroes the dual variable of every pair in its live set (src/tvl1flow.cpp:87-90) and, when
// zero_u is set, the flow too (coarsest level, :278-280).
__global__ void k_zero_fields(float *__restrict__ state, size_t plane0, size_t field_stride,
                              size_t set_stride, const PairCtl *__restrict__ ctl, size_t n4,
                              int first_field)
{
    const int b = blockIdx.z;
    const int f = first_field + blockIdx.y;
    float4 *dst = reinterpret_cast<float4 *>(state + (size_t) ctl[b].cur * set_stride +
                                             (size_t) f * field_stride + (size_t) b * plane0);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t) gridDim.x * blockDim.x)
        dst[i] = z;
}

// Row-band mode (one image split over several GPUs): after the all-reduce of the per-rank sums every
// rank applies the stopping rule of src/tvl1flow.cpp:113,162 to the same number.
__global__ void k_band_decide(PairCtl *ctl, LoopCtl *loop, const double *band_sum, double npix, double eps2,
                              int max_iter, int *stat_iters, double *stat_errs, int stat_slot,
                              unsigned long long *px_iters, unsigned long long own_pixels)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || !ctl->active) return;
    const double error = *band_sum / npix;
    const int n = ctl->n + 1;
    ctl->n = n;
    ctl->err = error;
    ctl->cur ^= 1;
    atomicAdd(px_iters, own_pixels);
    if (!(error > eps2 && n < max_iter)) {
        ctl->active = 0;
        stat_iters[stat_slot] = n;
        stat_errs[stat_slot] = error;
        loop->max_n = n;
        loop->active_pairs = 0;
    }
}

// start of a warp step: n = 0, error = INFINITY (src/tvl1flow.cpp:111-112)
__global__ void k_begin_warp(PairCtl *ctl, LoopCtl *loop, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        // first block of the warp step: as many iterations as the previous warp step of this pair
        // suggests (it rarely needs more than its predecessor), 1 when that one stopped at once
        const int n_old = ctl[b].n;
        ctl[b].nsteps = n_old >= 2 * kTbT ? kTbT : (n_old >= 4 ? 2 : 1);
        ctl[b].replay = 0;
        ctl[b].pzero = 0;
        ctl[b].active = 1;
        ctl[b].n = 0;
        ctl[b].arrive = 0u;
        ctl[b].err = INFINITY;
    }
    if (b == 0) { loop->active_pairs = B; loop->max_n = 0; }
}

__global__ void k_init_ctl(PairCtl *ctl, unsigned int *mm, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        ctl[b].cur = 0; ctl[b].active = 0; ctl[b].n = 0; ctl[b].arrive = 0u; ctl[b].err = 0.0;
        ctl[b].nsteps = 1; ctl[b].replay = 0;
        mm[2 * b] = 0xffffffffu;
        mm[2 * b + 1] = 0u;
    }
}

// dense <-> pitched copies of the flow (live set), and fp64 <-> fp32 conversion at the boundary
__global__ void k_export_flow(const float *__restrict__ state, size_t plane0, size_t field_stride,
                              size_t set_stride, const PairCtl *__restrict__ ctl, Level lv,
                              float *__restrict__ u1, float *__restrict__ u2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= lv.nx || i >= lv.ny) return;
    const int b = blockIdx.z;
    const float *s = state + (size_t) ctl[b].cur * set_stride + (size_t) b * plane0 + (size_t) i * lv.pitch + j;
    const size_t o = ((size_t) b * lv.ny + i) * lv.nx + j;
    u1[o] = s[0];
    u2[o] = s[field_stride];
}

__global__ void k_import_flow(float *__restrict__ state, size_t plane0, size_t field_stride,
                              size_t set_stride, const PairCtl *__restrict__ ctl, Level lv,
                              const float *__restrict__ u1, const float *__restrict__ u2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= lv.nx || i >= lv.ny) return;
    const int b = blockIdx.z;
    float *s = state + (size_t) ctl[b].cur * set_stride + (size_t) b * plane0 + (size_t) i * lv.pitch + j;
    const size_t o = ((size_t) b * lv.ny + i) * lv.nx + j;
    s[0] = u1[o];
    s[field_stride] = u2[o];
}

// dense [nimg][ny][nx] -> pitched planes (used where no blur precedes: single-scale entry, hooks)
__global__ void k_pack(const float *__restrict__ in, float *__restrict__ out, int nx, int ny,
                       int pitch, size_t out_stride)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    out[(size_t) blockIdx.z * out_stride + (size_t) i * pitch + j] =
        in[((size_t) blockIdx.z * ny + i) * nx + j];
}
__global__ void k_unpack(const float *__restrict__ in, float *__restrict__ out, int nx, int ny,
                         int pitch, size_t in_stride)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    out[((size_t) blockIdx.z * ny + i) * nx + j] =
        in[(size_t) blockIdx.z * in_stride + (size_t) i * pitch + j];
}

__global__ void k_f64_to_f32(const double *__restrict__ in, float *__restrict__ out, size_t n)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t) gridDim.x * blockDim.x)
        out[i] = (float) in[i];
}
// 8-bit frames (video, PGM): four pixels per thread
__global__ void k_u8_to_f32(const unsigned char *__restrict__ in, float *__restrict__ out, size_t n)
{
    const size_t n4 = n / 4;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) {
        const uchar4 v = reinterpret_cast<const uchar4 *>(in)[i];
        reinterpret_cast<float4 *>(out)[i] = make_float4((float) v.x, (float) v.y, (float) v.z, (float) v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[4 * n4 + threadIdx.x] = (float) in[4 * n4 + threadIdx.x];
}
__global__ void k_f32_to_f64(const float *__restrict__ in, double *__restrict__ out, size_t n)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t) gridDim.x * blockDim.x)
        out[i] = (double) in[i];
}

// Seeded pseudo-random content for the kernel-only benchmark (tvl1_bench_iterate): flow in
// [-3,3], dual in [-1,1], image gradients in [-20,20] with about 10% exact zeros (so all three
// branches of the thresholding step occur), rho_c in [-30,30], grad = Ix^2 + Iy^2.
__device__ __forceinline__ float hash_unit(unsigned long long i, unsigned int salt)
{
    unsigned long long z = i * 0x9E3779B97F4A7C15ull + ((unsigned long long) salt << 32 | 0x632BE5ABu);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float) (z >> 40) * (1.0f / 16777216.0f);
}
__global__ void k_fill_bench(float *__restrict__ state, float *__restrict__ consts, size_t field_stride,
                             size_t set_stride)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < field_stride;
         i += (size_t) gridDim.x * blockDim.x) {
        for (int set = 0; set < 2; set++) {
            float *s = state + set * set_stride + i;
            s[F_U1 * field_stride] = 6.f * hash_unit(i, 1) - 3.f;
            s[F_U2 * field_stride] = 6.f * hash_unit(i, 2) - 3.f;
            for (int f = F_P11; f <= F_P22; f++) s[f * field_stride] = 2.f * hash_unit(i, 3 + f) - 1.f;
        }
        const bool flat = hash_unit(i, 11) < 0.1f;
        const float ix = flat ? 0.f : 40.f * hash_unit(i, 12) - 20.f;
        const float iy = flat ? 0.f : 40.f * hash_unit(i, 13) - 20.f;
        consts[C_IX * field_stride + i] = ix;
        consts[C_IY * field_stride + i] = iy;
        consts[C_RHO * field_stride + i] = 60.f * hash_unit(i, 14) - 30.f;
        consts[C_GRAD * field_stride + i] = ix * ix + iy * iy;
    }
}

// ------------------------------------------------------------------------------------------------
// (b) warp + precompute
// ------------------------------------------------------------------------------------------------

// One kernel for src/tvl1flow.cpp:84 and :94-109:
//   centered_gradient(I1)                      src/operators.cpp:335-406
//   bicubic_interpolation_warp(I1 / I1x / I1y) src/bicubic_interpolation.cpp:352-374, :153-245
//   grad = I1wx^2 + I1wy^2,  rho_c = I1w - I1wx*u1 - I1wy*u2 - I0
// The gradient planes are never materialised: centered_gradient is linear, so the warped
// derivatives are taken from a 6x6 (corner-less, 32-tap) neighbourhood of I1 with index-clamped
// outer taps, which is exactly 0.5*(I1[clamp(k+1)] - I1[clamp(k-1)]) pushed through the same
// bicubic weights.  With border_out = true the result is non-zero iff
// 1 <= (int)(j+u1) <= nx-3 and 1 <= (int)(i+u2) <= ny-3 (all three warps share this test);
// (int) truncation of the non-negative coordinate equals j + floor(u1), formed exactly.
// |grad I1w|^2 (src/tvl1flow.cpp:100-104) with a fixed evaluation order, so that every kernel that
// needs it (stored by k_warp, recomputed by the cluster-resident iteration kernel) agrees bitwise.
__device__ __forceinline__ float grad_of(float ix, float iy)
{
    return __fmaf_rn(ix, ix, __fmul_rn(iy, iy));
}

// `fetch(r, c)` returns I1 at row y-2+r, column x-2+c with index clamping already applied.
// Keys (a = -1/2) weights built from the partition of unity and the first moment: 10 operations, and
// sum(w) == 1 up to one rounding.
__device__ __forceinline__ void keys_weights_pu(float t, float w[4])
{
    const float s = 1.0f - t;
    const float h = -0.5f * t * s;          // -t(1-t)/2
    w[0] = h * s;                           // -t(1-t)^2/2
    w[3] = h * t;                           // -t^2(1-t)/2
    w[2] = fmaf(-2.0f, w[3], t + w[0]);     // first moment: -w0 + w2 + 2 w3 = t
    w[1] = ((1.0f - w[0]) - w[2]) - w[3];
}

template <class Fetch>
__device__ __forceinline__ void warp_gather(Fetch fetch, float tx, float ty, float &w, float &wx, float &wy)
{
    float ax[4], ay[4];
    keys_weights_pu(tx, ax);
    keys_weights_pu(ty, ay);
    const float d2 = ax[0] - ax[2], d3 = ax[1] - ax[3];
    float rowI[6];
    float accx = 0.f;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        const float c1 = fetch(r, 1), c2 = fetch(r, 2), c3 = fetch(r, 3), c4 = fetch(r, 4);
        rowI[r] = fmaf(ax[3], c4, fmaf(ax[2], c3, fmaf(ax[1], c2, ax[0] * c1)));
        if (r >= 1 && r <= 4) {
            const float c0 = fetch(r, 0), c5 = fetch(r, 5);
            float dxr = ax[3] * c5;
            dxr = fmaf(ax[2], c4, dxr);
            dxr = fmaf(d3, c3, dxr);
            dxr = fmaf(d2, c2, dxr);
            dxr = fmaf(-ax[1], c1, dxr);
            dxr = fmaf(-ax[0], c0, dxr);
            accx = fmaf(ay[r - 1], dxr, accx);
        }
    }
    w = fmaf(ay[3], rowI[4], fmaf(ay[2], rowI[3], fmaf(ay[1], rowI[2], ay[0] * rowI[1])));
    wx = 0.5f * accx;
    float a = ay[0] * (rowI[2] - rowI[0]);
    a = fmaf(ay[1], rowI[3] - rowI[1], a);
    a = fmaf(ay[2], rowI[4] - rowI[2], a);
    a = fmaf(ay[3], rowI[5] - rowI[3], a);
    wy = 0.5f * a;
}

__device__ __noinline__ void warp_gather_global(const float *__restrict__ img1, int pitch, int nx, int ny,
                                                 int x, int y, float tx, float ty, float *out3)
{
    float w, wx, wy;
    warp_gather([&](int r, int c) {
        return __ldg(img1 + clampi(y - 2 + r, 0, ny - 1) * pitch + clampi(x - 2 + c, 0, nx - 1));
    }, tx, ty, w, wx, wy);
    out3[0] = w; out3[1] = wx; out3[2] = wy;
}

// Shared-memory-staged gather: a CTA owns a 64x16 tile of output pixels.  The box of I1 that covers
// every 6x6 neighbourhood whose integer sample offset is within +-kWarpM pixels of its pixel (tile
// grown by kWarpM+2 / kWarpM+3, x origin rounded down to a float4) is staged with cp.async
// (16-byte copies for interior tiles, index-clamped 4-byte copies on the image border) WHILE the
// flow and I0 of the tile are being loaded: one memory round trip per tile instead of two.
// Pixels whose flow exceeds the margin gather from global memory instead (same arithmetic).
constexpr int kWarpTW = 64, kWarpTH = 16, kWarpM = 8;
constexpr int kWarpBX = 12;                                    // box starts kWarpBX columns left of the tile
constexpr int kWarpBW = 88;                                    // >= kWarpBX + 64 + kWarpM + 3, multiple of 4
constexpr int kWarpBY = kWarpM + 2;
constexpr int kWarpBH = kWarpTH + 2 * kWarpM + 6;              // rows y0-M-2 .. y0+15+M+3

__device__ __forceinline__ void cp_async4(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned int) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned int) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256, 4)
k_warp(const __grid_constant__ CUtensorMap map_i1, const int use_tma,
        const float *__restrict__ I0, const float *__restrict__ I1, size_t img_stride,
        const float *__restrict__ state, size_t plane0, size_t field_stride, size_t set_stride,
        const PairCtl *__restrict__ ctl, float *__restrict__ consts, Level lv, int write_grad,
        int row_begin, int row_end)
{
    __shared__ __align__(128) float s_box[kWarpBH * kWarpBW];
    __shared__ unsigned long long mbar;
    const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
    const int tid = ty * 32 + tx;
    const int b = blockIdx.z;
    const int nx = lv.nx, ny = lv.ny, pitch = lv.pitch;
    const int X0 = blockIdx.x * kWarpTW, Y0 = row_begin + blockIdx.y * kWarpTH;
    const int bx0 = X0 - kWarpBX, by0 = Y0 - kWarpBY;
    const float *img1 = I1 + (size_t) b * img_stride;

    const bool box_inside = bx0 >= 0 && by0 >= 0 && bx0 + kWarpBW <= nx && by0 + kWarpBH <= ny;
    const bool tma = use_tma && box_inside;          // CTA-uniform
    if (tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(smem_u32(&mbar)), "r"((unsigned int) (kWarpBH * kWarpBW * sizeof(float))) : "memory");
            tma_load_3d(s_box, &map_i1, bx0, by0, b, &mbar);
        }
    } else if (box_inside) {
        const float *src = img1 + by0 * pitch + bx0;
        for (int t = tid; t < kWarpBH * (kWarpBW / 4); t += 256) {
            const int ly = t / (kWarpBW / 4), l4 = t - ly * (kWarpBW / 4);
            cp_async16(s_box + ly * kWarpBW + l4 * 4, src + ly * pitch + l4 * 4);
        }
    } else {
        for (int ly = ty; ly < kWarpBH; ly += 8) {
            const float *row = img1 + clampi(by0 + ly, 0, ny - 1) * pitch;
            for (int lx = tx; lx < kWarpBW; lx += 32)
                cp_async4(s_box + ly * kWarpBW + lx, row + clampi(bx0 + lx, 0, nx - 1));
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- flow and I0 of this thread's four pixels: (jj, ii), (jj+32, ii), (jj, ii+8), (jj+32, ii+8) ----
    const int jj = X0 + tx, ii = Y0 + ty;
    const size_t pair_off = (size_t) b * plane0;
    const float *pu1 = state + (size_t) ctl[b].cur * set_stride + pair_off + (size_t) ii * pitch + jj;
    const float *pu2 = pu1 + field_stride;
    const float *pi0 = I0 + (size_t) b * img_stride + (size_t) ii * pitch + jj;
    const bool full = X0 + kWarpTW <= nx && Y0 + kWarpTH <= row_end;      // CTA-uniform: no pixel outside
    const int dn = 8 * pitch;
    float u1[4], u2[4], i0v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int o = 32 * (q & 1) + dn * (q >> 1);
        const bool in = full || (jj + 32 * (q & 1) < nx && ii + 8 * (q >> 1) < row_end);
        u1[q] = in ? __ldg(pu1 + o) : 0.f;
        u2[q] = in ? __ldg(pu2 + o) : 0.f;
        i0v[q] = in ? __ldg(pi0 + o) : 0.f;
    }
    if (tma) {
        __syncthreads();                 // the barrier word is initialised before anybody polls it
        unsigned int done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }

    float *cIx = consts + (size_t) C_IX * field_stride + pair_off + (size_t) ii * pitch + jj;
    const float fjj = (float) jj, fii = (float) ii;          // pixel coordinates in fp32 (exact: far below 2^24)
    const float xmax = (float) (nx - 3), ymax = (float) (ny - 3);
    // One pixel.  INTERIOR (CTA-uniform): the staged box lies inside the image, so a sample whose taps lie inside the
    // box is valid by construction (2 <= sx <= nx-4, 2 <= sy <= ny-4) and the reference's validity test
    // (src/bicubic_interpolation.cpp: 1 <= (int)(j+u1) <= nx-3, ...) is only evaluated for the others.
    auto pixel = [&](auto interior, const int q) {
        constexpr bool INTERIOR = decltype(interior)::value;
        const int j = jj + 32 * (q & 1), i = ii + 8 * (q >> 1);
        if (!full && !(j < nx && i < row_end)) return;
        const float fu = floorf(u1[q]), fv = floorf(u2[q]);
        const float xf = (fjj + (float) (32 * (q & 1))) + fu, yf = (fii + (float) (8 * (q >> 1))) + fv;
        // (int) of a value outside the int range, or of a NaN, is never inside the box
        const int sx = (int) xf, sy = (int) yf;
        const int cx = sx - 2 - bx0, cy = sy - 2 - by0;          // box coordinates of tap (0,0)
        const bool inbox = (unsigned) cx <= (unsigned) (kWarpBW - 6) && (unsigned) cy <= (unsigned) (kWarpBH - 6);
        float w = 0.f, wx = 0.f, wy = 0.f;
        const float ftx = u1[q] - fu, fty = u2[q] - fv;
        if (inbox && (INTERIOR || (xf >= 1.0f && xf <= xmax && yf >= 1.0f && yf <= ymax))) {
            const float *base = s_box + cy * kWarpBW + cx;
            warp_gather([&](int r, int c) { return base[r * kWarpBW + c]; }, ftx, fty, w, wx, wy);
        } else if (!inbox && xf >= 1.0f && xf <= xmax && yf >= 1.0f && yf <= ymax) {
            float o3[3];
            warp_gather_global(img1, pitch, nx, ny, sx, sy, ftx, fty, o3);
            w = o3[0]; wx = o3[1]; wy = o3[2];
        }
        const int o = 32 * (q & 1) + dn * (q >> 1);
        cIx[o] = wx;
        cIx[field_stride + o] = wy;                 // C_IY
        cIx[2 * field_stride + o] = __fmaf_rn(-wy, u2[q], __fmaf_rn(-wx, u1[q], __fsub_rn(w, i0v[q])));   // C_RHO
        if (write_grad) cIx[3 * field_stride + o] = grad_of(wx, wy);
    };
    if (box_inside) {
#pragma unroll
        for (int q = 0; q < 4; q++) pixel(std::true_type{}, q);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) pixel(std::false_type{}, q);
    }
}

// ------------------------------------------------------------------------------------------------
// (c) fused primal-dual iteration, one iteration per launch (T = 1), register marching
// ------------------------------------------------------------------------------------------------
//
// One launch = one pass of the loop body src/tvl1flow.cpp:114-181 for every pair still iterating:
//   TH (:117-143), divergence x2 (src/operators.cpp:35-78), u update + error (:150-162),
//   forward_gradient x2 (src/operators.cpp:86-125), p update (:169-181), stopping rule (:113).
// v, div p and grad u are never materialised.
//
// Mapping: a warp owns a strip of 124 columns x R rows.  Lane l holds 4 consecutive pixels
// (float4); lanes 0..30 are owners, lane 31 only evaluates u_new for the column group to the right
// so that its left neighbour can take the forward x-difference by shuffle.  The warp marches down
// its strip keeping u_new of the current row and the p12/p22 row above in registers, so every
// plane element is loaded once per launch (plus one halo row per strip and one halo lane per warp).
// State is read from set `cur` and written to set `cur^1` (neighbouring strips read each other's
// old values, hence ping-pong).  Per-CTA error partials are reduced in a fixed order by the last
// CTA of each pair (deterministic, fp64), which also applies the stopping rule on the device.

// Row-band mode over peer memory (NVLink): every rank owns one mailbox in its own HBM that all ranks
// of the box map through CUDA IPC.  slot[e & 1][r] carries rank r's partial error sum of global
// iteration (epoch) e; `epoch` counts the exchanges this rank has taken part in.
constexpr int kMaxRanks = 8;
struct BandMailbox {
    double sum[2][kMaxRanks][kTbT];    // up to kTbT sums per exchange (one per iteration of a block)
    unsigned long long tag[2][kMaxRanks];
    unsigned long long epoch;
    int timed_out;
    int pad;
};

struct BandPeers {
    int enabled;                       // 0: not in peer-memory band mode
    int rank, world;
    int halo;                          // rows a band keeps current beyond each of its edges (kTbT, or 1)
    float *up_state;                   // state base of rank-1 (same layout as ours), or null
    float *dn_state;                   // state base of rank+1, or null
    BandMailbox *box[kMaxRanks];       // box[r] = rank r's mailbox (box[rank] is local memory)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All-to-all of `ns` per-rank partial sums through the mailboxes; called by ONE thread per rank and
// exchange.  vals[t] goes in, the sum over ranks (fixed rank order: same bits everywhere) comes back.
// This is the one synchronisation point of a band iteration / block: a rank publishes only after all
// its CTAs (and their halo pushes, fenced at system scope) are done, and nobody leaves before it holds
// every rank's sums, so the next launch finds its halo rows up to date.
__device__ __forceinline__ void band_all_to_all(const BandPeers &pp, double *vals, int ns)
{
    BandMailbox *mine = pp.box[pp.rank];
    const unsigned long long e = mine->epoch + 1;
    const int par = (int) (e & 1);
    for (int r = 0; r < pp.world; r++) {
        BandMailbox *bx = pp.box[r];
        for (int t = 0; t < ns; t++) bx->sum[par][pp.rank][t] = vals[t];
        st_release_sys(&bx->tag[par][pp.rank], e);
    }
    for (int t = 0; t < ns; t++) vals[t] = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < pp.world; r++) {
        while (ld_acquire_sys(&mine->tag[par][r]) != e) {
            // ~4 s in total, once: never hang the GPU on a rank that fell out of step
            if (mine->timed_out || clock64() - t0 > (1ll << 33)) { mine->timed_out = 1; break; }
        }
        for (int t = 0; t < ns; t++) vals[t] += *((volatile double *) &mine->sum[par][r][t]);
    }
    mine->epoch = e;
}

struct IterParams {
    float *state;
    const float *consts;
    PairCtl *ctl;
    double *partials;            // [B][parts_per_pair]
    LoopCtl *loop;
    cudaGraphConditionalHandle cond;   // while-node handle when launched from the solve graph, else 0
    int use_cond;
    // two-phase loop of big lock-step batches: the first while node (wide grid) ends as soon as no more
    // than bulk_min pairs still iterate, the second (narrow grid) when none does
    cudaGraphConditionalHandle cond_bulk;
    int bulk_min;                      // -1: single-phase loop
    int take_all;                      // this launch serves every active pair whatever its block length (a level's first
                                       // block through k_iterate_t2)
    int p_zero;                        // first iteration of a level: the duals are zero (src/tvl1flow.cpp:87-90) and
                                       // are neither read nor were they written by a zeroing pass (k_iterate_t1 only)
    double *tb_partials;               // [batch][tb_parts][kTbT] per-CTA error sums of k_iterate_tb
    int tb_parts;
    int batch;                         // pairs in the lock-step batch (plane index = field * batch + pair)
    int tb;                            // 1: pairs whose next block has more than one iteration belong
                                       // to k_iterate_tb; this kernel only takes the nsteps == 1 pairs.
                                       // 2: ... to k_iterate_t2 (two iterations per launch in registers)
    int tb_max;                        // longest block: kTbT (k_iterate_tb) or kT2T (k_iterate_t2)
    int row_begin, row_end;            // rows this launch owns (whole image: 0, ny; a row band otherwise)
    double *band_sum;                  // row-band mode: the rank's raw sum of squared updates goes here
                                       // and k_band_decide applies the stopping rule after the all-reduce
    BandPeers peers;                   // row-band mode over peer memory: halos and error travel inside the kernel
    int *stat_iters;             // [B][stat_stride]
    double *stat_errs;
    unsigned long long *px_iters;  // [level] pixel-iterations; [TVL1_MAX_LEVELS + level] launches
    size_t plane0, field_stride, set_stride;
    Level lv;
    int parts_per_pair;
    int stat_stride, stat_slot;
    int max_iter;
    int level;                   // pyramid level (statistics only)
    float l_t, theta, taut;
    double eps2;
};

// Stopping rule for a block of `ns` iterations of pair b that started at state set `cur`
// (src/tvl1flow.cpp:113: stop after the first iteration whose mean squared update is <= eps^2, or at
// the cap).  errs[t] = mean squared update of the block's iteration t.  With ns == 1 this is the plain
// rule.  With ns > 1 (temporal blocking) the block may overshoot the stopping point; then nothing
// is accepted -- the ping-pong set that still holds the block's start state stays current -- and the
// pair is marked for an exact replay of t+1 iterations, after which it stops.  Otherwise the number
// of iterations of the next block is predicted from the geometric decay of the error.
__device__ __forceinline__ void decide_block(const IterParams &P, PairCtl *ctl, int b, int cur, int ns,
                                             const double *errs, unsigned long long own_pixels)
{
    const int n0 = ctl->n;
    int stop_at = -1;
    if (ctl->replay) {
        stop_at = ns - 1;                                   // the replay was sized to stop exactly here
    } else {
        for (int t = 0; t < ns; t++)
            if (!(errs[t] > P.eps2 && n0 + t + 1 < P.max_iter)) { stop_at = t; break; }
    }
    ctl->arrive = 0u;
    if (stop_at >= 0 && stop_at < ns - 1) {                 // overshoot: replay exactly stop_at+1 iterations
        ctl->nsteps = stop_at + 1;
        ctl->replay = 1;
        ctl->pzero = P.p_zero;                              // ... from zero duals if that is what the block started from
        return;
    }
    ctl->pzero = 0;
    const double last = errs[ns - 1];
    const double prev = ns >= 2 ? errs[ns - 2] : ctl->err;
    const int n = n0 + ns;
    ctl->n = n;
    ctl->err = last;
    ctl->cur = cur ^ 1;
    ctl->replay = 0;
    atomicAdd(P.px_iters + P.level, (unsigned long long) ns * own_pixels);
    if (stop_at == ns - 1) {
        ctl->active = 0;
        ctl->nsteps = 1;
        P.stat_iters[(size_t) b * P.stat_stride + P.stat_slot] = n;
        P.stat_errs[(size_t) b * P.stat_stride + P.stat_slot] = last;
        atomicAdd(P.px_iters + 2 * kStatLevels + P.level, (unsigned long long) min(n, kLoopClip));
        atomicMax(&P.loop->max_n, n);
        const int left = atomicSub(&P.loop->active_pairs, 1) - 1;
        // the last pair to stop ends the device-side while loop of the solve graph
        if (P.use_cond) {
            if (left == P.bulk_min) cudaGraphSetConditional(P.cond_bulk, 0);
            if (left == 0) cudaGraphSetConditional(P.cond, 0);
        }
        return;
    }
    int next = 1;
    if (P.tb && prev < INFINITY) {          // no history yet (first iteration of the warp step): stay at 1
        next = P.tb_max;                    // error not falling: far from the stopping point
        if (last < prev && last > P.eps2) {
            // error ~ last * r^k: iterations until it is below eps^2 (the decay usually slows down, so
            // this under-estimates and a replay stays rare)
            const double k = log(P.eps2 / last) / log(last / prev);
            if (P.tb == 2) {
                // two-iteration blocks in registers: a block saves 54 of 120 B per pixel when the pair has two more
                // iterations to go and costs 66 B when it has one, so it pays from an even chance on
                if (k < 2.0) next = 1;
            } else if (k < (double) (kTbT + 1)) next = max(1, (int) k - 1);     // keep one in hand
        }
        next = max(1, min(next, P.max_iter - n));
    }
    ctl->nsteps = next;
}

struct Row4 {                    // one image row segment of 4 pixels, everything the update needs
    float4 u1, u2, ix, iy, rho, p11, p12, p21, p22;      // |grad|^2 is recomputed from ix, iy (grad_of)
};

// Hardware approximations (MUFU.RCP / MUFU.RSQ based, <= 1-2 ulp): the IEEE-rounded division and
// square root cost ~10 instructions plus a slow-path branch each and made this HBM-bound kernel
// issue-bound; their rounding difference is far below the fp32-vs-fp64 gap to the reference.
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The per-pixel arithmetic below is written with explicit rounding intrinsics (never re-associated
// or contracted differently by the compiler), so the streaming, temporally blocked and cluster-resident
// kernels produce the same bits for the same input -- which kernel serves a level depends on the batch
// size, and batching must not change the result.
__device__ __forceinline__ float th_coeff(float rho, float grad, float l_t)
{
    // src/tvl1flow.cpp:123-139: d = c * (I1wx, I1wy)
    const float thr = __fmul_rn(l_t, grad);
    const float q = __fmul_rn(-rho, fast_rcp(grad));    // unused (and possibly inf/nan) when grad ~ 0
    return (rho < -thr) ? l_t : ((rho > thr) ? -l_t : ((grad < kGradIsZero) ? 0.f : q));
}

// u_new of one pixel: thresholding step (:117-143), divergence of the dual variables
// (src/operators.cpp:35-78; the caller passes 0 for the terms the reference drops or reads as 0)
// and the primal update (:156-157).
__device__ __forceinline__ void primal_px(float u1, float u2, float ix, float iy, float rho_c, float grad,
                                          float p11c, float p11l, float p12c, float p12a, float p21c,
                                          float p21l, float p22c, float p22a, float l_t, float theta,
                                          float &o1, float &o2)
{
    const float rho = __fmaf_rn(ix, u1, __fmaf_rn(iy, u2, rho_c));      // two roundings
    const float c = th_coeff(rho, grad, l_t);
    const float v1 = __fmaf_rn(c, ix, u1), v2 = __fmaf_rn(c, iy, u2);
    const float d1 = __fadd_rn(__fsub_rn(p11c, p11l), __fsub_rn(p12c, p12a));
    const float d2 = __fadd_rn(__fsub_rn(p21c, p21l), __fsub_rn(p22c, p22a));
    o1 = __fmaf_rn(theta, d1, v1);
    o2 = __fmaf_rn(theta, d2, v2);
}

// squared update of one pixel, the summand of the stopping test (:159-160)
__device__ __forceinline__ float update_sq(float o1, float u1, float o2, float u2)
{
    const float e1 = __fsub_rn(o1, u1), e2 = __fsub_rn(o2, u2);
    return __fmaf_rn(e1, e1, __fmul_rn(e2, e2));
}

// dual update of one pixel from the forward differences of u_new (:169-181)
__device__ __forceinline__ void dual_px(float u1x, float u1y, float u2x, float u2y, float taut,
                                        float &p11, float &p12, float &p21, float &p22)
{
    const float g1 = fast_sqrt(__fmaf_rn(u1x, u1x, __fmul_rn(u1y, u1y)));
    const float g2 = fast_sqrt(__fmaf_rn(u2x, u2x, __fmul_rn(u2y, u2y)));
    const float i1 = fast_rcp(__fmaf_rn(taut, g1, 1.0f));
    const float i2 = fast_rcp(__fmaf_rn(taut, g2, 1.0f));
    p11 = __fmul_rn(__fmaf_rn(taut, u1x, p11), i1);
    p12 = __fmul_rn(__fmaf_rn(taut, u1y, p12), i1);
    p21 = __fmul_rn(__fmaf_rn(taut, u2x, p21), i2);
    p22 = __fmul_rn(__fmaf_rn(taut, u2y, p22), i2);
}

#define TVL1_F4_GET(v, k) ((k) == 0 ? (v).x : (k) == 1 ? (v).y : (k) == 2 ? (v).z : (v).w)

// Pair slots.  A lock-step batch keeps launching until its slowest pair has converged, so most launches
// of a big batch find most pairs already inactive -- and a grid with one z-slice per pair would spend
// the launch starting CTAs that exit at once (57 us for 256 x 272 empty CTAs, measured).  Instead
// grid.z = Z <= B pair SLOTS: the CTA in slot z serves pairs z, z+Z, z+2Z, ...  Every lane tests one of
// them, the ballot gives the CTA's work list, and the CTA walks the set bits.  `tb_kernel` selects
// which of the two iteration kernels the pair's next block belongs to.
// The work list lives in shared memory, not in registers: nothing of the loop stays live across
// the (register-hungry) body.  The host sizes Z so that 32 * Z >= batch: one ballot covers the slot.
// ROUNDS = false: the host sizes Z so that 32 * Z >= batch, one ballot covers the slot (the streaming
// kernel: a loop around its body costs it 50 registers).  ROUNDS = true: any Z; a narrow launch with
// fewer than batch / 32 slots goes round again.
template <bool ROUNDS, class Body>
__device__ __forceinline__ void for_each_pair_of_slot(const IterParams &P, bool tb_kernel, Body &&body)
{
    __shared__ unsigned int s_todo;
    const int per_round = 32 * (int) gridDim.z;
#pragma unroll 1
    for (int base = 0; base < (ROUNDS ? P.batch : 1); base += per_round) {
        const int mine = base + blockIdx.z + (threadIdx.x & 31) * gridDim.z;
        bool take = false;
        if (mine < P.batch) {
            // a pair's control word changes only after ALL its CTAs, this one included, have arrived:
            // every warp of the CTA reads the same values here
            const PairCtl *c = P.ctl + mine;
            const bool blocked = P.tb && c->nsteps > 1;
            take = c->active && (P.take_all || blocked == tb_kernel);
        }
        const unsigned int todo0 = __ballot_sync(0xffffffffu, take);
        if (todo0 == 0u) continue;                   // the common case of a late launch (CTA-uniform)
        if (threadIdx.x == 0) s_todo = todo0;
        __syncthreads();
        for (;;) {
            const unsigned int todo = *(volatile unsigned int *) &s_todo;
            if (todo == 0u) break;
            body(base + (int) blockIdx.z + (__ffs(todo) - 1) * (int) gridDim.z);
            __syncthreads();                         // the pair's shared-memory scratch is free again
            if (threadIdx.x == 0) { const unsigned int t = s_todo; s_todo = t & (t - 1u); }
            __syncthreads();
        }
        __syncthreads();                             // everybody has seen the empty list before it is refilled
    }
}

template <int R, int WY>
__device__ __forceinline__ void iterate_t1_pair(const IterParams &P, const int b)
{
    PairCtl *ctl = P.ctl + b;

    const int nx = P.lv.nx, ny = P.lv.ny, pitch = P.lv.pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cur = ctl->cur;
    const bool p_zero = P.p_zero || ctl->pzero;     // CTA-uniform: the duals of the start state read as zero
    const float *sin = P.state + (size_t) cur * P.set_stride + (size_t) b * P.plane0;
    float *sout = P.state + (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0;
    const float *cst = P.consts + (size_t) b * P.plane0;
    const size_t fs = P.field_stride;

    const int x0 = blockIdx.x * 124 + lane * 4;
    const int ys = P.row_begin + (blockIdx.y * WY + warp) * R;
    const int ye = min(ys + R, P.row_end);
    const bool in_alloc = x0 < pitch;               // float4 lies inside the row allocation
    const bool owner = in_alloc && lane < 31 && x0 < nx;

    float err = 0.f;

    if (ys < P.row_end) {                           // warp-uniform
        auto load_row = [&](int y, Row4 &r) {
            if (in_alloc) {
                const size_t o = (size_t) y * pitch + x0;
                r.u1 = ldg4(sin + F_U1 * fs + o);
                r.u2 = ldg4(sin + F_U2 * fs + o);
                if (p_zero) {                        // CTA-uniform
                    r.p11 = r.p12 = r.p21 = r.p22 = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    r.p11 = ldg4(sin + F_P11 * fs + o);
                    r.p12 = ldg4(sin + F_P12 * fs + o);
                    r.p21 = ldg4(sin + F_P21 * fs + o);
                    r.p22 = ldg4(sin + F_P22 * fs + o);
                }
                r.ix = ldg4(cst + C_IX * fs + o);
                r.iy = ldg4(cst + C_IY * fs + o);
                r.rho = ldg4(cst + C_RHO * fs + o);
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                r.u1 = r.u2 = r.p11 = r.p12 = r.p21 = r.p22 = r.ix = r.iy = r.rho = z;
            }
        };
        // u_new of one row; a12/a22 = p12/p22 of the row above (0 on the first image row)
        auto primal = [&](int y, const Row4 &r, const float4 &a12, const float4 &a22, float4 &n1,
                          float4 &n2, bool count) {
            // left neighbours of p11/p21: previous lane's .w, or a scalar load for lane 0
            float l11 = __shfl_up_sync(0xffffffffu, r.p11.w, 1);
            float l21 = __shfl_up_sync(0xffffffffu, r.p21.w, 1);
            if (lane == 0) {
                const bool has = x0 > 0 && !p_zero;
                const size_t o = (size_t) y * pitch + x0 - 1;
                l11 = has ? __ldg(sin + F_P11 * fs + o) : 0.f;
                l21 = has ? __ldg(sin + F_P21 * fs + o) : 0.f;
            }
            const bool last_row = (y == ny - 1);
            const bool tally = count && owner;
            float o1[4], o2[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float u1 = TVL1_F4_GET(r.u1, k), u2 = TVL1_F4_GET(r.u2, k);
                // divergence, src/operators.cpp:35-78: "+p1" dropped on the last column,
                // "+p2" on the last row, p[-1] = 0
                const bool last_col = (x0 + k >= nx - 1);
                primal_px(u1, u2, TVL1_F4_GET(r.ix, k), TVL1_F4_GET(r.iy, k), TVL1_F4_GET(r.rho, k),
                          grad_of(TVL1_F4_GET(r.ix, k), TVL1_F4_GET(r.iy, k)),
                          last_col ? 0.f : TVL1_F4_GET(r.p11, k), (k == 0) ? l11 : TVL1_F4_GET(r.p11, (k + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(r.p12, k), TVL1_F4_GET(a12, k),
                          last_col ? 0.f : TVL1_F4_GET(r.p21, k), (k == 0) ? l21 : TVL1_F4_GET(r.p21, (k + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(r.p22, k), TVL1_F4_GET(a22, k),
                          P.l_t, P.theta, o1[k], o2[k]);
                const float sq = update_sq(o1[k], u1, o2[k], u2);
                err = __fadd_rn(err, (tally && x0 + k < nx) ? sq : 0.f);
            }
            n1 = make_float4(o1[0], o1[1], o1[2], o1[3]);
            n2 = make_float4(o2[0], o2[1], o2[2], o2[3]);
        };
        // One step of the march: `c` (row y) is complete, `d` receives row y+1; then the forward
        // gradient of u_new (src/operators.cpp:86-125), the dual update (:169-181) and the stores
        // of row y.  Called with the roles of the two register sets alternating, so nothing is copied.
        auto step = [&](int y, Row4 &c, float4 &uc1, float4 &uc2, Row4 &d, float4 &ud1, float4 &ud2) {
            const bool has_below = (y + 1 < ny);
            if (has_below) {
                load_row(y + 1, d);
                primal(y + 1, d, c.p12, c.p22, ud1, ud2, y + 1 < ye);
            }
            const float r1 = __shfl_down_sync(0xffffffffu, uc1.x, 1);
            const float r2 = __shfl_down_sync(0xffffffffu, uc2.x, 1);
            float q11[4], q12[4], q21[4], q22[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool last_col = (x0 + k >= nx - 1);
                const float c1 = TVL1_F4_GET(uc1, k), c2 = TVL1_F4_GET(uc2, k);
                const float e1 = (k == 3) ? r1 : TVL1_F4_GET(uc1, (k + 1) & 3);
                const float e2 = (k == 3) ? r2 : TVL1_F4_GET(uc2, (k + 1) & 3);
                q11[k] = TVL1_F4_GET(c.p11, k); q12[k] = TVL1_F4_GET(c.p12, k);
                q21[k] = TVL1_F4_GET(c.p21, k); q22[k] = TVL1_F4_GET(c.p22, k);
                dual_px(last_col ? 0.f : e1 - c1, has_below ? TVL1_F4_GET(ud1, k) - c1 : 0.f,
                        last_col ? 0.f : e2 - c2, has_below ? TVL1_F4_GET(ud2, k) - c2 : 0.f,
                        P.taut, q11[k], q12[k], q21[k], q22[k]);
            }
            if (owner) {
                const size_t o = (size_t) y * pitch + x0;
                st4(sout + F_U1 * fs + o, uc1);
                st4(sout + F_U2 * fs + o, uc2);
                st4(sout + F_P11 * fs + o, make_float4(q11[0], q11[1], q11[2], q11[3]));
                st4(sout + F_P12 * fs + o, make_float4(q12[0], q12[1], q12[2], q12[3]));
                st4(sout + F_P21 * fs + o, make_float4(q21[0], q21[1], q21[2], q21[3]));
                st4(sout + F_P22 * fs + o, make_float4(q22[0], q22[1], q22[2], q22[3]));
                if (P.peers.enabled) {
                    // my first `halo` rows are the halo below of the band above, my last `halo` rows the
                    // halo above of the band below: same offsets in the neighbour's planes (NVLink stores)
                    const size_t po = (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0 + o;
                    float *d = nullptr;
                    if (y < P.row_begin + P.peers.halo) d = P.peers.up_state;
                    else if (y >= P.row_end - P.peers.halo) d = P.peers.dn_state;
                    // (a band shorter than 2*halo rows sends its overlap both ways)
                    float *d2 = (y < P.row_begin + P.peers.halo && y >= P.row_end - P.peers.halo) ? P.peers.dn_state : nullptr;
#pragma unroll 1
                    for (int k2 = 0; k2 < 2; k2++) {
                        float *dd = k2 ? d2 : d;
                        if (!dd) continue;
                        dd += po;
                        st4(dd + F_U1 * fs, uc1);
                        st4(dd + F_U2 * fs, uc2);
                        st4(dd + F_P11 * fs, make_float4(q11[0], q11[1], q11[2], q11[3]));
                        st4(dd + F_P12 * fs, make_float4(q12[0], q12[1], q12[2], q12[3]));
                        st4(dd + F_P21 * fs, make_float4(q21[0], q21[1], q21[2], q21[3]));
                        st4(dd + F_P22 * fs, make_float4(q22[0], q22[1], q22[2], q22[3]));
                    }
                }
            }
        };

        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        Row4 ra, rb;
        float4 a12 = zero4, a22 = zero4;
        if (ys > 0 && in_alloc && !p_zero) {
            const size_t o = (size_t) (ys - 1) * pitch + x0;
            a12 = ldg4(sin + F_P12 * fs + o);
            a22 = ldg4(sin + F_P22 * fs + o);
        }
        load_row(ys, ra);
        rb = ra;
        float4 ua1, ua2, ub1 = zero4, ub2 = zero4;
        primal(ys, ra, a12, a22, ua1, ua2, true);
        for (int y = ys; y < ye; y += 2) {
            step(y, ra, ua1, ua2, rb, ub1, ub2);
            if (y + 1 < ye) step(y + 1, rb, ub1, ub2, ra, ua1, ua2);
        }
    }

    // ---- error reduction: warp shuffle -> CTA -> fixed-order sum by the pair's last CTA ----
    __shared__ double s_part[32];
    __shared__ int s_last;
    double e = warp_sum((double) err);
    if (lane == 0) s_part[warp] = e;
    __syncthreads();
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < WY; w++) s += s_part[w];
        P.partials[(size_t) b * P.parts_per_pair + blk] = s;
        if (P.peers.enabled) __threadfence_system();    // halo rows pushed to the neighbours are out
        else __threadfence();
        const unsigned int t = atomicAdd(&ctl->arrive, 1u);
        s_last = (t == (unsigned int) nblk - 1u);
    }
    __syncthreads();
    if (!s_last) return;

    __threadfence();
    const volatile double *part = P.partials + (size_t) b * P.parts_per_pair;
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32 * WY) s += part[i];
    s = warp_sum(s);
    if (lane == 0) s_part[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < WY; w++) tot += s_part[w];
        if (P.band_sum) {            // row-band mode: the other ranks' rows are still missing
            *P.band_sum = tot;
            ctl->arrive = 0u;
            return;
        }
        if (P.peers.enabled) band_all_to_all(P.peers, &tot, 1);
        const double error = tot / ((double) nx * (double) ny);   // src/tvl1flow.cpp:162
        decide_block(P, ctl, b, cur, 1, &error,
                     (unsigned long long) nx * (unsigned long long) (P.row_end - P.row_begin));
    }
}

template <int R, int WY>
__global__ void __launch_bounds__(32 * WY)
k_iterate_t1(const IterParams P)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        atomicAdd(P.px_iters + kStatLevels + P.level, 1ull);
    for_each_pair_of_slot<false>(P, false, [&](int b) { iterate_t1_pair<R, WY>(P, b); });
}

// (c') TWO iterations per launch, in registers: the marching kernel above with a second iteration following
// the first one row behind.  For the lock-step batches whose streamed levels saturate HBM (where the
// shared-memory kernel k_iterate_tb loses: it trades bytes for instruction issue, 243 instructions per useful
// pixel-iteration against 110 here), this halves the bytes instead: every plane element is loaded once and
// stored once per TWO iterations -- (36 B in + 24 B out) x halo overhead = 33 B per pixel-iteration against 60.
// A warp owns 120 columns x R rows.  Lane l holds the 4 pixels at x = 120*bx - 4 + 4*l: lanes 1..30 own, lanes 0
// and 31 are the one-pixel-per-iteration halo either side (a float4 each).  Rows: the first iteration runs on
// rows ys-1 .. ye+1, the second on ys .. ye, stores on ys .. ye-1 (R + 3 rows loaded for R stored; the halo rows
// are the neighbouring strips' own rows, i.e. L2 hits).  Three register blocks (rows t, t+1, t+2) rotate:
//   A(t+2)  load row t+2 of the start state;  u' = first-iteration u_new  (needs p(t+1) of the start state)
//   B(t+1)  p' (t+1) from u'(t+1), u'(t+2), in place
//   C(t+1)  u''(t+1) from u'(t+1), p'(t+1), p'(t), in place
//   D(t)    p''(t) from p'(t), u''(t), u''(t+1); store row t.
// Same per-pixel functions as the other iteration kernels, hence the same bits; the error of each of the two
// iterations is reduced separately and decide_block accepts the block, or has its first iteration replayed.
constexpr int kT2W = 120;                 // owned columns per warp
constexpr int kT2T = 2;                   // iterations per launch
// STAGE: the rows of the start state reach the registers through a per-warp ring in shared memory filled by cp.async
// (kT2PF rows ahead, every lane only ever touches its own 16-byte slots: no barrier) instead of by loads that hold
// their destination registers for the whole HBM latency -- at 168 registers and 12 warps per SM the kernel was
// bound by that latency (DRAM at 74 % of peak, profiles/r3m_iterate_t2_full.csv), not by bandwidth.
constexpr int kT2PF = 2;                  // rows in flight ahead of the one being consumed
constexpr int kT2NS = kT2PF + 1;          // ring slots per warp
constexpr int kT2RowF4 = 9 * 32;          // float4 per ring slot: nine planes x 32 lanes
__host__ __device__ constexpr size_t t2_smem_bytes(int WY) { return (size_t) WY * kT2NS * kT2RowF4 * 16; }

template <int R, int WY, bool STAGE>
__device__ __forceinline__ void iterate_t2_pair(const IterParams &P, const int b, float4 *t2_ring)
{
    PairCtl *ctl = P.ctl + b;
    const int nx = P.lv.nx, ny = P.lv.ny, pitch = P.lv.pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cur = ctl->cur;
    const float *sin = P.state + (size_t) cur * P.set_stride + (size_t) b * P.plane0;
    float *sout = P.state + (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0;
    const float *cst = P.consts + (size_t) b * P.plane0;
    const size_t fs = P.field_stride;

    const int x0 = blockIdx.x * kT2W - 4 + lane * 4;
    const int ys = P.row_begin + (blockIdx.y * WY + warp) * R;
    const int ye = min(ys + R, P.row_end);
    const bool in_alloc = x0 >= 0 && x0 < pitch;
    const bool owner = in_alloc && lane >= 1 && lane <= 30 && x0 < nx;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    float err1 = 0.f, err2 = 0.f;

    if (ys < P.row_end) {                           // warp-uniform
        // ---- staged loads: this lane's 16-byte slots of the warp's ring ----
        float4 *ring = t2_ring + (size_t) warp * kT2NS * kT2RowF4 + lane;
        const int y_last = min(ye + 1, ny - 1);     // last row the march requests
        int seq = 0, y_next = 0;                    // rows consumed so far; next row to put in flight
        auto issue = [&](int y, int slot) {
            if (in_alloc && y <= y_last) {
                const size_t o = (size_t) y * pitch + x0;
                float4 *d = ring + slot * kT2RowF4;
                cp_async16((float *) (d + 0 * 32), sin + F_U1 * fs + o);
                cp_async16((float *) (d + 1 * 32), sin + F_U2 * fs + o);
                if (!P.p_zero) {
                    cp_async16((float *) (d + 2 * 32), sin + F_P11 * fs + o);
                    cp_async16((float *) (d + 3 * 32), sin + F_P12 * fs + o);
                    cp_async16((float *) (d + 4 * 32), sin + F_P21 * fs + o);
                    cp_async16((float *) (d + 5 * 32), sin + F_P22 * fs + o);
                }
                cp_async16((float *) (d + 6 * 32), cst + C_IX * fs + o);
                cp_async16((float *) (d + 7 * 32), cst + C_IY * fs + o);
                cp_async16((float *) (d + 8 * 32), cst + C_RHO * fs + o);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (STAGE) {
            y_next = max(ys - 1, 0);
#pragma unroll
            for (int i = 0; i < kT2PF; i++) issue(y_next++, i);
        }
        auto load_row = [&](int y, Row4 &r) {
            if (STAGE) {
                issue(y_next++, (seq + kT2PF) % kT2NS);
                asm volatile("cp.async.wait_group %0;" :: "n"(kT2PF) : "memory");
                const float4 *d = ring + (seq % kT2NS) * kT2RowF4;
                seq++;
                if (in_alloc) {
                    r.u1 = d[0 * 32]; r.u2 = d[1 * 32];
                    if (P.p_zero) {
                        r.p11 = r.p12 = r.p21 = r.p22 = zero4;
                    } else {
                        r.p11 = d[2 * 32]; r.p12 = d[3 * 32]; r.p21 = d[4 * 32]; r.p22 = d[5 * 32];
                    }
                    r.ix = d[6 * 32]; r.iy = d[7 * 32]; r.rho = d[8 * 32];
                } else {
                    r.u1 = r.u2 = r.p11 = r.p12 = r.p21 = r.p22 = r.ix = r.iy = r.rho = zero4;
                }
            } else if (in_alloc) {
                const size_t o = (size_t) y * pitch + x0;
                r.u1 = ldg4(sin + F_U1 * fs + o);
                r.u2 = ldg4(sin + F_U2 * fs + o);
                if (P.p_zero) {                      // launch-uniform: a level's first block (src/tvl1flow.cpp:87-90)
                    r.p11 = r.p12 = r.p21 = r.p22 = zero4;
                } else {
                    r.p11 = ldg4(sin + F_P11 * fs + o);
                    r.p12 = ldg4(sin + F_P12 * fs + o);
                    r.p21 = ldg4(sin + F_P21 * fs + o);
                    r.p22 = ldg4(sin + F_P22 * fs + o);
                }
                r.ix = ldg4(cst + C_IX * fs + o);
                r.iy = ldg4(cst + C_IY * fs + o);
                r.rho = ldg4(cst + C_RHO * fs + o);
            } else {
                r.u1 = r.u2 = r.p11 = r.p12 = r.p21 = r.p22 = r.ix = r.iy = r.rho = zero4;
            }
        };
        // u_new of row y from the row's u, p and constants, in place; a12 / a22 = p12 / p22 of the row above
        auto primal = [&](int y, Row4 &r, const float4 &a12, const float4 &a22, float &err, bool count) {
            // left neighbours of p11 / p21: the previous lane's .w (lane 0 is halo: its .x is never used)
            const float l11 = __shfl_up_sync(0xffffffffu, r.p11.w, 1);
            const float l21 = __shfl_up_sync(0xffffffffu, r.p21.w, 1);
            const bool last_row = (y == ny - 1);
            const bool tally = count && owner;
            float o1[4], o2[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float u1 = TVL1_F4_GET(r.u1, k), u2 = TVL1_F4_GET(r.u2, k);
                const bool last_col = (x0 + k >= nx - 1);
                primal_px(u1, u2, TVL1_F4_GET(r.ix, k), TVL1_F4_GET(r.iy, k), TVL1_F4_GET(r.rho, k),
                          grad_of(TVL1_F4_GET(r.ix, k), TVL1_F4_GET(r.iy, k)),
                          last_col ? 0.f : TVL1_F4_GET(r.p11, k), (k == 0) ? l11 : TVL1_F4_GET(r.p11, (k + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(r.p12, k), TVL1_F4_GET(a12, k),
                          last_col ? 0.f : TVL1_F4_GET(r.p21, k), (k == 0) ? l21 : TVL1_F4_GET(r.p21, (k + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(r.p22, k), TVL1_F4_GET(a22, k),
                          P.l_t, P.theta, o1[k], o2[k]);
                const float sq = update_sq(o1[k], u1, o2[k], u2);
                err = __fadd_rn(err, (tally && x0 + k < nx) ? sq : 0.f);
            }
            r.u1 = make_float4(o1[0], o1[1], o1[2], o1[3]);
            r.u2 = make_float4(o2[0], o2[1], o2[2], o2[3]);
        };
        // dual update of row y from the u_new of rows y (in c) and y+1 (in d): q = new p of row y
        auto dual = [&](int y, const Row4 &c, const Row4 &d, float4 &q11, float4 &q12, float4 &q21, float4 &q22) {
            const bool has_below = (y + 1 < ny);
            const float r1 = __shfl_down_sync(0xffffffffu, c.u1.x, 1);
            const float r2 = __shfl_down_sync(0xffffffffu, c.u2.x, 1);
            float a11[4], a12[4], a21[4], a22[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool last_col = (x0 + k >= nx - 1);
                const float c1 = TVL1_F4_GET(c.u1, k), c2 = TVL1_F4_GET(c.u2, k);
                const float e1 = (k == 3) ? r1 : TVL1_F4_GET(c.u1, (k + 1) & 3);
                const float e2 = (k == 3) ? r2 : TVL1_F4_GET(c.u2, (k + 1) & 3);
                a11[k] = TVL1_F4_GET(c.p11, k); a12[k] = TVL1_F4_GET(c.p12, k);
                a21[k] = TVL1_F4_GET(c.p21, k); a22[k] = TVL1_F4_GET(c.p22, k);
                dual_px(last_col ? 0.f : e1 - c1, has_below ? TVL1_F4_GET(d.u1, k) - c1 : 0.f,
                        last_col ? 0.f : e2 - c2, has_below ? TVL1_F4_GET(d.u2, k) - c2 : 0.f,
                        P.taut, a11[k], a12[k], a21[k], a22[k]);
            }
            q11 = make_float4(a11[0], a11[1], a11[2], a11[3]);
            q12 = make_float4(a12[0], a12[1], a12[2], a12[3]);
            q21 = make_float4(a21[0], a21[1], a21[2], a21[3]);
            q22 = make_float4(a22[0], a22[1], a22[2], a22[3]);
        };
        // A: row y of the start state arrives in r; first-iteration u_new in place (`above` = the row above, start state)
        auto stage_a = [&](int y, Row4 &r, const float4 &a12, const float4 &a22) {
            load_row(y, r);
            primal(y, r, a12, a22, err1, y >= ys && y < ye);
        };
        // B: first-iteration p of row y in place.  Left of the image (the halo lane of the first strip column) the
        // duals are the zeros the divergence reads there (src/operators.cpp:35-78), not an update of padding.
        auto stage_b = [&](int y, Row4 &c, const Row4 &d) {
            float4 q11, q12, q21, q22;
            dual(y, c, d, q11, q12, q21, q22);
            const bool outside = x0 < 0;                       // whole float4 left of column 0 (lane 0 of strip column 0)
            c.p11 = outside ? zero4 : q11; c.p12 = outside ? zero4 : q12;
            c.p21 = outside ? zero4 : q21; c.p22 = outside ? zero4 : q22;
        };
        // D: second-iteration p of row y and the stores of the row
        auto stage_d = [&](int y, const Row4 &c, const Row4 &d) {
            float4 q11, q12, q21, q22;
            dual(y, c, d, q11, q12, q21, q22);
            if (owner) {
                const size_t o = (size_t) y * pitch + x0;
                st4(sout + F_U1 * fs + o, c.u1);
                st4(sout + F_U2 * fs + o, c.u2);
                st4(sout + F_P11 * fs + o, q11);
                st4(sout + F_P12 * fs + o, q12);
                st4(sout + F_P21 * fs + o, q21);
                st4(sout + F_P22 * fs + o, q22);
            }
        };
        // one step of the march: output row t.  On entry `c` = row t (p', u''), `d` = row t+1 (start p, u', constants);
        // `e` receives row t+2.
        auto step = [&](int t, Row4 &c, Row4 &d, Row4 &e) {
            if (t + 2 < ny) stage_a(t + 2, e, d.p12, d.p22);
            if (t + 1 < ny) {
                stage_b(t + 1, d, e);
                primal(t + 1, d, c.p12, c.p22, err2, t + 1 < ye);       // C
            }
            stage_d(t, c, d);
        };

        Row4 r0, r1, r2;
        // prologue: rows ys-1, ys, ys+1 of the first iteration, p' of ys-1 and ys, u'' of ys
        float4 a12 = zero4, a22 = zero4;
        if (ys > 0) {
            if (ys > 1 && in_alloc && !P.p_zero) {
                const size_t o = (size_t) (ys - 2) * pitch + x0;
                a12 = ldg4(sin + F_P12 * fs + o);
                a22 = ldg4(sin + F_P22 * fs + o);
            }
            stage_a(ys - 1, r2, a12, a22);
            a12 = r2.p12; a22 = r2.p22;
        } else {
            r2.u1 = r2.u2 = r2.p11 = r2.p12 = r2.p21 = r2.p22 = r2.ix = r2.iy = r2.rho = zero4;
        }
        stage_a(ys, r0, a12, a22);
        if (ys + 1 < ny) stage_a(ys + 1, r1, r0.p12, r0.p22);
        else r1 = r0;
        if (ys > 0) stage_b(ys - 1, r2, r0);
        stage_b(ys, r0, r1);
        primal(ys, r0, r2.p12, r2.p22, err2, true);                     // C(ys): above = p'(ys-1), zeros on the first row
        for (int t = ys; t < ye; t += 3) {
            step(t, r0, r1, r2);
            if (t + 1 < ye) step(t + 1, r1, r2, r0);
            if (t + 2 < ye) step(t + 2, r2, r0, r1);
        }
    }

    // ---- per-iteration error sums: warp shuffle -> CTA -> fixed-order sum by the pair's last CTA ----
    __shared__ double s_part[kT2T][32];
    __shared__ double s_tot[kT2T];
    __shared__ int s_last;
    const double e1 = warp_sum((double) err1), e2 = warp_sum((double) err2);
    if (lane == 0) { s_part[0][warp] = e1; s_part[1][warp] = e2; }
    __syncthreads();
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    double *part = P.tb_partials + ((size_t) b * P.tb_parts) * kTbT;
    if (threadIdx.x == 0) {
        for (int t = 0; t < kT2T; t++) {
            double s = 0.0;
            for (int w = 0; w < WY; w++) s += s_part[t][w];
            part[(size_t) blk * kTbT + t] = s;
        }
        __threadfence();
        const unsigned int tk = atomicAdd(&ctl->arrive, 1u);
        s_last = (tk == (unsigned int) nblk - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile double *vp = part;
    for (int t = 0; t < kT2T; t++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblk; i += 32 * WY) s += vp[(size_t) i * kTbT + t];
        s = warp_sum(s);
        __syncthreads();
        if (lane == 0) s_part[0][warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < WY; w++) tot += s_part[0][w];
            s_tot[t] = tot / ((double) nx * (double) ny);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        decide_block(P, ctl, b, cur, kT2T, s_tot, (unsigned long long) nx * (unsigned long long) (P.row_end - P.row_begin));
}

template <int R, int WY, bool STAGE>
__global__ void __launch_bounds__(32 * WY, 3)
k_iterate_t2(const IterParams P)
{
    extern __shared__ __align__(16) float4 t2_ring[];      // [WY][kT2NS][9][32] when STAGE, else nothing
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        atomicAdd(P.px_iters + kStatLevels + P.level, 1ull);
    for_each_pair_of_slot<false>(P, true, [&](int b) { iterate_t2_pair<R, WY, STAGE>(P, b, t2_ring); });
}

// Row-band mode over peer memory: barrier of all ranks through the mailboxes (start of a solve: no rank
// may push halos into a neighbour that is still reading the previous solve's result).
__global__ void k_band_barrier(BandPeers pp)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double none = 0.0;
        band_all_to_all(pp, &none, 0);
    }
}

// Row-band mode over peer memory: all-gather of the flow after a split level.  Every rank stores its
// band of u1, u2 (live set) into the same rows of every other rank's planes (NVLink stores), then the
// ranks meet at a mailbox barrier, so when this kernel has finished on a rank that rank holds the whole
// flow of the level.  No NCCL call, no host round trip.
struct GatherParams {
    float *state[kMaxRanks];           // state base of every rank (state[rank] = ours)
    BandPeers peers;
    const PairCtl *ctl;
    unsigned int *ticket;              // zeroed; left zeroed
    size_t set_stride, field_stride;
    int pitch, row_begin, row_end;
};

__global__ void __launch_bounds__(256)
k_band_allgather(const GatherParams G)
{
    const size_t set_off = (size_t) G.ctl->cur * G.set_stride;
    const size_t first = (size_t) G.row_begin * G.pitch;
    const size_t n4 = (size_t) (G.row_end - G.row_begin) * G.pitch / 4;
    const float *mine = G.state[G.peers.rank] + set_off + first;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n4; i += (size_t) gridDim.x * blockDim.x) {
        const size_t f = i / n4, k = i - f * n4;                     // field F_U1 / F_U2, float4 index
        const size_t o = f * G.field_stride + 4 * k;
        const float4 v = ldg4(mine + o);
        for (int r = 0; r < G.peers.world; r++)
            if (r != G.peers.rank) st4(G.state[r] + set_off + first + o, v);
    }
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(G.ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence_system();
        *G.ticket = 0u;
        double none = 0.0;
        band_all_to_all(G.peers, &none, 0);
    }
}

// ------------------------------------------------------------------------------------------------
// (c') cluster-resident iteration kernel: the whole while loop of one warp step on chip
// ------------------------------------------------------------------------------------------------
//
// For pyramid levels small enough to live on chip, one thread-block CLUSTER owns one frame pair:
// each CTA keeps a band of rows of the six evolving planes (u1,u2,p11,p12,p21,p22) in shared
// memory, the per-warp constants of its pixels in registers, and the cluster runs the complete
// loop of src/tvl1flow.cpp:113-182 without touching HBM between iterations:
//   * state is brought in once with TMA bulk copies (cp.async.bulk + mbarrier) and written back
//     once with TMA bulk stores;
//   * band neighbours exchange their 1-row halos through distributed shared memory (the row of
//     u_new a CTA's upper neighbour needs for the forward y-difference, the row of p12/p22 its
//     lower neighbour needs for the divergence), pushed by the producer;
//   * the mean squared update is reduced per CTA, broadcast to every CTA of the cluster through
//     DSMEM and summed in rank order, so every CTA takes the same exact stop decision after every
//     single iteration (no replay, no host, bit-reproducible).
// Two barriers per iteration (cluster barrier, or __syncthreads for a 1-CTA cluster).
namespace cg = cooperative_groups;

constexpr int kResThreads = 512;
constexpr int kResQuads = 4;                       // float4 pixel groups per thread
constexpr int kResMaxCluster = 16;
constexpr size_t kResSmemLimit = 227 * 1024;

struct ResParams {
    float *state;
    const float *consts;
    PairCtl *ctl;
    int *stat_iters;
    double *stat_errs;
    unsigned long long *counters;    // [level] pixel-iterations; [kStatLevels + level] launches
    double *err_trace;               // optional [max_iter] per-iteration error of pair 0 (tests)
    size_t plane0, field_stride, set_stride;
    Level lv;
    int rows_per_cta;
    int stat_stride, stat_slot;
    int max_iter;
    int level;
    float l_t, theta, taut;
    double eps2;
};

__host__ __device__ inline size_t resident_smem_bytes(int pitch, int rows_per_cta)
{
    // u1,u2: rows+1 (halo below) | p11,p21: rows | p12,p22: rows+1 (halo above) | slots | mbarrier
    return (size_t) pitch * (6 * rows_per_cta + 4) * sizeof(float) + (kResMaxCluster + kResThreads / 32) * sizeof(double) + 16;
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__global__ void __launch_bounds__(kResThreads, 1)
k_iterate_resident(const ResParams P)
{
    extern __shared__ __align__(128) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int) cluster.num_blocks();
    const int rank = (int) cluster.block_rank();
    const int b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = P.lv.nx, ny = P.lv.ny, pitch = P.lv.pitch;
    const int RB = P.rows_per_cta;
    const int r0 = rank * RB;
    const int rows = min(RB, ny - r0);              // >= 1 by construction (host)
    const int qpr = pitch >> 2;
    const int nquads = rows * qpr;

    float *sU1 = smem;
    float *sU2 = sU1 + (size_t) (RB + 1) * pitch;
    float *sP11 = sU2 + (size_t) (RB + 1) * pitch;
    float *sP21 = sP11 + (size_t) RB * pitch;
    float *sP12 = sP21 + (size_t) RB * pitch;       // row 0 = halo above, own rows at 1..rows
    float *sP22 = sP12 + (size_t) (RB + 1) * pitch;
    double *sSlots = reinterpret_cast<double *>(sP22 + (size_t) (RB + 1) * pitch);   // [C] cluster partials
    double *sWarp = sSlots + kResMaxCluster;                                          // [kResThreads / 32] warp partials
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(sWarp + kResThreads / 32);

    PairCtl *ctl = P.ctl + b;
    const int cur = ctl->cur;
    const size_t fs = P.field_stride;
    const float *gin = P.state + (size_t) cur * P.set_stride + (size_t) b * P.plane0;
    float *gout = P.state + (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0;
    const float *cst = P.consts + (size_t) b * P.plane0;

    // ---- bring the band in: TMA bulk copies signalled on one mbarrier -------------------------
    const unsigned int band_bytes = (unsigned int) ((size_t) rows * pitch * sizeof(float));
    const unsigned int row_bytes = (unsigned int) (pitch * sizeof(float));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned int total = 6u * band_bytes + (r0 > 0 ? 2u * row_bytes : 0u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(total) : "memory");
        const size_t go = (size_t) r0 * pitch;
        float *dsts[6] = { sU1, sU2, sP11, sP12 + pitch, sP21, sP22 + pitch };   // order of enum Field
#pragma unroll
        for (int f = 0; f < 6; f++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(dsts[f])), "l"(gin + (size_t) f * fs + go), "r"(band_bytes), "r"(smem_u32(mbar)) : "memory");
        if (r0 > 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(sP12)), "l"(gin + (size_t) F_P12 * fs + go - pitch), "r"(row_bytes), "r"(smem_u32(mbar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(sP22)), "l"(gin + (size_t) F_P22 * fs + go - pitch), "r"(row_bytes), "r"(smem_u32(mbar)) : "memory");
        }
    }
    if (r0 == 0)
        for (int i = tid; i < pitch; i += kResThreads) sP12[i] = sP22[i] = 0.f;     // p[-1] = 0

    // ---- per-thread pixel groups and their constants (registers) -------------------------------
    // per group: band row (bits 16..30), first column (bits 0..15), valid (sign bit clear)
    int qpos[kResQuads];
    float4 cix[kResQuads], ciy[kResQuads], crho[kResQuads];
#pragma unroll
    for (int k = 0; k < kResQuads; k++) {
        const int q = tid + k * kResThreads;
        const bool ok = q < nquads;
        const int lr = ok ? q / qpr : 0;
        const int cq = ok ? q - lr * qpr : 0;
        qpos[k] = (lr << 16) | (cq * 4) | (ok ? 0 : (int) 0x80000000);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        cix[k] = ciy[k] = crho[k] = z;
        if (ok) {
            const size_t o = (size_t) (r0 + lr) * pitch + cq * 4;
            cix[k] = ldg4(cst + C_IX * fs + o);
            ciy[k] = ldg4(cst + C_IY * fs + o);
            crho[k] = ldg4(cst + C_RHO * fs + o);
        }
    }
    {   // wait for the bulk copies (phase 0 of the mbarrier)
        unsigned int done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
    }
    __syncthreads();

    // DSMEM views of the neighbours' halo rows
    float *upU1 = nullptr, *upU2 = nullptr, *dnP12 = nullptr, *dnP22 = nullptr;
    if (rank > 0) {             // upper neighbour always holds RB rows: its halo-below row is row RB
        upU1 = cluster.map_shared_rank(sU1 + (size_t) RB * pitch, rank - 1);
        upU2 = cluster.map_shared_rank(sU2 + (size_t) RB * pitch, rank - 1);
    }
    if (rank < C - 1) {
        dnP12 = cluster.map_shared_rank(sP12, rank + 1);
        dnP22 = cluster.map_shared_rank(sP22, rank + 1);
    }
    if (C > 1) cluster.sync();  // every CTA of the cluster is resident before any remote store

    const double npix = (double) nx * (double) ny;
    int n = 0;
    double error = INFINITY;                                            // src/tvl1flow.cpp:111-112
    while (true) {
        n++;
        // ---- phase A: thresholding, divergence, primal update, error (:117-161) ----------------
        float errp = 0.f;
#pragma unroll
        for (int k = 0; k < kResQuads; k++) {
            const bool ok = qpos[k] >= 0;
            const int lrow = (qpos[k] >> 16) & 0x7fff, x0 = qpos[k] & 0xffff;
            const int o = lrow * pitch + x0;
            const int gy = r0 + lrow;
            const float4 u1 = lds4(sU1 + o), u2 = lds4(sU2 + o);
            const float4 p11 = lds4(sP11 + o), p21 = lds4(sP21 + o);
            const float4 p12 = lds4(sP12 + o + pitch), p22 = lds4(sP22 + o + pitch);
            const float4 a12 = lds4(sP12 + o), a22 = lds4(sP22 + o);
            float l11 = __shfl_up_sync(0xffffffffu, p11.w, 1);
            float l21 = __shfl_up_sync(0xffffffffu, p21.w, 1);
            if (x0 == 0) { l11 = 0.f; l21 = 0.f; }
            else if (lane == 0) { l11 = sP11[o - 1]; l21 = sP21[o - 1]; }
            const bool last_row = (gy == ny - 1);
            float o1[4], o2[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float a = TVL1_F4_GET(u1, e), c = TVL1_F4_GET(u2, e);
                const float ix = TVL1_F4_GET(cix[k], e), iy = TVL1_F4_GET(ciy[k], e);
                const bool last_col = (x0 + e >= nx - 1);
                primal_px(a, c, ix, iy, TVL1_F4_GET(crho[k], e), grad_of(ix, iy),
                          last_col ? 0.f : TVL1_F4_GET(p11, e), (e == 0) ? l11 : TVL1_F4_GET(p11, (e + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(p12, e), TVL1_F4_GET(a12, e),
                          last_col ? 0.f : TVL1_F4_GET(p21, e), (e == 0) ? l21 : TVL1_F4_GET(p21, (e + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(p22, e), TVL1_F4_GET(a22, e),
                          P.l_t, P.theta, o1[e], o2[e]);
                const float sq = update_sq(o1[e], a, o2[e], c);
                errp = __fadd_rn(errp, (ok && x0 + e < nx) ? sq : 0.f);
            }
            if (ok) {
                const float4 n1 = make_float4(o1[0], o1[1], o1[2], o1[3]);
                const float4 n2 = make_float4(o2[0], o2[1], o2[2], o2[3]);
                st4(sU1 + o, n1);
                st4(sU2 + o, n2);
                if (lrow == 0 && rank > 0) {     // my first row is the upper neighbour's row below
                    st4(upU1 + x0, n1);
                    st4(upU2 + x0, n2);
                }
            }
        }
        {
            const double e = warp_sum((double) errp);
            if (lane == 0) sWarp[warp] = e;
        }
        if (C > 1) cluster.sync(); else __syncthreads();

        // ---- phase B: forward gradient of u_new, dual update (:165-181) -------------------------
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < kResThreads / 32; w++) s += sWarp[w];
            for (int r = 0; r < C; r++) *cluster.map_shared_rank(sSlots + rank, r) = s;
        }
#pragma unroll
        for (int k = 0; k < kResQuads; k++) {
            const bool ok = qpos[k] >= 0;
            const int lrow = (qpos[k] >> 16) & 0x7fff, x0 = qpos[k] & 0xffff;
            const int o = lrow * pitch + x0;
            const int gy = r0 + lrow;
            const float4 u1 = lds4(sU1 + o), u2 = lds4(sU2 + o);
            const float4 b1 = lds4(sU1 + o + pitch), b2 = lds4(sU2 + o + pitch);
            float r1 = __shfl_down_sync(0xffffffffu, u1.x, 1);
            float r2 = __shfl_down_sync(0xffffffffu, u2.x, 1);
            if (lane == 31 && x0 + 4 < pitch) { r1 = sU1[o + 4]; r2 = sU2[o + 4]; }
            const float4 p11 = lds4(sP11 + o), p21 = lds4(sP21 + o);
            const float4 p12 = lds4(sP12 + o + pitch), p22 = lds4(sP22 + o + pitch);
            const bool has_below = (gy + 1 < ny);
            float q11[4], q12[4], q21[4], q22[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool last_col = (x0 + e >= nx - 1);
                const float c1 = TVL1_F4_GET(u1, e), c2 = TVL1_F4_GET(u2, e);
                const float e1 = (e == 3) ? r1 : TVL1_F4_GET(u1, (e + 1) & 3);
                const float e2 = (e == 3) ? r2 : TVL1_F4_GET(u2, (e + 1) & 3);
                q11[e] = TVL1_F4_GET(p11, e); q12[e] = TVL1_F4_GET(p12, e);
                q21[e] = TVL1_F4_GET(p21, e); q22[e] = TVL1_F4_GET(p22, e);
                dual_px(last_col ? 0.f : e1 - c1, has_below ? TVL1_F4_GET(b1, e) - c1 : 0.f,
                        last_col ? 0.f : e2 - c2, has_below ? TVL1_F4_GET(b2, e) - c2 : 0.f,
                        P.taut, q11[e], q12[e], q21[e], q22[e]);
            }
            if (ok) {
                const float4 n12 = make_float4(q12[0], q12[1], q12[2], q12[3]);
                const float4 n22 = make_float4(q22[0], q22[1], q22[2], q22[3]);
                st4(sP11 + o, make_float4(q11[0], q11[1], q11[2], q11[3]));
                st4(sP21 + o, make_float4(q21[0], q21[1], q21[2], q21[3]));
                st4(sP12 + o + pitch, n12);
                st4(sP22 + o + pitch, n22);
                if (lrow == rows - 1 && rank < C - 1) {      // my last row is the lower neighbour's row above
                    st4(dnP12 + x0, n12);
                    st4(dnP22 + x0, n22);
                }
            }
        }
        if (C > 1) cluster.sync(); else __syncthreads();

        // ---- stopping rule (:113), identical in every CTA of the cluster -------------------------
        double tot = 0.0;
        for (int r = 0; r < C; r++) tot += sSlots[r];
        error = tot / npix;                                              // :162
        if (P.err_trace && b == 0 && rank == 0 && tid == 0) P.err_trace[n - 1] = error;
        if (!(error > P.eps2 && n < P.max_iter)) break;
    }

    // ---- write the band back (other ping-pong set) with TMA bulk stores --------------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const size_t go = (size_t) r0 * pitch;
        const float *srcs[6] = { sU1, sU2, sP11, sP12 + pitch, sP21, sP22 + pitch };
#pragma unroll
        for (int f = 0; f < 6; f++)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(gout + (size_t) f * fs + go), "r"(smem_u32(srcs[f])), "r"(band_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (rank == 0) {
            ctl->cur = cur ^ 1;
            ctl->n = n;
            ctl->err = error;
            ctl->active = 0;
            P.stat_iters[(size_t) b * P.stat_stride + P.stat_slot] = n;
            P.stat_errs[(size_t) b * P.stat_stride + P.stat_slot] = error;
            atomicAdd(P.counters + P.level, (unsigned long long) n * (unsigned long long) nx * (unsigned long long) ny);
            if (b == 0) atomicAdd(P.counters + kStatLevels + P.level, 1ull);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// (c'') temporally blocked streaming kernel: up to kTbT iterations per launch on 2-D TMA halo tiles
// ------------------------------------------------------------------------------------------------
//
// For levels too large to stay on chip.  A CTA owns a 56x24 tile; one elected thread pulls the 64x32
// box around it (halo kTbT on every side) of the six evolving planes and the three constant planes
// into shared memory with nine `cp.async.bulk.tensor.3d` copies on one mbarrier (out-of-image parts
// arrive as zeros, which is exactly the p[-1] = 0 rule of src/operators.cpp:35-78); then the CTA runs
// `nsteps` (<= kTbT) complete iterations of src/tvl1flow.cpp:114-181 in place in shared memory -- the
// region that is still exact shrinks by one pixel per side per iteration and ends at the owned tile
// -- and writes the tile back with float4 stores.  HBM traffic per iteration: (9*1.52 + 6)/nsteps
// planes instead of 15.  The error of every one of the nsteps iterations is reduced separately
// (owned pixels only), so the last CTA of the pair can apply the stopping rule to each of them and
// order an exact replay when the block overshot (decide_block).
constexpr int kTbBW = 64, kTbBH = 32;                                  // box
constexpr int kTbW = kTbBW - 2 * kTbT, kTbH = kTbBH - 2 * kTbT;        // owned tile: 56 x 24
constexpr int kTbThreads = 256;
constexpr int kTbPlane = kTbBW * kTbBH;
constexpr int kTbPlanes = 9;
constexpr size_t kTbSmemBytes = (size_t) kTbPlanes * kTbPlane * sizeof(float);

struct TbMaps {
    CUtensorMap state[2];      // dims (nx, ny, 6*B) of ping-pong set 0 / 1
    CUtensorMap consts;        // dims (nx, ny, 4*B)
};

// The iterations of k_iterate_tb on the shared-memory box.  INTERIOR = the box lies inside the image
// and touches neither its last column nor its last row: all boundary predicates fold away.
template <bool INTERIOR>
__device__ __forceinline__ void tb_iterations(const IterParams &P, int ns, int X0, int Y0, int nx, int ny,
                                              int own_rows, float *sU1, float *sU2, float *sP11, float *sP12, float *sP21,
                                              float *sP22, const float *sIx, const float *sIy, const float *sRho,
                                              double (*s_err)[kTbThreads / 32])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qx = tid & 15, ry = tid >> 4;                  // two groups of 4 pixels: rows ry and ry+16
    const int bx0 = qx * 4;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int t = 0; t < ns; t++) {
        // ---- phase A: thresholding, divergence, primal update, error ----------------------------
        float errp = 0.f;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int by = ry + 16 * k, o = by * kTbBW + bx0;
            const int gy = Y0 + by;
            const float4 u1 = lds4(sU1 + o), u2 = lds4(sU2 + o);
            const float4 p11 = lds4(sP11 + o), p21 = lds4(sP21 + o);
            const float4 p12 = lds4(sP12 + o), p22 = lds4(sP22 + o);
            const float4 a12 = by > 0 ? lds4(sP12 + o - kTbBW) : zero4;
            const float4 a22 = by > 0 ? lds4(sP22 + o - kTbBW) : zero4;
            const float4 ix = lds4(sIx + o), iy = lds4(sIy + o), rc = lds4(sRho + o);
            float l11 = __shfl_up_sync(0xffffffffu, p11.w, 1);
            float l21 = __shfl_up_sync(0xffffffffu, p21.w, 1);
            if (qx == 0) { l11 = 0.f; l21 = 0.f; }
            const bool last_row = !INTERIOR && (gy == ny - 1);
            const bool row_in = INTERIOR || (gy >= 0 && gy < ny);
            const bool row_owned = by >= kTbT && by < kTbT + own_rows;   // (a row band may end inside the tile)
            float o1[4], o2[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int gx = X0 + bx0 + e;
                const bool in_img = INTERIOR || (row_in && gx >= 0 && gx < nx);
                const bool last_col = !INTERIOR && gx >= nx - 1;
                const float a = TVL1_F4_GET(u1, e), c = TVL1_F4_GET(u2, e);
                const float gxv = TVL1_F4_GET(ix, e), gyv = TVL1_F4_GET(iy, e);
                primal_px(a, c, gxv, gyv, TVL1_F4_GET(rc, e), grad_of(gxv, gyv),
                          last_col ? 0.f : TVL1_F4_GET(p11, e), (e == 0) ? l11 : TVL1_F4_GET(p11, (e + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(p12, e), TVL1_F4_GET(a12, e),
                          last_col ? 0.f : TVL1_F4_GET(p21, e), (e == 0) ? l21 : TVL1_F4_GET(p21, (e + 3) & 3),
                          last_row ? 0.f : TVL1_F4_GET(p22, e), TVL1_F4_GET(a22, e),
                          P.l_t, P.theta, o1[e], o2[e]);
                if (!in_img) { o1[e] = a; o2[e] = c; }
                const bool owned = row_owned && in_img && bx0 + e >= kTbT && bx0 + e < kTbT + kTbW;
                errp = __fadd_rn(errp, owned ? update_sq(o1[e], a, o2[e], c) : 0.f);
            }
            st4(sU1 + o, make_float4(o1[0], o1[1], o1[2], o1[3]));
            st4(sU2 + o, make_float4(o2[0], o2[1], o2[2], o2[3]));
        }
        {
            const double e = warp_sum((double) errp);
            if (lane == 0) s_err[t][warp] = e;
        }
        __syncthreads();
        // ---- phase B: forward gradient of u_new, dual update --------------------------------------
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int by = ry + 16 * k, o = by * kTbBW + bx0;
            const int gy = Y0 + by;
            const float4 u1 = lds4(sU1 + o), u2 = lds4(sU2 + o);
            const float4 b1 = by + 1 < kTbBH ? lds4(sU1 + o + kTbBW) : zero4;
            const float4 b2 = by + 1 < kTbBH ? lds4(sU2 + o + kTbBW) : zero4;
            float r1 = __shfl_down_sync(0xffffffffu, u1.x, 1);
            float r2 = __shfl_down_sync(0xffffffffu, u2.x, 1);
            if (qx == 15) { r1 = 0.f; r2 = 0.f; }
            const float4 p11 = lds4(sP11 + o), p21 = lds4(sP21 + o);
            const float4 p12 = lds4(sP12 + o), p22 = lds4(sP22 + o);
            const bool has_below = INTERIOR || gy + 1 < ny;
            const bool row_in = INTERIOR || (gy >= 0 && gy < ny);
            float q11[4], q12[4], q21[4], q22[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int gx = X0 + bx0 + e;
                const bool in_img = INTERIOR || (row_in && gx >= 0 && gx < nx);
                const bool last_col = !INTERIOR && gx >= nx - 1;
                const float c1 = TVL1_F4_GET(u1, e), c2 = TVL1_F4_GET(u2, e);
                const float e1 = (e == 3) ? r1 : TVL1_F4_GET(u1, (e + 1) & 3);
                const float e2 = (e == 3) ? r2 : TVL1_F4_GET(u2, (e + 1) & 3);
                q11[e] = TVL1_F4_GET(p11, e); q12[e] = TVL1_F4_GET(p12, e);
                q21[e] = TVL1_F4_GET(p21, e); q22[e] = TVL1_F4_GET(p22, e);
                dual_px(last_col ? 0.f : e1 - c1, has_below ? TVL1_F4_GET(b1, e) - c1 : 0.f,
                        last_col ? 0.f : e2 - c2, has_below ? TVL1_F4_GET(b2, e) - c2 : 0.f,
                        P.taut, q11[e], q12[e], q21[e], q22[e]);
                if (!in_img) { q11[e] = q12[e] = q21[e] = q22[e] = 0.f; }   // outside the image p stays 0
            }
            st4(sP11 + o, make_float4(q11[0], q11[1], q11[2], q11[3]));
            st4(sP12 + o, make_float4(q12[0], q12[1], q12[2], q12[3]));
            st4(sP21 + o, make_float4(q21[0], q21[1], q21[2], q21[3]));
            st4(sP22 + o, make_float4(q22[0], q22[1], q22[2], q22[3]));
        }
        __syncthreads();
    }

}

// One block of up to kTbT iterations of pair b on this CTA's tile.  `mbar` was initialised by the
// kernel; `parity` is the phase this use of it completes.
__device__ __forceinline__ void tb_pair(const TbMaps &maps, const IterParams &P, const int b, float *tb_smem,
                                        double (*s_err)[kTbThreads / 32], double *s_tot,
                                        unsigned long long &mbar, int &s_last, const unsigned int parity)
{
    PairCtl *ctl = P.ctl + b;
    const int ns = min(ctl->nsteps, kTbT);
    const int cur = ctl->cur;
    const int nx = P.lv.nx, ny = P.lv.ny, pitch = P.lv.pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles cover the rows this launch owns (the whole image, or this rank's row band)
    const int X0 = blockIdx.x * kTbW - kTbT, Y0 = P.row_begin + blockIdx.y * kTbH - kTbT;
    const int own_rows = min(kTbH, P.row_end - (Y0 + kTbT));

    float *sU1 = tb_smem, *sU2 = sU1 + kTbPlane, *sP11 = sU2 + kTbPlane, *sP12 = sP11 + kTbPlane,
          *sP21 = sP12 + kTbPlane, *sP22 = sP21 + kTbPlane, *sIx = sP22 + kTbPlane, *sIy = sIx + kTbPlane,
          *sRho = sIy + kTbPlane;

    if (tid == 0) {
        // the previous pair's tile was read out of this memory with ordinary loads: order them
        // before the TMA writes (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(smem_u32(&mbar)), "r"((unsigned int) kTbSmemBytes) : "memory");
        const CUtensorMap *ms = &maps.state[cur];
#pragma unroll
        for (int f = 0; f < F_COUNT; f++)                   // smem plane order == enum Field
            tma_load_3d(tb_smem + f * kTbPlane, ms, X0, Y0, f * P.batch + b, &mbar);
        tma_load_3d(sIx, &maps.consts, X0, Y0, C_IX * P.batch + b, &mbar);
        tma_load_3d(sIy, &maps.consts, X0, Y0, C_IY * P.batch + b, &mbar);
        tma_load_3d(sRho, &maps.consts, X0, Y0, C_RHO * P.batch + b, &mbar);
    }
    {
        unsigned int done = 0;
        const long long t0 = clock64();
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
            if (!done && clock64() - t0 > (1ll << 31)) break;   // a bad descriptor must not hang the GPU
        }
    }

    {
        const bool interior = X0 >= 0 && Y0 >= 0 && X0 + kTbBW < nx && Y0 + kTbBH < ny;   // CTA-uniform
        if (interior)
            tb_iterations<true>(P, ns, X0, Y0, nx, ny, own_rows, sU1, sU2, sP11, sP12, sP21, sP22, sIx, sIy, sRho, s_err);
        else
            tb_iterations<false>(P, ns, X0, Y0, nx, ny, own_rows, sU1, sU2, sP11, sP12, sP21, sP22, sIx, sIy, sRho, s_err);
    }

    // ---- write the owned tile to the other ping-pong set ------------------------------------------
    {
        float *gout = P.state + (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0;
        const size_t fs = P.field_stride;
        for (int idx = tid; idx < (kTbW / 4) * kTbH; idx += kTbThreads) {
            const int row = idx / (kTbW / 4), q = idx - row * (kTbW / 4);
            const int by = kTbT + row, bx = kTbT + 4 * q;
            const int gy = Y0 + by, gx = X0 + bx;
            if (row >= own_rows || gx >= nx) continue;
            const int so = by * kTbBW + bx;
            const size_t go = (size_t) gy * pitch + gx;
#pragma unroll
            for (int f = 0; f < F_COUNT; f++) st4(gout + f * fs + go, lds4(tb_smem + f * kTbPlane + so));
            if (P.peers.enabled) {
                // rows within `halo` of a band edge are the neighbour's halo rows: the same values go to
                // the same offsets of its planes (NVLink stores), see iterate_t1_pair
                const size_t po = (size_t) (cur ^ 1) * P.set_stride + (size_t) b * P.plane0 + go;
                if (gy < P.row_begin + P.peers.halo && P.peers.up_state) {
#pragma unroll
                    for (int f = 0; f < F_COUNT; f++) st4(P.peers.up_state + po + f * fs, lds4(tb_smem + f * kTbPlane + so));
                }
                if (gy >= P.row_end - P.peers.halo && P.peers.dn_state) {
#pragma unroll
                    for (int f = 0; f < F_COUNT; f++) st4(P.peers.dn_state + po + f * fs, lds4(tb_smem + f * kTbPlane + so));
                }
            }
        }
    }

    // ---- per-iteration error sums: CTA -> fixed-order sum by the pair's last CTA -------------------
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    double *part = P.tb_partials + ((size_t) b * P.tb_parts) * kTbT;
    if (tid == 0) {
        for (int t = 0; t < ns; t++) {
            double s = 0.0;
            for (int w = 0; w < kTbThreads / 32; w++) s += s_err[t][w];
            part[(size_t) blk * kTbT + t] = s;
        }
        if (P.peers.enabled) __threadfence_system();    // halo rows pushed to the neighbours are out
        else __threadfence();
        const unsigned int tk = atomicAdd(&ctl->arrive, 1u);
        s_last = (tk == (unsigned int) nblk - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile double *vp = part;
    for (int t = 0; t < ns; t++) {
        double s = 0.0;
        for (int i = tid; i < nblk; i += kTbThreads) s += vp[(size_t) i * kTbT + t];
        s = warp_sum(s);
        __syncthreads();
        if (lane == 0) s_err[0][warp] = s;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < kTbThreads / 32; w++) tot += s_err[0][w];
            s_tot[t] = tot;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (P.peers.enabled) band_all_to_all(P.peers, s_tot, ns);     // the other bands' sums of this block
        for (int t = 0; t < ns; t++) s_tot[t] /= (double) nx * (double) ny;
        decide_block(P, ctl, b, cur, ns, s_tot,
                     (unsigned long long) nx * (unsigned long long) (P.row_end - P.row_begin));
    }
}

__global__ void __launch_bounds__(kTbThreads, 3)
k_iterate_tb(const __grid_constant__ TbMaps maps, const IterParams P)
{
    extern __shared__ __align__(128) float tb_smem[];
    __shared__ double s_err[kTbT][kTbThreads / 32];
    __shared__ double s_tot[kTbT];
    __shared__ unsigned long long mbar;
    __shared__ int s_last;

    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        atomicAdd(P.px_iters + kStatLevels + P.level, 1ull);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    __shared__ unsigned int s_parity;                // phase of the next use of mbar
    if (threadIdx.x == 0) s_parity = 0u;
    for_each_pair_of_slot<true>(P, true, [&](int b) {      // (a barrier separates this from the first read)
        const unsigned int parity = *(volatile unsigned int *) &s_parity;
        tb_pair(maps, P, b, tb_smem, s_err, s_tot, mbar, s_last, parity);
        __syncthreads();
        if (threadIdx.x == 0) s_parity ^= 1u;
    });
}

} // namespace tvl1
