// Micro-benchmark: scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100) throughput on one GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

__global__ void k_scalar(float *out, float a, float b)
{
    float x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; k++) x[k] = threadIdx.x * 0.001f + k;
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int k = 0; k < kChains; k++) x[k] = __fmaf_rn(x[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float *out, float a, float b)
{
    float2 x[kChains / 2];
#pragma unroll
    for (int k = 0; k < kChains / 2; k++) x[k] = make_float2(threadIdx.x * 0.001f + 2 * k, threadIdx.x * 0.001f + 2 * k + 1);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int k = 0; k < kChains / 2; k++) x[k] = __ffma2_rn(x[k], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains / 2; k++) s += x[k].x + x[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three varying register operands (what a real kernel has): x = x * y + z with y, z per thread
__global__ void k_scalar3(float *out, const float *in)
{
    float x[kChains], y[kChains], z[kChains];
#pragma unroll
    for (int k = 0; k < kChains; k++) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 32 + k]; z[k] = in[threadIdx.x + 64 + k]; }
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int k = 0; k < kChains; k++) x[k] = __fmaf_rn(x[k], y[k], z[k]);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed3(float *out, const float *in)
{
    float2 x[kChains / 2], y[kChains / 2], z[kChains / 2];
#pragma unroll
    for (int k = 0; k < kChains / 2; k++) {
        x[k] = make_float2(in[threadIdx.x + 2 * k], in[threadIdx.x + 2 * k + 1]);
        y[k] = make_float2(in[threadIdx.x + 32 + 2 * k], in[threadIdx.x + 33 + 2 * k]);
        z[k] = make_float2(in[threadIdx.x + 64 + 2 * k], in[threadIdx.x + 65 + 2 * k]);
    }
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int k = 0; k < kChains / 2; k++) x[k] = __ffma2_rn(x[k], y[k], z[k]);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains / 2; k++) s += x[k].x + x[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256;
    float *out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float *in;
    cudaMalloc(&in, sizeof(float) * 4096);
    cudaMemset(in, 0, sizeof(float) * 4096);
    for (int which = 0; which < 4; which++) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (which == 0) k_scalar<<<blocks, threads>>>(out, 0.999f, 0.001f);
            else if (which == 1) k_packed<<<blocks, threads>>>(out, 0.999f, 0.001f);
            else if (which == 2) k_scalar3<<<blocks, threads>>>(out, in);
            else k_packed3<<<blocks, threads>>>(out, in);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double fmas = (double) blocks * threads * kIters * kChains;
        printf("%s: %.3f ms, %.1f TFLOP/s (2 flop per FMA), %.1f FMA per clock per SM at 1.9 GHz\n",
               which == 0 ? "scalar FFMA, uniform multiplicand / addend " : which == 1 ? "packed FFMA2, uniform multiplicand / addend" : which == 2 ? "scalar FFMA, three register operands      " : "packed FFMA2, three register operands     ", best, 2 * fmas / best / 1e9, fmas / (best * 1e-3) / 1.9e9 / sms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
