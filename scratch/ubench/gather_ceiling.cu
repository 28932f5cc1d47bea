// Micro-benchmark of the two ceilings the projection kernels are measured against:
//   (a) shared-memory gather: one "update" = 4 bilinear taps (LDS.32) from a staged tile, conflict-free
//       addresses (lane i reads column base + i), nothing else but the accumulate  -> crossbar ceiling
//   (b) the same 4 taps with the arithmetic of a real update (floor by magic add, 2 weights, 3 lerps,
//       weighted accumulate: ~13 instructions)                                      -> issue ceiling
//   (c) 3 taps per update (the 3-row z-invariant BP loop: 6 taps per voxel pair) with its arithmetic
// Output: updates / clk / SM for each, at the clock the run actually had (SM cycles from clock64()).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_ceiling gather_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int PITCH = 68, ROWS = 46, THREADS = 256, ITERS = 4096;

__device__ __forceinline__ float lds(uint32_t a, int off)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a + off) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) k(float *out, float fu0, float fv0, float dv, long long *cycles)
{
    __shared__ float tile[ROWS * PITCH];
    for (int i = threadIdx.x; i < ROWS * PITCH; i += THREADS) tile[i] = (float)(i % 7);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc0 = 0.f, acc1 = 0.f;
    float fu = fu0 + lane * 0.8f, fv = fv0 + warp * 0.5f;
    const long long t0 = clock64();
    if (MODE == 0) {
        uint32_t a = base + 4u * (uint32_t)(lane + PITCH * warp);
#pragma unroll 8
        for (int it = 0; it < ITERS; ++it) {
            acc0 += lds(a, 0) + lds(a, 4);
            acc1 += lds(a, 4 * PITCH) + lds(a, 4 * PITCH + 4);
            a += 4u * PITCH;
            if (a >= base + 4u * PITCH * (ROWS - 2)) a -= 4u * PITCH * (ROWS - 2 - 8);
        }
    } else if (MODE == 1) {
        const float M = 12582912.0f;
#pragma unroll 8
        for (int it = 0; it < ITERS; ++it) {
            const float ru = __fadd_rd(fu, M), rv = __fadd_rd(fv, M);
            const float wu = fu - (ru - M), wv = fv - (rv - M);
            const uint32_t a = (__float_as_uint(rv) * (uint32_t)PITCH + __float_as_uint(ru)) * 4u + (base - 4u * 0x4B400000u * (PITCH + 1));
            const float p00 = lds(a, 0), p10 = lds(a, 4), p01 = lds(a, 4 * PITCH), p11 = lds(a, 4 * PITCH + 4);
            const float lo = fmaf(wu, p10 - p00, p00), hi = fmaf(wu, p11 - p01, p01);
            acc0 = fmaf(dv, fmaf(wv, hi - lo, lo), acc0);
            fv += dv;
            if (fv > (float)(ROWS - 3)) fv -= (float)(ROWS - 12);
        }
    } else {
        const float M = 12582912.0f;
        const float wu = 0.3f;
        const uint32_t cb = base - 4u * 0x4B400000u * PITCH + 4u * (uint32_t)lane;
#pragma unroll 8
        for (int it = 0; it < ITERS; it += 2) {  // one voxel pair: 3 rows x 2 columns
            const float rv = __fadd_rd(fv, M), rv1 = __fadd_rd(fv + dv, M);
            const float wv = fv - (rv - M), wv1 = fv + dv - (rv1 - M);
            const uint32_t a = __float_as_uint(rv) * (uint32_t)(4 * PITCH) + cb;
            const float p00 = lds(a, 0), p01 = lds(a, 4), p10 = lds(a, 4 * PITCH), p11 = lds(a, 4 * PITCH + 4);
            const float p20 = lds(a, 8 * PITCH), p21 = lds(a, 8 * PITCH + 4);
            const float h0 = fmaf(wu, p01 - p00, p00), h1 = fmaf(wu, p11 - p10, p10), h2 = fmaf(wu, p21 - p20, p20);
            const bool nxt = __float_as_uint(rv1) != __float_as_uint(rv);
            acc0 = fmaf(dv, fmaf(wv, h1 - h0, h0), acc0);
            acc1 = fmaf(dv, nxt ? fmaf(wv1, h2 - h1, h1) : fmaf(wv1, h1 - h0, h0), acc1);
            fv += 2.0f * dv;
            if (fv > (float)(ROWS - 4)) fv -= (float)(ROWS - 12);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int sms)
{
    const int ctas = 2 * sms;  // exactly one wave at 2 CTAs / SM, like the projection kernels
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * ctas * THREADS);
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    for (int rep = 0; rep < 3; ++rep) k<MODE><<<ctas, THREADS>>>(out, 3.1f, 2.2f, 0.8f, cyc);
    cudaDeviceSynchronize();
    long long *h = new long long[ctas];
    cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    // per SM: 2 CTAs x THREADS lanes x ITERS updates in `avg` cycles
    printf("%-44s %6.2f updates/clk/SM  (%.0f cycles per CTA, err %s)\n", name, 2.0 * THREADS * ITERS / avg, avg,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("(a) 4 LDS.32 taps per update, no arithmetic", p.multiProcessorCount);
    run<1>("(b) 4 taps + bilinear arithmetic", p.multiProcessorCount);
    run<2>("(c) 3-row loop: 6 taps per voxel pair", p.multiProcessorCount);
    return 0;
}
