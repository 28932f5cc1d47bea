// Texture arm of the interpolation-path comparison (north_star: "a texture-unit versus shared-memory
// interpolation choice justified by ncu counters"; VERDICT r01 row g).
//
// Same access pattern as the backprojector's inner loop (a warp = 32 consecutive x voxels whose detector
// column advances by ~0.8 pixels per lane, each thread walks a 32-voxel z run whose detector row advances by
// dv = 0.8 per voxel), one 2-D projection image per "angle" held in a cudaArray:
//   (t1) tex2D<float>, hardware bilinear filter (cudaFilterModeLinear): 1 fetch per update, 9-bit weights
//        (what ASTRA's cone_bp does; NOT usable for the 1e-5 fp64 target)
//   (t2) tex2Dgather (TLD4): the 2 x 2 footprint in one instruction + the exact fp32 lerp in the SM
//   (t3) 4 point-sampled tex2D fetches + fp32 lerp
//   (s)  the shared-memory arms, for the same run, are in gather_ceiling.cu
// Output: updates / clk / SM (SM cycles from clock64()), one wave of 2 CTAs per SM x 256 threads.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tex_arm tex_arm.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int THREADS = 256, ITERS = 4096, W = 768, H = 512;

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) k(cudaTextureObject_t tex_lin, cudaTextureObject_t tex_pt, float *out,
                                                 float fu0, float fv0, float dv, long long *cycles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // every CTA works on its own patch of the image, like a voxel tile's footprint
    const float fu = fu0 + lane * 0.8f + (float)((blockIdx.x * 37) % (W - 64));
    float fv = fv0 + warp * 0.5f + (float)((blockIdx.x * 11) % (H - 64));
    const float fv_wrap = fv + 40.0f;
    float acc = 0.f;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
            acc = fmaf(dv, tex2D<float>(tex_lin, fu + 0.5f, fv + 0.5f), acc);
        } else if (MODE == 1) {
            const float flu = floorf(fu), flv = floorf(fv);
            const float wu = fu - flu, wv = fv - flv;
            // gather returns (x0y1, x1y1, x1y0, x0y0) of the 2x2 footprint around the sample point
            const float4 g = tex2Dgather<float4>(tex_pt, flu + 1.0f, flv + 1.0f, 0);
            const float lo = fmaf(wu, g.z - g.w, g.w), hi = fmaf(wu, g.y - g.x, g.x);
            acc = fmaf(dv, fmaf(wv, hi - lo, lo), acc);
        } else {
            const float flu = floorf(fu), flv = floorf(fv);
            const float wu = fu - flu, wv = fv - flv;
            const float p00 = tex2D<float>(tex_pt, flu + 0.5f, flv + 0.5f), p10 = tex2D<float>(tex_pt, flu + 1.5f, flv + 0.5f);
            const float p01 = tex2D<float>(tex_pt, flu + 0.5f, flv + 1.5f), p11 = tex2D<float>(tex_pt, flu + 1.5f, flv + 1.5f);
            const float lo = fmaf(wu, p10 - p00, p00), hi = fmaf(wu, p11 - p01, p01);
            acc = fmaf(dv, fmaf(wv, hi - lo, lo), acc);
        }
        fv += dv;
        if (fv > fv_wrap) fv -= 32.0f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int sms, cudaTextureObject_t tl, cudaTextureObject_t tp)
{
    const int ctas = 2 * sms;
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * ctas * THREADS);
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    for (int rep = 0; rep < 3; ++rep) k<MODE><<<ctas, THREADS>>>(tl, tp, out, 3.1f, 2.2f, 0.8f, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    printf("%-58s %6.2f updates/clk/SM  (%.0f cycles per CTA, %s)\n", name, 2.0 * THREADS * ITERS / avg, avg,
           cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    std::vector<float> img((size_t)W * H);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (float)(i % 13) * 0.25f;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather);
    cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex_lin, tex_pt;
    td.filterMode = cudaFilterModeLinear;
    cudaCreateTextureObject(&tex_lin, &rd, &td, nullptr);
    td.filterMode = cudaFilterModePoint;
    cudaCreateTextureObject(&tex_pt, &rd, &td, nullptr);
    printf("setup: %s\n", cudaGetErrorString(cudaGetLastError()));
    run<0>("(t1) tex2D hardware bilinear (9-bit weights), 1 fetch", p.multiProcessorCount, tex_lin, tex_pt);
    run<1>("(t2) tex2Dgather (TLD4) + fp32 lerp", p.multiProcessorCount, tex_lin, tex_pt);
    run<2>("(t3) 4 point-sampled fetches + fp32 lerp", p.multiProcessorCount, tex_lin, tex_pt);
    return 0;
}
