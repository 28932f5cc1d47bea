// FDK pre-filter (reference: tomosipo/astra.py:374-406 -> astra.experimental.accumulate_FDK): the element-wise
// passes around the ramp filter's FFT, each fused into one kernel.
//   fdk_preweight_kernel  out[v][a][0..pitch) = proj[v][a][u] * cos-weight(v, a, u) * redundancy(a, u), zero beyond
//                         u = U: cosine weighting, Parker / full-circle redundancy weighting and the zero padding
//                         of the FFT input in one pass (one read of the projections, one write of the padded rows)
//   fdk_ramp_kernel       spectrum[row][k] *= G[k]                  (band-limited ramp, real and even)
//   fdk_scale_crop_kernel q[v][a][u] = filtered[v][a][u] * c[a]     (crop of the padded rows + per-angle constant:
//                         angular step, distance weights, filter pitch - see fdk_angle_constants in tsproj.cu)
#pragma once
#include <cuda_runtime.h>

namespace tsp {

struct FDKAngle {      // per angle, fp32 (derived in fp64 on the host)
    float pu, pv;      // detector pixel pitch along u, v
    float sdd;         // source - detector-plane distance
    float ppu, ppv;    // principal point, in pixels from the detector centre
    float scale;       // per-angle constant applied after the filter
};

__global__ void __launch_bounds__(256) fdk_preweight_kernel(const float *__restrict__ proj, float *__restrict__ out,
                                                            const FDKAngle *__restrict__ tab,
                                                            const float *__restrict__ redundancy, int det_u, int det_v,
                                                            int n_angles, int pitch)
{
    const int row = blockIdx.x;  // = v * n_angles + a
    const int a = row % n_angles, v = row / n_angles;
    const FDKAngle t = tab[a];
    const float vp = ((float)v + 0.5f - 0.5f * (float)det_v - t.ppv) * t.pv;
    const float base = t.sdd * t.sdd + vp * vp;
    const float *src = proj + (size_t)row * det_u;
    float *dst = out + (size_t)row * pitch;
    const float *red = redundancy ? redundancy + (size_t)a * det_u : nullptr;
    for (int u = threadIdx.x; u < pitch; u += blockDim.x) {
        float val = 0.0f;
        if (u < det_u) {
            const float up = ((float)u + 0.5f - 0.5f * (float)det_u - t.ppu) * t.pu;
            val = __ldg(src + u) * t.sdd * rsqrtf(base + up * up) * (red ? __ldg(red + u) : 0.5f);
        }
        dst[u] = val;
    }
}

__global__ void __launch_bounds__(256) fdk_ramp_kernel(float2 *__restrict__ spec, const float *__restrict__ G, int nfreq,
                                                       size_t n_rows)
{
    const size_t row = blockIdx.x;
    if (row >= n_rows) return;
    float2 *s = spec + row * nfreq;
    for (int k = threadIdx.x; k < nfreq; k += blockDim.x) {
        const float g = __ldg(G + k);
        float2 v = s[k];
        v.x *= g; v.y *= g;
        s[k] = v;
    }
}

__global__ void __launch_bounds__(256) fdk_scale_crop_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                             const FDKAngle *__restrict__ tab, int det_u, int n_angles,
                                                             int pitch)
{
    const int row = blockIdx.x;
    const float c = tab[row % n_angles].scale;
    const float *src = in + (size_t)row * pitch;
    float *dst = out + (size_t)row * det_u;
    for (int u = threadIdx.x; u < det_u; u += blockDim.x) dst[u] = src[u] * c;
}

}  // namespace tsp
