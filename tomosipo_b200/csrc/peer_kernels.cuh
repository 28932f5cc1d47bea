// Row-band exchange of the multi-GPU backprojection over peer memory (tomosipo_b200/distributed.py, bp_exchange =
// "rows"): every rank stores the detector rows each peer's z-slab reads straight into that peer's band buffer
// [rows][all angles][U], at the angle offset of its own block - NVLink stores from one kernel, no staging copy, no
// interleave pass on the receiving side.  The buffers are cudaMalloc'ed by the library and opened by the peers
// through CUDA IPC (tsp_peer_alloc / tsp_peer_open).
#pragma once
#include <cstdint>

constexpr int PUSH_MAX_JOBS = 16;  // one job per destination rank (own band included); more ranks -> several launches

struct PushJob {
    const float *src;     // first row of the band in this rank's angle block [V][A_own][U]
    float *dst;           // the same rows in the peer's buffer, at this rank's angle offset
    long long rows;       // rows of the band
    long long width;      // floats per row (A_own * U)
    long long src_pitch;  // floats between rows in the source (A_own * U)
    long long dst_pitch;  // floats between rows in the destination (A_all * U)
};

struct PushArgs {
    PushJob job[PUSH_MAX_JOBS];
};

// grid: (blocks per job, jobs).  VEC = 4: every pointer, width and pitch is a multiple of four floats.
template <int VEC>
__global__ void __launch_bounds__(512) push_rows_kernel(const __grid_constant__ PushArgs args)
{
    const PushJob &j = args.job[blockIdx.y];
    const long long w = j.width / VEC;
    const long long total = j.rows * w;
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC == 4) {
        // four independent 16-byte loads in flight per thread before the (fire-and-forget) remote stores
        for (; i + 3 * step < total; i += 4 * step) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long e = i + k * step, r = e / w, c = e - r * w;
                v[k] = __ldcs(reinterpret_cast<const float4 *>(j.src + r * j.src_pitch) + c);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long e = i + k * step, r = e / w, c = e - r * w;
                reinterpret_cast<float4 *>(j.dst + r * j.dst_pitch)[c] = v[k];
            }
        }
        for (; i < total; i += step) {
            const long long r = i / w, c = i - r * w;
            reinterpret_cast<float4 *>(j.dst + r * j.dst_pitch)[c] =
                __ldcs(reinterpret_cast<const float4 *>(j.src + r * j.src_pitch) + c);
        }
    } else {
        for (; i < total; i += step) {
            const long long r = i / w, c = i - r * w;
            j.dst[r * j.dst_pitch + c] = __ldcs(j.src + r * j.src_pitch + c);
        }
    }
}
