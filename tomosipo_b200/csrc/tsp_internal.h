// Internal structures shared by the host side and the kernels of libtsproj.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tsproj.h"

namespace tsp {

// One projection angle in the *normalised* frame (unit voxels, volume centred
// on the origin), with every 3-vector permuted to (march, p, q) order so a
// kernel never indexes a vector with a runtime axis id.  fp64: per-ray set-up
// is done in double precision, the march itself in fp32.
struct FPAngle {
    double o[3];   // cone: source position; parallel: ray direction
    double d0[3];  // lower-left *corner* of detector pixel (0, 0)
    double u[3];
    double v[3];
};

// Voxel -> detector map of one angle (normalised frame, (x, y, z, 1) order):
//   U = nu.X / dn.X,  V = nv.X / dn.X   (texel-centre convention: pixel i spans [i, i+1))
// cone:     dn is scaled so that 1/dn.X^2 is the ray-density weight
// parallel: dn = (0,0,0,1) and `weight` carries 1/|u x v|
struct BPAngle {
    double nu[4];
    double nv[4];
    double dn[4];
    double weight;
};

// One FP launch group: all angles that march along the same axis and read
// the same volume layout.
struct FPGroup {
    int march;           // 0 x, 1 y, 2 z  (volume axis marched)
    int p_axis, q_axis;  // in-slice axes; p is the memory-contiguous one
    bool transposed;     // reads the (z, x, y) copy instead of (z, y, x)
    bool columns;        // det_v is parallel to the q axis: fp_cols_kernel applies
    std::vector<int> angles;
    // TMA-staged kernel (fp_tma_kernel): angles paired with a neighbour whose footprint nearly
    // coincides (second = -1: single), and the staged box that bounds every pair's footprint.
    std::vector<int> pairs;  // 2 ints per pair
    int box_w = 0, box_h = 0;  // 0: group not eligible
    int rows_per_thread = 4;   // 4 or 8: a CTA covers 32 x (4 * rows_per_thread) pixels
};

// The tensor map is an opaque 128-byte, 64-byte aligned CUtensorMap.
struct alignas(64) TensorMapBlob { unsigned char bytes[128]; };
// The two descriptors of a launch (two box widths = two shared-memory pitches) travel as a
// __grid_constant__ kernel parameter: no device copy to keep alive, safe under stream capture.
struct alignas(64) TensorMapPair { TensorMapBlob m[2]; };

struct DeviceState {
    FPAngle *fp_angles = nullptr;  // [n_angles], permuted per angle
    int *fp_lists = nullptr;       // concatenated group lists
    int *fp_pairs = nullptr;       // concatenated group pair lists (2 ints per pair)
    std::vector<size_t> pair_offset;  // in pairs
    BPAngle *bp_angles = nullptr;  // [n_angles]
    std::vector<size_t> list_offset;
    // host-array pipeline (tsp_project with TSP_MEM_HOST): copy-in / copy-out streams
    cudaStream_t s_in = nullptr, s_out = nullptr;
    // private stream-ordered memory pool of this (projector, device): per-call scratch and host-array staging
    cudaMemPool_t pool = nullptr;
    bool owns_pool = true;      // false: the pool belongs to the projector this one is a sub-projector of
    DeviceState *pool_st = nullptr;  // ... and this is that projector's state (its release threshold is the one that counts)
    size_t pool_keep = 0;       // bytes this projector wants cached between calls (the threshold adds slack, see pool_keep_at_least)
    size_t pool_extra = 0;      // owner only: what its sub-projectors want on top (sum of their pool_keep)
    size_t pool_keep_base = 0;  // the part kept for device-array calls (transposed-volume scratch)
};

}  // namespace tsp

struct tsp_projector {
    tsp_projector *pool_owner = nullptr;  // sub-projectors allocate from their owner's memory pool
    tsp_geometry g;
    std::vector<double> vectors;
    double sigma[3];
    std::vector<tsp::FPAngle> fp_angles;
    std::vector<int> march_axis;
    std::vector<tsp::FPGroup> groups;
    std::vector<tsp::BPAngle> bp_angles;
    std::map<int, tsp::DeviceState> dev;
    std::mutex mu;
    // introspection counters: written by concurrent callers, hence atomic
    std::atomic<int64_t> launches{0};
    std::atomic<int> bp_uses_tma{0};
    std::atomic<int> fp_uses_transpose{0};
    std::atomic<int> fp_uses_tma{0};
    // Host-array pipeline: the problem cut into sub-problems (BP: z-slabs of the volume with the
    // detector rows their cone shadow covers; FP: detector row blocks), each with its own
    // sub-projector, so that H2D / D2H of one chunk overlaps the kernels of another.
    struct HostChunk {
        tsp_projector *sub = nullptr;
        int z0 = 0, z1 = 0, v0 = 0, v1 = 0;
    };
    std::vector<HostChunk> host_bp, host_fp;
    bool host_planned = false;
    // Supersampling through the staged kernels: sub-projectors on the refined geometry (z-slabs of the fine volume for
    // VoxelSuperSampling, row blocks of the fine detector for DetectorSuperSampling); z0 / z1 / v0 / v1 are COARSE indices.
    std::vector<HostChunk> ss_bp, ss_fp;
    bool ss_planned = false;
    std::atomic<int> host_pipelined{0};  // last host-array call ran the chunked pipeline
    std::atomic<int> host_ring{0};       // ... out of the bounded ring of chunk buffers (device memory budget exceeded)
    std::atomic<int> host_devices{0};    // ... on this many devices
};
