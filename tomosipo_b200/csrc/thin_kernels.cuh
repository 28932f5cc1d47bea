// Thin problems (BASELINE cfg 5: learned primal-dual slabs, 1 x N x N volume, one
// detector row, a batch of 16): the tiled kernels waste most of their threads
// on such shapes and the reference issues one tiny projector call per batch
// element (tomosipo/torch_support.py:49-53,70-74).  These kernels fold the
// whole batch into the grid (one launch per angle group / per backprojection)
// and map threads to what a slab has plenty of: detector columns x angles (FP)
// and in-plane voxels (BP).  Arithmetic = fp_kernel / bp_kernel.
#pragma once
#include "bp_kernels.cuh"
#include "fp_kernels.cuh"

namespace tsp {

constexpr int THIN_FP_ANGLES = 8;   // angles per CTA (threadIdx.y)
constexpr int THIN_BP_BATCH = 32;   // angles set up per block barrier
constexpr int THIN_MAX = 4;         // "thin" = at most this many detector rows (FP) / z slices (BP)

// grid: (det_u tiles of 32, angle tiles of 8, batch * det_v)
template <bool CONE>
__global__ void __launch_bounds__(32 * THIN_FP_ANGLES) fp_thin_kernel(const FPArgs P0, int n_list, size_t vol_bstride,
                                                                     size_t proj_bstride)
{
    const int b = blockIdx.z / P0.det_v, iv = blockIdx.z % P0.det_v;
    FPArgs P = P0;
    P.vol = P0.vol + (size_t)b * vol_bstride;
    P.proj = P0.proj + (size_t)b * proj_bstride;
    if (P.epi_mul) { P.epi_mul += (size_t)b * proj_bstride; P.epi_sub += (size_t)b * proj_bstride; }
    const int ai = blockIdx.y * THIN_FP_ANGLES + threadIdx.y;
    const int a = P.list[min(ai, n_list - 1)];
    const FPAngle g = P.angles[a];
    const int iu = blockIdx.x * 32 + threadIdx.x;
    const bool live = (ai < n_list) && (iu < P.det_u);
    const double cu = (double)iu + 0.5, cv = (double)iv + 0.5;
    const double pm = g.d0[0] + cu * g.u[0] + cv * g.v[0];
    const double pp = g.d0[1] + cu * g.u[1] + cv * g.v[1];
    const double pq = g.d0[2] + cu * g.u[2] + cv * g.v[2];
    float sum;
    if (CONE) sum = march_ray(P, live, pm - g.o[0], pp - g.o[1], pq - g.o[2], g.o[0], g.o[1], g.o[2]);
    else sum = march_ray(P, live, g.o[0], g.o[1], g.o[2], pm, pp, pq);
    if (live) fp_store(P, ((size_t)iv * P.n_angles + a) * P.det_u + iu, sum);
}

// grid: (x tiles of 32, y tiles of 8, batch); every thread owns one (x, y) column of <= THIN_MAX voxels
template <bool CONE>
__global__ void __launch_bounds__(BP_THREADS) bp_thin_kernel(const BPArgs P0, size_t vol_bstride, size_t proj_bstride)
{
    __shared__ float loc[THIN_BP_BATCH][16];  // au[3] bu | av[3] bv | ad[3] bd | weight
    BPArgs P = P0;
    P.vol = P0.vol + (size_t)blockIdx.z * vol_bstride;
    P.proj = P0.proj + (size_t)blockIdx.z * proj_bstride;
    if (P.epi_mul) P.epi_mul += (size_t)blockIdx.z * vol_bstride;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * BP_TX + tx;
    const int x0 = blockIdx.x * BP_TX, y0 = blockIdx.y * BP_TY;
    const int x1 = min(x0 + BP_TX, P.nx) - 1, y1 = min(y0 + BP_TY, P.ny) - 1;
    // voxel coordinates relative to the tile centre keep the fp32 affine maps accurate to ~1e-6 pixel
    const double xc = 0.5 * (x0 + x1) + 0.5 - 0.5 * P.nx;
    const double yc = 0.5 * (y0 + y1) + 0.5 - 0.5 * P.ny;
    const double zc = 0.0;
    const int x = x0 + tx, y = y0 + ty;
    const float dx = (float)((double)x + 0.5 - 0.5 * P.nx - xc);
    const float dy = (float)((double)y + 0.5 - 0.5 * P.ny - yc);
    const float dz0 = (float)(0.5 - 0.5 * P.nz);
    const bool in_xy = (x < P.nx) && (y < P.ny);
    const size_t row_pitch = (size_t)P.n_angles * P.det_u;

    float acc[THIN_MAX];
#pragma unroll
    for (int i = 0; i < THIN_MAX; ++i) acc[i] = 0.0f;

    for (int a0 = 0; a0 < P.n_angles; a0 += THIN_BP_BATCH) {
        const int na = min(THIN_BP_BATCH, P.n_angles - a0);
        __syncthreads();
        if (tid < na) {
            const BPAngle *ang = P.angles + a0 + tid;
            const double den_c = ang->dn[0] * xc + ang->dn[1] * yc + ang->dn[2] * zc + ang->dn[3];
            const double nu_c = ang->nu[0] * xc + ang->nu[1] * yc + ang->nu[2] * zc + ang->nu[3];
            const double nv_c = ang->nv[0] * xc + ang->nv[1] * yc + ang->nv[2] * zc + ang->nv[3];
            float *L = loc[tid];
            // texel-centre convention -> index coordinates: subtract 0.5 (x den)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                L[i] = (float)(ang->nu[i] - 0.5 * ang->dn[i]);
                L[4 + i] = (float)(ang->nv[i] - 0.5 * ang->dn[i]);
                L[8 + i] = (float)ang->dn[i];
            }
            L[3] = (float)(nu_c - 0.5 * den_c);
            L[7] = (float)(nv_c - 0.5 * den_c);
            L[11] = (float)den_c;
            L[12] = (float)ang->weight;
        }
        __syncthreads();
        if (!in_xy) continue;
        for (int j = 0; j < na; ++j) {
            const float *L = loc[j];
            float nu = fmaf(L[0], dx, fmaf(L[1], dy, fmaf(L[2], dz0, L[3])));
            float nv = fmaf(L[4], dx, fmaf(L[5], dy, fmaf(L[6], dz0, L[7])));
            float dn = CONE ? fmaf(L[8], dx, fmaf(L[9], dy, fmaf(L[10], dz0, L[11]))) : 1.0f;
            const float *src = P.proj + (size_t)(a0 + j) * P.det_u;
#pragma unroll
            for (int i = 0; i < THIN_MAX; ++i) {
                if (i < P.nz) {
                    float fu, fv, w2;
                    if (CONE) {
                        const float r = 1.0f / dn;
                        fu = nu * r; fv = nv * r; w2 = r * r;
                    } else {
                        fu = nu; fv = nv; w2 = L[12];
                    }
                    const float val = bp_sample_global(src, P.det_u, P.det_v, row_pitch, fu, fv);
                    if (val != 0.0f) acc[i] = fmaf(w2, val, acc[i]);
                    nu += L[2]; nv += L[6];
                    if (CONE) dn += L[10];
                }
            }
        }
    }
    if (in_xy) {
#pragma unroll
        for (int i = 0; i < THIN_MAX; ++i)
            if (i < P.nz) bp_store_one(P, ((size_t)i * P.ny + y) * P.nx + x, acc[i] * P.out_scale);
    }
}

}  // namespace tsp
