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
constexpr int THIN_BT = 4;          // batch items per thread in the batched instantiations
#ifndef THIN_BP_UNROLL
#define THIN_BP_UNROLL 4
#endif
constexpr int THIN_BP_ANGLES_IN_FLIGHT = THIN_BP_UNROLL;
#ifndef THIN_BP_UNROLL
#define THIN_BP_UNROLL 4
#endif

// All items of a batch share the geometry: a thread computes the tap offsets and weights of a
// sample once and applies them to BT batch items (BT = 1 or THIN_BT), so the index arithmetic --
// most of the instructions of a bounds-checked bilinear gather -- is amortised over the batch.
// Border rule without branches: the 2 x 2 tap window is shifted into the array
// (c = clamp(i, 0, n - 2)) and the two weights are re-assigned to the shifted positions, taps that
// fall outside get weight 0.  Needs n >= 2 along the contiguous axis (the host checks); an axis of
// length 1 (the single slice of a slab) has one row and one weight.

// Weights of the array elements c and c + 1 for a sample at index coordinate f, where
// c = clamp(floor(f), 0, n - 2): linear interpolation with zeros outside the array is the hat
// function max(0, 1 - |f - j|) of every element j, so the shifted window needs no case analysis.
__device__ __forceinline__ void thin_hat_weights(float f, int c, float &w0, float &w1)
{
    const float x = f - (float)c;
    w0 = fmaxf(0.0f, 1.0f - fabsf(x));
    w1 = fmaxf(0.0f, 1.0f - fabsf(x - 1.0f));
}

// A tap whose weight is exactly 0 (shifted window at a border, lanes outside their own slice interval) still
// multiplies what it loads: 0 * Inf and 0 * NaN are NaN, so a NON-FINITE value in the outermost two rows / columns of
// the array can reach samples that lie within one element outside the array, which ASTRA's border mode and the
// tiled kernels (TSP_NO_THIN=1) would leave finite.  Guarding it was measured twice on the learned-PD
// step (r02 GPU calls 11, 12): loads predicated on the weight 1.18 -> 1.48 ms, FMAs predicated on the weight
// 1.18 -> 1.44 ms - a quarter of the step for a case that needs non-finite border data.  The guard is therefore a
// build option (-DTHIN_STRICT_BORDERS); finite data and non-finite interior data behave identically either way.
__device__ __forceinline__ void thin_fma(float &acc, bool live, float w, float v)
{
#ifdef THIN_STRICT_BORDERS
    if (live) acc = fmaf(w, v, acc);
#else
    (void)live;
    acc = fmaf(w, v, acc);
#endif
}

// grid: (det_u tiles of 32, angle tiles of 8, batch groups * det_v)
template <bool CONE, int BT>
__global__ void __launch_bounds__(32 * THIN_FP_ANGLES) fp_thin_kernel(const FPArgs P, int n_list, int batch,
                                                                     size_t vol_bstride, size_t proj_bstride)
{
    const int bg = blockIdx.z / P.det_v, iv = blockIdx.z % P.det_v;
    const int b0 = bg * BT;
    const int nb = min(BT, batch - b0);
    const int ai = blockIdx.y * THIN_FP_ANGLES + threadIdx.y;
    const int iu = blockIdx.x * 32 + threadIdx.x;
    if (ai >= n_list) return;  // warp-uniform (threadIdx.y is the warp)
    const bool live = iu < P.det_u;
    const int a = P.list[ai];
    const FPAngle g = P.angles[a];
    const double cu = (double)iu + 0.5, cv = (double)iv + 0.5;
    const double pm = g.d0[0] + cu * g.u[0] + cv * g.v[0];
    const double pp = g.d0[1] + cu * g.u[1] + cv * g.v[1];
    const double pq = g.d0[2] + cu * g.u[2] + cv * g.v[2];
    const double dir_m = CONE ? pm - g.o[0] : g.o[0], dir_p = CONE ? pp - g.o[1] : g.o[1], dir_q = CONE ? pq - g.o[2] : g.o[2];
    const double org_m = CONE ? g.o[0] : pm, org_p = CONE ? g.o[1] : pp, org_q = CONE ? g.o[2] : pq;
    // same ray parametrisation as march_ray (fp_kernels.cuh)
    const double inv = 1.0 / dir_m;
    const double a_p = dir_p * inv, a_q = dir_q * inv;
    const float ap = (float)a_p, aq = (float)a_q;
    const float cp = (float)(org_p - a_p * org_m + 0.5 * P.n_p - 0.5);
    const float cq = (float)(org_q - a_q * org_m + 0.5 * P.n_q - 0.5);
    const float t0 = 0.5f - 0.5f * (float)P.n_m;
    const float scale = P.sigma_m * sqrtf(1.0f + ap * ap * P.rp2 + aq * aq * P.rq2);
    float f1, l1, f2, l2;
    k_interval(ap, cp, t0, -1.0f, (float)P.n_p, f1, l1);
    k_interval(aq, cq, t0, -1.0f, (float)P.n_q, f2, l2);
    int k_lo = clamp_f2i(fmaxf(f1, f2) - 1.0f, 0, P.n_m, false);
    int k_hi = clamp_f2i(fminf(l1, l2) + 1.0f, 0, P.n_m, true);  // exclusive
    if (!live) { k_lo = P.n_m; k_hi = 0; }
    // the warp walks the slices together (its lanes read neighbouring voxels of one slice: coalesced);
    // outside its own interval a lane's taps are outside the volume and weigh 0
    const int kA = warp_min_i(k_lo), kD = warp_max_i(k_hi);

    // 32-bit element offsets (the host takes this path only when a whole batch has < 2^31 elements)
    float acc[BT];
    uint32_t ob[BT];
    const float *vol0 = P.vol + (size_t)b0 * vol_bstride;
    asm volatile("" : "+l"(vol0));  // keep the base in a register pair (ptxas otherwise re-derives it per load)
#pragma unroll
    for (int j = 0; j < BT; ++j) {
        acc[j] = 0.0f;
        ob[j] = (uint32_t)min(j, nb - 1) * (uint32_t)vol_bstride;  // surplus items of the last group repeat a valid one
        asm volatile("" : "+r"(ob[j]));
    }
    const uint32_t sm32 = (uint32_t)P.stride_m, sq32 = (uint32_t)P.stride_q;
    const bool one_row = P.n_q < 2;
    const int rmax = max(P.n_q - 2, 0);
    for (int k = kA; k < kD; ++k) {
        const float t = (float)k + t0;
        const float fp = fmaf(ap, t, cp), fq = fmaf(aq, t, cq);  // (float -> int saturates; far-away samples weigh 0)
        const int c = min(max(__float2int_rd(fp), 0), P.n_p - 2);
        const int r = min(max(__float2int_rd(fq), 0), rmax);
        float wp0, wp1, wq0, wq1;
        thin_hat_weights(fp, c, wp0, wp1);
        thin_hat_weights(fq, r, wq0, wq1);
        if (one_row) wq1 = 0.0f;
        const uint32_t off = (uint32_t)k * sm32 + (uint32_t)r * sq32 + (uint32_t)c;
        const float w00 = wq0 * wp0, w01 = wq0 * wp1;
        const bool l00 = w00 != 0.0f, l01 = w01 != 0.0f;
#pragma unroll
        for (int j = 0; j < BT; ++j) {
            const float *s = vol0 + (ob[j] + off);
            const float v0 = __ldg(s), v1 = __ldg(s + 1);
            thin_fma(acc[j], l01, w01, v1);
            thin_fma(acc[j], l00, w00, v0);
        }
        if (wq1 != 0.0f) {
            const float w10 = wq1 * wp0, w11 = wq1 * wp1;
            const bool l10 = w10 != 0.0f, l11 = w11 != 0.0f;
#pragma unroll
            for (int j = 0; j < BT; ++j) {
                const float *s = vol0 + (ob[j] + off + sq32);
                const float v0 = __ldg(s), v1 = __ldg(s + 1);
                thin_fma(acc[j], l11, w11, v1);
                thin_fma(acc[j], l10, w10, v0);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int j = 0; j < BT; ++j)
        if (j < nb) fp_store(P, iv, a, iu, acc[j] * scale, (size_t)(b0 + j) * proj_bstride);
}

// grid: (x tiles of 32, y tiles of 8, batch groups); every thread owns one (x, y) column of
// <= THIN_MAX voxels in BT batch items
template <bool CONE, int BT>
__global__ void __launch_bounds__(BP_THREADS) bp_thin_kernel(const BPArgs P, int batch, size_t vol_bstride, size_t proj_bstride)
{
    __shared__ float loc[THIN_BP_BATCH][16];  // au[3] bu | av[3] bv | ad[3] bd | weight
    const int b0 = blockIdx.z * BT;
    const int nb = min(BT, batch - b0);
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
    const bool one_row = P.det_v < 2;
    const int rmax = max(P.det_v - 2, 0);

    // 32-bit element offsets (the host takes this path only when a whole batch has < 2^31 elements)
    float acc[THIN_MAX][BT];
    uint32_t ob[BT];
    const float *proj0 = P.proj + (size_t)b0 * proj_bstride;
    asm volatile("" : "+l"(proj0));  // keep the base in a register pair (ptxas otherwise re-derives it per load)
    const uint32_t rp32 = (uint32_t)row_pitch;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        ob[b] = (uint32_t)min(b, nb - 1) * (uint32_t)proj_bstride;  // surplus items repeat a valid one
        asm volatile("" : "+r"(ob[b]));
#pragma unroll
        for (int i = 0; i < THIN_MAX; ++i) acc[i][b] = 0.0f;
    }

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

#pragma unroll THIN_BP_ANGLES_IN_FLIGHT  // angles in flight (independent loads of consecutive angles overlap)
        for (int j = 0; j < na; ++j) {
            const float *L = loc[j];
            float nu = fmaf(L[0], dx, fmaf(L[1], dy, fmaf(L[2], dz0, L[3])));
            float nv = fmaf(L[4], dx, fmaf(L[5], dy, fmaf(L[6], dz0, L[7])));
            float dn = CONE ? fmaf(L[8], dx, fmaf(L[9], dy, fmaf(L[10], dz0, L[11]))) : 1.0f;
            const uint32_t aoff = (uint32_t)(a0 + j) * (uint32_t)P.det_u;
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
                    // (float -> int saturates; far-away samples weigh 0)
                    const int c = min(max(__float2int_rd(fu), 0), P.det_u - 2);
                    const int r = min(max(__float2int_rd(fv), 0), rmax);
                    float wu0, wu1, wv0, wv1;
                    thin_hat_weights(fu, c, wu0, wu1);
                    thin_hat_weights(fv, r, wv0, wv1);
                    if (one_row) wv1 = 0.0f;
                    wv0 *= w2; wv1 *= w2;
                    const uint32_t off = aoff + (uint32_t)r * rp32 + (uint32_t)c;
                    const float w00 = wv0 * wu0, w01 = wv0 * wu1;
                    const bool l00 = w00 != 0.0f, l01 = w01 != 0.0f;
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const float *s = proj0 + (ob[b] + off);
                        const float v0 = __ldg(s), v1 = __ldg(s + 1);
                        thin_fma(acc[i][b], l01, w01, v1);
                        thin_fma(acc[i][b], l00, w00, v0);
                    }
                    if (wv1 != 0.0f) {
                        const float w10 = wv1 * wu0, w11 = wv1 * wu1;
                        const bool l10 = w10 != 0.0f, l11 = w11 != 0.0f;
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            const float *s = proj0 + (ob[b] + off + rp32);
                            const float v0 = __ldg(s), v1 = __ldg(s + 1);
                            thin_fma(acc[i][b], l11, w11, v1);
                            thin_fma(acc[i][b], l10, w10, v0);
                        }
                    }
                    nu += L[2]; nv += L[6];
                    if (CONE) dn += L[10];
                }
            }
        }
    }
    if (in_xy) {
#pragma unroll
        for (int b = 0; b < BT; ++b)
            if (b < nb) {
                BPArgs Q = P;
                const size_t off = (size_t)(b0 + b) * vol_bstride;
                Q.vol = P.vol + off;
                if (P.epi_mul) Q.epi_mul = P.epi_mul + off;
#pragma unroll
                for (int i = 0; i < THIN_MAX; ++i)
                    if (i < P.nz) bp_store_one(Q, ((size_t)i * P.ny + y) * P.nx + x, acc[i][b] * P.out_scale);
            }
    }
}

}  // namespace tsp
