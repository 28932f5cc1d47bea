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

// All items of a batch share the geometry: a thread computes the tap offsets and weights of a
// sample once and applies them to BT batch items (BT = 1 or THIN_BT), so the index arithmetic --
// most of the instructions of a bounds-checked bilinear gather -- is amortised over the batch.
// Out-of-volume taps get weight 0 and a clamped (valid) offset; a row of taps whose weight is 0
// (always the case for one of the two rows of a single-slice slab) is skipped.

// grid: (det_u tiles of 32, angle tiles of 8, batch groups * det_v)
template <bool CONE, int BT>
__global__ void __launch_bounds__(32 * THIN_FP_ANGLES) fp_thin_kernel(const FPArgs P, int n_list, int batch,
                                                                     size_t vol_bstride, size_t proj_bstride)
{
    const int bg = blockIdx.z / P.det_v, iv = blockIdx.z % P.det_v;
    const int b0 = bg * BT;
    const int ai = blockIdx.y * THIN_FP_ANGLES + threadIdx.y;
    const int iu = blockIdx.x * 32 + threadIdx.x;
    if (ai >= n_list || iu >= P.det_u) return;
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
    const int k_lo = clamp_f2i(fmaxf(f1, f2) - 1.0f, 0, P.n_m, false);
    const int k_hi = clamp_f2i(fminf(l1, l2) + 1.0f, 0, P.n_m, true);  // exclusive

    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.0f;
    const float *__restrict__ vol = P.vol + (size_t)b0 * vol_bstride;
    const int nb = min(BT, batch - b0);
    for (int k = k_lo; k < k_hi; ++k) {
        const float t = (float)k + t0;
        const float fp = fmaf(ap, t, cp), fq = fmaf(aq, t, cq);
        const float flp = floorf(fp), flq = floorf(fq);
        const int ip = (int)flp, iq = (int)flq;
        const float wp = fp - flp, wq = fq - flq;
        const float wp0 = (ip >= 0 && ip < P.n_p) ? 1.0f - wp : 0.0f, wp1 = (ip >= -1 && ip + 1 < P.n_p) ? wp : 0.0f;
        const float wq0 = (iq >= 0 && iq < P.n_q) ? 1.0f - wq : 0.0f, wq1 = (iq >= -1 && iq + 1 < P.n_q) ? wq : 0.0f;
        const int c0 = min(max(ip, 0), P.n_p - 1), c1 = min(max(ip + 1, 0), P.n_p - 1);
        const int r0 = min(max(iq, 0), P.n_q - 1), r1 = min(max(iq + 1, 0), P.n_q - 1);
        const float *s0 = vol + (long long)k * P.stride_m + (long long)r0 * P.stride_q;
        const float *s1 = vol + (long long)k * P.stride_m + (long long)r1 * P.stride_q;
        if (wq0 != 0.0f) {
            const float w0 = wq0 * wp0, w1 = wq0 * wp1;
#pragma unroll
            for (int j = 0; j < BT; ++j)
                if (j < nb) acc[j] = fmaf(w0, __ldg(s0 + j * vol_bstride + c0), fmaf(w1, __ldg(s0 + j * vol_bstride + c1), acc[j]));
        }
        if (wq1 != 0.0f) {
            const float w0 = wq1 * wp0, w1 = wq1 * wp1;
#pragma unroll
            for (int j = 0; j < BT; ++j)
                if (j < nb) acc[j] = fmaf(w0, __ldg(s1 + j * vol_bstride + c0), fmaf(w1, __ldg(s1 + j * vol_bstride + c1), acc[j]));
        }
    }
    const size_t pix = ((size_t)iv * P.n_angles + a) * P.det_u + iu;
#pragma unroll
    for (int j = 0; j < BT; ++j)
        if (j < nb) {
            FPArgs Q = P;
            const size_t off = (size_t)(b0 + j) * proj_bstride;
            Q.proj = P.proj + off;
            if (P.epi_mul) { Q.epi_mul = P.epi_mul + off; Q.epi_sub = P.epi_sub + off; }
            fp_store(Q, pix, acc[j] * scale);
        }
}

// grid: (x tiles of 32, y tiles of 8, batch groups); every thread owns one (x, y) column of
// <= THIN_MAX voxels in BT batch items
template <bool CONE, int BT>
__global__ void __launch_bounds__(BP_THREADS) bp_thin_kernel(const BPArgs P, int batch, size_t vol_bstride, size_t proj_bstride)
{
    __shared__ float loc[THIN_BP_BATCH][16];  // au[3] bu | av[3] bv | ad[3] bd | weight
    const int b0 = blockIdx.z * BT;
    const int nb = min(BT, batch - b0);
    const float *__restrict__ proj = P.proj + (size_t)b0 * proj_bstride;
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

    float acc[THIN_MAX][BT];
#pragma unroll
    for (int i = 0; i < THIN_MAX; ++i)
#pragma unroll
        for (int j = 0; j < BT; ++j) acc[i][j] = 0.0f;

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
            const float *src = proj + (size_t)(a0 + j) * P.det_u;
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
                    if (fu > -1.0f && fu < (float)P.det_u && fv > -1.0f && fv < (float)P.det_v) {
                        const float flu = floorf(fu), flv = floorf(fv);
                        const int iu = (int)flu, iv = (int)flv;
                        const float wu = fu - flu, wv = fv - flv;
                        const float wu0 = iu >= 0 ? 1.0f - wu : 0.0f, wu1 = iu + 1 < P.det_u ? wu : 0.0f;
                        const float wv0 = (iv >= 0 ? 1.0f - wv : 0.0f) * w2, wv1 = (iv + 1 < P.det_v ? wv : 0.0f) * w2;
                        const int c0 = max(iu, 0), c1 = min(iu + 1, P.det_u - 1);
                        const float *s0 = src + (size_t)max(iv, 0) * row_pitch;
                        const float *s1 = src + (size_t)min(iv + 1, P.det_v - 1) * row_pitch;
                        if (wv0 != 0.0f) {
                            const float w0 = wv0 * wu0, w1 = wv0 * wu1;
#pragma unroll
                            for (int b = 0; b < BT; ++b)
                                if (b < nb)
                                    acc[i][b] = fmaf(w0, __ldg(s0 + b * proj_bstride + c0),
                                                     fmaf(w1, __ldg(s0 + b * proj_bstride + c1), acc[i][b]));
                        }
                        if (wv1 != 0.0f) {
                            const float w0 = wv1 * wu0, w1 = wv1 * wu1;
#pragma unroll
                            for (int b = 0; b < BT; ++b)
                                if (b < nb)
                                    acc[i][b] = fmaf(w0, __ldg(s1 + b * proj_bstride + c0),
                                                     fmaf(w1, __ldg(s1 + b * proj_bstride + c1), acc[i][b]));
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
