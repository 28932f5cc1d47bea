// Forward projection: ray-driven Joseph march with bilinear in-slice
// interpolation (semantics: SURVEY.md B.1; call site tomosipo/astra.py:147-153).
//
// Mapping: one thread per detector pixel, lanes of a warp along det_u, one
// angle per CTA.  All angles of a launch march along the same volume axis and
// read a layout whose in-slice axis `p` is memory-contiguous, so that a warp's
// four bilinear taps fall into one or two 128-byte lines per slice.  Angles
// whose in-slice axes are (y, z) read an (x<->y)-transposed copy of the volume
// made once per call by transpose_xy_kernel.
//
// Every output pixel is written exactly once: the thread walks *all* slices,
// unlike ASTRA's N/4 launches that read-modify-write the projections.
#pragma once
#include "tsp_internal.h"

namespace tsp {

constexpr int FP_BU = 32;  // det_u pixels per CTA (= warp width)
constexpr int FP_BV = 8;   // det_v pixels per CTA

// Multi-GPU row exchange fused into the projector's store (tomosipo_b200/distributed.py, tsp_fp_push): every value of
// detector row v in [lo[q], hi[q]) is also stored into rank q's band buffer - peer memory, NVLink stores - so the
// backprojection of the other ranks can start on this rank's angles without a separate exchange pass.
constexpr int FP_MAX_PEERS = 16;
struct FPPeers {
    float *base[FP_MAX_PEERS];  // rank q's band buffer [hi - lo][all angles][U], offset to this rank's first angle
    int lo[FP_MAX_PEERS], hi[FP_MAX_PEERS];
    long long pitch;            // floats between rows of the band buffers (all angles * U)
    int n;                      // 0: single-GPU store
};

struct FPArgs {
    const float *vol;     // volume in the layout of this group
    long long stride_m;   // elements between consecutive slices
    long long stride_q;   // elements between consecutive q rows (p stride is 1)
    int n_m, n_p, n_q;
    const FPAngle *angles;
    const int *list;      // angle ids of this group
    float *proj;
    int det_u, det_v, n_angles;
    int additive;         // 0: SET, 1: ADD to the output, 2: continue a segmented projection (sum, then epilogue, then SET)
    int det_ss;
    float sigma_m;        // voxel size along the march axis
    float rp2, rq2;       // (sigma_p / sigma_m)^2, (sigma_q / sigma_m)^2
    int offsets_fit_32bit;  // volume has < 2^31 elements: the interior loop may use 32-bit offsets
    // fused SIRT residual (tsp_sirt): when set, the stored value is epi_mul[i] * (value - epi_sub[i])
    const float *epi_sub;
    const float *epi_mul;
    FPPeers peers;
};

// The one place a projection value is written: SET, ADD, or the fused SIRT residual - and, on a multi-GPU job, the
// copies of the value in the band buffers of the ranks whose z-slabs read detector row iv.
// `batch_off`: element offset of a batch item in proj / epi_sub / epi_mul (thin kernels); the peers take batch item 0 only.
__device__ __forceinline__ void fp_store(const FPArgs &P, int iv, int a, int iu, float val, size_t batch_off = 0)
{
    const size_t col = (size_t)a * P.det_u + iu;
    const size_t idx = batch_off + (size_t)iv * P.n_angles * P.det_u + col;
    float *dst = P.proj + idx;
    if (P.additive == 2) val += *dst;  // a later segment of a segmented projection: the sum so far is in the output
    if (P.epi_mul) val = __ldg(P.epi_mul + idx) * (val - __ldg(P.epi_sub + idx));
    *dst = P.additive == 1 ? *dst + val : val;
    for (int q = 0; q < P.peers.n; ++q)
        if (iv >= P.peers.lo[q] && iv < P.peers.hi[q]) P.peers.base[q][(size_t)(iv - P.peers.lo[q]) * P.peers.pitch + col] = val;
}

__device__ __forceinline__ int warp_min_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// k-interval on which  lo < a * (k + t0) + c < hi.  Returns [k_first, k_last]
// as floats (may be empty / infinite); the caller rounds and clamps.
__device__ __forceinline__ void k_interval(float a, float c, float t0, float lo, float hi,
                                           float &k_first, float &k_last)
{
    if (fabsf(a) < 1e-12f) {
        const bool in = (c > lo) && (c < hi);
        k_first = in ? -1e30f : 1e30f;
        k_last = in ? 1e30f : -1e30f;
        return;
    }
    const float inv = 1.0f / a;
    float t1 = (lo - c) * inv, t2 = (hi - c) * inv;
    if (t1 > t2) { const float s = t1; t1 = t2; t2 = s; }
    k_first = t1 - t0;
    k_last = t2 - t0;
}

__device__ __forceinline__ int clamp_f2i(float x, int lo, int hi, bool round_up)
{
    x = fminf(fmaxf(x, (float)lo - 1.0f), (float)hi + 1.0f);
    const int i = round_up ? __float2int_ru(x) : __float2int_rd(x);
    return min(max(i, lo), hi);
}

// Line integral of one ray through the whole volume.  `live` = false lanes
// carry a harmless dummy ray so that warp-uniform loops stay in bounds.
__device__ __forceinline__ float march_ray(const FPArgs &P, bool live, double dir_m, double dir_p,
                                           double dir_q, double org_m, double org_p, double org_q)
{
    const double inv = 1.0 / dir_m;
    const double a_p = dir_p * inv, a_q = dir_q * inv;
    float ap = (float)a_p, aq = (float)a_q;
    // index-space intercepts: f = a * t_k + c, voxel i centred on f == i
    float cp = (float)(org_p - a_p * org_m + 0.5 * P.n_p - 0.5);
    float cq = (float)(org_q - a_q * org_m + 0.5 * P.n_q - 0.5);
    const float t0 = 0.5f - 0.5f * (float)P.n_m;
    const float scale = P.sigma_m * sqrtf(1.0f + ap * ap * P.rp2 + aq * aq * P.rq2);

    // slices on which the ray touches the volume, and on which all four taps
    // are guaranteed in bounds
    int k_lo = P.n_m, k_hi = 0, in_lo = 0, in_hi = P.n_m;
    if (live) {
        float f1, l1, f2, l2;
        k_interval(ap, cp, t0, -1.0f, (float)P.n_p, f1, l1);
        k_interval(aq, cq, t0, -1.0f, (float)P.n_q, f2, l2);
        k_lo = clamp_f2i(fmaxf(f1, f2) - 1.0f, 0, P.n_m, false);
        k_hi = clamp_f2i(fminf(l1, l2) + 1.0f, 0, P.n_m, true);  // exclusive
        k_interval(ap, cp, t0, 1.0f, (float)P.n_p - 2.0f, f1, l1);
        k_interval(aq, cq, t0, 1.0f, (float)P.n_q - 2.0f, f2, l2);
        in_lo = clamp_f2i(fmaxf(f1, f2), 0, P.n_m, true);
        in_hi = clamp_f2i(fminf(l1, l2), -1, P.n_m - 1, false) + 1;  // exclusive
        if (P.n_p < 4 || P.n_q < 4 || in_hi <= in_lo || !P.offsets_fit_32bit) { in_lo = P.n_m; in_hi = 0; }
        if (k_hi <= k_lo) live = false;
    }
    if (!live) {  // dummy: stays on an interior voxel for every slice
        ap = 0.0f; aq = 0.0f; cp = 1.0f; cq = 1.0f;
        k_lo = P.n_m; k_hi = 0; in_lo = 0; in_hi = P.n_m;
    }
    const int kA = warp_min_i(k_lo);
    const int kD = warp_max_i(k_hi);
    if (kD <= kA) return 0.0f;
    int kB = max(warp_max_i(in_lo), kA);
    int kC = min(warp_min_i(in_hi), kD);
    if (kB >= kC) { kB = kD; kC = kD; }

    float acc = 0.0f;
    const float *__restrict__ vol = P.vol;
    const long long sm = P.stride_m, sq = P.stride_q;

    auto careful = [&](int k_begin, int k_end) {
        for (int k = k_begin; k < k_end; ++k) {
            const float t = (float)k + t0;
            const float fp = fmaf(ap, t, cp), fq = fmaf(aq, t, cq);
            const float flp = floorf(fp), flq = floorf(fq);
            const int ip = (int)flp, iq = (int)flq;
            const float wp = fp - flp, wq = fq - flq;
            const bool p0 = (ip >= 0) && (ip < P.n_p), p1 = (ip >= -1) && (ip + 1 < P.n_p);
            const bool q0 = (iq >= 0) && (iq < P.n_q), q1 = (iq >= -1) && (iq + 1 < P.n_q);
            const float *s = vol + (long long)k * sm + (long long)iq * sq + ip;
            const float v00 = (p0 && q0) ? __ldg(s) : 0.0f;
            const float v10 = (p1 && q0) ? __ldg(s + 1) : 0.0f;
            const float v01 = (p0 && q1) ? __ldg(s + sq) : 0.0f;
            const float v11 = (p1 && q1) ? __ldg(s + sq + 1) : 0.0f;
            const float lo = fmaf(wp, v10 - v00, v00);
            const float hi = fmaf(wp, v11 - v01, v01);
            acc += fmaf(wq, hi - lo, lo);
        }
    };

    careful(kA, kB);
    {
        // Interior slices: no bounds checks; floor() by adding 1.5*2^23 with
        // round-down (valid for 0 <= f < 2^22): the sum's low mantissa bits are
        // floor(f).  Element offsets are 32-bit:
        // off = k*sm + iq*sq + ip, with the magic-number bias folded into koff.
        const float MAGIC = 12582912.0f;
        const uint32_t MBITS = 0x4B400000u;
        const uint32_t sq32 = (uint32_t)sq, sm32 = (uint32_t)sm;
        uint32_t koff = (uint32_t)kB * sm32 - MBITS * (sq32 + 1u);
        float t = (float)kB + t0;
#pragma unroll 4
        for (int k = kB; k < kC; ++k) {
            const float fp = fmaf(ap, t, cp), fq = fmaf(aq, t, cq);
            const float rp = __fadd_rd(fp, MAGIC), rq = __fadd_rd(fq, MAGIC);  // round-down add == floor
            const float wp = fp - (rp - MAGIC), wq = fq - (rq - MAGIC);
            const uint32_t off = __float_as_uint(rq) * sq32 + __float_as_uint(rp) + koff;
            const float *s0 = vol + off;
            const float *s1 = vol + (off + sq32);
            const float v00 = __ldg(s0), v10 = __ldg(s0 + 1);
            const float v01 = __ldg(s1), v11 = __ldg(s1 + 1);
            const float lo = fmaf(wp, v10 - v00, v00);
            const float hi = fmaf(wp, v11 - v01, v01);
            acc += fmaf(wq, hi - lo, lo);
            t += 1.0f;
            koff += sm32;
        }
    }
    careful(kC, kD);
    return live ? acc * scale : 0.0f;
}

template <bool CONE, bool SUPERSAMPLE>
__global__ void __launch_bounds__(FP_BU *FP_BV) fp_kernel(const FPArgs P)
{
    // Grid order (x fastest): det_u tile, angle, det_v tile.  CTAs that run
    // together share one det_v tile across many angles, i.e. (for the common
    // geometries) one thin slab of the volume, which then stays L2-resident.
    const int a = P.list[blockIdx.y];
    const FPAngle g = P.angles[a];
    const int iu = blockIdx.x * FP_BU + threadIdx.x;
    const int iv = blockIdx.z * FP_BV + threadIdx.y;
    const bool live = (iu < P.det_u) && (iv < P.det_v);
    const int ss = SUPERSAMPLE ? P.det_ss : 1;

    float sum = 0.0f;
    for (int sv = 0; sv < ss; ++sv) {
        for (int su = 0; su < ss; ++su) {
            const double cu = (double)iu + ((double)su + 0.5) / (double)ss;
            const double cv = (double)iv + ((double)sv + 0.5) / (double)ss;
            const double pm = g.d0[0] + cu * g.u[0] + cv * g.v[0];
            const double pp = g.d0[1] + cu * g.u[1] + cv * g.v[1];
            const double pq = g.d0[2] + cu * g.u[2] + cv * g.v[2];
            if (CONE)
                sum += march_ray(P, live, pm - g.o[0], pp - g.o[1], pq - g.o[2], g.o[0], g.o[1], g.o[2]);
            else
                sum += march_ray(P, live, g.o[0], g.o[1], g.o[2], pm, pp, pq);
        }
    }
    if (SUPERSAMPLE) sum /= (float)(ss * ss);
    if (live) fp_store(P, iv, a, iu, sum);
}

// ---------------------------------------------------------------------------
// Column-coherent variant: when the detector's v vector has no component along
// the marching axis or the contiguous in-slice axis p (every circular geometry:
// v parallel to the rotation axis), all pixels of one detector column share the
// in-slice p coordinate, its weight and its address part for every slice.  One
// thread then walks R consecutive rows of a column together: the p part is
// computed once per slice, each ray only adds its q (row) part and the taps.
constexpr int FP_COLS_R = 4;

__device__ __forceinline__ void careful_range(const FPArgs &P, float ap, float aq, float cp, float cq, float t0,
                                              int k_begin, int k_end, float &acc)
{
    const float *__restrict__ vol = P.vol;
    const long long sm = P.stride_m, sq = P.stride_q;
    for (int k = k_begin; k < k_end; ++k) {
        const float t = (float)k + t0;
        const float fp = fmaf(ap, t, cp), fq = fmaf(aq, t, cq);
        const float flp = floorf(fp), flq = floorf(fq);
        const int ip = (int)flp, iq = (int)flq;
        const float wp = fp - flp, wq = fq - flq;
        const bool p0 = (ip >= 0) && (ip < P.n_p), p1 = (ip >= -1) && (ip + 1 < P.n_p);
        const bool q0 = (iq >= 0) && (iq < P.n_q), q1 = (iq >= -1) && (iq + 1 < P.n_q);
        const float *s = vol + (long long)k * sm + (long long)iq * sq + ip;
        const float v00 = (p0 && q0) ? __ldg(s) : 0.0f;
        const float v10 = (p1 && q0) ? __ldg(s + 1) : 0.0f;
        const float v01 = (p0 && q1) ? __ldg(s + sq) : 0.0f;
        const float v11 = (p1 && q1) ? __ldg(s + sq + 1) : 0.0f;
        const float lo = fmaf(wp, v10 - v00, v00);
        const float hi = fmaf(wp, v11 - v01, v01);
        acc += fmaf(wq, hi - lo, lo);
    }
}

template <bool CONE>
__global__ void __launch_bounds__(FP_BU * FP_BV) fp_cols_kernel(const FPArgs P)
{
    constexpr int R = FP_COLS_R;
    const int a = P.list[blockIdx.y];
    const FPAngle g = P.angles[a];
    const int lane = threadIdx.x;
    const int iu = blockIdx.x * FP_BU + lane;
    const int iv0 = (blockIdx.z * FP_BV + threadIdx.y) * R;
    const bool live_u = iu < P.det_u;
    const float t0 = 0.5f - 0.5f * (float)P.n_m;

    // shared (march, p) part of the column; per-row q part (fp64 set-up)
    const double cu = (double)iu + 0.5, cv0 = (double)iv0 + 0.5;
    const double pm = g.d0[0] + cu * g.u[0] + cv0 * g.v[0];
    const double pp = g.d0[1] + cu * g.u[1] + cv0 * g.v[1];
    const double pq0 = g.d0[2] + cu * g.u[2] + cv0 * g.v[2];
    const double dir_m = CONE ? pm - g.o[0] : g.o[0];
    const double dir_p = CONE ? pp - g.o[1] : g.o[1];
    const double org_m = CONE ? g.o[0] : pm;
    const double org_p = CONE ? g.o[1] : pp;
    const double inv = 1.0 / dir_m;
    const double a_p = dir_p * inv;
    float ap = (float)a_p;
    float cp = (float)(org_p - a_p * org_m + 0.5 * P.n_p - 0.5);

    float aq[R], cq[R], scale[R], acc[R];
    bool live[R];
    int k_lo = P.n_m, k_hi = 0, in_lo = 0, in_hi = P.n_m;
    float pf, pl, pif, pil;  // p-range: touch [pf, pl], interior [pif, pil]
    k_interval(ap, cp, t0, -1.0f, (float)P.n_p, pf, pl);
    k_interval(ap, cp, t0, 1.0f, (float)P.n_p - 2.0f, pif, pil);
    bool any_live = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double pq = pq0 + (double)r * g.v[2];
        const double dir_q = CONE ? pq - g.o[2] : g.o[2];
        const double org_q = CONE ? g.o[2] : pq;
        const double a_q = dir_q * inv;
        aq[r] = (float)a_q;
        cq[r] = (float)(org_q - a_q * org_m + 0.5 * P.n_q - 0.5);
        scale[r] = P.sigma_m * sqrtf(1.0f + ap * ap * P.rp2 + aq[r] * aq[r] * P.rq2);
        acc[r] = 0.0f;
        live[r] = live_u && (iv0 + r < P.det_v);
        if (live[r]) {
            float f2, l2;
            k_interval(aq[r], cq[r], t0, -1.0f, (float)P.n_q, f2, l2);
            const int lo = clamp_f2i(fmaxf(pf, f2) - 1.0f, 0, P.n_m, false);
            const int hi = clamp_f2i(fminf(pl, l2) + 1.0f, 0, P.n_m, true);
            if (hi <= lo) {
                live[r] = false;
            } else {
                k_lo = min(k_lo, lo);
                k_hi = max(k_hi, hi);
                k_interval(aq[r], cq[r], t0, 1.0f, (float)P.n_q - 2.0f, f2, l2);
                int ilo = clamp_f2i(fmaxf(pif, f2), 0, P.n_m, true);
                int ihi = clamp_f2i(fminf(pil, l2), -1, P.n_m - 1, false) + 1;
                if (P.n_p < 4 || P.n_q < 4 || ihi <= ilo || !P.offsets_fit_32bit) { ilo = P.n_m; ihi = 0; }
                in_lo = max(in_lo, ilo);
                in_hi = min(in_hi, ihi);
                any_live = true;
            }
        }
        if (!live[r]) { aq[r] = 0.0f; cq[r] = 1.0f; }  // dummy row: stays on an interior voxel
    }
    if (!any_live) { ap = 0.0f; cp = 1.0f; }          // dummy column
    const int kA = warp_min_i(k_lo);
    const int kD = warp_max_i(k_hi);
    if (kD > kA) {
        int kB = max(warp_max_i(in_lo), kA);
        int kC = min(warp_min_i(in_hi), kD);
        if (kB >= kC) { kB = kD; kC = kD; }
#pragma unroll
        for (int r = 0; r < R; ++r) careful_range(P, ap, aq[r], cp, cq[r], t0, kA, kB, acc[r]);
        {
            const float MAGIC = 12582912.0f;
            const uint32_t MBITS = 0x4B400000u;
            const float *__restrict__ vol = P.vol;
            const uint32_t sq32 = (uint32_t)P.stride_q, sm32 = (uint32_t)P.stride_m;
            uint32_t koff = (uint32_t)kB * sm32 - MBITS * (sq32 + 1u);
            float t = (float)kB + t0;
#pragma unroll 2
            for (int k = kB; k < kC; ++k) {
                const float fp = fmaf(ap, t, cp);
                const float rp = __fadd_rd(fp, MAGIC);
                const float wp = fp - (rp - MAGIC);
                const uint32_t offp = __float_as_uint(rp) + koff;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float fq = fmaf(aq[r], t, cq[r]);
                    const float rq = __fadd_rd(fq, MAGIC);
                    const float wq = fq - (rq - MAGIC);
                    const uint32_t off = __float_as_uint(rq) * sq32 + offp;
                    const float *s0 = vol + off;
                    const float *s1 = vol + (off + sq32);
                    const float v00 = __ldg(s0), v10 = __ldg(s0 + 1);
                    const float v01 = __ldg(s1), v11 = __ldg(s1 + 1);
                    const float lo = fmaf(wp, v10 - v00, v00);
                    const float hi = fmaf(wp, v11 - v01, v01);
                    acc[r] += fmaf(wq, hi - lo, lo);
                }
                t += 1.0f;
                koff += sm32;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) careful_range(P, ap, aq[r], cp, cq[r], t0, kC, kD, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (live_u && iv0 + r < P.det_v) {
            const float val = live[r] ? acc[r] * scale[r] : 0.0f;
            fp_store(P, iv0 + r, a, iu, val);
        }
    }
}

// out[z][x][y] = in[z][y][x]; the rows of `out` have pitch ny_pad >= ny
__global__ void __launch_bounds__(256) transpose_xy_kernel(const float *__restrict__ in,
                                                            float *__restrict__ out, int nx, int ny, int ny_pad)
{
    __shared__ float tile[32][33];
    const size_t plane = (size_t)nx * ny * blockIdx.z;
    const size_t plane_out = (size_t)nx * ny_pad * blockIdx.z;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = x0 + threadIdx.x, y = y0 + r;
        if (x < nx && y < ny) tile[r][threadIdx.x] = in[plane + (size_t)y * nx + x];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int y = y0 + threadIdx.x, x = x0 + r;
        if (x < nx && y < ny) out[plane_out + (size_t)x * ny_pad + y] = tile[threadIdx.x][r];
    }
}

// Detector supersampling through the staged kernels: proj[v][a][u] = mean of the d x d pixels of a d-times finer
// detector (rows v0 .. v0 + nv of the coarse detector are in `fine`, [nv * d][A][U * d]); stored with fp_store, so SET,
// ADD and the fused SIRT residual behave as in the direct kernels.
__global__ void __launch_bounds__(256) fp_pool_kernel(const FPArgs P, const float *__restrict__ fine, int d, int v0, int nv)
{
    const int u = blockIdx.x * 32 + threadIdx.x;
    const int a = blockIdx.y * 8 + threadIdx.y;
    const int v = blockIdx.z;
    if (u >= P.det_u || a >= P.n_angles || v >= nv) return;
    const size_t fu = (size_t)P.det_u * d, frow = fu * P.n_angles;
    float sum = 0.0f;
    for (int sv = 0; sv < d; ++sv) {
        const float *src = fine + ((size_t)(v * d + sv)) * frow + (size_t)a * fu + (size_t)u * d;
        for (int su = 0; su < d; ++su) sum += __ldg(src + su);
    }
    fp_store(P, v0 + v, a, u, sum / (float)(d * d));
}

}  // namespace tsp
