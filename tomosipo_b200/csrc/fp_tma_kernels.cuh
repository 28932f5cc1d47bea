// Forward projection, TMA-staged variant (default when the volume layout meets
// TMA's alignment rules).  Same arithmetic as fp_kernel / fp_cols_kernel
// (Joseph march), different data path:
//
//   * a CTA owns a 32 (u) x 16 (v) detector tile for a PAIR of neighbouring
//     angles and marches through the slices of the volume together;
//   * per slice, the producer warp bounds the tile's footprint on that slice (the
//     bounding box of the 2 x 4 corner rays, exact for projective maps) and has
//     the TMA engine copy that box of the slice into a shared-memory ring stage;
//     voxels outside the volume arrive as zeros, which is the projector's border
//     rule - so there is no bounds-checked "careful" loop at all;
//   * the 8 consumer warps (16 columns x 2 angles per warp, 4 detector rows per thread) take their
//     four bilinear taps from shared memory (bank-granular, not line-granular
//     like the L1 path: ncu showed the LDG kernel bound by L1 wavefronts, ~2
//     128-byte lines per tap instruction);
//   * two neighbouring angles share the staged box (their footprints differ by a
//     few voxels), which halves the L2 -> shared-memory traffic per sample.
//
// A slice whose footprint does not fit the box (unusual geometries) is flagged by
// the producer and sampled with bounds-checked global loads instead.
#pragma once
#include "bp_kernels.cuh"  // mbarrier / TMA helpers
#include "fp_kernels.cuh"

namespace tsp {

constexpr int FPT_TU = 32;         // det_u pixels per CTA
// det_v rows per thread: template parameter R (4 or 8); a CTA covers 4 * R rows
__host__ __device__ constexpr int fpt_min_ctas(int r) { return r >= 8 ? 2 : 3; }
constexpr int FPT_CONSUMERS = 256;
constexpr int FPT_THREADS = FPT_CONSUMERS + 32;

struct FPTmaArgs {
    FPArgs a;               // volume (for the fallback path), dims, angle table, list, output
    const int *pairs;       // angle pairs of this launch: {a, b} with b = -1 for a single angle
    // Two staged-box variants that differ in row pitch (= box width): a warp's lanes sit on a few
    // neighbouring rows of the box; which pitch residue (mod 32 banks) keeps them on distinct
    // banks depends on whether column and row move together along the lanes
    // (scratch/bank_sim_fp2.py).  The producer picks the variant per CTA; tmap[v] has box_w[v].
    int box_w[2];           // elements, multiples of 4
    int box_h;
    // SPS consecutive slices share one ring stage (one barrier round trip, one control word, one TMA issue per
    // stage: the hand-off is ~26 % of the instructions of a one-slice stage).  The staged box is then
    // (box_w, SPS, box_h) or (box_w, box_h, SPS) elements and serves the footprints of all its slices.
    int sps;                // 1 or 2
    uint32_t row_stride4[2];  // bytes between consecutive q rows of a staged slice
    uint32_t slice_off4[2];   // bytes between the slices of a stage
    uint32_t magic_off[2];  // -4 * MAGIC_BITS * (row stride in words + 1) mod 2^32 (run-time on purpose, see BPArgs)
    int march_is_middle;    // tensor coordinates are (p, k, q) if set, (p, q, k) otherwise
    int stages;
    uint32_t stage_bytes;   // max box bytes rounded up to 128
    // Segment [m_begin, m_end) of the marching axis this launch integrates over.  All CTAs of one detector row tile read
    // the same slab of the volume (every in-plane position x the tile's q band): 42 MB at 512^3, where it lives in L2
    // across the angles (97.8 % hit rate), but 168 MB at 1024^3, where every angle pair re-read it from DRAM (39 % hits,
    // 2.5 TB per launch, the kernel at 6.2 TB/s = HBM-bound; r02 GPU call 38).  Large volumes are therefore projected
    // segment by segment, the later segments accumulating into the output (FPArgs::additive == 2).
    int m_begin, m_end;
};

struct FPRay {
    float ap, cp, aq, cq;
};

// Ray through the centre-relative detector coordinate (cu, cv) of angle g, as
// index-space lines  p(k) = ap * (k + t0) + cp,  q(k) = aq * (k + t0) + cq.
template <bool CONE>
__device__ __forceinline__ FPRay fpt_ray(const FPAngle &g, double cu, double cv, int n_p, int n_q)
{
    const double pm = g.d0[0] + cu * g.u[0] + cv * g.v[0];
    const double pp = g.d0[1] + cu * g.u[1] + cv * g.v[1];
    const double pq = g.d0[2] + cu * g.u[2] + cv * g.v[2];
    const double dir_m = CONE ? pm - g.o[0] : g.o[0];
    const double dir_p = CONE ? pp - g.o[1] : g.o[1];
    const double dir_q = CONE ? pq - g.o[2] : g.o[2];
    const double org_m = CONE ? g.o[0] : pm;
    const double org_p = CONE ? g.o[1] : pp;
    const double org_q = CONE ? g.o[2] : pq;
    const double inv = 1.0 / dir_m;
    const double a_p = dir_p * inv, a_q = dir_q * inv;
    FPRay r;
    r.ap = (float)a_p;
    r.aq = (float)a_q;
    r.cp = (float)(org_p - a_p * org_m + 0.5 * n_p - 0.5);
    r.cq = (float)(org_q - a_q * org_m + 0.5 * n_q - 0.5);
    return r;
}

template <int OFF>
__device__ __forceinline__ float fpt_lds(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

__device__ __forceinline__ void fpt_tma_box(uint32_t dst, const void *tmap, int c0, int c1, int c2, uint32_t bar)
{
    tma_load_box_3d(dst, tmap, c0, c1, c2, bar);
}

// Bounding box of the 8 corner rays on slice k: [pmin, pmax] x [qmin, qmax] (index coordinates).
__device__ __forceinline__ void fpt_slice_box(const FPRay (&c)[8], float t, float &pmin, float &pmax, float &qmin,
                                              float &qmax)
{
    pmin = qmin = 3.0e38f;
    pmax = qmax = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float p = fmaf(c[i].ap, t, c[i].cp), q = fmaf(c[i].aq, t, c[i].cq);
        pmin = fminf(pmin, p); pmax = fmaxf(pmax, p);
        qmin = fminf(qmin, q); qmax = fmaxf(qmax, q);
    }
    // consumers evaluate interior rays in fp32: allow for their rounding
    pmin -= 0.01f; qmin -= 0.01f; pmax += 0.01f; qmax += 0.01f;
}

// Consumer march over the hull slices [kA, kD) for pitch variant V (box_w[V] stays in a
// uniform register, so the second tap row is addressed as [a0 + UR]).
template <bool COLS, int V, int R, int SPS>
__device__ __forceinline__ void fpt_consume(const FPTmaArgs &A, int kA, int kD, float t0, uint32_t ctrl, uint32_t full,
                                            uint32_t empty, int lane, const float2 *ap2, const float2 *cp2,
                                            const float2 (&aq2)[R / 2], const float2 (&cq2)[R / 2], float2 (&acc2)[R / 2])
{
    // rows are kept as packed pairs (.x, .y) = (row 2h, row 2h + 1); with COLS, ap2[0].x / cp2[0].x hold
    // the column's shared p line
    const FPArgs &P = A.a;
    const float MAGIC = 12582912.0f;
    const uint32_t rs4 = A.row_stride4[V], so4 = A.slice_off4[V];
    constexpr int sps = SPS;
    float t = (float)kA + t0;
    int s = 0;
    uint32_t parity = 0u;
    for (int k = kA; k < kD; k += sps) {
        mbar_wait(full + 8u * s, parity);
        uint32_t sb, fit;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sb), "=r"(fit) : "r"(ctrl + 8u * s) : "memory");
        if (fit) {
            // Two detector rows per step with Blackwell's packed fp32x2 arithmetic (FFMA2 / FADD2: two
            // independent fp32 operations per issue slot - the kernel is issue-bound): 10.5 instead
            // of 16 instructions per sample.  Every pair (.x, .y) = (row r, row r + 1).
            const float2 M2 = make_float2(MAGIC, MAGIC), NEG1 = make_float2(-1.0f, -1.0f);
#pragma unroll
            for (int j = 0; j < sps; ++j) {  // a slice past the hull's end samples zeros (its box is staged all the same)
                const float tj = t + (float)j;
                const float2 t2 = make_float2(tj, tj);
                float2 wp2 = make_float2(0.0f, 0.0f);
                uint32_t offp0 = 0u, offp1 = 0u;
                if (COLS) {
                    const float fp = fmaf(ap2[0].x, tj, cp2[0].x);
                    const float rp = __fadd_rd(fp, MAGIC);
                    const float wp = fp - (rp - MAGIC);
                    wp2 = make_float2(wp, wp);
                    offp0 = offp1 = __float_as_uint(rp) * 4u + sb;
                }
#pragma unroll
                for (int h = 0; h < R / 2; ++h) {
                    if (!COLS) {
                        const float2 fp2 = __ffma2_rn(ap2[COLS ? 0 : h], t2, cp2[COLS ? 0 : h]);
                        const float2 rp2 = __fadd2_rd(fp2, M2);
                        wp2 = __fadd2_rn(fp2, __ffma2_rn(rp2, NEG1, M2));
                        offp0 = __float_as_uint(rp2.x) * 4u + sb;
                        offp1 = __float_as_uint(rp2.y) * 4u + sb;
                    }
                    const float2 fq2 = __ffma2_rn(aq2[h], t2, cq2[h]);
                    const float2 rq2 = __fadd2_rd(fq2, M2);                          // round-down add == floor
                    const float2 wq2 = __fadd2_rn(fq2, __ffma2_rn(rq2, NEG1, M2));  // fq - (rq - M)
                    const uint32_t a0 = __float_as_uint(rq2.x) * rs4 + offp0;
                    const uint32_t b0 = __float_as_uint(rq2.y) * rs4 + offp1;
                    const uint32_t a1 = a0 + rs4, b1 = b0 + rs4;
                    const float2 v00 = make_float2(fpt_lds<0>(a0), fpt_lds<0>(b0));
                    const float2 v10 = make_float2(fpt_lds<4>(a0), fpt_lds<4>(b0));
                    const float2 v01 = make_float2(fpt_lds<0>(a1), fpt_lds<0>(b1));
                    const float2 v11 = make_float2(fpt_lds<4>(a1), fpt_lds<4>(b1));
                    const float2 lo = __ffma2_rn(wp2, __ffma2_rn(v00, NEG1, v10), v00);
                    const float2 hi = __ffma2_rn(wp2, __ffma2_rn(v01, NEG1, v11), v01);
                    const float2 val = __ffma2_rn(wq2, __ffma2_rn(lo, NEG1, hi), lo);
                    acc2[h] = __fadd2_rn(acc2[h], val);
                }
                if (SPS > 1) sb += so4;
            }
        } else {
            const int k_end = min(k + sps, P.n_m);
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {
                careful_range(P, COLS ? ap2[0].x : ap2[COLS ? 0 : h].x, aq2[h].x, COLS ? cp2[0].x : cp2[COLS ? 0 : h].x,
                              cq2[h].x, t0, k, k_end, acc2[h].x);
                careful_range(P, COLS ? ap2[0].x : ap2[COLS ? 0 : h].y, aq2[h].y, COLS ? cp2[0].x : cp2[COLS ? 0 : h].y,
                              cq2[h].y, t0, k, k_end, acc2[h].y);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8u * s);
        if (++s == A.stages) { s = 0; parity ^= 1u; }
        t += (float)sps;
    }
}

template <bool CONE, bool COLS, int R, int SPS>
__global__ void __launch_bounds__(FPT_THREADS, fpt_min_ctas(R))
fp_tma_kernel(const FPTmaArgs A, const __grid_constant__ TensorMapPair tmaps)
{
    const TensorMapBlob *tmap = tmaps.m;
    constexpr int FPT_TV = 4 * R;
    const FPArgs &P = A.a;
    extern __shared__ __align__(128) unsigned char fpt_smem[];
    unsigned char *base = fpt_smem + ((128u - (smem_u32(fpt_smem) & 127u)) & 127u);
    const uint32_t bufs = smem_u32(base);
    const uint32_t ctrl = bufs + (uint32_t)A.stages * A.stage_bytes;  // per stage: {uint32 sb, uint32 fit}
    const uint32_t full = ctrl + 8u * A.stages;
    const uint32_t empty = full + 8u * A.stages;
    int *hull = reinterpret_cast<int *>(base + (size_t)A.stages * (A.stage_bytes + 24u));  // {kA, kD, variant}

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair_a = A.pairs[2 * blockIdx.y], pair_b = A.pairs[2 * blockIdx.y + 1];
    const int n_in_pair = pair_b >= 0 ? 2 : 1;
    const int u0 = blockIdx.x * FPT_TU, v0 = blockIdx.z * FPT_TV;
    const int u1 = min(u0 + FPT_TU, P.det_u) - 1, v1 = min(v0 + FPT_TV, P.det_v) - 1;
    const float t0 = 0.5f - 0.5f * (float)P.n_m;

    if (tid == 0) {
        for (int s = 0; s < A.stages; ++s) {
            mbar_init(full + 8u * s, 1);
            mbar_init(empty + 8u * s, FPT_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == FPT_CONSUMERS / 32) {
        // ------------------------------------------------------------ producer
        // lanes 0..7 each trace one corner ray (angle slot = lane >> 2), then everybody gets all 8
        FPRay mine;
        {
            const int slot = min((lane >> 2) & 1, n_in_pair - 1);
            const FPAngle g = P.angles[slot ? pair_b : pair_a];
            const double cu = (double)((lane & 1) ? u1 : u0) + 0.5;
            const double cv = (double)((lane & 2) ? v1 : v0) + 0.5;
            mine = fpt_ray<CONE>(g, cu, cv, P.n_p, P.n_q);
        }
        FPRay c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            c[i].ap = __shfl_sync(0xffffffffu, mine.ap, i);
            c[i].cp = __shfl_sync(0xffffffffu, mine.cp, i);
            c[i].aq = __shfl_sync(0xffffffffu, mine.aq, i);
            c[i].cq = __shfl_sync(0xffffffffu, mine.cq, i);
        }
        // hull of the slices on which the tile's footprint touches the volume
        int kA = P.n_m, kD = 0;
        for (int k0 = 0; k0 < P.n_m; k0 += 32) {
            const int k = k0 + lane;
            float pmin, pmax, qmin, qmax;
            fpt_slice_box(c, (float)k + t0, pmin, pmax, qmin, qmax);
            const bool needed = (k < P.n_m) && (pmax > -1.0f) && (pmin < (float)P.n_p) && (qmax > -1.0f) &&
                                (qmin < (float)P.n_q);
            const unsigned m = __ballot_sync(0xffffffffu, needed);
            if (m) {
                kA = min(kA, k0 + __ffs(m) - 1);
                kD = max(kD, k0 + 32 - __clz(m));
            }
        }
        kA = max(kA, A.m_begin);
        kD = min(kD, A.m_end);
        // a stage holds SPS consecutive slices: in a segmented projection it must not reach into the next segment
        // (segment lengths are multiples of SPS; slices below the hull add nothing)
        if (A.m_end - A.m_begin < P.n_m) kA = A.m_begin + (kA - A.m_begin) / SPS * SPS;
        if (kD <= kA) { kA = 0; kD = 0; }
        // pitch variant: do column and row move together along u (at the middle of the hull)?
        int variant = 0;
        {
            const float tm = 0.5f * (float)(kA + kD) + t0;
            const float dp = (fmaf(c[1].ap, tm, c[1].cp) + fmaf(c[3].ap, tm, c[3].cp)) -
                             (fmaf(c[0].ap, tm, c[0].cp) + fmaf(c[2].ap, tm, c[2].cp));
            const float dq = (fmaf(c[1].aq, tm, c[1].cq) + fmaf(c[3].aq, tm, c[3].cq)) -
                             (fmaf(c[0].aq, tm, c[0].cq) + fmaf(c[2].aq, tm, c[2].cq));
            variant = (dp * dq < 0.0f) ? 1 : 0;
        }
        if (lane == 0) { hull[0] = kA; hull[1] = kD; hull[2] = variant; }
        __syncthreads();

        int s = 0;
        uint32_t parity = 1u;
        const int bw = A.box_w[variant];
        constexpr int sps = SPS;
        const int rs = (int)(A.row_stride4[variant] >> 2);  // words between q rows of a staged slice
        const uint32_t moff = A.magic_off[variant];
        const uint32_t box_bytes = (uint32_t)bw * (uint32_t)A.box_h * 4u * (uint32_t)sps;
        const TensorMapBlob *tm = tmap + variant;
        for (int k0 = kA; k0 < kD; k0 += 32 * sps) {
            const int k = k0 + lane * sps;  // this lane bounds the stage that starts at slice k
            float pmin, pmax, qmin, qmax;
            fpt_slice_box(c, (float)k + t0, pmin, pmax, qmin, qmax);
            if (sps == 2) {  // union with the footprint on slice k + 1
                float pmin1, pmax1, qmin1, qmax1;
                fpt_slice_box(c, (float)(k + 1) + t0, pmin1, pmax1, qmin1, qmax1);
                pmin = fminf(pmin, pmin1); pmax = fmaxf(pmax, pmax1);
                qmin = fminf(qmin, qmin1); qmax = fmaxf(qmax, qmax1);
            }
            // clamp far-away boxes so that the float -> int conversions are safe
            pmin = fmaxf(pmin, -1.0e6f); qmin = fmaxf(qmin, -1.0e6f);
            pmax = fminf(pmax, 1.0e6f); qmax = fminf(qmax, 1.0e6f);
            int p0 = __float2int_rd(pmin);
            p0 -= ((p0 % 4) + 4) % 4;  // TMA: innermost coordinate on a 16-byte boundary
            const int q0 = __float2int_rd(qmin);
            const int fit = (__float2int_rd(pmax) + 2 - p0 <= bw) && (__float2int_rd(qmax) + 2 - q0 <= A.box_h) &&
                            (pmax >= pmin) && (qmax >= qmin);
            const int nj = min(32, (kD - k0 + sps - 1) / sps);
            for (int j = 0; j < nj; ++j) {
                const int p0j = __shfl_sync(0xffffffffu, p0, j);
                const int q0j = __shfl_sync(0xffffffffu, q0, j);
                const int fitj = __shfl_sync(0xffffffffu, fit, j);
                mbar_wait(empty + 8u * s, parity);
                if (lane == 0) {
                    const uint32_t dst = bufs + (uint32_t)s * A.stage_bytes;
                    const uint32_t sb = dst - 4u * (uint32_t)(q0j * rs + p0j) + moff;
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ctrl + 8u * s), "r"(sb), "r"((uint32_t)fitj)
                                 : "memory");
                    if (fitj) {
                        mbar_arrive_expect_tx(full + 8u * s, box_bytes);
                        if (A.march_is_middle) fpt_tma_box(dst, tm, p0j, k0 + j * sps, q0j, full + 8u * s);
                        else fpt_tma_box(dst, tm, p0j, q0j, k0 + j * sps, full + 8u * s);
                    } else {
                        mbar_arrive(full + 8u * s);
                    }
                }
                __syncwarp();
                if (++s == A.stages) { s = 0; parity ^= 1u; }
            }
        }
        return;
    }

    // --------------------------------------------------------------- consumers
    // A warp = 16 detector columns x the 2 angles of the pair (lanes 0-15 / 16-31): the two
    // half-warps sample nearly the same voxels, so a tap instruction touches <= ~32 distinct
    // words (32 consecutive columns of one angle span 40-60 words: always >= 2 wavefronts).
    const int slot = lane >> 4;
    const bool slot_live = slot < n_in_pair;
    const int a = (slot_live && slot) ? pair_b : pair_a;
    const FPAngle g = P.angles[a];
    const int iu = u0 + (warp & 1) * 16 + (lane & 15);
    const int iv0 = v0 + (warp >> 1) * R;
    // out-of-detector lanes / rows shadow the tile's last pixel: their taps stay inside the staged box
    const double cu = (double)min(iu, u1) + 0.5;

    float2 ap2[COLS ? 1 : R / 2], cp2[COLS ? 1 : R / 2], aq2[R / 2], cq2[R / 2], acc2[R / 2];
#pragma unroll
    for (int h = 0; h < R / 2; ++h) {
        const FPRay r0 = fpt_ray<CONE>(g, cu, (double)min(iv0 + 2 * h, v1) + 0.5, P.n_p, P.n_q);
        const FPRay r1 = fpt_ray<CONE>(g, cu, (double)min(iv0 + 2 * h + 1, v1) + 0.5, P.n_p, P.n_q);
        if (!COLS || h == 0) {
            ap2[COLS ? 0 : h] = make_float2(r0.ap, r1.ap);
            cp2[COLS ? 0 : h] = make_float2(r0.cp, r1.cp);
        }
        aq2[h] = make_float2(r0.aq, r1.aq);
        cq2[h] = make_float2(r0.cq, r1.cq);
        acc2[h] = make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    const int kA = hull[0], kD = hull[1], variant = hull[2];
    if (variant) fpt_consume<COLS, 1, R, SPS>(A, kA, kD, t0, ctrl, full, empty, lane, ap2, cp2, aq2, cq2, acc2);
    else fpt_consume<COLS, 0, R, SPS>(A, kA, kD, t0, ctrl, full, empty, lane, ap2, cp2, aq2, cq2, acc2);

    if (slot_live && iu < P.det_u) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (iv0 + r < P.det_v) {
                const float apr = COLS ? ap2[0].x : ((r & 1) ? ap2[COLS ? 0 : r / 2].y : ap2[COLS ? 0 : r / 2].x);
                const float aqr = (r & 1) ? aq2[r / 2].y : aq2[r / 2].x;
                const float scale = P.sigma_m * sqrtf(1.0f + apr * apr * P.rp2 + aqr * aqr * P.rq2);
                const float val = ((r & 1) ? acc2[r / 2].y : acc2[r / 2].x) * scale;
                fp_store(P, iv0 + r, a, iu, val);
            }
        }
    }
}

}  // namespace tsp
