// Backprojection: voxel-driven, bilinear detector interpolation, ray-density
// weight (semantics: SURVEY.md B.2; call site tomosipo/astra.py:147-153).
//
// A CTA owns a BP_TX x BP_TY x ZPT voxel tile and loops over *all* angles,
// so each voxel is written exactly once (ASTRA: one launch per 32 angles, each
// read-modify-writing the volume).  Per angle batch the detector footprint of
// the tile is staged into shared memory (zero-filled outside the detector,
// which is ASTRA's border mode) and sampled with full-precision fp32 weights
// in tile-local coordinates: the per-(tile, angle) affine maps are re-centred
// on the tile in fp64 by a few set-up threads, which keeps the fractional
// weights accurate to ~1e-6 irrespective of detector size.
#pragma once
#include "tsp_internal.h"

namespace tsp {

constexpr int BP_TX = 32;   // voxels along x per CTA (= warp width)
constexpr int BP_TY = 8;    // voxels along y per CTA
constexpr int BP_K = 8;     // angles staged per batch
constexpr int BP_WU = 64;   // staged footprint: max columns
constexpr int BP_PITCH = BP_WU + 1;
constexpr int BP_THREADS = BP_TX * BP_TY;
// staged footprint: max rows, as a function of the z voxels per thread (ZPT)
__host__ __device__ constexpr int bp_wv(int zpt) { return zpt + 14; }
__host__ __device__ constexpr size_t bp_smem_bytes(int zpt)
{
    return (size_t)BP_K * bp_wv(zpt) * BP_PITCH * sizeof(float);
}

// BP_SMEM_CLAMP: the footprint was clipped to the detector (+4 pixel zero margin), so
// buffer coordinates are clamped into the zero margin before sampling.
// BP_SMEM_B: like BP_SMEM, staged with the alternative row pitch (see BP_TMA_PITCH_B).
enum { BP_SKIP = 0, BP_SMEM = 1, BP_GLOBAL = 2, BP_SMEM_CLAMP = 3, BP_SMEM_B = 4 };

struct BPArgs {
    const float *proj;
    float *vol;
    int nx, ny, nz;
    int det_u, det_v, n_angles;
    const BPAngle *angles;
    float out_scale;  // voxel volume
    int additive;
    int vox_ss;
    // -4 * BP_MAGIC_BITS * (pitch + 1) mod 2^32 for the launched kernel's footprint pitch.  Passed
    // as a run-time value on purpose: as a literal, ptxas re-associates it out of the address
    // register and re-adds it in front of every shared-memory tap (3 extra instructions per update).
    uint32_t magic_off;
    uint32_t magic_off_b;  // same for the alternative pitch
    int no_rows3;          // tuning aid (TSP_BP_NO_ROWS3): disable the row-sharing z-invariant loops
    int rows_loop;         // which row-sharing loop z-invariant angles take: 3 = row walk (default), 2 = 3-row pairs (TSP_BP_ROWS)
    // fused SIRT update (tsp_sirt): when set, vol[i] -= epi_mul[i] * value instead of a plain store
    const float *epi_mul;
};

// The one place a voxel value is written: SET, ADD, or the fused SIRT update.
__device__ __forceinline__ void bp_store_one(const BPArgs &P, size_t idx, float v)
{
    float *dst = P.vol + idx;
    if (P.epi_mul) *dst = *dst - __ldg(P.epi_mul + idx) * v;
    else *dst = P.additive ? *dst + v : v;
}

// Per-(tile, angle) set-up result, tile-local: for a voxel at offset
// (dx, dy, dz) from the tile centre,
//   column = (Bu + Au.d) / (Bd + Ad.d),   row = (Bv + Av.d) / (Bd + Ad.d)
// are *buffer* coordinates (BP_SMEM) or detector index coordinates (BP_GLOBAL).
struct BPLocal {
    float au[3], bu;
    float av[3], bv;
    float ad[3], bd;
    int u_lo, v_lo, wu, wv;
    int mode;
    float weight;
    int z_invariant;  // column and magnification do not change along the z run
    int pad;
};

__device__ __forceinline__ float bp_sample_global(const float *__restrict__ proj, int det_u, int det_v,
                                                  size_t row_pitch, float fu, float fv)
{
    if (!(fu > -1.0f && fu < (float)det_u && fv > -1.0f && fv < (float)det_v)) return 0.0f;
    const float flu = floorf(fu), flv = floorf(fv);
    const int iu = (int)flu, iv = (int)flv;
    const float wu = fu - flu, wv = fv - flv;
    const bool u0 = iu >= 0, u1 = iu + 1 < det_u, v0 = iv >= 0, v1 = iv + 1 < det_v;
    const float *s = proj + (long long)iv * (long long)row_pitch + iu;
    const float p00 = (u0 && v0) ? __ldg(s) : 0.0f;
    const float p10 = (u1 && v0) ? __ldg(s + 1) : 0.0f;
    const float p01 = (u0 && v1) ? __ldg(s + row_pitch) : 0.0f;
    const float p11 = (u1 && v1) ? __ldg(s + row_pitch + 1) : 0.0f;
    const float lo = fmaf(wu, p10 - p00, p00);
    const float hi = fmaf(wu, p11 - p01, p01);
    return fmaf(wv, hi - lo, lo);
}

// Second half of the set-up, shared by both front ends: from the centre values (fp64) and the bounding box of the
// tile's footprint to the tile-local map, the staged box and the loop the consumers take.
__device__ __forceinline__ void bp_finish_setup(const BPArgs &P, const BPAngle *__restrict__ ang, double den_c, double nu_c,
                                                double nv_c, double umin, double umax, double vmin, double vmax,
                                                double dmin, double dmax, double hz, int max_rows, int max_cols,
                                                int alt_cols, int u_align, bool allow_rows3, int rows_loop, BPLocal *out)
{
    BPLocal L;
    int mode = BP_SMEM;
    // The projective map is monotone over the box only if den keeps its sign.
    const bool regular = (dmin > 0.0 || dmax < 0.0) && isfinite(umin) && isfinite(umax) &&
                         isfinite(vmin) && isfinite(vmax) &&
                         fmin(fabs(dmin), fabs(dmax)) > 1e-6 * fmax(fabs(dmin), fabs(dmax));
    double off_u = 0.5, off_v = 0.5;  // texel-centre convention -> index coordinates
    int u_lo = 0, v_lo = 0, wu = 0, wv = 0;
    if (!regular) {
        mode = BP_GLOBAL;
    } else {
        const double cu0 = -4.0, cv0 = -4.0, cu1 = (double)P.det_u + 4.0, cv1 = (double)P.det_v + 4.0;
        if (umin < cu0 || vmin < cv0 || umax > cu1 || vmax > cv1) mode = BP_SMEM_CLAMP;
        umin = fmax(umin, cu0); vmin = fmax(vmin, cv0);
        umax = fmin(umax, cu1); vmax = fmin(vmax, cv1);
        if (umax < umin || vmax < vmin) {
            mode = BP_SKIP;  // footprint entirely off the detector
        } else {
            u_lo = (int)floor(umin - 0.5) - 1;
            // TMA needs the innermost box coordinate on a 16-byte boundary (measured: any
            // other start faults with "illegal instruction"); u_align = 4 floats there.
            u_lo -= ((u_lo % u_align) + u_align) % u_align;
            v_lo = (int)floor(vmin - 0.5) - 1;
            wu = (int)floor(umax - 0.5) + 3 - u_lo;
            wv = (int)floor(vmax - 0.5) + 3 - v_lo;
            if (wu > max_cols || wv > max_rows) {
                mode = BP_GLOBAL;
            } else {
                // Along a warp (x) the column moves by dU/dx per lane and the row by dV/dx.  With
                // a row pitch = +4 banks, lanes on different rows collide only if column and row
                // move in opposite directions; with pitch = -4 banks only if they move together
                // (scratch/bank_sim.py: 1.37 wavefronts per tap with one fixed pitch, 1.02 with
                // the pitch chosen by this sign).  Staging picks the pitch per (tile, angle).
                const double su = ang->nu[0] * den_c - nu_c * ang->dn[0];
                const double sv = ang->nv[0] * den_c - nv_c * ang->dn[0];
                if (mode == BP_SMEM && alt_cols > 0 && wu <= alt_cols && su * sv < 0.0) mode = BP_SMEM_B;
                off_u += (double)u_lo;
                off_v += (double)v_lo;
            }
        }
    }
    // (Folding the run's first dz into bu / bv / bd - 6 instead of 9 FMAs per thread and angle - and reading the map
    // through an opaque base register without the S2R that ptxas re-issues per use were both measured SLOWER:
    // 48.9 / 48.7 vs 48.0 ms in one run, r02 GPU call 9.  The unrolled loop's schedule is what matters.)
    const double au2 = ang->nu[2] - off_u * ang->dn[2], av2 = ang->nv[2] - off_v * ang->dn[2];
    L.au[0] = (float)(ang->nu[0] - off_u * ang->dn[0]);
    L.au[1] = (float)(ang->nu[1] - off_u * ang->dn[1]);
    L.au[2] = (float)au2;
    L.bu = (float)(nu_c - off_u * den_c);
    L.av[0] = (float)(ang->nv[0] - off_v * ang->dn[0]);
    L.av[1] = (float)(ang->nv[1] - off_v * ang->dn[1]);
    L.av[2] = (float)av2;
    L.bv = (float)(nv_c - off_v * den_c);
    L.ad[0] = (float)ang->dn[0];
    L.ad[1] = (float)ang->dn[1];
    L.ad[2] = (float)ang->dn[2];
    L.bd = (float)den_c;
    L.u_lo = u_lo; L.v_lo = v_lo; L.wu = wu; L.wv = wv;
    L.mode = mode;
    L.weight = (float)ang->weight;
    // U (and the cone magnification) independent of z over this tile?  Exact zeros for
    // detectors whose rows are parallel to the z axis (all circular geometries).
    L.z_invariant = ((fabs(au2) + 64.0 * fabs(ang->dn[2])) * (2.0 * hz + 1.0) < 1e-7 * fabs(den_c)) ? 1 : 0;
    if (L.z_invariant && mode != BP_GLOBAL && mode != BP_SKIP && allow_rows3) {
        // row step per voxel, dv = av2 / den, largest where |den| is smallest on the tile; the 3-row
        // loop also reads row j+2 of the last voxel pair: one spare row must fit the staged box
        const double dvmax = fabs(av2) / fmin(fabs(dmin), fabs(dmax));
        const bool positive = (av2 >= 0.0) == (den_c > 0.0);
        // (the row walk reads no spare row)
        if (positive && dvmax <= 0.9999 && (rows_loop == 3 || wv + 1 <= max_rows)) L.z_invariant = rows_loop;
    }
    L.pad = 0;
    *out = L;
}

// Eight threads per angle: each projects one corner of the tile's voxel-centre
// box; shuffles reduce the bounding box; lane 0 writes the local map.
__device__ __forceinline__ void bp_setup(const BPArgs &P, const BPAngle *__restrict__ ang, int corner,
                                         double xc, double yc, double zc, double hx, double hy,
                                         double hz, int max_rows, int max_cols, int alt_cols, int u_align, bool allow_rows3, int rows_loop, BPLocal *out)
{
    const double den_c = ang->dn[0] * xc + ang->dn[1] * yc + ang->dn[2] * zc + ang->dn[3];
    const double nu_c = ang->nu[0] * xc + ang->nu[1] * yc + ang->nu[2] * zc + ang->nu[3];
    const double nv_c = ang->nv[0] * xc + ang->nv[1] * yc + ang->nv[2] * zc + ang->nv[3];
    const double dx = (corner & 1) ? hx : -hx;
    const double dy = (corner & 2) ? hy : -hy;
    const double dz = (corner & 4) ? hz : -hz;
    const double den = den_c + ang->dn[0] * dx + ang->dn[1] * dy + ang->dn[2] * dz;
    const double uu = (nu_c + ang->nu[0] * dx + ang->nu[1] * dy + ang->nu[2] * dz) / den;
    const double vv = (nv_c + ang->nv[0] * dx + ang->nv[1] * dy + ang->nv[2] * dz) / den;
    double umin = uu, umax = uu, vmin = vv, vmax = vv, dmin = den, dmax = den;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        umin = fmin(umin, __shfl_xor_sync(0xffffffffu, umin, o));
        umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (corner != 0) return;
    bp_finish_setup(P, ang, den_c, nu_c, nv_c, umin, umax, vmin, vmax, dmin, dmax, hz, max_rows, max_cols, alt_cols, u_align,
                    allow_rows3, rows_loop, out);
}

// One lane = one angle (the TMA kernel's producer: 32 angles per pass).  The centre values and everything that enters
// the fractional weights stay fp64; the footprint's bounding box - which only has to be right to a fraction of the
// one-pixel margin the staged box carries - comes from the 8 corners in fp32, widened by 1/128 pixel.  No shuffles,
// no fp64 divisions: ~6 producer instructions per (tile, angle) instead of ~100 for the 8-lanes-per-angle front end
// (measured: the producer's set-up cost 7 % of the kernel, r02 GPU call 4).
__device__ __forceinline__ void bp_setup_lane(const BPArgs &P, const BPAngle *__restrict__ ang, double xc, double yc,
                                              double zc, float hx, float hy, float hz, int max_rows, int max_cols,
                                              int alt_cols, int u_align, bool allow_rows3, int rows_loop, BPLocal *out)
{
    const double den_c = ang->dn[0] * xc + ang->dn[1] * yc + ang->dn[2] * zc + ang->dn[3];
    const double nu_c = ang->nu[0] * xc + ang->nu[1] * yc + ang->nu[2] * zc + ang->nu[3];
    const double nv_c = ang->nv[0] * xc + ang->nv[1] * yc + ang->nv[2] * zc + ang->nv[3];
    const float dc = (float)den_c, uc = (float)nu_c, vc = (float)nv_c;
    const float ddx = hx * (float)ang->dn[0], ddy = hy * (float)ang->dn[1], ddz = hz * (float)ang->dn[2];
    const float dux = hx * (float)ang->nu[0], duy = hy * (float)ang->nu[1], duz = hz * (float)ang->nu[2];
    const float dvx = hx * (float)ang->nv[0], dvy = hy * (float)ang->nv[1], dvz = hz * (float)ang->nv[2];
    float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f, dmin = 3.0e38f, dmax = -3.0e38f;
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float sx = (c & 1) ? 1.0f : -1.0f, sy = (c & 2) ? 1.0f : -1.0f, sz = (c & 4) ? 1.0f : -1.0f;
        const float den = dc + sx * ddx + sy * ddy + sz * ddz;
        const float r = 1.0f / den;
        const float uu = (uc + sx * dux + sy * duy + sz * duz) * r;
        const float vv = (vc + sx * dvx + sy * dvy + sz * dvz) * r;
        finite = finite && isfinite(uu) && isfinite(vv);
        umin = fminf(umin, uu); umax = fmaxf(umax, uu);
        vmin = fminf(vmin, vv); vmax = fmaxf(vmax, vv);
        dmin = fminf(dmin, den); dmax = fmaxf(dmax, den);
    }
    const float eps_u = 0.0078125f + 1e-6f * fmaxf(fabsf(umin), fabsf(umax));
    const float eps_v = 0.0078125f + 1e-6f * fmaxf(fabsf(vmin), fabsf(vmax));
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    bp_finish_setup(P, ang, den_c, nu_c, nv_c, finite ? (double)(umin - eps_u) : nan, (double)(umax + eps_u), (double)(vmin - eps_v),
                    (double)(vmax + eps_v), (double)dmin, (double)dmax, (double)hz, max_rows, max_cols, alt_cols, u_align,
                    allow_rows3, rows_loop, out);
}

template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr float BP_MAGIC = 12582912.0f;      // 1.5 * 2^23: a round-down add leaves floor(f) in the low mantissa bits (0 <= f < 2^22)
constexpr uint32_t BP_MAGIC_BITS = 0x4B400000u;

// Inner loop over the register-resident z run, sampling the staged footprint.
// `sbase` is the shared-memory byte address of the footprint buffer.
template <bool CONE, bool CLAMP, int ZPT, int PITCH>
__device__ __forceinline__ void bp_tile_loop(uint32_t sbase, float nu, float nv, float dn, float su, float sv,
                                             float sd, float umax, float vmax, float wpar, uint32_t magic_off,
                                             float (&acc)[ZPT])
{
    // byte address = sbase + 4 * ((rv_bits - M) * PITCH + (ru_bits - M))
    const uint32_t cbase = sbase + magic_off;
#pragma unroll
    for (int i = 0; i < ZPT; ++i) {
        float fu, fv, w2;
        if (CONE) {
            const float r = rcp_approx(dn);
            fu = nu * r; fv = nv * r; w2 = r * r;
        } else {
            fu = nu; fv = nv; w2 = wpar;
        }
        if (CLAMP) {  // NaN-safe: fmaxf/fminf return the non-NaN operand
            fu = fminf(fmaxf(fu, 0.0f), umax);
            fv = fminf(fmaxf(fv, 0.0f), vmax);
        }
        const float ru = __fadd_rd(fu, BP_MAGIC), rv = __fadd_rd(fv, BP_MAGIC);  // round-down add == floor
        const float wu = fu - (ru - BP_MAGIC), wv = fv - (rv - BP_MAGIC);
        const uint32_t a = (__float_as_uint(rv) * (uint32_t)PITCH + __float_as_uint(ru)) * 4u + cbase;
        const float p00 = lds_f32<0>(a), p10 = lds_f32<4>(a);
        const float p01 = lds_f32<4 * PITCH>(a), p11 = lds_f32<4 * PITCH + 4>(a);
        const float lo = fmaf(wu, p10 - p00, p00);
        const float hi = fmaf(wu, p11 - p01, p01);
        const float val = fmaf(wv, hi - lo, lo);
        // behind-the-source voxels (w2 = inf) over an empty footprint must stay 0
        acc[i] = (CONE && CLAMP) ? ((val != 0.0f) ? fmaf(w2, val, acc[i]) : acc[i]) : fmaf(w2, val, acc[i]);
        nu += su; nv += sv;
        if (CONE) dn += sd;
    }
}

// Same, for angles whose detector column and magnification are constant along
// z (detector rows parallel to the z axis): the column, its weight and the
// ray-density weight are computed once per (x, y); only the row moves.
template <bool CONE, int ZPT, int PITCH>
__device__ __forceinline__ void bp_tile_loop_zinv(uint32_t sbase, float nu, float nv, float dn, float sv,
                                                  float wpar, uint32_t magic_off, float (&acc)[ZPT])
{
    float r = 1.0f, w2 = wpar;
    if (CONE) { r = rcp_approx(dn); w2 = r * r; }
    const float fu = nu * r;
    const float ru = __fadd_rd(fu, BP_MAGIC);
    const float wu = fu - (ru - BP_MAGIC);
    const uint32_t cbase = sbase + magic_off + 4u * __float_as_uint(ru);
    float fv = nv * r;
    const float dv = sv * r;
    if (ZPT % 2 == 0) {
        // Two voxels of the run per step with Blackwell's packed fp32x2 arithmetic (FFMA2 / FADD2:
        // two independent fp32 operations per issue slot - the kernel is issue-bound): 10.5
        // instead of 16 instructions per update.  Every pair (.x, .y) = (voxel i, voxel i + 1).
        const float2 M2 = make_float2(BP_MAGIC, BP_MAGIC), NEG1 = make_float2(-1.0f, -1.0f);
        const float2 wu2 = make_float2(wu, wu), w22 = make_float2(w2, w2), step2 = make_float2(2.0f * dv, 2.0f * dv);
        float2 fv2 = make_float2(fv, fv + dv);
#pragma unroll
        for (int i = 0; i < ZPT; i += 2) {
            const float2 rv2 = __fadd2_rd(fv2, M2);                       // round-down add == floor
            const float2 wv2 = __fadd2_rn(fv2, __ffma2_rn(rv2, NEG1, M2));  // fv - (rv - M)
            const uint32_t a0 = __float_as_uint(rv2.x) * (uint32_t)(4 * PITCH) + cbase;
            const uint32_t a1 = __float_as_uint(rv2.y) * (uint32_t)(4 * PITCH) + cbase;
            const float2 p00 = make_float2(lds_f32<0>(a0), lds_f32<0>(a1));
            const float2 p10 = make_float2(lds_f32<4>(a0), lds_f32<4>(a1));
            const float2 p01 = make_float2(lds_f32<4 * PITCH>(a0), lds_f32<4 * PITCH>(a1));
            const float2 p11 = make_float2(lds_f32<4 * PITCH + 4>(a0), lds_f32<4 * PITCH + 4>(a1));
            const float2 lo = __ffma2_rn(wu2, __ffma2_rn(p00, NEG1, p10), p00);
            const float2 hi = __ffma2_rn(wu2, __ffma2_rn(p01, NEG1, p11), p01);
            const float2 val = __ffma2_rn(wv2, __ffma2_rn(lo, NEG1, hi), lo);
            const float2 a2 = __ffma2_rn(w22, val, make_float2(acc[i], acc[i + 1]));
            acc[i] = a2.x; acc[i + 1] = a2.y;
            fv2 = __fadd2_rn(fv2, step2);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < ZPT; ++i) {
        const float rv = __fadd_rd(fv, BP_MAGIC);
        const float wv = fv - (rv - BP_MAGIC);
        const uint32_t a = __float_as_uint(rv) * (uint32_t)(4 * PITCH) + cbase;
        const float p00 = lds_f32<0>(a), p10 = lds_f32<4>(a);
        const float p01 = lds_f32<4 * PITCH>(a), p11 = lds_f32<4 * PITCH + 4>(a);
        const float lo = fmaf(wu, p10 - p00, p00);
        const float hi = fmaf(wu, p11 - p01, p01);
        const float val = fmaf(wv, hi - lo, lo);
        acc[i] = fmaf(w2, val, acc[i]);
        fv += dv;
    }
}

// z-invariant angles whose row coordinate advances by 0 <= dv <= 1 per voxel (the detector
// pixel is at least as tall as a projected voxel - every configuration of BASELINE.json):
// voxel i + 1 samples row pair (j, j+1) or (j+1, j+2) of voxel i's row j, so a pair of voxels
// needs 3 rows x 2 columns = 6 shared-memory taps instead of 8.  The kernel is bound by the
// shared-memory pipe after the packed-arithmetic rewrite; this trades two taps for one select.
template <bool CONE, int ZPT, int PITCH>
__device__ __forceinline__ void bp_tile_loop_zinv3(uint32_t sbase, float nu, float nv, float dn, float sv,
                                                   float wpar, uint32_t magic_off, float (&acc)[ZPT])
{
    static_assert(ZPT % 2 == 0, "pairs of voxels");
    float r = 1.0f, w2 = wpar;
    if (CONE) { r = rcp_approx(dn); w2 = r * r; }
    const float fu = nu * r;
    const float ru = __fadd_rd(fu, BP_MAGIC);
    const float wu = fu - (ru - BP_MAGIC);
    const uint32_t cbase = sbase + magic_off + 4u * __float_as_uint(ru);
    const float fv = nv * r;
    const float dv = sv * r;
    const float2 M2 = make_float2(BP_MAGIC, BP_MAGIC), NEG1 = make_float2(-1.0f, -1.0f);
    const float2 wu2 = make_float2(wu, wu), w22 = make_float2(w2, w2), step2 = make_float2(2.0f * dv, 2.0f * dv);
    float2 fv2 = make_float2(fv, fv + dv);
#pragma unroll
    for (int i = 0; i < ZPT; i += 2) {
        const float2 rv2 = __fadd2_rd(fv2, M2);
        const float2 wv2 = __fadd2_rn(fv2, __ffma2_rn(rv2, NEG1, M2));
        const uint32_t a = __float_as_uint(rv2.x) * (uint32_t)(4 * PITCH) + cbase;
        const bool next_row = __float_as_uint(rv2.y) != __float_as_uint(rv2.x);
        // rows j, j+1 (packed) and j+2 (scalar), two columns each
        const float2 p0 = make_float2(lds_f32<0>(a), lds_f32<4 * PITCH>(a));
        const float2 p1 = make_float2(lds_f32<4>(a), lds_f32<4 * PITCH + 4>(a));
        const float q0 = lds_f32<8 * PITCH>(a), q1 = lds_f32<8 * PITCH + 4>(a);
        const float2 h01 = __ffma2_rn(wu2, __ffma2_rn(p0, NEG1, p1), p0);   // (h(j), h(j+1))
        const float h2 = fmaf(wu, q1 - q0, q0);                             // h(j+2)
        const float2 h12 = make_float2(h01.y, h2);
        const float2 d = __ffma2_rn(h01, NEG1, h12);                        // (h1 - h0, h2 - h1)
        float2 val = __ffma2_rn(wv2, d, h01);                               // (voxel i, voxel i+1 if it moved on a row)
        const float same = fmaf(wv2.y, d.x, h01.x);                         // voxel i+1 if it stayed on row j
        val.y = next_row ? val.y : same;
        const float2 a2 = __ffma2_rn(w22, val, make_float2(acc[i], acc[i + 1]));
        acc[i] = a2.x; acc[i + 1] = a2.y;
        fv2 = __fadd2_rn(fv2, step2);
    }
}

// Same precondition (z-invariant, 0 <= dv <= 1), walking the detector ROWS instead: voxel i
// samples rows (j_i, j_i + 1) with j_i - j_{i-1} in {0, 1}, so the only row it can need that the
// previous voxel did not have is j_i + 1.  h(j) = lerp(p[j][c], p[j][c+1], wu) of that row is
// formed once (2 taps, one packed lerp per voxel pair) and h(j_i) is selected from the previous
// voxel's two values: 2 shared-memory taps and 9.5 instructions per voxel instead of 3 / 11.
// The loop is pitch-agnostic (the row pitch is a register), so one copy serves both TMA variants.
// (A variant with the row coordinate as a 32-bit fixed-point phase - carry = "advance one row", weight = I2FP(phase) -
// moves 5 of the 11 FMA-pipe instructions per voxel pair to the integer pipe but needs 22 instead of 19 instructions:
// 51.5 vs 47.7 ms, r02 GPU call 17.  The kernel is bound by instruction issue, not by either math pipe.)
template <bool CONE, int ZPT>
__device__ __forceinline__ void bp_tile_loop_rows(uint32_t sbase, uint32_t pitch4, float nu, float nv, float dn, float sv,
                                                  float wpar, uint32_t magic_off, float (&acc)[ZPT])
{
    static_assert(ZPT % 2 == 0, "pairs of voxels");
    float r = 1.0f, w2 = wpar;
    if (CONE) { r = rcp_approx(dn); w2 = r * r; }
    const float fu = nu * r;
    const float ru = __fadd_rd(fu, BP_MAGIC);
    const float wu = fu - (ru - BP_MAGIC);
    const uint32_t cbase = sbase + magic_off + 4u * __float_as_uint(ru);  // row j, column c of the first tap
    const uint32_t cbase1 = cbase + pitch4;                               // row j + 1
    const float fv = nv * r;
    const float dv = sv * r;
    const float2 M2 = make_float2(BP_MAGIC, BP_MAGIC), NEG1 = make_float2(-1.0f, -1.0f);
    const float2 wu2 = make_float2(wu, wu), w22 = make_float2(w2, w2), step2 = make_float2(2.0f * dv, 2.0f * dv);
    float2 fv2 = make_float2(fv, fv + dv);
    // state carried from voxel i - 1: its row index (as magic-float bits) and h of its two rows
    const float rv0 = __fadd_rd(fv, BP_MAGIC);
    uint32_t prev = __float_as_uint(rv0);
    float h0p, h1p;
    {
        const uint32_t a = prev * pitch4 + cbase;
        const float t0 = lds_f32<0>(a), t1 = lds_f32<4>(a);
        h0p = fmaf(wu, t1 - t0, t0);
        h1p = h0p;
    }
#pragma unroll
    for (int i = 0; i < ZPT; i += 2) {
        const float2 rv2 = __fadd2_rd(fv2, M2);
        const float2 wv2 = __fadd2_rn(fv2, __ffma2_rn(rv2, NEG1, M2));
        const uint32_t bx = __float_as_uint(rv2.x), by = __float_as_uint(rv2.y);
        const uint32_t a0 = bx * pitch4 + cbase1, a1 = by * pitch4 + cbase1;
        const float2 q0 = make_float2(lds_f32<0>(a0), lds_f32<0>(a1));
        const float2 q1 = make_float2(lds_f32<4>(a0), lds_f32<4>(a1));
        const float2 hn = __ffma2_rn(wu2, __ffma2_rn(q0, NEG1, q1), q0);  // h(j_i + 1), h(j_{i+1} + 1)
        float2 hl;                                                        // h(j_i), h(j_{i+1})
        hl.x = (bx != prev) ? h1p : h0p;
        hl.y = (by != bx) ? hn.x : hl.x;
        const float2 d = __ffma2_rn(hl, NEG1, hn);
        const float2 val = __ffma2_rn(wv2, d, hl);
        const float2 a2 = __ffma2_rn(w22, val, make_float2(acc[i], acc[i + 1]));
        acc[i] = a2.x; acc[i + 1] = a2.y;
        prev = by; h0p = hl.y; h1p = hn.y;
        fv2 = __fadd2_rn(fv2, step2);
    }
}

// One angle's contribution to a thread's z run.  `L` lives in shared memory;
// `sbase` is the shared byte address of the staged footprint (row pitch PITCH).
template <bool CONE, int ZPT, int PITCH, int PITCH_B>
__device__ __forceinline__ void bp_accumulate_angle(const BPArgs &P, const BPLocal &L, uint32_t sbase, int angle,
                                                    float dx, float dy, float dz0, size_t row_pitch,
                                                    float (&acc)[ZPT])
{
    const int mode = L.mode;
    if (mode == BP_SKIP) return;
    float nu = fmaf(L.au[0], dx, fmaf(L.au[1], dy, fmaf(L.au[2], dz0, L.bu)));
    float nv = fmaf(L.av[0], dx, fmaf(L.av[1], dy, fmaf(L.av[2], dz0, L.bv)));
    float dn = CONE ? fmaf(L.ad[0], dx, fmaf(L.ad[1], dy, fmaf(L.ad[2], dz0, L.bd))) : 1.0f;
    const float su = L.au[2], sv = L.av[2], sd = L.ad[2];
    const float wpar = L.weight;
    if (ZPT % 2 == 0 && L.z_invariant == 3) {  // only ever set for BP_SMEM / BP_SMEM_B
        const bool b = PITCH_B != PITCH && mode == BP_SMEM_B;
        bp_tile_loop_rows<CONE, (ZPT % 2 == 0 ? ZPT : 2)>(sbase, b ? 4u * PITCH_B : 4u * PITCH, nu, nv, dn, sv, wpar, b ? P.magic_off_b : P.magic_off,
                                                          reinterpret_cast<float (&)[(ZPT % 2 == 0 ? ZPT : 2)]>(acc));
    } else if (mode == BP_SMEM) {
        if (ZPT % 2 == 0 && L.z_invariant == 2) bp_tile_loop_zinv3<CONE, (ZPT % 2 == 0 ? ZPT : 2), PITCH>(sbase, nu, nv, dn, sv, wpar, P.magic_off, reinterpret_cast<float (&)[(ZPT % 2 == 0 ? ZPT : 2)]>(acc));
        else if (L.z_invariant) bp_tile_loop_zinv<CONE, ZPT, PITCH>(sbase, nu, nv, dn, sv, wpar, P.magic_off, acc);
        else bp_tile_loop<CONE, false, ZPT, PITCH>(sbase, nu, nv, dn, su, sv, sd, 0.0f, 0.0f, wpar, P.magic_off, acc);
    } else if (PITCH_B != PITCH && mode == BP_SMEM_B) {
        if (ZPT % 2 == 0 && L.z_invariant == 2) bp_tile_loop_zinv3<CONE, (ZPT % 2 == 0 ? ZPT : 2), PITCH_B>(sbase, nu, nv, dn, sv, wpar, P.magic_off_b, reinterpret_cast<float (&)[(ZPT % 2 == 0 ? ZPT : 2)]>(acc));
        else if (L.z_invariant) bp_tile_loop_zinv<CONE, ZPT, PITCH_B>(sbase, nu, nv, dn, sv, wpar, P.magic_off_b, acc);
        else bp_tile_loop<CONE, false, ZPT, PITCH_B>(sbase, nu, nv, dn, su, sv, sd, 0.0f, 0.0f, wpar, P.magic_off_b, acc);
    } else if (mode == BP_SMEM_CLAMP) {
        bp_tile_loop<CONE, true, ZPT, PITCH>(sbase, nu, nv, dn, su, sv, sd, (float)L.wu - 1.5f, (float)L.wv - 1.5f,
                                             wpar, P.magic_off, acc);
    } else {
        const float *src = P.proj + (size_t)angle * P.det_u;
#pragma unroll
        for (int i = 0; i < ZPT; ++i) {
            float fu, fv, w2;
            if (CONE) {
                const float r = 1.0f / dn;
                fu = nu * r; fv = nv * r; w2 = r * r;
            } else {
                fu = nu; fv = nv; w2 = wpar;
            }
            const float val = bp_sample_global(src, P.det_u, P.det_v, row_pitch, fu, fv);
            if (val != 0.0f) acc[i] = fmaf(w2, val, acc[i]);
            nu += su; nv += sv;
            if (CONE) dn += sd;
        }
    }
}

template <int ZPT>
__device__ __forceinline__ void bp_store(const BPArgs &P, int x, int y, int z0, const float (&acc)[ZPT])
{
#pragma unroll
    for (int i = 0; i < ZPT; ++i) {
        const int z = z0 + i;
        if (z < P.nz) {
            bp_store_one(P, ((size_t)z * P.ny + y) * P.nx + x, acc[i] * P.out_scale);
        }
    }
}

template <bool CONE, int BP_ZPT>
__global__ void __launch_bounds__(BP_THREADS) bp_kernel(const BPArgs P)
{
    constexpr int BP_WV = bp_wv(BP_ZPT);
    extern __shared__ __align__(16) float bp_dyn_smem[];
    float(*buf)[BP_WV * BP_PITCH] = reinterpret_cast<float(*)[BP_WV * BP_PITCH]>(bp_dyn_smem);
    __shared__ BPLocal loc[BP_K];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * BP_TX + tx;
    const int x0 = blockIdx.x * BP_TX, y0 = blockIdx.y * BP_TY, z0 = blockIdx.z * BP_ZPT;
    // z is not clipped to the volume: every thread walks its whole z run (only the store is guarded)
    const int x1 = min(x0 + BP_TX, P.nx) - 1, y1 = min(y0 + BP_TY, P.ny) - 1, z1 = z0 + BP_ZPT - 1;
    // voxel-centre box of the tile in the normalised frame
    const double xc = 0.5 * (x0 + x1) + 0.5 - 0.5 * P.nx, hx = 0.5 * (x1 - x0);
    const double yc = 0.5 * (y0 + y1) + 0.5 - 0.5 * P.ny, hy = 0.5 * (y1 - y0);
    const double zc = 0.5 * (z0 + z1) + 0.5 - 0.5 * P.nz, hz = 0.5 * (z1 - z0);

    const int x = x0 + tx, y = y0 + ty;
    const float dx = (float)((double)x + 0.5 - 0.5 * P.nx - xc);
    const float dy = (float)((double)y + 0.5 - 0.5 * P.ny - yc);
    const float dz0 = (float)((double)z0 + 0.5 - 0.5 * P.nz - zc);
    const bool in_xy = (x < P.nx) && (y < P.ny);
    const size_t row_pitch = (size_t)P.n_angles * P.det_u;

    float acc[BP_ZPT];
#pragma unroll
    for (int i = 0; i < BP_ZPT; ++i) acc[i] = 0.0f;

    for (int a0 = 0; a0 < P.n_angles; a0 += BP_K) {
        const int na = min(BP_K, P.n_angles - a0);
        __syncthreads();  // previous batch fully consumed
        if (tid < 8 * BP_K) {
            const int j = tid >> 3;
            // clamp so that all 8 lanes of a group take part in the shuffles
            const int a = a0 + min(j, na - 1);
            bp_setup(P, P.angles + a, tid & 7, xc, yc, zc, hx, hy, hz, BP_WV, BP_WU, 0, 1, false, 0, &loc[j]);
        }
        __syncthreads();
        // stage footprints: warps over rows, lanes over columns
        for (int j = 0; j < na; ++j) {
            if (loc[j].mode != BP_SMEM && loc[j].mode != BP_SMEM_CLAMP) continue;  // (BP_SMEM_B never occurs here)
            const int u_lo = loc[j].u_lo, v_lo = loc[j].v_lo, wu = loc[j].wu, wv = loc[j].wv;
            const float *src = P.proj + (size_t)(a0 + j) * P.det_u;
            for (int r = ty; r < wv; r += BP_TY) {
                const int gv = v_lo + r;
                const bool vin = (gv >= 0) && (gv < P.det_v);
                for (int c = tx; c < wu; c += BP_TX) {
                    const int gu = u_lo + c;
                    float val = 0.0f;
                    if (vin && gu >= 0 && gu < P.det_u) val = __ldg(src + (size_t)gv * row_pitch + gu);
                    buf[j][r * BP_PITCH + c] = val;
                }
            }
        }
        __syncthreads();
        if (!in_xy) continue;
        for (int j = 0; j < na; ++j) {
            bp_accumulate_angle<CONE, BP_ZPT, BP_PITCH, BP_PITCH>(P, loc[j], (uint32_t)__cvta_generic_to_shared(buf[j]), a0 + j,
                                                        dx, dy, dz0, row_pitch, acc);
        }
    }
    if (in_xy) bp_store<BP_ZPT>(P, x, y, z0, acc);
}

// ---------------------------------------------------------------------------
// TMA-staged, warp-specialised variant (the default when the projection array
// satisfies TMA's alignment rules: 16-byte aligned base, det_u % 4 == 0).
//
//   warp 8 (producer): per angle, projects the tile's corners (fp64), writes the
//       tile-local map, and has the TMA engine copy the footprint box
//       proj[v_lo : v_lo+WV, angle, u_lo : u_lo+68] into a ring stage
//       (cp.async.bulk.tensor.3d; out-of-detector elements arrive as zeros,
//       which is exactly the projector's border rule).
//   warps 0-7 (consumers): wait on the stage's "full" mbarrier, accumulate the
//       angle into their register-resident z runs, release the stage.
// No block-wide barrier and no staging instructions in the compute warps.
// Row pitch of a staged footprint = TMA box width.  64 + 4: consecutive rows start 4 banks
// apart, so lanes that sit in the same detector column on different rows (angles whose u
// axis is nearly perpendicular to x) do not collide on a bank (ncu, pitch 64: 43 % excess
// shared wavefronts on every tap).
constexpr int BP_TMA_PITCH = 68;
constexpr int BP_TMA_PITCH_B = 60;  // alternative: consecutive rows start 4 banks *earlier*
constexpr int BP_TMA_CONSUMERS = BP_TX * BP_TY;
// A second helper warp (the "publisher") waits on the TMA barriers and publishes "angles ready" in a plain
// shared-memory word, which the consumers poll with ld.shared (29 cycles) instead of mbarrier.try_wait (90+ cycles,
// paid by all eight consumer warps at the same moment): 48.5 -> 47.7 ms at cfg 3 (r02 GPU call 5).
// -DBP_NO_PUBLISHER restores consumers that wait on the barrier themselves.
#ifndef BP_NO_PUBLISHER
#define BP_PUBLISHER 1
#endif
#ifdef BP_PUBLISHER
constexpr int BP_TMA_HELPERS = 2;
#else
constexpr int BP_TMA_HELPERS = 1;
#endif
constexpr int BP_TMA_THREADS = BP_TMA_CONSUMERS + 32 * BP_TMA_HELPERS;
// "tall" variant (ZPT = 64): a 32 x 16 x 64 tile, one CTA of 16 consumer warps per SM; the per-angle work of a thread
// (map, reciprocal, hand-off: ~67 instructions) is spread over twice the voxels
__host__ __device__ constexpr int bp_tma_ty(int zpt) { return zpt >= 64 ? 16 : BP_TY; }
__host__ __device__ constexpr int bp_tma_threads(int zpt) { return BP_TX * bp_tma_ty(zpt) + 32 * BP_TMA_HELPERS; }
#ifndef BP_TMA_STAGES_Z32
#define BP_TMA_STAGES_Z32 4  // ring depth at 32 voxels per thread (tuning: -DBP_TMA_STAGES_Z32=n)
#endif
__host__ __device__ constexpr int bp_tma_stages(int zpt) { return zpt >= 32 ? BP_TMA_STAGES_Z32 : (zpt >= 24 ? 4 : (zpt >= 16 ? 6 : 8)); }
__host__ __device__ constexpr size_t bp_tma_box_bytes(int zpt) { return (size_t)bp_wv(zpt) * BP_TMA_PITCH * 4; }
// ring stages are 128-byte aligned (TMA destination alignment)
__host__ __device__ constexpr size_t bp_tma_stage_bytes(int zpt) { return (bp_tma_box_bytes(zpt) + 127) / 128 * 128; }
__host__ __device__ constexpr size_t bp_tma_smem_bytes(int zpt)
{
    return bp_tma_stages(zpt) * (bp_tma_stage_bytes(zpt) + sizeof(BPLocal) + 16) + 128 + 16 + 2 * 20 * 33 * 4;
}
__host__ __device__ constexpr int bp_tma_min_ctas(int zpt) { return zpt >= 64 ? 1 : (zpt >= 32 ? 2 : 3); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_box_3d(uint32_t dst, const void *tmap, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

template <bool CONE, int ZPT>
__global__ void __launch_bounds__(bp_tma_threads(ZPT), bp_tma_min_ctas(ZPT))
bp_tma_kernel(const BPArgs P, const __grid_constant__ TensorMapPair tmaps)  // m[0]: pitch 68 boxes, m[1]: pitch 60
{
    const TensorMapBlob *tmap = tmaps.m;
    constexpr int WV = bp_wv(ZPT);
    constexpr int STAGES = bp_tma_stages(ZPT);
    constexpr uint32_t STAGE_BYTES = (uint32_t)bp_tma_stage_bytes(ZPT);
    constexpr uint32_t BOX_BYTES = (uint32_t)bp_tma_box_bytes(ZPT);
    constexpr uint32_t BOX_BYTES_B = (uint32_t)(bp_wv(ZPT) * BP_TMA_PITCH_B * 4);

    extern __shared__ __align__(128) unsigned char bp_tma_smem[];
    // carve: [stages x footprint] [stages x BPLocal] [full barriers] [empty barriers];
    // the base is aligned by an offset (not an integer round trip) so that the compiler
    // keeps the shared state space of every pointer derived from it
    // carve: [stages x footprint] [stages x BPLocal] [full barriers] [empty barriers] [ready] [producer scratch];
    // the base is aligned by an offset (not an integer round trip) so that the compiler
    // keeps the shared state space of every pointer derived from it
    unsigned char *base = bp_tma_smem + ((128u - (smem_u32(bp_tma_smem) & 127u)) & 127u);
    BPLocal *loc = reinterpret_cast<BPLocal *>(base + (size_t)STAGES * STAGE_BYTES);
    const uint32_t bufs = smem_u32(base);
    const uint32_t loc0 = smem_u32(loc);
    const uint32_t full = smem_u32(loc + STAGES);
    const uint32_t empty = full + 8u * STAGES;
    const uint32_t ready = empty + 8u * STAGES;  // BP_PUBLISHER: number of angles whose footprint has landed
    constexpr uint32_t PEND_BYTES = 20u * 33u * 4u;
    const uint32_t pend = ready + 16u;           // producer scratch: 2 x (20 words x 33) (see the producer)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int TY = bp_tma_ty(ZPT);
    const int x0 = blockIdx.x * BP_TX, y0 = blockIdx.y * TY, z0 = blockIdx.z * ZPT;
    // z is not clipped to the volume: every thread walks its whole z run (only the store is guarded)
    const int x1 = min(x0 + BP_TX, P.nx) - 1, y1 = min(y0 + TY, P.ny) - 1, z1 = z0 + ZPT - 1;
    const double xc = 0.5 * (x0 + x1) + 0.5 - 0.5 * P.nx, hx = 0.5 * (x1 - x0);
    const double yc = 0.5 * (y0 + y1) + 0.5 - 0.5 * P.ny, hy = 0.5 * (y1 - y0);
    const double zc = 0.5 * (z0 + z1) + 0.5 - 0.5 * P.nz, hz = 0.5 * (z1 - z0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + 8u * s, 1);                       // the producer's arrive(+expect_tx)
            mbar_init(empty + 8u * s, TY);                     // one arrival per consumer warp
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ready), "r"(0u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

#ifdef BP_PUBLISHER
    if (warp == TY + 1) {
        // ----------------------------------------------------------- publisher
        int s = 0;
        uint32_t parity = 0u;
        for (int angle = 0; angle < P.n_angles; ++angle) {
            mbar_wait(full + 8u * s, parity);
            if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ready), "r"((uint32_t)angle + 1u) : "memory");
            if (++s == STAGES) { s = 0; parity ^= 1u; }
        }
        return;
    }
#endif
    if (warp == TY) {
        // ------------------------------------------------------------ producer
        // A pass = 32 angles, one per lane (bp_setup_lane).  The lane's map goes to a transposed scratch table
        // `pend` (word f of angle g at [f * 33 + g]: conflict-free both ways); when angle g's ring stage is free,
        // lanes 0-19 copy its 20 words into the stage's BPLocal and lane 0 issues the TMA.  The next pass is set up
        // right after the first angle of the current one has been issued - the moment the ring is full and the
        // consumers have several angles of work queued - so issuing is never held up by the set-up arithmetic.
        constexpr int LW = (int)(sizeof(BPLocal) / 4);
        static_assert(sizeof(BPLocal) == 80, "BPLocal is copied as 20 words");
        const float hxf = (float)hx, hyf = (float)hy, hzf = (float)hz;
        auto setup_pass = [&](int a0, uint32_t pend_base, int &mode, int &u_lo, int &v_lo) {
            BPLocal L;
            bp_setup_lane(P, P.angles + min(a0 + lane, P.n_angles - 1), xc, yc, zc, hxf, hyf, hzf, WV, BP_TMA_PITCH, BP_TMA_PITCH_B, 4,
                          !P.no_rows3, P.rows_loop, &L);
            const uint32_t w[LW] = {__float_as_uint(L.au[0]), __float_as_uint(L.au[1]), __float_as_uint(L.au[2]), __float_as_uint(L.bu),
                                    __float_as_uint(L.av[0]), __float_as_uint(L.av[1]), __float_as_uint(L.av[2]), __float_as_uint(L.bv),
                                    __float_as_uint(L.ad[0]), __float_as_uint(L.ad[1]), __float_as_uint(L.ad[2]), __float_as_uint(L.bd),
                                    (uint32_t)L.u_lo, (uint32_t)L.v_lo, (uint32_t)L.wu, (uint32_t)L.wv,
                                    (uint32_t)L.mode, __float_as_uint(L.weight), (uint32_t)L.z_invariant, 0u};
#pragma unroll
            for (int f = 0; f < LW; ++f)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(pend_base + 4u * (uint32_t)(f * 33 + lane)), "r"(w[f]) : "memory");
            mode = L.mode; u_lo = L.u_lo; v_lo = L.v_lo;
        };
        int s = 0;
        uint32_t parity = 1u;  // waiting on a fresh "empty" barrier with parity 1 passes at once
        int mode = 0, u_lo = 0, v_lo = 0, mode_n = 0, u_lo_n = 0, v_lo_n = 0;
        uint32_t pb = 0;
        setup_pass(0, pend, mode, u_lo, v_lo);
        __syncwarp();
        for (int a0 = 0; a0 < P.n_angles; a0 += 32) {
            const int n = min(32, P.n_angles - a0);
            const uint32_t pend_cur = pend + pb * PEND_BYTES;
            for (int g = 0; g < n; ++g) {
                const int angle = a0 + g;
                mbar_wait(empty + 8u * s, parity);
                const uint32_t loc_s = loc0 + (uint32_t)s * (uint32_t)sizeof(BPLocal);
                if (lane < LW) {
                    uint32_t v;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(pend_cur + 4u * (uint32_t)(lane * 33 + g)) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(loc_s + 4u * (uint32_t)lane), "r"(v) : "memory");
                }
                const int mode_g = __shfl_sync(0xffffffffu, mode, g);
                const int u_lo_g = __shfl_sync(0xffffffffu, u_lo, g), v_lo_g = __shfl_sync(0xffffffffu, v_lo, g);
                __syncwarp();  // the stage's map is complete before the footprint is requested
                if (lane == 0) {
                    if (mode_g == BP_SMEM || mode_g == BP_SMEM_CLAMP) {
                        mbar_arrive_expect_tx(full + 8u * s, BOX_BYTES);
                        tma_load_box_3d(bufs + (uint32_t)s * STAGE_BYTES, tmap, u_lo_g, angle, v_lo_g, full + 8u * s);
                    } else if (mode_g == BP_SMEM_B) {
                        mbar_arrive_expect_tx(full + 8u * s, BOX_BYTES_B);
                        tma_load_box_3d(bufs + (uint32_t)s * STAGE_BYTES, tmap + 1, u_lo_g, angle, v_lo_g, full + 8u * s);
                    } else {
                        mbar_arrive(full + 8u * s);
                    }
                }
                if (++s == STAGES) { s = 0; parity ^= 1u; }
                if (g == 0 && a0 + 32 < P.n_angles) setup_pass(a0 + 32, pend + (pb ^ 1u) * PEND_BYTES, mode_n, u_lo_n, v_lo_n);
            }
            __syncwarp();
            mode = mode_n; u_lo = u_lo_n; v_lo = v_lo_n;
            pb ^= 1u;
        }
        return;
    }

    // --------------------------------------------------------------- consumers
    const int tx = lane, ty = warp;
    const int x = x0 + tx, y = y0 + ty;
    const float dx = (float)((double)x + 0.5 - 0.5 * P.nx - xc);
    const float dy = (float)((double)y + 0.5 - 0.5 * P.ny - yc);
    const float dz0 = (float)((double)z0 + 0.5 - 0.5 * P.nz - zc);
    const bool in_xy = (x < P.nx) && (y < P.ny);
    const size_t row_pitch = (size_t)P.n_angles * P.det_u;

    float acc[ZPT];
#pragma unroll
    for (int i = 0; i < ZPT; ++i) acc[i] = 0.0f;

    int s = 0;
    uint32_t parity = 0u;
    for (int angle = 0; angle < P.n_angles; ++angle) {
#ifdef BP_PUBLISHER
        {
            uint32_t r;
            do {
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(ready) : "memory");
            } while (r <= (uint32_t)angle);
        }
#else
        mbar_wait(full + 8u * s, parity);
#endif
        if (in_xy)
            bp_accumulate_angle<CONE, ZPT, BP_TMA_PITCH, BP_TMA_PITCH_B>(P, loc[s], bufs + (uint32_t)s * STAGE_BYTES, angle, dx, dy, dz0,
                                                                         row_pitch, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8u * s);
        if (++s == STAGES) { s = 0; parity ^= 1u; }
    }
    if (in_xy) bp_store<ZPT>(P, x, y, z0, acc);
}

// Voxel supersampling (rare, API parity with VoxelSuperSampling > 1): one
// thread per voxel, direct bounds-checked gathers.
template <bool CONE>
__global__ void __launch_bounds__(256) bp_supersample_kernel(const BPArgs P)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= P.nx || y >= P.ny) return;
    const int ss = P.vox_ss;
    const size_t row_pitch = (size_t)P.n_angles * P.det_u;
    float acc = 0.0f;
    for (int a = 0; a < P.n_angles; ++a) {
        const BPAngle *g = P.angles + a;
        const float *src = P.proj + (size_t)a * P.det_u;
        for (int sz = 0; sz < ss; ++sz)
            for (int sy = 0; sy < ss; ++sy)
                for (int sx = 0; sx < ss; ++sx) {
                    const double px = x + (sx + 0.5) / ss - 0.5 * P.nx;
                    const double py = y + (sy + 0.5) / ss - 0.5 * P.ny;
                    const double pz = z + (sz + 0.5) / ss - 0.5 * P.nz;
                    const double den = g->dn[0] * px + g->dn[1] * py + g->dn[2] * pz + g->dn[3];
                    const double r = 1.0 / den;
                    const float fu = (float)((g->nu[0] * px + g->nu[1] * py + g->nu[2] * pz + g->nu[3]) * r - 0.5);
                    const float fv = (float)((g->nv[0] * px + g->nv[1] * py + g->nv[2] * pz + g->nv[3]) * r - 0.5);
                    const float w = CONE ? (float)(r * r) : (float)g->weight;
                    const float val = bp_sample_global(src, P.det_u, P.det_v, row_pitch, fu, fv);
                    if (val != 0.0f) acc = fmaf(w, val, acc);
                }
    }
    acc *= P.out_scale / (float)(ss * ss * ss);
    bp_store_one(P, ((size_t)z * P.ny + y) * P.nx + x, acc);
}

// Voxel supersampling through the staged kernels: vol[z][y][x] = sum of the d^3 voxels of a d-times finer grid (each
// fine voxel already carries its own, d^3 times smaller, voxel volume, so the sum is voxel volume x the mean over the
// sub-voxel centres).  Coarse slices z0 .. z0 + nzs are in `fine`, [nzs * d][ny * d][nx * d]; stored with
// bp_store_one (SET, ADD, fused SIRT update).
__global__ void __launch_bounds__(256) bp_pool_kernel(const BPArgs P, const float *__restrict__ fine, int d, int z0, int nzs)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= P.nx || y >= P.ny || z >= nzs) return;
    const size_t fx = (size_t)P.nx * d, fplane = fx * P.ny * d;
    float sum = 0.0f;
    for (int sz = 0; sz < d; ++sz)
        for (int sy = 0; sy < d; ++sy) {
            const float *src = fine + (size_t)(z * d + sz) * fplane + (size_t)(y * d + sy) * fx + (size_t)x * d;
            for (int sx = 0; sx < d; ++sx) sum += __ldg(src + sx);
        }
    bp_store_one(P, ((size_t)(z0 + z) * P.ny + y) * P.nx + x, sum);
}

}  // namespace tsp
