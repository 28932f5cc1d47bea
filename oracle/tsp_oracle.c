/*
 * tsp_oracle.c -- CPU restatement of the projector arithmetic behind
 * tomosipo's projection path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py are the only callers.
 * Nothing under tomosipo_b200/ may import, link or execute it.
 *
 * PARITY STATUS: pinned as far as the reference allows; ASTRA's own BP / cone-FP
 * values are "parity unpinned".  The arithmetic itself lives in the ASTRA
 * toolbox (astra-toolbox >= 2.0, un-vendored dependency of the reference,
 * setup.py:15), which is absent from /root/reference and from this image.
 * Pinned: the forward projector against the one real ASTRA output the
 * reference ships (notebooks/cupy.ipynb cell 4, tests/test_oracle.py) and
 * against closed-form line integrals, cone and parallel (tests/test_oracle_pins.py);
 * the backprojector's voxel -> detector map against the reference's own
 * project_point (golden vectors + tests/geometry/test_cone_vec.py:143-173); its
 * weight against closed forms, and - the open question of SURVEY.md B.2 - shown
 * to be the exact adjoint's weight times the cosine of the ray's obliquity.
 * Not pinned (no such number exists in the reference): any backprojection or
 * cone-beam value computed by ASTRA itself.  Everything else is a restatement
 * of ASTRA's published cuda3d semantics as summarised in SURVEY.md Appendix B.
 *
 * What is restated, and which reference call site consumes it:
 *   - geometry normalisation (unit voxels, centred volume):
 *       tomosipo/geometry/volume.py:286-301 (extent -> create_vol_geom),
 *       tomosipo/geometry/cone_vec.py:174-191 and parallel_vec.py:175-192
 *       (12-column vectors in (x,y,z) order: [src|ray, det centre, u, v]).
 *   - FP  (ray-driven Joseph, bilinear in-slice interpolation, zero border):
 *       tomosipo/astra.py:147-153 with "FP".
 *   - BP  (voxel-driven, bilinear detector interpolation, ray-density weight):
 *       tomosipo/astra.py:147-153 with "BP".
 *   - SET / ADD modes: tomosipo/astra.py:132-135,151.
 *   - Voxel / detector supersampling: tomosipo/astra.py:84-97.
 *
 * The file is compiled twice (REAL = double and REAL = float); the float
 * build is the multi-threaded CPU baseline that bench.py times.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif
#ifndef SUFFIX
#define SUFFIX f64
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* Same field order as tsp_geometry in include/tsproj.h so that one ctypes
 * Structure serves both libraries. */
typedef struct {
    int32_t kind; /* 0 = cone_vec, 1 = parallel3d_vec */
    int32_t nx, ny, nz;
    double win_min[3]; /* x, y, z */
    double win_max[3];
    int32_t det_rows, det_cols, n_angles;
    const double *vectors; /* n_angles x 12, ASTRA (x,y,z) order */
    int32_t voxel_supersampling, detector_supersampling;
} oracle_geometry;

typedef struct {
    double p[3];  /* source position (cone) or ray direction (parallel) */
    double dc[3]; /* detector centre */
    double u[3], v[3];
    double area;  /* |u x v| in physical units */
} nangle;

static double det3(const double *a, const double *b, const double *c)
{
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) +
           a[2] * (b[0] * c[1] - b[1] * c[0]);
}

/* translate by -centre, divide each component by the voxel
 * size of its axis.  Afterwards the volume is [-N/2, N/2] per axis. */
static void normalise(const oracle_geometry *g, int a, double sigma[3], nangle *o)
{
    const double *w = g->vectors + 12 * (size_t)a;
    int n[3] = {g->nx, g->ny, g->nz};
    double pu[3], pv[3];
    for (int i = 0; i < 3; ++i) {
        sigma[i] = (g->win_max[i] - g->win_min[i]) / n[i];
        double c = 0.5 * (g->win_max[i] + g->win_min[i]);
        o->p[i] = (g->kind == 0) ? (w[i] - c) / sigma[i] : w[i] / sigma[i];
        o->dc[i] = (w[3 + i] - c) / sigma[i];
        o->u[i] = w[6 + i] / sigma[i];
        o->v[i] = w[9 + i] / sigma[i];
        pu[i] = w[6 + i];
        pv[i] = w[9 + i];
    }
    double cx = pu[1] * pv[2] - pu[2] * pv[1];
    double cy = pu[2] * pv[0] - pu[0] * pv[2];
    double cz = pu[0] * pv[1] - pu[1] * pv[0];
    o->area = sqrt(cx * cx + cy * cy + cz * cz);
}

/* One marching axis per angle, from the central ray, in the normalised
 * frame; ties resolve x before y before z. */
static int marching_axis(const oracle_geometry *g, const nangle *n)
{
    double c[3];
    for (int i = 0; i < 3; ++i)
        c[i] = fabs(g->kind == 0 ? n->p[i] - n->dc[i] : n->p[i]);
    if (c[0] >= c[1] && c[0] >= c[2]) return 0;
    if (c[1] >= c[0] && c[1] >= c[2]) return 1;
    return 2;
}

int FN(oracle_marching_axes)(const oracle_geometry *g, int32_t *axes)
{
    for (int a = 0; a < g->n_angles; ++a) {
        double s[3];
        nangle n;
        normalise(g, a, s, &n);
        axes[a] = marching_axis(g, &n);
    }
    return 0;
}

/* ASTRA-emulation mode.  ASTRA samples through the texture unit with
 * hardware linear filtering, whose interpolation weights are 9-bit fixed point (1.8 format: the
 * fractional position is rounded to a multiple of 1/256; CUDA programming guide, "Texture
 * Fetching / Linear Filtering").  bits = 8 reproduces that; bits = 0 (default) is exact.  Used by
 * the tests to bound how far an exact-weight projector can be from ASTRA (north_star: <= 1e-3). */
static int FN(g_weight_bits) = 0;
void FN(oracle_set_weight_bits)(int bits) { FN(g_weight_bits) = bits; }
static inline REAL quant_weight(REAL w)
{
    const int bits = FN(g_weight_bits);
    if (bits <= 0) return w;
    const double q = (double)(1 << bits);
    return (REAL)(floor((double)w * q + 0.5) / q);
}

/* ------------------------------------------------------------------ FP -- */

static inline REAL vol_at(const REAL *vol, const int n[3], int ix, int iy, int iz)
{
    if (ix < 0 || iy < 0 || iz < 0 || ix >= n[0] || iy >= n[1] || iz >= n[2]) return 0;
    return vol[((size_t)iz * n[1] + iy) * n[0] + ix];
}

/* Line integral of one ray: origin o, direction d (normalised frame). */
static REAL joseph_ray(const REAL *vol, const int n[3], const double sigma[3], int m,
                       const double o[3], const double d[3])
{
    const int p = (m == 0) ? 1 : 0;
    const int q = (m == 2) ? 1 : 2;
    const REAL ap = (REAL)(d[p] / d[m]);
    const REAL aq = (REAL)(d[q] / d[m]);
    const REAL bp = (REAL)(o[p] - (d[p] / d[m]) * o[m]);
    const REAL bq = (REAL)(o[q] - (d[q] / d[m]) * o[m]);
    const REAL hp = (REAL)(0.5 * n[p] - 0.5), hq = (REAL)(0.5 * n[q] - 0.5);

    /* Slices outside [k0, k1) sample entirely outside the volume. */
    int k0 = 0, k1 = n[m];
    REAL acc = 0;
    for (int k = k0; k < k1; ++k) {
        const REAL t = (REAL)(k + 0.5 - 0.5 * n[m]);
        const REAL fp = ap * t + bp + hp;
        const REAL fq = aq * t + bq + hq;
        if (!(fp > -1 && fp < n[p] && fq > -1 && fq < n[q])) continue;
        const REAL flp = (REAL)floor((double)fp), flq = (REAL)floor((double)fq);
        const int ip = (int)flp, iq = (int)flq;
        const REAL wp = quant_weight(fp - flp), wq = quant_weight(fq - flq);
        int i0[3], i1[3], i2[3], i3[3];
        i0[m] = i1[m] = i2[m] = i3[m] = k;
        i0[p] = ip;     i0[q] = iq;
        i1[p] = ip + 1; i1[q] = iq;
        i2[p] = ip;     i2[q] = iq + 1;
        i3[p] = ip + 1; i3[q] = iq + 1;
        const REAL v00 = vol_at(vol, n, i0[0], i0[1], i0[2]);
        const REAL v10 = vol_at(vol, n, i1[0], i1[1], i1[2]);
        const REAL v01 = vol_at(vol, n, i2[0], i2[1], i2[2]);
        const REAL v11 = vol_at(vol, n, i3[0], i3[1], i3[2]);
        const REAL lo = v00 + wp * (v10 - v00);
        const REAL hi = v01 + wp * (v11 - v01);
        acc += lo + wq * (hi - lo);
    }
    const double rp = sigma[p] / sigma[m], rq = sigma[q] / sigma[m];
    const double len = sigma[m] * sqrt(1.0 + (double)ap * ap * rp * rp + (double)aq * aq * rq * rq);
    return acc * (REAL)len;
}

int FN(oracle_fp)(const oracle_geometry *g, const REAL *vol, REAL *proj, int additive)
{
    const int n[3] = {g->nx, g->ny, g->nz};
    const int U = g->det_cols, V = g->det_rows, A = g->n_angles;
    const int ss = g->detector_supersampling < 1 ? 1 : g->detector_supersampling;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int a = 0; a < A; ++a) {
        for (int iv = 0; iv < V; ++iv) {
            double sigma[3];
            nangle na;
            normalise(g, a, sigma, &na);
            const int m = marching_axis(g, &na);
            for (int iu = 0; iu < U; ++iu) {
                REAL sum = 0;
                for (int sv = 0; sv < ss; ++sv) {
                    for (int su = 0; su < ss; ++su) {
                        const double cu = iu + (su + 0.5) / ss - 0.5 * U;
                        const double cv = iv + (sv + 0.5) / ss - 0.5 * V;
                        double px[3], o[3], d[3];
                        for (int i = 0; i < 3; ++i)
                            px[i] = na.dc[i] + cu * na.u[i] + cv * na.v[i];
                        if (g->kind == 0) {
                            for (int i = 0; i < 3; ++i) { o[i] = na.p[i]; d[i] = px[i] - na.p[i]; }
                        } else {
                            for (int i = 0; i < 3; ++i) { o[i] = px[i]; d[i] = na.p[i]; }
                        }
                        sum += joseph_ray(vol, n, sigma, m, o, d);
                    }
                }
                sum /= (REAL)(ss * ss);
                REAL *dst = proj + ((size_t)iv * A + a) * U + iu;
                *dst = additive ? *dst + sum : sum;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ BP -- */

static inline REAL proj_at(const REAL *proj, int U, int V, int A, int a, int iu, int iv)
{
    if (iu < 0 || iv < 0 || iu >= U || iv >= V) return 0;
    return proj[((size_t)iv * A + a) * U + iu];
}

typedef struct {
    REAL nu[4], nv[4], dn[4]; /* x, y, z, constant */
    REAL weight;              /* parallel: 1/|u x v|; cone: folded into dn */
} bp_coef;

/* SURVEY.md B.2.  Cone:  U = det(s-d, v, s-x)/det(u, v, s-x),
 *                        V = det(u, s-d, s-x)/det(u, v, s-x),
 *                        w = det(u,v,s-d)^2 / (|u x v| det(u,v,s-x)^2).
 * Parallel:              U = det(x-d, v, r)/det(u, v, r),
 *                        V = det(u, x-d, r)/det(u, v, r),   w = 1/|u x v|.
 * d is the detector corner (pixel (0,0) lower-left), |u x v| is physical. */
typedef struct { double nu[4], nv[4], dn[4], weight; } bp_coef_d;

static void bp_coefficients_d(const oracle_geometry *g, const nangle *na, bp_coef_d *c)
{
    const int U = g->det_cols, V = g->det_rows;
    double d[3], e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) d[i] = na->dc[i] - 0.5 * U * na->u[i] - 0.5 * V * na->v[i];
    if (g->kind == 0) {
        double sd[3];
        for (int i = 0; i < 3; ++i) sd[i] = na->p[i] - d[i];
        const double scale = sqrt(na->area) / fabs(det3(na->u, na->v, sd));
        for (int i = 0; i < 3; ++i) {
            c->nu[i] = -det3(sd, na->v, e[i]) * scale;
            c->nv[i] = -det3(na->u, sd, e[i]) * scale;
            c->dn[i] = -det3(na->u, na->v, e[i]) * scale;
        }
        c->nu[3] = det3(sd, na->v, na->p) * scale;
        c->nv[3] = det3(na->u, sd, na->p) * scale;
        c->dn[3] = det3(na->u, na->v, na->p) * scale;
        c->weight = 1;
    } else {
        const double den = det3(na->u, na->v, na->p);
        for (int i = 0; i < 3; ++i) {
            c->nu[i] = det3(e[i], na->v, na->p) / den;
            c->nv[i] = det3(na->u, e[i], na->p) / den;
            c->dn[i] = 0;
        }
        c->nu[3] = -det3(d, na->v, na->p) / den;
        c->nv[3] = -det3(na->u, d, na->p) / den;
        c->dn[3] = 1;
        c->weight = 1.0 / na->area;
    }
}

static void bp_coefficients(const oracle_geometry *g, const nangle *na, bp_coef *c)
{
    bp_coef_d d;
    bp_coefficients_d(g, na, &d);
    for (int i = 0; i < 4; ++i) { c->nu[i] = (REAL)d.nu[i]; c->nv[i] = (REAL)d.nv[i]; c->dn[i] = (REAL)d.dn[i]; }
    c->weight = (REAL)d.weight;
}

int FN(oracle_bp)(const oracle_geometry *g, REAL *vol, const REAL *proj, int additive)
{
    const int n[3] = {g->nx, g->ny, g->nz};
    const int U = g->det_cols, V = g->det_rows, A = g->n_angles;
    const int ss = g->voxel_supersampling < 1 ? 1 : g->voxel_supersampling;
    bp_coef *coef = (bp_coef *)malloc(sizeof(bp_coef) * (size_t)A);
    double sigma[3];
    for (int a = 0; a < A; ++a) {
        nangle na;
        normalise(g, a, sigma, &na);
        bp_coefficients(g, &na, &coef[a]);
    }
    const REAL vox = (REAL)(sigma[0] * sigma[1] * sigma[2]);
#pragma omp parallel for collapse(2) schedule(static)
    for (int iz = 0; iz < n[2]; ++iz) {
        for (int iy = 0; iy < n[1]; ++iy) {
            for (int ix = 0; ix < n[0]; ++ix) {
                REAL acc = 0;
                for (int a = 0; a < A; ++a) {
                    const bp_coef *c = &coef[a];
                    for (int sz = 0; sz < ss; ++sz)
                        for (int sy = 0; sy < ss; ++sy)
                            for (int sx = 0; sx < ss; ++sx) {
                                const REAL x = (REAL)(ix + (sx + 0.5) / ss - 0.5 * n[0]);
                                const REAL y = (REAL)(iy + (sy + 0.5) / ss - 0.5 * n[1]);
                                const REAL z = (REAL)(iz + (sz + 0.5) / ss - 0.5 * n[2]);
                                const REAL den = c->dn[0] * x + c->dn[1] * y + c->dn[2] * z + c->dn[3];
                                const REAL r = 1 / den;
                                const REAL fu = (c->nu[0] * x + c->nu[1] * y + c->nu[2] * z + c->nu[3]) * r - (REAL)0.5;
                                const REAL fv = (c->nv[0] * x + c->nv[1] * y + c->nv[2] * z + c->nv[3]) * r - (REAL)0.5;
                                if (!(fu > -1 && fu < U && fv > -1 && fv < V)) continue;
                                const REAL flu = (REAL)floor((double)fu), flv = (REAL)floor((double)fv);
                                const int iu = (int)flu, iv = (int)flv;
                                const REAL wu = quant_weight(fu - flu), wv = quant_weight(fv - flv);
                                const REAL p00 = proj_at(proj, U, V, A, a, iu, iv);
                                const REAL p10 = proj_at(proj, U, V, A, a, iu + 1, iv);
                                const REAL p01 = proj_at(proj, U, V, A, a, iu, iv + 1);
                                const REAL p11 = proj_at(proj, U, V, A, a, iu + 1, iv + 1);
                                const REAL lo = p00 + wu * (p10 - p00);
                                const REAL hi = p01 + wu * (p11 - p01);
                                const REAL w = (g->kind == 0) ? r * r : c->weight;
                                acc += w * (lo + wv * (hi - lo));
                            }
                }
                acc *= vox / (REAL)(ss * ss * ss);
                REAL *dst = vol + ((size_t)iz * n[1] + iy) * n[0] + ix;
                *dst = additive ? *dst + acc : acc;
            }
        }
    }
    free(coef);
    return 0;
}


/* ---------------------------------------------------- full-size spot checks -- */
/* Entry points for checking a GPU result at the benchmark sizes (BASELINE.json configs[2], [3]) in
 * seconds: the arrays are the float32 arrays the GPU saw, the arithmetic is fp64, and only part of
 * the output is formed - FP for a list of angles, BP for a voxel window.  Compiled into the f64
 * build only. */
#ifdef ORACLE_MIXED

static inline double volf_at(const float *vol, const int n[3], long ix, long iy, long iz)
{
    if (ix < 0 || iy < 0 || iz < 0 || ix >= n[0] || iy >= n[1] || iz >= n[2]) return 0;
    return (double)vol[((size_t)iz * n[1] + iy) * n[0] + ix];
}

static double joseph_ray_mixed(const float *vol, const int n[3], const double sigma[3], int m, const double o[3],
                               const double d[3])
{
    const int p = (m == 0) ? 1 : 0;
    const int q = (m == 2) ? 1 : 2;
    const double ap = d[p] / d[m], aq = d[q] / d[m];
    const double bp = o[p] - ap * o[m], bq = o[q] - aq * o[m];
    const double hp = 0.5 * n[p] - 0.5, hq = 0.5 * n[q] - 0.5;
    double acc = 0;
    for (int k = 0; k < n[m]; ++k) {
        const double t = k + 0.5 - 0.5 * n[m];
        const double fp = ap * t + bp + hp, fq = aq * t + bq + hq;
        if (!(fp > -1 && fp < n[p] && fq > -1 && fq < n[q])) continue;
        const double flp = floor(fp), flq = floor(fq);
        const long ip = (long)flp, iq = (long)flq;
        const double wp = fp - flp, wq = fq - flq;
        long i0[3], i1[3], i2[3], i3[3];
        i0[m] = i1[m] = i2[m] = i3[m] = k;
        i0[p] = ip;     i0[q] = iq;
        i1[p] = ip + 1; i1[q] = iq;
        i2[p] = ip;     i2[q] = iq + 1;
        i3[p] = ip + 1; i3[q] = iq + 1;
        const double v00 = volf_at(vol, n, i0[0], i0[1], i0[2]), v10 = volf_at(vol, n, i1[0], i1[1], i1[2]);
        const double v01 = volf_at(vol, n, i2[0], i2[1], i2[2]), v11 = volf_at(vol, n, i3[0], i3[1], i3[2]);
        const double lo = v00 + wp * (v10 - v00), hi = v01 + wp * (v11 - v01);
        acc += lo + wq * (hi - lo);
    }
    const double rp = sigma[p] / sigma[m], rq = sigma[q] / sigma[m];
    return acc * sigma[m] * sqrt(1.0 + ap * ap * rp * rp + aq * aq * rq * rq);
}

/* out[i][j][u] = (A vol)[row_list[i]][angle_list[j]][u]; detector supersampling 1 */
int oracle_fp_angles_mixed(const oracle_geometry *g, const float *vol, const int32_t *angle_list, int n_list,
                           const int32_t *row_list, int n_rows, double *out)
{
    const int n[3] = {g->nx, g->ny, g->nz};
    const int U = g->det_cols, V = g->det_rows;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int j = 0; j < n_list; ++j) {
        for (int i = 0; i < n_rows; ++i) {
            double sigma[3];
            nangle na;
            normalise(g, angle_list[j], sigma, &na);
            const int m = marching_axis(g, &na);
            const int iv = row_list[i];
            for (int iu = 0; iu < U; ++iu) {
                const double cu = iu + 0.5 - 0.5 * U, cv = iv + 0.5 - 0.5 * V;
                double px[3], o[3], d[3];
                for (int k = 0; k < 3; ++k) px[k] = na.dc[k] + cu * na.u[k] + cv * na.v[k];
                if (g->kind == 0) {
                    for (int k = 0; k < 3; ++k) { o[k] = na.p[k]; d[k] = px[k] - na.p[k]; }
                } else {
                    for (int k = 0; k < 3; ++k) { o[k] = px[k]; d[k] = na.p[k]; }
                }
                out[((size_t)i * n_list + j) * U + iu] = joseph_ray_mixed(vol, n, sigma, m, o, d);
            }
        }
    }
    return 0;
}

/* Where a point of the WORLD frame (x, y, z: the frame of the window and the vectors) lands on the
 * detector of one angle, and the backprojection weight it gets there (voxel volume not included):
 * out = {U, V, w} with pixel (iv, iu) spanning [iu, iu+1) x [iv, iv+1).  This is the (U, V) level
 * of the backprojector; tests compare it with the reference's own project_point
 * (tomosipo/geometry/cone_vec.py:306-326, parallel_vec.py:313-330, tests/geometry/test_cone_vec.py:143-173). */
int oracle_bp_map(const oracle_geometry *g, int angle, const double *xyz, double *out)
{
    double sigma[3];
    nangle na;
    bp_coef_d c;
    normalise(g, angle, sigma, &na);
    bp_coefficients_d(g, &na, &c);
    double x[3];
    for (int i = 0; i < 3; ++i) x[i] = (xyz[i] - 0.5 * (g->win_min[i] + g->win_max[i])) / sigma[i];
    const double den = c.dn[0] * x[0] + c.dn[1] * x[1] + c.dn[2] * x[2] + c.dn[3];
    out[0] = (c.nu[0] * x[0] + c.nu[1] * x[1] + c.nu[2] * x[2] + c.nu[3]) / den;
    out[1] = (c.nv[0] * x[0] + c.nv[1] * x[1] + c.nv[2] * x[2] + c.nv[3]) / den;
    out[2] = (g->kind == 0) ? 1.0 / (den * den) : c.weight;
    return 0;
}

/* out[z - z0][y - y0][x - x0] = (A^T proj)[z][y][x] for the voxel window lo <= index < hi
 * (lo / hi in x, y, z order); voxel supersampling 1 */
int oracle_bp_window_mixed(const oracle_geometry *g, const float *proj, const int32_t *lo, const int32_t *hi, double *out)
{
    const int n[3] = {g->nx, g->ny, g->nz};
    const long U = g->det_cols, V = g->det_rows, A = g->n_angles;
    bp_coef_d *coef = (bp_coef_d *)malloc(sizeof(bp_coef_d) * (size_t)A);
    double sigma[3];
    for (int a = 0; a < A; ++a) {
        nangle na;
        normalise(g, a, sigma, &na);
        bp_coefficients_d(g, &na, &coef[a]);
    }
    const double vox = sigma[0] * sigma[1] * sigma[2];
    const int wx = hi[0] - lo[0], wy = hi[1] - lo[1];
#pragma omp parallel for collapse(2) schedule(static)
    for (int iz = lo[2]; iz < hi[2]; ++iz) {
        for (int iy = lo[1]; iy < hi[1]; ++iy) {
            for (int ix = lo[0]; ix < hi[0]; ++ix) {
                const double x = ix + 0.5 - 0.5 * n[0], y = iy + 0.5 - 0.5 * n[1], z = iz + 0.5 - 0.5 * n[2];
                double acc = 0;
                for (long a = 0; a < A; ++a) {
                    const bp_coef_d *c = &coef[a];
                    const double den = c->dn[0] * x + c->dn[1] * y + c->dn[2] * z + c->dn[3];
                    const double r = 1 / den;
                    const double fu = (c->nu[0] * x + c->nu[1] * y + c->nu[2] * z + c->nu[3]) * r - 0.5;
                    const double fv = (c->nv[0] * x + c->nv[1] * y + c->nv[2] * z + c->nv[3]) * r - 0.5;
                    if (!(fu > -1 && fu < U && fv > -1 && fv < V)) continue;
                    const double flu = floor(fu), flv = floor(fv);
                    const long iu = (long)flu, iv = (long)flv;
                    const double wu = fu - flu, wv = fv - flv;
#define PF(uu, vv) (((uu) < 0 || (vv) < 0 || (uu) >= U || (vv) >= V) ? 0.0 : (double)proj[((size_t)(vv) * A + a) * U + (uu)])
                    const double p00 = PF(iu, iv), p10 = PF(iu + 1, iv), p01 = PF(iu, iv + 1), p11 = PF(iu + 1, iv + 1);
#undef PF
                    const double l0 = p00 + wu * (p10 - p00), h0 = p01 + wu * (p11 - p01);
                    acc += ((g->kind == 0) ? r * r : c->weight) * (l0 + wv * (h0 - l0));
                }
                out[((size_t)(iz - lo[2]) * wy + (iy - lo[1])) * wx + (ix - lo[0])] = acc * vox;
            }
        }
    }
    free(coef);
    return 0;
}
#endif /* ORACLE_MIXED */
