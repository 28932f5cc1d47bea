// libtsproj: C ABI + host side of the B200-native projector (see include/tsproj.h).
//
// Host work per projector (once, fp64): normalise the geometry to unit voxels
// (SURVEY.md B.0), pick one marching axis per angle, group angles by
// (marching axis, volume layout), and derive the affine voxel->detector maps
// for the backprojector.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "bp_kernels.cuh"
#include "fp_kernels.cuh"
#include "fdk_kernels.cuh"
#include "peer_kernels.cuh"
#include "fp_tma_kernels.cuh"
#include "thin_kernels.cuh"
#include "tsp_internal.h"

using namespace tsp;

// ----------------------------------------------------------------- errors --
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(TSP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

// --------------------------------------------------------------- geometry --
static double det3(const double *a, const double *b, const double *c)
{
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) +
           a[2] * (b[0] * c[1] - b[1] * c[0]);
}

struct NormAngle {
    double p[3], dc[3], u[3], v[3];
    double area;  // |u x v| in physical units
};

static void normalise_angle(const tsp_projector *pr, int a, NormAngle &n)
{
    const tsp_geometry &g = pr->g;
    const double *w = pr->vectors.data() + 12 * (size_t)a;
    double pu[3], pv[3];
    for (int i = 0; i < 3; ++i) {
        const double s = pr->sigma[i];
        const double c = 0.5 * (g.win_min[i] + g.win_max[i]);
        n.p[i] = (g.kind == TSP_KIND_CONE_VEC) ? (w[i] - c) / s : w[i] / s;
        n.dc[i] = (w[3 + i] - c) / s;
        n.u[i] = w[6 + i] / s;
        n.v[i] = w[9 + i] / s;
        pu[i] = w[6 + i];
        pv[i] = w[9 + i];
    }
    const double cx = pu[1] * pv[2] - pu[2] * pv[1];
    const double cy = pu[2] * pv[0] - pu[0] * pv[2];
    const double cz = pu[0] * pv[1] - pu[1] * pv[0];
    n.area = std::sqrt(cx * cx + cy * cy + cz * cz);
}

static int pick_marching_axis(int kind, const NormAngle &n)
{
    double c[3];
    for (int i = 0; i < 3; ++i) c[i] = std::fabs(kind == TSP_KIND_CONE_VEC ? n.p[i] - n.dc[i] : n.p[i]);
    if (c[0] >= c[1] && c[0] >= c[2]) return 0;
    if (c[1] >= c[0] && c[1] >= c[2]) return 1;
    return 2;
}

static void build_bp_angle(const tsp_geometry &g, const NormAngle &n, BPAngle &o)
{
    const double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double d[3];
    for (int i = 0; i < 3; ++i) d[i] = n.dc[i] - 0.5 * g.det_cols * n.u[i] - 0.5 * g.det_rows * n.v[i];
    if (g.kind == TSP_KIND_CONE_VEC) {
        double sd[3];
        for (int i = 0; i < 3; ++i) sd[i] = n.p[i] - d[i];
        // 1/den^2 must equal det(u,v,s-d)^2 / (|u x v| det(u,v,s-x)^2)
        const double k = std::sqrt(n.area) / std::fabs(det3(n.u, n.v, sd));
        for (int i = 0; i < 3; ++i) {
            o.nu[i] = -det3(sd, n.v, e[i]) * k;
            o.nv[i] = -det3(n.u, sd, e[i]) * k;
            o.dn[i] = -det3(n.u, n.v, e[i]) * k;
        }
        o.nu[3] = det3(sd, n.v, n.p) * k;
        o.nv[3] = det3(n.u, sd, n.p) * k;
        o.dn[3] = det3(n.u, n.v, n.p) * k;
        o.weight = 1.0;
    } else {
        const double den = det3(n.u, n.v, n.p);
        for (int i = 0; i < 3; ++i) {
            o.nu[i] = det3(e[i], n.v, n.p) / den;
            o.nv[i] = det3(n.u, e[i], n.p) / den;
            o.dn[i] = 0.0;
        }
        o.nu[3] = -det3(d, n.v, n.p) / den;
        o.nv[3] = -det3(n.u, d, n.p) / den;
        o.dn[3] = 1.0;
        o.weight = 1.0 / n.area;
    }
}


// ------------------------------------------- FP footprint boxes (TMA path) --
// Host mirror of fpt_ray / fpt_slice_box (fp_tma_kernels.cuh): bounds, for one
// launch group, the box that the footprint of any 32 x 16 detector tile of an
// angle pair needs on any slice, and pairs neighbouring angles whose footprints
// nearly coincide.  An underestimate only costs speed (the kernel flags slices
// that do not fit and samples them from global memory).
struct HostRay { double ap, cp, aq, cq; };

static HostRay host_ray(bool cone, const FPAngle &g, double cu, double cv, int n_p, int n_q)
{
    const double pm = g.d0[0] + cu * g.u[0] + cv * g.v[0];
    const double pp = g.d0[1] + cu * g.u[1] + cv * g.v[1];
    const double pq = g.d0[2] + cu * g.u[2] + cv * g.v[2];
    const double dir_m = cone ? pm - g.o[0] : g.o[0], dir_p = cone ? pp - g.o[1] : g.o[1],
                 dir_q = cone ? pq - g.o[2] : g.o[2];
    const double org_m = cone ? g.o[0] : pm, org_p = cone ? g.o[1] : pp, org_q = cone ? g.o[2] : pq;
    const double a_p = dir_p / dir_m, a_q = dir_q / dir_m;
    return {a_p, org_p - a_p * org_m + 0.5 * n_p - 0.5, a_q, org_q - a_q * org_m + 0.5 * n_q - 0.5};
}

// Box (columns, rows) needed by the angle set {a, b} (b < 0: only a) over sampled tiles and slices.
static void fp_pair_need(const tsp_projector *pr, int a, int b, int n_m, int n_p, int n_q, int tile_v, int &need_w,
                         int &need_h)
{
    const tsp_geometry &g = pr->g;
    const bool cone = g.kind == TSP_KIND_CONE_VEC;
    const int tu = (g.det_cols + FPT_TU - 1) / FPT_TU, tv = (g.det_rows + tile_v - 1) / tile_v;
    const int ut[3] = {0, tu / 2, tu - 1}, vt[3] = {0, tv / 2, tv - 1};
    const double t0 = 0.5 - 0.5 * n_m;
    const int kstep = std::max(1, n_m / 48);
    need_w = need_h = 0;
    for (int iu = 0; iu < 3; ++iu)
        for (int iv = 0; iv < 3; ++iv) {
            const int u0 = ut[iu] * FPT_TU, v0 = vt[iv] * tile_v;
            const int u1 = std::min(u0 + FPT_TU, g.det_cols) - 1, v1 = std::min(v0 + tile_v, g.det_rows) - 1;
            HostRay c[8];
            int nc = 0;
            for (int slot = 0; slot < (b >= 0 ? 2 : 1); ++slot)
                for (int k = 0; k < 4; ++k)
                    c[nc++] = host_ray(cone, pr->fp_angles[slot ? b : a], ((k & 1) ? u1 : u0) + 0.5,
                                       ((k & 2) ? v1 : v0) + 0.5, n_p, n_q);
            for (int k = 0; k < n_m; k += kstep) {
                const double t = std::min(k, n_m - 1) + t0;
                double pmin = 1e300, pmax = -1e300, qmin = 1e300, qmax = -1e300;
                for (int i = 0; i < nc; ++i) {
                    const double p = c[i].ap * t + c[i].cp, q = c[i].aq * t + c[i].cq;
                    pmin = std::min(pmin, p); pmax = std::max(pmax, p);
                    qmin = std::min(qmin, q); qmax = std::max(qmax, q);
                }
                if (!(pmax > -1.0 && pmin < n_p && qmax > -1.0 && qmin < n_q)) continue;  // slice not touched
                if (!(std::isfinite(pmin) && std::isfinite(pmax) && std::isfinite(qmin) && std::isfinite(qmax))) continue;
                // columns floor(pmin) .. floor(pmax) + 1 at the worst phase of pmin on the voxel grid,
                // + 3 for aligning the start down to 16 bytes; spans vary smoothly between the
                // sampled tiles / slices: 2 % + 0.05 slack
                const double w = std::floor(1.02 * (pmax - pmin) + 0.07) + 3 + 3;
                const double h = std::floor(1.02 * (qmax - qmin) + 0.07) + 3;
                if (w < 1e6) need_w = std::max(need_w, (int)w);
                if (h < 1e6) need_h = std::max(need_h, (int)h);
            }
        }
}

static void plan_fp_tma_group(const tsp_projector *pr, FPGroup &grp)
{
    const tsp_geometry &g = pr->g;
    const int n[3] = {g.nx, g.ny, g.nz};
    const int n_m = n[grp.march], n_p = n[grp.p_axis], n_q = n[grp.q_axis];
    grp.pairs.clear();
    grp.box_w = grp.box_h = 0;
    // rows per thread: 8 amortises the per-slice synchronisation over twice the samples; small
    // detectors keep 4 so that the grid still fills the GPU.  TSP_FP_R overrides (tuning aid).
    // (the row blocks of the host pipeline are 32-128 rows of a wide detector: judged by the CTA count, not by the rows)
    {
        const long long ctas8 = (long long)((g.det_cols + FPT_TU - 1) / FPT_TU) * ((g.det_rows + 31) / 32) *
                                (long long)((grp.angles.size() + 1) / 2);
        grp.rows_per_thread = (g.det_rows >= 32 && ctas8 >= 8LL * 148) ? 8 : 4;
    }
    if (const char *e = getenv("TSP_FP_R")) grp.rows_per_thread = atoi(e) == 8 ? 8 : 4;
    const int tile_v = 4 * grp.rows_per_thread;
    int box_w = 0, box_h = 0;
    const size_t na = grp.angles.size();
    for (size_t i = 0; i < na;) {
        int w1, h1;
        fp_pair_need(pr, grp.angles[i], -1, n_m, n_p, n_q, tile_v, w1, h1);
        bool paired = false;
        if (i + 1 < na) {
            int w2, h2;
            fp_pair_need(pr, grp.angles[i], grp.angles[i + 1], n_m, n_p, n_q, tile_v, w2, h2);
            // pair only when sharing the box is a clear win: union at most ~25 % larger than one footprint
            if ((double)w2 * h2 <= 1.25 * (double)w1 * h1 + 64.0) {
                grp.pairs.push_back(grp.angles[i]);
                grp.pairs.push_back(grp.angles[i + 1]);
                box_w = std::max(box_w, w2); box_h = std::max(box_h, h2);
                i += 2;
                paired = true;
            }
        }
        if (!paired) {
            grp.pairs.push_back(grp.angles[i]);
            grp.pairs.push_back(-1);
            box_w = std::max(box_w, w1); box_h = std::max(box_h, h1);
            i += 1;
        }
    }
    box_w = (box_w + 3) / 4 * 4;  // TMA boxes are whole 16-byte units
    // test hook: an undersized box makes the kernel flag slices as unfit and sample them from
    // global memory (the results must not change)
    if (const char *e = getenv("TSP_FP_BOX_SHRINK")) {
        box_w = std::max(8, box_w - 4 * atoi(e));
        box_h = std::max(3, box_h - atoi(e));
    }
    if (box_w < 8) box_w = 8;
    if (box_w > 256 || box_h > 256) return;                 // TMA box limit
    if ((size_t)box_w * box_h * 4 > 32 * 1024) return;      // keep >= 3 ring stages
    grp.box_w = box_w;
    grp.box_h = box_h;
}

static int validate(const tsp_geometry *g)
{
    if (!g) return fail(TSP_ERR_INVALID, "geometry is NULL");
    if (g->kind != TSP_KIND_CONE_VEC && g->kind != TSP_KIND_PARALLEL_VEC)
        return fail(TSP_ERR_INVALID, "unknown geometry kind %d", g->kind);
    if (g->nx < 1 || g->ny < 1 || g->nz < 1)
        return fail(TSP_ERR_INVALID, "volume shape must be positive, got (%d, %d, %d)", g->nz, g->ny, g->nx);
    if (g->det_rows < 1 || g->det_cols < 1 || g->n_angles < 1)
        return fail(TSP_ERR_INVALID, "detector shape / angle count must be positive, got (%d, %d, %d)",
                    g->det_rows, g->n_angles, g->det_cols);
    if (!g->vectors) return fail(TSP_ERR_INVALID, "vectors is NULL");
    for (int i = 0; i < 3; ++i)
        if (!(g->win_max[i] > g->win_min[i]))
            return fail(TSP_ERR_INVALID, "volume window must have positive extent on axis %d", i);
    if (g->voxel_supersampling < 1 || g->detector_supersampling < 1)
        return fail(TSP_ERR_INVALID, "supersampling factors must be >= 1");
    for (size_t i = 0; i < (size_t)g->n_angles * 12; ++i)
        if (!std::isfinite(g->vectors[i])) return fail(TSP_ERR_INVALID, "vectors contain a non-finite value");
    return TSP_OK;
}

static int create_projector_internal(const tsp_geometry *geometry, const int *axes_override, tsp_projector **out);

extern "C" int tsp_projector_create(const tsp_geometry *geometry, tsp_projector **out)
{
    return create_projector_internal(geometry, nullptr, out);
}

// axes_override: per-angle FP marching axes to use instead of picking them from the central ray
// (sub-projectors of the host pipeline must march like their parent, whose central ray differs)
static int create_projector_internal(const tsp_geometry *geometry, const int *axes_override, tsp_projector **out)
{
    if (!out) return fail(TSP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (int rc = validate(geometry)) return rc;
    tsp_projector *pr = new tsp_projector();
    pr->g = *geometry;
    pr->vectors.assign(geometry->vectors, geometry->vectors + (size_t)geometry->n_angles * 12);
    pr->g.vectors = pr->vectors.data();
    const tsp_geometry &g = pr->g;
    const int n[3] = {g.nx, g.ny, g.nz};
    for (int i = 0; i < 3; ++i) pr->sigma[i] = (g.win_max[i] - g.win_min[i]) / n[i];

    const int A = g.n_angles;
    pr->fp_angles.resize(A);
    pr->bp_angles.resize(A);
    pr->march_axis.resize(A);

    // group key: march axis * 2 + transposed
    std::map<int, FPGroup> groups;
    for (int a = 0; a < A; ++a) {
        NormAngle na;
        normalise_angle(pr, a, na);
        build_bp_angle(g, na, pr->bp_angles[a]);
        const int m = axes_override ? axes_override[a] : pick_marching_axis(g.kind, na);
        pr->march_axis[a] = m;
        // The in-slice axis the warp's lanes (det_u) run along must be the
        // contiguous one: x in the native (z,y,x) layout, y in the (z,x,y) copy.
        int p, q;
        bool transposed;
        if (m == 0) { p = 1; q = 2; transposed = true; }
        else if (m == 1) { p = 0; q = 2; transposed = false; }
        else {
            transposed = std::fabs(na.u[1]) > std::fabs(na.u[0]);
            p = transposed ? 1 : 0;
            q = transposed ? 0 : 1;
        }
        // detector rows parallel to the q axis: all pixels of a column share their (march, p) part
        const double vn = std::sqrt(na.v[0] * na.v[0] + na.v[1] * na.v[1] + na.v[2] * na.v[2]);
        const bool columns = (std::fabs(na.v[m]) + std::fabs(na.v[p])) <= 1e-12 * vn && !getenv("TSP_FP_NO_COLS");
        FPGroup &grp = groups[(m * 2 + (transposed ? 1 : 0)) * 2 + (columns ? 1 : 0)];
        grp.march = m; grp.p_axis = p; grp.q_axis = q; grp.transposed = transposed; grp.columns = columns;
        grp.angles.push_back(a);
        FPAngle &f = pr->fp_angles[a];
        const int perm[3] = {m, p, q};
        for (int i = 0; i < 3; ++i) {
            const int s = perm[i];
            f.o[i] = na.p[s];
            f.u[i] = na.u[s];
            f.v[i] = na.v[s];
            f.d0[i] = na.dc[s] - 0.5 * g.det_cols * na.u[s] - 0.5 * g.det_rows * na.v[s];
        }
    }
    // Split every group by footprint width: the staged box of a launch is sized for its widest footprint (angles near
    // the diagonals of the slice grid), and every CTA pays the TMA fill of that box - a third of the kernel's
    // shared-memory traffic.  Near-axis angles get a launch of their own with a narrower box (and more ring stages).
    for (auto &kv : groups) {
        FPGroup &grp = kv.second;
        const int nn[3] = {g.nx, g.ny, g.nz};
        const int tile_v = 32;
        std::vector<int> w(grp.angles.size());
        int wmin = INT_MAX, wmax = 0;
        for (size_t i = 0; i < grp.angles.size(); ++i) {
            int h;
            fp_pair_need(pr, grp.angles[i], -1, nn[grp.march], nn[grp.p_axis], nn[grp.q_axis], tile_v, w[i], h);
            wmin = std::min(wmin, w[i]); wmax = std::max(wmax, w[i]);
        }
        int classes = 1;
        // (measured at cfg 3 with two slices per stage: 1 class 42.9 ms, 2 classes 43.3, 3 classes 43.5 - the tails of
        // the extra launches cost more than the narrower boxes save; kept as a tuning aid, TSP_FP_CLASSES=n)
        if (const char *e = getenv("TSP_FP_CLASSES")) classes = std::max(1, std::min(4, atoi(e)));
        if (grp.angles.size() < 96 || wmax - wmin < 12) classes = 1;
        if (classes == 1) {
            pr->groups.push_back(std::move(grp));
            continue;
        }
        std::vector<FPGroup> sub(classes, grp);
        for (auto &sg : sub) sg.angles.clear();
        for (size_t i = 0; i < grp.angles.size(); ++i) {
            const int c = std::min(classes - 1, (w[i] - wmin) * classes / (wmax - wmin + 1));
            sub[c].angles.push_back(grp.angles[i]);
        }
        for (auto &sg : sub)
            if (!sg.angles.empty()) pr->groups.push_back(std::move(sg));
    }
    for (FPGroup &grp : pr->groups) plan_fp_tma_group(pr, grp);
    if (getenv("TSP_DEBUG"))
        for (const FPGroup &grp : pr->groups)
            fprintf(stderr, "[tsp] plan: fp group march=%d transposed=%d cols=%d: %zu angles in %zu pairs, box need %dx%d, R=%d\n",
                    grp.march, (int)grp.transposed, (int)grp.columns, grp.angles.size(), grp.pairs.size() / 2, grp.box_w, grp.box_h,
                    grp.rows_per_thread);
    *out = pr;
    return TSP_OK;
}

static void free_device_state(tsp_projector *pr)
{
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    for (auto &kv : pr->dev) {
        if (cudaSetDevice(kv.first) != cudaSuccess) continue;
        cudaFree(kv.second.fp_angles);
        cudaFree(kv.second.fp_lists);
        cudaFree(kv.second.fp_pairs);
        cudaFree(kv.second.bp_angles);
        if (kv.second.s_in) cudaStreamDestroy(kv.second.s_in);
        if (kv.second.s_out) cudaStreamDestroy(kv.second.s_out);
        if (kv.second.pool && kv.second.owns_pool) cudaMemPoolDestroy(kv.second.pool);
    }
    cudaSetDevice(cur);
    pr->dev.clear();
}

extern "C" void tsp_projector_destroy(tsp_projector *pr)
{
    if (!pr) return;
    for (auto &c : pr->host_bp) tsp_projector_destroy(c.sub);
    for (auto &c : pr->host_fp) tsp_projector_destroy(c.sub);
    for (auto &c : pr->ss_bp) tsp_projector_destroy(c.sub);
    for (auto &c : pr->ss_fp) tsp_projector_destroy(c.sub);
    pr->host_bp.clear();
    pr->host_fp.clear();
    pr->ss_bp.clear();
    pr->ss_fp.clear();
    if (!pr->dev.empty()) free_device_state(pr);
    delete pr;
}

extern "C" int tsp_projector_get_info(const tsp_projector *pr, tsp_projector_info *info)
{
    if (!pr || !info) return fail(TSP_ERR_INVALID, "NULL argument");
    memset(info, 0, sizeof *info);
    info->n_angles = pr->g.n_angles;
    for (int m : pr->march_axis) {
        if (m == 0) ++info->n_march_x;
        else if (m == 1) ++info->n_march_y;
        else ++info->n_march_z;
    }
    for (int i = 0; i < 3; ++i) info->voxel_size[i] = pr->sigma[i];
    info->kernel_launches = pr->launches;
    info->bp_uses_tma = pr->bp_uses_tma;
    info->fp_uses_transpose = pr->fp_uses_transpose;
    info->fp_uses_tma = pr->fp_uses_tma;
    info->host_pipelined = pr->host_pipelined;
    info->host_ring = pr->host_ring;
    info->host_devices = pr->host_devices;
    return TSP_OK;
}

extern "C" int tsp_projector_marching_axes(const tsp_projector *pr, int32_t *axes)
{
    if (!pr || !axes) return fail(TSP_ERR_INVALID, "NULL argument");
    for (size_t i = 0; i < pr->march_axis.size(); ++i) axes[i] = pr->march_axis[i];
    return TSP_OK;
}

extern "C" int tsp_projector_bp_map(const tsp_projector *pr, int angle, const double *xyz, double *out)
{
    if (!pr || !xyz || !out) return fail(TSP_ERR_INVALID, "NULL argument");
    if (angle < 0 || angle >= pr->g.n_angles) return fail(TSP_ERR_INVALID, "angle %d out of range", angle);
    const BPAngle &m = pr->bp_angles[angle];
    double x[3];
    for (int i = 0; i < 3; ++i) x[i] = (xyz[i] - 0.5 * (pr->g.win_min[i] + pr->g.win_max[i])) / pr->sigma[i];
    const double den = m.dn[0] * x[0] + m.dn[1] * x[1] + m.dn[2] * x[2] + m.dn[3];
    out[0] = (m.nu[0] * x[0] + m.nu[1] * x[1] + m.nu[2] * x[2] + m.nu[3]) / den;
    out[1] = (m.nv[0] * x[0] + m.nv[1] * x[1] + m.nv[2] * x[2] + m.nv[3]) / den;
    out[2] = pr->g.kind == TSP_KIND_CONE_VEC ? 1.0 / (den * den) : m.weight;
    return TSP_OK;
}

extern "C" int tsp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int tsp_cuda_available(void) { return tsp_device_count() > 0 ? 1 : 0; }
extern "C" int tsp_version(void) { return TSP_VERSION; }
extern "C" const char *tsp_last_error(void) { return g_last_error.c_str(); }

// ------------------------------------------------------------ device side --
static int get_device_state(tsp_projector *pr, int device, DeviceState **out)
{
    std::lock_guard<std::mutex> lock(pr->mu);
    auto it = pr->dev.find(device);
    if (it != pr->dev.end()) {
        *out = &it->second;
        return TSP_OK;
    }
    DeviceState st;
    const size_t A = pr->g.n_angles;
    CUDA_TRY(cudaMalloc(&st.fp_angles, A * sizeof(FPAngle)));
    CUDA_TRY(cudaMalloc(&st.bp_angles, A * sizeof(BPAngle)));
    CUDA_TRY(cudaMalloc(&st.fp_lists, A * sizeof(int)));
    CUDA_TRY(cudaMemcpy(st.fp_angles, pr->fp_angles.data(), A * sizeof(FPAngle), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(st.bp_angles, pr->bp_angles.data(), A * sizeof(BPAngle), cudaMemcpyHostToDevice));
    std::vector<int> lists;
    for (const FPGroup &grp : pr->groups) {
        st.list_offset.push_back(lists.size());
        lists.insert(lists.end(), grp.angles.begin(), grp.angles.end());
    }
    CUDA_TRY(cudaMemcpy(st.fp_lists, lists.data(), lists.size() * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> pairs;
    for (const FPGroup &grp : pr->groups) {
        st.pair_offset.push_back(pairs.size() / 2);
        pairs.insert(pairs.end(), grp.pairs.begin(), grp.pairs.end());
    }
    CUDA_TRY(cudaMalloc(&st.fp_pairs, std::max<size_t>(1, pairs.size()) * sizeof(int)));
    CUDA_TRY(cudaMemcpy(st.fp_pairs, pairs.data(), pairs.size() * sizeof(int), cudaMemcpyHostToDevice));
    // Private pool: caches one transposed-volume scratch between calls, nothing more (see pool_alloc).  Sub-projectors
    // (host pipeline chunks, supersampling slabs) allocate from the pool of the projector they belong to - one cache
    // per user-visible projector, not one per chunk.
    const int ny_pad = (pr->g.ny + 3) / 4 * 4;
    const uint64_t base = (uint64_t)pr->g.nz * pr->g.nx * ny_pad * sizeof(float);
    if (pr->pool_owner) {
        DeviceState *ost = nullptr;
        if (int rc = get_device_state(pr->pool_owner, device, &ost)) return rc;
        st.pool = ost->pool;
        st.owns_pool = false;
        st.pool_st = ost;
    } else {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&st.pool, &props) == cudaSuccess) {
            uint64_t keep = base + base / 4 + ((uint64_t)32 << 20);
            if (const char *e = getenv("TSP_POOL_KEEP_MB")) keep = (uint64_t)std::max(0LL, atoll(e)) << 20;
            cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep);
            st.pool_keep = st.pool_keep_base = (size_t)base;
        } else {
            cudaGetLastError();
            st.pool = nullptr;  // falls back to the device's default pool, untouched
        }
    }
    *out = &(pr->dev[device] = st);
    return TSP_OK;
}


// ---- device memory of a call: a private stream-ordered pool per (projector, device) ------------------
// The default pool is left alone (raising its release threshold is a process-wide side effect
// and keeps memory away from other allocators).  The private pool keeps at most `keep` bytes cached between
// calls - enough for the per-call scratch of an iterative loop (the transposed volume copy of launch_fp) -
// and hands everything above that back to the driver at the next synchronisation point.
static cudaError_t pool_alloc(DeviceState *st, void **p, size_t bytes, cudaStream_t stream)
{
    if (st->pool) return cudaMallocFromPoolAsync(p, bytes, st->pool, stream);
    return cudaMallocAsync(p, bytes, stream);
}

// Host-array calls stage whole arrays (or a ring of chunks) on the device: let the pool keep that much between calls,
// so that an iterative loop over host arrays (the README SIRT of the reference) does not map and unmap gigabytes per
// call (measured: 1807 -> 1127 GUPS end to end without this).  The bound is what this projector itself needs for one
// call; tsp_projector_destroy gives everything back.
static void pool_set_threshold(DeviceState *owner)
{
    // the pool reserves whole 2 MB granules per allocation: a threshold equal to the bytes requested is exceeded by the
    // rounding, and the excess block would be unmapped and mapped again on every call (measured: 1.8 instead of 0.65 ms
    // per 18 MB host-array call, r02 GPU call 10) - keep a quarter plus 32 MB of slack
    const uint64_t total = (uint64_t)owner->pool_keep + owner->pool_extra;
    uint64_t keep = total + total / 4 + ((uint64_t)32 << 20);
    if (cudaMemPoolSetAttribute(owner->pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
}

static void pool_keep_at_least(DeviceState *st, size_t bytes)
{
    if (bytes <= st->pool_keep) return;
    if (st->pool_st) {  // a sub-projector: its scratch comes on top of what the owner and the other sub-projectors keep
        if (!st->pool_st->pool) return;
        st->pool_st->pool_extra += bytes - st->pool_keep;
        st->pool_keep = bytes;
        pool_set_threshold(st->pool_st);
        return;
    }
    if (!st->pool) return;
    st->pool_keep = bytes;
    pool_set_threshold(st);
}

// Per-call scratch from the projector's private pool, released (stream-ordered) on every exit path.
struct PoolScratch {
    void *p = nullptr;
    cudaStream_t stream = nullptr;
    ~PoolScratch() { if (p) cudaFreeAsync(p, stream); }
};

// defined in the TMA section below
static bool make_tensor_map_3d(const float *base, const uint64_t dims[3], const uint64_t stride_bytes[2],
                               const uint32_t box[3], TensorMapBlob *out);

// supersampling through the staged kernels (defined with the sub-projector machinery further down)
static int launch_fp_supersampled(tsp_projector *pr, DeviceState *st, const float *vol, float *proj, int additive, cudaStream_t stream,
                                  const float *epi_sub, const float *epi_mul);
static int launch_bp_supersampled(tsp_projector *pr, DeviceState *st, float *vol, const float *proj, int additive, cudaStream_t stream,
                                  const float *epi_mul);

template <bool CONE, bool COLS, int R, int SPS>
static int launch_fp_tma_one(dim3 grid, size_t smem, cudaStream_t stream, const FPTmaArgs &A, const TensorMapPair &tmap)
{
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && configured[dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(fp_tma_kernel<CONE, COLS, R, SPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    fp_tma_kernel<CONE, COLS, R, SPS><<<grid, FPT_THREADS, smem, stream>>>(A, tmap);
    return TSP_OK;
}

static inline bool last_seg(int seg, int segments) { return seg == segments - 1; }

static int launch_fp(tsp_projector *pr, DeviceState *st, const float *vol, float *proj, int additive,
                     cudaStream_t stream, const float *epi_sub = nullptr, const float *epi_mul = nullptr, int batch = 1,
                     const float *vol_t_ext = nullptr,  // caller-made (x <-> y)-transposed copy (tsp_transpose_slices)
                     const FPPeers *peers = nullptr)    // multi-GPU: band buffers the store also writes (tsp_fp_push)
{
    const tsp_geometry &g = pr->g;
    if (peers && (batch != 1 || additive || g.detector_supersampling > 1))
        return fail(TSP_ERR_INVALID, "the fused row exchange takes one plain (SET, no supersampling) forward projection");
    const int n[3] = {g.nx, g.ny, g.nz};
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    const size_t npix = (size_t)g.det_rows * g.n_angles * g.det_cols;
    // thin detectors (cfg 5 slabs): one launch per angle group covers the whole batch (thin_kernels.cuh)
    const bool thin = g.det_rows <= THIN_MAX && g.detector_supersampling == 1 && !getenv("TSP_NO_THIN") &&
                      (long long)batch * g.det_rows <= 65535 && (size_t)batch * g.nz * (g.nx + 4) * (g.ny + 4) < (1ull << 31) &&  // 32-bit offsets, either layout
                     
                      g.nx >= 2 && g.ny >= 2;  // the thin kernels shift their tap window into the array
    if (!thin && batch > 1) {
        for (int b = 0; b < batch; ++b)
            if (int rc = launch_fp(pr, st, vol + b * nvox, proj + b * npix, additive, stream,
                                   epi_sub ? epi_sub + b * npix : nullptr, epi_mul ? epi_mul + b * npix : nullptr, 1))
                return rc;
        return TSP_OK;
    }

    if (g.detector_supersampling > 1 && !thin && !getenv("TSP_SS_DIRECT")) {
        const int rc = launch_fp_supersampled(pr, st, vol, proj, additive, stream, epi_sub, epi_mul);
        if (rc != 1) return rc;  // 1: not applicable, take the direct kernels
    }
    bool need_t = false;
    for (const FPGroup &grp : pr->groups) need_t |= grp.transposed;
    // the transpose grid covers nz * batch planes: a batch too tall for it goes item by item
    if (need_t && batch > 1 && (long long)g.nz * batch > 65535) {
        for (int b = 0; b < batch; ++b)
            if (int rc = launch_fp(pr, st, vol + b * nvox, proj + b * npix, additive, stream,
                                   epi_sub ? epi_sub + b * npix : nullptr, epi_mul ? epi_mul + b * npix : nullptr, 1))
                return rc;
        return TSP_OK;
    }
    if (need_t && g.nz > 65535) return fail(TSP_ERR_INVALID, "nz exceeds the transpose grid");
    float *vol_t = nullptr;
    PoolScratch scratch;
    const int ny_pad = (g.ny + 3) / 4 * 4;  // row pitch of the transposed copy: whole 16-byte units (TMA stride rule)
    if (need_t && vol_t_ext && batch == 1) {
        vol_t = const_cast<float *>(vol_t_ext);
    } else if (need_t) {
        const size_t scratch_bytes = (size_t)batch * g.nz * g.nx * ny_pad * sizeof(float);
        if (scratch_bytes > st->pool_keep_base) {  // batched calls: keep the batch's scratch cached between calls
            pool_keep_at_least(st, st->pool_keep + (scratch_bytes - st->pool_keep_base));
            st->pool_keep_base = scratch_bytes;
        }
        CUDA_TRY(pool_alloc(st, &scratch.p, scratch_bytes, stream));
        scratch.stream = stream;
        vol_t = (float *)scratch.p;
        dim3 grid((g.nx + 31) / 32, (g.ny + 31) / 32, g.nz * batch), block(32, 8);  // batch items are contiguous planes
        transpose_xy_kernel<<<grid, block, 0, stream>>>(vol, vol_t, g.nx, g.ny, ny_pad);
        ++pr->launches;
    }
    pr->fp_uses_transpose = need_t ? 1 : 0;
    int used_tma = 0;

    // element strides of x, y, z in the two layouts
    const long long stride_native[3] = {1, g.nx, (long long)g.nx * g.ny};
    const long long stride_transp[3] = {ny_pad, 1, (long long)g.nx * ny_pad};
    for (size_t gi = 0; gi < pr->groups.size(); ++gi) {
        const FPGroup &grp = pr->groups[gi];
        const long long *stride = grp.transposed ? stride_transp : stride_native;
        FPArgs P;
        P.vol = grp.transposed ? vol_t : vol;
        P.stride_m = stride[grp.march];
        P.stride_q = stride[grp.q_axis];
        P.n_m = n[grp.march]; P.n_p = n[grp.p_axis]; P.n_q = n[grp.q_axis];
        P.angles = st->fp_angles;
        P.list = st->fp_lists + st->list_offset[gi];
        P.proj = proj;
        P.det_u = g.det_cols; P.det_v = g.det_rows; P.n_angles = g.n_angles;
        P.additive = additive;
        P.epi_sub = epi_sub; P.epi_mul = epi_mul;
        if (peers) P.peers = *peers;
        else P.peers.n = 0;
        P.det_ss = g.detector_supersampling;
        P.sigma_m = (float)pr->sigma[grp.march];
        const double rp = pr->sigma[grp.p_axis] / pr->sigma[grp.march];
        const double rq = pr->sigma[grp.q_axis] / pr->sigma[grp.march];
        P.rp2 = (float)(rp * rp);
        P.rq2 = (float)(rq * rq);
        P.offsets_fit_32bit = (size_t)g.nz * g.nx * std::max(g.ny, ny_pad) < (1ull << 31) ? 1 : 0;
        const bool cone = g.kind == TSP_KIND_CONE_VEC;
        const bool ss = g.detector_supersampling > 1;
        if (thin) {
            P.det_ss = 1;
            const int na = (int)grp.angles.size();
            if ((na + THIN_FP_ANGLES - 1) / THIN_FP_ANGLES > 65535) return fail(TSP_ERR_INVALID, "too many angles for the thin FP grid");
            const int bt = batch >= THIN_BT ? THIN_BT : 1;  // batch items per thread
            const int groups = (batch + bt - 1) / bt;
            dim3 tgrid((g.det_cols + 31) / 32, (na + THIN_FP_ANGLES - 1) / THIN_FP_ANGLES, groups * g.det_rows);
            dim3 tblock(32, THIN_FP_ANGLES);
            const size_t vstride = grp.transposed ? (size_t)g.nz * g.nx * ny_pad : nvox;
#define TSP_THIN_FP(C) (bt == 1 ? fp_thin_kernel<C, 1><<<tgrid, tblock, 0, stream>>>(P, na, batch, vstride, npix) \
                                : fp_thin_kernel<C, THIN_BT><<<tgrid, tblock, 0, stream>>>(P, na, batch, vstride, npix))
            if (cone) TSP_THIN_FP(true);
            else TSP_THIN_FP(false);
#undef TSP_THIN_FP
            ++pr->launches;
            continue;
        }
        // ---- TMA-staged kernel (default): needs a 16-byte aligned base and row pitch
        if (!ss && grp.box_w > 0 && !getenv("TSP_FP_NO_TMA") && grp.pairs.size() / 2 <= 65535) {
            const int n_second = grp.transposed ? g.nx : g.ny;  // layout dims: (p, second, nz)
            const int pitch_p = grp.transposed ? ny_pad : g.nx;
            const bool middle = grp.march != 2;                // marching along the middle layout axis?
            const uint64_t dims[3] = {(uint64_t)P.n_p, (uint64_t)n_second, (uint64_t)g.nz};
            const uint64_t strides[2] = {(uint64_t)pitch_p * 4, (uint64_t)pitch_p * 4 * (uint64_t)n_second};
            // Slices per ring stage: 2 halves the per-slice hand-off (one barrier round trip, control word and TMA issue
            // per stage; measured 46.7 -> 43.0 ms per FP at cfg 3); it needs three such stages in shared memory.
            // TSP_FP_SPS=1: one slice per stage.
            const int mult_of = middle ? 1 : 0;
            int sps = 2, need_w = 0, box_h = 0;
            int bw[2] = {0, 0};
            if (const char *e = getenv("TSP_FP_SPS")) sps = atoi(e) == 1 ? 1 : 2;
            if (P.n_m < 2) sps = 1;
            const size_t budget = (size_t)(grp.rows_per_thread == 8 ? 112 : 72) * 1024;
            for (;; sps = 1) {
                // the footprint moves by at most one voxel per slice in p and in q (the marching axis dominates the ray)
                need_w = (grp.box_w + (sps - 1) + 3) / 4 * 4;
                box_h = grp.box_h + (sps - 1);
                // two pitch variants: box widths >= the needed width for which the row stride of a staged slice
                // (box width x slices per stage when the marching axis is the middle tensor dimension) has a residue
                // mod 32 banks of {4, 8} (column and row move together along a warp) or {24, 28} (opposite)
                const int mult = mult_of ? sps : 1;
                bw[0] = bw[1] = 0;
                for (int w = need_w; w <= need_w + 32 && !(bw[0] && bw[1]); w += 4) {
                    const int r = (w * mult) & 31;
                    if (!bw[0] && (r == 4 || r == 8)) bw[0] = w;
                    if (!bw[1] && (r == 24 || r == 28)) bw[1] = w;
                }
                if (getenv("TSP_FP_ONE_PITCH")) bw[0] = bw[1] = need_w;
                // (the wider box costs L2 -> shared traffic, which has headroom: 26 % of the crossbar peak
                // at cfg 3, while the shared-memory pipe is the busiest unit of this kernel)
                if (!bw[0]) bw[0] = bw[1] ? bw[1] : need_w;
                if (!bw[1]) bw[1] = bw[0];
                if (bw[0] > 256 || bw[1] > 256) bw[0] = bw[1] = need_w;
                size_t stage = ((size_t)std::max(bw[0], bw[1]) * box_h * sps * 4 + 127) / 128 * 128 + 24;
                if (sps == 1 || 3 * stage + 160 <= budget) break;
                // tall boxes (row blocks far from the mid-plane: the host pipeline's sub-projectors): one pitch for both
                // variants - a few more bank conflicts for half of the CTAs - rather than one slice per stage
                bw[0] = bw[1] = std::min(bw[0], bw[1]);
                stage = ((size_t)bw[0] * box_h * sps * 4 + 127) / 128 * 128 + 24;
                if (3 * stage + 160 <= budget) break;
            }
            TensorMapPair slot;
            bool ok = box_h <= 256;
            for (int v = 0; v < 2 && ok; ++v) {
                const uint32_t box[3] = {(uint32_t)bw[v], middle ? (uint32_t)sps : (uint32_t)box_h, middle ? (uint32_t)box_h : (uint32_t)sps};
                ok = make_tensor_map_3d(P.vol, dims, strides, box, &slot.m[v]);
            }
            if (ok) {
                FPTmaArgs T;
                T.a = P;
                T.pairs = st->fp_pairs + 2 * st->pair_offset[gi];
                T.box_h = box_h;
                T.sps = sps;
                for (int v = 0; v < 2; ++v) {
                    T.box_w[v] = bw[v];
                    const uint32_t rs = (uint32_t)bw[v] * (middle ? (uint32_t)sps : 1u);  // words between q rows of one slice
                    T.row_stride4[v] = 4u * rs;
                    T.slice_off4[v] = 4u * (middle ? (uint32_t)bw[v] : (uint32_t)bw[v] * (uint32_t)box_h);
                    T.magic_off[v] = 0u - 4u * 0x4B400000u * (rs + 1u);
                }
                T.march_is_middle = middle ? 1 : 0;
                T.stage_bytes = ((uint32_t)std::max(bw[0], bw[1]) * box_h * sps * 4u + 127u) / 128u * 128u;
                const int R = grp.rows_per_thread;
                int stages = (int)(((R == 8 ? (sps == 2 ? 112u : 108u) : 72u) * 1024u) / (T.stage_bytes + 24u));
                if (const char *e = getenv("TSP_FP_STAGES")) stages = atoi(e);
                T.stages = std::max(2, std::min(stages, 12));
                if (getenv("TSP_DEBUG"))
                    fprintf(stderr, "[tsp] fp group march=%d transposed=%d cols=%d R=%d: %zu pairs, box need %dx%d, pitches %d/%d, %d slices per stage, %d stages of %u B\n",
                            grp.march, (int)grp.transposed, (int)grp.columns, R, grp.pairs.size() / 2, grp.box_w, grp.box_h,
                            bw[0], bw[1], sps, T.stages, T.stage_bytes);
                const size_t smem = 128 + (size_t)T.stages * (T.stage_bytes + 24) + 16;
                dim3 tgrid((g.det_cols + FPT_TU - 1) / FPT_TU, (unsigned)(grp.pairs.size() / 2),
                           (g.det_rows + 4 * R - 1) / (4 * R));
                // segments of the marching axis (FPTmaArgs::m_begin): as many as keep the slab one detector row tile
                // reads - every in-plane position x the staged box's q rows - within 72 MB (cfg 3: 53 MB, one segment,
                // 97.8 % L2 hits; cfg 4: 210 MB -> 3 segments: 811 / 705 / 697 / 716 / 766 ms with 1 / 2 / 3 / 4 / 6, r02 GPU call 39)
                const double slab_bytes = (double)P.n_p * P.n_m * box_h * 4.0;
                int segments = (int)std::min(8.0, std::ceil(slab_bytes / (72.0 * 1024 * 1024)));
                if (const char *e = getenv("TSP_FP_SEGMENTS")) segments = atoi(e);
                segments = std::max(1, std::min(segments, std::max(1, P.n_m / 32)));
                for (int seg = 0; seg < segments; ++seg) {
                    T.m_begin = (int)((long long)P.n_m * seg / segments) & ~1;  // even: stages of two slices stay inside
                    T.m_end = last_seg(seg, segments) ? P.n_m : (int)((long long)P.n_m * (seg + 1) / segments) & ~1;
                    // the epilogue and the peers' copies belong to the last segment, which holds the whole sum
                    const bool last = seg == segments - 1;
                    T.a.additive = seg == 0 ? P.additive : (P.additive == 1 ? 1 : 2);
                    T.a.epi_sub = last ? P.epi_sub : nullptr;
                    T.a.epi_mul = last ? P.epi_mul : nullptr;
                    T.a.peers.n = last ? P.peers.n : 0;
                    int rc;
#define TSP_FPT_S(C, L, S) (R == 8 ? launch_fp_tma_one<C, L, 8, S>(tgrid, smem, stream, T, slot) \
                                   : launch_fp_tma_one<C, L, 4, S>(tgrid, smem, stream, T, slot))
#define TSP_FPT(C, L) (sps == 2 ? TSP_FPT_S(C, L, 2) : TSP_FPT_S(C, L, 1))
                    if (cone) rc = grp.columns ? TSP_FPT(true, true) : TSP_FPT(true, false);
                    else rc = grp.columns ? TSP_FPT(false, true) : TSP_FPT(false, false);
#undef TSP_FPT
#undef TSP_FPT_S
                    if (rc) return rc;
                    ++pr->launches;
                }
                used_tma = 1;
                continue;
            }
        }
        // gridDim.y is limited to 65535: chunk the angle list
        for (size_t off = 0; off < grp.angles.size(); off += 65535) {
            const int na = (int)std::min<size_t>(65535, grp.angles.size() - off);
            FPArgs Q = P;
            Q.list = P.list + off;
            dim3 block(FP_BU, FP_BV);
            if (grp.columns && !ss) {
                const int rows_per_cta = FP_BV * FP_COLS_R;
                dim3 cgrid((g.det_cols + FP_BU - 1) / FP_BU, na, (g.det_rows + rows_per_cta - 1) / rows_per_cta);
                if (cone) fp_cols_kernel<true><<<cgrid, block, 0, stream>>>(Q);
                else fp_cols_kernel<false><<<cgrid, block, 0, stream>>>(Q);
                ++pr->launches;
                continue;
            }
            dim3 grid((g.det_cols + FP_BU - 1) / FP_BU, na, (g.det_rows + FP_BV - 1) / FP_BV);
            if (cone && !ss) fp_kernel<true, false><<<grid, block, 0, stream>>>(Q);
            else if (cone && ss) fp_kernel<true, true><<<grid, block, 0, stream>>>(Q);
            else if (!cone && !ss) fp_kernel<false, false><<<grid, block, 0, stream>>>(Q);
            else fp_kernel<false, true><<<grid, block, 0, stream>>>(Q);
            ++pr->launches;
        }
    }
    pr->fp_uses_tma = used_tma;
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

// z voxels per thread: long register runs amortise per-angle set-up and footprint staging
// (cfg 3: 32 -> 61.0 ms, 16 -> 69.8 ms); thin or small volumes use short runs.  TSP_BP_ZPT
// overrides (tuning aid).
static int bp_zpt_choice(int nx, int ny, int nz)
{
    if (const char *e = getenv("TSP_BP_ZPT")) {
        const int v = atoi(e);
        if (v == 1 || v == 4 || v == 8 || v == 16 || v == 24 || v == 32 || v == 64) return v;
    }
    // the tall tile (32 x 16 x 64, one CTA per SM): 46.2 vs 47.8 ms at cfg 3 (r02 GPU call 18) - when the volume is
    // at most 1/8 padding along z and gives every SM at least 6 such CTAs with a last wave that is not mostly empty
    // (TMA kernel only; the z-chunks of the multi-GPU operator qualify, the 64-slice slabs of the host pipeline do not)
    if (!getenv("TSP_BP_NO_TALL") && (nz % 64 == 0 || nz % 64 >= 56)) {
        const long long ctas = (long long)((nx + BP_TX - 1) / BP_TX) * ((ny + 15) / 16) * ((nz + 63) / 64);
        const long long waves = (ctas + 147) / 148;
        if (ctas >= 6LL * 148 && ctas * 100 >= waves * 148 * 93) return 64;
    }
    // longest run that (a) is not mostly padding and (b) still leaves >= 4 CTAs per SM to balance
    const int cand[5] = {32, 16, 8, 4, 1};
    const long long tiles_xy = (long long)((nx + BP_TX - 1) / BP_TX) * ((ny + BP_TY - 1) / BP_TY);
    int fallback = 1;
    for (int i = 0; i < 5; ++i) {
        const int z = cand[i];
        if (z > 1 && nz < (3 * z) / 4) continue;  // run mostly outside the volume
        if (fallback == 1) fallback = z;
        if (tiles_xy * ((nz + z - 1) / z) >= 4LL * 148) return z;
    }
    return std::min(fallback, 8);
}

template <bool CONE, int ZPT>
static int launch_bp_one(dim3 grid, dim3 block, cudaStream_t stream, const BPArgs &P)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = bp_smem_bytes(ZPT);
    if (dev < 64 && !configured[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(bp_kernel<CONE, ZPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    bp_kernel<CONE, ZPT><<<grid, block, smem, stream>>>(P);
    return TSP_OK;
}

static int launch_bp_variant(bool cone, int zpt, dim3 grid, dim3 block, cudaStream_t stream, const BPArgs &P)
{
#define TSP_BP_CASE(Z)                                                          \
    case Z:                                                                     \
        return cone ? launch_bp_one<true, Z>(grid, block, stream, P)            \
                    : launch_bp_one<false, Z>(grid, block, stream, P);
    switch (zpt) {
        TSP_BP_CASE(1)
        TSP_BP_CASE(4)
        TSP_BP_CASE(8)
        TSP_BP_CASE(16)
        TSP_BP_CASE(32)
    }
#undef TSP_BP_CASE
    return fail(TSP_ERR_INVALID, "unsupported z run %d", zpt);
}

// ------------------------------------------------------------ TMA staging --
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

static PFN_tensorMapEncodeTiled tensor_map_encoder()
{
    static PFN_tensorMapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tensorMapEncodeTiled)p;
        else
            cudaGetLastError();
    }
    return fn;
}

static bool make_tensor_map_3d(const float *base, const uint64_t dims[3], const uint64_t stride_bytes[2],
                               const uint32_t box[3], TensorMapBlob *out)
{
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "CUtensorMap is 128 bytes");
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
    if ((stride_bytes[0] & 15u) != 0 || (stride_bytes[1] & 15u) != 0) return false;
    if (stride_bytes[0] >= (1ull << 40) || stride_bytes[1] >= (1ull << 40)) return false;
    if (box[0] > 256 || box[1] > 256 || box[2] > 256 || (box[0] & 3u) != 0) return false;
    const cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
    const cuuint64_t st[2] = {stride_bytes[0], stride_bytes[1]};
    const cuuint32_t bx[3] = {box[0], box[1], box[2]};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base),
                     d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Tensor map over the projection stack viewed as (u, angle, v), box = cols x 1 x rows.
static bool make_proj_tensor_map(const float *proj, int det_u, int n_angles, int det_v, int box_cols, int box_rows,
                                 TensorMapBlob *out)
{
    const uint64_t dims[3] = {(uint64_t)det_u, (uint64_t)n_angles, (uint64_t)det_v};
    const uint64_t strides[2] = {(uint64_t)det_u * 4, (uint64_t)det_u * 4 * (uint64_t)n_angles};
    const uint32_t box[3] = {(uint32_t)box_cols, 1u, (uint32_t)box_rows};
    return make_tensor_map_3d(proj, dims, strides, box, out);
}

template <bool CONE, int ZPT>
static int launch_bp_tma_one(dim3 grid, cudaStream_t stream, const BPArgs &P, const TensorMapPair &tmap)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = bp_tma_smem_bytes(ZPT);
    if (dev < 64 && !configured[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(bp_tma_kernel<CONE, ZPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    bp_tma_kernel<CONE, ZPT><<<grid, bp_tma_threads(ZPT), smem, stream>>>(P, tmap);
    return TSP_OK;
}

static int launch_bp_tma_variant(bool cone, int zpt, dim3 grid, cudaStream_t stream, const BPArgs &P,
                                 const TensorMapPair &tmap)
{
#define TSP_BP_CASE(Z)                                                             \
    case Z:                                                                        \
        return cone ? launch_bp_tma_one<true, Z>(grid, stream, P, tmap)            \
                    : launch_bp_tma_one<false, Z>(grid, stream, P, tmap);
    switch (zpt) {
        TSP_BP_CASE(1)
        TSP_BP_CASE(4)
        TSP_BP_CASE(8)
        TSP_BP_CASE(16)
        TSP_BP_CASE(24)
        TSP_BP_CASE(32)
        TSP_BP_CASE(64)
    }
#undef TSP_BP_CASE
    return fail(TSP_ERR_INVALID, "unsupported z run %d", zpt);
}

static int launch_bp(tsp_projector *pr, DeviceState *st, float *vol, const float *proj, int additive,
                     cudaStream_t stream, const float *epi_mul = nullptr, int batch = 1)
{
    const tsp_geometry &g = pr->g;
    const size_t nvox_b = (size_t)g.nx * g.ny * g.nz;
    const size_t npix_b = (size_t)g.det_rows * g.n_angles * g.det_cols;
    // thin volumes (cfg 5 slabs): one launch covers the whole batch (thin_kernels.cuh)
    const bool thin = g.nz <= THIN_MAX && g.voxel_supersampling == 1 && !getenv("TSP_NO_THIN") && batch <= 65535 &&
                      (size_t)batch * npix_b < (1ull << 31) &&  // 32-bit offsets
                      g.det_cols >= 2;  // the thin kernels shift their tap window into the array
    if (!thin && batch > 1) {
        for (int b = 0; b < batch; ++b)
            if (int rc = launch_bp(pr, st, vol + b * nvox_b, proj + b * npix_b, additive, stream,
                                   epi_mul ? epi_mul + b * nvox_b : nullptr, 1))
                return rc;
        return TSP_OK;
    }
    BPArgs P;
    P.proj = proj; P.vol = vol;
    P.nx = g.nx; P.ny = g.ny; P.nz = g.nz;
    P.det_u = g.det_cols; P.det_v = g.det_rows; P.n_angles = g.n_angles;
    P.angles = st->bp_angles;
    P.out_scale = (float)(pr->sigma[0] * pr->sigma[1] * pr->sigma[2]);
    P.additive = additive;
    P.vox_ss = g.voxel_supersampling;
    P.magic_off = 0;
    P.magic_off_b = 0;
    P.no_rows3 = getenv("TSP_BP_NO_ROWS3") ? 1 : 0;
    P.rows_loop = 3;
    if (const char *e = getenv("TSP_BP_ROWS")) P.rows_loop = atoi(e) == 2 ? 2 : 3;
    P.epi_mul = epi_mul;
    const bool cone = g.kind == TSP_KIND_CONE_VEC;
    int used_tma = 0;
    if (thin) {
        const int gy = (g.ny + BP_TY - 1) / BP_TY;
        if (gy > 65535) return fail(TSP_ERR_INVALID, "volume too large for the BP grid");
        const int bt = batch >= THIN_BT ? THIN_BT : 1;  // batch items per thread
        dim3 grid((g.nx + BP_TX - 1) / BP_TX, gy, (batch + bt - 1) / bt), block(BP_TX, BP_TY);
#define TSP_THIN_BP(C) (bt == 1 ? bp_thin_kernel<C, 1><<<grid, block, 0, stream>>>(P, batch, nvox_b, npix_b) \
                                : bp_thin_kernel<C, THIN_BT><<<grid, block, 0, stream>>>(P, batch, nvox_b, npix_b))
        if (cone) TSP_THIN_BP(true);
        else TSP_THIN_BP(false);
#undef TSP_THIN_BP
    } else if (g.voxel_supersampling > 1) {
        if (!getenv("TSP_SS_DIRECT")) {
            const int rc = launch_bp_supersampled(pr, st, vol, proj, additive, stream, epi_mul);
            if (rc != 1) return rc;  // 1: not applicable, take the direct kernel
        }
        if (g.nz > 65535) return fail(TSP_ERR_INVALID, "voxel supersampling supports nz <= 65535");
        dim3 grid((g.nx + 31) / 32, (g.ny + 7) / 8, g.nz), block(32, 8);
        if (cone) bp_supersample_kernel<true><<<grid, block, 0, stream>>>(P);
        else bp_supersample_kernel<false><<<grid, block, 0, stream>>>(P);
    } else {
        int zpt = bp_zpt_choice(g.nx, g.ny, g.nz);
        TensorMapPair tmap;
        const bool use_tma = !getenv("TSP_BP_NO_TMA") &&
                             make_proj_tensor_map(proj, g.det_cols, g.n_angles, g.det_rows, BP_TMA_PITCH, bp_wv(zpt), &tmap.m[0]) &&
                             make_proj_tensor_map(proj, g.det_cols, g.n_angles, g.det_rows, BP_TMA_PITCH_B, bp_wv(zpt), &tmap.m[1]);
        if (!use_tma && zpt > 32) zpt = 32;  // the tall tile exists for the TMA kernel only
        const int ty = use_tma ? bp_tma_ty(zpt) : BP_TY;
        const int gz = (g.nz + zpt - 1) / zpt;
        const int gy = (g.ny + ty - 1) / ty;
        if (gz > 65535 || gy > 65535) return fail(TSP_ERR_INVALID, "volume too large for the BP grid");
        dim3 grid((g.nx + BP_TX - 1) / BP_TX, gy, gz), block(BP_TX, BP_TY);
        P.magic_off = 0u - 4u * BP_MAGIC_BITS * (uint32_t)((use_tma ? BP_TMA_PITCH : BP_PITCH) + 1);
        P.magic_off_b = 0u - 4u * BP_MAGIC_BITS * (uint32_t)(BP_TMA_PITCH_B + 1);
        if (use_tma) {
            // the descriptors travel as a __grid_constant__ parameter
            if (int rc = launch_bp_tma_variant(cone, zpt, grid, stream, P, tmap)) return rc;
        } else {
            if (int rc = launch_bp_variant(cone, zpt, grid, block, stream, P)) return rc;
        }
        used_tma = use_tma ? 1 : 0;
    }
    ++pr->launches;
    pr->bp_uses_tma = used_tma;
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return -1;
        if (prev != device) {
            if (cudaSetDevice(device) != cudaSuccess) return -1;
            active = true;
        }
        return 0;
    }
    ~DeviceGuard()
    {
        if (active) cudaSetDevice(prev);
    }
};


// ------------------------------------------------- host-array pipeline ----
// What ASTRA's CompositeGeometryManager does for host arrays (reference doc/topics/operator.rst:226-261;
// SURVEY.md 8f rank 4): cut the job into sub-problems so that transfers overlap the kernels.
//   BP: z-slabs of the volume; slab k needs only the detector rows its cone shadow covers, which are
//       contiguous in the (v, angle, u) layout -> H2D rows(k) | BP_k | D2H slab(k) run as a 3-stage pipeline;
//   FP: detector row blocks; block j needs only the slices its rays can reach (contiguous in z)
//       -> H2D slices(j) | FP_j | D2H rows(j).
// Every sub-problem is an ordinary projector on a sub-geometry (same vectors, shifted detector centre,
// cropped volume window) that marches along its parent's axes, so the results are those of the
// single-shot path up to fp32 summation order.
static tsp_projector *make_sub_projector(const tsp_projector *pr, int z0, int z1, int v0, int v1)
{
    tsp_geometry g = pr->g;
    std::vector<double> vec = pr->vectors;
    const double sz = pr->sigma[2];
    g.nz = z1 - z0;
    g.win_min[2] = pr->g.win_min[2] + z0 * sz;
    g.win_max[2] = pr->g.win_min[2] + z1 * sz;
    g.det_rows = v1 - v0;
    const double shift = 0.5 * (v0 + v1) - 0.5 * pr->g.det_rows;  // new detector centre, in rows from the old one
    for (int a = 0; a < g.n_angles; ++a)
        for (int i = 0; i < 3; ++i) vec[12 * (size_t)a + 3 + i] += shift * vec[12 * (size_t)a + 9 + i];
    g.vectors = vec.data();
    tsp_projector *sub = nullptr;
    if (create_projector_internal(&g, pr->march_axis.data(), &sub) != TSP_OK) return nullptr;
    sub->pool_owner = pr->pool_owner ? pr->pool_owner : const_cast<tsp_projector *>(pr);
    return sub;
}

// ---- supersampling through the staged kernels -----------------------------------------------------------------------
// VoxelSuperSampling d: the mean over d^3 sub-voxel centres (reference doc/topics/operator.rst:77-81) is the
// backprojection onto a d-times finer grid, summed over each coarse voxel.  DetectorSuperSampling d: the mean over
// d^2 sub-rays (:83-86) is the forward projection onto a d-times finer detector, averaged over each coarse pixel.
// Both therefore run the ordinary TMA kernels on a refined geometry, slab by slab / row block by row block through a
// bounded scratch buffer, followed by a pooling kernel (the direct kernels - one thread per voxel with fp64
// sub-voxel arithmetic, plain-load FP - remain as the fall-back and as the cross-check, TSP_SS_DIRECT=1).
static tsp_projector *make_fine_projector(const tsp_projector *pr, int vox_d, int det_d)
{
    tsp_geometry g = pr->g;
    std::vector<double> vec = pr->vectors;
    g.nx *= vox_d; g.ny *= vox_d; g.nz *= vox_d;
    g.det_rows *= det_d; g.det_cols *= det_d;
    for (int a = 0; a < g.n_angles; ++a)
        for (int i = 0; i < 6; ++i) vec[12 * (size_t)a + 6 + i] /= det_d;  // pixel vectors; the detector centre stays
    g.voxel_supersampling = g.detector_supersampling = 1;
    g.vectors = vec.data();
    tsp_projector *fine = nullptr;
    if (create_projector_internal(&g, pr->march_axis.data(), &fine) != TSP_OK) return nullptr;
    // the slabs / row blocks cut from it belong to `pr` (the fine projector itself is only a template)
    fine->pool_owner = pr->pool_owner ? pr->pool_owner : const_cast<tsp_projector *>(pr);
    return fine;
}

static bool plan_supersampling(tsp_projector *pr)
{
    std::lock_guard<std::mutex> lock(pr->mu);
    if (pr->ss_planned) return !pr->ss_bp.empty() || !pr->ss_fp.empty();
    pr->ss_planned = true;
    const tsp_geometry &g = pr->g;
    const size_t budget = (size_t)256 << 20;  // scratch bytes per sub-problem
    if (g.voxel_supersampling > 1) {
        const int d = g.voxel_supersampling;
        tsp_projector *fine = make_fine_projector(pr, d, 1);
        if (fine) {
            const size_t fine_slice = (size_t)g.nx * d * g.ny * d * d * sizeof(float);  // bytes of one COARSE slice
            int s = (int)std::max<size_t>(1, std::min<size_t>(32, budget / std::max<size_t>(1, fine_slice)));
            for (int z0 = 0; z0 < g.nz; z0 += s) {
                tsp_projector::HostChunk c;
                c.z0 = z0; c.z1 = std::min(g.nz, z0 + s); c.v0 = 0; c.v1 = g.det_rows;
                c.sub = make_sub_projector(fine, c.z0 * d, c.z1 * d, 0, g.det_rows);
                if (!c.sub) break;
                pr->ss_bp.push_back(c);
            }
            tsp_projector_destroy(fine);
            if (pr->ss_bp.empty() || pr->ss_bp.back().z1 != g.nz) {
                for (auto &c : pr->ss_bp) tsp_projector_destroy(c.sub);
                pr->ss_bp.clear();
            }
        }
    }
    if (g.detector_supersampling > 1) {
        const int d = g.detector_supersampling;
        tsp_projector *fine = make_fine_projector(pr, 1, d);
        if (fine) {
            const size_t fine_row = (size_t)g.n_angles * g.det_cols * d * d * sizeof(float);  // bytes of one COARSE row
            int r = (int)std::max<size_t>(1, std::min<size_t>(32, budget / std::max<size_t>(1, fine_row)));
            for (int v0 = 0; v0 < g.det_rows; v0 += r) {
                tsp_projector::HostChunk c;
                c.v0 = v0; c.v1 = std::min(g.det_rows, v0 + r); c.z0 = 0; c.z1 = g.nz;
                c.sub = make_sub_projector(fine, 0, g.nz, c.v0 * d, c.v1 * d);
                if (!c.sub) break;
                pr->ss_fp.push_back(c);
            }
            tsp_projector_destroy(fine);
            if (pr->ss_fp.empty() || pr->ss_fp.back().v1 != g.det_rows) {
                for (auto &c : pr->ss_fp) tsp_projector_destroy(c.sub);
                pr->ss_fp.clear();
            }
        }
    }
    return !pr->ss_bp.empty() || !pr->ss_fp.empty();
}

static int launch_bp_supersampled(tsp_projector *pr, DeviceState *st, float *vol, const float *proj, int additive, cudaStream_t stream,
                                  const float *epi_mul)
{
    plan_supersampling(pr);
    if (pr->ss_bp.empty()) return 1;
    const tsp_geometry &g = pr->g;
    const int d = g.voxel_supersampling;
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    size_t max_elems = 0;
    for (const auto &c : pr->ss_bp) max_elems = std::max(max_elems, (size_t)(c.z1 - c.z0) * d * g.ny * d * g.nx * d);
    PoolScratch scratch;
    scratch.stream = stream;
    pool_keep_at_least(st, st->pool_keep_base + max_elems * sizeof(float));
    CUDA_TRY(pool_alloc(st, &scratch.p, max_elems * sizeof(float), stream));
    BPArgs P;
    memset(&P, 0, sizeof P);
    P.vol = vol; P.nx = g.nx; P.ny = g.ny; P.nz = g.nz;
    P.additive = additive; P.epi_mul = epi_mul;
    for (const auto &c : pr->ss_bp) {
        DeviceState *sst = nullptr;
        if (int rc = get_device_state(c.sub, device, &sst)) return rc;
        const int64_t l0 = c.sub->launches;
        if (int rc = launch_bp(c.sub, sst, (float *)scratch.p, proj, 0, stream)) return rc;  // the same scratch: stream order
        pr->launches += c.sub->launches - l0 + 1;
        pr->bp_uses_tma = c.sub->bp_uses_tma.load();
        dim3 grid((g.nx + 31) / 32, (g.ny + 7) / 8, c.z1 - c.z0), block(32, 8);
        bp_pool_kernel<<<grid, block, 0, stream>>>(P, (const float *)scratch.p, d, c.z0, c.z1 - c.z0);
    }
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

static int launch_fp_supersampled(tsp_projector *pr, DeviceState *st, const float *vol, float *proj, int additive, cudaStream_t stream,
                                  const float *epi_sub, const float *epi_mul)
{
    plan_supersampling(pr);
    if (pr->ss_fp.empty()) return 1;
    const tsp_geometry &g = pr->g;
    const int d = g.detector_supersampling;
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    size_t max_elems = 0;
    for (const auto &c : pr->ss_fp) max_elems = std::max(max_elems, (size_t)(c.v1 - c.v0) * d * g.n_angles * g.det_cols * d);
    PoolScratch scratch;
    scratch.stream = stream;
    pool_keep_at_least(st, st->pool_keep_base + max_elems * sizeof(float));
    CUDA_TRY(pool_alloc(st, &scratch.p, max_elems * sizeof(float), stream));
    FPArgs P;
    memset(&P, 0, sizeof P);
    P.proj = proj; P.det_u = g.det_cols; P.det_v = g.det_rows; P.n_angles = g.n_angles;
    P.additive = additive; P.epi_sub = epi_sub; P.epi_mul = epi_mul;
    for (const auto &c : pr->ss_fp) {
        DeviceState *sst = nullptr;
        if (int rc = get_device_state(c.sub, device, &sst)) return rc;
        const int64_t l0 = c.sub->launches;
        if (int rc = launch_fp(c.sub, sst, vol, (float *)scratch.p, 0, stream)) return rc;
        pr->launches += c.sub->launches - l0 + 1;
        pr->fp_uses_tma = c.sub->fp_uses_tma.load(); pr->fp_uses_transpose = c.sub->fp_uses_transpose.load();
        dim3 grid((g.det_cols + 31) / 32, (g.n_angles + 7) / 8, c.v1 - c.v0), block(32, 8);
        if ((g.n_angles + 7) / 8 > 65535 || c.v1 - c.v0 > 65535) return fail(TSP_ERR_INVALID, "too many angles for the pooling grid");
        fp_pool_kernel<<<grid, block, 0, stream>>>(P, (const float *)scratch.p, d, c.v0, c.v1 - c.v0);
    }
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

// detector rows [v0, v1) that voxels of the slab z0 <= z < z1 can touch (bilinear taps included)
static void slab_row_range(const tsp_projector *pr, int z0, int z1, int &v0, int &v1)
{
    const tsp_geometry &g = pr->g;
    double lo = 1e300, hi = -1e300;
    bool all = false;
    for (int a = 0; a < g.n_angles && !all; ++a) {
        const BPAngle &m = pr->bp_angles[a];
        int sign = 0;
        for (int c = 0; c < 8; ++c) {
            const double x = (c & 1) ? 0.5 * g.nx : -0.5 * g.nx, y = (c & 2) ? 0.5 * g.ny : -0.5 * g.ny;
            const double z = ((c & 4) ? z1 : z0) - 0.5 * g.nz;
            const double den = m.dn[0] * x + m.dn[1] * y + m.dn[2] * z + m.dn[3];
            const int sg = den > 0 ? 1 : -1;
            if (den == 0.0 || (sign && sg != sign)) { all = true; break; }
            sign = sg;
            const double v = (m.nv[0] * x + m.nv[1] * y + m.nv[2] * z + m.nv[3]) / den;
            if (!std::isfinite(v)) { all = true; break; }
            lo = std::min(lo, v); hi = std::max(hi, v);
        }
    }
    if (all) { v0 = 0; v1 = g.det_rows; return; }
    v0 = (int)std::max(0.0, std::floor(lo) - 2.0);
    v1 = (int)std::min((double)g.det_rows, std::ceil(hi) + 2.0);
    if (v1 <= v0) { v0 = std::min(std::max(v0, 0), g.det_rows - 1); v1 = v0 + 1; }  // shadow off the detector
}


// 2-D distance from point q to segment [a, b]
static double dist_point_segment(const double q[2], const double a[2], const double b[2])
{
    const double ab[2] = {b[0] - a[0], b[1] - a[1]}, aq[2] = {q[0] - a[0], q[1] - a[1]};
    const double l2 = ab[0] * ab[0] + ab[1] * ab[1];
    double t = l2 > 0 ? (aq[0] * ab[0] + aq[1] * ab[1]) / l2 : 0.0;
    t = std::min(1.0, std::max(0.0, t));
    const double dx = aq[0] - t * ab[0], dy = aq[1] - t * ab[1];
    return std::sqrt(dx * dx + dy * dy);
}

// volume slices [z0, z1) that rays of the detector rows v0 <= v < v1 can touch: a guaranteed bound
// from the steepest / shallowest ray elevation of the block and the horizontal reach of the volume
static void block_z_range(const tsp_projector *pr, int v0, int v1, int &z0, int &z1)
{
    const tsp_geometry &g = pr->g;
    const double rxy = 0.5 * std::sqrt((double)g.nx * g.nx + (double)g.ny * g.ny) + 1.0;
    double lo = 1e300, hi = -1e300;
    bool all = false;
    for (int a = 0; a < g.n_angles && !all; ++a) {
        NormAngle n;
        normalise_angle(pr, a, n);
        double pix[4][3];
        for (int c = 0; c < 4; ++c) {
            const double cu = ((c & 1) ? g.det_cols : 0) - 0.5 * g.det_cols, cv = ((c & 2) ? v1 : v0) - 0.5 * g.det_rows;
            for (int i = 0; i < 3; ++i) pix[c][i] = n.dc[i] + cu * n.u[i] + cv * n.v[i];
        }
        if (g.kind == TSP_KIND_CONE_VEC) {
            const double *s = n.p;
            double dzmin = 1e300, dzmax = -1e300, hmax = 0.0;
            for (int c = 0; c < 4; ++c) {
                dzmin = std::min(dzmin, pix[c][2] - s[2]); dzmax = std::max(dzmax, pix[c][2] - s[2]);
                hmax = std::max(hmax, std::hypot(pix[c][0] - s[0], pix[c][1] - s[1]));
            }
            // min horizontal distance from the source to the block's parallelogram (corner order 0,1,3,2)
            const int ord[4] = {0, 1, 3, 2};
            double hmin = 1e300;
            bool inside = true;
            double sgn = 0.0;
            for (int e = 0; e < 4; ++e) {
                const double *p0 = pix[ord[e]], *p1 = pix[ord[(e + 1) & 3]];
                hmin = std::min(hmin, dist_point_segment(s, p0, p1));
                const double cr = (p1[0] - p0[0]) * (s[1] - p0[1]) - (p1[1] - p0[1]) * (s[0] - p0[0]);
                if (cr != 0.0) { if (sgn == 0.0) sgn = cr; else if ((cr > 0) != (sgn > 0)) inside = false; }
            }
            const double ds = std::hypot(s[0], s[1]);
            if (inside || hmin < 1e-9 * (hmax + 1.0)) { all = true; break; }
            const double dnear = std::max(0.0, ds - rxy), dfar = ds + rxy;
            const double slo = std::min(dzmin / hmin, dzmin / hmax), shi = std::max(dzmax / hmin, dzmax / hmax);
            const double cand[4] = {slo * dnear, slo * dfar, shi * dnear, shi * dfar};
            for (double c : cand) { lo = std::min(lo, s[2] + c); hi = std::max(hi, s[2] + c); }
        } else {
            const double *r = n.p;
            const double h = std::hypot(r[0], r[1]);
            if (h < 1e-9 * std::fabs(r[2])) { all = true; break; }
            for (int c = 0; c < 4; ++c) {
                const double reach = (std::hypot(pix[c][0], pix[c][1]) + rxy) / h * std::fabs(r[2]);
                lo = std::min(lo, pix[c][2] - reach); hi = std::max(hi, pix[c][2] + reach);
            }
        }
        if (!std::isfinite(lo) || !std::isfinite(hi)) all = true;
    }
    if (all) { z0 = 0; z1 = g.nz; return; }
    z0 = (int)std::max(0.0, std::floor(lo + 0.5 * g.nz) - 2.0);
    z1 = (int)std::min((double)g.nz, std::ceil(hi + 0.5 * g.nz) + 2.0);
    if (z1 <= z0) { z0 = std::min(std::max(z0, 0), g.nz - 1); z1 = z0 + 1; }  // block sees nothing of the volume
}

static int chunk_size(int n)
{
    int k = 8;  // chunks per job; TSP_HOST_CHUNKS overrides (tuning aid)
    if (const char *e = getenv("TSP_HOST_CHUNKS")) k = std::max(1, atoi(e));
    int c = (n + k - 1) / k;
    c = (c + 31) / 32 * 32;
    return std::max(32, c);
}

static bool plan_host_pipeline(tsp_projector *pr)
{
    std::lock_guard<std::mutex> lock(pr->mu);
    if (pr->host_planned) return !pr->host_bp.empty();
    pr->host_planned = true;
    const tsp_geometry &g = pr->g;
    if (g.nz < 64 || g.det_rows < 64) return false;
    const int cz = chunk_size(g.nz), cv = chunk_size(g.det_rows);
    std::vector<tsp_projector::HostChunk> bp, fp;
    bool ok = true;
    // uniform chunks, except where a transfer cannot hide behind a kernel: the slab the BP starts with (its
    // rows must arrive before any kernel runs) and the row block the FP ends with (its download follows the
    // last kernel) are halved
    std::vector<std::pair<int, int>> zr, vr;
    // (Graded sizes - 32 / 64 at the two exposed ends of the pipeline, chunks of up to n / 4 in between - were measured
    // at cfg 3: 1 801 GUPS against 1 892 for these uniform chunks, r02 GPU call 23: the coarse middle chunks pipeline
    // worse than their fewer launches save.)
    for (int z0 = 0; z0 < g.nz; z0 += cz) zr.push_back({z0, std::min(g.nz, z0 + cz)});
    for (int v0 = 0; v0 < g.det_rows; v0 += cv) vr.push_back({v0, std::min(g.det_rows, v0 + cv)});
    if (!getenv("TSP_HOST_UNIFORM")) {
        size_t mid = 0;
        for (size_t i = 0; i < zr.size(); ++i)
            if (zr[i].first <= (g.nz - 1) / 2 && (g.nz - 1) / 2 < zr[i].second) mid = i;
        if (zr.size() > 1 && zr[mid].second - zr[mid].first >= 64) {
            const int h = (zr[mid].first + zr[mid].second) / 2;
            const std::pair<int, int> hi{h, zr[mid].second};
            zr[mid].second = h;
            zr.insert(zr.begin() + mid + 1, hi);
        }
        if (vr.size() > 1 && vr.back().second - vr.back().first >= 64) {
            const int h = (vr.back().first + vr.back().second) / 2;
            const std::pair<int, int> hi{h, vr.back().second};
            vr.back().second = h;
            vr.push_back(hi);
        }
    }
    for (size_t i = 0; i < zr.size() && ok; ++i) {
        tsp_projector::HostChunk c;
        c.z0 = zr[i].first; c.z1 = zr[i].second;
        slab_row_range(pr, c.z0, c.z1, c.v0, c.v1);
        c.sub = make_sub_projector(pr, c.z0, c.z1, c.v0, c.v1);
        ok = c.sub != nullptr;
        bp.push_back(c);
    }
    for (size_t i = 0; i < vr.size() && ok; ++i) {
        tsp_projector::HostChunk c;
        c.v0 = vr[i].first; c.v1 = vr[i].second;
        block_z_range(pr, c.v0, c.v1, c.z0, c.z1);
        c.sub = make_sub_projector(pr, c.z0, c.z1, c.v0, c.v1);
        ok = c.sub != nullptr;
        fp.push_back(c);
    }
    if (!ok) {
        for (auto &c : bp) tsp_projector_destroy(c.sub);
        for (auto &c : fp) tsp_projector_destroy(c.sub);
        return false;
    }
    // BP: start with the slab whose shadow is the smallest (fewest rows to upload before the first kernel can
    // run: for a cone beam the central one) and work outwards, so that each later slab only adds a few rows
    if (!getenv("TSP_HOST_BP_INORDER") && !bp.empty()) {
        size_t first = 0;
        for (size_t i = 1; i < bp.size(); ++i)
            if (bp[i].v1 - bp[i].v0 < bp[first].v1 - bp[first].v0) first = i;
        std::vector<tsp_projector::HostChunk> ord;
        ord.push_back(bp[first]);
        for (size_t d = 1; d < bp.size(); ++d) {
            if (first + d < bp.size()) ord.push_back(bp[first + d]);
            if (first >= d) ord.push_back(bp[first - d]);
        }
        bp.swap(ord);
    }
    if (getenv("TSP_DEBUG")) {
        for (auto &c : bp) fprintf(stderr, "[tsp] host BP chunk: z [%d, %d) <- rows [%d, %d)\n", c.z0, c.z1, c.v0, c.v1);
        for (auto &c : fp) fprintf(stderr, "[tsp] host FP chunk: rows [%d, %d) <- z [%d, %d)\n", c.v0, c.v1, c.z0, c.z1);
    }
    pr->host_bp = std::move(bp);
    pr->host_fp = std::move(fp);
    return true;
}

extern "C" int tsp_projector_host_plan(tsp_projector *pr, int direction, int32_t *out, int max_chunks)
{
    if (!pr || (!out && max_chunks > 0)) return fail(TSP_ERR_INVALID, "NULL argument");
    if (direction != TSP_FP && direction != TSP_BP) return fail(TSP_ERR_INVALID, "direction must be TSP_FP or TSP_BP");
    if (!plan_host_pipeline(pr)) return 0;
    const std::vector<tsp_projector::HostChunk> &chunks = direction == TSP_FP ? pr->host_fp : pr->host_bp;
    for (size_t k = 0; k < chunks.size() && (int)k < max_chunks; ++k) {
        out[4 * k + 0] = chunks[k].z0; out[4 * k + 1] = chunks[k].z1;
        out[4 * k + 2] = chunks[k].v0; out[4 * k + 3] = chunks[k].v1;
    }
    return (int)chunks.size();
}

// Everything a host-array call acquires on one device; released on every exit path.
struct PipeResources {
    cudaStream_t stream = nullptr;  // the stream the buffers are allocated / freed on
    std::vector<void *> bufs;
    std::vector<cudaEvent_t> events;
    cudaStream_t own_stream = nullptr;
    ~PipeResources()
    {
        // (nothing to wait for when nothing was acquired: device-array calls stay asynchronous and capturable)
        for (void *b : bufs) cudaFreeAsync(b, stream);
        if (!bufs.empty() || !events.empty()) cudaStreamSynchronize(stream);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
    cudaError_t alloc(DeviceState *st, float **p, size_t n_floats)
    {
        void *q = nullptr;
        cudaError_t e = pool_alloc(st, &q, std::max<size_t>(n_floats, 1) * sizeof(float), stream);
        if (e == cudaSuccess) { bufs.push_back(q); *p = (float *)q; }
        return e;
    }
    cudaError_t event(cudaEvent_t *ev)
    {
        cudaError_t e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e == cudaSuccess) events.push_back(*ev);
        return e;
    }
};

// Device-memory budget of a host-array call, in bytes (0 = unlimited): TSP_HOST_MEM_CAP_MB, else what the device
// has free minus a reserve.  Above it the pipeline runs out of a ring of chunk buffers instead of whole arrays.
static size_t host_mem_cap()
{
    if (const char *e = getenv("TSP_HOST_MEM_CAP_MB")) return (size_t)std::max(1LL, atoll(e)) << 20;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return 0; }
    return fr - std::min(fr / 8, (size_t)4 << 30);
}

// One FP (SET) or BP (SET) between host arrays on ONE device, over the chunks `order` (indices into the plan):
// H2D(k+1) | kernels(k) | D2H(k-1) on three streams.
//   input axis  = detector rows for BP, volume slices for FP   (what a chunk reads;  unit = one row / one slice)
//   output axis = volume slices for BP, detector rows for FP   (what a chunk writes, disjoint between chunks)
// Two memory modes:
//   range mode (default): one input and one output buffer spanning the union of the chunks' ranges; an input
//       row / slice is uploaded once, by the first chunk that needs it;
//   ring mode (when range mode would exceed `cap` bytes): RING input and RING output buffers of the largest
//       chunk's size; every chunk uploads its whole input range (neighbouring chunks' inputs overlap, so up to
//       ~2x the upload traffic), and a buffer is reused once the chunk that held it has finished with it.
//       Device memory is then bounded by 3 x (largest chunk input + output) whatever the size of the arrays -
//       the out-of-core path of ASTRA's CompositeGeometryManager (reference doc/topics/operator.rst:226-261).
static int project_host_chunks(tsp_projector *pr, int device, int direction, float *vol, float *proj,
                               const std::vector<int> &order, cudaStream_t user_stream, size_t cap, bool *used_ring)
{
    const tsp_geometry &g = pr->g;
    const size_t row = (size_t)g.n_angles * g.det_cols, slice = (size_t)g.nx * g.ny;
    const bool fp = direction == TSP_FP;
    const std::vector<tsp_projector::HostChunk> &chunks = fp ? pr->host_fp : pr->host_bp;
    const size_t in_unit = fp ? slice : row, out_unit = fp ? row : slice;
    float *const host_in = fp ? vol : proj, *const host_out = fp ? proj : vol;
    auto in0 = [&](const tsp_projector::HostChunk &c) { return fp ? c.z0 : c.v0; };
    auto in1 = [&](const tsp_projector::HostChunk &c) { return fp ? c.z1 : c.v1; };
    auto out0 = [&](const tsp_projector::HostChunk &c) { return fp ? c.v0 : c.z0; };
    auto out1 = [&](const tsp_projector::HostChunk &c) { return fp ? c.v1 : c.z1; };
    if (order.empty()) return TSP_OK;

    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    if (!st->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking));
    if (!st->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking));

    int lo_in = INT_MAX, hi_in = 0, lo_out = INT_MAX, hi_out = 0, max_in = 0, max_out = 0;
    for (int k : order) {
        const auto &c = chunks[k];
        lo_in = std::min(lo_in, in0(c)); hi_in = std::max(hi_in, in1(c));
        lo_out = std::min(lo_out, out0(c)); hi_out = std::max(hi_out, out1(c));
        max_in = std::max(max_in, in1(c) - in0(c)); max_out = std::max(max_out, out1(c) - out0(c));
    }
    const size_t range_bytes = ((size_t)(hi_in - lo_in) * in_unit + (size_t)(hi_out - lo_out) * out_unit) * sizeof(float);
    const bool ring = cap != 0 && range_bytes > cap;
    constexpr int RING = 3;
    if (used_ring) *used_ring = ring;
    if (ring && (size_t)RING * ((size_t)max_in * in_unit + (size_t)max_out * out_unit) * sizeof(float) > cap)
        return fail(TSP_ERR_NOMEM, "host-array pipeline: %d ring buffers of the largest chunk (%zu MB) exceed the device memory budget of %zu MB; "
                    "raise TSP_HOST_CHUNKS", RING, ((size_t)max_in * in_unit + (size_t)max_out * out_unit) * sizeof(float) >> 20, cap >> 20);

    PipeResources res;
    res.stream = user_stream;
    float *din[RING] = {nullptr, nullptr, nullptr}, *dout[RING] = {nullptr, nullptr, nullptr};
    const int nbuf = ring ? RING : 1;
    pool_keep_at_least(st, st->pool_keep_base + (ring ? (size_t)RING * ((size_t)max_in * in_unit + (size_t)max_out * out_unit) * sizeof(float)
                                                      : range_bytes));
    for (int i = 0; i < nbuf; ++i) {
        CUDA_TRY(res.alloc(st, &din[i], (size_t)(ring ? max_in : hi_in - lo_in) * in_unit));
        CUDA_TRY(res.alloc(st, &dout[i], (size_t)(ring ? max_out : hi_out - lo_out) * out_unit));
    }
    const size_t n = order.size();
    std::vector<cudaEvent_t> ev_in(n), ev_done(n), ev_out(n);
    cudaEvent_t ev_start;
    CUDA_TRY(res.event(&ev_start));
    for (size_t i = 0; i < n; ++i) {
        CUDA_TRY(res.event(&ev_in[i]));
        CUDA_TRY(res.event(&ev_done[i]));
        CUDA_TRY(res.event(&ev_out[i]));
    }
    cudaStream_t stream = user_stream;
    CUDA_TRY(cudaEventRecord(ev_start, stream));  // buffers allocated, earlier work on `stream` ordered before us
    CUDA_TRY(cudaStreamWaitEvent(st->s_in, ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(st->s_out, ev_start, 0));

    std::vector<char> uploaded(ring ? 0 : (size_t)(hi_in - lo_in), 0);
    for (size_t i = 0; i < n; ++i) {
        const auto &c = chunks[order[i]];
        const int slot = ring ? (int)(i % RING) : 0;
        // base pointers such that element (unit u of the full array) lives at base + u * unit
        float *in_base = ring ? din[slot] - (size_t)in0(c) * in_unit : din[0] - (size_t)lo_in * in_unit;
        float *out_base = ring ? dout[slot] - (size_t)out0(c) * out_unit : dout[0] - (size_t)lo_out * out_unit;
        if (ring) {
            // the slot's previous tenant (chunk i - RING) must be done: its kernels with the input buffer ...
            if (i >= RING) CUDA_TRY(cudaStreamWaitEvent(st->s_in, ev_done[i - RING], 0));
            CUDA_TRY(cudaMemcpyAsync(in_base + (size_t)in0(c) * in_unit, host_in + (size_t)in0(c) * in_unit,
                                     (size_t)(in1(c) - in0(c)) * in_unit * sizeof(float), cudaMemcpyHostToDevice, st->s_in));
        } else {
            // upload what no earlier chunk brought in (maximal runs)
            for (int u = in0(c); u < in1(c);) {
                if (uploaded[u - lo_in]) { ++u; continue; }
                int e = u;
                while (e < in1(c) && !uploaded[e - lo_in]) uploaded[e++ - lo_in] = 1;
                CUDA_TRY(cudaMemcpyAsync(in_base + (size_t)u * in_unit, host_in + (size_t)u * in_unit,
                                         (size_t)(e - u) * in_unit * sizeof(float), cudaMemcpyHostToDevice, st->s_in));
                u = e;
            }
        }
        CUDA_TRY(cudaEventRecord(ev_in[i], st->s_in));
        CUDA_TRY(cudaStreamWaitEvent(stream, ev_in[i], 0));
        // ... and its download out of the output buffer
        if (ring && i >= RING) CUDA_TRY(cudaStreamWaitEvent(stream, ev_out[i - RING], 0));
        DeviceState *sst = nullptr;
        if (int rc = get_device_state(c.sub, device, &sst)) return rc;
        const int64_t l0 = c.sub->launches;
        float *dvol_c = fp ? in_base + (size_t)c.z0 * slice : out_base + (size_t)c.z0 * slice;
        float *dproj_c = fp ? out_base + (size_t)c.v0 * row : in_base + (size_t)c.v0 * row;
        int rc = fp ? launch_fp(c.sub, sst, dvol_c, dproj_c, 0, stream) : launch_bp(c.sub, sst, dvol_c, dproj_c, 0, stream);
        pr->launches += c.sub->launches - l0;
        if (fp) { pr->fp_uses_tma = c.sub->fp_uses_tma.load(); pr->fp_uses_transpose = c.sub->fp_uses_transpose.load(); }
        else pr->bp_uses_tma = c.sub->bp_uses_tma.load();
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(ev_done[i], stream));
        CUDA_TRY(cudaStreamWaitEvent(st->s_out, ev_done[i], 0));
        CUDA_TRY(cudaMemcpyAsync(host_out + (size_t)out0(c) * out_unit, out_base + (size_t)out0(c) * out_unit,
                                 (size_t)(out1(c) - out0(c)) * out_unit * sizeof(float), cudaMemcpyDeviceToHost, st->s_out));
        CUDA_TRY(cudaEventRecord(ev_out[i], st->s_out));
    }
    // the user's stream completes after the copy-out stream; the destructor of `res` then releases everything
    CUDA_TRY(cudaStreamWaitEvent(stream, ev_out[n - 1], 0));
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaStreamSynchronize(st->s_in));
    return TSP_OK;
}

static int project_host_pipelined(tsp_projector *pr, int device, int direction, float *vol, float *proj, cudaStream_t stream)
{
    const std::vector<tsp_projector::HostChunk> &chunks = direction == TSP_FP ? pr->host_fp : pr->host_bp;
    std::vector<int> order(chunks.size());
    for (size_t k = 0; k < chunks.size(); ++k) order[k] = (int)k;
    bool ring = false;
    const int rc = project_host_chunks(pr, device, direction, vol, proj, order, stream, host_mem_cap(), &ring);
    pr->host_ring = ring ? 1 : 0;
    return rc;
}

// Host arrays over several devices - what `astra.set_gpu_index([0, 1, ...])` switches on in the reference
// (doc/topics/operator.rst:233-245): the chunks of the plan are independent sub-problems with disjoint outputs, so
// each device takes a contiguous run of them (contiguous in the output axis: neighbouring chunks share input rows /
// slices) and runs the single-device pipeline over its share from its own host thread.
extern "C" int tsp_project_multi(tsp_projector *pr, int direction, int additive, void *vol, void *proj, const int *devices,
                                 int n_devices)
{
    if (!pr) return fail(TSP_ERR_INVALID, "projector is NULL");
    if (!vol || !proj) return fail(TSP_ERR_INVALID, "vol / proj pointer is NULL");
    if (direction != TSP_FP && direction != TSP_BP) return fail(TSP_ERR_INVALID, "direction must be TSP_FP or TSP_BP");
    if (!devices || n_devices < 1) return fail(TSP_ERR_INVALID, "need at least one device");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    for (int i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", devices[i], ndev);
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(TSP_ERR_INVALID, "device %d listed twice", devices[i]);
    }
    // additive calls, small problems and single-device lists take the ordinary path on the first device
    if (n_devices == 1 || additive || !plan_host_pipeline(pr))
        return tsp_project(pr, direction, additive, vol, proj, 1, TSP_MEM_HOST, devices[0], nullptr);
    const std::vector<tsp_projector::HostChunk> &chunks = direction == TSP_FP ? pr->host_fp : pr->host_bp;
    // chunks sorted along the output axis, cut into n_devices contiguous shares of (nearly) equal output size
    std::vector<int> sorted(chunks.size());
    for (size_t k = 0; k < chunks.size(); ++k) sorted[k] = (int)k;
    auto key = [&](int k) { return direction == TSP_FP ? chunks[k].v0 : chunks[k].z0; };
    auto len = [&](int k) { return direction == TSP_FP ? chunks[k].v1 - chunks[k].v0 : chunks[k].z1 - chunks[k].z0; };
    std::sort(sorted.begin(), sorted.end(), [&](int a, int b) { return key(a) < key(b); });
    long long total = 0;
    for (int k : sorted) total += len(k);
    std::vector<std::vector<int>> share(n_devices);
    long long acc = 0;
    for (int k : sorted) {
        const int d = (int)std::min<long long>(n_devices - 1, (acc + len(k) / 2) * n_devices / std::max(1LL, total));
        share[d].push_back(k);
        acc += len(k);
    }
    std::vector<int> rcs(n_devices, TSP_OK);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> threads;
    for (int d = 0; d < n_devices; ++d) {
        threads.emplace_back([&, d]() {
            if (share[d].empty()) return;
            DeviceGuard guard;
            if (guard.enter(devices[d]) != 0) { rcs[d] = TSP_ERR_CUDA; errs[d] = "cannot switch device"; return; }
            cudaStream_t s = nullptr;
            if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { rcs[d] = TSP_ERR_CUDA; errs[d] = "cannot create a stream"; return; }
            rcs[d] = project_host_chunks(pr, devices[d], direction, (float *)vol, (float *)proj, share[d], s, host_mem_cap(), nullptr);
            if (rcs[d]) errs[d] = g_last_error;  // thread-local message of the worker
            cudaStreamDestroy(s);
        });
    }
    for (auto &t : threads) t.join();
    pr->host_pipelined = 1;
    pr->host_devices = n_devices;
    for (int d = 0; d < n_devices; ++d)
        if (rcs[d]) return fail(rcs[d], "device %d: %s", devices[d], errs[d].c_str());
    return TSP_OK;
}

extern "C" int tsp_project(tsp_projector *pr, int direction, int additive, void *vol, void *proj, int batch,
                           int memory_kind, int device, void *cuda_stream)
{
    if (!pr) return fail(TSP_ERR_INVALID, "projector is NULL");
    if (!vol || !proj) return fail(TSP_ERR_INVALID, "vol / proj pointer is NULL");
    if (direction != TSP_FP && direction != TSP_BP) return fail(TSP_ERR_INVALID, "direction must be TSP_FP or TSP_BP");
    if (batch < 1) return fail(TSP_ERR_INVALID, "batch must be >= 1");
    if (memory_kind != TSP_MEM_HOST && memory_kind != TSP_MEM_DEVICE)
        return fail(TSP_ERR_INVALID, "memory_kind must be TSP_MEM_HOST or TSP_MEM_DEVICE");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);

    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    const tsp_geometry &g = pr->g;
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    const size_t npix = (size_t)g.det_rows * g.n_angles * g.det_cols;

    float *dvol = (float *)vol, *dproj = (float *)proj;
    pr->host_pipelined = 0;
    size_t pipeline_min = 64u << 20;  // smaller jobs: one upload, one launch, one download
    if (const char *e = getenv("TSP_HOST_PIPELINE_MIN_MB")) pipeline_min = (size_t)atoll(e) << 20;
    if (memory_kind == TSP_MEM_HOST && batch == 1 && !additive && (nvox + npix) * sizeof(float) >= pipeline_min &&
        !getenv("TSP_HOST_NO_PIPELINE") && plan_host_pipeline(pr)) {
        pr->host_pipelined = 1;
        pr->host_devices = 1;
        return project_host_pipelined(pr, device, direction, (float *)vol, (float *)proj, stream);
    }
    PipeResources res;  // releases the staging buffers on every exit path
    res.stream = stream;
    if (memory_kind == TSP_MEM_HOST) {
        pool_keep_at_least(st, st->pool_keep_base + (nvox + npix) * batch * sizeof(float));
        CUDA_TRY(res.alloc(st, &dvol, nvox * batch));
        CUDA_TRY(res.alloc(st, &dproj, npix * batch));
        // inputs, and the destination too when accumulating
        if (direction == TSP_FP || additive)
            CUDA_TRY(cudaMemcpyAsync(dvol, vol, nvox * batch * sizeof(float), cudaMemcpyHostToDevice, stream));
        if (direction == TSP_BP || additive)
            CUDA_TRY(cudaMemcpyAsync(dproj, proj, npix * batch * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    int rc = TSP_OK;
    if (direction == TSP_FP) rc = launch_fp(pr, st, dvol, dproj, additive, stream, nullptr, nullptr, batch);
    else rc = launch_bp(pr, st, dvol, dproj, additive, stream, nullptr, batch);
    if (memory_kind == TSP_MEM_HOST && rc == TSP_OK) {
        if (direction == TSP_FP)
            CUDA_TRY(cudaMemcpyAsync(proj, dproj, npix * batch * sizeof(float), cudaMemcpyDeviceToHost, stream));
        else
            CUDA_TRY(cudaMemcpyAsync(vol, dvol, nvox * batch * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return rc;
}

// ------------------------------------------------------------------- SIRT --
extern "C" int tsp_sirt(tsp_projector *pr, void *x, const void *y, const void *R, const void *C, void *y_tmp,
                        int iterations, int device, void *cuda_stream)
{
    if (!pr || !x || !y || !R || !C || !y_tmp) return fail(TSP_ERR_INVALID, "NULL argument");
    if (iterations < 0) return fail(TSP_ERR_INVALID, "iterations must be >= 0");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const tsp_geometry &g = pr->g;
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;
    const size_t npix = (size_t)g.det_rows * g.n_angles * g.det_cols;
    // Each iteration is two launches of the projection kernels with fused epilogues:
    //   FP stores   y_tmp = R * (A x - y)        (fp_store)
    //   BP stores   x    -= C * (A^T y_tmp)      (bp_store_one)
    // so the residual / update passes of the reference loop (README.md:162-163,
    // notebooks/sirt_benchmark.py:130-136) cost no extra HBM traffic.
    int rc = TSP_OK;
    for (int it = 0; it < iterations && rc == TSP_OK; ++it) {
        rc = launch_fp(pr, st, (const float *)x, (float *)y_tmp, 0, stream, (const float *)y, (const float *)R);
        if (rc) break;
        rc = launch_bp(pr, st, (float *)x, (const float *)y_tmp, 0, stream, (const float *)C);
    }
    (void)nvox; (void)npix;
    if (rc == TSP_OK) CUDA_TRY(cudaGetLastError());
    return rc;
}

extern "C" int tsp_project_fused(tsp_projector *pr, int direction, void *vol, void *proj, const void *sub, const void *mul,
                                 int device, void *cuda_stream)
{
    if (!pr || !vol || !proj || !mul) return fail(TSP_ERR_INVALID, "NULL argument");
    if (direction != TSP_FP && direction != TSP_BP) return fail(TSP_ERR_INVALID, "direction must be TSP_FP or TSP_BP");
    if ((direction == TSP_FP) != (sub != nullptr))
        return fail(TSP_ERR_INVALID, "sub is required for TSP_FP and must be NULL for TSP_BP");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    if (direction == TSP_FP)
        return launch_fp(pr, st, (const float *)vol, (float *)proj, 0, stream, (const float *)sub, (const float *)mul);
    return launch_bp(pr, st, (float *)vol, (const float *)proj, 0, stream, (const float *)mul);
}

// ---- forward projection from a caller-made transposed copy (multi-GPU: tomosipo_b200/distributed.py) ----------------
extern "C" int tsp_fp_transposed_elems(const tsp_projector *pr, int64_t *elems)
{
    if (!pr || !elems) return fail(TSP_ERR_INVALID, "NULL argument");
    bool need_t = false;
    for (const FPGroup &grp : pr->groups) need_t |= grp.transposed;
    const int ny_pad = (pr->g.ny + 3) / 4 * 4;
    *elems = need_t ? (int64_t)pr->g.nz * pr->g.nx * ny_pad : 0;
    return TSP_OK;
}

extern "C" int tsp_transpose_slices(tsp_projector *pr, const void *vol, void *vol_t, int z0, int z1, int device, void *cuda_stream)
{
    if (!pr || !vol || !vol_t) return fail(TSP_ERR_INVALID, "NULL argument");
    const tsp_geometry &g = pr->g;
    if (z0 < 0 || z1 > g.nz || z1 < z0) return fail(TSP_ERR_INVALID, "slice range [%d, %d) outside the volume", z0, z1);
    if (z1 == z0) return TSP_OK;
    if (z1 - z0 > 65535) return fail(TSP_ERR_INVALID, "too many slices for one transpose launch");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    const int ny_pad = (g.ny + 3) / 4 * 4;
    dim3 grid((g.nx + 31) / 32, (g.ny + 31) / 32, z1 - z0), block(32, 8);
    transpose_xy_kernel<<<grid, block, 0, (cudaStream_t)cuda_stream>>>((const float *)vol + (size_t)z0 * g.nx * g.ny,
                                                                       (float *)vol_t + (size_t)z0 * g.nx * ny_pad, g.nx, g.ny, ny_pad);
    ++pr->launches;
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

// ------------------------------------------------------------ peer memory --
static int enter_device(DeviceGuard &guard, int device)
{
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    return TSP_OK;
}

extern "C" int tsp_peer_alloc(size_t bytes, int device, void **ptr, void *handle64)
{
    if (!ptr || !handle64 || bytes == 0) return fail(TSP_ERR_INVALID, "NULL argument / empty buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard;
    if (int rc = enter_device(guard, device)) return rc;
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(TSP_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, sizeof h);
    *ptr = p;
    return TSP_OK;
}

extern "C" int tsp_peer_open(const void *handle64, int device, void **ptr)
{
    if (!ptr || !handle64) return fail(TSP_ERR_INVALID, "NULL argument");
    DeviceGuard guard;
    if (int rc = enter_device(guard, device)) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TSP_OK;
}

extern "C" int tsp_peer_close(void *ptr, int device)
{
    if (!ptr) return TSP_OK;
    DeviceGuard guard;
    if (int rc = enter_device(guard, device)) return rc;
    CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return TSP_OK;
}

extern "C" int tsp_peer_free(void *ptr, int device)
{
    if (!ptr) return TSP_OK;
    DeviceGuard guard;
    if (int rc = enter_device(guard, device)) return rc;
    CUDA_TRY(cudaFree(ptr));
    return TSP_OK;
}

extern "C" int tsp_push_rows(tsp_projector *pr, int n_jobs, const void *const *src, void *const *dst, const int64_t *rows,
                             const int64_t *width, const int64_t *src_pitch, const int64_t *dst_pitch, int device,
                             void *cuda_stream)
{
    if (n_jobs < 0 || (n_jobs > 0 && (!src || !dst || !rows || !width || !src_pitch || !dst_pitch)))
        return fail(TSP_ERR_INVALID, "NULL argument");
    DeviceGuard guard;
    if (int rc = enter_device(guard, device)) return rc;
    for (int first = 0; first < n_jobs; first += PUSH_MAX_JOBS) {
        PushArgs args;
        int n = 0;
        bool vec4 = true;
        long long most = 0;
        for (int k = first; k < n_jobs && n < PUSH_MAX_JOBS; ++k) {
            if (rows[k] < 0 || width[k] < 0 || src_pitch[k] < width[k] || dst_pitch[k] < width[k])
                return fail(TSP_ERR_INVALID, "job %d: rows %lld, width %lld, pitches %lld / %lld", k, (long long)rows[k],
                            (long long)width[k], (long long)src_pitch[k], (long long)dst_pitch[k]);
            if (rows[k] == 0 || width[k] == 0) continue;
            if (!src[k] || !dst[k]) return fail(TSP_ERR_INVALID, "job %d: NULL buffer", k);
            PushJob &j = args.job[n++];
            j = {(const float *)src[k], (float *)dst[k], rows[k], width[k], src_pitch[k], dst_pitch[k]};
            vec4 = vec4 && width[k] % 4 == 0 && src_pitch[k] % 4 == 0 && dst_pitch[k] % 4 == 0 &&
                   (uintptr_t)src[k] % 16 == 0 && (uintptr_t)dst[k] % 16 == 0;
            most = std::max<long long>(most, rows[k] * width[k]);
        }
        if (n == 0) continue;
        // enough CTAs to fill the SMs across all jobs, no more than one pass of the largest job needs
        const long long per_cta = 512LL * (vec4 ? 16 : 4);
        const int bx = (int)std::max<long long>(1, std::min<long long>((most + per_cta - 1) / per_cta, (4 * 148 + n - 1) / n));
        dim3 grid(bx, n);
        if (vec4)
            push_rows_kernel<4><<<grid, 512, 0, (cudaStream_t)cuda_stream>>>(args);
        else
            push_rows_kernel<1><<<grid, 512, 0, (cudaStream_t)cuda_stream>>>(args);
        if (pr) ++pr->launches;
        CUDA_TRY(cudaGetLastError());
    }
    return TSP_OK;
}

extern "C" int tsp_fp_pre_transposed(tsp_projector *pr, const void *vol, const void *vol_t, void *proj, const void *sub,
                                     const void *mul, int device, void *cuda_stream)
{
    if (!pr || !vol || !proj) return fail(TSP_ERR_INVALID, "NULL argument");
    if ((sub != nullptr) != (mul != nullptr)) return fail(TSP_ERR_INVALID, "sub and mul go together");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    return launch_fp(pr, st, (const float *)vol, (float *)proj, 0, (cudaStream_t)cuda_stream, (const float *)sub, (const float *)mul, 1,
                     (const float *)vol_t);
}

extern "C" int tsp_fp_push(tsp_projector *pr, const void *vol, const void *vol_t, void *proj, const void *sub, const void *mul,
                           int n_peers, void *const *peer_base, const int32_t *row_lo, const int32_t *row_hi,
                           int64_t peer_pitch, int device, void *cuda_stream)
{
    if (!pr || !vol || !proj) return fail(TSP_ERR_INVALID, "NULL argument");
    if ((sub != nullptr) != (mul != nullptr)) return fail(TSP_ERR_INVALID, "sub and mul go together");
    if (n_peers < 0 || n_peers > FP_MAX_PEERS) return fail(TSP_ERR_INVALID, "%d destinations (at most %d)", n_peers, FP_MAX_PEERS);
    if (n_peers > 0 && (!peer_base || !row_lo || !row_hi)) return fail(TSP_ERR_INVALID, "NULL argument");
    const tsp_geometry &g = pr->g;
    if (n_peers > 0 && peer_pitch < (int64_t)g.n_angles * g.det_cols)
        return fail(TSP_ERR_INVALID, "band pitch %lld shorter than this projector's %d angles x %d columns", (long long)peer_pitch,
                    g.n_angles, g.det_cols);
    FPPeers peers;
    peers.n = 0;
    peers.pitch = peer_pitch;
    for (int q = 0; q < n_peers; ++q) {
        if (row_hi[q] <= row_lo[q]) continue;  // this destination reads none of the rows
        if (row_lo[q] < 0 || row_hi[q] > g.det_rows || !peer_base[q])
            return fail(TSP_ERR_INVALID, "destination %d: rows [%d, %d) of %d, buffer %p", q, row_lo[q], row_hi[q], g.det_rows, peer_base[q]);
        peers.base[peers.n] = (float *)peer_base[q];
        peers.lo[peers.n] = row_lo[q];
        peers.hi[peers.n] = row_hi[q];
        ++peers.n;
    }
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    return launch_fp(pr, st, (const float *)vol, (float *)proj, 0, (cudaStream_t)cuda_stream, (const float *)sub, (const float *)mul, 1,
                     (const float *)vol_t, peers.n ? &peers : nullptr);
}

// -------------------------------------------------------------------- FDK --
// Per-angle constants of a circular cone-beam scan, recovered from the vectors (fp64): pitches, source - detector-plane
// distance, principal point, and the constant that turns the library's backprojection of the filtered rows into
//   f = integral d_beta * redundancy * SOD^2 / (SOD - depth)^2 * q :
// the backprojector applies V_vox * SDD^2 / (|u||v| (SOD - depth)^2), the filter is defined at the isocentre pitch
// tau = |u| SOD / SDD, hence  scale = d_beta * SOD^2 |u||v| / (SDD^2 V_vox tau).  (The redundancy weight - 1/2 for a
// full circle, Parker's for a short scan - is applied before the filter by fdk_preweight_kernel.)
static int fdk_angle_constants(const tsp_projector *pr, const double *angle_weights, std::vector<FDKAngle> &tab)
{
    const tsp_geometry &g = pr->g;
    if (g.kind != TSP_KIND_CONE_VEC) return fail(TSP_ERR_INVALID, "FDK needs a cone-beam projection geometry");
    const double vox = pr->sigma[0] * pr->sigma[1] * pr->sigma[2];
    tab.resize(g.n_angles);
    for (int a = 0; a < g.n_angles; ++a) {
        const double *w = pr->vectors.data() + 12 * (size_t)a;
        const double *src = w, *det = w + 3, *u = w + 6, *v = w + 9;
        const double pu = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        const double pv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(nn > 0.0) || !(pu > 0.0) || !(pv > 0.0)) return fail(TSP_ERR_INVALID, "degenerate detector at angle %d", a);
        double h = 0.0, hs = 0.0;
        for (int i = 0; i < 3; ++i) {
            n[i] /= nn;
            h += (det[i] - src[i]) * n[i];
            hs += (0.5 * (g.win_min[i] + g.win_max[i]) - src[i]) * n[i];
        }
        const double sdd = std::fabs(h), sod = std::fabs(hs);
        double ppu = 0.0, ppv = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double foot = src[i] + h * n[i] - det[i];
            ppu += foot * u[i];
            ppv += foot * v[i];
        }
        FDKAngle &t = tab[a];
        t.pu = (float)pu; t.pv = (float)pv; t.sdd = (float)sdd;
        t.ppu = (float)(ppu / (pu * pu)); t.ppv = (float)(ppv / (pv * pv));
        const double dbeta = angle_weights ? angle_weights[a] : 2.0 * M_PI / g.n_angles;
        const double tau = pu * sod / sdd;
        t.scale = (float)(dbeta * sod * sod * pu * pv / (sdd * sdd * vox * tau));
    }
    return TSP_OK;
}

// stage: 0 = pre-weight + pad   (in: proj [V][A][U],            out: padded rows [V][A][pitch])
//        1 = ramp multiply      (in/out: spectrum float2 [V][A][pitch], pitch = nfft / 2 + 1 bins; aux = nfft)
//        2 = crop + scale       (in: filtered rows [V][A][pitch], out: q [V][A][U])
extern "C" int tsp_fdk_stage(tsp_projector *pr, int stage, const void *in, void *out, int pitch, int aux,
                             const void *redundancy, const double *angle_weights, int device, void *cuda_stream)
{
    if (!pr || !out || (stage != 1 && !in)) return fail(TSP_ERR_INVALID, "NULL argument");
    if (stage < 0 || stage > 2) return fail(TSP_ERR_INVALID, "stage must be 0, 1 or 2");
    const int ndev = tsp_device_count();
    if (ndev == 0) return fail(TSP_ERR_CUDA, "no CUDA device available (libtsproj has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TSP_ERR_INVALID, "device %d out of range [0, %d)", device, ndev);
    const tsp_geometry &g = pr->g;
    if (pitch < (stage == 1 ? 2 : g.det_cols)) return fail(TSP_ERR_INVALID, "pitch %d too small", pitch);
    const long long rows = (long long)g.det_rows * g.n_angles;
    if (rows > 2147483647LL) return fail(TSP_ERR_INVALID, "too many detector rows x angles for the FDK grid");
    DeviceGuard guard;
    if (guard.enter(device) != 0) return fail(TSP_ERR_CUDA, "cannot switch to device %d", device);
    DeviceState *st = nullptr;
    if (int rc = get_device_state(pr, device, &st)) return rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    PoolScratch tab_d;
    tab_d.stream = stream;
    if (stage != 1) {
        std::vector<FDKAngle> tab;
        if (int rc = fdk_angle_constants(pr, angle_weights, tab)) return rc;
        CUDA_TRY(pool_alloc(st, &tab_d.p, tab.size() * sizeof(FDKAngle), stream));
        // pageable source: the copy is staged by the runtime before the call returns, `tab` may go out of scope
        CUDA_TRY(cudaMemcpyAsync(tab_d.p, tab.data(), tab.size() * sizeof(FDKAngle), cudaMemcpyHostToDevice, stream));
    }
    if (stage == 0) {
        fdk_preweight_kernel<<<(unsigned)rows, 256, 0, stream>>>((const float *)in, (float *)out, (const FDKAngle *)tab_d.p,
                                                                 (const float *)redundancy, g.det_cols, g.det_rows, g.n_angles, pitch);
    } else if (stage == 1) {
        // frequency response of the band-limited ramp h[0] = 1/4, h[n odd] = -1 / (pi n)^2 on an nfft-periodic grid
        const int nfft = aux, nfreq = pitch;
        if (nfft < 2 * g.det_cols || nfreq != nfft / 2 + 1) return fail(TSP_ERR_INVALID, "ramp: need nfft >= 2 U and nfft / 2 + 1 bins");
        std::vector<float> G(nfreq);
        for (int k = 0; k < nfreq; ++k) {
            double acc = 0.25;
            for (int n = 1; n <= nfft / 2; n += 2) {
                const double hn = -1.0 / (M_PI * M_PI * (double)n * (double)n);
                // n and nfft - n are both odd-indexed images of the same tap unless n == nfft / 2
                acc += (2 * n == nfft ? 1.0 : 2.0) * hn * std::cos(2.0 * M_PI * (double)k * (double)n / (double)nfft);
            }
            G[k] = (float)acc;
        }
        PoolScratch G_d;
        G_d.stream = stream;
        CUDA_TRY(pool_alloc(st, &G_d.p, G.size() * sizeof(float), stream));
        CUDA_TRY(cudaMemcpyAsync(G_d.p, G.data(), G.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        fdk_ramp_kernel<<<(unsigned)rows, 256, 0, stream>>>((float2 *)out, (const float *)G_d.p, nfreq, (size_t)rows);
    } else {
        fdk_scale_crop_kernel<<<(unsigned)rows, 256, 0, stream>>>((const float *)in, (float *)out, (const FDKAngle *)tab_d.p,
                                                                  g.det_cols, g.n_angles, pitch);
    }
    ++pr->launches;
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

// ------------------------------------------------------ pinned host buffers --
// Page-locked host buffers for the arrays the host-array path itself creates (operator outputs, the float32 copies
// of float64 inputs: tomosipo_b200/links/numpy.py).  cudaMemcpyAsync from / to pageable memory is staged through a
// driver bounce buffer and does not overlap anything; the README loop of the reference (README.md:150-164, BASELINE
// configs[0]) allocates a fresh output per call, so freed buffers are kept on a size-keyed free list (bounded by
// TSP_PINNED_CACHE_MB, default 1024) - cudaHostAlloc itself costs about as much as the pageable copy it avoids.
namespace {
std::mutex g_pin_mu;
std::multimap<size_t, void *> g_pin_free;      // size -> buffer
std::map<void *, size_t> g_pin_live;           // buffer -> size
size_t g_pin_cached = 0;
size_t pin_round(size_t b) { return (b + 65535) / 65536 * 65536; }
}  // namespace

extern "C" void *tsp_host_alloc(size_t bytes)
{
    if (bytes == 0 || tsp_device_count() == 0) return nullptr;
    const size_t sz = pin_round(bytes);
    {
        std::lock_guard<std::mutex> lock(g_pin_mu);
        auto it = g_pin_free.find(sz);
        if (it != g_pin_free.end()) {
            void *p = it->second;
            g_pin_free.erase(it);
            g_pin_cached -= sz;
            g_pin_live[p] = sz;
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, sz, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_pin_mu);
    g_pin_live[p] = sz;
    return p;
}

extern "C" void tsp_host_free(void *p)
{
    if (!p) return;
    size_t sz = 0;
    {
        std::lock_guard<std::mutex> lock(g_pin_mu);
        auto it = g_pin_live.find(p);
        if (it == g_pin_live.end()) return;  // not ours
        sz = it->second;
        g_pin_live.erase(it);
        size_t cap = (size_t)1024 << 20;
        if (const char *e = getenv("TSP_PINNED_CACHE_MB")) cap = (size_t)std::max(0LL, atoll(e)) << 20;
        if (g_pin_cached + sz <= cap) {
            g_pin_free.emplace(sz, p);
            g_pin_cached += sz;
            return;
        }
    }
    cudaFreeHost(p);
}
