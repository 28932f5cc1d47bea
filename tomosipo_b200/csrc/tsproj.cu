// libtsproj: C ABI + host side of the B200-native projector (see include/tsproj.h).
//
// Host work per projector (once, fp64): normalise the geometry to unit voxels
// (SURVEY.md B.0), pick one marching axis per angle, group angles by
// (marching axis, volume layout), and derive the affine voxel->detector maps
// for the backprojector.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bp_kernels.cuh"
#include "fp_kernels.cuh"
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

extern "C" int tsp_projector_create(const tsp_geometry *geometry, tsp_projector **out)
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
        const int m = pick_marching_axis(g.kind, na);
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
    for (auto &kv : groups) pr->groups.push_back(std::move(kv.second));
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
        cudaFree(kv.second.bp_angles);
        cudaFree(kv.second.tmap_ring);
    }
    cudaSetDevice(cur);
    pr->dev.clear();
}

extern "C" void tsp_projector_destroy(tsp_projector *pr)
{
    if (!pr) return;
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
    return TSP_OK;
}

extern "C" int tsp_projector_marching_axes(const tsp_projector *pr, int32_t *axes)
{
    if (!pr || !axes) return fail(TSP_ERR_INVALID, "NULL argument");
    for (size_t i = 0; i < pr->march_axis.size(); ++i) axes[i] = pr->march_axis[i];
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
    CUDA_TRY(cudaMalloc(&st.tmap_ring, DeviceState::kTmapSlots * sizeof(TensorMapBlob)));
    CUDA_TRY(cudaMemcpy(st.fp_angles, pr->fp_angles.data(), A * sizeof(FPAngle), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(st.bp_angles, pr->bp_angles.data(), A * sizeof(BPAngle), cudaMemcpyHostToDevice));
    std::vector<int> lists;
    for (const FPGroup &grp : pr->groups) {
        st.list_offset.push_back(lists.size());
        lists.insert(lists.end(), grp.angles.begin(), grp.angles.end());
    }
    CUDA_TRY(cudaMemcpy(st.fp_lists, lists.data(), lists.size() * sizeof(int), cudaMemcpyHostToDevice));
    // keep freed scratch (the transposed volume copy) cached in the pool
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = &(pr->dev[device] = st);
    return TSP_OK;
}

static int launch_fp(tsp_projector *pr, DeviceState *st, const float *vol, float *proj, int additive,
                     cudaStream_t stream)
{
    const tsp_geometry &g = pr->g;
    const int n[3] = {g.nx, g.ny, g.nz};
    const size_t nvox = (size_t)g.nx * g.ny * g.nz;

    bool need_t = false;
    for (const FPGroup &grp : pr->groups) need_t |= grp.transposed;
    float *vol_t = nullptr;
    if (need_t) {
        CUDA_TRY(cudaMallocAsync(&vol_t, nvox * sizeof(float), stream));
        dim3 grid((g.nx + 31) / 32, (g.ny + 31) / 32, g.nz), block(32, 8);
        transpose_xy_kernel<<<grid, block, 0, stream>>>(vol, vol_t, g.nx, g.ny);
        ++pr->launches;
    }
    pr->fp_uses_transpose = need_t ? 1 : 0;

    // element strides of x, y, z in the two layouts
    const long long stride_native[3] = {1, g.nx, (long long)g.nx * g.ny};
    const long long stride_transp[3] = {g.ny, 1, (long long)g.nx * g.ny};
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
        P.det_ss = g.detector_supersampling;
        P.sigma_m = (float)pr->sigma[grp.march];
        const double rp = pr->sigma[grp.p_axis] / pr->sigma[grp.march];
        const double rq = pr->sigma[grp.q_axis] / pr->sigma[grp.march];
        P.rp2 = (float)(rp * rp);
        P.rq2 = (float)(rq * rq);
        P.offsets_fit_32bit = nvox < (1ull << 31) ? 1 : 0;
        const bool cone = g.kind == TSP_KIND_CONE_VEC;
        const bool ss = g.detector_supersampling > 1;
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
    if (vol_t) CUDA_TRY(cudaFreeAsync(vol_t, stream));
    CUDA_TRY(cudaGetLastError());
    return TSP_OK;
}

// z voxels per thread: long register runs amortise per-angle set-up and footprint staging;
// thin volumes (slabs, cfg 5) use short runs.  TSP_BP_ZPT overrides (tuning aid).
static int bp_zpt_choice(int nz)
{
    if (const char *e = getenv("TSP_BP_ZPT")) {
        const int v = atoi(e);
        if (v == 1 || v == 4 || v == 8 || v == 16) return v;
    }
    if (nz >= 12) return 16;
    if (nz > 4) return 8;
    if (nz > 1) return 4;
    return 1;
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

// Tensor map over the projection stack viewed as (u, angle, v), box = 64 x 1 x rows.
static bool make_proj_tensor_map(const float *proj, int det_u, int n_angles, int det_v, int box_rows, TensorMapBlob *out)
{
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "CUtensorMap is 128 bytes");
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(proj) & 15u) != 0 || (det_u & 3) != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)det_u, (cuuint64_t)n_angles, (cuuint64_t)det_v};
    const cuuint64_t strides[2] = {(cuuint64_t)det_u * 4, (cuuint64_t)det_u * 4 * (cuuint64_t)n_angles};
    const cuuint32_t box[3] = {(cuuint32_t)BP_TMA_PITCH, 1u, (cuuint32_t)box_rows};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (strides[1] >= (1ull << 40)) return false;
    CUresult r = enc(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(proj),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool CONE, int ZPT>
static int launch_bp_tma_one(dim3 grid, cudaStream_t stream, const BPArgs &P, const TensorMapBlob *tmap)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = bp_tma_smem_bytes(ZPT);
    if (dev < 64 && !configured[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(bp_tma_kernel<CONE, ZPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    bp_tma_kernel<CONE, ZPT><<<grid, BP_TMA_THREADS, smem, stream>>>(P, tmap);
    return TSP_OK;
}

static int launch_bp_tma_variant(bool cone, int zpt, dim3 grid, cudaStream_t stream, const BPArgs &P,
                                 const TensorMapBlob *tmap)
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
    }
#undef TSP_BP_CASE
    return fail(TSP_ERR_INVALID, "unsupported z run %d", zpt);
}

static int launch_bp(tsp_projector *pr, DeviceState *st, float *vol, const float *proj, int additive,
                     cudaStream_t stream)
{
    const tsp_geometry &g = pr->g;
    BPArgs P;
    P.proj = proj; P.vol = vol;
    P.nx = g.nx; P.ny = g.ny; P.nz = g.nz;
    P.det_u = g.det_cols; P.det_v = g.det_rows; P.n_angles = g.n_angles;
    P.angles = st->bp_angles;
    P.out_scale = (float)(pr->sigma[0] * pr->sigma[1] * pr->sigma[2]);
    P.additive = additive;
    P.vox_ss = g.voxel_supersampling;
    const bool cone = g.kind == TSP_KIND_CONE_VEC;
    int used_tma = 0;
    if (g.voxel_supersampling > 1) {
        if (g.nz > 65535) return fail(TSP_ERR_INVALID, "voxel supersampling supports nz <= 65535");
        dim3 grid((g.nx + 31) / 32, (g.ny + 7) / 8, g.nz), block(32, 8);
        if (cone) bp_supersample_kernel<true><<<grid, block, 0, stream>>>(P);
        else bp_supersample_kernel<false><<<grid, block, 0, stream>>>(P);
    } else {
        const int zpt = bp_zpt_choice(g.nz);
        const int gz = (g.nz + zpt - 1) / zpt;
        const int gy = (g.ny + BP_TY - 1) / BP_TY;
        if (gz > 65535 || gy > 65535) return fail(TSP_ERR_INVALID, "volume too large for the BP grid");
        dim3 grid((g.nx + BP_TX - 1) / BP_TX, gy, gz), block(BP_TX, BP_TY);
        TensorMapBlob tmap;
        const bool use_tma = !getenv("TSP_BP_NO_TMA") &&
                             make_proj_tensor_map(proj, g.det_cols, g.n_angles, g.det_rows, bp_wv(zpt), &tmap);
        if (use_tma) {
            // The descriptor lives in device memory (a small ring per device, so that
            // back-to-back asynchronous calls never overwrite a descriptor in use).
            TensorMapBlob *slot = st->tmap_ring + (st->tmap_next++ % DeviceState::kTmapSlots);
            CUDA_TRY(cudaMemcpyAsync(slot, &tmap, sizeof tmap, cudaMemcpyHostToDevice, stream));
            if (int rc = launch_bp_tma_variant(cone, zpt, grid, stream, P, slot)) return rc;
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
    if (memory_kind == TSP_MEM_HOST) {
        CUDA_TRY(cudaMallocAsync(&dvol, nvox * batch * sizeof(float), stream));
        CUDA_TRY(cudaMallocAsync(&dproj, npix * batch * sizeof(float), stream));
        // inputs, and the destination too when accumulating
        if (direction == TSP_FP || additive)
            CUDA_TRY(cudaMemcpyAsync(dvol, vol, nvox * batch * sizeof(float), cudaMemcpyHostToDevice, stream));
        if (direction == TSP_BP || additive)
            CUDA_TRY(cudaMemcpyAsync(dproj, proj, npix * batch * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    int rc = TSP_OK;
    for (int b = 0; b < batch && rc == TSP_OK; ++b) {
        if (direction == TSP_FP) rc = launch_fp(pr, st, dvol + b * nvox, dproj + b * npix, additive, stream);
        else rc = launch_bp(pr, st, dvol + b * nvox, dproj + b * npix, additive, stream);
    }
    if (memory_kind == TSP_MEM_HOST) {
        if (rc == TSP_OK) {
            if (direction == TSP_FP)
                CUDA_TRY(cudaMemcpyAsync(proj, dproj, npix * batch * sizeof(float), cudaMemcpyDeviceToHost, stream));
            else
                CUDA_TRY(cudaMemcpyAsync(vol, dvol, nvox * batch * sizeof(float), cudaMemcpyDeviceToHost, stream));
        }
        cudaFreeAsync(dvol, stream);
        cudaFreeAsync(dproj, stream);
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return rc;
}

// ------------------------------------------------------------------- SIRT --
__global__ void sirt_residual_kernel(float *__restrict__ y_tmp, const float *__restrict__ y,
                                     const float *__restrict__ R, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y_tmp[i] = R[i] * (y_tmp[i] - y[i]);
}
__global__ void sirt_update_kernel(float *__restrict__ x, const float *__restrict__ x_tmp,
                                   const float *__restrict__ C, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] -= C[i] * x_tmp[i];
}

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
    float *x_tmp = nullptr;
    CUDA_TRY(cudaMallocAsync(&x_tmp, nvox * sizeof(float), stream));
    int rc = TSP_OK;
    for (int it = 0; it < iterations && rc == TSP_OK; ++it) {
        rc = launch_fp(pr, st, (const float *)x, (float *)y_tmp, 0, stream);
        if (rc) break;
        sirt_residual_kernel<<<148 * 8, 256, 0, stream>>>((float *)y_tmp, (const float *)y, (const float *)R, npix);
        rc = launch_bp(pr, st, x_tmp, (const float *)y_tmp, 0, stream);
        if (rc) break;
        sirt_update_kernel<<<148 * 8, 256, 0, stream>>>((float *)x, x_tmp, (const float *)C, nvox);
        pr->launches += 2;
    }
    cudaFreeAsync(x_tmp, stream);
    if (rc == TSP_OK) CUDA_TRY(cudaGetLastError());
    return rc;
}
