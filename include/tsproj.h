/*
 * tsproj.h -- C ABI of libtsproj.so, the B200-native (sm_100a) 3D tomographic
 * projector that replaces the ASTRA calls behind tomosipo's projection path.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout of ahendriksen/tomosipo v0.6.0).
 *
 * Conventions
 *   - every function returns 0 on success or a negative tsp_status; the
 *     message of the last failure on the calling thread is tsp_last_error().
 *   - volumes are dense float32 [nz][ny][nx]; projections are dense float32
 *     [det_rows][n_angles][det_cols]  (tomosipo/links/base.py:21-31).
 *   - the library never allocates or frees caller buffers.
 *   - TSP_MEM_DEVICE calls are asynchronous on the given CUDA stream;
 *     TSP_MEM_HOST calls return after the result is back in host memory
 *     (ASTRA's synchronous contract, tomosipo/astra.py:146-153).
 *   - a projector is immutable after creation: TSP_MEM_DEVICE calls may be issued
 *     concurrently from several host threads / streams (per-call scratch comes
 *     from the stream-ordered pool, TMA descriptors travel by value), and they
 *     capture into CUDA graphs.
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with TSP_ERR_CUDA.
 */
#ifndef TSPROJ_H
#define TSPROJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSP_VERSION 100

typedef enum {
    TSP_OK = 0,
    TSP_ERR_INVALID = -1, /* bad argument / geometry: maps to ValueError */
    TSP_ERR_CUDA = -2,    /* CUDA runtime failure: maps to RuntimeError */
    TSP_ERR_NOMEM = -3
} tsp_status;

enum { TSP_KIND_CONE_VEC = 0, TSP_KIND_PARALLEL_VEC = 1 };
enum { TSP_FP = 0, TSP_BP = 1 };
enum { TSP_MEM_HOST = 0, TSP_MEM_DEVICE = 1 };

/*
 * The two dicts tomosipo hands to astra.create_projector("cuda3d", pg, vg,
 * options) (tomosipo/astra.py:90-98), flattened:
 *   vg  = VolumeGeometry.to_astra()        tomosipo/geometry/volume.py:286-301
 *         GridColCount = nx, GridRowCount = ny, GridSliceCount = nz,
 *         option.WindowMin/Max{X,Y,Z} = win_min / win_max (x, y, z order)
 *   pg  = {Cone,Parallel}VectorGeometry.to_astra()
 *         tomosipo/geometry/cone_vec.py:174-191, parallel_vec.py:175-192
 *         Vectors[a] = [src|ray (3), det centre (3), det_u (3), det_v (3)],
 *         each triple in ASTRA (x, y, z) order; DetectorRowCount = det_rows,
 *         DetectorColCount = det_cols.
 *   options = {VoxelSuperSampling, DetectorSuperSampling}.
 * Circular 'cone' / 'parallel3d' dicts are converted to vectors by the caller
 * (astra.geom_2vec semantics; see tomosipo_b200/astra_compat.py).
 * The library copies everything it needs; the caller keeps ownership.
 */
typedef struct tsp_geometry {
    int32_t kind;
    int32_t nx, ny, nz;
    double win_min[3];
    double win_max[3];
    int32_t det_rows, det_cols, n_angles;
    const double *vectors; /* n_angles x 12 */
    int32_t voxel_supersampling, detector_supersampling;
} tsp_geometry;

typedef struct tsp_projector tsp_projector;

/* Replaces astra.create_projector("cuda3d", ...)  tomosipo/astra.py:90-98.
 * Host-only; no CUDA call is made until the first tsp_project. */
int tsp_projector_create(const tsp_geometry *geometry, tsp_projector **out);

/* The reference never frees projector ids (tomosipo/Operator.py:167-172);
 * this library lets the owner do so. */
void tsp_projector_destroy(tsp_projector *projector);

/*
 * Replaces astra.experimental.direct_FPBP3D(projector, vol, proj, mode, "FP"|"BP")
 * (tomosipo/astra.py:147-153).
 *   direction   TSP_FP: proj (+)= A vol;  TSP_BP: vol (+)= A^T proj
 *   additive    0 = MODE_SET (overwrite), 1 = MODE_ADD      astra.py:132-135
 *   vol, proj   float32; `batch` consecutive dense volumes / projection
 *               stacks (batch == 1 is the reference behaviour; batch > 1
 *               serves the leading-dimension loop of
 *               tomosipo/torch_support.py:49-53,70-74 in one call)
 *   memory_kind TSP_MEM_HOST  (numpy / CPU tensors; ASTRA's ndarray path) or
 *               TSP_MEM_DEVICE (astra.data3d.GPULink(ptr, x, y, z, pitch = 4x),
 *               tomosipo/links/torch.py:94-106)
 *   device      CUDA device ordinal owning the pointers (links/torch.py:150-152);
 *               for TSP_MEM_HOST the device to compute on
 *   cuda_stream cudaStream_t to launch on; NULL = legacy default stream
 */
int tsp_project(tsp_projector *projector, int direction, int additive, void *vol, void *proj,
                int batch, int memory_kind, int device, void *cuda_stream);

/*
 * Host arrays over several GPUs: what `astra.set_gpu_index([0, 1, ...])` switches on for ndarray inputs in the
 * reference (doc/topics/operator.rst:226-245, "the ASTRA-toolbox will automatically divide all projection and
 * backprojection over all four GPUs").  vol / proj are HOST arrays (pinned memory overlaps best); the chunks of the
 * host-array plan are dealt out to `devices` and each device runs the copy / compute pipeline over its share from
 * its own host thread.  Returns when the result is in host memory.  Additive calls and problems too small to be
 * chunked run on devices[0].
 *
 * Device memory of every TSP_MEM_HOST call is bounded: when the two arrays would not fit the budget
 * (TSP_HOST_MEM_CAP_MB, default: the device's free memory minus a reserve) the pipeline works out of a ring of three
 * chunk-sized buffer pairs instead of whole arrays (out-of-core, as ASTRA's CompositeGeometryManager splits jobs).
 */
int tsp_project_multi(tsp_projector *projector, int direction, int additive, void *vol, void *proj,
                      const int *devices, int n_devices);

/* Introspection used by the tests and by bench.py.  The *_uses_* / host_* fields describe the LAST call on the
 * projector; with concurrent callers they are informational only. */
typedef struct tsp_projector_info {
    int32_t n_angles;
    int32_t n_march_x, n_march_y, n_march_z; /* FP marching-axis census */
    double voxel_size[3];                     /* x, y, z */
    int64_t kernel_launches;                  /* kernels launched so far by this projector */
    int32_t bp_uses_tma;                      /* last BP used TMA-staged footprints */
    int32_t fp_uses_transpose;                /* last FP built an (x<->y) transposed volume */
    int32_t fp_uses_tma;                      /* last FP used the TMA-staged kernel for >= 1 angle group */
    int32_t host_pipelined;                   /* last TSP_MEM_HOST call ran the chunked copy/compute pipeline */
    int32_t host_ring;                        /* ... out of a bounded ring of chunk buffers (out-of-core mode) */
    int32_t host_devices;                     /* ... on this many devices (tsp_project_multi) */
} tsp_projector_info;
int tsp_projector_get_info(const tsp_projector *projector, tsp_projector_info *info);

/* Plan of the host-array pipeline (TSP_MEM_HOST calls on large problems are cut into sub-problems
 * whose transfers overlap the kernels: what ASTRA's CompositeGeometryManager does for ndarray inputs,
 * reference doc/topics/operator.rst:226-261).  Chunk k of `direction` works on volume slices
 * [z0, z1) and detector rows [v0, v1): out[4k .. 4k+3] = z0, z1, v0, v1, in execution order.
 * Returns the number of chunks (0: this geometry is not pipelined), writes at most max_chunks of
 * them.  Host-only (no CUDA call); used by the tests to check the bounds against the geometry. */
int tsp_projector_host_plan(tsp_projector *projector, int direction, int32_t *out, int max_chunks);

/* The backprojector's voxel -> detector map of one angle, evaluated on the host from the very table
 * the kernels read: for a point xyz of the world frame (the frame of win_min / win_max and of the
 * vectors; x, y, z order) out = {U, V, w}, detector pixel (iv, iu) spanning [iu, iu+1) x [iv, iv+1)
 * and w the ray-density weight (voxel volume not included).  This is what the reference exposes as
 * {Cone,Parallel}VectorGeometry.project_point (tomosipo/geometry/cone_vec.py:306-326,
 * parallel_vec.py:313-330; known answers tests/geometry/test_cone_vec.py:143-173), which the parity
 * tests hold it against.  Host-only (no CUDA call). */
int tsp_projector_bp_map(const tsp_projector *projector, int angle, const double *xyz, double *out);

/* Per-angle FP marching axis (0 = x, 1 = y, 2 = z), for parity checks. */
int tsp_projector_marching_axes(const tsp_projector *projector, int32_t *axes);

/* Replaces astra.use_cuda()  (reference tests/__init__.py:7). */
int tsp_cuda_available(void);
int tsp_device_count(void);
int tsp_version(void);
const char *tsp_last_error(void);

/*
 * Fused SIRT iteration on device-resident data (the
 * loop of notebooks/sirt_benchmark.py:130-136):
 *     y_tmp = R * (A x - y);   x -= C * A^T y_tmp
 * x, C: [nz][ny][nx];  y, R, y_tmp: [det_rows][n_angles][det_cols].
 * Runs `iterations` iterations asynchronously on the stream.
 */
int tsp_sirt(tsp_projector *projector, void *x, const void *y, const void *R, const void *C,
             void *y_tmp, int iterations, int device, void *cuda_stream);

/*
 * One half of that iteration on device-resident data, for callers that own the
 * loop (the z-slab / angle-block sharded SIRT of tomosipo_b200/distributed.py,
 * where a collective sits between the two halves):
 *   TSP_FP:  proj = mul * (A vol - sub)   sub, mul: [det_rows][n_angles][det_cols]
 *            (the `y_tmp = A(x); y_tmp -= y; y_tmp *= R` lines of
 *            notebooks/sirt_benchmark.py:131-133 formed in the projector's store)
 *   TSP_BP:  vol -= mul * (A^T proj)      mul: [nz][ny][nx]; sub must be NULL
 *            (`x_tmp = A.T(y_tmp); x_tmp *= C; x -= x_tmp`, :134-136)
 * Asynchronous on the stream.
 */
int tsp_project_fused(tsp_projector *projector, int direction, void *vol, void *proj, const void *sub,
                      const void *mul, int device, void *cuda_stream);

/*
 * Forward projection from a caller-made transposed copy.  Angles that march along x read an (x <-> y)-transposed copy
 * of the volume, which tsp_project makes per call.  A caller that assembles the volume piecewise - the multi-GPU
 * operator, whose z-chunks arrive one all-gather at a time (tomosipo_b200/distributed.py) - can transpose each piece as
 * it arrives, behind the transfers still in flight, and hand the finished copy in:
 *   tsp_fp_transposed_elems   number of floats of the transposed copy ([nz][nx][ny rounded up to 4]); 0 = not needed
 *   tsp_transpose_slices      slices [z0, z1) of `vol` (dense [nz][ny][nx]) into `vol_t`; asynchronous on the stream
 *   tsp_fp_pre_transposed     proj = A vol, or mul * (A vol - sub) when sub / mul are given (see tsp_project_fused);
 *                             vol_t may be NULL (then the library transposes itself, as tsp_project does)
 */
int tsp_fp_transposed_elems(const tsp_projector *projector, int64_t *elems);
int tsp_transpose_slices(tsp_projector *projector, const void *vol, void *vol_t, int z0, int z1, int device,
                         void *cuda_stream);
int tsp_fp_pre_transposed(tsp_projector *projector, const void *vol, const void *vol_t, void *proj, const void *sub,
                          const void *mul, int device, void *cuda_stream);

/*
 * Peer memory for the multi-GPU operator (one process per GPU; the reference has no multi-GPU path for device arrays:
 * "you must distribute the data over multiple GPUs yourself", doc/topics/operator.rst:246-249).  Its backprojection
 * by row exchange (tomosipo_b200/distributed.py) needs, on every rank, a band of detector rows of every other rank's
 * angle block; the ranks store those rows straight into each other's band buffers over NVLink:
 *   tsp_peer_alloc   cudaMalloc a buffer on `device` and export its 64-byte CUDA IPC handle
 *   tsp_peer_open    map another process's buffer (its handle) into this process, peer access enabled lazily
 *   tsp_peer_close   unmap it;  tsp_peer_free: release an own buffer (after the peers have closed it)
 *   tsp_push_rows    n_jobs strided copies in one kernel, asynchronous on the stream: job k copies rows[k] rows of
 *                    width[k] floats from src[k] (row pitch src_pitch[k] floats) to dst[k] (row pitch dst_pitch[k]);
 *                    dst may be peer memory.  `projector` (may be NULL) only counts the launch.
 *   tsp_fp_push      tsp_fp_pre_transposed whose store also writes every value of detector row v in
 *                    [row_lo[q], row_hi[q]) to peer_base[q][(v - row_lo[q]) * peer_pitch + angle * det_cols + u]
 *                    (q < n_peers <= 16; peer_base[q] already points at this rank's first angle in rank q's band
 *                    buffer): the exchange rides on the projector's own stores, tile by tile, and needs no pass of its
 *                    own.  SET mode, no detector supersampling.
 * Completion across ranks is the caller's business (a collective on the same stream).
 */
int tsp_peer_alloc(size_t bytes, int device, void **ptr, void *handle64);
int tsp_peer_open(const void *handle64, int device, void **ptr);
int tsp_peer_close(void *ptr, int device);
int tsp_peer_free(void *ptr, int device);
int tsp_fp_push(tsp_projector *projector, const void *vol, const void *vol_t, void *proj, const void *sub, const void *mul,
                int n_peers, void *const *peer_base, const int32_t *row_lo, const int32_t *row_hi, int64_t peer_pitch,
                int device, void *cuda_stream);
int tsp_push_rows(tsp_projector *projector, int n_jobs, const void *const *src, void *const *dst, const int64_t *rows,
                  const int64_t *width, const int64_t *src_pitch, const int64_t *dst_pitch, int device,
                  void *cuda_stream);

/*
 * Page-locked host buffers for arrays the host-array path creates itself (the operator's outputs and the float32
 * copies of float64 inputs; reference tomosipo/links/numpy.py:26-32,121-144 allocates them with numpy): transfers
 * from / to pageable memory do not overlap the kernels.  Freed buffers are cached by size (TSP_PINNED_CACHE_MB,
 * default 1024).  tsp_host_alloc returns NULL without a CUDA device - the caller then uses ordinary memory.
 */
void *tsp_host_alloc(size_t bytes);
void tsp_host_free(void *buffer);

/*
 * The element-wise passes of FDK around the ramp filter's FFT, replacing the filtering half of
 * astra.experimental.accumulate_FDK (tomosipo/astra.py:404-406); the backprojection half is tsp_project(TSP_BP).
 * Device arrays, asynchronous on the stream.  All rows are [det_rows][n_angles][pitch]:
 *   stage 0  out[..][0 .. pitch) = proj * cosine weight * redundancy, zero beyond det_cols (the FFT's zero padding);
 *            in = projections [det_rows][n_angles][det_cols]; redundancy: device float [n_angles][det_cols]
 *            (Parker weights of a short scan) or NULL = 1/2 (full circle)
 *   stage 1  out (float2 spectrum, pitch = aux / 2 + 1 bins of an aux-point real FFT) *= band-limited ramp; in unused
 *   stage 2  out [det_rows][n_angles][det_cols] = in[..][0 .. det_cols) * per-angle constant; angle_weights: host
 *            doubles [n_angles], the angular step d_beta per angle in radians, or NULL = 2 pi / n_angles
 */
int tsp_fdk_stage(tsp_projector *projector, int stage, const void *in, void *out, int pitch, int aux,
                  const void *redundancy, const double *angle_weights, int device, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TSPROJ_H */
