"""ctypes binding of ``libtsproj.so`` (C ABI: ``include/tsproj.h``).

This is the process-internal FFI boundary that replaces the Cython calls
``astra.create_projector`` / ``astra.experimental.direct_FPBP3D`` of the
reference (``tomosipo/astra.py:90-98,147-153``).  There is no CPU fallback:
if the library is missing, or no CUDA device is usable, projection raises.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# TSPROJ_LIB: load another build of the same sources (kernel tuning experiments)
LIB_PATH = os.environ.get("TSPROJ_LIB") or os.path.join(_PKG_DIR, "libtsproj.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

KIND_CONE_VEC = 0
KIND_PARALLEL_VEC = 1
FP, BP = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

ERR_INVALID = -1
ERR_CUDA = -2
ERR_NOMEM = -3


class tsp_geometry(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("nz", ctypes.c_int32),
        ("win_min", ctypes.c_double * 3),
        ("win_max", ctypes.c_double * 3),
        ("det_rows", ctypes.c_int32),
        ("det_cols", ctypes.c_int32),
        ("n_angles", ctypes.c_int32),
        ("vectors", ctypes.POINTER(ctypes.c_double)),
        ("voxel_supersampling", ctypes.c_int32),
        ("detector_supersampling", ctypes.c_int32),
    ]


class tsp_projector_info(ctypes.Structure):
    _fields_ = [
        ("n_angles", ctypes.c_int32),
        ("n_march_x", ctypes.c_int32),
        ("n_march_y", ctypes.c_int32),
        ("n_march_z", ctypes.c_int32),
        ("voxel_size", ctypes.c_double * 3),
        ("kernel_launches", ctypes.c_int64),
        ("bp_uses_tma", ctypes.c_int32),
        ("fp_uses_transpose", ctypes.c_int32),
        ("fp_uses_tma", ctypes.c_int32),
        ("host_pipelined", ctypes.c_int32),
        ("host_ring", ctypes.c_int32),
        ("host_devices", ctypes.c_int32),
    ]


#: every symbol include/tsproj.h declares
EXPORTED_SYMBOLS = (
    "tsp_projector_create",
    "tsp_projector_destroy",
    "tsp_project",
    "tsp_projector_get_info",
    "tsp_projector_marching_axes",
    "tsp_cuda_available",
    "tsp_device_count",
    "tsp_version",
    "tsp_last_error",
    "tsp_sirt",
    "tsp_project_fused",
    "tsp_projector_host_plan",
    "tsp_projector_bp_map",
    "tsp_project_multi",
    "tsp_fdk_stage",
    "tsp_host_alloc",
    "tsp_host_free",
    "tsp_fp_transposed_elems",
    "tsp_transpose_slices",
    "tsp_fp_pre_transposed",
    "tsp_peer_alloc",
    "tsp_peer_open",
    "tsp_peer_close",
    "tsp_peer_free",
    "tsp_push_rows",
    "tsp_fp_push",
)


def build(force=False, verbose=False):
    """Compile ``libtsproj.so`` in-tree with nvcc for sm_100a."""
    cmd = ["make", "-C", CSRC_DIR] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libtsproj.so failed")
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (once).  Raises ``ImportError`` if absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C tomosipo_b200/csrc`. tomosipo_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        L.tsp_projector_create.argtypes = [ctypes.POINTER(tsp_geometry), ctypes.POINTER(vp)]
        L.tsp_projector_create.restype = ctypes.c_int
        L.tsp_projector_destroy.argtypes = [vp]
        L.tsp_projector_destroy.restype = None
        L.tsp_project.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, vp]
        L.tsp_project.restype = ctypes.c_int
        L.tsp_projector_get_info.argtypes = [vp, ctypes.POINTER(tsp_projector_info)]
        L.tsp_projector_get_info.restype = ctypes.c_int
        L.tsp_projector_marching_axes.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
        L.tsp_projector_marching_axes.restype = ctypes.c_int
        L.tsp_cuda_available.restype = ctypes.c_int
        L.tsp_device_count.restype = ctypes.c_int
        L.tsp_version.restype = ctypes.c_int
        L.tsp_last_error.restype = ctypes.c_char_p
        L.tsp_sirt.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        L.tsp_sirt.restype = ctypes.c_int
        L.tsp_project_fused.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp]
        L.tsp_project_fused.restype = ctypes.c_int
        L.tsp_projector_host_plan.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        L.tsp_projector_host_plan.restype = ctypes.c_int
        L.tsp_project_multi.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.tsp_project_multi.restype = ctypes.c_int
        L.tsp_fdk_stage.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, vp,
                                    ctypes.POINTER(ctypes.c_double), ctypes.c_int, vp]
        L.tsp_fdk_stage.restype = ctypes.c_int
        L.tsp_fp_transposed_elems.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
        L.tsp_fp_transposed_elems.restype = ctypes.c_int
        L.tsp_transpose_slices.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.tsp_transpose_slices.restype = ctypes.c_int
        L.tsp_fp_pre_transposed.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.tsp_fp_pre_transposed.restype = ctypes.c_int
        L.tsp_host_alloc.argtypes = [ctypes.c_size_t]
        L.tsp_host_alloc.restype = vp
        L.tsp_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp), ctypes.c_char_p]
        L.tsp_peer_alloc.restype = ctypes.c_int
        L.tsp_peer_open.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
        L.tsp_peer_open.restype = ctypes.c_int
        L.tsp_peer_close.argtypes = [vp, ctypes.c_int]
        L.tsp_peer_close.restype = ctypes.c_int
        L.tsp_peer_free.argtypes = [vp, ctypes.c_int]
        L.tsp_peer_free.restype = ctypes.c_int
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.tsp_push_rows.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), i64p, i64p, i64p, i64p,
                                    ctypes.c_int, vp]
        L.tsp_push_rows.restype = ctypes.c_int
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.tsp_fp_push.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.POINTER(vp), i32p, i32p, ctypes.c_int64,
                                  ctypes.c_int, vp]
        L.tsp_fp_push.restype = ctypes.c_int
        L.tsp_host_free.argtypes = [vp]
        L.tsp_host_free.restype = None
        f64p = ctypes.POINTER(ctypes.c_double)
        L.tsp_projector_bp_map.argtypes = [vp, ctypes.c_int, f64p, f64p]
        L.tsp_projector_bp_map.restype = ctypes.c_int
        _lib = L
        return _lib


def _check(rc):
    if rc == 0:
        return
    msg = lib().tsp_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


class _PinnedOwner:
    """Returns a page-locked buffer to the library's cache when the last array on it is gone."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        ptr, self.ptr = self.ptr, None
        if ptr and _lib is not None:
            try:
                _lib.tsp_host_free(ctypes.c_void_p(ptr))
            except Exception:  # interpreter shutdown
                pass


#: arrays smaller than this are not worth a page-locked buffer
PINNED_MIN_BYTES = 1 << 20


def pinned_empty(shape, dtype=np.float32):
    """Uninitialised array in page-locked host memory (``tsp_host_alloc``), or an ordinary ``np.empty`` when the
    array is small, pinning is switched off (``TSP_NO_PINNED=1``) or no CUDA device / library is available."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < PINNED_MIN_BYTES or os.environ.get("TSP_NO_PINNED"):
        return np.empty(shape, dtype=dtype)
    try:
        ptr = lib().tsp_host_alloc(nbytes)
    except (ImportError, OSError):
        ptr = None
    if not ptr:
        return np.empty(shape, dtype=dtype)
    buf = (ctypes.c_byte * nbytes).from_address(ptr)
    buf._owner = _PinnedOwner(ptr)  # np.frombuffer keeps `buf` (and with it the owner) alive as the array's base
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def cuda_available():
    """Replacement of ``astra.use_cuda()`` (reference ``tests/__init__.py:7``)."""
    try:
        return bool(lib().tsp_cuda_available())
    except (ImportError, OSError):
        return False


class Projector:
    """Owner of one ``tsp_projector`` handle.

    ``vectors`` are ASTRA 12-column rows in (x, y, z) order, ``window`` is
    ((minx, maxx), (miny, maxy), (minz, maxz)) -- exactly the content of the
    dicts ``create_astra_projector`` builds in the reference.
    """

    def __init__(self, kind, vol_shape_zyx, window_xyz, det_shape_vu, vectors,
                 voxel_supersampling=1, detector_supersampling=1):
        vec = np.ascontiguousarray(vectors, dtype=np.float64)
        if vec.ndim != 2 or vec.shape[1] != 12:
            raise ValueError(f"Expected vectors of shape (num_angles, 12). Got {vec.shape}")
        g = tsp_geometry()
        g.kind = int(kind)
        g.nz, g.ny, g.nx = (int(s) for s in vol_shape_zyx)
        for i in range(3):
            g.win_min[i] = float(window_xyz[i][0])
            g.win_max[i] = float(window_xyz[i][1])
        g.det_rows, g.det_cols = int(det_shape_vu[0]), int(det_shape_vu[1])
        g.n_angles = vec.shape[0]
        g.vectors = vec.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        g.voxel_supersampling = int(voxel_supersampling)
        g.detector_supersampling = int(detector_supersampling)
        handle = ctypes.c_void_p()
        _check(lib().tsp_projector_create(ctypes.byref(g), ctypes.byref(handle)))
        self._handle = handle
        self.kind = int(kind)
        self.vol_shape = (g.nz, g.ny, g.nx)
        self.proj_shape = (g.det_rows, g.n_angles, g.det_cols)
        self.n_angles = g.n_angles

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and _lib is not None:
            try:
                _lib.tsp_projector_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    def project(self, direction, additive, vol_ptr, proj_ptr, memory_kind, device=0, stream=0, batch=1):
        """Raw call: pointers are integers (host or device addresses)."""
        _check(lib().tsp_project(self._handle, int(direction), int(bool(additive)), ctypes.c_void_p(vol_ptr),
                                 ctypes.c_void_p(proj_ptr), int(batch), int(memory_kind), int(device),
                                 ctypes.c_void_p(stream)))

    def project_multi(self, direction, additive, vol_ptr, proj_ptr, devices):
        """HOST arrays divided over several GPUs (``astra.set_gpu_index([...])`` in the reference)."""
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        _check(lib().tsp_project_multi(self._handle, int(direction), int(bool(additive)), ctypes.c_void_p(vol_ptr),
                                       ctypes.c_void_p(proj_ptr), devs, len(devices)))

    def fp_transposed_elems(self):
        """Floats of the (x <-> y)-transposed volume copy the forward projector reads (0: none of its angles needs one)."""
        n = ctypes.c_int64(0)
        _check(lib().tsp_fp_transposed_elems(self._handle, ctypes.byref(n)))
        return int(n.value)

    def fp_push(self, vol_ptr, vol_t_ptr, proj_ptr, sub_ptr, mul_ptr, peers, pitch, device=0, stream=0):
        """``fp_pre_transposed`` whose store also writes the detector rows ``[lo, hi)`` of every ``(base_ptr, lo, hi)``
        in ``peers`` into that (peer-memory) band buffer with row pitch ``pitch`` floats (``tsp_fp_push``)."""
        vp = ctypes.c_void_p
        n = len(peers)
        base = (vp * n)(*[vp(b) for b, _, _ in peers])
        lo = (ctypes.c_int32 * n)(*[int(a) for _, a, _ in peers])
        hi = (ctypes.c_int32 * n)(*[int(b) for _, _, b in peers])
        _check(lib().tsp_fp_push(self._handle, vp(vol_ptr), vp(vol_t_ptr), vp(proj_ptr), vp(sub_ptr), vp(mul_ptr), n, base, lo, hi,
                                 int(pitch), int(device), vp(stream)))

    def push_rows(self, jobs, device=0, stream=0):
        """One kernel of strided row copies (``tsp_push_rows``); destinations may be peer memory."""
        push_rows(jobs, device=device, stream=stream, projector=self)

    def transpose_slices(self, vol_ptr, vol_t_ptr, z0, z1, device=0, stream=0):
        vp = ctypes.c_void_p
        _check(lib().tsp_transpose_slices(self._handle, vp(vol_ptr), vp(vol_t_ptr), int(z0), int(z1), int(device), vp(stream)))

    def fp_pre_transposed(self, vol_ptr, vol_t_ptr, proj_ptr, sub_ptr=None, mul_ptr=None, device=0, stream=0):
        """``proj = A vol`` (or ``mul * (A vol - sub)``) reading the caller's transposed copy ``vol_t_ptr``."""
        vp = ctypes.c_void_p
        _check(lib().tsp_fp_pre_transposed(self._handle, vp(vol_ptr), vp(vol_t_ptr), vp(proj_ptr), vp(sub_ptr), vp(mul_ptr),
                                           int(device), vp(stream)))

    def fdk_stage(self, stage, in_ptr, out_ptr, pitch, aux, redundancy_ptr, angle_weights, device=0, stream=0):
        """One element-wise pass of the FDK pre-filter on device pointers (``tsp_fdk_stage`` in include/tsproj.h)."""
        vp = ctypes.c_void_p
        aw = None
        if angle_weights is not None:
            aw = np.ascontiguousarray(angle_weights, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        _check(lib().tsp_fdk_stage(self._handle, int(stage), vp(in_ptr), vp(out_ptr), int(pitch), int(aux),
                                   vp(redundancy_ptr), aw, int(device), vp(stream)))

    def sirt(self, x_ptr, y_ptr, r_ptr, c_ptr, ytmp_ptr, iterations, device=0, stream=0):
        vp = ctypes.c_void_p
        _check(lib().tsp_sirt(self._handle, vp(x_ptr), vp(y_ptr), vp(r_ptr), vp(c_ptr), vp(ytmp_ptr),
                              int(iterations), int(device), vp(stream)))

    def project_fused(self, direction, vol_ptr, proj_ptr, sub_ptr, mul_ptr, device=0, stream=0):
        """``proj = mul * (A vol - sub)`` (FP) or ``vol -= mul * A^T proj`` (BP, ``sub_ptr`` = None) on device pointers."""
        vp = ctypes.c_void_p
        _check(lib().tsp_project_fused(self._handle, int(direction), vp(vol_ptr), vp(proj_ptr), vp(sub_ptr), vp(mul_ptr),
                                       int(device), vp(stream)))

    def host_plan(self, direction):
        """``[(z0, z1, v0, v1)]`` of the host-array pipeline for FP / BP, in execution order ([] = not pipelined)."""
        n = lib().tsp_projector_host_plan(self._handle, int(direction), None, 0)
        if n < 0:
            _check(n)
        buf = (ctypes.c_int32 * (4 * max(n, 1)))()
        n = lib().tsp_projector_host_plan(self._handle, int(direction), buf, n)
        return [tuple(buf[4 * k: 4 * k + 4]) for k in range(n)]

    def bp_map(self, angle, xyz):
        """(U, V, weight) of a world-frame point (x, y, z) on the detector of ``angle`` (host-only)."""
        p = np.ascontiguousarray(xyz, dtype=np.float64)
        out = np.zeros(3)
        f64p = ctypes.POINTER(ctypes.c_double)
        _check(lib().tsp_projector_bp_map(self._handle, int(angle), p.ctypes.data_as(f64p), out.ctypes.data_as(f64p)))
        return out

    def info(self):
        info = tsp_projector_info()
        _check(lib().tsp_projector_get_info(self._handle, ctypes.byref(info)))
        return info

    def marching_axes(self):
        axes = np.zeros(self.n_angles, dtype=np.int32)
        _check(lib().tsp_projector_marching_axes(self._handle, axes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return axes


# ------------------------------------------------------------ peer memory --
def peer_alloc(nbytes, device):
    """``(device pointer, 64-byte CUDA IPC handle)`` of a new cudaMalloc'ed buffer (``tsp_peer_alloc``)."""
    ptr = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    _check(lib().tsp_peer_alloc(int(nbytes), int(device), ctypes.byref(ptr), handle))
    return int(ptr.value), handle.raw


def peer_open(handle, device):
    """Device pointer of another process's buffer mapped into this one (``tsp_peer_open``)."""
    ptr = ctypes.c_void_p()
    _check(lib().tsp_peer_open(bytes(handle), int(device), ctypes.byref(ptr)))
    return int(ptr.value)


def peer_close(ptr, device):
    _check(lib().tsp_peer_close(ctypes.c_void_p(ptr), int(device)))


def peer_free(ptr, device):
    _check(lib().tsp_peer_free(ctypes.c_void_p(ptr), int(device)))


def push_rows(jobs, device=0, stream=0, projector=None):
    """``jobs``: ``[(src_ptr, dst_ptr, rows, width, src_pitch, dst_pitch)]`` in floats; one launch, asynchronous."""
    n = len(jobs)
    vp = ctypes.c_void_p
    src = (vp * n)(*[vp(j[0]) for j in jobs])
    dst = (vp * n)(*[vp(j[1]) for j in jobs])
    cols = [(ctypes.c_int64 * n)(*[int(j[k]) for j in jobs]) for k in (2, 3, 4, 5)]
    _check(lib().tsp_push_rows(projector._handle if projector is not None else None, n, src, dst, *cols, int(device), vp(stream)))


class DeviceBuffer:
    """A float32 device buffer owned by the library, visible to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr, shape):
        self.ptr, self.shape = int(ptr), tuple(int(v) for v in shape)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 2,
                                         "strides": None}
