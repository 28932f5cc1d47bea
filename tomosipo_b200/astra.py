"""Backend boundary: geometry conversion and direct projection.

API mirror of the reference's ``tomosipo/astra.py`` -- the file whose ASTRA
calls this package replaces:

=============================================  ================================
reference (ASTRA)                              here (libtsproj C ABI)
=============================================  ================================
``astra.create_projector('cuda3d', pg, vg)``   ``tsp_projector_create``
``astra.experimental.direct_FPBP3D(...)``      ``tsp_project``
``astra.data3d.GPULink(ptr, x, y, z, pitch)``  ``links.base.RawBuffer``
=============================================  ================================

``to_astra`` / ``from_astra`` still produce / accept ASTRA-format dicts, so
geometries interoperate with code written against ASTRA.
"""
import numpy as np

import tomosipo_b200 as ts
from . import _backend, astra_compat


# ------------------------------------------------------------ geometry dicts --
def from_astra(astra_geom):
    """Import a 3D ASTRA volume or projection geometry dict."""
    if not isinstance(astra_geom, dict):
        raise TypeError(
            f"Currently, tomosipo only supports importing ASTRA geometries. "
            f"Objects of type {type(astra_geom)} are not supported. "
            f"Perhaps you meant to use `ts.to_astra'? "
        )
    if "GridSliceCount" in astra_geom:
        return ts.geometry.volume.from_astra(astra_geom)
    return ts.geometry.conversion.from_astra_projection_geometry(astra_geom)


def to_astra(x):
    """Convert a volume or projection geometry to its ASTRA dict."""
    try:
        return x.to_astra()
    except AttributeError:
        raise TypeError(
            f"The object of type {type(x)} does not support conversion to ASTRA."
            f"Perhaps you meant to use `ts.from_astra'? "
        )


# ------------------------------------------------------------------ projector --
def _projection_vectors(astra_pg):
    """(kind, vectors) of an ASTRA projection dict; circular dicts go through geom_2vec."""
    kind = astra_pg["type"]
    if kind in ("cone", "parallel3d"):
        astra_pg = astra_compat.geom_2vec(astra_pg)
        kind = astra_pg["type"]
    if kind == "cone_vec":
        return _backend.KIND_CONE_VEC, astra_pg["Vectors"]
    if kind == "parallel3d_vec":
        return _backend.KIND_PARALLEL_VEC, astra_pg["Vectors"]
    raise ValueError(f"Projection geometry of type '{kind}' cannot be used to create a projector.")


def create_astra_projector(volume_geometry, projection_geometry, *, voxel_supersampling=1,
                           detector_supersampling=1):
    """Create the backend projector for a (volume, projection) geometry pair.

    Same inputs as the reference (``tomosipo/astra.py:80-98``): an axis-aligned
    ``VolumeGeometry`` and any projection geometry; both are first converted to
    their ASTRA dicts, so the C library sees exactly what ASTRA would.
    Returns an opaque handle (:class:`tomosipo_b200._backend.Projector`).
    """
    vg, pg = volume_geometry, projection_geometry
    assert isinstance(vg, ts.geometry.VolumeGeometry)
    avg, apg = vg.to_astra(), pg.to_astra()
    kind, vectors = _projection_vectors(apg)
    opt = avg["option"]
    window = tuple((opt[f"WindowMin{a}"], opt[f"WindowMax{a}"]) for a in "XYZ")
    return _backend.Projector(
        kind,
        (avg["GridSliceCount"], avg["GridRowCount"], avg["GridColCount"]),
        window,
        (apg["DetectorRowCount"], apg["DetectorColCount"]),
        vectors,
        voxel_supersampling=voxel_supersampling,
        detector_supersampling=detector_supersampling,
    )


def direct_project(projector, vol_link, proj_link, forward=None, additive=False):
    """Forward- or back-project between two linked arrays, in place.

    Mirrors ``tomosipo/astra.py:104-153``: ``forward`` must be given; incompatible
    links (different devices) raise ``ValueError``; ``additive`` selects ASTRA's
    MODE_ADD instead of MODE_SET.  CUDA arrays are processed asynchronously on
    their current stream; host arrays synchronously (H2D, kernels, D2H).
    """
    if forward is None:
        raise ValueError("project must be given a forward argument (True/False).")
    if not ts.links.are_compatible(vol_link, proj_link):
        raise ValueError(
            "Cannot perform ASTRA projection on volume and projection data, because they are not compatible. "
            "Usually, this indicates that the data are located on different computing devices. "
        )
    with vol_link.context():
        vol, proj = vol_link.linked_data, proj_link.linked_data
        if tuple(vol.shape) != tuple(projector.vol_shape) or tuple(proj.shape) != tuple(projector.proj_shape):
            raise ValueError(
                f"Projector expects volume {projector.vol_shape} and projections {projector.proj_shape}; "
                f"got {tuple(vol.shape)} and {tuple(proj.shape)}."
            )
        if vol.kind != proj.kind:
            raise ValueError("Cannot project between host and device memory.")
        on_device = vol.kind == "device"
        direction = _backend.FP if forward else _backend.BP
        if on_device:
            # links of different array libraries pass are_compatible() undetermined: the kernel would then read a
            # foreign-device pointer, or race with the producer of the other array on its stream
            if vol.device != proj.device:
                raise ValueError(
                    f"Cannot project between arrays on different GPUs (volume on {vol.device}, projections on {proj.device}).")
            if vol.stream != proj.stream:
                _order_streams(vol.device, first=proj.stream, then=vol.stream)
            projector.project(direction, additive, vol.ptr, proj.ptr, _backend.MEM_DEVICE, device=vol.device,
                              stream=vol.stream)
            if vol.stream != proj.stream:
                _order_streams(vol.device, first=vol.stream, then=proj.stream)
        elif len(_gpu_index) > 1:
            projector.project_multi(direction, additive, vol.ptr, proj.ptr, _gpu_index)
        else:
            projector.project(direction, additive, vol.ptr, proj.ptr, _backend.MEM_HOST,
                              device=_gpu_index[0] if _gpu_index else _default_device(), stream=0)


def _order_streams(device, first, then):
    """Work submitted to CUDA stream ``then`` from now on runs after what ``first`` holds now (raw stream handles)."""
    import torch

    with torch.cuda.device(device):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.ExternalStream(first, device=device) if first else torch.cuda.default_stream(device))
        (torch.cuda.ExternalStream(then, device=device) if then else torch.cuda.default_stream(device)).wait_event(ev)


#: GPUs used for HOST arrays; see :func:`set_gpu_index`
_gpu_index = []


def set_gpu_index(gpus):
    """Choose the GPU(s) that project host (NumPy / CPU tensor) arrays.

    Drop-in for ``astra.set_gpu_index`` as the reference documents it (``doc/topics/operator.rst:226-245``): with a
    list of several GPUs every projection and backprojection of host arrays is divided over all of them
    (``tsp_project_multi``); a single index selects that GPU.  Arrays that live on a GPU are always processed where
    they are.  ``None`` / ``[]`` restores the default (torch's current device, else GPU 0).
    """
    global _gpu_index
    if gpus is None:
        gpus = []
    elif isinstance(gpus, (int, np.integer)):
        gpus = [int(gpus)]
    gpus = [int(g) for g in gpus]
    if len(set(gpus)) != len(gpus) or any(g < 0 for g in gpus):
        raise ValueError(f"Expected a list of distinct, non-negative GPU indices. Got {gpus}")
    _gpu_index = gpus


def _default_device():
    """Device used for host arrays: torch's current device when torch is loaded, else 0."""
    import sys

    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available():
        return torch.cuda.current_device()
    return 0


def direct_fp(projector, vol_data, proj_data, additive=False):
    """``proj (+)= A vol`` on linked arrays (``tomosipo/astra.py:156-184``)."""
    return direct_project(projector, vol_data, proj_data, forward=True, additive=additive)


def direct_bp(projector, vol_data, proj_data, additive=False):
    """``vol (+)= A^T proj`` on linked arrays (``tomosipo/astra.py:187-215``)."""
    return direct_project(projector, vol_data, proj_data, forward=False, additive=additive)


# ------------------------------------------------- legacy Data-based interface --
def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def project(*data, voxel_supersampling=1, detector_supersampling=1, forward=None, additive=False, projector=None):
    """All-to-all projection between ``Data`` volumes and ``Data`` projections.

    Legacy interface of the reference (``tomosipo/astra.py:221-290``, built on
    ``astra.experimental.do_composite``): every volume is projected onto every
    projection dataset (forward), or every projection dataset is back-projected
    into every volume (backward); contributions accumulate.  ``projector``: a
    pre-built projector (``ts.operator(...).astra_projector``) to use instead
    of creating one; as in the reference it serves a single volume / projection
    pair, for which its shapes must match.
    """
    if forward is None:
        raise ValueError("project must be given a forward argument (True/False).")
    vols = [d for d in data if d.is_volume()]
    projs = [d for d in data if d.is_projection()]
    if not vols or not projs:
        raise ValueError("Expected at least one projection dataset and one volume dataset")
    if projector is not None and (len(vols) != 1 or len(projs) != 1):
        raise ValueError("A pre-built projector serves exactly one volume and one projection dataset.")
    targets = projs if forward else vols
    for i, t in enumerate(targets):
        first = not additive
        for s in (vols if forward else projs):
            v, p = (s, t) if forward else (t, s)
            proj_handle = projector
            if proj_handle is None:
                proj_handle = ts.operator(v.geometry, p.geometry, voxel_supersampling=voxel_supersampling,
                                          detector_supersampling=detector_supersampling).astra_projector
            direct_project(proj_handle, v.link, p.link, forward=forward, additive=not first)
            first = False


def forward(*data, voxel_supersampling=1, detector_supersampling=1, additive=False, projector=None):
    """Legacy forward projection of ``Data`` objects (``tomosipo/astra.py:293-330``)."""
    project(*data, voxel_supersampling=voxel_supersampling, detector_supersampling=detector_supersampling,
            forward=True, additive=additive, projector=projector)


def backward(*data, voxel_supersampling=1, detector_supersampling=1, additive=False, projector=None):
    """Legacy backprojection of ``Data`` objects (``tomosipo/astra.py:333-371``)."""
    project(*data, voxel_supersampling=voxel_supersampling, detector_supersampling=detector_supersampling,
            forward=False, additive=additive, projector=projector)


def fdk(vol_data, proj_data, *, voxel_supersampling=1, detector_supersampling=1):
    """FDK reconstruction of ``proj_data`` into ``vol_data`` (``tomosipo/astra.py:374-406``).

    The reference calls ``astra.experimental.accumulate_FDK``, which ADDS the reconstruction to the
    volume dataset's array; so does this (start from zeros for a plain reconstruction).  Cosine
    weighting and the ramp filter run on the GPU (:func:`tomosipo_b200.algorithms.fdk`), followed by
    the library's backprojector.
    """
    from .algorithms import fdk as _fdk

    op = ts.operator(vol_data.geometry, proj_data.geometry, voxel_supersampling=voxel_supersampling,
                     detector_supersampling=detector_supersampling)
    rec = _fdk(op, proj_data.data)
    dst = vol_data.data
    if isinstance(dst, np.ndarray):
        dst += rec if isinstance(rec, np.ndarray) else rec.cpu().numpy()
    else:
        import torch

        dst += torch.as_tensor(rec).to(dst.device)
    return vol_data
