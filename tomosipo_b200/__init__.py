"""tomosipo_b200 -- B200-native 3D tomographic projector behind the tomosipo API.

``import tomosipo_b200 as ts`` gives the top-level names of tomosipo 0.6.0
(reference ``tomosipo/__init__.py:10-41``); ``A = ts.operator(vg, pg)``,
``A(x)``, ``A.T(y)`` run hand-written sm_100a kernels through the C ABI in
``include/tsproj.h`` instead of the ASTRA toolbox.
"""
__version__ = "0.6.0+b200.1"

# Fundamental tolerance of floating-point equality checks (reference __init__.py:41).
epsilon = 1e-8

from . import types  # noqa: E402
from . import utils  # noqa: E402
from . import vector_calc  # noqa: E402
from . import geometry  # noqa: E402
from .geometry.volume import volume  # noqa: E402
from .geometry.volume_vec import volume_vec  # noqa: E402
from .geometry.cone import cone  # noqa: E402
from .geometry.cone_vec import cone_vec  # noqa: E402
from .geometry.parallel_vec import parallel_vec  # noqa: E402
from .geometry.parallel import parallel  # noqa: E402
from .geometry.transform import (  # noqa: E402
    translate,
    scale,
    rotate,
    reflect,
    to_perspective,
    from_perspective,
)
from . import links  # noqa: E402
from .links.base import link  # noqa: E402
from . import astra  # noqa: E402
from .astra import from_astra, to_astra  # noqa: E402
from .Data import data  # noqa: E402
from . import Operator  # noqa: E402
from .Operator import operator  # noqa: E402
from .geometry.concatenate import concatenate  # noqa: E402
from . import phantom  # noqa: E402
from ._backend import cuda_available  # noqa: E402
