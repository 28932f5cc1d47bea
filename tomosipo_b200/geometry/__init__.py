"""Geometry model (volumes, projection geometries, transforms); host-side fp64."""
from . import transform
from . import volume_vec
from . import volume
from . import base_projection
from . import det_vec
from . import cone_vec
from . import cone
from . import parallel_vec
from . import parallel
from . import conversion

from .cone import random_cone
from .cone_vec import random_cone_vec
from .parallel_vec import random_parallel_vec
from .parallel import random_parallel
from .volume import random_volume
from .volume_vec import random_volume_vec
from .transform import random_transform

from .base_projection import ProjectionGeometry
from .cone import ConeGeometry
from .cone_vec import ConeVectorGeometry
from .parallel import ParallelGeometry
from .parallel_vec import ParallelVectorGeometry
from .volume import VolumeGeometry
from .volume_vec import VolumeVectorGeometry
from .transform import Transform

from .base_projection import is_projection, is_cone, is_parallel
from .volume import is_volume
