"""pm3d — host side of the B200-native PixelMek 3D raycasting hot path.

The product is libpm3d.so (pixelmek-3d_b200/csrc: hand-written sm_100a CUDA behind the C-ABI
of include/pm3d.h).  This package is the thin host mirror of the reference's interfaces for
that path (camera, level map, texture handler, sprites, line-of-sight queries) plus the
synthetic workloads.  It never computes the path on the CPU: without the library or a GPU
every compute call raises.
"""
from ._lib import LIB_PATH, Pm3dError, load  # noqa: F401
from .context import Context, Group, comm_unique_id, generate_wall_collision_lines  # noqa: F401
from .scene import SceneData, premultiply  # noqa: F401
