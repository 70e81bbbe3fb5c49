"""ctypes binding of libpm3d.so (include/pm3d.h).  No torch, no numpy-side compute: every
compute call goes to the CUDA library, and a missing library or a missing GPU is an error —
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libpm3d.so")

PM3D_OUT_DEVICE = 1
PM3D_IN_DEVICE = 2
PM3D_NO_SYNC = 4
OPT_KEEP_COLUMNS = 1
OPT_FORCE_PAINTER = 2
OPT_PIPELINE = 3
OPT_OVERLAP = 4
OPT_LANES = 5
OPT_DDA_JUMP = 6
PRESENT_COPY, PRESENT_CRT, PRESENT_CRT_CURVE = 0, 1, 2


class Pm3dError(RuntimeError):
    pass


class View(C.Structure):
    """pm3d_view"""
    _fields_ = [
        ("pos_x", C.c_double), ("pos_y", C.c_double), ("cam_z_px", C.c_double),
        ("dir_x", C.c_double), ("dir_y", C.c_double), ("plane_x", C.c_double), ("plane_y", C.c_double),
        ("pitch_px", C.c_int32), ("_reserved", C.c_int32 * 3),
        ("render_distance", C.c_double), ("light_falloff", C.c_double), ("global_illumination", C.c_double),
        ("min_light", C.c_uint8 * 4), ("max_light", C.c_uint8 * 4),
    ]


assert C.sizeof(View) == 104

VIEW_DTYPE = np.dtype(
    [
        ("pos_x", "<f8"), ("pos_y", "<f8"), ("cam_z_px", "<f8"), ("dir_x", "<f8"), ("dir_y", "<f8"),
        ("plane_x", "<f8"), ("plane_y", "<f8"), ("pitch_px", "<i4"), ("_reserved", "<i4", (3,)),
        ("render_distance", "<f8"), ("light_falloff", "<f8"), ("global_illumination", "<f8"),
        ("min_light", "u1", (4,)), ("max_light", "u1", (4,)),
    ]
)
assert VIEW_DTYPE.itemsize == 104

SPRITE_DTYPE = np.dtype(
    [
        ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("scale", "<f8"), ("illumination", "<f8"),
        ("sheet", "<i4"), ("rx0", "<i4"), ("ry0", "<i4"), ("rx1", "<i4"), ("ry1", "<i4"),
        ("anchor", "u1"), ("focusable", "u1"), ("_pad", "u1", (2,)),
    ]
)
assert SPRITE_DTYPE.itemsize == 64

COLUMN_DTYPE = np.dtype(
    [
        ("perp_dist", "<f8"), ("line_height", "<i8"), ("draw_start", "<i8"), ("draw_end", "<i8"),
        ("map_x", "<i4"), ("map_y", "<i4"), ("tex_x", "<i4"), ("tile", "<i2"), ("side", "u1"), ("hit", "u1"),
        ("tint", "u1", (4,)), ("_tail", "u1", (4,)),
    ]
)
assert COLUMN_DTYPE.itemsize == 56

CONV_DTYPE = np.dtype([("distance", "<f8"), ("point_x", "<f8"), ("point_y", "<f8"), ("sprite", "<i4"), ("_pad", "<i4")])
assert CONV_DTYPE.itemsize == 32

MOVER_DTYPE = np.dtype(
    [
        ("pos_x", "<f8"), ("pos_y", "<f8"), ("pos_z", "<f8"), ("move_x", "<f8"), ("move_y", "<f8"), ("move_z", "<f8"),
        ("radius", "<f8"), ("height", "<f8"), ("self", "<i4"), ("parent", "<i4"),
        ("anchor", "u1"), ("destroyed", "u1"), ("_pad", "u1", (6,)),
    ]
)
assert MOVER_DTYPE.itemsize == 80

COLLIDER_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("radius", "<f8"), ("height", "<f8"),
     ("anchor", "u1"), ("destroyed", "u1"), ("_pad", "u1", (6,))]
)
assert COLLIDER_DTYPE.itemsize == 48

MOVE_RESULT_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("collided", "<i4"), ("n_collisions", "<i4"), ("first_entity", "<i4"),
     ("n_points", "<i4"), ("first_x", "<f8"), ("first_y", "<f8"), ("first_z", "<f8"), ("near_x", "<f8"), ("near_y", "<f8")]
)
assert MOVE_RESULT_DTYPE.itemsize == 80

CLUTTER_DEF_DTYPE = np.dtype(
    [("frequency", "<f8"), ("scale", "<f8"), ("sheet", "<i4"), ("rx0", "<i4"), ("ry0", "<i4"), ("rx1", "<i4"),
     ("ry1", "<i4"), ("_pad", "<i4")]
)
assert CLUTTER_DEF_DTYPE.itemsize == 40
CLUTTER_MAX_ANGLES, CLUTTER_MAX_DEFS, CLUTTER_MAX_DISTANCE = 21, 64, 190.0
CLUTTER_QUERY_DTYPE = np.dtype(
    [("cam_x", "<f8"), ("cam_y", "<f8"), ("n_angles", "<i4"), ("_pad", "<i4"), ("cs", "<f8", (CLUTTER_MAX_ANGLES, 2))]
)
assert CLUTTER_QUERY_DTYPE.itemsize == 360
CLUTTER_SPRITE_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("cell_x", "<i4"), ("cell_y", "<i4"), ("def", "<i4"), ("_pad", "<i4")]
)
assert CLUTTER_SPRITE_DTYPE.itemsize == 32

LIST_QUERY_DTYPE = np.dtype([("cam_x", "<f8"), ("cam_y", "<f8"), ("target", "<i4"), ("_pad", "<i4")])
assert LIST_QUERY_DTYPE.itemsize == 24

# every symbol include/pm3d.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "pm3d_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "pm3d_destroy": (None, [_P]),
    "pm3d_last_error": (C.c_char_p, [_P]),
    "pm3d_version": (C.c_int, []),
    "pm3d_device_count": (C.c_int, []),
    "pm3d_sync": (C.c_int, [_P]),
    "pm3d_upload_grid": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pm3d_upload_wall_textures": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "pm3d_upload_atlas": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "pm3d_upload_floor_tiles": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "pm3d_set_box_tiles": (C.c_int, [_P, C.c_int, C.c_int]),
    "pm3d_upload_sprite_sheets": (C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    "pm3d_set_view": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pm3d_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "pm3d_render": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_uint32]),
    "pm3d_render_resident": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "pm3d_download_results": (C.c_int, [_P, _P, _P]),
    "pm3d_download_columns": (C.c_int, [_P, C.c_int, _P, _P]),
    "pm3d_frame_checksums": (C.c_int, [_P, C.c_int, _P, C.c_uint32]),
    "pm3d_last_frame_device": (_P, [_P]),
    "pm3d_upload_segments": (C.c_int, [_P, _P, C.c_int]),
    "pm3d_los_batch": (C.c_int, [_P, _P, C.c_int64, _P, C.c_uint32]),
    "pm3d_move_batch": (C.c_int, [_P, _P, C.c_int64, C.c_double, _P, _P, C.c_uint32]),
    "pm3d_set_present": (C.c_int, [_P, C.c_int, C.c_int, C.c_double]),
    "pm3d_present": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_uint32]),
    "pm3d_upload_colliders": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "pm3d_valid_move_batch": (C.c_int, [_P, _P, C.c_int64, C.c_double, _P, C.c_uint32]),
    "pm3d_upload_clutter": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int64]),
    "pm3d_clutter_batch": (C.c_int, [_P, _P, C.c_int, C.c_double, C.c_int, _P, _P, _P, C.c_uint32]),
    "pm3d_go_rand_cooked": (C.c_int, [_P]),
    "pm3d_raycast_sprites_batch": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_double, _P, _P, C.c_int, _P, _P, C.c_int, _P, _P, _P,
                                             C.POINTER(C.c_int32), C.c_uint32]),
    "pm3d_generate_wall_collision_lines": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, _P, C.c_int]),
    "pm3d_device_alloc": (C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    "pm3d_device_free": (C.c_int, [_P, _P]),
    "pm3d_host_alloc": (C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    "pm3d_host_free": (C.c_int, [_P, _P]),
    "pm3d_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pm3d_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pm3d_timer_start": (C.c_int, [_P, C.c_int]),
    "pm3d_timer_stop": (C.c_int, [_P, C.c_int]),
    "pm3d_timer_elapsed_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "pm3d_ktimer_begin": (C.c_int, [_P, C.c_int]),
    "pm3d_ktimer_collect": (C.c_int, [_P, _P, C.c_int]),
    "pm3d_launch_count": (C.c_int64, [_P]),
    "pm3d_flush_l2": (C.c_int, [_P]),
    "pm3d_comm_unique_id": (C.c_int, [_P]),
    "pm3d_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "pm3d_comm_destroy": (C.c_int, [_P]),
    "pm3d_comm_rank": (C.c_int, [_P]),
    "pm3d_comm_size": (C.c_int, [_P]),
    "pm3d_shard_view_count": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "pm3d_shard_ray_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pm3d_comm_all_gather": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pm3d_comm_render_views": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "pm3d_comm_los_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    "pm3d_group_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "pm3d_group_destroy": (None, [_P]),
    "pm3d_group_last_error": (C.c_char_p, [_P]),
    "pm3d_group_size": (C.c_int, [_P]),
    "pm3d_group_ctx": (_P, [_P, C.c_int]),
    "pm3d_group_sync": (C.c_int, [_P]),
    "pm3d_group_render_views": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "pm3d_group_los_batch": (C.c_int, [_P, _P, C.c_int64, _P]),
    "pm3d_frames_capture_begin": (C.c_int, [_P]),
    "pm3d_frames_capture_end": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "pm3d_frames_replay": (C.c_int, [_P, C.c_int]),
    "pm3d_frames_graph_destroy": (C.c_int, [_P, C.c_int]),
    "pm3d_find_path_batch": (C.c_int, [_P, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]),
    "pm3d_selfcheck_dda": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pm3d_selfcheck_floor": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pm3d_selfcheck_division": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load libpm3d.so; raises Pm3dError (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pm3dError(
            f"{LIB_PATH} is missing: build it with `make -C pixelmek-3d_b200/csrc` "
            "(or __graft_entry__.build()); pm3d has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)
