"""Context: one libpm3d ctx (one GPU, one stream).  Thin, typed wrapper over the C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (LIST_QUERY_DTYPE, CLUTTER_DEF_DTYPE, CLUTTER_QUERY_DTYPE, CLUTTER_SPRITE_DTYPE, COLLIDER_DTYPE, COLUMN_DTYPE, CONV_DTYPE, MOVE_RESULT_DTYPE, MOVER_DTYPE, SPRITE_DTYPE, VIEW_DTYPE,
                   Pm3dError, ptr)
from .scene import SceneData


class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.pm3d_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise Pm3dError(f"pm3d_create failed ({rc}): {self.lib.pm3d_last_error(None).decode()}")
        self.h = h
        self.device = device
        self.scene: SceneData | None = None
        self.w = self.h_px = 0
        self._keep = []
        self._no_sprites = np.zeros(0, dtype=SPRITE_DTYPE)
        self._res_ring, self._res_pos = [], 0   # pinned (rects, conv) result buffers for enqueue-only renders

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise Pm3dError(f"{what} failed ({rc}): {self.lib.pm3d_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pm3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.pm3d_sync(self.h), "pm3d_sync")
        self._keep = []

    @property
    def launches(self) -> int:
        return int(self.lib.pm3d_launch_count(self.h))

    # ------------------------------------------------------------------ mission-time uploads
    def upload_scene(self, s: SceneData):
        grid = np.ascontiguousarray(s.grid, dtype=np.uint16)
        ls, mw, mh = grid.shape
        self._check(self.lib.pm3d_upload_grid(self.h, ptr(grid), ls, mw, mh, int(s.num_levels)), "pm3d_upload_grid")
        ti = np.ascontiguousarray(s.wall_tex_index, dtype=np.int32)
        t0 = np.ascontiguousarray(s.wall_tile_side0, dtype=np.int32)
        t1 = np.ascontiguousarray(s.wall_tile_side1, dtype=np.int32)
        self._check(self.lib.pm3d_upload_wall_textures(self.h, ptr(ti), ptr(t0), ptr(t1), len(ti)), "pm3d_upload_wall_textures")
        atlas = np.ascontiguousarray(s.atlas, dtype=np.uint8)
        self._check(self.lib.pm3d_upload_atlas(self.h, ptr(atlas), atlas.shape[0], atlas.shape[1]), "pm3d_upload_atlas")
        ft = np.ascontiguousarray(s.floor_tiles, dtype=np.int16)
        self._check(self.lib.pm3d_upload_floor_tiles(self.h, ptr(ft), ft.shape[0], ft.shape[1]), "pm3d_upload_floor_tiles")
        self._check(self.lib.pm3d_set_box_tiles(self.h, int(s.sky_tile), int(s.floorbox_tile)), "pm3d_set_box_tiles")
        sheets = [np.ascontiguousarray(x, dtype=np.uint8) for x in s.sheets]
        if sheets:
            sizes = np.array([x.size for x in sheets], dtype=np.int64)
            off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
            pix = np.concatenate([x.reshape(-1) for x in sheets])
            sw = np.array([x.shape[1] for x in sheets], dtype=np.int32)
            sh = np.array([x.shape[0] for x in sheets], dtype=np.int32)
            self._check(self.lib.pm3d_upload_sprite_sheets(self.h, ptr(pix), ptr(off), ptr(sw), ptr(sh), len(sheets)),
                        "pm3d_upload_sprite_sheets")
        else:
            self._check(self.lib.pm3d_upload_sprite_sheets(self.h, None, None, None, None, 0), "pm3d_upload_sprite_sheets")
        self.scene = s

    def set_view(self, w: int, h: int, max_views: int = 1, max_sprites_total: int = 2048):
        self._check(self.lib.pm3d_set_view(self.h, w, h, max_views, max_sprites_total), "pm3d_set_view")
        self.w, self.h_px = w, h
        # a device-to-host copy into PAGEABLE memory blocks the caller until the stream reaches it, so
        # PM3D_NO_SYNC renders write their small results (rects, convergence) into a ring of pinned buffers
        self._res_ring, self._res_pos = [], 0
        if max(max_sprites_total, 1) * 16 + max_views * 32 <= (1 << 20):
            for _ in range(8):
                self._res_ring.append((self.host_alloc(max(max_sprites_total, 1) * 16, np.int32, (max(max_sprites_total, 1), 4)),
                                       self.host_alloc(max_views * CONV_DTYPE.itemsize, CONV_DTYPE, (max_views,))))

    def set_option(self, option: int, value: int):
        self._check(self.lib.pm3d_set_option(self.h, option, value), "pm3d_set_option")

    # ------------------------------------------------------------------ frame path
    def render(self, views: np.ndarray, sprites: np.ndarray | None = None, sprite_offset: np.ndarray | None = None,
               want_frame: bool = True, out_device_ptr: int | None = None, no_sync: bool = False,
               out: np.ndarray | None = None):
        """Camera.Update + Draw for len(views) views.  Returns dict(rgba, rects, conv).
        `out` (optional) is a host buffer for the frames, e.g. pinned memory from host_alloc."""
        if not (isinstance(views, np.ndarray) and views.dtype == VIEW_DTYPE and views.flags.c_contiguous):
            views = np.ascontiguousarray(views, dtype=VIEW_DTYPE)
        views = views.reshape(-1)
        nv = len(views)
        if sprites is None:
            sprites = self._no_sprites
        elif not (isinstance(sprites, np.ndarray) and sprites.dtype == SPRITE_DTYPE and sprites.flags.c_contiguous):
            sprites = np.ascontiguousarray(sprites, dtype=SPRITE_DTYPE).reshape(-1)
        if sprite_offset is None:
            if nv != 1:
                raise Pm3dError("sprite_offset is required for more than one view")
            sprite_offset = np.array([0, len(sprites)], dtype=np.int32)
        elif not (isinstance(sprite_offset, np.ndarray) and sprite_offset.dtype == np.int32 and sprite_offset.flags.c_contiguous):
            sprite_offset = np.ascontiguousarray(sprite_offset, dtype=np.int32)
        ns = int(sprite_offset[-1])
        if no_sync and self._res_ring:
            # pinned slot, valid until 8 more renders were enqueued
            rects, conv = self._res_ring[self._res_pos]
            self._res_pos = (self._res_pos + 1) % len(self._res_ring)
            rects, conv = rects[:max(ns, 1)], conv[:nv]
        else:
            rects = np.empty((max(ns, 1), 4), dtype=np.int32)
            conv = np.empty(nv, dtype=CONV_DTYPE)
        flags = 0
        rgba = None
        out_p = None
        if out_device_ptr is not None:
            flags |= _lib.PM3D_OUT_DEVICE
            out_p = C.c_void_p(out_device_ptr)
        elif want_frame:
            rgba = out if out is not None else np.empty((nv, self.h_px, self.w, 4), dtype=np.uint8)
            if rgba.nbytes < nv * self.h_px * self.w * 4:
                raise Pm3dError("out buffer too small")
            out_p = ptr(rgba)
        if no_sync:
            flags |= _lib.PM3D_NO_SYNC
            self._keep.append((views, sprites, sprite_offset, rects, conv, rgba))  # alive until sync()
        self._check(
            self.lib.pm3d_render(self.h, ptr(views), nv, ptr(sprites) if ns else None, ptr(sprite_offset), out_p,
                                 ptr(rects), ptr(conv), flags),
            "pm3d_render",
        )
        return {"rgba": rgba, "rects": rects[:ns], "conv": conv}

    def render_resident(self, d_views: int, n_views: int, d_sprites: int, d_offsets: int, n_sprites_total: int,
                        d_out: int):
        """Enqueue one frame batch whose inputs and output already live in HBM (device pointers)."""
        self._check(
            self.lib.pm3d_render_resident(self.h, C.c_void_p(d_views), n_views, C.c_void_p(d_sprites) if d_sprites else None,
                                          C.c_void_p(d_offsets), n_sprites_total, C.c_void_p(d_out)),
            "pm3d_render_resident",
        )

    def download_results(self, n_sprites: int, n_views: int):
        rects = np.zeros((max(n_sprites, 1), 4), dtype=np.int32)
        conv = np.zeros(n_views, dtype=CONV_DTYPE)
        self._check(self.lib.pm3d_download_results(self.h, ptr(rects), ptr(conv)), "pm3d_download_results")
        return rects[:n_sprites], conv

    def ktimer_begin(self, max_frames: int):
        self._check(self.lib.pm3d_ktimer_begin(self.h, int(max_frames)), "pm3d_ktimer_begin")

    def ktimer_collect(self, cap_frames: int) -> np.ndarray:
        """[frames][3] ms: column cast, view setup, compose."""
        out = np.zeros((cap_frames, 3), dtype=np.float32)
        n = self.lib.pm3d_ktimer_collect(self.h, ptr(out), int(cap_frames))
        if n < 0:
            self._check(n, "pm3d_ktimer_collect")
        return out[:n]

    def download_columns(self, view: int = 0):
        s = self.scene
        cols = np.zeros((int(s.num_levels), self.w), dtype=COLUMN_DTYPE)
        zbuf = np.zeros(self.w, dtype=np.float64)
        self._check(self.lib.pm3d_download_columns(self.h, view, ptr(cols), ptr(zbuf)), "pm3d_download_columns")
        return cols, zbuf

    def frame_checksums(self, n_views: int) -> np.ndarray:
        out = np.zeros(n_views, dtype=np.uint64)
        self._check(self.lib.pm3d_frame_checksums(self.h, n_views, ptr(out), 0), "pm3d_frame_checksums")
        return out

    # ------------------------------------------------------------------ presentation step
    def set_present(self, out_w: int, out_h: int, render_scale: float):
        """Screen size + g.renderScale (game/config.go:239-251); after set_view."""
        self._check(self.lib.pm3d_set_present(self.h, int(out_w), int(out_h), float(render_scale)), "pm3d_set_present")
        self.present_w, self.present_h = int(out_w), int(out_h)

    def present(self, mode: int, n_views: int = 1, d_src: int | None = None, d_out: int | None = None):
        """rayScreen -> overlayScreen (game/scene_game.go:169-185) of the last render (or of d_src)."""
        if d_out is not None:
            self._check(self.lib.pm3d_present(self.h, C.c_void_p(d_src) if d_src else None, n_views, int(mode),
                                              C.c_void_p(d_out), _lib.PM3D_OUT_DEVICE | _lib.PM3D_NO_SYNC), "pm3d_present")
            return None
        out = np.empty((n_views, self.present_h, self.present_w, 4), dtype=np.uint8)
        self._check(self.lib.pm3d_present(self.h, C.c_void_p(d_src) if d_src else None, n_views, int(mode), ptr(out), 0),
                    "pm3d_present")
        return out

    # ------------------------------------------------------------------ LOS / collision path
    def upload_segments(self, segs: np.ndarray):
        segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
        self._check(self.lib.pm3d_upload_segments(self.h, ptr(segs), len(segs)), "pm3d_upload_segments")

    def los_batch(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 4)
        out = np.zeros(len(rays), dtype=np.uint8)
        self._check(self.lib.pm3d_los_batch(self.h, ptr(rays), len(rays), ptr(out), 0), "pm3d_los_batch")
        return out

    # ------------------------------------------------------------------ recorded frame sequences (one CUDA graph)
    def frames_capture_begin(self):
        self._check(self.lib.pm3d_frames_capture_begin(self.h), "pm3d_frames_capture_begin")

    def frames_capture_end(self) -> int:
        g = C.c_int(-1)
        self._check(self.lib.pm3d_frames_capture_end(self.h, C.byref(g)), "pm3d_frames_capture_end")
        return int(g.value)

    def frames_replay(self, graph: int):
        self._check(self.lib.pm3d_frames_replay(self.h, int(graph)), "pm3d_frames_replay")

    def frames_graph_destroy(self, graph: int):
        self._check(self.lib.pm3d_frames_graph_destroy(self.h, int(graph)), "pm3d_frames_graph_destroy")

    def find_path_batch(self, queries: np.ndarray, max_points: int = 4096):
        """Pathing.FindPath for a batch of (start x, y, finish x, y) queries over level 0 of the uploaded grid:
        (points [n][max_points][2], count [n]; -1 = no path)."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 4)
        n = len(q)
        pts = np.zeros((n, max_points, 2), dtype=np.float64)
        cnt = np.zeros(n, dtype=np.int32)
        self._check(self.lib.pm3d_find_path_batch(self.h, ptr(q), n, int(max_points), ptr(pts), ptr(cnt), 0), "pm3d_find_path_batch")
        return pts, cnt

    def move_batch(self, moves: np.ndarray, clip: float):
        moves = np.ascontiguousarray(moves, dtype=np.float64).reshape(-1, 4)
        n = len(moves)
        pos = np.zeros((n, 2), dtype=np.float64)
        col = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.pm3d_move_batch(self.h, ptr(moves), n, float(clip), ptr(pos), ptr(col), 0), "pm3d_move_batch")
        return pos, col

    def upload_colliders(self, colliders: np.ndarray, player_index: int = -1):
        """g.player.Unit + the collidable sprites, in the order the host walks them."""
        colliders = np.ascontiguousarray(colliders, dtype=COLLIDER_DTYPE).reshape(-1)
        self._check(self.lib.pm3d_upload_colliders(self.h, ptr(colliders) if len(colliders) else None, len(colliders),
                                                   int(player_index)), "pm3d_upload_colliders")

    def valid_move_batch(self, movers: np.ndarray, clip: float) -> np.ndarray:
        """Game.getValidMove(entity, x, y, z, checkAlternate=false) for every mover."""
        movers = np.ascontiguousarray(movers, dtype=MOVER_DTYPE).reshape(-1)
        out = np.zeros(len(movers), dtype=MOVE_RESULT_DTYPE)
        self._check(self.lib.pm3d_valid_move_batch(self.h, ptr(movers), len(movers), float(clip), ptr(out), 0),
                    "pm3d_valid_move_batch")
        return out

    # ------------------------------------------------------------------ clutter
    def upload_clutter(self, defs: np.ndarray, cell_class: np.ndarray, class_mask: np.ndarray, map_seed: int):
        """Map.Clutter + the floor-path class of every cell + Map.Seed (after upload_grid)."""
        defs = np.ascontiguousarray(defs, dtype=CLUTTER_DEF_DTYPE).reshape(-1)
        cell_class = np.ascontiguousarray(cell_class, dtype=np.uint8)
        class_mask = np.ascontiguousarray(class_mask, dtype=np.uint64).reshape(-1)
        self._check(self.lib.pm3d_upload_clutter(self.h, ptr(defs) if len(defs) else None, len(defs), ptr(cell_class),
                                                 ptr(class_mask), len(class_mask), int(map_seed)), "pm3d_upload_clutter")

    def clutter_batch(self, queries: np.ndarray, clutter_distance: float, max_per_view: int, want_sprites: bool = False):
        """ClutterHandler.Update for every camera of `queries`: (counts, records[n][max_per_view][, sprites])."""
        queries = np.ascontiguousarray(queries, dtype=CLUTTER_QUERY_DTYPE).reshape(-1)
        n = len(queries)
        counts = np.zeros(n, dtype=np.int32)
        recs = np.zeros((n, max_per_view), dtype=CLUTTER_SPRITE_DTYPE)
        sprites = np.zeros((n, max_per_view), dtype=SPRITE_DTYPE) if want_sprites else None
        self._check(self.lib.pm3d_clutter_batch(self.h, ptr(queries), n, float(clutter_distance), int(max_per_view),
                                                ptr(counts), ptr(recs), ptr(sprites), 0), "pm3d_clutter_batch")
        return (counts, recs, sprites) if want_sprites else (counts, recs)

    def raycast_sprites_batch(self, world: np.ndarray, queries: np.ndarray, render_distance: float, clutter=None,
                              extras=None, max_total: int | None = None):
        """Game.getRaycastSprites for every camera of `queries`.  clutter = (counts[n], sprites[n][stride]) as
        clutter_batch(..., want_sprites=True) returns them; extras = list of per-camera pm3d_sprite arrays.
        Returns (sprites, sprite_offset, index): the concatenated lists pm3d_render takes."""
        world = np.ascontiguousarray(world, dtype=SPRITE_DTYPE).reshape(-1)
        queries = np.ascontiguousarray(queries, dtype=LIST_QUERY_DTYPE).reshape(-1)
        n = len(queries)
        cl_cnt = cl_spr = ex_off = ex_spr = None
        stride = 0
        if clutter is not None:
            cl_cnt = np.ascontiguousarray(clutter[0], dtype=np.int32).reshape(n)
            cl_spr = np.ascontiguousarray(clutter[1], dtype=SPRITE_DTYPE).reshape(n, -1)
            stride = cl_spr.shape[1]
        if extras is not None:
            ex_off = np.concatenate([[0], np.cumsum([len(e) for e in extras])]).astype(np.int32)
            ex_spr = np.ascontiguousarray(np.concatenate(list(extras) + [np.zeros(0, SPRITE_DTYPE)]), dtype=SPRITE_DTYPE)
        if max_total is None:
            max_total = n * len(world) + (int(np.minimum(cl_cnt, stride).clip(0).sum()) if cl_cnt is not None else 0) + \
                (int(ex_off[-1]) if ex_off is not None else 0)
        max_total = max(int(max_total), 1)
        out = np.zeros(max_total, dtype=SPRITE_DTYPE)
        off = np.zeros(n + 1, dtype=np.int32)
        idx = np.zeros(max_total, dtype=np.int32)
        total = C.c_int32(0)
        self._check(self.lib.pm3d_raycast_sprites_batch(self.h, ptr(world) if len(world) else None, len(world), ptr(queries), n,
                                                        float(render_distance), ptr(cl_spr), ptr(cl_cnt), stride, ptr(ex_spr), ptr(ex_off),
                                                        max_total, ptr(out), ptr(off), ptr(idx), C.byref(total), 0),
                    "pm3d_raycast_sprites_batch")
        return out[:total.value], off, idx[:total.value]

    # ------------------------------------------------------------------ device memory + timers
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.pm3d_device_alloc(self.h, int(nbytes), C.byref(p)), "pm3d_device_alloc")
        return int(p.value)

    def device_free(self, p: int):
        self._check(self.lib.pm3d_device_free(self.h, C.c_void_p(p)), "pm3d_device_free")

    def host_alloc(self, nbytes: int, dtype=np.uint8, shape=None) -> np.ndarray:
        """Pinned host buffer as a numpy array (kept alive by this ctx)."""
        p = C.c_void_p()
        self._check(self.lib.pm3d_host_alloc(self.h, int(nbytes), C.byref(p)), "pm3d_host_alloc")
        buf = (C.c_uint8 * int(nbytes)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8).view(dtype)
        return arr.reshape(shape) if shape is not None else arr

    def h2d(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self.lib.pm3d_memcpy_h2d(self.h, C.c_void_p(dst), ptr(src), src.nbytes), "pm3d_memcpy_h2d")

    def d2h(self, dst: np.ndarray, src: int):
        self._check(self.lib.pm3d_memcpy_d2h(self.h, ptr(dst), C.c_void_p(src), dst.nbytes), "pm3d_memcpy_d2h")

    def timer_start(self, slot: int = 0):
        self._check(self.lib.pm3d_timer_start(self.h, slot), "pm3d_timer_start")

    def timer_stop(self, slot: int = 0):
        self._check(self.lib.pm3d_timer_stop(self.h, slot), "pm3d_timer_stop")

    def timer_ms(self, slot: int = 0) -> float:
        ms = C.c_float()
        self._check(self.lib.pm3d_timer_elapsed_ms(self.h, slot, C.byref(ms)), "pm3d_timer_elapsed_ms")
        return float(ms.value)

    def flush_l2(self):
        self._check(self.lib.pm3d_flush_l2(self.h), "pm3d_flush_l2")

    def selfcheck_floor(self, seed: int, blocks: int, iters: int) -> tuple[int, int, int]:
        """(mismatches, operand sets tested, sets the fast path accepted) of the fast floor cast on the device."""
        bad, n, acc = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.pm3d_selfcheck_floor(self.h, int(seed), int(blocks), int(iters), C.byref(bad), C.byref(n),
                                                  C.byref(acc)), "pm3d_selfcheck_floor")
        return int(bad.value), int(n.value), int(acc.value)

    def selfcheck_dda(self, seed: int, blocks: int, steps: int) -> tuple[int, int, int]:
        """(mismatches, states tested, states the certificate accepted) of the closed-form DDA jump on the device."""
        bad, n, acc = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.pm3d_selfcheck_dda(self.h, int(seed), int(blocks), int(steps), C.byref(bad), C.byref(n),
                                                C.byref(acc)), "pm3d_selfcheck_dda")
        return int(bad.value), int(n.value), int(acc.value)

    def selfcheck_division(self, seed: int, blocks: int, iters: int) -> tuple[int, int]:
        """(mismatches, pairs tested) of the division-free floor-cast quotient against `/` on the device."""
        bad, n = C.c_uint64(), C.c_uint64()
        self._check(self.lib.pm3d_selfcheck_division(self.h, int(seed), int(blocks), int(iters), C.byref(bad), C.byref(n)),
                    "pm3d_selfcheck_division")
        return int(bad.value), int(n.value)


def generate_wall_collision_lines(level0: np.ndarray, has_boundary_wall: bool, clip: float) -> np.ndarray:
    """Host mirror of Map.GenerateWallCollisionLines (game/model/map.go:811-851); needs no GPU."""
    lib = _lib.load()
    level0 = np.ascontiguousarray(level0, dtype=np.uint16)
    w, h = level0.shape
    n = lib.pm3d_generate_wall_collision_lines(ptr(level0), w, h, int(bool(has_boundary_wall)), float(clip), None, 0)
    if n < 0:
        raise Pm3dError(f"pm3d_generate_wall_collision_lines failed ({n})")
    out = np.zeros((max(n, 1), 4), dtype=np.float64)
    lib.pm3d_generate_wall_collision_lines(ptr(level0), w, h, int(bool(has_boundary_wall)), float(clip), ptr(out), n)
    return out[:n]


# ---------------------------------------------------------------------- multi-GPU (csrc/comm.cu)
def comm_unique_id() -> bytes:
    """pm3d_comm_unique_id: 128 bytes rank 0 hands to the other ranks."""
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    rc = lib.pm3d_comm_unique_id(buf)
    if rc != 0:
        raise Pm3dError(f"pm3d_comm_unique_id failed ({rc}): {lib.pm3d_last_error(None).decode()}")
    return bytes(buf)


def _comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
    """Collective: every rank's ctx joins the library-owned NCCL communicator."""
    buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
    self._check(self.lib.pm3d_comm_init(self.h, buf, int(rank), int(n_ranks)), "pm3d_comm_init")


def _comm_render_views(self, d_views: int, n_local: int, d_sprites: int, d_offsets: int, n_sprites_total: int, d_out: int,
                       n_views_global: int, d_all_conv: int | None, d_all_sums: int | None):
    self._check(self.lib.pm3d_comm_render_views(self.h, C.c_void_p(d_views), n_local, C.c_void_p(d_sprites) if d_sprites else None,
                                                C.c_void_p(d_offsets), n_sprites_total, C.c_void_p(d_out), n_views_global,
                                                C.c_void_p(d_all_conv) if d_all_conv else None,
                                                C.c_void_p(d_all_sums) if d_all_sums else None), "pm3d_comm_render_views")


def _comm_los_batch(self, d_rays: int, n_local: int, n_global: int, d_all_out: int):
    self._check(self.lib.pm3d_comm_los_batch(self.h, C.c_void_p(d_rays), n_local, n_global, C.c_void_p(d_all_out)), "pm3d_comm_los_batch")


def _comm_all_gather(self, d_send: int, d_recv: int, bytes_per_rank: int):
    self._check(self.lib.pm3d_comm_all_gather(self.h, C.c_void_p(d_send), C.c_void_p(d_recv), int(bytes_per_rank)), "pm3d_comm_all_gather")


Context.comm_init = _comm_init
Context.comm_render_views = _comm_render_views
Context.comm_los_batch = _comm_los_batch
Context.comm_all_gather = _comm_all_gather
Context.comm_size = property(lambda self: int(self.lib.pm3d_comm_size(self.h)))
Context.comm_rank = property(lambda self: int(self.lib.pm3d_comm_rank(self.h)))


class Group:
    """pm3d_group: one process driving n GPUs — one ctx per device plus the communicator over them."""

    def __init__(self, devices):
        self.lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(x) for x in devices])
        g = C.c_void_p()
        rc = self.lib.pm3d_group_create(devs, len(devices), C.byref(g))
        if rc != 0:
            raise Pm3dError(f"pm3d_group_create failed ({rc}): {self.lib.pm3d_group_last_error(None).decode()}")
        self.g = g
        self.ctxs = []
        for i in range(len(devices)):
            c = Context.__new__(Context)
            c.lib, c.h, c.device, c.scene, c.w, c.h_px = self.lib, C.c_void_p(self.lib.pm3d_group_ctx(g, i)), devices[i], None, 0, 0
            c._keep, c._no_sprites, c._res_ring, c._res_pos = [], np.zeros(0, dtype=SPRITE_DTYPE), [], 0
            c.close = lambda: None   # owned by the group
            self.ctxs.append(c)

    def _check(self, rc, what):
        if rc != 0:
            raise Pm3dError(f"{what} failed ({rc}): {self.lib.pm3d_group_last_error(self.g).decode()}")

    def close(self):
        if getattr(self, "g", None):
            for c in self.ctxs:
                c.h = None
            self.lib.pm3d_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def each(self, fn):
        """Replicated set-up: fn(ctx) on every device's ctx (uploads, set_view, options)."""
        for c in self.ctxs:
            fn(c)

    def render_views(self, views: np.ndarray, sprites: np.ndarray | None, sprite_offset: np.ndarray | None):
        views = np.ascontiguousarray(views, dtype=VIEW_DTYPE).reshape(-1)
        n = len(views)
        conv = np.zeros(n, dtype=CONV_DTYPE)
        sums = np.zeros(n, dtype=np.uint64)
        if sprites is not None:
            sprites = np.ascontiguousarray(sprites, dtype=SPRITE_DTYPE).reshape(-1)
            sprite_offset = np.ascontiguousarray(sprite_offset, dtype=np.int32)
        self._check(self.lib.pm3d_group_render_views(self.g, ptr(views), n, ptr(sprites) if sprites is not None and len(sprites) else None,
                                                     ptr(sprite_offset) if sprites is not None else None, ptr(conv), ptr(sums)),
                    "pm3d_group_render_views")
        return conv, sums

    def los_batch(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 4)
        out = np.zeros(len(rays), dtype=np.uint8)
        self._check(self.lib.pm3d_group_los_batch(self.g, ptr(rays), len(rays), ptr(out)), "pm3d_group_los_batch")
        return out
