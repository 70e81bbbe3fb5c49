#!/usr/bin/env python
"""bench.py — measurement of the B200-native PixelMek 3D raycasting hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Headline workload (BASELINE.json configs[1], "c2"): single camera 1920x1080 on the bundled
proving_grounds map with floor casting and 200 sprites.  One STEP = one frame = one pass of
the hot path (column cast -> view setup -> compose).  A ring of RING distinct poses/output
framebuffers (> L2) is cycled so successive frames never hit the previous frame's lines.
With N GPUs every rank renders its own K frames of independent cameras (weak scaling; the path
has no exchange step — the only collective is one all_gather of the per-frame checksums).

  value        frames/s, inputs (views, sprites) and outputs resident in HBM, CUDA-event timed
  e2e          frames/s through the public API (pm3d_render) with HOST buffers: H2D of the
               view + sprite snapshot and D2H of the 8.3 MB frame + rects + convergence per step
  roofline     compose kernel: algorithmic bytes (4*W*H per launch) / its mean CUDA-event
               duration inside the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle port (oracle/, all host cores) on a bounded sample of the same frames
  extra        the batched sweeps: c4 (4096 views of 320x200) and c5 (16 Mi LOS rays), sharded

--impl reference times the reference's CPU implementation of the path.  The Go binary cannot
be built here (no Go toolchain, raycaster-go not vendored: DESIGN.md), so this arm runs the
oracle port (kind "port") with all host threads on the same config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "pixelmek-3d_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

W_PX, H_PX, N_SPRITES, RING = 1920, 1080, 200, 24
LANES = 4  # PM3D_OPT_LANES of the resident throughput measurement (independent frames in flight)
FRAME_BYTES = 4 * W_PX * H_PX  # algorithmic bytes of one frame


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reason_bits = [], 0
        self.sm_max = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        names = {
            0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
            0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
        }
        reasons = [n for b, n in names.items() if self.reason_bits & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.samples)}


def view_to_orc(O, v, w, h):
    return O.make_camera(dict(
        pos_x=v["pos_x"], pos_y=v["pos_y"], cam_z_px=v["cam_z_px"], dir_x=v["dir_x"], dir_y=v["dir_y"],
        plane_x=v["plane_x"], plane_y=v["plane_y"], pitch_px=v["pitch_px"], w=w, h=h,
        render_distance=v["render_distance"], light_falloff=v["light_falloff"],
        global_illumination=v["global_illumination"], min_light=v["min_light"], max_light=v["max_light"]))


def cpu_frames(cfg, budget_s: float, threads: int, max_frames: int | None = None, min_frames: int = 3):
    """The oracle port on the same frames: returns (frames, seconds)."""
    from oracle import oracle as O

    scene = O.Scene(cfg["scene"])
    off = cfg["offsets"]
    n = 0
    # one untimed frame: page in, spin up the OpenMP team
    O.render(scene, view_to_orc(O, cfg["views"][0], cfg["w"], cfg["h"]), cfg["sprites"][off[0]:off[1]].astype(O.SPRITE_DTYPE), threads)
    t0 = time.perf_counter()
    while True:
        r = n % len(cfg["views"])
        O.render(scene, view_to_orc(O, cfg["views"][r], cfg["w"], cfg["h"]),
                 cfg["sprites"][off[r]:off[r + 1]].astype(O.SPRITE_DTYPE), threads)
        n += 1
        el = time.perf_counter() - t0
        if max_frames is not None:
            if n >= max_frames:
                break
        elif n >= min_frames and el >= budget_s:
            break
    return n, time.perf_counter() - t0


def base_line(args, world, value, ms_per_step):
    return {
        "metric": "frames/s at 1920x1080 (walls+floor+200 sprites, one camera per GPU)",
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "c2: proving_grounds 500x500x3, 1920x1080, floor cast, 200 sprites, fov 70, render distance 100 cells",
            "rays_per_frame": W_PX * 3, "pixels_per_frame": W_PX * H_PX, "ring": RING,
            "l2": f"ring of {RING} distinct poses and output framebuffers = {RING * FRAME_BYTES / 1e6:.0f} MB > 126 MB L2",
        },
    }


# --------------------------------------------------------------------------------------- reference arm


def run_reference(args, rank, world):
    if rank != 0:
        return
    from pm3d import workloads as Wk

    threads = os.cpu_count() or 1
    cfg = Wk.config_c2(W_PX, H_PX, N_SPRITES, n_poses=RING)
    cpu_frames(cfg, 0.0, threads, max_frames=max(args.warmup, 1))
    n, sec = cpu_frames(cfg, 0.0, threads, max_frames=args.steps)
    value = n / sec
    line = base_line(args, world, value, 1e3 * sec / n)
    line["impl"] = "reference"
    line["n_gpus"] = world
    line["parallelism"] = f"CPU oracle port, {threads} OpenMP threads (rank 0 only)"
    line["cpu_baseline"] = {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                            "sample": f"{n} full 1920x1080 frames of the c2 ring; the Go reference cannot be built here (no Go toolchain), this is the C restatement in oracle/"}
    line["e2e"] = {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    line["gpu_launches"] = 0
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- our arm


def sweep_c4(ctx, rank, world, dist, torch, dev):
    """4096 views of 320x200 sharded i % world through pm3d_comm_render_views: every rep renders this rank's views,
    checksums them and all-gathers (convergence, checksum) of ALL views to every rank — the gather is inside the
    timed region, enqueued by the library behind the kernels."""
    from pm3d import shard, workloads as Wk

    n_views, w, h = 4096, 320, 200
    cfg = Wk.config_c4(n_views, w, h)
    mine = shard.view_shard(n_views, rank, world)
    views = cfg["views"][mine]
    off_all = cfg["offsets"]
    recs = [cfg["sprites"][off_all[i]:off_all[i + 1]] for i in mine]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.int32)
    sprites = np.concatenate(recs)
    nv, ns = len(views), len(sprites)
    ctx.upload_scene(cfg["scene"])
    ctx.set_view(w, h, nv, ns)
    d_views, d_spr, d_off = ctx.device_alloc(views.nbytes), ctx.device_alloc(sprites.nbytes), ctx.device_alloc(offs.nbytes)
    ctx.h2d(d_views, views); ctx.h2d(d_spr, sprites); ctx.h2d(d_off, offs)
    fb = nv * w * h * 4
    d_out = ctx.device_alloc(fb)
    d_conv, d_sums = ctx.device_alloc(n_views * 32), ctx.device_alloc(n_views * 8)
    call = lambda: ctx.comm_render_views(d_views, nv, d_spr, d_off, ns, d_out, n_views, d_conv, d_sums)
    for _ in range(3):
        call()
    ctx.sync()
    reps = 10   # (a sweep is 0.3 - 1.3 ms: ten of them dilute the ranks' skew at the start of the timed region)
    if world > 1:
        dist.barrier()
    ctx.flush_l2()
    ctx.ktimer_begin(reps)
    ctx.timer_start(2)
    for _ in range(reps):
        call()
    ctx.timer_stop(2)
    ms = ctx.timer_ms(2) / reps
    kt = ctx.ktimer_collect(reps)
    sums = np.zeros(n_views, dtype=np.uint64)   # every rank holds every view's checksum
    ctx.sync(); ctx.d2h(sums, d_sums)
    total = int(sums.view(np.int64).sum()) & 0xFFFFFFFFFFFFFFFF
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([total & 0x7FFFFFFFFFFFFFFF], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lo_, hi_ = tot.clone(), tot.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        if int(lo_.item()) != int(hi_.item()):
            raise RuntimeError("c4: the ranks hold different gathered checksums")
    for p in (d_views, d_spr, d_off, d_out, d_conv, d_sums):
        ctx.device_free(p)
    ms_max = float(t.item())
    peak, _ = hbm_peak()
    compose_ms = float(kt[:, 2].mean())
    return {
        "workload": "c4: 4096 views 320x200, 50 nearest sprites each, sharded i % n_gpus; per rep render + checksum + "
                    "ncclAllGather of (convergence, checksum) of all views to every rank (pm3d_comm_render_views), gather timed",
        "views_per_s": n_views / (ms_max * 1e-3), "ms_per_sweep": ms_max, "views_per_gpu": nv,
        "compose_ms": compose_ms, "front_ms": float(kt[:, 0].mean()),
        "hbm_frac_compose": (fb / (compose_ms * 1e-3)) / 1e9 / peak,
        "hbm_frac_sweep": (fb / (ms * 1e-3)) / 1e9 / peak,
        "checksum_of_checksums": f"{total:016x}",
    }


def sweep_c3(ctx):
    """c3: single camera 3840x2160, synthetic 512x512 map (5 % walls), 3 and 8 raycast levels, infinite render
    distance, pitch swept over [-pi/8, pi/4] (8 poses, ring of 8 output frames = 265 MB > L2).  A single camera does
    not shard: every rank would render the same frames, so this is rank 0's GPU (replicas only)."""
    from pm3d import workloads as Wk

    w, h, ring = 3840, 2160, 8
    fb = w * h * 4
    peak, _ = hbm_peak()
    out = {"workload": "c3: 3840x2160, synthetic 512x512 map 5% walls + boundary, renderDistance -1, pitch swept over "
                       "[-pi/8, pi/4] (8 poses, 8 output frames = 265 MB > L2), no sprites; one GPU (replicas only)"}
    d_out = ctx.device_alloc(ring * fb)
    for levels in (3, 8):
        cfg = Wk.config_c3(levels=levels, w=w, h=h)
        x, y, z, heading, _ = cfg["poses"][0]
        poses = [(x, y, z, heading + 0.37 * i, -math.pi / 8 + (math.pi / 4 + math.pi / 8) * i / (ring - 1)) for i in range(ring)]
        views = Wk.make_views(cfg["scene"], w, h, poses, render_distance=-1.0)
        ctx.upload_scene(cfg["scene"])
        ctx.set_view(w, h, 1, 1)
        d_views = ctx.device_alloc(views.nbytes)
        d_off = ctx.device_alloc(8)
        ctx.h2d(d_views, views); ctx.h2d(d_off, np.zeros(2, dtype=np.int32))
        frame = lambda i: ctx.render_resident(d_views + (i % ring) * views.itemsize, 1, 0, d_off, 0, d_out + (i % ring) * fb)
        reps = 64

        def timed(lanes):
            ctx.set_option(__import__("pm3d")._lib.OPT_LANES, lanes)
            for i in range(ring):
                frame(i)
            ctx.sync()
            ctx.timer_start(7)
            for i in range(reps):
                frame(i)
            ctx.timer_stop(7)
            return ctx.timer_ms(7) / reps

        ms = timed(1)
        ms_lanes = timed(LANES)
        ctx.set_option(__import__("pm3d")._lib.OPT_LANES, 1)
        ctx.set_option(__import__("pm3d")._lib.OPT_OVERLAP, 0)   # per-kernel durations: one kernel at a time
        ctx.ktimer_begin(ring)
        for i in range(ring):
            frame(i)
        kt = ctx.ktimer_collect(ring)
        ctx.set_option(__import__("pm3d")._lib.OPT_OVERLAP, 1)
        ctx.device_free(d_views); ctx.device_free(d_off)
        compose_ms = float(kt[:, 2].mean())
        out[f"levels_{levels}"] = {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "mrays_per_s": w * levels / (ms * 1e-3) / 1e6,
                                   "mpixels_per_s": w * h / (ms * 1e-3) / 1e6, "compose_ms": compose_ms,
                                   "front_ms": float(kt[:, 0].mean()), "hbm_frac_compose": fb / (compose_ms * 1e-3) / 1e9 / peak,
                                   "hbm_frac_frame": fb / (ms * 1e-3) / 1e9 / peak,
                                   f"frames_per_s_{LANES}_lanes": 1e3 / ms_lanes,
                                   f"hbm_frac_frame_{LANES}_lanes": fb / (ms_lanes * 1e-3) / 1e9 / peak}
    ctx.device_free(d_out)
    return out


def sweep_c5(ctx, rank, world, dist, torch, dev):
    """16 Mi LOS rays on a 256x256 map, contiguous shards, through pm3d_comm_los_batch: every rep answers this rank's
    range and all-gathers the 1 B/ray answers of ALL rays to every rank inside the timed region."""
    from pm3d import shard, workloads as Wk

    n_rays = 1 << 24
    lo, hi = shard.ray_shard(n_rays, rank, world)
    grid = Wk.synthetic_levels(256, 256, 1, 0.10, 42)
    from pm3d import generate_wall_collision_lines

    segs = generate_wall_collision_lines(grid[0], True, Wk.CLIP_DISTANCE)
    rays = Wk.los_rays(grid[0], n_rays, 43)[lo:hi]
    n = len(rays)
    ctx.upload_segments(segs)
    d_rays, d_out = ctx.device_alloc(rays.nbytes), ctx.device_alloc(n_rays)
    ctx.h2d(d_rays, rays)
    call = lambda: ctx.comm_los_batch(d_rays, n, n_rays, d_out)
    for _ in range(2):
        call()
    ctx.sync()
    if world > 1:
        dist.barrier()
    reps = 10   # (a batch is 0.2 - 0.75 ms)
    ctx.timer_start(3)
    for _ in range(reps):
        call()
    ctx.timer_stop(3)
    ms = ctx.timer_ms(3) / reps
    out = np.zeros(n_rays, dtype=np.uint8)   # every rank holds every ray's answer
    ctx.d2h(out, d_out)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    visible = torch.tensor([int(out.sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vlo, vhi = visible.clone(), visible.clone()
        dist.all_reduce(vlo, op=dist.ReduceOp.MIN); dist.all_reduce(vhi, op=dist.ReduceOp.MAX)
        if int(vlo.item()) != int(vhi.item()):
            raise RuntimeError("c5: the ranks hold different gathered answers")
    ctx.device_free(d_rays); ctx.device_free(d_out)
    ms_max = float(t.item())
    peak, _ = hbm_peak()
    return {
        "workload": f"c5: 16777216 LOS rays, 256x256 map 10% walls, {len(segs)} collision segments, contiguous shards; per rep "
                    "k_los_fast + k_los + all-gather of the 16 MiB of answers to every rank (pm3d_comm_los_batch), gather timed",
        "mrays_per_s": n_rays / (ms_max * 1e-3) / 1e6, "ms_per_batch": ms_max, "rays_per_gpu": n,
        "hbm_frac": (33.0 * n / (ms * 1e-3)) / 1e9 / peak, "visible_rays": int(visible.item()),
    }


def sweep_astar(ctx, rank, world, dist, torch, dev, with_cpu):
    """Batched Pathing.FindPath (game/model/pathing.go:123-186): 4096 AI units of proving_grounds each asking for a
    path to a target within 60 cells, contiguous shards, results resident; returns dict for 'extra'."""
    import ctypes as C

    from pm3d import shard, workloads as Wk

    n_all, cap = 4096, 1024
    g = Wk.load_golden("proving_grounds")
    lvl = g["levels"][0]
    scene, _ = Wk.proving_grounds_scene()
    ctx.upload_scene(scene)
    q_all = Wk.path_queries(lvl, n_all, 5, max_manhattan=60)
    lo, hi = shard.ray_shard(n_all, rank, world)
    q = np.ascontiguousarray(q_all[lo:hi])
    n = len(q)
    d_q, d_p, d_c = ctx.device_alloc(max(q.nbytes, 32)), ctx.device_alloc(max(n, 1) * cap * 16), ctx.device_alloc(max(n, 1) * 4)
    ctx.h2d(d_q, q)
    call = lambda: ctx._check(ctx.lib.pm3d_find_path_batch(ctx.h, C.c_void_p(d_q), n, cap, C.c_void_p(d_p), C.c_void_p(d_c), 7), "pm3d_find_path_batch")
    call()
    ctx.sync()
    if world > 1:
        dist.barrier()
    reps = 3
    ctx.timer_start(5)
    for _ in range(reps):
        call()
    ctx.timer_stop(5)
    ms = ctx.timer_ms(5) / reps
    cnt = np.zeros(n, dtype=np.int32)
    ctx.d2h(cnt, d_c)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([int((cnt > 0).sum()), int(cnt[cnt > 0].sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    for p_ in (d_q, d_p, d_c):
        ctx.device_free(p_)
    ms_max = float(t.item())
    out = {
        "workload": f"astar: {n_all} FindPath queries on proving_grounds (500x500, targets within 60 cells), contiguous shards; "
                    "go-astar's exact queue order + Bezier smoothing, one search per thread",
        "paths_per_s": n_all / (ms_max * 1e-3), "ms_per_batch": ms_max, "queries_per_gpu": n,
        "paths_found": int(tot[0].item()), "curve_points": int(tot[1].item()),
    }
    if with_cpu:
        from oracle import oracle as O

        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        O.find_path_batch(lvl, q_all[:512], cap=cap, threads=threads)
        out["cpu_port_paths_per_s"] = 512 / (time.perf_counter() - t0)
        out["cpu_port_threads"] = threads
    return out


def sweep_moves(ctx, rank, world, dist, torch, dev):
    """Full getValidMove batch (game/collision.go:69-355): 4 Mi movers against the c5 map's outlines and 512
    colliders, contiguous shards; returns dict for 'extra'."""
    import ctypes as C

    from pm3d import generate_wall_collision_lines, shard, workloads as Wk

    n_all = 1 << 22
    grid = Wk.synthetic_levels(256, 256, 1, 0.10, 42)
    segs = generate_wall_collision_lines(grid[0], True, Wk.CLIP_DISTANCE)
    colliders, movers = Wk.move_workload(grid[0], n_all, 512, 77)
    lo, hi = shard.ray_shard(n_all, rank, world)
    movers = movers[lo:hi]
    n = len(movers)
    ctx.upload_scene(Wk.scene_from_grid(grid, 1))
    ctx.upload_segments(segs)
    ctx.upload_colliders(colliders, 0)
    d_in, d_out = ctx.device_alloc(movers.nbytes), ctx.device_alloc(n * 80)
    ctx.h2d(d_in, movers)
    flags = 1 | 2 | 4
    call = lambda: ctx._check(ctx.lib.pm3d_valid_move_batch(ctx.h, C.c_void_p(d_in), n, Wk.CLIP_DISTANCE, C.c_void_p(d_out), flags), "pm3d_valid_move_batch")
    call(); ctx.sync()
    if world > 1:
        dist.barrier()
    reps = 3
    ctx.timer_start(4)
    for _ in range(reps):
        call()
    ctx.timer_stop(4)
    ms = ctx.timer_ms(4) / reps
    out = np.zeros(n, dtype=__import__("pm3d")._lib.MOVE_RESULT_DTYPE)
    ctx.d2h(out, d_out)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    cnt = torch.tensor([int(out["collided"].sum()), int((out["n_collisions"] > 0).sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ctx.device_free(d_in); ctx.device_free(d_out)
    ms_max = float(t.item())
    peak, _ = hbm_peak()
    return {
        "workload": f"moves: {n_all} getValidMove(checkAlternate=false) calls, 256x256 map 10% walls, {len(segs)} segments, 512 colliders",
        "mmoves_per_s": n_all / (ms_max * 1e-3) / 1e6, "ms_per_batch": ms_max, "movers_per_gpu": n,
        "hbm_frac": (160.0 * n / (ms * 1e-3)) / 1e9 / peak, "collided": int(cnt[0].item()), "entity_hits": int(cnt[1].item()),
    }


def sweep_clutter(ctx, rank, world, dist, torch, dev, with_cpu):
    """ClutterHandler.Update (game/clutter.go:25-117) for the 4096 c4 cameras on the bundled map with its own clutter
    list and seed, fov 70, clutterDistance 25 cells (the reference default), views sharded i % n_gpus."""
    import ctypes as C
    import time

    from pm3d import clutter as cl, shard, workloads as Wk
    L = __import__("pm3d")._lib

    scene, g = Wk.proving_grounds_scene()
    n_all = 4096
    poses = Wk.free_cell_poses(scene, n_all, 4096)
    clutter = [cl.MapClutter(str(im), float(f), float(h), str(m) or None) for im, m, f, h in
               zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]
    sheets = {c.image: (3, int(w), int(h)) for c, (w, h) in zip(clutter, g["clutter_image_size"])}
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(x) for x in g["floor_images"]], sheets)
    mine = shard.view_shard(n_all, rank, world)
    q = cl.clutter_queries([(poses[i][0], poses[i][1], poses[i][3]) for i in mine], 70.0)
    n, cap, cdist = len(q), 21 * 26 * len(clutter), 25.0
    ctx.upload_scene(scene)
    ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
    d_q, d_cnt = ctx.device_alloc(q.nbytes), ctx.device_alloc(4 * n)
    d_rec, d_spr = ctx.device_alloc(32 * n * cap), ctx.device_alloc(64 * n * cap)
    ctx.h2d(d_q, q)
    flags = L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC
    call = lambda: ctx._check(ctx.lib.pm3d_clutter_batch(ctx.h, C.c_void_p(d_q), n, cdist, cap, C.c_void_p(d_cnt), C.c_void_p(d_rec),
                                                         C.c_void_p(d_spr), flags), "pm3d_clutter_batch")
    call(); ctx.sync()
    if world > 1:
        dist.barrier()
    reps = 20
    ctx.timer_start(6)
    for _ in range(reps):
        call()
    ctx.timer_stop(6)
    ms = ctx.timer_ms(6) / reps
    counts = np.zeros(n, dtype=np.int32)
    ctx.d2h(counts, d_cnt)
    for p_ in (d_q, d_cnt, d_rec, d_spr):
        ctx.device_free(p_)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([int(counts.sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    out = {"workload": f"clutter: ClutterHandler.Update for {n_all} cameras on proving_grounds (2 clutter types, seed {int(g['map_seed'])}), "
                       "fov 70, clutterDistance 25 cells = 546 samples per camera, sharded i % n_gpus",
           "updates_per_s": n_all / (ms_max * 1e-3), "ms_per_batch": ms_max, "cameras_per_gpu": n, "clutter_sprites": int(tot.item()),
           "bytes_per_camera": 360 + 4 + 96 * float(tot.item()) / n_all,
           "hbm_frac": (360 + 4 + 96 * float(tot.item()) / n_all) * n / (ms * 1e-3) / 1e9 / hbm_peak()[0]}
    if with_cpu:
        from oracle import oracle as O
        k, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 2.0 and k < n:
            O.clutter_update(scene.grid[0], cls, masks, defs.astype(O.CLUTTER_DEF_DTYPE), int(g["map_seed"]), cdist,
                             q[k:k + 1].astype(O.CLUTTER_QUERY_DTYPE))
            k += 1
        out["cpu_port_updates_per_s_1_thread"] = k / (time.perf_counter() - t0)
    return out


def sweep_tick(ctx, rank, world, dist, torch, dev):
    """One headless "tick" for 4096 unit cameras with every intermediate in HBM: ClutterHandler.Update
    (pm3d_clutter_batch) -> getRaycastSprites (pm3d_raycast_sprites_batch: the map's 408 sprites within the render
    distance + each camera's clutter) -> Camera.Update/Draw at 320x200 (pm3d_render_resident); views sharded i % n."""
    import ctypes as C

    from pm3d import clutter as cl, shard, sprites as sp, workloads as Wk
    L = __import__("pm3d")._lib

    scene, g = Wk.proving_grounds_scene()
    scene.sheets = list(scene.sheets) + [Wk.sprite_sheet(18, 18, 41), Wk.sprite_sheet(24, 24, 42)]
    n_all, w, h, rd, cdist = 4096, 320, 200, 100.0, 25.0
    poses = Wk.free_cell_poses(scene, n_all, 4096)
    clutter = [cl.MapClutter(str(im), float(f), float(hh), str(m) or None) for im, m, f, hh in
               zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]
    sheets = {clutter[0].image: (len(scene.sheets) - 2, 18, 18), clutter[1].image: (len(scene.sheets) - 1, 24, 24)}
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(x) for x in g["floor_images"]], sheets)
    world_sprites = Wk.golden_sprites(scene, g)
    mine = shard.view_shard(n_all, rank, world)
    my_poses = [poses[i] for i in mine]
    views = Wk.make_views(scene, w, h, my_poses, render_distance=rd)
    cq = cl.clutter_queries([(p[0], p[1], p[3]) for p in my_poses], 70.0)
    lq = sp.list_queries([(p[0], p[1]) for p in my_poses])
    n, stride = len(views), 21 * 26 * len(clutter)
    max_total = n * (len(world_sprites) + stride)
    ctx.upload_scene(scene)
    ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
    ctx.set_view(w, h, n, max_total)
    sizes = dict(views=views.nbytes, cq=cq.nbytes, lq=lq.nbytes, world=world_sprites.nbytes, cnt=4 * n, rec=32 * n * stride,
                 cspr=64 * n * stride, spr=64 * max_total, off=4 * (n + 1), out=n * w * h * 4, conv=32 * n_all, sums=8 * n_all)
    d = {k: ctx.device_alloc(v) for k, v in sizes.items()}
    ctx.h2d(d["views"], views); ctx.h2d(d["cq"], cq); ctx.h2d(d["lq"], lq); ctx.h2d(d["world"], world_sprites)
    flags = L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC
    total = [0]

    def tick(first=False):
        ctx._check(ctx.lib.pm3d_clutter_batch(ctx.h, d["cq"], n, cdist, stride, d["cnt"], d["rec"], d["cspr"], flags), "clutter")
        ctx._check(ctx.lib.pm3d_raycast_sprites_batch(ctx.h, d["world"], len(world_sprites), d["lq"], n, rd, d["cspr"], d["cnt"], stride,
                                                      None, None, max_total, d["spr"], d["off"], None, None, flags), "list")
        if first:   # the list total is a pure function of the poses: read it once, the render call is told the same every tick
            off = np.zeros(n + 1, dtype=np.int32)
            ctx.sync(); ctx.d2h(off, d["off"])
            total[0] = int(off[-1])
        ctx.comm_render_views(d["views"], n, d["spr"], d["off"], total[0], d["out"], n_all, d["conv"], d["sums"])

    tick(first=True); tick(); ctx.sync()
    if world > 1:
        dist.barrier()
    reps = 5
    ctx.timer_start(8)
    for _ in range(reps):
        tick()
    ctx.timer_stop(8)
    ms = ctx.timer_ms(8) / reps
    sums = np.zeros(n_all, dtype=np.uint64)
    ctx.sync(); ctx.d2h(sums, d["sums"])       # gathered by the library: every rank holds all of them
    for p_ in d.values():
        ctx.device_free(p_)
    ctx.set_view(w, h, 1, 1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([total[0]], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    csum = int(sums.view(np.int64).sum()) & 0xFFFFFFFFFFFFFFFF
    ms_max = float(t.item())
    return {"workload": f"tick: {n_all} cameras at {w}x{h} on proving_grounds, per tick clutter update -> sprite list (408 map sprites within "
                        f"{rd:.0f} cells + clutter) -> frame -> all-gather of every view's (convergence, checksum), all resident in HBM, sharded i % n_gpus",
            "ticks_per_s": 1e3 / ms_max, "views_per_s": n_all / (ms_max * 1e-3), "ms_per_tick": ms_max,
            "sprites_listed": int(tot.item()), "checksum_of_checksums": f"{csum:016x}"}


def sweep_present(ctx, torch):
    """Presentation step at 1920x1080 from a 960x540 rayScreen (renderScale 0.5), CRT shader; rank 0's GPU."""
    from pm3d import workloads as Wk

    cfg = Wk.config_c2(960, 540, 60, n_poses=1)
    ctx.upload_scene(cfg["scene"])
    ctx.set_view(960, 540, 1, 60)
    ctx.render(cfg["views"], cfg["sprites"], cfg["offsets"], want_frame=False)
    ctx.set_present(1920, 1080, 0.5)
    ring = 24
    d_out = ctx.device_alloc(ring * FRAME_BYTES)
    out = {}
    for name, mode in (("copy", 0), ("crt", 1), ("crt_curve", 2)):
        for i in range(5):
            ctx.present(mode, 1, d_out=d_out + (i % ring) * FRAME_BYTES)
        ctx.sync()
        reps = 200
        ctx.timer_start(5)
        for i in range(reps):
            ctx.present(mode, 1, d_out=d_out + (i % ring) * FRAME_BYTES)
        ctx.timer_stop(5)
        ms = ctx.timer_ms(5) / reps
        peak, _ = hbm_peak()
        out[name] = {"us_per_frame": ms * 1e3, "hbm_frac": FRAME_BYTES / (ms * 1e-3) / 1e9 / peak}
    ctx.device_free(d_out)
    out["workload"] = "present: 960x540 rayScreen -> 1920x1080 screen (renderScale 0.5), ring of 24 outputs > L2; includes launch gaps"
    return out


def run_ours(args, rank, world, local_rank):
    import torch

    import pm3d
    from pm3d import workloads as Wk

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = pm3d.Context(local_rank)
    # the library's own NCCL communicator (csrc/comm.cu) gathers the compact results of the sharded sweeps; torch only
    # carries the 128-byte id from rank 0 to the others (and reduces the timings)
    uid = [pm3d.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)

    # ---- workload: every rank its own ring of cameras
    # (the same ring on every rank: weak scaling measures the box, not which rank drew the heaviest camera poses —
    #  with the reference's translucent trees a frame is 22-80 us of compose depending on the pose)
    cfg = Wk.config_c2(W_PX, H_PX, N_SPRITES, n_poses=RING, seed=12345)
    ctx.upload_scene(cfg["scene"])
    ctx.set_view(W_PX, H_PX, 1, N_SPRITES)
    views, sprites, off = cfg["views"], cfg["sprites"], cfg["offsets"]
    slot_off = np.array([0, N_SPRITES], dtype=np.int32)
    d_views = ctx.device_alloc(views.nbytes)
    d_spr = ctx.device_alloc(sprites.nbytes)
    d_off = ctx.device_alloc(slot_off.nbytes)
    d_out = ctx.device_alloc(RING * FRAME_BYTES)
    ctx.h2d(d_views, views); ctx.h2d(d_spr, sprites); ctx.h2d(d_off, slot_off)
    vsz, ssz = views.dtype.itemsize, sprites.dtype.itemsize

    def frame(i):
        r = i % RING
        ctx.render_resident(d_views + r * vsz, 1, d_spr + int(off[r]) * ssz, d_off, N_SPRITES, d_out + r * FRAME_BYTES)

    for i in range(max(args.warmup, 3)):
        frame(i)
    ctx.sync()

    # ---- timed region (device resident).  Measured twice: one lane (frames strictly one after the other: front of
    #      frame i+1 beside compose of frame i), then PM3D_OPT_LANES = LANES, where the library deals the independent
    #      frames of the ring round-robin to LANES lanes (own streams + intermediate buffers) so that several frames
    #      are in flight; `value` is the latter, the former is reported as `single_lane`.
    def timed_resident(lanes, use_graph=False):
        """K frames of the ring.  use_graph: the RING frames are recorded once (pm3d_frames_capture_*: one CUDA graph
        over all lanes and streams) and replayed, one launch per RING frames, the K % RING frames left over are plain
        calls; else every frame is enqueued with its own calls (the headline: the driver's short runs have K < RING)."""
        ctx.set_option(pm3d._lib.OPT_LANES, lanes)
        graph = None
        if use_graph:
            ctx.frames_capture_begin()
            for i in range(RING):
                frame(i)
            graph = ctx.frames_capture_end()

        def run(n):
            full, rest = (n // RING, n % RING) if graph is not None else (0, n)
            for _ in range(full):
                ctx.frames_replay(graph)
            for i in range(rest):
                frame(i)

        run(max(args.warmup, 3, RING))
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        ctx.timer_start(0)
        run(args.steps)
        ctx.timer_stop(0)
        ms = ctx.timer_ms(0)
        ctx.sync()
        torch.cuda.synchronize()
        launches = ctx.launches - l0
        if graph is not None:
            ctx.frames_graph_destroy(graph)
        return ms, launches

    single_ms, _ = timed_resident(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    elapsed_ms, launches = timed_resident(LANES)
    sampler.stop()
    graph_ms = timed_resident(LANES, use_graph=True)[0] if args.steps >= RING else None
    ctx.set_option(pm3d._lib.OPT_LANES, 1)
    # per-kernel durations: the same frames again with an event pair around each kernel (the event
    # records cost ~1 us each, so they are kept out of the region `value` is computed from)
    kt_frames = min(args.steps, 1024)
    ctx.set_option(pm3d._lib.OPT_OVERLAP, 0)  # kernels one after the other, as ncu sees them: durations without a co-running launch
    ctx.ktimer_begin(kt_frames)
    for i in range(kt_frames):
        frame(i)
    kt = ctx.ktimer_collect(kt_frames)
    sums = ctx.frame_checksums(1)  # last frame: compact result
    ctx.set_option(pm3d._lib.OPT_OVERLAP, 1)
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allsum = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allsum, torch.from_numpy(sums.view(np.int64)).to(dev))
    elapsed_ms = float(t.item())
    value = world * args.steps / (elapsed_ms * 1e-3)
    t1 = torch.tensor([single_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
    single_ms = float(t1.item())

    # ---- e2e through the public API (pm3d_render) with HOST buffers: every step copies the view +
    #      sprite snapshot host->device and the frame + rects + convergence device->host.
    #      (a) one stream sync per frame (what a game loop does); (b) PM3D_OPT_PIPELINE=3: same copies,
    #      the D2H of frame i overlaps the kernels of frame i+1 (throughput of headless batches).
    # (its own step count, stated in the field: a 3-deep copy pipeline needs more than the driver's 20 steps to show its rate)
    e2e_steps = max(3, min(max(args.steps, 200), 300))
    host_ring = [ctx.host_alloc(FRAME_BYTES, np.uint8, (1, H_PX, W_PX, 4)) for _ in range(3)]

    def e2e_loop(n, piped):
        for i in range(n):
            r = i % RING
            ctx.render(views[r:r + 1], sprites[off[r]:off[r + 1]], out=host_ring[i % 3], no_sync=piped)
        ctx.sync()

    def timed(piped):
        e2e_loop(3, piped)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, piped)
        dt = time.perf_counter() - t0
        if os.environ.get("PM3D_BENCH_DEBUG"):
            print(f"[rank {rank}] e2e piped={piped}: {e2e_steps / dt:.0f} frames/s on this rank", file=sys.stderr, flush=True)
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * e2e_steps / float(te.item())

    e2e_sync = timed(False)
    ctx.set_option(pm3d._lib.OPT_PIPELINE, 3)
    e2e_piped = timed(True)
    ctx.set_option(pm3d._lib.OPT_PIPELINE, 1)
    e2e_value = max(e2e_sync, e2e_piped)

    # ---- roofline of the dominant kernel (compose)
    peak, peak_kind = hbm_peak()
    compose_ms = float(kt[:, 2].mean()) if len(kt) else float("nan")
    achieved = FRAME_BYTES / (compose_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("k_compose_c2_dram_bytes_per_launch")
        except Exception:
            traffic = None

    line = base_line(args, world, value, elapsed_ms / args.steps)
    line["lanes"] = LANES
    if graph_ms is not None:
        tg = torch.tensor([graph_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        line["graph_replay"] = {"value": world * args.steps / (float(tg.item()) * 1e-3), "unit": "frames/s",
                                "note": f"same frames, the ring recorded once into one CUDA graph (pm3d_frames_capture_begin/_end) and replayed: one launch per {RING} "
                                        "frames instead of ~8 runtime calls per frame; a replay is one item of the ctx stream, so the ring's last frames "
                                        "drain before the next replay starts (a bubble per ring) - it pays when the HOST is the limit (many ranks per box), not on one GPU"}
    line["parallelism"] = (f"{world} ring(s) of the same {RING} camera poses, one ring per GPU; no data-path collective; frames are "
                                     f"enqueued back to back and dealt round-robin to {LANES} lanes of one ctx (PM3D_OPT_LANES: own streams "
                                     "and intermediate buffers per lane, so the cast/setup and compose launches of several frames overlap)")
    line["single_lane"] = {"value": world * args.steps / (single_ms * 1e-3), "unit": "frames/s", "ms_per_step": single_ms / args.steps,
                           "note": "PM3D_OPT_LANES = 1: frames strictly in order, only the cast/setup launch of frame i+1 overlaps the "
                                   "compose launch of frame i"}
    line["clocks"] = sampler.summary()
    line["e2e"] = {"value": e2e_value, "unit": "frames/s",
                   "h2d_bytes_per_step": int(vsz + N_SPRITES * ssz + 8),
                   "d2h_bytes_per_step": int(FRAME_BYTES + N_SPRITES * 16 + 32), "steps": e2e_steps,
                   "sync_per_frame": e2e_sync, "pipelined_depth3": e2e_piped,
                   "note": "pm3d_render with pinned host buffers; value = the better of per-frame sync and the "
                           "3-deep pipeline (same copies per step, D2H of frame i overlaps the kernels of frame i+1); "
                           "the pipelined figure sits at the PCIe Gen5 x16 D2H ceiling of the 8.3 MB frame"}
    line["gpu_launches"] = int(launches * world)
    line["roofline"] = {"bound": "hbm", "kernel": "k_compose", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650 GB/s",
                        "algorithmic_bytes_per_launch": FRAME_BYTES, "kernel_ms": compose_ms,
                        "note": "kernel_ms = mean CUDA-event duration of k_compose over the same frames with the front/compose overlap "
                                "switched off (one kernel at a time, like the ncu launch list); in the timed region of `value` the "
                                "front launch of frame i+1 runs beside it, so ms_per_step < kernel_ms + front.  The kernel is "
                                "instruction/latency bound, not HBM bound (DESIGN.md section 5)",
                        "other_kernels_ms": {"k_front (column cast + view setup)": float(kt[:, 0].mean()) if len(kt) else None}}
    line["mrays_per_s"] = value * W_PX * 3 / 1e6
    line["mpixels_per_s"] = value * W_PX * H_PX / 1e6

    # ---- batched sweeps (c4 views, c5 LOS rays), sharded over the same ranks
    extra = {}
    if not args.no_sweeps:
        try:
            if rank == 0:
                extra["c3"] = sweep_c3(ctx)
            extra["c4"] = sweep_c4(ctx, rank, world, dist, torch, dev)
            extra["c5"] = sweep_c5(ctx, rank, world, dist, torch, dev)
            extra["moves"] = sweep_moves(ctx, rank, world, dist, torch, dev)
            extra["clutter"] = sweep_clutter(ctx, rank, world, dist, torch, dev, rank == 0 and world == 1 and not args.no_cpu)
            extra["tick"] = sweep_tick(ctx, rank, world, dist, torch, dev)
            extra["astar"] = sweep_astar(ctx, rank, world, dist, torch, dev, rank == 0 and world == 1 and not args.no_cpu)
            extra["present"] = sweep_present(ctx, torch)
        except Exception as e:
            # a failed sweep is a failed run: the headline line is still printed (with the error), then every rank
            # leaves with a non-zero status WITHOUT tearing NCCL down over a possibly poisoned CUDA context
            import traceback

            traceback.print_exc()
            extra["error"] = repr(e)
            line["extra"] = extra
            line["error"] = f"sweep failed on rank {rank}: {e!r}"
            if rank == 0:
                print(json.dumps(line), flush=True)
            sys.stderr.flush()
            os._exit(1)
    if extra:
        # the sharded sweeps as first-class numbers: short, numeric, ahead of `extra` (which a tail-truncating reader
        # may lose), repeated inside `roofline` and on stderr
        sw = {"n_gpus": world}
        if "c4" in extra:
            sw.update(c4_views_per_s=round(extra["c4"]["views_per_s"], 1), c4_ms_per_sweep=round(extra["c4"]["ms_per_sweep"], 4),
                      c4_hbm_frac_sweep=round(extra["c4"]["hbm_frac_sweep"], 4), c4_views_per_gpu=extra["c4"]["views_per_gpu"])
        if "c5" in extra:
            sw.update(c5_grays_per_s=round(extra["c5"]["mrays_per_s"] / 1e3, 3), c5_ms_per_batch=round(extra["c5"]["ms_per_batch"], 4),
                      c5_hbm_frac=round(extra["c5"]["hbm_frac"], 4))
        if "tick" in extra:
            sw.update(tick_views_per_s=round(extra["tick"]["views_per_s"], 1), tick_ms=round(extra["tick"]["ms_per_tick"], 4))
        if "moves" in extra:
            sw.update(moves_g_per_s=round(extra["moves"]["mmoves_per_s"] / 1e3, 3))
        if "clutter" in extra:
            sw.update(clutter_m_updates_per_s=round(extra["clutter"]["updates_per_s"] / 1e6, 2))
        if "astar" in extra:
            sw.update(astar_k_paths_per_s=round(extra["astar"]["paths_per_s"] / 1e3, 2))
        sw["gather"] = "in the timed region (library-owned NCCL communicator, pm3d_comm_*)"
        line["sharded"] = sw
        line["roofline"]["sweeps"] = sw
        if rank == 0:
            print("[pm3d-sweeps] " + json.dumps(sw), file=sys.stderr, flush=True)
    line["extra"] = extra

    # ---- CPU baseline beside it (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n, sec = cpu_frames(cfg, 10.0, threads)
        n1, sec1 = cpu_frames(cfg, 3.0, 1)
        line["cpu_baseline"] = {"value": n / sec, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{n} full 1920x1080 frames of the same ring in {sec:.1f} s (oracle/ C restatement, OpenMP); single thread: {n1 / sec1:.2f} frames/s",
                                "single_thread_value": n1 / sec1}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweeps", action="store_true", help="skip the c4/c5 batched sweeps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
