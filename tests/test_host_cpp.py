"""The C++ mirror of raycaster.Camera (pixelmek-3d_b200/host/raycaster_camera.hpp) driving the C-ABI.

CPU: the demo compiles, links against libpm3d.so and fails loudly without a GPU (no CPU fallback).
GPU: the frame it renders through NewCamera/setters/Update/Draw equals the oracle's, byte for byte."""
import math
import os
import subprocess

import numpy as np
import pytest

import pm3d
from conftest import orc_camera
from oracle import oracle as O
from pm3d import camera as cam
from pm3d.scene import SceneData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pixelmek-3d_b200", "host")
DEMO = os.path.join(HOST, "camera_demo")


def build_demo():
    src = [os.path.join(HOST, "camera_demo.cpp"), os.path.join(HOST, "raycaster_camera.hpp")]
    if not os.path.exists(DEMO) or any(os.path.getmtime(s) > os.path.getmtime(DEMO) for s in src):
        subprocess.check_call([
            "g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", DEMO, src[0],
            "-L" + os.path.join(ROOT, "pixelmek-3d_b200"), "-lpm3d", "-Wl,-rpath,$ORIGIN/.."])
    return DEMO


def make_tile(size, seed, holes):
    yy, xx = np.mgrid[0:size, 0:size]
    im = np.zeros((size, size, 4), dtype=np.uint8)
    im[..., 0] = (xx * 3 + seed * 40) & 255
    im[..., 1] = (yy * 2 + seed * 70) & 255
    im[..., 2] = ((xx ^ yy) + seed * 20) & 255
    im[..., 3] = 255
    if holes:
        im[((xx % 32) > 20) & ((yy % 32) > 20)] = 0
    return im


def test_cpp_mirror_builds_and_refuses_to_run_without_gpu(ensure_built):
    demo = build_demo()
    if pm3d.load().pm3d_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([demo], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_frame_matches_oracle(gpu_ctx, tmp_path):
    demo = build_demo()
    out = str(tmp_path / "frame.bin")
    r = subprocess.run([demo, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H, T = 320, 200, 64
    frame = np.fromfile(out, dtype=np.uint8).reshape(H, W, 4)

    # the same scene through the Python mirror's flattening and the oracle
    levels = np.zeros((2, 12, 10), dtype=np.uint16)
    levels[0, :, 0] = levels[0, :, 9] = 1
    levels[0, 0, :] = levels[0, 11, :] = 1
    levels[0, 7, 4] = 2
    levels[1, 7, 4] = 2
    levels[0, 8, 6] = 1
    wall1, wall2, wall2x = make_tile(T, 1, False), make_tile(T, 2, True), make_tile(T, 3, False)
    floor_a, floor_b, sky, box = make_tile(T, 4, False), make_tile(T, 5, False), make_tile(T, 6, False), make_tile(T, 7, False)
    sheet = make_tile(48, 8, True)[:40]
    xs, ys = np.mgrid[0:12, 0:10]
    atlas = np.stack([wall1, wall2, wall2x, floor_a, floor_b, sky, box])
    scene = SceneData(
        grid=levels, num_levels=2, wall_tex_index=np.array([1, 2], np.int32), wall_tile_side0=np.array([0, 1], np.int32),
        wall_tile_side1=np.array([0, 2], np.int32), floor_tiles=np.where((xs + ys) % 3 == 0, 4, 3).astype(np.int16),
        atlas=atlas, sky_tile=5, floorbox_tile=6, sheets=[sheet])
    st = cam.CameraState(W, H)
    st.pos, st.pos_z, st.heading, st.pitch_angle = (3.3, 4.6), 0.6, 0.2, 0.05
    st.fov_angle, st.fov_depth, st.render_distance = math.radians(70.0), 1.0, 50.0
    st.light_falloff, st.global_illumination, st.min_light, st.max_light = -60.0, 120.0, (40, 50, 60), (255, 240, 230)
    spr = np.zeros(2, dtype=O.SPRITE_DTYPE)
    spr["x"], spr["y"], spr["z"], spr["scale"] = [5.5, 6.2], [5.0, 3.9], [0.0, 0.3], [0.7, 0.5]
    spr["illumination"], spr["focusable"], spr["sheet"] = 10.0, 1, 0
    spr["rx0"], spr["ry0"], spr["rx1"], spr["ry1"] = [0, 4], [0, 2], [48, 40], [40, 38]
    ref = O.render(O.Scene(scene), orc_camera(st.view_record()[0], W, H), spr)
    bad = int((frame != ref["rgba"]).any(axis=-1).sum())
    assert bad == 0, f"{bad} pixels differ between the C++ mirror's frame and the oracle"
    want = {-1: "none", 0: "a", 1: "b"}[ref["convergence"]["sprite"]]
    assert f"convergence sprite {want}" in r.stdout
    rc = ref["rects"][0]
    assert f"[{rc[0]} {rc[1]} {rc[2]} {rc[3]}]" in r.stdout


# ---- the C++ mirror of Game.lineOfSight / Game.getValidMove(checkAlternate=true) (host/game_collision.hpp)
COL_DEMO = os.path.join(HOST, "collision_demo")


def build_collision_demo():
    src = [os.path.join(HOST, "collision_demo.cpp"), os.path.join(HOST, "game_collision.hpp")]
    if not os.path.exists(COL_DEMO) or any(os.path.getmtime(s) > os.path.getmtime(COL_DEMO) for s in src):
        subprocess.check_call([
            "g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", COL_DEMO, src[0],
            "-L" + os.path.join(ROOT, "pixelmek-3d_b200"), "-lpm3d", "-Wl,-rpath,$ORIGIN/.."])
    return COL_DEMO


def test_cpp_collision_mirror_builds(ensure_built):
    demo = build_collision_demo()
    if pm3d.load().pm3d_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([demo, "missing.bin", "out.bin"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_get_valid_move_with_check_alternate_matches_oracle(gpu_ctx, tmp_path):
    """pm3d::Game::getValidMove(entity, x, y, z, true) — the checkAlternate recursion of game/collision.go:204-306 in
    C++ over pm3d_valid_move_batch — against the oracle's recursion, and lineOfSight through the same class."""
    from pm3d import workloads as Wk
    from test_move_gpu import random_colliders, random_movers

    demo = build_collision_demo()
    g = Wk.load_golden("debug_land")
    lvl = np.ascontiguousarray(g["levels"][:1])
    scene = Wk.scene_from_grid(lvl, 1)
    segs = np.ascontiguousarray(g["segments"], dtype=np.float64)
    rng = np.random.default_rng(17)
    c = random_colliders(rng, lvl[0], 90).astype(pm3d._lib.COLLIDER_DTYPE)
    m = random_movers(rng, lvl[0], c, 1200, 2.5).astype(pm3d._lib.MOVER_DTYPE)
    m["radius"] = np.maximum(m["radius"], 0.05)
    m["height"] = np.maximum(m["height"], 0.05)
    m["self"] = -1; m["parent"] = -1; m["destroyed"] = 0
    hv = np.stack([rng.random(len(m)) * 2 * math.pi - math.pi, rng.random(len(m)) * 0.2], axis=1)
    scen = str(tmp_path / "scenario.bin")
    with open(scen, "wb") as f:
        np.array([lvl.shape[1], lvl.shape[2], len(segs), len(c), len(m)], dtype=np.int32).tofile(f)
        np.ascontiguousarray(lvl[0], dtype=np.uint16).tofile(f)
        segs.tofile(f)
        c.tofile(f)
        for k in range(len(m)):
            m[k:k + 1].tofile(f)
            hv[k].tofile(f)
    out = str(tmp_path / "results.bin")
    r = subprocess.run([demo, scen, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rec_dt = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("collided", "<i4"), ("n", "<i4"), ("first", "<i4"), ("_pad", "<i4"),
                       ("fx", "<f8"), ("fy", "<f8"), ("fz", "<f8")])
    raw = np.fromfile(out, dtype=np.uint8)
    got = raw[:len(m) * rec_dt.itemsize].view(rec_dt)
    los = raw[len(m) * rec_dt.itemsize:]
    osc = O.Scene(scene)
    n_alt = 0
    for k in range(len(m)):
        want = O.valid_move_alternate(osc, segs, c.astype(O.COLLIDER_DTYPE), 0, Wk.CLIP_DISTANCE, m[k:k + 1].astype(O.MOVER_DTYPE),
                                      float(hv[k, 0]), float(hv[k, 1]), True)
        assert (got[k]["x"], got[k]["y"], got[k]["z"], got[k]["collided"]) == (want["x"], want["y"], want["z"], int(want["collided"])), k
        assert got[k]["n"] == (1 if want["n_collisions"] > 0 else 0)
        if got[k]["n"]:
            assert (got[k]["first"], got[k]["fx"], got[k]["fy"], got[k]["fz"]) == (want["first_entity"], want["first_x"], want["first_y"], want["first_z"])
        n_alt += int(want["n_points"] > 0)
    assert n_alt > 20
    rays = np.stack([m["pos_x"], m["pos_y"], c["x"][np.arange(len(m)) % len(c)], c["y"][np.arange(len(m)) % len(c)]], axis=1)
    assert np.array_equal(los, O.los_batch(segs, rays, threads=4))
