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
