"""Generates tests/golden/*.npz from the reference's bundled map/mission YAML.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  It imports the Python ORACLE restatement (oracle/mapgen.py) of
  Map.generateMapLevels            game/model/map.go:434-554
  TextureHandler.loadMapTextures   game/texture/texture.go:145-192
  Map.GenerateWallCollisionLines   game/model/map.go:811-851
and freezes their outputs for the two bundled maps, together with the map constants the
path needs (lighting, texture table, sprite placements).  The fixtures are DATA derived from
the reference's YAML by this script; no reference source is copied.

spriteFill placements use Go's math/rand (game/model/map.go:364-399: rng.Seed(m.Seed + n), then
RandFloat64In per coordinate and height, game/model/rand.go:36-44).  The oracle restates that generator
(oracle/pm3d_oracle_clutter.c, pinned against the library's published seed-1 outputs in
tests/test_clutter_cpu.py), so the fill positions frozen here are the ones the game itself generates.
Clutter (Map.Clutter, Map.Seed) rides along for the clutter tests.

    python tests/golden/make_fixtures.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import mapgen  # noqa: E402

REF = "/root/reference/game/resources"
METERS_PER_UNIT = 20.0  # game/model/model.go:12
CLIP_DISTANCE = 0.01    # game/game.go:33

def png_size(path):
    with open(path, "rb") as f:
        head = f.read(24)
    return int.from_bytes(head[16:20], "big"), int.from_bytes(head[20:24], "big")


def build(map_name: str, mission_name: str | None):
    with open(os.path.join(REF, "maps", map_name + ".yaml")) as f:
        m = yaml.safe_load(f)
    n_levels = int(m.get("numRaycastLevels") or 0)
    gen = m.get("generateLevels")
    if gen and gen.get("mapSize"):
        levels = mapgen.generate_map_levels(gen, n_levels)
        has_boundary = bool((gen.get("boundaryWall") or {}).get("image"))
        boundary_image = (gen.get("boundaryWall") or {}).get("image", "")
    else:
        levels = np.array(m["levels"], dtype=np.uint16)
        has_boundary = False
        boundary_image = ""
        if n_levels == 0:
            n_levels = levels.shape[0]
    size_x, size_y = levels.shape[1], levels.shape[2]
    floor_grid, floor_images = mapgen.floor_texture_grid(m["flooring"], size_x, size_y)
    segs = mapgen.generate_wall_collision_lines(levels[0], has_boundary, CLIP_DISTANCE)

    # texture table: index -> (image, sideX, sideY)   (game/model/map.go:100-115)
    tex_idx, tex_image, tex_sx, tex_sy = [], [], [], []
    for k, v in sorted((m.get("textures") or {}).items()):
        tex_idx.append(int(k))
        tex_image.append((v or {}).get("image", "") or "")
        tex_sx.append((v or {}).get("sideX", "") or "")
        tex_sy.append((v or {}).get("sideY", "") or "")
    if has_boundary:
        tex_idx.append(mapgen.BOUNDARY_TEX)
        tex_image.append(boundary_image)
        tex_sx.append("")
        tex_sy.append("")

    # sprites: explicit positions, fills (Go math/rand through the oracle's restatement), stamps
    by_id = {s["id"]: s for s in (m.get("sprites") or []) if s.get("id")}
    spr = []  # (x, y, z, scale, image, anchor)

    def add(s, x, y, height):
        anchor = {"top": 2, "center": 1}.get((s.get("anchor") or "bottom"), 0)
        spr.append((float(x), float(y), float(s.get("zPosition") or 0.0), float(height) / METERS_PER_UNIT, s["image"], anchor))

    for s in m.get("sprites") or []:
        for pos in s.get("positions") or []:
            add(s, pos[0], pos[1], s.get("height") or 0)
    from oracle import oracle as O
    map_seed = int(m.get("seed") or 0)
    for n, fill in enumerate(m.get("spriteFill") or []):
        rng = O.GoRand(map_seed + n)                           # rng.Seed(m.Seed + int64(n)), map.go:372
        (x0, y0), (x1, y1) = [[float(v) for v in p] for p in fill["rect"]]
        hr = fill.get("heightRange") or []
        s = by_id[fill["sprite"]]
        for _ in range(int(fill["quantity"])):
            fx = x0 + (x1 - x0) * rng.float64()                 # RandFloat64In, rand.go:36-44
            fy = y0 + (y1 - y0) * rng.float64()
            height = float(hr[0]) + (float(hr[1]) - float(hr[0])) * rng.float64() if len(hr) == 2 else 0.0
            add(s, fx, fy, height)
    for stamp in m.get("spriteStamps") or []:
        for px, py in stamp["positions"]:
            for ref in stamp["sprites"]:
                s = by_id[ref["sprite"]]
                for sx, sy in ref.get("positions") or []:
                    add(s, px + sx, py + sy, ref.get("height") or 0)
    images = sorted({s[4] for s in spr})
    image_sizes = [png_size(os.path.join(REF, "sprites", im)) for im in images]

    out = {
        "levels": levels, "num_raycast_levels": np.int32(n_levels), "has_boundary_wall": np.bool_(has_boundary),
        "floor_grid": floor_grid, "floor_images": np.array(floor_images), "segments": segs,
        "tex_index": np.array(tex_idx, dtype=np.int32), "tex_image": np.array(tex_image),
        "tex_side_x": np.array(tex_sx), "tex_side_y": np.array(tex_sy),
        "floor_box": np.array(m["floorBox"]["image"]), "sky_box": np.array(m["skyBox"]["image"]),
        "light_falloff": np.float64(m["lighting"]["falloff"]), "light_illumination": np.float64(m["lighting"]["illumination"]),
        "min_light": np.array(m["lighting"]["minLightRGB"], dtype=np.uint8),
        "max_light": np.array(m["lighting"]["maxLightRGB"], dtype=np.uint8),
        "sprite_xyzs": np.array([s[:4] for s in spr], dtype=np.float64).reshape(-1, 4),
        "sprite_image": np.array([images.index(s[4]) for s in spr], dtype=np.int32),
        "sprite_anchor": np.array([s[5] for s in spr], dtype=np.uint8),
        "sprite_images": np.array(images), "sprite_image_sizes": np.array(image_sizes, dtype=np.int32).reshape(-1, 2),
        "map_seed": np.int64(map_seed),
    }
    clutter = m.get("clutter") or []
    if clutter:
        out["clutter_image"] = np.array([c["image"] for c in clutter])
        out["clutter_match"] = np.array([c.get("floorPathMatch") or "" for c in clutter])
        out["clutter_frequency"] = np.array([float(c["frequency"]) for c in clutter], dtype=np.float64)
        out["clutter_height"] = np.array([float(c["height"]) for c in clutter], dtype=np.float64)
        out["clutter_image_size"] = np.array([png_size(os.path.join(REF, "sprites", c["image"])) for c in clutter],
                                             dtype=np.int32).reshape(-1, 2)
    if mission_name:
        with open(os.path.join(REF, "missions", mission_name + ".yaml")) as f:
            ms = yaml.safe_load(f)
        out["drop_zone"] = np.array(ms["dropZone"]["position"], dtype=np.float64)
        out["drop_heading_deg"] = np.float64(ms["dropZone"]["heading"])
        units = []
        for kind in ("mechs", "vehicles", "vtols", "infantry", "emplacements"):
            for u in ms.get(kind) or []:
                units.append([float(u["position"][0]), float(u["position"][1]), float(u.get("zPosition") or 0.0)])
        out["mission_units"] = np.array(units, dtype=np.float64).reshape(-1, 3)
    return out


def main():
    for map_name, mission in (("proving_grounds", "trial_day"), ("debug_land", None)):
        data = build(map_name, mission)
        path = os.path.join(HERE, map_name + ".npz")
        np.savez_compressed(path, **data)
        lv = data["levels"]
        print(f"{path}: levels {lv.shape}, wall cells per level {[int((lv[i] > 0).sum()) for i in range(lv.shape[0])]}, "
              f"{len(data['segments'])} collision segments, {len(data['sprite_xyzs'])} sprites, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
