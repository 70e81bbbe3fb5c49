"""Generates tests/golden/proving_grounds_pixels.npz: the reference's REAL wall / floor / sky pixels and sprite
sheets for the bundled proving_grounds map and the trial_day mission, decoded with PIL.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box).  The fixture is DATA
derived from the reference's PNG assets by this script; no reference source is copied.  Pixels are stored as decoded
(straight, non-premultiplied RGBA8); pm3d.workloads premultiplies them the way the game's loaders do
(ebiten.NewImageFromImage for walls / sprites, image.RGBA through draw.Draw for the floor cast:
game/resources/resources.go:215-275).

What it holds:
  tiles        every 256x256 image the map's texture table, flooring, floorBox and skyBox name
               (game/resources/maps/proving_grounds.yaml), in a fixed order, with their paths
  sheets       shrubbery/tree_0, tree_1 (the map's static sprites), mechs/adder (6x1 strip, 240x42), rocks/rock_0,
               shrubbery/bush_0 (the map's clutter) and the three other mechs of missions/trial_day.yaml

    python tests/golden/make_pixels.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/game/resources"


def rgba(path):
    return np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)


def main():
    g = np.load(os.path.join(HERE, "proving_grounds.npz"))
    names = []
    for arr in (g["tex_image"], g["tex_side_x"], g["tex_side_y"], g["floor_images"], [g["floor_box"]], [g["sky_box"]]):
        for nm in arr:
            nm = str(nm)
            if nm and nm not in names:
                names.append(nm)
    tiles = np.stack([rgba(os.path.join(REF, "textures", nm)) for nm in names])
    assert tiles.shape[1:] == (256, 256, 4), tiles.shape
    sheet_names = ["shrubbery/tree_0.png", "shrubbery/tree_1.png", "mechs/adder.png", "rocks/rock_0.png", "shrubbery/bush_0.png",
                   "mechs/warhawk.png", "mechs/executioner.png", "mechs/timber_wolf.png"]
    out = {"tile_names": np.array(names), "tiles": tiles, "sheet_names": np.array(sheet_names)}
    for i, nm in enumerate(sheet_names):
        out[f"sheet_{i}"] = rgba(os.path.join(REF, "sprites", nm))
    path = os.path.join(HERE, "proving_grounds_pixels.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(names), "tiles:", names, "; sheets:",
          [(nm, out[f"sheet_{i}"].shape[:2]) for i, nm in enumerate(sheet_names)])
    for i, nm in enumerate(sheet_names):
        a = out[f"sheet_{i}"][..., 3]
        print(f"  {nm}: alpha 0: {(a == 0).sum()}, 255: {(a == 255).sum()}, translucent: {((a > 0) & (a < 255)).sum()}")
    for i, nm in enumerate(names):
        a = tiles[i][..., 3]
        print(f"  {nm}: alpha 0: {(a == 0).sum()}, translucent: {((a > 0) & (a < 255)).sum()}")


if __name__ == "__main__":
    main()
