"""CPU ORACLE (test infrastructure) — Python restatement of the reference's one-time,
host-side map preparation that the hot path's results depend on.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (pixelmek-3d_b200/) never does.

PARITY UNPINNED: the reference has no tests or golden vectors; these
functions follow the cited Go sources line by line, but were never run against the Go binary.

  generate_map_levels            game/model/map.go:434-554   [V]
  floor_texture_grid             game/texture/texture.go:145-192 [V]
  generate_wall_collision_lines  game/model/map.go:556-884   [V]
  geom.Rect / Line.Angle / LineFromAngle / Distance: raycaster-go geom [R]
"""
from __future__ import annotations

import math

import numpy as np

NORTH, EAST, SOUTH, WEST = 0, 1, 2, 3
BOUNDARY_TEX = 32767  # math.MaxInt16, map.go:458


def _line_points(x1, y1, x2, y2):
    """Cells visited by the 0.1-step walk along a segment (map.go:534-545, texture.go:174-181)."""
    angle = math.atan2(y2 - y1, x2 - x1)  # geom.Line.Angle [R]
    dist = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)  # geom.Distance [R]
    d = 0.0
    while d <= dist:
        px = x1 + d * math.cos(angle)  # geom.LineFromAngle [R]
        py = y1 + d * math.sin(angle)
        yield int(px), int(py)
        d += 0.1


def generate_map_levels(gen: dict, num_raycast_levels: int) -> np.ndarray:
    """Map.generateMapLevels (map.go:434-554). Returns levels[level][x][y] as uint16."""
    size_x, size_y = gen["mapSize"]
    if size_x <= 0 or size_y <= 0:
        raise ValueError("map X/Y size must both be greater than zero")
    levels = np.zeros((num_raycast_levels, size_x, size_y), dtype=np.int64)

    boundary = gen.get("boundaryWall") or {}
    if len(boundary.get("image", "") or "") > 0:  # HasBoundaryWall map.go:229
        lvl = levels[0]
        lvl[0, :] = BOUNDARY_TEX
        lvl[size_x - 1, :] = BOUNDARY_TEX
        lvl[:, 0] = BOUNDARY_TEX
        lvl[:, size_y - 1] = BOUNDARY_TEX

    for prefab in gen.get("prefabs") or []:
        layers = prefab["layers"]
        n_layers = len(layers)
        if n_layers == 0 or len(prefab["positions"]) == 0:
            raise ValueError("prefab must have at least one layer and one position")
        if n_layers > num_raycast_levels:
            raise ValueError("prefab cannot have more layers than numRaycastLevels")
        for pos_x, pos_y in prefab["positions"]:
            for i in range(n_layers):
                i_rev = n_layers - i - 1  # map.go:493
                rows = layers[i]
                for y in range(len(rows)):
                    y_rev = len(rows) - y - 1  # map.go:496
                    for x in range(len(rows[y])):
                        if x + pos_x >= size_x or y_rev + pos_y >= size_y:
                            continue
                        levels[i_rev, x + pos_x, y_rev + pos_y] = rows[y][x]

    for wall in gen.get("walls") or []:
        tex, height = wall["texture"], wall["height"]
        if tex < 0 or height < 1 or height > num_raycast_levels:
            raise ValueError("bad generated wall")
        for segments in wall["lines"]:
            prev = None
            for seg in segments:
                point = (float(seg[0]), float(seg[1]))
                if prev is not None:
                    for cx, cy in _line_points(prev[0], prev[1], point[0], point[1]):
                        levels[0:height, cx, cy] = tex
                prev = point
    return levels.astype(np.uint16)


def floor_texture_grid(flooring: dict, size_x: int, size_y: int):
    """TextureHandler.loadMapTextures floor part (texture.go:145-192).

    Returns (grid[x][y] of int16 index into `images`, images) with index 0 = default floor."""
    images = [flooring.get("default", "")]
    grid = np.zeros((size_x, size_y), dtype=np.int16)
    for pathing in flooring.get("pathing") or []:
        images.append(pathing["image"])
        idx = len(images) - 1
        for rect in pathing.get("rects") or []:
            (x0, y0), (x1, y1) = rect
            grid[x0 : x1 + 1, y0 : y1 + 1] = idx  # inclusive bounds texture.go:161-165
        for segments in pathing.get("lines") or []:
            prev = None
            for seg in segments:
                point = (float(seg[0]), float(seg[1]))
                if prev is not None:
                    for cx, cy in _line_points(prev[0], prev[1], point[0], point[1]):
                        grid[cx, cy] = idx
                prev = point
    return grid, images


# --------------------------------------------------------------------------- L0


def _geom_rect(x, y, w, h):
    """geom.Rect [R]: four edges, counter-clockwise from (x, y)."""
    return [
        [x, y, x, y + h],
        [x, y + h, x + w, y + h],
        [x + w, y + h, x + w, y],
        [x + w, y, x, y],
    ]


def _create_wall_line(x, y, d):
    """createWallLine (map.go:854-884)."""
    x1 = x2 = float(x)
    y1 = y2 = float(y)
    if d == NORTH:
        y2 += 1
    elif d == EAST:
        y1 += 1
        x2 += 1
        y2 += 1
    elif d == SOUTH:
        x1 += 1
        y1 += 1
        x2 += 1
    elif d == WEST:
        x1 += 1
    return {"line": [x1, y1, x2, y2], "dir": d, "cancelled": False, "visited": False}


_STEP = {NORTH: (0, 1), EAST: (1, 0), SOUTH: (0, -1), WEST: (-1, 0)}
_DIAG = {NORTH: (-1, 1), EAST: (1, 1), SOUTH: (1, -1), WEST: (-1, -1)}


def generate_wall_collision_lines(level0: np.ndarray, has_boundary_wall: bool, clip: float) -> np.ndarray:
    """Map.GenerateWallCollisionLines (map.go:811-851). Returns float64 [n][4] = x1,y1,x2,y2."""
    w, h = level0.shape
    if w == 0 or h == 0:
        return np.zeros((0, 4))
    lines = []
    if has_boundary_wall:  # map.go:820-827
        lines += _geom_rect(1 + clip, 1 + clip, float(w - 2) - 2 * clip, float(h - 2) - 2 * clip)
    lines += _geom_rect(clip, clip, float(w) - 2 * clip, float(h) - 2 * clip)  # map.go:830-834

    # Phase 1 (map.go:662-696): 4 directed border lines per wall cell, cancel opposite duplicates
    cells = {}
    wall_set = {}
    for x in range(w):
        for y in range(h):
            if has_boundary_wall and (x == 0 or y == 0 or x == w - 1 or y == h - 1):
                continue
            if level0[x, y] == 0:
                continue
            cb = {"lines": {d: _create_wall_line(x, y, d) for d in (NORTH, EAST, SOUTH, WEST)}, "visited": False}
            cells[(x, y)] = cb
            for dl in cb["lines"].values():
                l = dl["line"]
                opp = (l[2], l[3], l[0], l[1])
                if opp in wall_set:
                    dl["cancelled"] = True
                    wall_set[opp]["cancelled"] = True
                else:
                    wall_set[tuple(l)] = dl

    def cell_at(i, j):
        if i < 0 or j < 0 or i >= w or j >= h:
            return None
        return cells.get((i, j))

    # Phase 2 (map.go:698-776): walk outlines clockwise into groups
    groups = []
    for y in range(h):
        for x in range(w):
            cell = cells.get((x, y))
            if cell is None or cell["visited"] or all(l["cancelled"] for l in cell["lines"].values()):
                continue
            group = []
            groups.append(group)
            line_dir = NORTH
            og = cell["lines"][line_dir]["line"]
            og_x, og_y = og[0], og[1]
            i, j = x, y
            start_dir = line_dir
            while True:
                cell["visited"] = True
                line = cell["lines"].get(line_dir)
                if line is None or line["cancelled"] or line["visited"]:
                    line_dir = (line_dir + 1) % 4
                    if line_dir == start_dir:
                        break  # "unable to escape cell lines"
                    continue
                start_dir = line_dir
                line["visited"] = True
                group.append(line)
                di, dj = _STEP[line_dir]
                peek = cell_at(i + di, j + dj)
                if peek is not None:
                    pl = peek["lines"].get(line_dir)
                    if pl is not None and not pl["cancelled"] and not pl["visited"]:
                        cell = peek
                        i, j = i + di, j + dj
                        continue
                di, dj = _DIAG[line_dir]
                diag = cell_at(i + di, j + dj)
                if diag is not None:
                    diag_dir = (line_dir - 1) % 4  # Prev()
                    dl = diag["lines"].get(diag_dir)
                    if dl is not None and not dl["cancelled"] and not dl["visited"]:
                        cell = diag
                        line_dir = diag_dir
                        i, j = i + di, j + dj
                        continue
                if line["line"][2] == og_x and line["line"][3] == og_y:
                    break

    # Phase 3 (map.go:783-809): merge consecutive same-direction lines
    for group in groups:
        if not group:
            continue
        cur = group[0]
        prev_dir = cur["dir"]
        merged = [cur["line"]]
        for wl in group[1:]:
            if wl["dir"] == prev_dir:
                cur["line"][2], cur["line"][3] = wl["line"][2], wl["line"][3]
            else:
                cur = wl
                merged.append(cur["line"])
            prev_dir = wl["dir"]
        lines += merged
    return np.array(lines, dtype=np.float64).reshape(-1, 4)
