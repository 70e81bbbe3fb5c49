"""Test configuration.

-m "not gpu": the oracle against known answers and the committed fixtures, the host logic,
and that libpm3d.so loads and exports every symbol include/pm3d.h declares (no compute).
-m gpu: the parity tests proper — the CUDA path, called through the C-ABI, against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pixelmek-3d_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def orc_camera(view, w, h):
    """pm3d_view record -> oracle camera (same numbers, nothing recomputed)."""
    from oracle import oracle as O

    return O.make_camera(dict(
        pos_x=view["pos_x"], pos_y=view["pos_y"], cam_z_px=view["cam_z_px"], dir_x=view["dir_x"], dir_y=view["dir_y"],
        plane_x=view["plane_x"], plane_y=view["plane_y"], pitch_px=view["pitch_px"], w=w, h=h,
        render_distance=view["render_distance"], light_falloff=view["light_falloff"],
        global_illumination=view["global_illumination"], min_light=view["min_light"], max_light=view["max_light"],
    ))


@pytest.fixture(scope="session")
def ensure_built():
    """Build libpm3d.so (nvcc cross-compiles without a GPU) and the oracle if missing."""
    import __graft_entry__ as ge

    ge.build(only_if_missing=True)
    return True


@pytest.fixture(scope="session")
def gpu_ctx(ensure_built):
    import pm3d

    if pm3d.load().pm3d_device_count() < 1:
        pytest.skip("no CUDA device")
    ctx = pm3d.Context(0)
    yield ctx
    ctx.close()


def assert_columns_equal(gpu_cols, orc_cols):
    """Bit-exact comparison of wall-cast records (hit cell, side, texX, distances, tint)."""
    for k in ("map_x", "map_y", "side", "hit", "tex_x", "tile", "line_height", "draw_start", "draw_end"):
        a, b = gpu_cols[k], orc_cols[k]
        assert np.array_equal(a, b), f"{k}: {int((a != b).sum())} mismatches"
    a = gpu_cols["perp_dist"].view(np.uint64)
    b = orc_cols["perp_dist"].view(np.uint64)
    assert np.array_equal(a, b), f"perp_dist bits: {int((a != b).sum())} mismatches"
    assert np.array_equal(gpu_cols["tint"], orc_cols["tint"])
