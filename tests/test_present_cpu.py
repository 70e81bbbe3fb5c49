"""CPU tests of the presentation-step oracle (oracle/pm3d_oracle_present.c) against hand-derived answers."""
import math

import numpy as np

from oracle import oracle as O


def test_scale_mapping_known_answers():
    src = np.zeros((2, 3, 4), dtype=np.uint8)
    src[..., 0] = np.arange(6).reshape(2, 3) * 10 + 1
    src[..., 3] = 255
    out = O.present(src, 6, 4, 0.5, 0)           # exact 2x: every source texel becomes a 2x2 block
    assert np.array_equal(out, np.repeat(np.repeat(src, 2, axis=0), 2, axis=1))
    out = O.present(src, 8, 4, 0.5, 0)           # screen wider than the scaled image: transparent black beyond it
    assert np.array_equal(out[:, :6], np.repeat(np.repeat(src, 2, axis=0), 2, axis=1)) and not out[:, 6:].any()
    out = O.present(src, 3, 2, 1.0, 0)
    assert np.array_equal(out, src)


def test_crt_uniform_image_interior_pixel():
    """Uniform opaque grey, no curvature: every tap reads the same texel, so the shader reduces to a closed form."""
    f = np.float32
    src = np.full((64, 64, 4), 128, dtype=np.uint8)
    src[..., 3] = 255
    out = O.present(src, 64, 64, 1.0, 1)
    X, Y = 20, 31                                 # even column: the column mask is 1 - 0.65 * clamp((0.5 - 1) * 2) = 1
    t = f(128) / f(255)
    uy, ux = (f(Y) + f(0.5)) / f(64), (f(X) + f(0.5)) / f(64)
    vig = f(16) * ux * uy * (f(1) - ux) * (f(1) - uy)
    vp = f(math.pow(float(vig), 0.3))
    scans = min(max(f(0.35) + f(0.35) * f(math.sin(float(uy * f(64) * f(1.5)))), f(0)), f(1))
    s = f(math.pow(float(scans), 1.7))
    want = []
    for bleed, chan in ((0.08, 0.95), (0.05, 1.05), (0.08, 0.95)):
        col = t + f(0.05)
        col = col + f(bleed) * t
        v = min(max(col * f(0.6) + f(0.4) * col * col, f(0)), f(1))
        v = v * vp * f(chan) * f(2.8) * (f(0.4) + f(0.7) * s)
        want.append(int(math.floor(float(min(max(v, f(0)), f(1)) * f(255) + f(0.5)))))
    got = out[Y, X]
    assert got[3] == 255 and all(abs(int(got[c]) - want[c]) <= 1 for c in range(3)), (got, want)
    # odd columns are darkened by the column mask: 1 - 0.65 * clamp((1.5 - 1) * 2, 0, 1) = 0.35
    assert all(abs(int(out[Y, X + 1][c]) - 0.35 * want[c]) <= 2 for c in range(3))
    # curved mode blacks out the corners (uv leaves [0, 1])
    assert not O.present(src, 64, 64, 1.0, 2)[0, 0, :3].any()
