"""Turns baseline/go_harness's golden_c1.bin (written by the real Go raycaster, see baseline/go_harness/README.md)
into tests/golden/go_c1.npz.  Not runnable in this image: there is no Go toolchain to produce the input."""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"PM3DGOLD", "not a go_harness dump"
    w, h, levels, n_tex, n_floor, n_spr = struct.unpack_from("<6i", raw, 8)
    o = 8 + 24
    tex = np.frombuffer(raw, dtype="<i4", count=4 * n_tex, offset=o).reshape(n_tex, 4); o += 16 * n_tex
    floor = np.frombuffer(raw, dtype="<i4", count=2 * n_floor, offset=o).reshape(n_floor, 2); o += 8 * n_floor
    rects = np.frombuffer(raw, dtype="<i4", count=4 * n_spr, offset=o).reshape(n_spr, 4); o += 16 * n_spr
    conv = np.frombuffer(raw, dtype="<f8", count=3, offset=o); o += 24
    (conv_sprite,) = struct.unpack_from("<i", raw, o)
    out = os.path.join(HERE, "go_c1.npz")
    np.savez_compressed(out, w=w, h=h, levels=levels, texture_at=tex, floor_texture_at=floor, sprite_rects=rects,
                        convergence=conv, convergence_sprite=conv_sprite)
    print(out, tex.shape, floor.shape, rects.shape)


if __name__ == "__main__":
    main(sys.argv[1])
