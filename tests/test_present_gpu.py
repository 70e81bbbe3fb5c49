"""GPU parity of the presentation step (pm3d_present; game/scene_game.go:169-185 + crt.kage) against the oracle.
The scale and the copy are integer work: bit-exact.  The CRT shader is float32 with powf / sinf, where CUDA's and
the C library's routines differ in the last bits: the tolerance is the north star's, <= 1 LSB on <= 0.1 % of
the channel values (alpha and every sampling position are exact)."""
import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def frame(ctx, w, h):
    c = W.config_c2(w=w, h=h, n_sprites=60, n_poses=1)
    ctx.upload_scene(c["scene"])
    ctx.set_view(w, h, 1, 60)
    return ctx.render(c["views"], c["sprites"], c["offsets"])["rgba"][0]


@pytest.mark.parametrize("w,h,ow,oh,scale", [(512, 384, 1024, 768, 0.5), (768, 576, 1024, 768, 0.75), (640, 360, 640, 360, 1.0),
                                             (333, 187, 1001, 563, 1 / 3), (614, 460, 1024, 768, 0.6)])
def test_scale_copy_is_exact(gpu_ctx, w, h, ow, oh, scale):
    src = frame(gpu_ctx, w, h)
    gpu_ctx.set_present(ow, oh, scale)
    got = gpu_ctx.present(pm3d._lib.PRESENT_COPY)[0]
    want = O.present(src, ow, oh, scale, 0)
    assert np.array_equal(got, want)
    # independent statement of the mapping: output pixel (X, Y) shows source texel floor((X + .5) * scale)
    xs = np.floor((np.arange(ow) + 0.5) * scale).astype(int)
    ys = np.floor((np.arange(oh) + 0.5) * scale).astype(int)
    ok = (xs[None, :] < w) & (ys[:, None] < h)
    ref = np.where(ok[..., None], src[np.minimum(ys, h - 1)][:, np.minimum(xs, w - 1)], 0)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("w,h,ow,oh,scale", [(512, 384, 1024, 768, 0.5), (768, 576, 1024, 768, 0.75), (640, 360, 640, 360, 1.0),
                                             (333, 187, 1001, 563, 1 / 3)])
def test_crt_shader_within_one_lsb(gpu_ctx, mode, w, h, ow, oh, scale):
    src = frame(gpu_ctx, w, h)
    gpu_ctx.set_present(ow, oh, scale)
    got = gpu_ctx.present(mode)[0].astype(np.int16)
    want = O.present(src, ow, oh, scale, mode).astype(np.int16)
    diff = np.abs(got - want)
    assert diff.max() <= 1, f"max channel difference {diff.max()}"
    share = float((diff > 0).mean())
    assert share <= 1e-3, f"{share:.2e} of the channel values differ by 1 LSB"
    assert (got[..., 3] == 255).all()
    assert got[..., :3].max() > 100          # not a black frame
    if mode == 2:                            # the curved screen leaves black corners
        assert got[0, 0, :3].max() == 0


def test_present_batch_resident_and_errors(gpu_ctx):
    ctx = gpu_ctx
    c = W.config_c4(n_views=6, w=320, h=200)
    ctx.upload_scene(c["scene"])
    ctx.set_view(320, 200, 6, len(c["sprites"]))
    with pytest.raises(pm3d.Pm3dError, match="set_present"):
        ctx.present(0, 6)
    res = ctx.render(c["views"], c["sprites"], c["offsets"])
    ctx.set_present(640, 400, 0.5)
    got = ctx.present(pm3d._lib.PRESENT_CRT, 6)
    for v in range(6):
        want = O.present(res["rgba"][v], 640, 400, 0.5, 1).astype(np.int16)
        assert np.abs(got[v].astype(np.int16) - want).max() <= 1
    # device-resident output equals the host output
    d_out = ctx.device_alloc(6 * 640 * 400 * 4)
    ctx.present(pm3d._lib.PRESENT_CRT, 6, d_out=d_out)
    ctx.sync()
    back = np.zeros_like(got)
    ctx.d2h(back, d_out)
    ctx.device_free(d_out)
    assert np.array_equal(back, got)
    with pytest.raises(pm3d.Pm3dError, match="mode"):
        ctx.present(7, 6)
