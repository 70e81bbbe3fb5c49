/*
 * pm3d_oracle_present.c — CPU restatement of the step right after the raycast in GameScene.Draw
 * (game/scene_game.go:169-185): the nearest-neighbour scale of rayScreen onto the screen-sized renderScreen by
 * 1/renderScale, then either a plain copy to overlayScreen or the CRT shader
 * (game/resources/shaders/crt.kage, driven by game/render/effects/crt.go:40-53).
 *
 * TEST INFRASTRUCTURE ONLY (see pm3d_oracle.h).  PARITY UNPINNED: in the reference both steps are ebiten GPU
 * draws (ebiten v2.9.9, not on disk); nothing in the repository pins their pixels.
 *   [V] the shader arithmetic follows crt.kage line by line (citations below), in float32 like a Kage shader;
 *   [D] decisions where the GPU pipeline is not pinned by any file:
 *       - scale: destination pixel (X, Y) shows source texel floor((X + 0.5) * renderScale) (pixel centres,
 *         FilterNearest), transparent black outside the source; evaluated in float64 on the host;
 *       - imageSrc0At(pos) = texel floor(pos) when origin <= pos < origin + size, else vec4(0); origin = (0, 0)
 *         (renderScreen is an off-screen render target, not an atlas entry);
 *       - pow / sin are the C library's powf / sinf; a fragment whose uv leaves [0, 1] (curved mode) is black
 *         (crt.kage:70-72 multiplies by 0; pow of a negative vignette is NaN on some GPUs, 0 here);
 *       - output = round-to-nearest of clamp(col, 0, 1) * 255, alpha 255 (vec4(col, 1), source-over).
 * Floating point: the GPU kernel and this file run the same float32 operation sequence without contraction;
 * they differ only in powf / sinf, so the tests allow <= 1 LSB on a small share of channel values.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pm3d_oracle.h"

/* floor((X + 0.5) * scale) for X in [0, n_out), -1 when the texel lies outside [0, n_src) */
void orc_present_axis_map(int32_t n_out, int32_t n_src, double render_scale, int32_t *map) {
    for (int32_t i = 0; i < n_out; i++) {
        const double s = floor(((double)i + 0.5) * render_scale);
        map[i] = (s >= 0 && s < (double)n_src) ? (int32_t)s : -1;
    }
}

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* imageSrc0At(uv * size + origin) over the scaled image; channel c of the premultiplied texel, in [0, 1] */
static inline float sample(const uint8_t *src, int32_t sw, const int32_t *xmap, const int32_t *ymap, int32_t ow,
                           int32_t oh, float u, float v, int c) {
    const float px = u * (float)ow, py = v * (float)oh;
    if (!(px >= 0.0f && px < (float)ow && py >= 0.0f && py < (float)oh)) return 0.0f;
    const int32_t ix = (int32_t)floorf(px), iy = (int32_t)floorf(py);
    const int32_t sx = xmap[ix], sy = ymap[iy];
    if (sx < 0 || sy < 0) return 0.0f;
    return (float)src[((size_t)sy * sw + sx) * 4 + c] / 255.0f;
}

static inline uint8_t to_u8(float v) {
    v = clampf(v, 0.0f, 1.0f);
    return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
}

/* mode 0: copy of the scaled image; 1: CRT shader, ShowCurve = 0; 2: CRT shader, ShowCurve = 1 */
void orc_present(const uint8_t *src, int32_t sw, int32_t sh, int32_t ow, int32_t oh, double render_scale, int32_t mode,
                 uint8_t *out, int32_t threads) {
    int32_t *xmap = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ow + oh));
    int32_t *ymap = xmap + ow;
    orc_present_axis_map(ow, sw, render_scale, xmap);
    orc_present_axis_map(oh, sh, render_scale, ymap);
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int32_t Y = 0; Y < oh; Y++) {
        for (int32_t X = 0; X < ow; X++) {
            uint8_t *o = out + ((size_t)Y * ow + X) * 4;
            if (mode == 0) {
                const int32_t sx = xmap[X], sy = ymap[Y];
                if (sx < 0 || sy < 0) { o[0] = o[1] = o[2] = o[3] = 0; }
                else memcpy(o, src + ((size_t)sy * sw + sx) * 4, 4);
                continue;
            }
            /* crt.kage:30-33: q := (src - origin) / size */
            const float fx = (float)X + 0.5f, fy = (float)Y + 0.5f;
            float ux = fx / (float)ow, uy = fy / (float)oh;
            if (mode == 2) { /* crt.kage:18-27 curve() */
                ux = (ux - 0.5f) * 2.0f; uy = (uy - 0.5f) * 2.0f;
                ux *= 1.1f; uy *= 1.1f;
                ux *= 1.0f + powf(fabsf(uy) / 5.0f, 2.0f);
                uy *= 1.0f + powf(fabsf(ux) / 4.0f, 2.0f);
                ux = ux / 2.0f + 0.5f; uy = uy / 2.0f + 0.5f;
                ux = ux * 0.92f + 0.04f; uy = uy * 0.92f + 0.04f;
            }
            float col[3];
            /* crt.kage:41-48 main colour + bleed */
            col[0] = sample(src, sw, xmap, ymap, ow, oh, ux + 0.0002f, uy + 0.0002f, 0) + 0.05f;
            col[1] = sample(src, sw, xmap, ymap, ow, oh, ux + 0.0000f, uy - 0.0005f, 1) + 0.05f;
            col[2] = sample(src, sw, xmap, ymap, ow, oh, ux - 0.0005f, uy + 0.0000f, 2) + 0.05f;
            col[0] += 0.08f * sample(src, sw, xmap, ymap, ow, oh, 0.01875f + (ux + 0.0002f), -0.02025f + (uy + 0.0002f), 0);
            col[1] += 0.05f * sample(src, sw, xmap, ymap, ow, oh, -0.0165f + (ux + 0.0000f), -0.015f + (uy - 0.0005f), 1);
            col[2] += 0.08f * sample(src, sw, xmap, ymap, ow, oh, -0.015f + (ux - 0.0005f), -0.0135f + (uy + 0.0000f), 2);
            const int outside = ux < 0.0f || ux > 1.0f || uy < 0.0f || uy > 1.0f;   /* crt.kage:70-72 */
            /* crt.kage:51 level adjustment, :54-57 vignette, :60-62 scanlines, :74 column mask */
            const float vig = 16.0f * ux * uy * (1.0f - ux) * (1.0f - uy);
            const float vp = outside ? 0.0f : powf(vig, 0.3f);
            const float scans = clampf(0.35f + 0.35f * sinf(uy * (float)oh * 1.5f), 0.0f, 1.0f);
            const float s = powf(scans, 1.7f);
            const float mask = 1.0f - 0.65f * clampf((fmodf(fx, 2.0f) - 1.0f) * 2.0f, 0.0f, 1.0f);
            static const float chan[3] = {0.95f, 1.05f, 0.95f};
            for (int c = 0; c < 3; c++) {
                float v = clampf(col[c] * 0.6f + 0.4f * col[c] * col[c], 0.0f, 1.0f);
                v *= vp;
                v *= chan[c];
                v *= 2.8f;
                v *= 0.4f + 0.7f * s;
                if (outside) v *= 0.0f;
                v *= mask;
                o[c] = to_u8(v);
            }
            o[3] = 255;
        }
    }
    free(xmap);
}
