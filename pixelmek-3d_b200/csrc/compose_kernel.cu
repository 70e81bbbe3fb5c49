// compose_kernel.cu — K2 floor cast + K3 Camera.Draw: the per-pixel pass that writes the
// RGBA8 framebuffer (call sites game/scene_game.go:161,167).
//
// Mapping: a warp owns one 64-column group x WARP_ROWS consecutive rows and walks them two rows
// at a time (16 lanes x 4 adjacent pixels per row, so every lane ends a row with ONE 128-bit
// store and a warp store covers two 256 B row segments).  Warps are independent: nothing is
// staged per CTA and there is no __syncthreads — the per-column records are read straight from
// the struct-of-arrays the cast wrote (a lane's four columns are one contiguous 32/64 B load,
// L1-resident after the first row pair), and the only shared memory is a warp-private list of
// the sprites that touch the warp's rows.
// Before any of that exists, a warp no sprite touches writes its sky-only and floor-only row pairs
// in two tight loops with their own mapping (compose_runs: 2 columns x 2 rows per lane, per-column
// parameters in registers, row records by shuffle) and returns if nothing is left; compose_warp is
// the general path for what remains.  Batches of views issue their row blocks horizon-first across
// ALL views (pm3d_launch_compose): the horizon blocks are the expensive ones.
//
// Layer order is the reference's painter's order (DESIGN.md "frame composition"):
//   sky / floor box -> walls (top level .. 0) -> textured floor -> sprites far .. near.
// Every layer is premultiplied source-over with one rounding per layer, so the colour after
// the last OPAQUE texel does not depend on anything behind it.  The background (box, walls,
// floor) is therefore resolved FRONT TO BACK and stops at the first opaque texel:
//   * rows at or below every floor start of the group (agg.fs_max, from the cast) evaluate the
//     four floor casts of a lane in straight line (4 independent binary64 chains);
//   * rows above every wall slice and floor start of the group (agg) are sky only;
//   * other rows resolve floor, then walls level by level for the four pixels together;
//   * a pixel that meets a translucent texel (0 < a < 255) is redone by the literal painter loop.
// Sprites are then blended over the background far -> near, exactly as the painter does.  Each
// lane keeps a private bitmask of the warp's sprites that overlap ITS four columns, so lanes walk
// different sprites concurrently (sprites pile up on the horizon rows).
// PM3D_OPT_FORCE_PAINTER forces the literal loop everywhere; tests require identical bytes.
// The warp that owns the screen-centre pixel also picks the convergence sprite
// (Camera.GetConvergenceSprite, game/scene_game.go:164).
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include <vector>

#include "pm3d_internal.cuh"

#ifndef COMPOSE_RUN_MIN_ROWS
#define COMPOSE_RUN_MIN_ROWS 2  // shapes with at least this many rows per warp take sky / floor runs: all of them (a 2-row warp is one row pair)
#endif
#ifndef COMPOSE_TP_LIST
#define COMPOSE_TP_LIST 256  // warp-private sprite list of the throughput shapes (entries)
#endif
#ifndef COMPOSE_TP_LATENCY_LIST
#define COMPOSE_TP_LATENCY_LIST 32  // sprite list of the >= 8-row shapes in latency mode (records in shared memory): 3.5 KB per warp
#endif
#ifndef COMPOSE_TP_LATENCY
#define COMPOSE_TP_LATENCY 1        // batches with sprites: the >= 8-row shapes walk sprites in latency mode
#endif
#ifndef COMPOSE_TP_WARPS
#define COMPOSE_TP_WARPS 1   // warps per CTA of the 16-row shape (large batches): a slot is not held for a slower sibling warp
#endif
#ifndef COMPOSE_SKY_NP
#define COMPOSE_SKY_NP 1    // row pairs per step of a sky run
#endif
#ifndef COMPOSE_FLOOR_NP
#define COMPOSE_FLOOR_NP 1  // row pairs per step of a floor run
#endif

// rows per warp (WARP_ROWS, a template parameter: 2, 4, 8 or 16 = 1 .. 8 iterations of 2 rows) trades the
// per-warp prologue against the length of the serial chain: single frames want many short warps, large
// batches of views want few long ones.
// Warp-private sprite list.  Throughput mode (batches without sprites to speak of): up to 256 entries, the draw records
// stay in global memory.  Latency mode (frames and batches with sprites): up to LATENCY_LIST (2/4-row shapes) or
// COMPOSE_TP_LATENCY_LIST (8/16-row shapes) entries whose record prefix (everything
// sprite_index reads: the first 80 bytes of a DrawRec) is copied to shared memory while the list is built, so the
// lanes' front-to-back walks — chains of dependent loads on the rows where sprites pile up, the slowest warps of the
// launch — read their records with an LDS instead of an L2 round trip.
#define DRAW_PREFIX_VEC 5
static_assert(offsetof(DrawRec, ty_hi) + 2 <= DRAW_PREFIX_VEC * 16, "sprite_index fields must sit in the copied prefix");

struct ComposeParams {
    const pm3d_view *views;
    const int2 *col_span;
    const uint2 *col_tt;
    const ulonglong2 *col_ps;
    const double *flo_xw, *flo_yw, *flo_dw, *flo_rc, *flo_dx, *flo_dy;
    const int32_t *flo_fs;
    const int4 *agg;
    const double *zbuf;
    const RowRec *row_rec;
    const DrawRec *draw_rec;
    const int32_t *sprite_offset;
    const uint16_t *bin_list;
    const int4 *bin_rect;
    const int4 *bin_count;
    const uint32_t *atlas;
    const int16_t *floor_tiles;
    const uint32_t *sheet_pixels;
    const pm3d_sprite *sprites;
    pm3d_convergence *conv;
    uint8_t *out;
    uint64_t xstep;  // span_step(tex_size, w): box texel column of screen column x
    int w, h, num_levels, tex_size, tex_shift, map_w, map_h, n_groups, sky_tile, floorbox_tile, force_painter;
    int row_major;  // grid = (groups, views, row blocks) instead of (groups, row blocks, views)
#ifdef PM3D_COMPOSE_STATS
    unsigned long long *task_log;  // tools builds: per warp {start ns, end ns, SM, first row} (tools/exp_tasklog.py)
#endif
};

// texel (tx, ty) of atlas tile `tile`
__device__ __forceinline__ uint32_t atlas_fetch(const ComposeParams &p, uint32_t tile, uint32_t ty, uint32_t tx) {
    return __ldg(p.atlas + ((((tile << p.tex_shift) | ty) << p.tex_shift) | tx));  // 32-bit texel index (pm3d_upload_atlas checks the size)
}

// screen rect of a listed sprite, clamped to int16 (screens are <= 16384 wide/high, so clamping never
// changes a comparison against a screen coordinate)
struct __align__(8) Rect16 { short x0, x1, y0, y1; };
__device__ __forceinline__ Rect16 make_rect16(int4 r) {
    Rect16 o;
    o.x0 = (short)max(-32768, min(32767, r.x)); o.x1 = (short)max(-32768, min(32767, r.y));
    o.y0 = (short)max(-32768, min(32767, r.z)); o.y1 = (short)max(-32768, min(32767, r.w));
    return o;
}
template <int MAXN, bool RECS>
struct WarpList {
    uint4 rec[RECS ? MAXN : 1][DRAW_PREFIX_VEC];
    Rect16 rect[MAXN];  // dsx, dex, dsy, dey, paint order
    uint16_t id[MAXN];
};

// ---- one sprite layer at pixel (x, y); returns texel (premultiplied) or 0 when it does not cover
// (not inlined: it has ~16 call sites and the kernel is instruction-fetch sensitive)
__device__ __noinline__ uint32_t sprite_texel(const DrawRec &d, const uint32_t *__restrict__ sheet_pixels, int x,
                                                 int y, double z) {
    if (x < d.dsx || x >= d.dex || x <= 0 || y < d.dsy || y >= d.dey) return 0u;
    if (!(d.transformY < z)) return 0u;
    const uint32_t k = (uint32_t)(x - d.dsx);
    const uint32_t tx = (uint32_t)__umul64hi((uint64_t)k * d.tex_w, d.xmagic);
    if (tx < d.tx_lo || tx >= d.tx_hi) return 0u;
    uint32_t ty = (uint32_t)((d.ypos0 + (uint64_t)(int64_t)y * d.ystep2) >> 33);
    if (ty > d.tex_h - 1) ty = d.tex_h - 1;
    if (ty < d.ty_lo || ty >= d.ty_hi) return 0u;
    return __ldg(sheet_pixels + d.sheet_px + (uint64_t)ty * d.sheet_w + tx);
}

// index of the texel sprite `d` shows at pixel (x, y) inside sheet_pixels, or ~0 when it does not cover the pixel
// (inlined twice in the front-to-back walk so that the loads of two sprites are in flight together)
#define NO_TEXEL 0xffffffffffffffffull
#ifndef SPRITE_BATCH
#define SPRITE_BATCH 2   // sprites fetched together per step of the latency-mode walk in the 80-register shapes
#endif
__device__ __forceinline__ uint64_t sprite_index(const DrawRec &d, int x, int y, double z) {
    if (x < d.dsx || x >= d.dex || x <= 0 || y < d.dsy || y >= d.dey) return NO_TEXEL;
    if (!(d.transformY < z)) return NO_TEXEL;
    const uint32_t k = (uint32_t)(x - d.dsx);
    const uint32_t tx = (uint32_t)__umul64hi((uint64_t)k * d.tex_w, d.xmagic);
    if (tx < d.tx_lo || tx >= d.tx_hi) return NO_TEXEL;
    uint32_t ty = (uint32_t)((d.ypos0 + (uint64_t)(int64_t)y * d.ystep2) >> 33);
    if (ty > d.tex_h - 1) ty = d.tex_h - 1;
    if (ty < d.ty_lo || ty >= d.ty_hi) return NO_TEXEL;
    return d.sheet_px + (uint64_t)ty * d.sheet_w + tx;
}

__device__ __forceinline__ uint32_t floor_tint(uint32_t px, uint32_t t) {
    // pixel.R = uint8(float64(pixel.R) * float64(pixelSt.R) / 256); alpha unchanged
    const uint32_t rr = ((px & 255u) * (t & 255u)) >> 8;
    const uint32_t gg = (((px >> 8) & 255u) * ((t >> 8) & 255u)) >> 8;
    const uint32_t bb = (((px >> 16) & 255u) * ((t >> 16) & 255u)) >> 8;
    return rr | (gg << 8) | (bb << 16) | (px & 0xff000000u);
}

// ---- floor layer of one pixel given its column parameters: texel already tinted (>>8), 0 if none
// weight := currentDist / distWall, correctly rounded either way
__device__ __forceinline__ double floor_weight(const RowRec &r, double dw, double rc) {
    return ((r.info & 0x20000000u) && rc != 0.0) ? div_by_rcp(r.cd, dw, rc) : r.cd / dw;
}

__device__ __forceinline__ uint32_t floor_texel(const ComposeParams &p, double fxw, double fyw, double dw, double rc, int fs,
                                                const RowRec &r, double posx, double posy, int y) {
    if (y < fs || !(r.info & 0x80000000u)) return 0u;
    const double weight = floor_weight(r, dw, rc);
    const double fx = weight * fxw + (1.0 - weight) * posx;
    const double fy = weight * fyw + (1.0 - weight) * posy;
    if (!(fx >= 0.0) || !(fy >= 0.0)) return 0u;
    if (!(fx < (double)p.map_w) || !(fy < (double)p.map_h)) return 0u;  // int(fx) >= map_w  <=>  fx >= map_w
    const int tx = (int)(fx * (double)p.tex_size), ty = (int)(fy * (double)p.tex_size);
    const int ft = __ldg(p.floor_tiles + (tx >> p.tex_shift) * p.map_h + (ty >> p.tex_shift));
    if (ft < 0) return 0u;
    const int m = p.tex_size - 1;
    const uint32_t px = atlas_fetch(p, (uint32_t)ft, (uint32_t)(ty & m), (uint32_t)(tx & m));
    if ((px >> 24) == 0) return 0u;
    return floor_tint(px, r.tint);
}

__device__ __forceinline__ uint32_t wall_texel(const ComposeParams &p, uint32_t tile_texx, ulonglong2 ps, int y) {
    uint32_t ty = (uint32_t)((ps.x + (uint64_t)(int64_t)y * ps.y) >> 33);
    if (ty > (uint32_t)p.tex_size - 1) ty = p.tex_size - 1;
    const uint32_t tile = tile_texx >> 16, tx = tile_texx & 0xffffu;
    return atlas_fetch(p, tile, ty, tx);
}

__device__ __forceinline__ uint32_t box_texel(const ComposeParams &p, uint32_t info, uint32_t u) {
    const int tile = (info & 0x40000000u) ? p.sky_tile : p.floorbox_tile;
    if (tile < 0) return 0u;
    const uint32_t v = info & 0xffffu;
    return atlas_fetch(p, (uint32_t)tile, v, u);
}

// tinted opaque texel over anything: (c*tint + 127)/255, alpha 255
__device__ __forceinline__ uint32_t shade_opaque(uint32_t t, uint32_t tint) {
    if (tint == 0xffffffffu) return t;  // (c*255 + 127)/255 == c
    return blend_over(0u, t, tint);
}

// literal painter's loop over the BACKGROUND layers of pixel (x, y) — the definition
__device__ __forceinline__ uint32_t background_painter(const ComposeParams &p, int view, int x, int y, uint32_t u) {
    const pm3d_view &c = p.views[view];
    const uint32_t box_tint = (uint32_t)c.max_light[0] | ((uint32_t)c.max_light[1] << 8) |
                              ((uint32_t)c.max_light[2] << 16) | 0xff000000u;
    const RowRec r = p.row_rec[(size_t)view * p.h + y];
    uint32_t dst = 0u;
    dst = blend_over(dst, box_texel(p, r.info, u), box_tint);
    for (int l = p.num_levels - 1; l >= 0; l--) {
        const size_t i = ((size_t)view * p.num_levels + l) * p.w + x;
        const int2 span = p.col_span[i];
        if (y >= span.x && y < span.y) {
            const uint2 tt = p.col_tt[i];
            dst = blend_over(dst, wall_texel(p, tt.x, p.col_ps[i], y), tt.y);
        }
    }
    const size_t fi = (size_t)view * p.w + x;
    dst = blend_over(dst, floor_texel(p, p.flo_xw[fi], p.flo_yw[fi], p.flo_dw[fi], p.flo_rc[fi], p.flo_fs[fi], r, c.pos_x, c.pos_y, y),
                     0xffffffffu);
    return dst;
}

// four consecutive array elements of a lane's columns; 16 B vector loads when the row pitch keeps
// them aligned (w % 4 == 0) and all four columns exist
__device__ __forceinline__ void load4(const double *base, size_t i, bool vec, int nvalid, double out[4]) {
    if (vec) {
        const double2 a = *reinterpret_cast<const double2 *>(base + i), b = *reinterpret_cast<const double2 *>(base + i + 2);
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] = base[i + (j < nvalid ? j : 0)];
    }
}
__device__ __forceinline__ void load4(const int32_t *base, size_t i, bool vec, int nvalid, int out[4]) {
    if (vec) {
        const int4 a = *reinterpret_cast<const int4 *>(base + i);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] = base[i + (j < nvalid ? j : 0)];
    }
}
__device__ __forceinline__ void load4(const int2 *base, size_t i, bool vec, int nvalid, int2 out[4]) {
    if (vec) {
        const int4 a = *reinterpret_cast<const int4 *>(base + i), b = *reinterpret_cast<const int4 *>(base + i + 2);
        out[0] = make_int2(a.x, a.y); out[1] = make_int2(a.z, a.w); out[2] = make_int2(b.x, b.y); out[3] = make_int2(b.z, b.w);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] = base[i + (j < nvalid ? j : 0)];
    }
}

// Row blocks are issued from the horizon outwards: the rows around it carry the sprites and the mixed
// wall/floor/sky rows, i.e. the longest CTAs, and a single frame is only a few waves of CTAs — started
// last they would be the tail of the launch.  (n row blocks of block_rows rows; k = blockIdx.y)
__device__ __forceinline__ int compose_yblock(int h, int pitch_px, int block_rows, int n, int k) {
    int cb = (h / 2 + pitch_px) / block_rows;
    cb = cb < 0 ? 0 : (cb > n - 1 ? n - 1 : cb);
    const int lo = cb, hi = n - 1 - cb;                      // blocks above / below the horizon block
    const int m = lo < hi ? lo : hi;
    if (k <= 2 * m) return (k & 1) ? cb - ((k + 1) >> 1) : cb + (k >> 1);
    return lo < hi ? k : n - 1 - k;
}
// ---- RUNS: tight loops over row pairs that are sky only / floor only for a whole 64-column group (k_compose decides
// which; shapes with many rows per warp).  Inside a run a lane owns 2 adjacent columns x 2 rows per pair and NP pairs per
// step, so everything per column — the three fast-floor-cast parameters, the box texel columns — is a handful of
// registers loaded ONCE per run (the general loop re-reads six 16 B vectors per row pair), the row records come out of
// the warp's own registers by shuffle (lane l holds row l), the texel loads of all NP pairs of a step are in flight
// together, and none of the general loop's state (sprite list, masks, slice spans, per-pixel flags) is live.  The
// arithmetic is that of the two fast branches of the general loop, value for value; whatever a step cannot prove for
// ALL lanes (a nil floor tile, a translucent texel, a zero fraction in floor_fast_index, a row beyond the fast cast's
// weight bound) ends the run and the general loop takes the row pairs from there.  Both return the first pair not written.
template <int NP>
__device__ __forceinline__ bool sky_step(const ComposeParams &p, const RowRec &mine, uint2 *o, const uint32_t (&u)[2], uint32_t box_tint, int it) {
    uint32_t t[NP][2][2];
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t inf = __shfl_sync(0xffffffffu, mine.info, 2 * (it + q) + k);
            const int tile = (inf & 0x40000000u) ? p.sky_tile : p.floorbox_tile;
            ok = ok && tile >= 0;
#pragma unroll
            for (int j = 0; j < 2; j++) t[q][k][j] = atlas_fetch(p, (uint32_t)max(tile, 0), inf & 0xffffu, u[j]);
        }
    uint32_t all = 0xffffffffu;
#pragma unroll
    for (int q = 0; q < NP; q++) all &= t[q][0][0] & t[q][0][1] & t[q][1][0] & t[q][1][1];
    ok = ok && all >= 0xff000000u;
    if (!__all_sync(0xffffffffu, ok)) return false;
    const int opitch = p.w >> 1;
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int k = 0; k < 2; k++)
            o[(size_t)(2 * (it + q) + k) * opitch] = make_uint2(shade_opaque(t[q][k][0], box_tint), shade_opaque(t[q][k][1], box_tint));
    return true;
}
template <int N_IT>
__device__ __forceinline__ int run_sky(const ComposeParams &p, const RowRec &mine, uint2 *o, int xa, uint32_t box_tint, int end) {
    const uint32_t tm = (uint32_t)p.tex_size - 1u;
    uint32_t u[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t uu = (uint32_t)((((uint64_t)(2 * (int64_t)(xa + j) + 1)) * p.xstep) >> 33);
        u[j] = uu > tm ? tm : uu;
    }
    int it = 0;
    if (N_IT >= 2 && COMPOSE_SKY_NP >= 2) {
#pragma unroll 1
        for (; it + 2 <= end; it += 2)
            if (!sky_step<2>(p, mine, o, u, box_tint, it)) break;
    }
#pragma unroll 1
    for (; it < end; it++)
        if (!sky_step<1>(p, mine, o, u, box_tint, it)) break;
    return it;
}

struct FloorRun {
    double2 cr, dx, dy;
    double cd_max, ptx, pty;
    uint32_t bx, by, tm;
};
template <int NP>
__device__ __forceinline__ bool floor_step(const ComposeParams &p, const FloorRun &f, const RowRec &mine, uint2 *o, int it) {
    uint32_t tx[NP][2][2], ty[NP][2][2];
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int row = 2 * (it + q) + k;
            const double cd = __shfl_sync(0xffffffffu, mine.cd, row);
            const uint32_t inf = __shfl_sync(0xffffffffu, mine.info, row);
            ok = ok && (inf & 0x80000000u) && (inf & 0x20000000u) && cd <= f.cd_max;
            const double w0 = cd * f.cr.x, w1 = cd * f.cr.y;
            ok = floor_fast_index(w0, f.dx.x, f.ptx, f.bx, tx[q][k][0]) && ok;
            ok = floor_fast_index(w0, f.dy.x, f.pty, f.by, ty[q][k][0]) && ok;
            ok = floor_fast_index(w1, f.dx.y, f.ptx, f.bx, tx[q][k][1]) && ok;
            ok = floor_fast_index(w1, f.dy.y, f.pty, f.by, ty[q][k][1]) && ok;
        }
    uint32_t t[NP][2][2];
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int k = 0; k < 2; k++) t[q][k][0] = t[q][k][1] = 0u;
    if (ok) {
        int ft[NP][2][2], any = 0;
#pragma unroll
        for (int q = 0; q < NP; q++)
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    ft[q][k][j] = __ldg(p.floor_tiles + (tx[q][k][j] >> p.tex_shift) * p.map_h + (ty[q][k][j] >> p.tex_shift));
                    any |= ft[q][k][j];
                }
        if (any >= 0) {
#pragma unroll
            for (int q = 0; q < NP; q++)
#pragma unroll
                for (int k = 0; k < 2; k++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        t[q][k][j] = atlas_fetch(p, (uint32_t)ft[q][k][j], ty[q][k][j] & f.tm, tx[q][k][j] & f.tm);
        }
    }
    uint32_t all = 0xffffffffu;  // (the texels are all zero when a step above failed)
#pragma unroll
    for (int q = 0; q < NP; q++) all &= t[q][0][0] & t[q][0][1] & t[q][1][0] & t[q][1][1];
    if (!__all_sync(0xffffffffu, all >= 0xff000000u)) return false;
    const int opitch = p.w >> 1;
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t tint = __shfl_sync(0xffffffffu, mine.tint, 2 * (it + q) + k);
            o[(size_t)(2 * (it + q) + k) * opitch] = make_uint2(floor_tint(t[q][k][0], tint), floor_tint(t[q][k][1], tint));
        }
    return true;
}
template <int N_IT>
__device__ __forceinline__ int run_floor(const ComposeParams &p, const pm3d_view &c, const RowRec &mine, uint2 *o, size_t fa, int it, int end) {
    FloorRun f;
    const double2 cw = *reinterpret_cast<const double2 *>(p.flo_dw + fa);
    f.cr = *reinterpret_cast<const double2 *>(p.flo_rc + fa);
    f.dx = *reinterpret_cast<const double2 *>(p.flo_dx + fa);
    f.dy = *reinterpret_cast<const double2 *>(p.flo_dy + fa);
    // largest row distance for which both columns keep |weight| <= 4 (negative: a column is not eligible)
    f.cd_max = (f.cr.x != 0.0 && f.cr.y != 0.0) ? 4.0 * fmin(fabs(cw.x), fabs(cw.y)) : -1.0;
    const double texd = (double)p.tex_size;
    f.ptx = c.pos_x * texd; f.pty = c.pos_y * texd;  // texel-unit camera position
    f.bx = (uint32_t)min((uint64_t)p.map_w << p.tex_shift, (uint64_t)0x80000000u);
    f.by = (uint32_t)min((uint64_t)p.map_h << p.tex_shift, (uint64_t)0x80000000u);
    f.tm = (uint32_t)p.tex_size - 1u;
    if (N_IT >= 2 && COMPOSE_FLOOR_NP >= 2) {
#pragma unroll 1
        for (; it + 2 <= end; it += 2)
            if (!floor_step<2>(p, f, mine, o, it)) break;
    }
#pragma unroll 1
    for (; it < end; it++)
        if (!floor_step<1>(p, f, mine, o, it)) break;
    return it;
}

// group aggregates written by the cast: rows >= fs_max are floor candidates in every column, rows < sky_below are
// above every wall slice and every floor start, no column of the group has a z-buffer value above zmax
__device__ __forceinline__ void compose_aggregates(const ComposeParams &p, int view, int group, int &fs_max, int &sky_below, double &zmax) {
    // every lane reads the same 2*L contiguous records (broadcast transactions at immediate offsets from one base:
    // cheaper than a strided read followed by ten shuffles); level 0 carries the floor starts and the z bound
    const int L = p.num_levels;
    const int4 *ag = p.agg + (size_t)((unsigned)(view * p.n_groups + group) * (unsigned)(2 * L));
    const int4 a0 = __ldg(ag), a1 = __ldg(ag + 1);
    fs_max = max(a0.z, a1.z);
    sky_below = min(min(a0.x, a0.y), min(a1.x, a1.y));
    zmax = (double)fmaxf(__int_as_float(a0.w), __int_as_float(a1.w));
    if (L <= 4) {
        // the usual stacks (the bundled maps have 3 levels): no loop; a slot past the last record re-reads the last
        // one (min / max do not mind), so there is nothing to predicate
        const int last = 2 * L - 1;
#pragma unroll
        for (int k = 2; k < 8; k++) {
            const int4 a = __ldg(ag + min(k, last));
            sky_below = min(sky_below, min(a.x, a.y));
            fs_max = max(fs_max, a.z);
        }
    } else {
#pragma unroll 2
        for (int k = 2; k < 2 * L; k++) {
            const int4 a = __ldg(ag + k);
            sky_below = min(sky_below, min(a.x, a.y));
            fs_max = max(fs_max, a.z);
        }
    }
}

// The runs of one warp (whole warp, converged): returns the bit mask of the row pairs it wrote.  When no sprite of the
// group touches the warp's rows, the leading row pairs that are sky only (above every slice of the group) and the
// trailing ones that are floor only (at or below every floor start) are written by run_sky / run_floor.
template <int WARP_ROWS>
__device__ __forceinline__ uint32_t compose_runs(const ComposeParams &p, const pm3d_view &c, int view, int group, int lane, int wrow0, int wrow1,
                                                 int fs_max, int sky_below, uint32_t box_tint) {
    constexpr int N_IT = WARP_ROWS / 2;
    const int gx0 = group * GROUP_COLS;
    if (p.force_painter || gx0 + GROUP_COLS > p.w || (p.w & 3) != 0) return 0u;
    const int4 gi = p.bin_count[view * p.n_groups + group];  // entries, min dsy, max dey
    if (gi.x > 0 && gi.z > wrow0 && gi.y < wrow1) return 0u;
    const int xa = gx0 + 2 * lane;  // this lane's two columns inside a run
    uint2 *o = reinterpret_cast<uint2 *>(reinterpret_cast<uint32_t *>(p.out) + ((size_t)view * p.h + wrow0) * p.w + xa);
    const int last = p.h - 1 - wrow0;  // last existing row of the view, relative to the warp's first
    // lane l holds the record of the warp's row l; the loops fetch them by shuffle (no load in the loops)
    const RowRec mine = p.row_rec[(size_t)view * p.h + wrow0 + min(lane, last)];
    const int n_pairs = min(N_IT, (last + 1) >> 1);  // complete row pairs of this warp
    // sky run: row pairs whose rows are all above every slice and floor start of the group
    const int sky_pairs = min(n_pairs, max(min(sky_below, p.h) - wrow0, 0) >> 1);
    int it = 0;
    if (sky_pairs > 0) it = run_sky<N_IT>(p, mine, o, xa, box_tint, sky_pairs);
    uint32_t done = (1u << it) - 1u;  // the sky run wrote row pairs [0, it)
    // floor run: from the first row pair at or below every floor start of the group to the warp's last one
    const int it_floor = max(it, (max(fs_max - wrow0, 0) + 1) >> 1);
    if (it_floor < n_pairs) {
        it = run_floor<N_IT>(p, c, mine, o, (size_t)view * p.w + xa, it_floor, n_pairs);
        done |= ((1u << it) - 1u) & ~((1u << it_floor) - 1u);  // the floor run wrote row pairs [it_floor, it)
    }
    return done;
}
// every row pair written: nothing is left for this warp — unless it owns the screen centre (convergence)
__device__ __forceinline__ bool compose_runs_complete(const ComposeParams &p, uint32_t done, int gx0, int wrow0, int wrow1) {
    const int n_it = (wrow1 - wrow0 + 1) >> 1;
    const int cx = p.w / 2, cy = p.h / 2;
    return done == (1u << n_it) - 1u && !(cx >= gx0 && cx < gx0 + GROUP_COLS && cy >= wrow0 && cy < wrow1);
}

#ifndef LATENCY_LIST
#define LATENCY_LIST 64  // 24 KB of shared memory per 4-warp CTA: 8 CTAs per SM fit (96 entries cap it at 6)
#endif
// One warp's work: 64 columns x WARP_ROWS rows starting at row wrow0 of column group `group` of view `view`.  `wl`,
// `xm`, `hit` are the warp's private shared-memory scratch (sprite list, per-lane sprite masks, translucent-hit masks).
// Called by the whole warp; lanes may return at different points (the caller re-converges the warp).
template <int WARP_ROWS, bool LATENCY, int SB, int MAX_WARP_SPRITES>
__device__ __forceinline__ void compose_warp(const ComposeParams &p, WarpList<MAX_WARP_SPRITES, LATENCY> &wl,
                                             uint32_t (&xm)[MAX_WARP_SPRITES / 32][32], uint32_t (&hit)[4][LATENCY ? MAX_WARP_SPRITES / 32 : 1][32],
                                             const int group, const int view, const int wrow0, const int lane) {
    static_assert(MAX_WARP_SPRITES % 32 == 0, "the per-lane masks are whole 32-bit words");
    constexpr int WARP_WORDS = MAX_WARP_SPRITES / 32;
    const pm3d_view &c = p.views[view];
    const int wrow1 = min(wrow0 + WARP_ROWS, p.h);
    const int gx0 = group * GROUP_COLS;
    const uint32_t box_tint = (uint32_t)c.max_light[0] | ((uint32_t)c.max_light[1] << 8) |
                              ((uint32_t)c.max_light[2] << 16) | 0xff000000u;

    // ---- group aggregates written by the cast: rows >= fs_max are floor candidates in every column,
    //      rows < sky_below are above every wall slice and every floor start
    int fs_max, sky_below;
    double zmax_group;  // no column of the group has a z-buffer value above this
    compose_aggregates(p, view, group, fs_max, sky_below, zmax_group);

    // ---- RUNS (shapes with many rows per warp), before anything else of the warp exists: the sky-only and floor-only
    //      row pairs of a warp no sprite touches are written by tight loops (compose_runs).  A warp whose rows are all
    //      done returns: no sprite list, no masks, no per-column records; the general loop below takes the row pairs no
    //      run has written (bit mask run_done).
    constexpr bool RUNS = WARP_ROWS >= COMPOSE_RUN_MIN_ROWS;
    uint32_t run_done = 0u;
    if (RUNS) {
        run_done = compose_runs<WARP_ROWS>(p, c, view, group, lane, wrow0, wrow1, fs_max, sky_below, box_tint);
        if (compose_runs_complete(p, run_done, gx0, wrow0, wrow1)) return;
    }

    // ---- warp-private sprite list: entries of the group list whose rows intersect this warp's rows
    const int s0 = p.sprite_offset[view];
    const int nv = p.sprite_offset[view + 1] - s0;
    const DrawRec *draw = p.draw_rec + s0;
    const uint16_t *glist = p.bin_list + (size_t)s0 * p.n_groups + (size_t)group * nv;
    const int4 *grect = p.bin_rect + (size_t)s0 * p.n_groups + (size_t)group * nv;
    const int4 ginfo = p.bin_count[view * p.n_groups + group];  // entries, min dsy, max dey
    const int gcnt = ginfo.x;
    // The list keeps the NEAREST MAX_WARP_SPRITES entries (the group list is in paint order, far -> near).  When more
    // qualify, the farther ones stay behind in the group list as entries [0, far_cut): the front-to-back walk only
    // reaches them for a pixel that sees through every listed sprite, and the repaint takes them first.
    int n_list = 0, far_cut = 0;
    bool overflow = false;
    if (gcnt > 0 && ginfo.z > wrow0 && ginfo.y < wrow1) {
        auto qualifies = [&](int i, uint16_t &id, int4 &rc) -> bool {
            if (i >= gcnt) return false;
            id = glist[i];
            rc = grect[i];  // dsx, dex, dsy, dey
            // rows of this warp, and z-cull: a sprite at or behind the farthest wall hit of the group's columns fails
            // `transformY < zBuffer[x]` in every one of them (camera facing a wall: whole stacks of sprites behind it
            // drop out here)
            return rc.w > wrow0 && rc.z < wrow1 && __ldg(&draw[id].transformY) < zmax_group;
        };
        int skip = 0;
        if (gcnt > MAX_WARP_SPRITES) {  // only then can the list overflow: count first
            int n_total = 0;
            for (int base = 0; base < gcnt; base += 32) {
                uint16_t id; int4 rc;
                n_total += __popc(__ballot_sync(0xffffffffu, qualifies(base + lane, id, rc)));
            }
            skip = max(0, n_total - MAX_WARP_SPRITES);
            overflow = skip > 0;
        }
        int rank = 0;
        for (int base = 0; base < gcnt; base += 32) {
            const int i = base + lane;
            uint16_t id = 0;
            int4 rc = make_int4(0, 0, 0, 0);
            const bool keep = qualifies(i, id, rc);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int r = rank + __popc(m & ((1u << lane) - 1u));
            if (overflow) {
                const unsigned first = __ballot_sync(0xffffffffu, keep && r == skip);
                if (first) far_cut = __shfl_sync(0xffffffffu, i, __ffs(first) - 1);
            }
            if (keep && r >= skip) {
                const int pos = r - skip;
                wl.id[pos] = id; wl.rect[pos] = make_rect16(rc);
                if (LATENCY) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(draw + id);
#pragma unroll
                    for (int q = 0; q < DRAW_PREFIX_VEC; q++) wl.rec[pos][q] = __ldg(src + q);
                }
            }
            rank += __popc(m);
        }
        n_list = rank - skip;
        __syncwarp();
    }

    // draw record of list entry i as far as sprite_index / tint need it
    auto list_rec = [&](int i) -> const DrawRec & {
        return LATENCY ? *reinterpret_cast<const DrawRec *>(&wl.rec[i][0]) : draw[wl.id[i]];
    };

    // ---- convergence (one warp per view): wall hit of the centre ray, overridden by the nearest
    //      focusable sprite covering the centre pixel in front of that wall; ties -> lower list index
    {
        const int cx = p.w / 2, cy = p.h / 2;
        if (cx >= gx0 && cx < gx0 + GROUP_COLS && cy >= wrow0 && cy < wrow1) {
            const double zc = p.zbuf[(size_t)view * p.w + cx];
            int best = -1;
            double best_d = 0.0;
            const int cnt = overflow ? gcnt : n_list;
            for (int i = lane; i < cnt; i += 32) {
                const DrawRec &d = draw[overflow ? glist[i] : wl.id[i]];
                if ((d.flags & 1u) && cx > 0 && cx >= d.dsx && cx < d.dex && cy >= d.dsy && cy < d.dey && d.transformY < zc) {
                    if (best < 0 || d.dist < best_d || (d.dist == best_d && d.src < best)) { best = d.src; best_d = d.dist; }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                const double od = __shfl_xor_sync(0xffffffffu, best_d, o);
                if (ob >= 0 && (best < 0 || od < best_d || (od == best_d && ob < best))) { best = ob; best_d = od; }
            }
            if (lane == 0) {
                pm3d_convergence cv;
                const double cameraX = 2.0 * (double)cx / (double)p.w - 1.0;
                const double rdx = c.dir_x + c.plane_x * cameraX, rdy = c.dir_y + c.plane_y * cameraX;
                cv.distance = zc;
                cv.point_x = c.pos_x + zc * rdx;
                cv.point_y = c.pos_y + zc * rdy;
                cv.sprite = -1;
                cv._pad = 0;
                if (best >= 0) {
                    const pm3d_sprite &sp = p.sprites[s0 + best];
                    cv.sprite = best;
                    cv.distance = best_d;
                    cv.point_x = sp.x;
                    cv.point_y = sp.y;
                }
                p.conv[view] = cv;
            }
        }
    }

    // ---- per-lane mask of the warp's sprites that overlap this lane's four columns (lanes then walk different
    //      sprites concurrently instead of every lane scanning the whole list).  Built transposed: lane i turns sprite
    //      i's column range into a 16-bit mask of the column quads it touches, and sixteen ballots hand every lane its
    //      word — ~50 instructions per 32 sprites instead of a 32-step loop per lane (which was a third of a two-row
    //      warp's instructions).  Before the early return of off-screen lanes: the ballots want the whole warp.
    const int n_words = (n_list + 31) >> 5;
    {
        const int lx_ = lane & 15;
#pragma unroll 1
        for (int wd = 0; wd < n_words; wd++) {
            const int i = wd * 32 + lane;
            uint32_t q = 0u;
            if (i < n_list) {
                const Rect16 rc = wl.rect[i];
                // quad k covers columns [gx0 + 4k, gx0 + 4k + 3]; it overlaps [rc.x0, rc.x1) iff
                // 4k >= rc.x0 - gx0 - 3 and 4k <= rc.x1 - gx0 - 1
                const int a = (int)rc.x0 - gx0 - 3, b = (int)rc.x1 - gx0 - 1;
                const int lo = a <= 0 ? 0 : (a + 3) >> 2;
                const int hi = b < 0 ? -1 : min(15, b >> 2);
                if (lo <= hi) q = ((2u << hi) - 1u) & ~((1u << lo) - 1u);
            }
            uint32_t m = 0u;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const unsigned bal = __ballot_sync(0xffffffffu, (q >> k) & 1u);
                if (lx_ == k) m = bal;
            }
            xm[wd][lane] = m;
        }
    }

    // ---- this lane's four columns
    const int lx = lane & 15, ly = lane >> 4;
    const int x0 = gx0 + lx * 4;
    if (x0 >= p.w) return;
    const int nvalid = min(4, p.w - x0);
    const bool full4 = nvalid == 4;
    const bool vec = ((p.w & 3) == 0) && full4;
    const bool fast_ok = !p.force_painter && full4;
    const int tmask = p.tex_size - 1;
    const double texd = (double)p.tex_size;
    const size_t fi0 = (size_t)view * p.w + x0;
    uint32_t *out32 = reinterpret_cast<uint32_t *>(p.out) + (size_t)view * p.h * p.w;

    // fast floor cast: texel-unit camera position, map bounds in texels, and the largest row distance for which
    // every column of this lane keeps |weight| <= 4 (negative when a column is not eligible)
    // (the camera position is read where a floor row needs it: half the warps of a frame never do)
    const uint32_t boundx = (uint32_t)min((uint64_t)p.map_w << p.tex_shift, (uint64_t)0x80000000u);
    const uint32_t boundy = (uint32_t)min((uint64_t)p.map_h << p.tex_shift, (uint64_t)0x80000000u);
    double lane_cd_max = -2.0;  // -2: not computed yet (only warps that reach a floor row need it)

    // box texel column of each of this lane's pixels: once per warp when it walks many rows, else by the rows that
    // show the box (a two-row warp below the horizon never needs it)
    auto box_col = [&](int j) -> uint32_t {
        const uint32_t u = (uint32_t)((((uint64_t)(2 * (int64_t)(x0 + j) + 1)) * p.xstep) >> 33);
        return u > (uint32_t)tmask ? (uint32_t)tmask : u;
    };
    uint32_t bu[4] = {0u, 0u, 0u, 0u};
    if (WARP_ROWS >= 8) {
#pragma unroll
        for (int j = 0; j < 4; j++) bu[j] = box_col(j);
    }
    auto box_u = [&](int j) -> uint32_t { return WARP_ROWS >= 8 ? bu[j] : box_col(j); };

    double zb[4] = {0.0, 0.0, 0.0, 0.0};  // z-buffer of this lane's columns (rows share it)
    if (n_list > 0) load4(p.zbuf, fi0, vec, nvalid, zb);

#pragma unroll 1
    for (int it = 0; it < WARP_ROWS / 2; it++) {
        if (RUNS && ((run_done >> it) & 1u)) continue;  // this row pair was written by a run
        const int y = wrow0 + it * 2 + ly;
        if (y >= p.h) break;
        const RowRec r = p.row_rec[(size_t)view * p.h + y];
        uint32_t px[4];
        bool done = false;
        const unsigned full = (1u << nvalid) - 1u;

        // ---------- sprites, front to back: the nearest opaque texel of a pixel hides everything behind it
        unsigned s_got = 0u;    // pixels whose nearest opaque sprite texel is in s_top
        unsigned s_trans = 0u;  // pixels that met a translucent sprite texel in front of it
        uint32_t s_top[4] = {0u, 0u, 0u, 0u};
        uint32_t s_stop = 0u;   // list index (one byte per pixel) of the sprite whose opaque texel is in s_top
        const bool sprite_f2b = n_list > 0 && !p.force_painter;
        unsigned s_far = 0u;    // pixels whose opaque texel was found among the far entries (not in the list)
        if (sprite_f2b && LATENCY) {
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int wq = 0; wq < WARP_WORDS; wq++) hit[j][wq][lane] = 0u;
            // Latency mode (few rows per warp, single frames).  SB sprites per step: a lane's walk is a chain of
            // dependent loads (record -> texel), so the records and texels of the next few candidates are fetched
            // together and then resolved nearest first.
            int wd = n_words - 1;
            uint32_t m = xm[wd][lane];
            auto next_sprite = [&]() -> int {  // nearest remaining sprite of this lane that crosses row y, or -1
                for (;;) {
                    while (m == 0u) {
                        if (--wd < 0) return -1;
                        m = xm[wd][lane];
                    }
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int i = wd * 32 + bit;
                    const Rect16 rc = wl.rect[i];
                    if (y >= rc.y0 && y < rc.y1) return i;
                }
            };
#pragma unroll 1
            while (s_got != full) {
                int ids[SB];
#pragma unroll
                for (int b = 0; b < SB; b++) ids[b] = (b == 0 || ids[b - 1] >= 0) ? next_sprite() : -1;
                if (ids[0] < 0) break;
                const DrawRec *dr[SB];
#pragma unroll
                for (int b = 0; b < SB; b++) dr[b] = &list_rec(ids[b] < 0 ? ids[0] : ids[b]);
                uint32_t tx[SB][4];
#pragma unroll
                for (int b = 0; b < SB; b++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const bool need = j < nvalid && !((s_got >> j) & 1u) && ids[b] >= 0;
                        const uint64_t a = need ? sprite_index(*dr[b], x0 + j, y, zb[j]) : NO_TEXEL;
                        tx[b][j] = a != NO_TEXEL ? __ldg(p.sheet_pixels + a) : 0u;
                    }
                }
#pragma unroll
                for (int b = 0; b < SB; b++) {  // resolve nearest first
                    const uint32_t tint = dr[b]->tint;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((s_got >> j) & 1u) continue;
                        const uint32_t a = tx[b][j] >> 24;
                        if (a == 255u) { s_top[j] = blend_over(0u, tx[b][j], tint); s_got |= 1u << j; s_stop |= (uint32_t)ids[b] << (8 * j); }
                        else if (a != 0u) { s_trans |= 1u << j; hit[j][ids[b] >> 5][lane] |= 1u << (ids[b] & 31); }
                    }
                }
            }
        }
        if (sprite_f2b && !LATENCY) {
            // throughput mode (large batches of views): one sprite at a time, no speculative fetches
#pragma unroll 1
            for (int wd = n_words - 1; wd >= 0; wd--) {
                uint32_t m = xm[wd][lane];
                while (m && s_got != full) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int i = wd * 32 + bit;
                    const Rect16 rc = wl.rect[i];
                    if (y < rc.y0 || y >= rc.y1) continue;
                    const DrawRec &d = list_rec(i);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (j >= nvalid || ((s_got >> j) & 1u)) continue;
                        const uint32_t t = sprite_texel(d, p.sheet_pixels, x0 + j, y, zb[j]);
                        const uint32_t a = t >> 24;
                        if (a == 255u) { s_top[j] = blend_over(0u, t, d.tint); s_got |= 1u << j; s_stop |= (uint32_t)i << (8 * j); }
                        else if (a != 0u) s_trans |= 1u << j;
                    }
                }
            }
        }
        if (overflow && !p.force_painter) {
            // pixels that saw through every listed sprite go on with the far entries of the group list, nearest first
            for (int i = far_cut - 1; i >= 0 && s_got != full; i--) {
                const int4 rc = grect[i];
                if (y < rc.z || y >= rc.w || x0 + 3 < rc.x || x0 >= rc.y) continue;
                const DrawRec &d = draw[glist[i]];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j >= nvalid || ((s_got >> j) & 1u)) continue;
                    const uint32_t t = sprite_texel(d, p.sheet_pixels, x0 + j, y, zb[j]);
                    const uint32_t a = t >> 24;
                    if (a == 255u) { s_top[j] = blend_over(0u, t, d.tint); s_got |= 1u << j; s_far |= 1u << j; }
                    else if (a != 0u) s_trans |= 1u << j;
                }
            }
        }
        // a pixel with an opaque sprite texel needs no background: whatever is translucent in front of that texel is
        // blended over it below (when the list overflowed the repaint starts from the background instead)
        // (a translucent pixel whose opaque texel is a far entry is repainted from the background like one without any)
        const unsigned s_base = (sprite_f2b && LATENCY) ? (s_got & ~(s_far & s_trans)) : (s_got & ~s_trans);
        const bool need_bg = s_base != full;
        if (!need_bg) done = true;

        if (done) {
            // every pixel of this lane is covered by an opaque sprite texel
        } else if (fast_ok && y >= fs_max && (r.info & 0x80000000u)) {
            // ---------- every pixel of this lane is a floor candidate: four fast casts in straight line
            //            (pm3d_internal.cuh "fast floor cast"); anything doubtful goes to the mixed path below,
            //            which runs the reference sequence
            if (lane_cd_max == -2.0) {
                double cw[4], cr[4];
                load4(p.flo_dw, fi0, vec, 4, cw);
                load4(p.flo_rc, fi0, vec, 4, cr);
                lane_cd_max = -1.0;
                if (cr[0] != 0.0 && cr[1] != 0.0 && cr[2] != 0.0 && cr[3] != 0.0)
                    lane_cd_max = 4.0 * fmin(fmin(fabs(cw[0]), fabs(cw[1])), fmin(fabs(cw[2]), fabs(cw[3])));
            }
            if ((r.info & 0x20000000u) && r.cd <= lane_cd_max) {
                double cr[4], dx[4], dy[4];
                load4(p.flo_rc, fi0, vec, 4, cr);
                load4(p.flo_dx, fi0, vec, 4, dx);
                load4(p.flo_dy, fi0, vec, 4, dy);
                uint32_t tx[4], ty[4];
                bool ok = true;
                const double ptx = c.pos_x * texd, pty = c.pos_y * texd;  // texel-unit camera position
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double wq = r.cd * cr[j];
                    ok = floor_fast_index(wq, dx[j], ptx, boundx, tx[j]) && ok;
                    ok = floor_fast_index(wq, dy[j], pty, boundy, ty[j]) && ok;
                }
                if (ok) {
                    int ft[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        ft[j] = __ldg(p.floor_tiles + (tx[j] >> p.tex_shift) * p.map_h + (ty[j] >> p.tex_shift));
                    if ((ft[0] | ft[1] | ft[2] | ft[3]) >= 0) {
                        uint32_t t[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            t[j] = atlas_fetch(p, (uint32_t)ft[j], ty[j] & (uint32_t)tmask, tx[j] & (uint32_t)tmask);
                        if ((t[0] & t[1] & t[2] & t[3]) >= 0xff000000u) {
#pragma unroll
                            for (int j = 0; j < 4; j++) px[j] = floor_tint(t[j], r.tint);
                            done = true;
                        }
                    }
                }
            }
        } else if (fast_ok && y < sky_below) {
            // ---------- above every wall slice and floor start of the group: sky / floor box only
            uint32_t t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) t[j] = box_texel(p, r.info, box_u(j));
            if ((t[0] & t[1] & t[2] & t[3]) >= 0xff000000u) {
#pragma unroll
                for (int j = 0; j < 4; j++) px[j] = shade_opaque(t[j], box_tint);
                done = true;
            }
        }
        if (!done) {
            // ---------- mixed row: floor, then walls level by level, then the box — four pixels together
            unsigned pending = (1u << nvalid) - 1u;         // pixels not resolved yet
            unsigned slow = p.force_painter ? pending : 0u;  // pixels that met a translucent texel
            pending &= ~slow;
#pragma unroll
            for (int j = 0; j < 4; j++) px[j] = 0u;
            if (pending) {
                int fs[4];
                load4(p.flo_fs, fi0, vec, nvalid, fs);
                bool any_floor = false;
#pragma unroll
                for (int j = 0; j < 4; j++) any_floor = any_floor || (((pending >> j) & 1u) && y >= fs[j]);
                if (any_floor && (r.info & 0x80000000u)) {
                    double cw[4], cr[4], cx[4], cy[4];
                    load4(p.flo_dw, fi0, vec, nvalid, cw);
                    load4(p.flo_rc, fi0, vec, nvalid, cr);
                    load4(p.flo_xw, fi0, vec, nvalid, cx);
                    load4(p.flo_yw, fi0, vec, nvalid, cy);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (!((pending >> j) & 1u)) continue;
                        const uint32_t t = floor_texel(p, cx[j], cy[j], cw[j], cr[j], fs[j], r, c.pos_x, c.pos_y, y);
                        const uint32_t a = t >> 24;
                        if (a == 255u) { px[j] = t; pending &= ~(1u << j); }  // (t*255 + 127)/255 == t
                        else if (a != 0u) { slow |= 1u << j; pending &= ~(1u << j); }
                    }
                }
            }
            for (int l = 0; l < p.num_levels && pending; l++) {
                const size_t ci = ((size_t)view * p.num_levels + l) * p.w + x0;
                int2 span[4];
                load4(p.col_span, ci, vec, nvalid, span);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (!((pending >> j) & 1u) || y < span[j].x || y >= span[j].y) continue;
                    const uint2 tt = p.col_tt[ci + j];
                    const uint32_t t = wall_texel(p, tt.x, p.col_ps[ci + j], y);
                    const uint32_t a = t >> 24;
                    if (a == 255u) { px[j] = shade_opaque(t, tt.y); pending &= ~(1u << j); }
                    else if (a != 0u) { slow |= 1u << j; pending &= ~(1u << j); }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((pending >> j) & 1u) px[j] = blend_over(0u, box_texel(p, r.info, box_u(j)), box_tint);
            if (slow) {
                // one inlined copy of the painter (a call would force the kernel parameters into local memory)
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    if (!((slow >> j) & 1u)) continue;
                    const uint32_t u = box_u(j);
                    const uint32_t v = background_painter(p, view, x0 + j, y, u);
                    if (j == 0) px[0] = v; else if (j == 1) px[1] = v; else if (j == 2) px[2] = v; else px[3] = v;
                }
            }
        }

        // ---------- combine: opaque sprite texels win; a pixel that met translucent texels gets them blended far ->
        //            near over its nearest opaque layer (sprite texel or background) — the layers behind an opaque
        //            texel cannot show, so this is the painter's result.  When the list overflowed or the painter is
        //            forced, every sprite is blended over the background.
        // (a warp without sprites — most warps of a batch of views — skips all of it on one uniform branch)
        unsigned repaint = 0u;
        if (n_list > 0 || overflow) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((s_base >> j) & 1u) px[j] = s_top[j];
            repaint = p.force_painter ? (n_list > 0 ? full : 0u) : s_trans;
        }
        if (repaint) {
            if (overflow) {
                // the far entries first (paint order), for the pixels repainted from the background
                for (int i = 0; i < far_cut; i++) {
                    const int4 rc = grect[i];
                    if (y < rc.z || y >= rc.w || x0 + 3 < rc.x || x0 >= rc.y) continue;
                    const DrawRec &d = draw[glist[i]];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (!(((repaint & ~s_base) >> j) & 1u)) continue;
                        px[j] = blend_over(px[j], sprite_texel(d, p.sheet_pixels, x0 + j, y, zb[j]), d.tint);
                    }
                }
            }
            if (LATENCY && !p.force_painter) {
                // the translucent texels the walk met, far -> near (ascending list index), over the opaque base
#pragma unroll 1
                for (int wd = 0; wd < n_words; wd++) {
                    uint32_t hm[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) hm[j] = ((repaint >> j) & 1u) ? hit[j][wd][lane] : 0u;
                    uint32_t m = hm[0] | hm[1] | hm[2] | hm[3];
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1u;
                        const DrawRec &d = list_rec(wd * 32 + bit);
                        uint32_t t[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {  // the four texels are fetched together
                            const uint64_t a = ((hm[j] >> bit) & 1u) ? sprite_index(d, x0 + j, y, zb[j]) : NO_TEXEL;
                            t[j] = a != NO_TEXEL ? __ldg(p.sheet_pixels + a) : 0u;
                        }
                        const uint32_t tint = d.tint;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if ((hm[j] >> bit) & 1u) px[j] = blend_over(px[j], t[j], tint);
                    }
                }
            } else
#pragma unroll 1
            for (int wd = 0; wd < n_words; wd++) {
                uint32_t m = xm[wd][lane];
                while (m) {
                    const int i = wd * 32 + __ffs(m) - 1;
                    m &= m - 1u;
                    const Rect16 rc = wl.rect[i];
                    if (y < rc.y0 || y >= rc.y1) continue;
                    if (LATENCY) {
                        unsigned todo = 0u;  // pixels being repainted whose opaque base lies behind sprite i
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (((repaint >> j) & 1u) && (!((s_base >> j) & 1u) || i > (int)((s_stop >> (8 * j)) & 255u))) todo |= 1u << j;
                        if (!todo) continue;
                        const DrawRec &d = list_rec(i);
                        uint32_t t[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {  // the four texels are fetched together
                            const uint64_t a = ((todo >> j) & 1u) ? sprite_index(d, x0 + j, y, zb[j]) : NO_TEXEL;
                            t[j] = a != NO_TEXEL ? __ldg(p.sheet_pixels + a) : 0u;
                        }
                        const uint32_t tint = d.tint;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if ((todo >> j) & 1u) px[j] = blend_over(px[j], t[j], tint);
                    } else {
                        // throughput mode: every sprite over the background, small code
                        const DrawRec &d = list_rec(i);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (!((repaint >> j) & 1u)) continue;
                            px[j] = blend_over(px[j], sprite_texel(d, p.sheet_pixels, x0 + j, y, zb[j]), d.tint);
                        }
                    }
                }
            }
        }

        uint32_t *o = out32 + (size_t)y * p.w + x0;
        if (vec) {
            *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
            for (int j = 0; j < nvalid; j++) o[j] = px[j];
        }
    }
}

// COMPOSE_WARPS warps per CTA: 4 for single frames, 2 for large batches (a CTA's slot is held until its slowest warp
// ends, and small CTAs waste fewer warp slots); MIN_BLOCKS is given for 4-warp CTAs and scaled.
// LATENCY: single frames with sprites (see the sprite list above)
template <int WARP_ROWS, int MIN_BLOCKS, int COMPOSE_WARPS, bool LATENCY = false>
__global__ void __launch_bounds__(COMPOSE_WARPS * 32, MIN_BLOCKS * 4 / COMPOSE_WARPS) k_compose(const __grid_constant__ ComposeParams p) {
    constexpr int COMPOSE_ROWS = COMPOSE_WARPS * WARP_ROWS;
    constexpr int MAX_WARP_SPRITES = LATENCY ? (WARP_ROWS >= 8 ? COMPOSE_TP_LATENCY_LIST : LATENCY_LIST) : COMPOSE_TP_LIST;
    constexpr int WARP_WORDS = MAX_WARP_SPRITES / 32;
    // the 64-register / 8-CTA shapes (frames sharing the chip: PM3D_OPT_LANES > 1) walk one sprite per step: the second
    // sprite's records and texels in flight cost them spills, and the other lanes' warps hide the latency anyway
    // (c2, 4 lanes: 55.6 k -> 62.1 k frames/s; a frame that has the chip to itself is the same either way)
    constexpr int SB = (LATENCY && MIN_BLOCKS >= 8) ? 1 : SPRITE_BATCH;
    __shared__ WarpList<MAX_WARP_SPRITES, LATENCY> s_list[COMPOSE_WARPS];
    __shared__ uint32_t s_xm[COMPOSE_WARPS][WARP_WORDS][32];  // per-lane sprite masks (kept out of registers and out of unrolled code)
    // latency mode: per pixel, the listed sprites whose texel was TRANSLUCENT in the front-to-back walk (bit = list
    // index).  The repaint of such a pixel blends exactly these over its opaque base instead of testing every sprite
    // of the lane again (the bundled trees are mostly fully transparent texels around a few translucent ones)
    __shared__ uint32_t s_hit[LATENCY ? COMPOSE_WARPS : 1][4][LATENCY ? WARP_WORDS : 1][32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int view = p.row_major ? blockIdx.y : blockIdx.z;
    const int yblock = compose_yblock(p.h, p.views[view].pitch_px, COMPOSE_ROWS, p.row_major ? gridDim.z : gridDim.y, p.row_major ? blockIdx.z : blockIdx.y);
    const int wrow0 = yblock * COMPOSE_ROWS + warp * WARP_ROWS;
    if (wrow0 >= p.h) return;
#ifdef PM3D_COMPOSE_STATS
    unsigned long long t0 = 0;
    if (p.task_log) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
    compose_warp<WARP_ROWS, LATENCY, SB, MAX_WARP_SPRITES>(p, s_list[warp], s_xm[warp], s_hit[LATENCY ? warp : 0], blockIdx.x, view, wrow0, lane);
#ifdef PM3D_COMPOSE_STATS
    __syncwarp();
    if (p.task_log && lane == 0) {
        unsigned long long t1; unsigned int sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        unsigned long long *e = p.task_log + 4 * ((((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * COMPOSE_WARPS + warp);
        e[0] = t0; e[1] = t1; e[2] = sm; e[3] = (unsigned long long)wrow0;
    }
#endif
}

// ---- per-view FNV-1a style checksum (order-independent combine of per-word hashes: a sum mod 2^64, so the split
//      into chunks and the order of the atomic adds do not matter)
__device__ __forceinline__ uint64_t word_hash(uint64_t i, uint32_t v) {
    uint64_t hsh = 1469598103934665603ull ^ i;
    hsh *= 1099511628211ull;
    hsh ^= v;
    hsh *= 1099511628211ull;
    return hsh;
}
constexpr int CHECKSUM_CHUNK = 256 * 4 * 8;  // words per CTA: 8 uint4 per thread
__global__ void __launch_bounds__(256) k_checksum(const uint32_t *frames, uint64_t *out, size_t words_per_view) {
    const int view = blockIdx.y;
    const uint32_t *f = frames + (size_t)view * words_per_view;
    const size_t lo = (size_t)blockIdx.x * CHECKSUM_CHUNK, hi = min(lo + (size_t)CHECKSUM_CHUNK, words_per_view);
    uint64_t acc = 0;
    if ((((uintptr_t)f & 15u) == 0) && hi - lo == (size_t)CHECKSUM_CHUNK) {
        const uint4 *f4 = reinterpret_cast<const uint4 *>(f + lo);
#pragma unroll 4
        for (int k = 0; k < 8; k++) {
            const int q = k * 256 + threadIdx.x;
            const uint4 v = __ldg(f4 + q);
            const size_t i = lo + (size_t)q * 4;
            acc += word_hash(i, v.x) + word_hash(i + 1, v.y) + word_hash(i + 2, v.z) + word_hash(i + 3, v.w);
        }
    } else {
        for (size_t i = lo + threadIdx.x; i < hi; i += 256) acc += word_hash(i, f[i]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ uint64_t red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < 8; w++) t += red[w];
        atomicAdd(reinterpret_cast<unsigned long long *>(out + view), (unsigned long long)t);
    }
}

int pm3d_launch_compose(pm3d_ctx *ctx, int n_views, uint8_t *out) {
    ComposeParams p;
    p.views = ctx->cur_views;
    p.col_span = ctx->col_span; p.col_tt = ctx->col_tt; p.col_ps = ctx->col_ps;
    p.flo_xw = ctx->flo_xw; p.flo_yw = ctx->flo_yw; p.flo_dw = ctx->flo_dw; p.flo_rc = ctx->flo_rc; p.flo_dx = ctx->flo_dx; p.flo_dy = ctx->flo_dy; p.flo_fs = ctx->flo_fs;
    p.agg = ctx->agg;
    p.zbuf = ctx->zbuf;
    p.row_rec = ctx->row_rec;
    p.draw_rec = ctx->draw_rec;
    p.sprite_offset = ctx->cur_offset;
    p.bin_list = ctx->bin_list;
    p.bin_rect = ctx->bin_rect;
    p.bin_count = ctx->bin_count;
    p.atlas = reinterpret_cast<const uint32_t *>(ctx->atlas);
    p.floor_tiles = ctx->floor_tiles;
    p.sheet_pixels = reinterpret_cast<const uint32_t *>(ctx->sheet_pixels);
    p.sprites = ctx->cur_sprites;
    p.conv = ctx->conv;
    p.out = out;
    p.xstep = ((uint64_t)ctx->tex_size << 32) / (uint64_t)ctx->w;
    p.w = ctx->w; p.h = ctx->h; p.num_levels = ctx->num_levels; p.tex_size = ctx->tex_size; p.tex_shift = ctx->tex_shift;
    p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.n_groups = ctx->n_groups;
    p.sky_tile = ctx->sky_tile; p.floorbox_tile = ctx->floorbox_tile; p.force_painter = ctx->force_painter;
    p.row_major = 0;
    // rows per warp: enough warps to fill the chip a few times over, otherwise as long as possible
    const long long warps8 = (long long)ctx->n_groups * ((ctx->h + 7) / 8) * n_views;
    const long long resident = (long long)ctx->sm_count * 24;
    int rows = warps8 >= 64 * resident ? 16 : (warps8 >= 4 * resident ? 8 : (warps8 >= 2 * resident ? 4 : 2));  // 4K single frame: 8
    // with sprites a small batch (up to ~350 views of 320x200) is better off with the 4-row warps of the single-frame
    // shapes (more, shorter warps for the sprite-heavy horizon rows); from there on the 8-row shape, which walks sprites
    // in latency mode too (256 views: 4 rows 99 us, 8 rows 122 us; 512 views: 168 / 152 us; 1024 views: 304 / 256 us)
    if (ctx->cur_n_sprites > 0 && rows == 8 && warps8 < 12 * resident) rows = 4;
    // 64 registers / 8 CTAs per SM for the >= 8-row (large batch) shapes, and for single frames when several of them
    // share the chip (PM3D_OPT_LANES > 1): a few spills buy a third more resident warps, which is what hides the texel
    // latency at saturation (3-4 lanes: +12 % frames/s); a frame that has the chip to itself is 3 % faster with 80
    const bool shared_chip = ctx->n_lanes > 1 && ctx->n_sets > 1;
    int minb = (rows >= 8 || (rows == 2 && shared_chip)) ? 8 : 6;
    if (const char *e = getenv("PM3D_COMPOSE_ROWS")) {  // tuning overrides (tools/); anything but a built shape is ignored
        const int v = atoi(e);
        if (v == 2 || v == 4 || v == 8 || v == 16) rows = v;
    }
    if (const char *e = getenv("PM3D_COMPOSE_MINB")) {
        const int v = atoi(e);
        if (v == 6 || v == 8) minb = v;
    }
    const int cta_warps = rows >= 8 ? 2 : 4;
    const bool with_sprites = ctx->cur_n_sprites > 0;  // latency-mode sprite walk (records in shared memory)
    dim3 block(cta_warps * 32), grid(ctx->n_groups, (ctx->h + cta_warps * rows - 1) / (cta_warps * rows), n_views);  // (x, row blocks, views)
    cudaStream_t cstream = ctx->cur_cstream ? ctx->cur_cstream : ctx->stream;  // the lane of this frame batch
    pm3d_kt_mark(ctx, 2, 0, cstream);
    // batches of views are issued ROW BLOCK BY ROW BLOCK from the horizon outwards, every view's block of one rank
    // before any block of the next: the horizon blocks carry the sprites and the mixed rows — 5 to 30 times the time of
    // a sky or floor block (per-task timings: mean 56 us, up to 0.9 ms, against 7 us) — and issued view by view the
    // last views' horizon blocks were a tail of a few dozen busy warps a third as long as the whole launch
    p.row_major = n_views > 1 && n_views <= 65535;
    if (p.row_major) grid = dim3(grid.x, n_views, grid.y);
#ifdef PM3D_COMPOSE_STATS
    // tools builds only: PM3D_COMPOSE_LOG=path dumps per-warp start / end times of this launch (it synchronises)
    p.task_log = nullptr;
    const char *log_path = getenv("PM3D_COMPOSE_LOG");
    const size_t log_warps = (size_t)ctx->n_groups * ((ctx->h + rows - 1) / rows + 4) * n_views;
    if (log_path) { PM3D_CUDA(ctx, cudaMalloc((void **)&p.task_log, log_warps * 32)); PM3D_CUDA(ctx, cudaMemset(p.task_log, 0, log_warps * 32)); }
#endif
    if (rows == 16) {
        const int yb = (ctx->h + COMPOSE_TP_WARPS * 16 - 1) / (COMPOSE_TP_WARPS * 16);
        dim3 g16 = p.row_major ? dim3(ctx->n_groups, n_views, yb) : dim3(ctx->n_groups, yb, n_views);
        if (with_sprites && COMPOSE_TP_LATENCY) k_compose<16, 8, COMPOSE_TP_WARPS, true><<<g16, COMPOSE_TP_WARPS * 32, 0, cstream>>>(p);
        else k_compose<16, 8, COMPOSE_TP_WARPS><<<g16, COMPOSE_TP_WARPS * 32, 0, cstream>>>(p);
    }
    else if (rows == 8 && minb == 8 && with_sprites && COMPOSE_TP_LATENCY) k_compose<8, 8, 2, true><<<grid, block, 0, cstream>>>(p);
    else if (rows == 8 && minb == 8) k_compose<8, 8, 2><<<grid, block, 0, cstream>>>(p);
    else if (rows == 8) k_compose<8, 6, 2><<<grid, block, 0, cstream>>>(p);
    else if (rows == 4 && minb == 8) k_compose<4, 8, 4><<<grid, block, 0, cstream>>>(p);
    else if (rows == 4 && with_sprites) k_compose<4, 6, 4, true><<<grid, block, 0, cstream>>>(p);
    else if (rows == 4) k_compose<4, 6, 4><<<grid, block, 0, cstream>>>(p);
    else if (minb == 8 && with_sprites) k_compose<2, 8, 4, true><<<grid, block, 0, cstream>>>(p);
    else if (minb == 8) k_compose<2, 8, 4><<<grid, block, 0, cstream>>>(p);
    else if (with_sprites) k_compose<2, 6, 4, true><<<grid, block, 0, cstream>>>(p);
    else k_compose<2, 6, 4><<<grid, block, 0, cstream>>>(p);
    pm3d_kt_mark(ctx, 2, 1, cstream);
#ifdef PM3D_COMPOSE_STATS
    if (p.task_log) {
        std::vector<unsigned long long> hlog(log_warps * 4);
        PM3D_CUDA(ctx, cudaStreamSynchronize(cstream));
        PM3D_CUDA(ctx, cudaMemcpy(hlog.data(), p.task_log, hlog.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(log_path, "wb")) { fwrite(hlog.data(), 8, hlog.size(), f); fclose(f); }
        cudaFree(p.task_log);
    }
#endif
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}

int pm3d_launch_checksum(pm3d_ctx *ctx, int n_views, const uint8_t *frames) {
    const size_t words = (size_t)ctx->w * ctx->h;
    PM3D_CUDA(ctx, cudaMemsetAsync(ctx->checksums, 0, (size_t)n_views * sizeof(uint64_t), ctx->stream));
    dim3 grid((unsigned)((words + CHECKSUM_CHUNK - 1) / CHECKSUM_CHUNK), (unsigned)n_views);
    k_checksum<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(frames), ctx->checksums, words);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
