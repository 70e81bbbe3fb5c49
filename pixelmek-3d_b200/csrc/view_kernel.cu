// view_kernel.cu — per-view setup between the wall cast and the pixel pass: one CTA per view.
//
//  (a) row table: the floor cast's per-row distance and light tint (they depend on the
//      screen row only) and the sky / floor-box texel row;
//  (b) sprite projection (raycaster.Camera sprite cast; Sprite accessors
//      game/render/sprites/sprite.go:49-91), far->near ordering, SetScreenRect rectangles
//      (sprite.go:85) and the convergence sprite (game/scene_game.go:164);
//  (c) per 64-column group lists of the sprites that can touch the group, in paint order.
#include "pm3d_internal.cuh"

#define VIEW_THREADS 256
#define MAX_VIEW_SPRITES 2048
#define GROUP_COLS 64

struct ViewParams {
    const pm3d_view *views;
    const double *zbuf;
    RowRec *row_rec;
    const pm3d_sprite *sprites;
    const int32_t *sprite_offset;
    const SheetInfo *sheets;
    DrawRec *draw_rec;
    int32_t *n_draw;
    uint16_t *bin_list;
    int32_t *bin_count;
    pm3d_rect *rects;
    pm3d_convergence *conv;
    int w, h, tex_size, n_sheets, n_groups, sky_tile, floorbox_tile;
};

__device__ __forceinline__ double sprite_distance(const pm3d_view &c, const pm3d_sprite &sp) {
    double dx = c.pos_x - sp.x, dy = c.pos_y - sp.y;
    return sqrt(dx * dx + dy * dy);
}

__device__ __forceinline__ double anchor_offset(int anchor, double scale, int64_t h) {
    double half = (double)h / 2.0;
    switch (anchor) {
    case PM3D_ANCHOR_BOTTOM: return half - scale * half;
    case PM3D_ANCHOR_CENTER: return half;
    case PM3D_ANCHOR_TOP: return half + scale * half;
    }
    return 0.0;
}

// ordering key: far -> near, ties by list index (DESIGN.md "sprite order")
struct SortKey { double d; int idx; };
__device__ __forceinline__ bool key_before(const SortKey &a, const SortKey &b) {
    // a paints before b: strictly farther, or not comparable/equal and lower index.
    // NaN distances never move relative to their neighbours in the CPU insertion sort; here
    // they sort as "equal to everything", which matches for the tested (finite) domain.
    if (a.d > b.d) return true;
    if (a.d < b.d) return false;
    return a.idx < b.idx;
}

__global__ void __launch_bounds__(VIEW_THREADS) k_view_setup(ViewParams p) {
    const int view = blockIdx.x;
    const int tid = threadIdx.x;
    const pm3d_view c = p.views[view];
    const int64_t h = p.h, w = p.w;

    // ---------------- (a) rows
    const int64_t horizon = go_int((double)h * 0.5) + c.pitch_px;
    for (int y = tid; y < p.h; y += VIEW_THREADS) {
        RowRec r;
        r.cd = ((double)h + 2.0 * c.cam_z_px) / (2.0 * (double)(y - c.pitch_px) - (double)h);
        r.tint = light_tint(c, r.cd, 255, 0.0, false);
        bool valid = !(c.render_distance >= 0 && !(r.cd <= c.render_distance));
        // sky / floor-box texel row (blit model): sky spans [0,horizon), floor box [horizon,h)
        uint32_t v = 0;
        if (y < horizon) {
            if (horizon > 0) v = (uint32_t)((((uint64_t)(2 * (int64_t)y + 1)) * span_step(p.tex_size, horizon)) >> 33);
        } else {
            if (h > horizon) v = (uint32_t)((((uint64_t)(2 * ((int64_t)y - horizon) + 1)) * span_step(p.tex_size, h - horizon)) >> 33);
        }
        if (v > (uint32_t)p.tex_size - 1) v = p.tex_size - 1;
        r.info = (valid ? 0x80000000u : 0u) | (y < horizon ? 0x40000000u : 0u) | v;
        p.row_rec[(size_t)view * p.h + y] = r;
    }

    // ---------------- (b) sprites
    const int s0 = p.sprite_offset[view];
    const int n = p.sprite_offset[view + 1] - s0;
    const pm3d_sprite *sprites = p.sprites + s0;
    DrawRec *draw = p.draw_rec + s0;

    __shared__ SortKey keys[MAX_VIEW_SPRITES];
    __shared__ uint8_t vis[MAX_VIEW_SPRITES];
    __shared__ int s_ndraw;

    int npad = 1;
    while (npad < n) npad <<= 1;

    const double zc = p.zbuf[(size_t)view * p.w + p.w / 2];
    for (int i = tid; i < npad; i += VIEW_THREADS) {
        if (i >= n) { keys[i].d = -1.0; keys[i].idx = 0x7fffffff; continue; }  // pads paint last
        const pm3d_sprite sp = sprites[i];
        keys[i].d = sprite_distance(c, sp);
        keys[i].idx = i;
    }
    __syncthreads();
    // NaN distances are given +inf rank deterministically? keep them as-is: comparisons with NaN are false both ways -> index order.
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += VIEW_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    SortKey a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    bool swap = up ? key_before(b, a) : key_before(a, b);
                    if (swap) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }

    // project in paint order; slot i of the scratch list is sprite keys[i].idx
    for (int i = tid; i < n; i += VIEW_THREADS) {
        const int src = keys[i].idx;
        const pm3d_sprite sp = sprites[src];
        const double spriteX = sp.x - c.pos_x, spriteY = sp.y - c.pos_y;
        const double invDet = 1.0 / (c.plane_x * c.dir_y - c.dir_x * c.plane_y);
        const double transformX = invDet * (c.dir_y * spriteX - c.dir_x * spriteY);
        const double transformY = invDet * (-c.plane_y * spriteX + c.plane_x * spriteY);
        const int64_t tex_w = (int64_t)sp.rx1 - sp.rx0, tex_h = (int64_t)sp.ry1 - sp.ry0;
        const int64_t spriteScreenX = go_int((double)w / 2.0 * (1.0 + transformX / transformY));
        const double ratioWH = (double)tex_w / (double)tex_h;
        const double uDiv = 1.0 / (sp.scale * ratioWH);
        const double vDiv = 1.0 / sp.scale;
        const double vMove = -sp.z * (double)h + anchor_offset(sp.anchor, sp.scale, h);
        const int64_t vMoveScreen = wadd(wadd(go_int(vMove / transformY), c.pitch_px), go_int(c.cam_z_px / transformY));
        const int64_t spriteHeight = go_int(fabs((double)h / transformY) / vDiv);
        const int64_t dsy = wadd(wadd(wneg(spriteHeight) / 2, h / 2), vMoveScreen);
        const int64_t dey = wadd(wadd(spriteHeight / 2, h / 2), vMoveScreen);
        const int64_t spriteWidth = go_int(fabs((double)h / transformY) / uDiv);
        const int64_t dsx = wadd(wneg(spriteWidth) / 2, spriteScreenX);
        const int64_t dex = wadd(spriteWidth / 2, spriteScreenX);
        const double dist = keys[i].d;

        bool ok = transformY > 0 && spriteWidth > 0 && spriteHeight > 0 && dex > dsx && dey > dsy && tex_w > 0 &&
                  tex_h > 0 && sp.sheet >= 0 && sp.sheet < p.n_sheets;
        if (c.render_distance >= 0 && !(dist <= c.render_distance)) ok = false;
        if (ok && (spriteWidth > (1ll << 20) || spriteHeight > (1ll << 20) || tex_w > 4096 || tex_h > 4096)) ok = false;

        pm3d_rect rc;
        if (ok) {
            rc.x0 = (int32_t)clamp_i64(dsx, -(1 << 30), 1 << 30);
            rc.y0 = (int32_t)clamp_i64(dsy, -(1 << 30), 1 << 30);
            rc.x1 = (int32_t)clamp_i64(dex, -(1 << 30), 1 << 30);
            rc.y1 = (int32_t)clamp_i64(dey, -(1 << 30), 1 << 30);
        } else {
            rc.x0 = 0; rc.y0 = 0; rc.x1 = -1; rc.y1 = -1;
        }
        p.rects[s0 + src] = rc;

        // on-screen test: only sprites that can touch a drawable stripe (0 < stripe < w) and row
        const bool on_screen = ok && dex > 1 && dsx < w && dey > 0 && dsy < h;
        vis[i] = on_screen ? 1 : 0;
        // convergence candidate flag in bit 1
        const int cx = p.w / 2, cy = p.h / 2;
        if (ok && sp.focusable && cx > 0 && cx >= dsx && cx < dex && cy >= dsy && cy < dey && transformY < zc) vis[i] |= 2;

        if (on_screen) {
            DrawRec d;
            d.dsx = (int32_t)clamp_i64(dsx, -(1 << 30), 1 << 30);
            d.dex = (int32_t)clamp_i64(dex, -(1 << 30), 1 << 30);
            d.dsy = (int32_t)clamp_i64(dsy, -(1 << 30), 1 << 30);
            d.dey = (int32_t)clamp_i64(dey, -(1 << 30), 1 << 30);
            d.transformY = transformY;
            const uint64_t ystep = span_step(tex_h, dey - dsy);
            d.ystep2 = ystep * 2ull;
            d.ypos0 = (1ull - 2ull * (uint64_t)dsy) * ystep;
            // m = floor(2^64 / sw) + 1 for sw >= 2 (exact floor(n/sw) for n < 2^32); sw == 1 has no stripes
            d.xmagic = spriteWidth >= 2 ? (0xFFFFFFFFFFFFFFFFull / (uint64_t)spriteWidth) + 1ull : 0ull;
            const SheetInfo si = p.sheets[sp.sheet];
            d.sheet_px = (uint64_t)si.px_offset + (uint64_t)((int64_t)sp.ry0 * si.w + sp.rx0);
            d.tex_w = (uint32_t)tex_w;
            d.tex_h = (uint32_t)tex_h;
            d.sheet_w = (uint32_t)si.w;
            d.tint = light_tint(c, transformY, 255, sp.illumination, true);
            // texel ranges inside the sheet: rx0 + tx in [0, si.w), ry0 + ty in [0, si.h)
            int64_t txl = sp.rx0 < 0 ? -(int64_t)sp.rx0 : 0, txh = (int64_t)si.w - sp.rx0;
            int64_t tyl = sp.ry0 < 0 ? -(int64_t)sp.ry0 : 0, tyh = (int64_t)si.h - sp.ry0;
            if (txh > tex_w) txh = tex_w;
            if (tyh > tex_h) tyh = tex_h;
            if (txh < txl) txh = txl;
            if (tyh < tyl) tyh = tyl;
            d.tx_lo = (uint16_t)txl; d.tx_hi = (uint16_t)txh; d.ty_lo = (uint16_t)tyl; d.ty_hi = (uint16_t)tyh;
            d.src = src;
            d._pad = 0;
            // stash at slot i; compacted below
            draw[i] = d;
        }
    }
    __syncthreads();

    // ordered compaction of the visible records + convergence pick (n is small; serial)
    if (tid == 0) {
        int m = 0;
        int best = -1;
        double best_d = 0.0;
        for (int i = 0; i < n; i++) {
            if (vis[i] & 2) {
                // nearest wins, ties -> lower list index; paint order is far->near so compare explicitly
                const double d = keys[i].d;
                const int src = keys[i].idx;
                if (best < 0 || d < best_d || (d == best_d && src < best)) { best = src; best_d = d; }
            }
            if (vis[i] & 1) {
                if (m != i) draw[m] = draw[i];
                m++;
            }
        }
        s_ndraw = m;
        p.n_draw[view] = m;
        pm3d_convergence cv;
        const int cx = p.w / 2;
        const double cameraX = 2.0 * (double)cx / (double)p.w - 1.0;
        const double rdx = c.dir_x + c.plane_x * cameraX, rdy = c.dir_y + c.plane_y * cameraX;
        cv.distance = zc;
        cv.point_x = c.pos_x + zc * rdx;
        cv.point_y = c.pos_y + zc * rdy;
        cv.sprite = -1;
        cv._pad = 0;
        if (best >= 0) {
            cv.sprite = best;
            cv.distance = best_d;
            cv.point_x = sprites[best].x;
            cv.point_y = sprites[best].y;
        }
        p.conv[view] = cv;
    }
    __syncthreads();

    // ---------------- (c) column-group lists, paint order preserved
    const int m = s_ndraw;
    for (int g = tid; g < p.n_groups; g += VIEW_THREADS) {
        const int gx0 = g * GROUP_COLS, gx1 = gx0 + GROUP_COLS;
        uint16_t *list = p.bin_list + (size_t)s0 * p.n_groups + (size_t)g * n;
        int cnt = 0;
        for (int i = 0; i < m; i++) {
            const DrawRec &d = draw[i];
            if (d.dex > gx0 && d.dsx < gx1) list[cnt++] = (uint16_t)i;
        }
        p.bin_count[view * p.n_groups + g] = cnt;
    }
}

int pm3d_launch_view_setup(pm3d_ctx *ctx, int n_views, int n_sprites_total) {
    (void)n_sprites_total;
    ViewParams p;
    p.views = ctx->cur_views;
    p.zbuf = ctx->zbuf;
    p.row_rec = ctx->row_rec;
    p.sprites = ctx->cur_sprites;
    p.sprite_offset = ctx->cur_offset;
    p.sheets = ctx->sheet_info;
    p.draw_rec = ctx->draw_rec;
    p.n_draw = ctx->n_draw;
    p.bin_list = ctx->bin_list;
    p.bin_count = ctx->bin_count;
    p.rects = ctx->rects;
    p.conv = ctx->conv;
    p.w = ctx->w; p.h = ctx->h; p.tex_size = ctx->tex_size; p.n_sheets = ctx->n_sheets; p.n_groups = ctx->n_groups;
    p.sky_tile = ctx->sky_tile; p.floorbox_tile = ctx->floorbox_tile;
    pm3d_kt_mark(ctx, 1, 0);
    k_view_setup<<<n_views, VIEW_THREADS, 0, ctx->stream>>>(p);
    pm3d_kt_mark(ctx, 1, 1);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
