/*
 * pm3d_oracle.c — CPU ORACLE (test infrastructure; see pm3d_oracle.h header comment).
 * PARITY UNPINNED against the Go binary: restatement of reference call-site semantics [V],
 * the raycaster-go v1.12.0 published algorithm [R], and this repo's blit definition [D].
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include "pm3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- Go integer model */

/* Go int(float64) on amd64 = CVTTSD2SQ: truncation; NaN / out of range -> INT64_MIN. */
int64_t orc_go_int(double v) {
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;
}
/* Go int arithmetic wraps silently. */
static inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
static inline int64_t wneg(int64_t a) { return (int64_t)(0 - (uint64_t)a); }
static inline int64_t wmul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
/* geom.ClampInt [R] */
static inline int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---------------------------------------------------------------- blit model [D]
 * The reference never computes wall/sprite RGBA on the CPU: raycaster-go issues one
 * nearest-filtered ebiten DrawImage per column slice with a colour scale (SURVEY §7 hard
 * part 3).  This repo defines that blit in exact integer arithmetic:
 *   - textures are premultiplied RGBA8 (what ebiten.Image / image.RGBA hold);
 *   - a source span of n texels stretched over len destination pixels samples, at
 *     destination pixel k (centre k+0.5), texel ((2k+1) * floor(n*2^32/len)) >> 33;
 *   - colour scale + source-over is one rounding: out = (c*tint + dst*(255-a) + 127)/255.
 */
static inline uint64_t span_step(int64_t n_src, int64_t len_dst) {
    return ((uint64_t)n_src << 32) / (uint64_t)len_dst;
}
static inline int64_t span_texel(int64_t k, uint64_t step) { return (int64_t)((((uint64_t)(2 * k + 1)) * step) >> 33); }

static inline void blend_over(uint8_t *dst, const uint8_t *src, const uint8_t *tint) {
    uint32_t a = src[3];
    if (a == 0) return; /* premultiplied: colour is 0 too, identity */
    uint32_t ia = 255u - a;
    for (int ch = 0; ch < 3; ch++) {
        uint32_t v = ((uint32_t)src[ch] * tint[ch] + (uint32_t)dst[ch] * ia + 127u) / 255u;
        dst[ch] = (uint8_t)(v > 255u ? 255u : v);
    }
    uint32_t va = (a * 255u + (uint32_t)dst[3] * ia + 127u) / 255u;
    dst[3] = (uint8_t)(va > 255u ? 255u : va);
}
static const uint8_t TINT_ONE[4] = {255, 255, 255, 255};

/* ---------------------------------------------------------------- lighting [R]
 * st = {255,255,255} (minus 12 on side 0 for walls), then per channel
 * byte(ClampInt(int(float64(st) + sqrt(dist)*lightFalloff + globalIllumination (+ sprite illumination)),
 *               minLight, maxLight)).                                                   */
static inline void light_tint(const orc_camera *c, double dist, int base, double extra, int use_extra,
                              uint8_t out[4]) {
    double shadow = sqrt(dist) * c->light_falloff;
    for (int ch = 0; ch < 3; ch++) {
        double v = (double)base + shadow + c->global_illumination;
        if (use_extra) v = v + extra;
        out[ch] = (uint8_t)clamp_i64(orc_go_int(v), c->min_light[ch], c->max_light[ch]);
    }
    out[3] = 255;
}

/* ---------------------------------------------------------------- callbacks [V] */

/* Map.Level(l): levels above the stored ones repeat the top one (map.go:257-264). */
static inline const uint16_t *level_grid(const orc_scene *s, int l) {
    int sl = l < s->stored_levels ? l : s->stored_levels - 1;
    return s->grid + (size_t)sl * s->map_w * s->map_h;
}
/* TextureHandler.TextureAt (texture.go:69-97) + MapTexture.GetImage (map.go:106-115). */
static inline int texture_at(const orc_scene *s, int64_t x, int64_t y, int level, int side) {
    if (!(x >= 0 && x < s->map_w && y >= 0 && y < s->map_h)) return -1;
    int tex = level_grid(s, level)[x * s->map_h + y];
    if (tex <= 0) return -1;
    return s->wall_lut[tex * 2 + side];
}
/* TextureHandler.FloorTextureAt (texture.go:119-133); past the map edge the Go code panics,
 * the oracle defines that as "no texel" (SURVEY §8a row F). */
static inline int floor_texture_at(const orc_scene *s, int64_t x, int64_t y) {
    if (x < 0 || y < 0 || x >= s->map_w || y >= s->map_h) return -1;
    return s->floor_tiles[x * s->map_h + y];
}

/* ---------------------------------------------------------------- K2 wall + floor cast [R]
 * raycaster-go Camera.castLevel for one column of one level.  floor_img (may be NULL) is
 * the camera's horizontal buffer: [h][w][4], written for level 0 only.                  */
static void cast_column(const orc_scene *s, const orc_camera *c, int x, int level, orc_column *rec,
                        double *zbuf, uint8_t *floor_img) {
    const uint16_t *grid = level_grid(s, level);
    const int64_t h = c->h;
    double cameraX = 2.0 * (double)x / (double)c->w - 1.0;
    double rayDirX = c->dir_x + c->plane_x * cameraX;
    double rayDirY = c->dir_y + c->plane_y * cameraX;
    double rayPosX = c->pos_x, rayPosY = c->pos_y;
    int64_t mapX = orc_go_int(rayPosX), mapY = orc_go_int(rayPosY);
    double deltaDistX = fabs(1.0 / rayDirX), deltaDistY = fabs(1.0 / rayDirY);
    double sideDistX, sideDistY, perpWallDist;
    int64_t stepX, stepY;
    int hit = 0, side = -1;

    if (rayDirX < 0) { stepX = -1; sideDistX = (rayPosX - (double)mapX) * deltaDistX; }
    else             { stepX = 1;  sideDistX = ((double)mapX + 1.0 - rayPosX) * deltaDistX; }
    if (rayDirY < 0) { stepY = -1; sideDistY = (rayPosY - (double)mapY) * deltaDistY; }
    else             { stepY = 1;  sideDistY = ((double)mapY + 1.0 - rayPosY) * deltaDistY; }

    while (hit == 0) {
        if (sideDistX < sideDistY) { sideDistX += deltaDistX; mapX += stepX; side = 0; }
        else                       { sideDistY += deltaDistY; mapY += stepY; side = 1; }
        if (mapX >= 0 && mapY >= 0 && mapX < s->map_w && mapY < s->map_h) {
            if (grid[mapX * s->map_h + mapY] > 0) hit = 1;
        } else {
            hit = 2; /* left the grid: clamp the cell back in range */
            if (mapX < 0) mapX = 0; else if (mapX > s->map_w - 1) mapX = s->map_w - 1;
            if (mapY < 0) mapY = 0; else if (mapY > s->map_h - 1) mapY = s->map_h - 1;
        }
    }

    if (side == 0) perpWallDist = ((double)mapX - rayPosX + (1.0 - (double)stepX) / 2.0) / rayDirX;
    else           perpWallDist = ((double)mapY - rayPosY + (1.0 - (double)stepY) / 2.0) / rayDirY;

    int64_t lineHeight = orc_go_int((double)h / perpWallDist);
    /* drawStart := (-lineHeight/2 + h/2) + pitch + int(camZ/perpWallDist) - lineHeight*levelNum */
    int64_t drawStart = wadd(wneg(lineHeight) / 2, h / 2);
    drawStart = wadd(drawStart, c->pitch_px);
    drawStart = wadd(drawStart, orc_go_int(c->cam_z_px / perpWallDist));
    drawStart = wsub(drawStart, wmul(lineHeight, level));
    int64_t drawEnd = wadd(drawStart, lineHeight);

    int tile = texture_at(s, mapX, mapY, level, side);

    double wallX;
    if (side == 0) wallX = rayPosY + perpWallDist * rayDirY;
    else           wallX = rayPosX + perpWallDist * rayDirX;
    wallX -= floor(wallX);
    int64_t texX = orc_go_int(wallX * (double)s->tex_size);
    if (side == 0 && rayDirX > 0) texX = s->tex_size - texX - 1;
    if (side == 1 && rayDirY < 0) texX = s->tex_size - texX - 1;

    /* [D] render distance: a slice farther than renderDistance is not drawn; the z-buffer
     * still records the real hit distance. */
    if (c->render_distance >= 0 && !(perpWallDist <= c->render_distance)) tile = -1;
    /* [D] a texture column outside the tile (NaN wallX) draws nothing instead of panicking */
    if (texX < 0 || texX >= s->tex_size) tile = -1;

    rec->perp_dist = perpWallDist;
    rec->line_height = lineHeight;
    rec->draw_start = drawStart;
    rec->draw_end = drawEnd;
    rec->map_x = (int32_t)mapX;
    rec->map_y = (int32_t)mapY;
    rec->tex_x = (int32_t)clamp_i64(texX, INT32_MIN, INT32_MAX);
    rec->tile = (int16_t)tile;
    rec->side = (uint8_t)side;
    rec->hit = (uint8_t)hit;
    light_tint(c, perpWallDist, side == 0 ? 255 - 12 : 255, 0.0, 0, rec->tint);

    if (level != 0) return;
    zbuf[x] = perpWallDist;
    if (!floor_img) return;

    /* ---- floor cast, level 0 only [R] ---- */
    if (drawEnd < 0) drawEnd = h; /* "becomes < 0 when the integer overflows" */
    double floorXWall, floorYWall;
    if (side == 0 && rayDirX > 0)      { floorXWall = (double)mapX;         floorYWall = (double)mapY + wallX; }
    else if (side == 0 && rayDirX < 0) { floorXWall = (double)mapX + 1.0;   floorYWall = (double)mapY + wallX; }
    else if (side == 1 && rayDirY > 0) { floorXWall = (double)mapX + wallX; floorYWall = (double)mapY; }
    else                               { floorXWall = (double)mapX + wallX; floorYWall = (double)mapY + 1.0; }
    double distWall = perpWallDist;

    /* drawEnd >= 0 here, so y >= 1; the INT64_MAX guard only avoids signed wrap in C */
    for (int64_t y = drawEnd >= h ? h : drawEnd + 1; y < h; y++) {
        double currentDist = ((double)h + 2.0 * c->cam_z_px) / (2.0 * (double)(y - c->pitch_px) - (double)h);
        if (c->render_distance >= 0 && !(currentDist <= c->render_distance)) continue;
        double weight = currentDist / distWall;
        double fx = weight * floorXWall + (1.0 - weight) * rayPosX;
        double fy = weight * floorYWall + (1.0 - weight) * rayPosY;
        if (!(fx >= 0.0) || !(fy >= 0.0)) continue;
        int64_t tx = orc_go_int(fx * (double)s->tex_size), ty = orc_go_int(fy * (double)s->tex_size);
        int64_t cx = orc_go_int(fx), cy = orc_go_int(fy);
        int ft = floor_texture_at(s, cx, cy);
        if (ft < 0) continue;
        int64_t ftx = tx % s->tex_size, fty = ty % s->tex_size;
        if (ftx < 0 || fty < 0) continue;
        /* image.RGBA.PixOffset(x, y): row-major, stride 4*tex (resources.go:215-243) */
        const uint8_t *px = s->atlas + (((size_t)ft * s->tex_size + fty) * s->tex_size + ftx) * 4;
        if (px[3] == 0) continue;
        uint8_t st[4];
        light_tint(c, currentDist, 255, 0.0, 0, st);
        uint8_t *o = floor_img + ((size_t)y * c->w + x) * 4;
        /* pixel.R = uint8(float64(pixel.R) * float64(pixelSt.R) / 256) */
        o[0] = (uint8_t)(((uint32_t)px[0] * st[0]) >> 8);
        o[1] = (uint8_t)(((uint32_t)px[1] * st[1]) >> 8);
        o[2] = (uint8_t)(((uint32_t)px[2] * st[2]) >> 8);
        o[3] = px[3];
    }
}

void orc_cast_walls(const orc_scene *s, const orc_camera *c, orc_column *columns, double *zbuf) {
    for (int l = 0; l < s->num_levels; l++)
        for (int x = 0; x < c->w; x++) cast_column(s, c, x, l, &columns[(size_t)l * c->w + x], zbuf, NULL);
}

/* ---------------------------------------------------------------- sprites [R] */

typedef struct {
    int visible;       /* any chance of drawing (transformY > 0, non-empty span) */
    double transformY; /* depth */
    double dist;       /* euclidean distance camera -> sprite */
    int64_t dsx, dex, dsy, dey; /* drawStartX/EndX/StartY/EndY (unclamped) */
    int64_t sprite_w;
    int64_t tex_w, tex_h;
    uint64_t ystep;
    uint8_t tint[4];
} sprite_proj;

static double sprite_distance(const orc_camera *c, const orc_sprite *sp) {
    double dx = c->pos_x - sp->x, dy = c->pos_y - sp->y;
    return sqrt(dx * dx + dy * dy);
}

/* far -> near.  [R] raycaster-go comb-sorts by distance (unstable); [D] ties are broken by
 * list index so the order is a pure function of the input. */
void orc_sprite_order(const orc_camera *c, const orc_sprite *sprites, int32_t n, int32_t *order) {
    double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { order[i] = i; d[i] = sprite_distance(c, &sprites[i]); }
    /* stable insertion-free merge sort is overkill: simple stable binary-insertion sort */
    for (int i = 1; i < n; i++) {
        int32_t oi = order[i];
        double di = d[oi];
        int j = i - 1;
        /* move while strictly nearer than the one being inserted (NaN compares false: stays put) */
        while (j >= 0 && d[order[j]] < di) { order[j + 1] = order[j]; j--; }
        order[j + 1] = oi;
    }
    free(d);
}

/* getAnchorVerticalOffset [R] */
static double anchor_offset(int anchor, double scale, int64_t h) {
    double half = (double)h / 2.0;
    switch (anchor) {
    case ORC_ANCHOR_BOTTOM: return half - scale * half;
    case ORC_ANCHOR_CENTER: return half;
    case ORC_ANCHOR_TOP: return half + scale * half;
    }
    return 0.0;
}

static void project_sprite(const orc_scene *s, const orc_camera *c, const orc_sprite *sp, sprite_proj *p) {
    const int64_t h = c->h, w = c->w;
    memset(p, 0, sizeof(*p));
    double spriteX = sp->x - c->pos_x, spriteY = sp->y - c->pos_y;
    double invDet = 1.0 / (c->plane_x * c->dir_y - c->dir_x * c->plane_y);
    double transformX = invDet * (c->dir_y * spriteX - c->dir_x * spriteY);
    double transformY = invDet * (-c->plane_y * spriteX + c->plane_x * spriteY);
    p->transformY = transformY;
    p->dist = sprite_distance(c, sp);
    p->tex_w = (int64_t)sp->rx1 - sp->rx0;
    p->tex_h = (int64_t)sp->ry1 - sp->ry0;

    int64_t spriteScreenX = orc_go_int((double)w / 2.0 * (1.0 + transformX / transformY));
    double ratioWH = (double)p->tex_w / (double)p->tex_h;
    double uDiv = 1.0 / (sp->scale * ratioWH);
    double vDiv = 1.0 / sp->scale;
    double vMove = -sp->z * (double)h + anchor_offset(sp->anchor, sp->scale, h);
    int64_t vMoveScreen = wadd(wadd(orc_go_int(vMove / transformY), c->pitch_px), orc_go_int(c->cam_z_px / transformY));

    int64_t spriteHeight = orc_go_int(fabs((double)h / transformY) / vDiv);
    p->dsy = wadd(wadd(wneg(spriteHeight) / 2, h / 2), vMoveScreen);
    p->dey = wadd(wadd(spriteHeight / 2, h / 2), vMoveScreen);
    int64_t spriteWidth = orc_go_int(fabs((double)h / transformY) / uDiv);
    p->sprite_w = spriteWidth;
    p->dsx = wadd(wneg(spriteWidth) / 2, spriteScreenX);
    p->dex = wadd(spriteWidth / 2, spriteScreenX);

    light_tint(c, transformY, 255, sp->illumination, 1, p->tint);

    int ok = transformY > 0 && spriteWidth > 0 && spriteHeight > 0 && p->dex > p->dsx && p->dey > p->dsy &&
             p->tex_w > 0 && p->tex_h > 0 && sp->sheet >= 0 && sp->sheet < s->n_sheets;
    /* [D] render distance also culls sprites (the game pre-culls: game/sprites.go:113-114) */
    if (c->render_distance >= 0 && !(p->dist <= c->render_distance)) ok = 0;
    /* [D] spans beyond 2^20 pixels (camera all but inside the billboard plane) and texture
     * rects beyond 4096 texels are not drawn: keeps (stripe-drawStartX)*texW below 2^32 */
    if (ok && (spriteWidth > ((int64_t)1 << 20) || spriteHeight > ((int64_t)1 << 20) || p->tex_w > 4096 ||
               p->tex_h > 4096))
        ok = 0;
    p->visible = ok;
    if (ok) p->ystep = span_step(p->tex_h, p->dey - p->dsy);
}

/* texture column for screen stripe [R]: int(256*(stripe-drawStartX)*texW/spriteWidth)/256 */
static inline int64_t sprite_tex_x(const sprite_proj *p, int64_t stripe) {
    int64_t k = stripe - p->dsx;
    return (256 * k * p->tex_w / p->sprite_w) / 256;
}

/* ---------------------------------------------------------------- K2+K3+K4 frame */

static void draw_box(const orc_scene *s, const orc_camera *c, uint8_t *rgba, int tile, int64_t y0, int64_t y1) {
    /* [R] drawTexture(screen, tex, texRect(0,0,tex,tex), dstRect(0,y0,w,y1), maxLightRGB) */
    if (tile < 0 || y1 <= y0) return;
    uint8_t tint[4] = {c->max_light[0], c->max_light[1], c->max_light[2], 255};
    uint64_t ystep = span_step(s->tex_size, y1 - y0), xstep = span_step(s->tex_size, c->w);
    int64_t ya = y0 < 0 ? 0 : y0, yb = y1 > c->h ? c->h : y1;
    for (int64_t y = ya; y < yb; y++) {
        int64_t ty = span_texel(y - y0, ystep);
        if (ty > s->tex_size - 1) ty = s->tex_size - 1;
        for (int64_t x = 0; x < c->w; x++) {
            int64_t tx = span_texel(x, xstep);
            if (tx > s->tex_size - 1) tx = s->tex_size - 1;
            blend_over(rgba + ((size_t)y * c->w + x) * 4,
                       s->atlas + (((size_t)tile * s->tex_size + ty) * s->tex_size + tx) * 4, tint);
        }
    }
}

void orc_render(const orc_scene *s, const orc_camera *c, const orc_sprite *sprites, int32_t n_sprites,
                uint8_t *rgba, orc_rect *rects, orc_convergence *conv, orc_column *columns_out, double *zbuf_out,
                int32_t threads) {
    const int w = c->w, h = c->h, L = s->num_levels;
    orc_column *columns = columns_out ? columns_out : (orc_column *)malloc(sizeof(orc_column) * (size_t)L * w);
    double *zbuf = zbuf_out ? zbuf_out : (double *)malloc(sizeof(double) * (size_t)w);
    uint8_t *floor_img = (uint8_t *)calloc((size_t)w * h, 4);
    if (threads < 1) threads = 1;
    (void)threads;

    /* ---- Camera.Update: cast every level (raycaster-go runs one goroutine per level) ---- */
#pragma omp parallel for collapse(2) schedule(dynamic, 16) num_threads(threads)
    for (int l = 0; l < L; l++)
        for (int x = 0; x < w; x++) cast_column(s, c, x, l, &columns[(size_t)l * w + x], zbuf, floor_img);

    /* ---- sprite casting ---- */
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sprites > 0 ? n_sprites : 1));
    sprite_proj *proj = (sprite_proj *)malloc(sizeof(sprite_proj) * (size_t)(n_sprites > 0 ? n_sprites : 1));
    orc_sprite_order(c, sprites, n_sprites, order);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n_sprites; i++) project_sprite(s, c, &sprites[i], &proj[i]);

    /* convergence [D]: the centre ray's wall hit, overridden by the nearest focusable sprite
     * that covers the screen centre and passes the z test there. */
    const int cx = w / 2, cy = h / 2;
    if (conv) {
        double cameraX = 2.0 * (double)cx / (double)w - 1.0;
        double rdx = c->dir_x + c->plane_x * cameraX, rdy = c->dir_y + c->plane_y * cameraX;
        conv->distance = zbuf[cx];
        conv->point_x = c->pos_x + zbuf[cx] * rdx;
        conv->point_y = c->pos_y + zbuf[cx] * rdy;
        conv->sprite = -1;
        conv->_pad = 0;
    }
    for (int i = 0; i < n_sprites; i++) {
        const sprite_proj *p = &proj[i];
        /* SetScreenRect: alwaysSetSpriteScreenRect=true (mission.go:76) -> rect whenever the
         * sprite is in front of the camera, nil otherwise. image.Rect canonicalises. */
        if (rects) {
            if (p->visible) {
                rects[i].x0 = (int32_t)clamp_i64(p->dsx, -(1 << 30), 1 << 30);
                rects[i].y0 = (int32_t)clamp_i64(p->dsy, -(1 << 30), 1 << 30);
                rects[i].x1 = (int32_t)clamp_i64(p->dex, -(1 << 30), 1 << 30);
                rects[i].y1 = (int32_t)clamp_i64(p->dey, -(1 << 30), 1 << 30);
            } else {
                rects[i].x0 = 0; rects[i].y0 = 0; rects[i].x1 = -1; rects[i].y1 = -1;
            }
        }
        if (conv && p->visible && sprites[i].focusable && cx > 0 && cx >= p->dsx && cx < p->dex && cy >= p->dsy &&
            cy < p->dey && p->transformY < zbuf[cx]) {
            if (conv->sprite < 0 || p->dist < proj[conv->sprite].dist) {
                conv->sprite = i;
                conv->distance = p->dist;
                conv->point_x = sprites[i].x;
                conv->point_y = sprites[i].y;
            }
        }
    }

    /* ---- Camera.Draw: painter's order [R] ---- */
    memset(rgba, 0, (size_t)w * h * 4);
    int64_t horizon = (int64_t)orc_go_int((double)h * 0.5) + c->pitch_px;
    draw_box(s, c, rgba, s->floorbox_tile, horizon, h);
    draw_box(s, c, rgba, s->sky_tile, 0, horizon);

    /* walls: for each column, levels top -> 0 */
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (int x = 0; x < w; x++) {
        for (int l = L - 1; l >= 0; l--) {
            const orc_column *r = &columns[(size_t)l * w + x];
            if (r->tile < 0 || r->line_height <= 0 || r->draw_end <= r->draw_start) continue;
            uint64_t step = span_step(s->tex_size, r->line_height);
            int64_t ya = r->draw_start < 0 ? 0 : r->draw_start, yb = r->draw_end > h ? h : r->draw_end;
            const uint8_t *tile = s->atlas + (size_t)r->tile * s->tex_size * s->tex_size * 4;
            for (int64_t y = ya; y < yb; y++) {
                int64_t ty = span_texel(y - r->draw_start, step);
                if (ty > s->tex_size - 1) ty = s->tex_size - 1;
                blend_over(rgba + ((size_t)y * w + x) * 4, tile + ((size_t)ty * s->tex_size + r->tex_x) * 4, r->tint);
            }
        }
    }
    /* textured floor image over the walls */
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) blend_over(rgba + ((size_t)y * w + x) * 4, floor_img + ((size_t)y * w + x) * 4, TINT_ONE);

    /* sprites far -> near; inside one sprite columns are independent */
    for (int oi = 0; oi < n_sprites; oi++) {
        int i = order[oi];
        const sprite_proj *p = &proj[i];
        if (!p->visible) continue;
        const orc_sprite *sp = &sprites[i];
        const uint8_t *sheet = s->sheet_pixels + s->sheet_offset[sp->sheet];
        const int64_t sw = s->sheet_w[sp->sheet], sh = s->sheet_h[sp->sheet];
        int64_t xa = p->dsx < 1 ? 1 : p->dsx, xb = p->dex > w ? w : p->dex; /* stripe > 0 && stripe < w */
        int64_t ya = p->dsy < 0 ? 0 : p->dsy, yb = p->dey > h ? h : p->dey;
#pragma omp parallel for schedule(static) num_threads(threads) if (xb - xa > 64)
        for (int64_t stripe = xa; stripe < xb; stripe++) {
            if (!(p->transformY < zbuf[stripe])) continue;
            int64_t tx = sprite_tex_x(p, stripe);
            if (tx < 0 || tx >= p->tex_w) continue;
            int64_t sx = sp->rx0 + tx;
            if (sx < 0 || sx >= sw) continue;
            for (int64_t y = ya; y < yb; y++) {
                int64_t ty = span_texel(y - p->dsy, p->ystep);
                if (ty > p->tex_h - 1) ty = p->tex_h - 1;
                int64_t sy = sp->ry0 + ty;
                if (sy < 0 || sy >= sh) continue;
                blend_over(rgba + ((size_t)y * w + stripe) * 4, sheet + ((size_t)sy * sw + sx) * 4, p->tint);
            }
        }
    }

    free(order);
    free(proj);
    free(floor_img);
    if (!columns_out) free(columns);
    if (!zbuf_out) free(zbuf);
}

/* ---------------------------------------------------------------- LOS / collision */

/* geom.LineIntersection [R] (raycaster-go geom; Wikipedia two-point form).  Endpoints inclusive. */
int orc_line_intersection(const double *l1, const double *l2, double *px, double *py) {
    double x1 = l1[0], y1 = l1[1], x2 = l1[2], y2 = l1[3];
    double x3 = l2[0], y3 = l2[1], x4 = l2[2], y4 = l2[3];
    double denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    double tNum = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4);
    double uNum = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3));
    if (denom == 0) return 0;
    double t = tNum / denom;
    if (t > 1 || t < 0) return 0;
    double u = uNum / denom;
    if (u > 1 || u < 0) return 0;
    if (px) *px = x1 + t * (x2 - x1);
    if (py) *py = y1 + t * (y2 - y1);
    return 1;
}

/* image.Rect(x0,y0,x1,y1) canonicalises so Min <= Max. */
static inline void go_rect(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t r[4]) {
    if (x0 > x1) { int64_t t = x0; x0 = x1; x1 = t; }
    if (y0 > y1) { int64_t t = y0; y0 = y1; y1 = t; }
    r[0] = x0; r[1] = y0; r[2] = x1; r[3] = y1;
}

/* Game.lineOfSight (game/collision.go:40-65) [V], including the integer bounding-rect
 * pre-filter at :49-56 exactly as written. */
static uint8_t los_one(const double *segs, int32_t n_segs, const double *ray) {
    int64_t lr[4], br[4];
    go_rect(orc_go_int(ray[0]), orc_go_int(ray[1]), orc_go_int(ray[2]), orc_go_int(ray[3]), lr);
    for (int32_t i = 0; i < n_segs; i++) {
        const double *b = segs + (size_t)i * 4;
        go_rect(orc_go_int(b[0]), orc_go_int(b[1]), orc_go_int(b[2]), orc_go_int(b[3]), br);
        if ((br[0] > lr[2] && br[1] > lr[3]) || (br[2] < lr[0] && br[3] < lr[1])) continue;
        if (orc_line_intersection(ray, b, NULL, NULL)) return 0;
    }
    return 1;
}

void orc_los_batch(const double *segs, int32_t n_segs, const double *rays, int64_t n_rays, uint8_t *out,
                   int32_t threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (int64_t i = 0; i < n_rays; i++) out[i] = los_one(segs, n_segs, rays + (size_t)i * 4);
}

/* Wall part of Game.getValidMove (game/collision.go:69-355) [V] for checkAlternate=false,
 * an entity with positive radius/height and no sprite colliders, 2-D move (z unchanged):
 *   :78  no-op move -> unchanged, no collision
 *   :93-97 every collision-map segment hit by the move segment is an intersection point
 *   :202-203,:307-311 any intersection -> stay put, collided
 *   :314-342 clamp to the map, then the level-0 cell test. */
void orc_move_walls_batch(const orc_scene *s, const double *segs, int32_t n_segs, const double *moves,
                          int64_t n_moves, double clip, double *out_pos, uint8_t *out_collided, int32_t *out_hits,
                          int32_t threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (int64_t i = 0; i < n_moves; i++) {
        const double *m = moves + (size_t)i * 4;
        double posX = m[0], posY = m[1], newX = m[2], newY = m[3];
        int32_t hits = 0;
        uint8_t collided = 0;
        if (posX == newX && posY == newY) {
            out_pos[i * 2] = posX; out_pos[i * 2 + 1] = posY; out_collided[i] = 0; out_hits[i] = 0;
            continue;
        }
        for (int32_t k = 0; k < n_segs; k++)
            if (orc_line_intersection(m, segs + (size_t)k * 4, NULL, NULL)) hits++;
        if (hits > 0) {
            out_pos[i * 2] = posX; out_pos[i * 2 + 1] = posY; out_collided[i] = 1; out_hits[i] = hits;
            continue;
        }
        int64_t ix = orc_go_int(newX), iy = orc_go_int(newY);
        if (ix < 0 || newX < 0) { newX = clip; ix = 0; }
        else if (ix >= s->map_w) { newX = (double)s->map_w - clip; ix = orc_go_int(newX); }
        if (iy < 0 || newY < 0) { newY = clip; iy = 0; }
        else if (iy >= s->map_h) { newY = (double)s->map_h - clip; iy = orc_go_int(newY); }
        if (ix > s->map_w - 1) ix = s->map_w - 1; /* clip == 0 would index past the edge (Go panics) */
        if (iy > s->map_h - 1) iy = s->map_h - 1;
        if (s->grid[ix * s->map_h + iy] <= 0) { posX = newX; posY = newY; }
        else collided = 1;
        out_pos[i * 2] = posX; out_pos[i * 2 + 1] = posY; out_collided[i] = collided; out_hits[i] = 0;
    }
}
