// cast_kernel.cu — K2 wall cast: one thread per (view, level, screen column) ray.
//
// Replaces the per-column DDA inside raycaster.Camera.Update (call site
// game/scene_game.go:161) together with the Map.Level / TextureHandler.TextureAt callbacks
// (game/model/map.go:257-264, game/texture/texture.go:69-97), which arrive here as the
// flattened grid and the (texIndex, side) -> tile LUT.
//
// Arithmetic: IEEE binary64, no FMA contraction (-fmad=false), Go integer semantics via
// go_int/wadd/...; the visited cell sequence is therefore identical to the CPU reference's.
#include "pm3d_internal.cuh"

struct CastParams {
    const pm3d_view *views;
    const uint16_t *grid;
    const int16_t *wall_lut;
    ColRec *col_rec;
    FloorRec *floor_rec;
    double *zbuf;
    pm3d_column *dbg;  // may be null
    int w, h, map_w, map_h, stored_levels, num_levels, tex_size;
};

template <bool DEBUG>
__global__ void __launch_bounds__(128) k_cast_columns(CastParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int level = blockIdx.y;
    const int view = blockIdx.z;
    if (x >= p.w) return;
    const pm3d_view c = p.views[view];
    const int sl = level < p.stored_levels ? level : p.stored_levels - 1;  // Map.Level()
    const uint16_t *__restrict__ grid = p.grid + (size_t)sl * p.map_w * p.map_h;
    const int64_t h = p.h;

    const double cameraX = 2.0 * (double)x / (double)p.w - 1.0;
    const double rayDirX = c.dir_x + c.plane_x * cameraX;
    const double rayDirY = c.dir_y + c.plane_y * cameraX;
    const double rayPosX = c.pos_x, rayPosY = c.pos_y;
    int64_t mapX = go_int(rayPosX), mapY = go_int(rayPosY);
    const double deltaDistX = fabs(1.0 / rayDirX), deltaDistY = fabs(1.0 / rayDirY);
    double sideDistX, sideDistY;
    int stepX, stepY;
    if (rayDirX < 0) { stepX = -1; sideDistX = (rayPosX - (double)mapX) * deltaDistX; }
    else             { stepX = 1;  sideDistX = ((double)mapX + 1.0 - rayPosX) * deltaDistX; }
    if (rayDirY < 0) { stepY = -1; sideDistY = (rayPosY - (double)mapY) * deltaDistY; }
    else             { stepY = 1;  sideDistY = ((double)mapY + 1.0 - rayPosY) * deltaDistY; }

    int hit = 0, side = -1;
    // The start cell may be far outside the grid (Go ints are 64-bit); once the walk is
    // inside int32 range all cell arithmetic is 32-bit.
    if (mapX < -2147480000ll || mapX > 2147480000ll || mapY < -2147480000ll || mapY > 2147480000ll) {
        // the very first step leaves the grid: same outcome as the reference loop
        if (sideDistX < sideDistY) { mapX += stepX; side = 0; } else { mapY += stepY; side = 1; }
        hit = 2;
        mapX = mapX < 0 ? 0 : (mapX > p.map_w - 1 ? p.map_w - 1 : mapX);
        mapY = mapY < 0 ? 0 : (mapY > p.map_h - 1 ? p.map_h - 1 : mapY);
    }
    int mx = (int)mapX, my = (int)mapY;
    while (hit == 0) {
        if (sideDistX < sideDistY) { sideDistX += deltaDistX; mx += stepX; side = 0; }
        else                       { sideDistY += deltaDistY; my += stepY; side = 1; }
        if ((unsigned)mx < (unsigned)p.map_w && (unsigned)my < (unsigned)p.map_h) {
            if (__ldg(grid + mx * p.map_h + my) > 0) hit = 1;
        } else {
            hit = 2;
            mx = mx < 0 ? 0 : (mx > p.map_w - 1 ? p.map_w - 1 : mx);
            my = my < 0 ? 0 : (my > p.map_h - 1 ? p.map_h - 1 : my);
        }
    }

    double perpWallDist;
    if (side == 0) perpWallDist = ((double)mx - rayPosX + (1.0 - (double)stepX) / 2.0) / rayDirX;
    else           perpWallDist = ((double)my - rayPosY + (1.0 - (double)stepY) / 2.0) / rayDirY;

    const int64_t lineHeight = go_int((double)h / perpWallDist);
    int64_t drawStart = wadd(wneg(lineHeight) / 2, h / 2);
    drawStart = wadd(drawStart, c.pitch_px);
    drawStart = wadd(drawStart, go_int(c.cam_z_px / perpWallDist));
    drawStart = wsub(drawStart, wmul(lineHeight, level));
    const int64_t drawEnd = wadd(drawStart, lineHeight);

    // TextureAt(mapX, mapY, level, side): mx,my are in range here
    int tile = -1;
    {
        const int tex = __ldg(grid + mx * p.map_h + my);
        if (tex > 0) tile = __ldg(p.wall_lut + tex * 2 + side);
    }

    double wallX;
    if (side == 0) wallX = rayPosY + perpWallDist * rayDirY;
    else           wallX = rayPosX + perpWallDist * rayDirX;
    wallX -= floor(wallX);
    int64_t texX = go_int(wallX * (double)p.tex_size);
    if (side == 0 && rayDirX > 0) texX = p.tex_size - texX - 1;
    if (side == 1 && rayDirY < 0) texX = p.tex_size - texX - 1;

    if (c.render_distance >= 0 && !(perpWallDist <= c.render_distance)) tile = -1;
    if (texX < 0 || texX >= p.tex_size) tile = -1;

    const uint32_t tint = light_tint(c, perpWallDist, side == 0 ? 255 - 12 : 255, 0.0, false);

    // ---- compose-side record
    ColRec r;
    r.ya = 0; r.yb = 0; r.tile_texx = 0; r.tint = tint; r.pos0 = 0; r.step2 = 0;
    if (tile >= 0 && lineHeight > 0 && drawEnd > drawStart) {
        const uint64_t step = span_step(p.tex_size, lineHeight);
        r.ya = (int32_t)(drawStart < 0 ? 0 : (drawStart > h ? h : drawStart));
        r.yb = (int32_t)(drawEnd > h ? h : (drawEnd < 0 ? 0 : drawEnd));
        r.tile_texx = ((uint32_t)tile << 16) | (uint32_t)texX;
        r.step2 = step * 2ull;
        r.pos0 = (1ull - 2ull * (uint64_t)drawStart) * step;  // (2*(0 - drawStart) + 1) * step, mod 2^64
    }
    const size_t ci = ((size_t)view * p.num_levels + level) * p.w + x;
    p.col_rec[ci] = r;

    if (DEBUG) {
        pm3d_column d;
        d.perp_dist = perpWallDist;
        d.line_height = lineHeight;
        d.draw_start = drawStart;
        d.draw_end = drawEnd;
        d.map_x = mx;
        d.map_y = my;
        d.tex_x = (int32_t)clamp_i64(texX, INT32_MIN, INT32_MAX);
        d.tile = (int16_t)tile;
        d.side = (uint8_t)side;
        d.hit = (uint8_t)hit;
        d.tint[0] = tint & 255u; d.tint[1] = (tint >> 8) & 255u; d.tint[2] = (tint >> 16) & 255u; d.tint[3] = 255;
        p.dbg[ci] = d;
    }

    if (level != 0) return;
    p.zbuf[(size_t)view * p.w + x] = perpWallDist;

    // ---- floor-cast parameters of this column
    FloorRec f;
    if (side == 0 && rayDirX > 0)      { f.fxw = (double)mx;         f.fyw = (double)my + wallX; }
    else if (side == 0 && rayDirX < 0) { f.fxw = (double)mx + 1.0;   f.fyw = (double)my + wallX; }
    else if (side == 1 && rayDirY > 0) { f.fxw = (double)mx + wallX; f.fyw = (double)my; }
    else                               { f.fxw = (double)mx + wallX; f.fyw = (double)my + 1.0; }
    f.dw = perpWallDist;
    int64_t de = drawEnd;
    if (de < 0) de = h;  // "becomes < 0 when the integer overflows"
    f.fs = (int32_t)(de >= h ? h : de + 1);
    f._pad = 0;
    p.floor_rec[(size_t)view * p.w + x] = f;
}

int pm3d_launch_cast(pm3d_ctx *ctx, int n_views) {
    CastParams p;
    p.views = ctx->cur_views;
    p.grid = ctx->grid;
    p.wall_lut = ctx->wall_lut;
    p.col_rec = ctx->col_rec;
    p.floor_rec = ctx->floor_rec;
    p.zbuf = ctx->zbuf;
    p.dbg = ctx->keep_columns ? ctx->dbg_cols : nullptr;
    p.w = ctx->w; p.h = ctx->h; p.map_w = ctx->map_w; p.map_h = ctx->map_h;
    p.stored_levels = ctx->stored_levels; p.num_levels = ctx->num_levels; p.tex_size = ctx->tex_size;
    dim3 block(128), grid((ctx->w + 127) / 128, ctx->num_levels, n_views);
    pm3d_kt_mark(ctx, 0, 0);
    if (p.dbg) k_cast_columns<true><<<grid, block, 0, ctx->stream>>>(p);
    else k_cast_columns<false><<<grid, block, 0, ctx->stream>>>(p);
    pm3d_kt_mark(ctx, 0, 1);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
