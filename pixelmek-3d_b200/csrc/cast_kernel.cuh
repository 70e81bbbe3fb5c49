// cast_kernel.cuh — K2 wall cast: one thread per (view, level, screen column) ray.
//
// Replaces the per-column DDA inside raycaster.Camera.Update (call site
// game/scene_game.go:161) together with the Map.Level / TextureHandler.TextureAt callbacks
// (game/model/map.go:257-264, game/texture/texture.go:69-97), which arrive here as the
// flattened grid and the (texIndex, side) -> tile LUT.
//
// Arithmetic: IEEE binary64, no FMA contraction (-fmad=false), Go integer semantics via
// go_int/wadd/...; the visited cell sequence is therefore identical to the CPU reference's.
//
// Empty-space skipping without changing that sequence: `dist` holds, per cell, the Chebyshev
// distance D to the nearest wall cell or map edge.  Every DDA step moves one cell along one
// axis, so the D-1 steps after a cell with distance D cannot reach a wall or leave the grid:
// they are executed as pure register arithmetic (the same sideDist additions, in the same
// order) with no memory access and no bounds test, and only every D-th step probes the grid.
// On the bundled 500x500 map the upper levels are almost empty, which turns ~700 dependent
// L2 loads per ray into 3-4.
#pragma once
#include "pm3d_internal.cuh"
#ifdef PM3D_FRONT_STATS
__device__ unsigned long long g_front_stats[8];  // (experiments only) 0 jump fallbacks, 1 jump rays, 2 slowest cast thread, 3 slowest view CTA
#endif

struct CastParams {
    const pm3d_view *views;
    const uint16_t *grid;
    const double *sparse_cells;   // [stored_levels][PM3D_SPARSE_MAX][2]
    const int32_t *sparse_count;  // [stored_levels], -1 = dense
    const uint8_t *dist;  // per stored level: Chebyshev distance to the nearest wall cell or map edge (0 = wall)
    const uint16_t *qdist;  // per stored level and step quadrant: DDA steps to the nearest wall cell / map exit (0 = wall)
    int jump;             // 1: closed-form jumps between probes (dda_jump), 0: the sequential loop only
    const int16_t *wall_lut;
    int2 *col_span;
    uint2 *col_tt;
    ulonglong2 *col_ps;
    double *flo_xw, *flo_yw, *flo_dw, *flo_rc, *flo_dx, *flo_dy;
    int32_t *flo_fs;
    int4 *agg;
    double *zbuf;
    pm3d_column *dbg;  // may be null
    int w, h, map_w, map_h, stored_levels, num_levels, tex_size, n_groups;
};

// Called by all 64 threads of a (view, level, group) pair of warps; `x` may be >= w in the last group
// (those lanes cast column w-1 again and store nothing) so that the warp-wide aggregate stays converged.
template <bool DEBUG>
__device__ __forceinline__ void cast_ray(const CastParams &p, const int view, const int level, const int group,
                                         const int x_in) {
    const bool live = x_in < p.w;
    const int x = live ? x_in : p.w - 1;
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

    // ---- sparse upper levels.  A level above the ground with a handful of wall cells is mostly crossed without a
    // hit: the ray leaves the grid, draws nothing and (not being level 0) feeds neither the z-buffer nor the floor
    // cast.  The walk visits only cells the ray passes through (its rounding can shift a border crossing by
    // ~1e-10 cells, never by a cell), so a ray whose line stays more than 2 cells away from the centre of every
    // wall cell of the level cannot hit one; it skips the walk and stores the "nothing drawn" record.  NaN or
    // infinite operands fail the comparisons and take the walk.  (Not with PM3D_OPT_KEEP_COLUMNS: the
    // per-column debug record wants the exit cell.)
    if (!DEBUG && level > 0) {
        const int ns = __ldg(p.sparse_count + sl);
        if (ns >= 0) {
            const double *cells = p.sparse_cells + (size_t)sl * PM3D_SPARSE_MAX * 2;
            const double d2 = rayDirX * rayDirX + rayDirY * rayDirY;
            bool clear = d2 > 0.0 && d2 < 1e300 && fabs(rayPosX) < 1e9 && fabs(rayPosY) < 1e9;
            for (int i = 0; i < ns && clear; i++) {
                const double wx = __ldg(cells + 2 * i) - rayPosX, wy = __ldg(cells + 2 * i + 1) - rayPosY;
                const double dot = wx * rayDirX + wy * rayDirY;          // |w| |d| cos
                const double cross = wx * rayDirY - wy * rayDirX;        // |w| |d| sin: distance to the line = |cross| / |d|
                const bool far_from_line = cross * cross > 4.0 * d2;     // more than 2 cells off the line
                const bool behind = dot < 0.0 && wx * wx + wy * wy > 4.0;  // behind the camera and not next to it
                clear = far_from_line || behind;
            }
            if (clear) {
                const size_t ci = ((size_t)view * p.num_levels + level) * p.w + x;
                if (live) {
                    p.col_span[ci] = make_int2(0, 0);
                    p.col_tt[ci] = make_uint2(0u, 0u);
                    p.col_ps[ci] = make_ulonglong2(0ull, 0ull);
                }
                hit = 3;
            }
        }
    }
    // The start cell may be far outside the grid (Go ints are 64-bit); once the walk is
    // inside int32 range all cell arithmetic is 32-bit.
    if (mapX < -2147480000ll || mapX > 2147480000ll || mapY < -2147480000ll || mapY > 2147480000ll) {
        // the very first step leaves the grid: same outcome as the reference loop
        if (sideDistX < sideDistY) { mapX += stepX; side = 0; } else { mapY += stepY; side = 1; }
        hit = 2;
        mapX = mapX < 0 ? 0 : (mapX > p.map_w - 1 ? p.map_w - 1 : mapX);
        mapY = mapY < 0 ? 0 : (mapY > p.map_h - 1 ? p.map_h - 1 : mapY);
    }
    int agg_ya = 0x7fffffff, agg_fs_min = 0x7fffffff, agg_fs_max = 0;
    float agg_zmax = -INFINITY;  // level 0: upper bound of the z-buffer over the warp's columns (sprite culling)
    if (hit != 3) {  // (skipped rays hold the defaults: nothing drawn)
    int mx = (int)mapX, my = (int)mapY;
    const uint8_t *__restrict__ dist = p.dist + (size_t)sl * p.map_w * p.map_h;
    // free-run length at the start cell; the reference never tests the start cell itself, so a
    // camera inside a wall (D = 0) or outside the grid simply probes on its first step
    int D = 1;
    if (hit == 0 && (unsigned)mx < (unsigned)p.map_w && (unsigned)my < (unsigned)p.map_h) {
        D = __ldg(dist + mx * p.map_h + my);
        if (D == 0) D = 1;
    }
    // ---- jump mode.  Between two probes the steps cannot hit anything, so only the NUMBER of x and y steps among
    // them matters: dda_jump gives the state after N steps in closed form with a certificate that the reference's
    // sequence of rounded additions and comparisons reaches exactly that state.  The probe distance comes from the
    // quadrant field (steps to the nearest wall cell or map exit for a ray walking in this ray's quadrant), so an
    // open map costs a handful of jumps instead of ~700 dependent add/compare steps.  Any failed certificate
    // (near ties, non-finite operands) restarts the ray in the sequential loop below.
    if (p.jump && hit == 0 && (unsigned)mx < (unsigned)p.map_w && (unsigned)my < (unsigned)p.map_h) {
        const uint16_t *__restrict__ qd =
            p.qdist + ((size_t)sl * 4 + ((stepX > 0 ? 1 : 0) | (stepY > 0 ? 2 : 0))) * p.map_w * p.map_h;
        int Dq = __ldg(qd + mx * p.map_h + my);
        // axis-parallel rays: a side distance that is inf or NaN never wins (or always loses) a comparison
        const bool x_dead = !(sideDistX < INFINITY) || sideDistY != sideDistY;   // only y steps
        const bool y_dead = !x_dead && sideDistY == INFINITY && deltaDistX < 1e290;  // only x steps
        const double c0 = sideDistY - sideDistX, inv = 1.0 / (deltaDistX + deltaDistY);
        int N = 0, jx = mx, jy = my, jside = -1, jhit = 0;
        bool ok = Dq > 0;  // (a camera inside a wall cell: the sequential loop)
        while (ok) {
            N += Dq;
            int i, j;
            if (x_dead) { i = 0; j = N; }
            else if (y_dead) { i = N; j = 0; }
            else dda_jump_split(deltaDistY, c0, inv, N, i, j);
            // the landing cell's probe distance is requested BEFORE the certificate is evaluated (an in-range load is
            // harmless if the certificate fails): the L2 round trip overlaps ~10 dependent binary64 operations
            jx = mx + i * stepX; jy = my + j * stepY;
            const bool inside = (unsigned)jx < (unsigned)p.map_w && (unsigned)jy < (unsigned)p.map_h;
            const int Dn = inside ? (int)__ldg(qd + jx * p.map_h + jy) : 0;
            if (x_dead) jside = 1;
            else if (y_dead) jside = 0;
            else ok = dda_jump_certify(sideDistX, sideDistY, deltaDistX, deltaDistY, N, i, j, jside);
            if (!ok) break;
            if (inside) {
                Dq = Dn;
                if (Dq == 0) { jhit = 1; break; }
            } else {
                jhit = 2;
                jx = jx < 0 ? 0 : (jx > p.map_w - 1 ? p.map_w - 1 : jx);
                jy = jy < 0 ? 0 : (jy > p.map_h - 1 ? p.map_h - 1 : jy);
                break;
            }
        }
        if (ok) { mx = jx; my = jy; side = jside; hit = jhit; }
#ifdef PM3D_FRONT_STATS
        atomicAdd(&g_front_stats[1], 1ull);
        if (!ok) atomicAdd(&g_front_stats[0], 1ull);
#endif
    }
    while (hit == 0) {
        for (int k = D - 1; k > 0; --k) {  // blind steps: cannot hit, cannot leave
            if (sideDistX < sideDistY) { sideDistX += deltaDistX; mx += stepX; side = 0; }
            else                       { sideDistY += deltaDistY; my += stepY; side = 1; }
        }
        if (sideDistX < sideDistY) { sideDistX += deltaDistX; mx += stepX; side = 0; }
        else                       { sideDistY += deltaDistY; my += stepY; side = 1; }
        if ((unsigned)mx < (unsigned)p.map_w && (unsigned)my < (unsigned)p.map_h) {
            D = __ldg(dist + mx * p.map_h + my);
            if (D == 0) hit = 1;
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

    // ---- compose-side records
    int ya = 0, yb = 0;
    uint32_t tile_texx = 0;
    uint64_t pos0 = 0, step2 = 0;
    if (tile >= 0 && lineHeight > 0 && drawEnd > drawStart) {
        const uint64_t step = span_step(p.tex_size, lineHeight);
        ya = (int32_t)(drawStart < 0 ? 0 : (drawStart > h ? h : drawStart));
        yb = (int32_t)(drawEnd > h ? h : (drawEnd < 0 ? 0 : drawEnd));
        tile_texx = ((uint32_t)tile << 16) | (uint32_t)texX;
        step2 = step * 2ull;
        pos0 = (1ull - 2ull * (uint64_t)drawStart) * step;  // (2*(0 - drawStart) + 1) * step, mod 2^64
    }
    const size_t ci = ((size_t)view * p.num_levels + level) * p.w + x;
    if (live) {
        p.col_span[ci] = make_int2(ya, yb);
        p.col_tt[ci] = make_uint2(tile_texx, tint);
        p.col_ps[ci] = make_ulonglong2(pos0, step2);
    }
    agg_ya = (live && ya < yb) ? ya : 0x7fffffff;

    if (DEBUG && live) {
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

    if (level == 0) {
        // ---- floor-cast parameters of this column
        double fxw, fyw;
        if (side == 0 && rayDirX > 0)      { fxw = (double)mx;         fyw = (double)my + wallX; }
        else if (side == 0 && rayDirX < 0) { fxw = (double)mx + 1.0;   fyw = (double)my + wallX; }
        else if (side == 1 && rayDirY > 0) { fxw = (double)mx + wallX; fyw = (double)my; }
        else                               { fxw = (double)mx + wallX; fyw = (double)my + 1.0; }
        int64_t de = drawEnd;
        if (de < 0) de = h;  // "becomes < 0 when the integer overflows"
        const int fs = (int32_t)(de >= h ? h : de + 1);
        if (live) {
            const size_t fi = (size_t)view * p.w + x;
            p.zbuf[fi] = perpWallDist;
            p.flo_xw[fi] = fxw; p.flo_yw[fi] = fyw; p.flo_dw[fi] = perpWallDist; p.flo_fs[fi] = fs;
            // fast floor cast (compose_kernel.cu): texel coordinates t = P + w*(X1 - P) in texel units.  Eligible
            // columns have a mid-range wall distance and |coordinates| <= 2^24 texels (error bound in DESIGN.md)
            const double S = (double)p.tex_size, lim = 16777216.0;
            const double X1 = fxw * S, Y1 = fyw * S, PX = rayPosX * S, PY = rayPosY * S;
            const bool elig = exp_mid(perpWallDist) && fabs(X1) <= lim && fabs(Y1) <= lim && fabs(PX) <= lim && fabs(PY) <= lim;
            p.flo_rc[fi] = elig ? 1.0 / perpWallDist : 0.0;
            p.flo_dx[fi] = X1 - PX;
            p.flo_dy[fi] = Y1 - PY;
            agg_fs_min = fs; agg_fs_max = fs;
            agg_zmax = __double2float_ru(perpWallDist);  // NaN stays NaN and is ignored by fmaxf below
        }
    }
    }
    // ---- aggregate over this warp's 32 columns (half a group)
    for (int o = 16; o > 0; o >>= 1) {
        agg_ya = min(agg_ya, __shfl_xor_sync(0xffffffffu, agg_ya, o));
        agg_fs_min = min(agg_fs_min, __shfl_xor_sync(0xffffffffu, agg_fs_min, o));
        agg_fs_max = max(agg_fs_max, __shfl_xor_sync(0xffffffffu, agg_fs_max, o));
        agg_zmax = fmaxf(agg_zmax, __shfl_xor_sync(0xffffffffu, agg_zmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        const int half = (threadIdx.x >> 5) & 1;
        // [view][group][level][half]: the 2*L records of a group are contiguous for the compose warps that read them all
        p.agg[(((size_t)view * p.n_groups + group) * p.num_levels + level) * 2 + half] =
            make_int4(agg_ya, agg_fs_min, agg_fs_max, __float_as_int(agg_zmax));
    }
}
