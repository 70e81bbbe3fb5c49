// compose_kernel.cu — K2 floor cast + K3 Camera.Draw: the per-pixel pass that writes the
// RGBA8 framebuffer (call sites game/scene_game.go:161,167).
//
// Mapping: a warp covers one 64-column group x 2 rows (16 lanes x 4 adjacent pixels, so each
// lane ends a row with ONE 128-bit store and a warp store covers two 256 B row segments);
// a CTA = 4 warps = COMPOSE_ROWS consecutive rows of one group.  The group's wall-slice,
// floor and z-buffer records and the CTA's row table are staged in shared memory once.
//
// Layer order is the reference's painter's order (DESIGN.md "frame composition"):
//   sky / floor box -> walls (top level .. 0) -> textured floor -> sprites far .. near.
// Because every layer is premultiplied source-over with one rounding per layer, the result
// after the last OPAQUE texel does not depend on anything behind it.  The fast path
// therefore resolves layers front to back and stops at the first opaque texel; a pixel that
// meets a translucent texel (0 < a < 255) on the way is re-done by the literal painter loop
// (PM3D_OPT_FORCE_PAINTER forces that loop everywhere; tests require identical output).
#include "pm3d_internal.cuh"

#define GROUP_COLS 64
#define COMPOSE_WARPS 4
#define COMPOSE_ROWS 32  // rows per CTA (8 per warp-pair-row slot: 4 warps x 2 rows x 4 iterations)
#define MAX_BAND_SPRITES 256

struct ComposeParams {
    const pm3d_view *views;
    const ColRec *col_rec;
    const FloorRec *floor_rec;
    const double *zbuf;
    const RowRec *row_rec;
    const DrawRec *draw_rec;
    const int32_t *sprite_offset;
    const uint16_t *bin_list;
    const int32_t *bin_count;
    const uint32_t *atlas;
    const int16_t *floor_tiles;
    const uint32_t *sheet_pixels;
    uint8_t *out;
    int w, h, num_levels, tex_size, tex_shift, map_w, map_h, n_groups, sky_tile, floorbox_tile, force_painter;
};

struct Smem {
    FloorRec floor[GROUP_COLS];
    double zbuf[GROUP_COLS];
    RowRec rows[COMPOSE_ROWS];
    uint16_t box_u[GROUP_COLS];
    uint16_t band[MAX_BAND_SPRITES];
    int n_band;
    int band_overflow;
};

// ---- one sprite layer at pixel (x, y); returns texel (premultiplied) or 0 when it does not cover
__device__ __forceinline__ uint32_t sprite_texel(const DrawRec &d, const uint32_t *__restrict__ sheet_pixels, int x,
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

// ---- floor layer at (column record f, row record r): returns texel already tinted (>>8), 0 if none
__device__ __forceinline__ uint32_t floor_texel(const ComposeParams &p, const FloorRec &f, const RowRec &r,
                                                double posx, double posy, int y) {
    if (y < f.fs || !(r.info & 0x80000000u)) return 0u;
    const double weight = r.cd / f.dw;
    const double fx = weight * f.fxw + (1.0 - weight) * posx;
    const double fy = weight * f.fyw + (1.0 - weight) * posy;
    if (!(fx >= 0.0) || !(fy >= 0.0)) return 0u;
    if (!(fx < (double)p.map_w) || !(fy < (double)p.map_h)) return 0u;  // int(fx) >= map_w  <=>  fx >= map_w
    const int tx = (int)(fx * (double)p.tex_size), ty = (int)(fy * (double)p.tex_size);
    const int cx = tx >> p.tex_shift, cy = ty >> p.tex_shift;
    const int ft = __ldg(p.floor_tiles + cx * p.map_h + cy);
    if (ft < 0) return 0u;
    const int m = p.tex_size - 1;
    const uint32_t px = __ldg(p.atlas + ((((size_t)ft << p.tex_shift) + (size_t)(ty & m)) << p.tex_shift) + (size_t)(tx & m));
    if ((px >> 24) == 0) return 0u;
    const uint32_t t = r.tint;
    const uint32_t rr = ((px & 255u) * (t & 255u)) >> 8;
    const uint32_t gg = (((px >> 8) & 255u) * ((t >> 8) & 255u)) >> 8;
    const uint32_t bb = (((px >> 16) & 255u) * ((t >> 16) & 255u)) >> 8;
    return rr | (gg << 8) | (bb << 16) | (px & 0xff000000u);
}

__device__ __forceinline__ uint32_t wall_texel(const ComposeParams &p, const ColRec &c, int y) {
    uint32_t ty = (uint32_t)((c.pos0 + (uint64_t)(int64_t)y * c.step2) >> 33);
    if (ty > (uint32_t)p.tex_size - 1) ty = p.tex_size - 1;
    const uint32_t tile = c.tile_texx >> 16, tx = c.tile_texx & 0xffffu;
    return __ldg(p.atlas + ((((size_t)tile << p.tex_shift) + ty) << p.tex_shift) + tx);
}

__device__ __forceinline__ uint32_t box_texel(const ComposeParams &p, const RowRec &r, uint32_t u) {
    const int tile = (r.info & 0x40000000u) ? p.sky_tile : p.floorbox_tile;
    if (tile < 0) return 0u;
    const uint32_t v = r.info & 0xffffu;
    return __ldg(p.atlas + ((((size_t)tile << p.tex_shift) + v) << p.tex_shift) + u);
}

// literal painter's loop (the definition)
__device__ __noinline__ uint32_t pixel_painter(const ComposeParams &p, const Smem &s, const ColRec *scol, int lc,
                                               int x, int y, const RowRec &r, const DrawRec *draw, double posx,
                                               double posy, uint32_t box_tint) {
    uint32_t dst = 0u;
    dst = blend_over(dst, box_texel(p, r, s.box_u[lc]), box_tint);
    for (int l = p.num_levels - 1; l >= 0; l--) {
        const ColRec c = scol[l * GROUP_COLS + lc];
        if (y >= c.ya && y < c.yb) dst = blend_over(dst, wall_texel(p, c, y), c.tint);
    }
    dst = blend_over(dst, floor_texel(p, s.floor[lc], r, posx, posy, y), 0xffffffffu);
    const double z = s.zbuf[lc];
    for (int i = 0; i < s.n_band; i++) {
        const DrawRec &d = draw[s.band[i]];
        dst = blend_over(dst, sprite_texel(d, p.sheet_pixels, x, y, z), d.tint);
    }
    return dst;
}

// front-to-back resolve; sets *ok=false when a translucent texel was met
__device__ __forceinline__ uint32_t pixel_fast(const ComposeParams &p, const Smem &s, const ColRec *scol, int lc, int x,
                                               int y, const RowRec &r, const DrawRec *draw, double posx, double posy,
                                               uint32_t box_tint, bool *ok) {
    const double z = s.zbuf[lc];
    for (int i = s.n_band - 1; i >= 0; i--) {
        const DrawRec &d = draw[s.band[i]];
        const uint32_t t = sprite_texel(d, p.sheet_pixels, x, y, z);
        const uint32_t a = t >> 24;
        if (a == 255u) return blend_over(0u, t, d.tint);
        if (a != 0u) { *ok = false; return 0u; }
    }
    {
        const uint32_t t = floor_texel(p, s.floor[lc], r, posx, posy, y);
        const uint32_t a = t >> 24;
        if (a == 255u) return t;  // (t*255 + 127)/255 == t
        if (a != 0u) { *ok = false; return 0u; }
    }
    for (int l = 0; l < p.num_levels; l++) {
        const ColRec c = scol[l * GROUP_COLS + lc];
        if (y >= c.ya && y < c.yb) {
            const uint32_t t = wall_texel(p, c, y);
            const uint32_t a = t >> 24;
            if (a == 255u) return blend_over(0u, t, c.tint);
            if (a != 0u) { *ok = false; return 0u; }
        }
    }
    return blend_over(0u, box_texel(p, r, s.box_u[lc]), box_tint);
}

__global__ void __launch_bounds__(COMPOSE_WARPS * 32) k_compose(ComposeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    ColRec *scol = reinterpret_cast<ColRec *>(smem_raw + ((sizeof(Smem) + 15) & ~(size_t)15));

    const int group = blockIdx.x, view = blockIdx.z;
    const int row0 = blockIdx.y * COMPOSE_ROWS;
    const int tid = threadIdx.x;
    const int gx0 = group * GROUP_COLS;
    const pm3d_view &c = p.views[view];
    const double posx = c.pos_x, posy = c.pos_y;
    const uint32_t box_tint = (uint32_t)c.max_light[0] | ((uint32_t)c.max_light[1] << 8) |
                              ((uint32_t)c.max_light[2] << 16) | 0xff000000u;

    // ---- stage the group's records
    for (int i = tid; i < GROUP_COLS * p.num_levels; i += blockDim.x) {
        const int l = i / GROUP_COLS, lc = i - l * GROUP_COLS;
        ColRec r;
        r.ya = 0; r.yb = 0; r.tile_texx = 0; r.tint = 0; r.pos0 = 0; r.step2 = 0;
        if (gx0 + lc < p.w) r = p.col_rec[((size_t)view * p.num_levels + l) * p.w + gx0 + lc];
        scol[i] = r;
    }
    if (tid < GROUP_COLS) {
        const int x = gx0 + tid;
        if (x < p.w) {
            s.floor[tid] = p.floor_rec[(size_t)view * p.w + x];
            s.zbuf[tid] = p.zbuf[(size_t)view * p.w + x];
            uint32_t u = (uint32_t)((((uint64_t)(2 * (int64_t)x + 1)) * span_step(p.tex_size, p.w)) >> 33);
            if (u > (uint32_t)p.tex_size - 1) u = p.tex_size - 1;
            s.box_u[tid] = (uint16_t)u;
        } else {
            FloorRec f; f.fxw = 0; f.fyw = 0; f.dw = 1; f.fs = 0x7fffffff; f._pad = 0;
            s.floor[tid] = f;
            s.zbuf[tid] = 0.0;
            s.box_u[tid] = 0;
        }
    }
    if (tid < COMPOSE_ROWS) {
        const int y = row0 + tid;
        RowRec r; r.cd = 0; r.tint = 0; r.info = 0;
        if (y < p.h) r = p.row_rec[(size_t)view * p.h + y];
        s.rows[tid] = r;
    }
    // ---- band sprite list: entries of the group list whose rows intersect this CTA (order kept)
    const int s0 = p.sprite_offset[view];
    const int nv = p.sprite_offset[view + 1] - s0;
    const DrawRec *draw = p.draw_rec + s0;
    if (tid < 32) {
        const uint16_t *list = p.bin_list + (size_t)s0 * p.n_groups + (size_t)group * nv;
        const int cnt = p.bin_count[view * p.n_groups + group];
        int nb = 0, overflow = 0;
        const int y0 = row0, y1 = min(row0 + COMPOSE_ROWS, p.h);
        for (int base = 0; base < cnt; base += 32) {
            const int i = base + tid;
            bool keep = false;
            uint16_t id = 0;
            if (i < cnt) {
                id = list[i];
                const DrawRec &d = draw[id];
                keep = d.dey > y0 && d.dsy < y1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int pos = nb + __popc(m & ((1u << tid) - 1u));
            if (keep) {
                if (pos < MAX_BAND_SPRITES) s.band[pos] = id; else overflow = 1;
            }
            nb += __popc(m);
        }
        overflow = __any_sync(0xffffffffu, overflow);
        if (tid == 0) { s.n_band = min(nb, MAX_BAND_SPRITES); s.band_overflow = overflow; }
    }
    __syncthreads();

    // ---- pixels
    const int warp = tid >> 5, lane = tid & 31;
    const int lx = lane & 15, ly = lane >> 4;
    const int lc0 = lx * 4;
    const int x0 = gx0 + lc0;
    if (x0 >= p.w) return;
    uint32_t *out32 = reinterpret_cast<uint32_t *>(p.out) + (size_t)view * p.h * p.w;
    const bool vec_ok = ((p.w & 3) == 0);

#pragma unroll 1
    for (int it = 0; it < COMPOSE_ROWS / (COMPOSE_WARPS * 2); it++) {
        const int ry = (it * COMPOSE_WARPS + warp) * 2 + ly;
        const int y = row0 + ry;
        if (y >= p.h) break;
        const RowRec r = s.rows[ry];
        uint32_t px[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int lc = lc0 + j, x = x0 + j;
            uint32_t v = 0u;
            if (x < p.w) {
                bool ok = !p.force_painter && !s.band_overflow;
                if (ok) v = pixel_fast(p, s, scol, lc, x, y, r, draw, posx, posy, box_tint, &ok);
                if (!ok) {
                    if (s.band_overflow) {
                        // more sprites than the band list holds: walk the whole group list instead
                        uint32_t dst = 0u;
                        dst = blend_over(dst, box_texel(p, r, s.box_u[lc]), box_tint);
                        for (int l = p.num_levels - 1; l >= 0; l--) {
                            const ColRec cr = scol[l * GROUP_COLS + lc];
                            if (y >= cr.ya && y < cr.yb) dst = blend_over(dst, wall_texel(p, cr, y), cr.tint);
                        }
                        dst = blend_over(dst, floor_texel(p, s.floor[lc], r, posx, posy, y), 0xffffffffu);
                        const uint16_t *list = p.bin_list + (size_t)s0 * p.n_groups + (size_t)group * nv;
                        const int cnt = p.bin_count[view * p.n_groups + group];
                        const double z = s.zbuf[lc];
                        for (int i = 0; i < cnt; i++) {
                            const DrawRec &d = draw[list[i]];
                            dst = blend_over(dst, sprite_texel(d, p.sheet_pixels, x, y, z), d.tint);
                        }
                        v = dst;
                    } else {
                        v = pixel_painter(p, s, scol, lc, x, y, r, draw, posx, posy, box_tint);
                    }
                }
            }
            px[j] = v;
        }
        uint32_t *o = out32 + (size_t)y * p.w + x0;
        if (vec_ok && x0 + 3 < p.w) {
            *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
            for (int j = 0; j < 4; j++)
                if (x0 + j < p.w) o[j] = px[j];
        }
    }
}

// ---- per-view FNV-1a style checksum (order-independent combine of per-word hashes)
__global__ void k_checksum(const uint32_t *frames, uint64_t *out, size_t words_per_view) {
    const int view = blockIdx.x;
    const uint32_t *f = frames + (size_t)view * words_per_view;
    uint64_t acc = 0;
    for (size_t i = threadIdx.x; i < words_per_view; i += blockDim.x) {
        uint64_t hsh = 1469598103934665603ull ^ (uint64_t)i;
        hsh *= 1099511628211ull;
        hsh ^= f[i];
        hsh *= 1099511628211ull;
        acc += hsh;
    }
    __shared__ uint64_t red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if (threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[view] = red[0];
}

int pm3d_launch_compose(pm3d_ctx *ctx, int n_views, uint8_t *out) {
    ComposeParams p;
    p.views = ctx->cur_views;
    p.col_rec = ctx->col_rec;
    p.floor_rec = ctx->floor_rec;
    p.zbuf = ctx->zbuf;
    p.row_rec = ctx->row_rec;
    p.draw_rec = ctx->draw_rec;
    p.sprite_offset = ctx->cur_offset;
    p.bin_list = ctx->bin_list;
    p.bin_count = ctx->bin_count;
    p.atlas = reinterpret_cast<const uint32_t *>(ctx->atlas);
    p.floor_tiles = ctx->floor_tiles;
    p.sheet_pixels = reinterpret_cast<const uint32_t *>(ctx->sheet_pixels);
    p.out = out;
    p.w = ctx->w; p.h = ctx->h; p.num_levels = ctx->num_levels; p.tex_size = ctx->tex_size; p.tex_shift = ctx->tex_shift;
    p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.n_groups = ctx->n_groups;
    p.sky_tile = ctx->sky_tile; p.floorbox_tile = ctx->floorbox_tile; p.force_painter = ctx->force_painter;
    const size_t smem = ((sizeof(Smem) + 15) & ~(size_t)15) + (size_t)GROUP_COLS * ctx->num_levels * sizeof(ColRec);
    if (smem > 48 * 1024) {
        PM3D_CUDA(ctx, cudaFuncSetAttribute(k_compose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 block(COMPOSE_WARPS * 32), grid(ctx->n_groups, (ctx->h + COMPOSE_ROWS - 1) / COMPOSE_ROWS, n_views);
    pm3d_kt_mark(ctx, 2, 0);
    k_compose<<<grid, block, smem, ctx->stream>>>(p);
    pm3d_kt_mark(ctx, 2, 1);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}

int pm3d_launch_checksum(pm3d_ctx *ctx, int n_views, const uint8_t *frames) {
    k_checksum<<<n_views, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(frames), ctx->checksums,
                                                 (size_t)ctx->w * ctx->h);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
