// front_kernel.cu — the frame path's first launch: the per-column wall cast (cast_kernel.cuh) and the
// per-view setup (view_kernel.cuh) are independent of each other, so they run as CTAs of ONE grid:
// CTAs [0, n_views) set up the sprites of one view each (issued first: the longest CTAs of a single frame),
// the next n_cast CTAs cast four (view, level, 64-column group) units of 64 rays each,
// the rest build the per-row tables (VIEW_THREADS rows of a view per CTA).  The compose kernel consumes all three.
#include "cast_kernel.cuh"
#include "view_kernel.cuh"  // (after cast_kernel.cuh: PM3D_FRONT_STATS counters)

#ifdef PM3D_FRONT_STATS
// (experiments only: tools/exp_front_stats.py builds the library with -DPM3D_FRONT_STATS)
extern "C" int pm3d_debug_front_stats(pm3d_ctx *ctx, unsigned long long *out, int reset) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_front_stats, sizeof(g_front_stats));
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g_front_stats, z, sizeof(z)); }
    return 0;
}
#endif
#ifndef FRONT_MIN_BLOCKS
#define FRONT_MIN_BLOCKS 4
#endif
template <bool DEBUG>
__global__ void __launch_bounds__(VIEW_THREADS, FRONT_MIN_BLOCKS) k_front(CastParams cp, ViewParams vp, int n_cast_blocks, long long n_pairs, int n_views) {
#ifdef PM3D_FRONT_STATS
    const long long t0 = clock64();
#endif
    // the view setups come first in the grid: one CTA each, the longest-running ones of a single frame
    const int b = (int)blockIdx.x;
    if (b < n_views) {
        view_setup(vp, b);
#ifdef PM3D_FRONT_STATS
        atomicMax(&g_front_stats[3], (unsigned long long)(clock64() - t0));
#endif
    } else if (b < n_views + n_cast_blocks) {
        // a pair of warps = the 64 columns of one (view, level, group)
        const long long pair = (long long)(b - n_views) * (VIEW_THREADS / 64) + (threadIdx.x >> 6);
        if (pair >= n_pairs) return;
        const int group = (int)(pair % cp.n_groups);
        const long long vl = pair / cp.n_groups;
        cast_ray<DEBUG>(cp, (int)(vl / cp.num_levels), (int)(vl % cp.num_levels), group, group * GROUP_COLS + (threadIdx.x & 63));
#ifdef PM3D_FRONT_STATS
        atomicMax(&g_front_stats[2], (unsigned long long)(clock64() - t0));
#endif
    } else {
        // row tables: VIEW_THREADS rows of one view per CTA
        const int r = b - n_views - n_cast_blocks, row_blocks = (vp.h + VIEW_THREADS - 1) / VIEW_THREADS;
        view_rows(vp, r / row_blocks, (r % row_blocks) * VIEW_THREADS);
    }
}

int pm3d_launch_front(pm3d_ctx *ctx, int n_views, cudaStream_t stream) {
    CastParams p;
    p.views = ctx->cur_views;
    p.grid = ctx->grid;
    p.dist = ctx->dist;
    p.qdist = ctx->qdist;
    p.jump = ctx->dda_jump;
    p.sparse_cells = ctx->sparse_cells; p.sparse_count = ctx->sparse_count;
    p.wall_lut = ctx->wall_lut;
    p.col_span = ctx->col_span; p.col_tt = ctx->col_tt; p.col_ps = ctx->col_ps;
    p.flo_xw = ctx->flo_xw; p.flo_yw = ctx->flo_yw; p.flo_dw = ctx->flo_dw; p.flo_rc = ctx->flo_rc; p.flo_dx = ctx->flo_dx; p.flo_dy = ctx->flo_dy; p.flo_fs = ctx->flo_fs;
    p.agg = ctx->agg;
    p.zbuf = ctx->zbuf;
    p.dbg = ctx->keep_columns ? ctx->dbg_cols : nullptr;
    p.w = ctx->w; p.h = ctx->h; p.map_w = ctx->map_w; p.map_h = ctx->map_h;
    p.stored_levels = ctx->stored_levels; p.num_levels = ctx->num_levels; p.tex_size = ctx->tex_size; p.n_groups = ctx->n_groups;

    ViewParams v;
    v.views = ctx->cur_views;
    v.row_rec = ctx->row_rec;
    v.sprites = ctx->cur_sprites;
    v.sprite_offset = ctx->cur_offset;
    v.sheets = ctx->sheet_info;
    v.draw_rec = ctx->draw_rec;
    v.n_draw = ctx->n_draw;
    v.bin_list = ctx->bin_list;
    v.bin_rect = ctx->bin_rect;
    v.bin_count = ctx->bin_count;
    v.rects = ctx->rects;
    v.conv = ctx->conv;
    v.w = ctx->w; v.h = ctx->h; v.tex_size = ctx->tex_size; v.n_sheets = ctx->n_sheets; v.n_groups = ctx->n_groups;
    v.sky_tile = ctx->sky_tile; v.floorbox_tile = ctx->floorbox_tile;
    v.status = ctx->status_d; v.n_sprites_total = ctx->cur_n_sprites;

    const long long n_pairs = (long long)n_views * ctx->num_levels * ctx->n_groups;
    const long long n_cast = (n_pairs + VIEW_THREADS / 64 - 1) / (VIEW_THREADS / 64);
    const long long row_blocks = (long long)n_views * ((ctx->h + VIEW_THREADS - 1) / VIEW_THREADS);
    if (n_cast + n_views + row_blocks > 0x7fffffffll) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "too many rays for one launch");
    const unsigned grid = (unsigned)(n_cast + n_views + row_blocks);
    pm3d_kt_mark(ctx, 0, 0, stream);
    if (p.dbg) k_front<true><<<grid, VIEW_THREADS, 0, stream>>>(p, v, (int)n_cast, n_pairs, n_views);
    else k_front<false><<<grid, VIEW_THREADS, 0, stream>>>(p, v, (int)n_cast, n_pairs, n_views);
    pm3d_kt_mark(ctx, 0, 1, stream);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
