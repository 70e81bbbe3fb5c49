// pm3d_api.cu — the C-ABI of libpm3d.so (include/pm3d.h): ctx lifetime, mission-time uploads,
// the frame path orchestration and the device-memory / timing helpers.  Host code only; the
// kernels live in cast_kernel.cu, view_kernel.cu, compose_kernel.cu and los_kernel.cu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pm3d_internal.cuh"

static thread_local std::string g_create_error;

int pm3d_fail(pm3d_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

// a recorded frame sequence holds the device pointers of the configuration it was captured under: any reallocation
// of mission data or frame buffers retires the recordings (pm3d_frames_replay then fails with PM3D_ERR_ARG)
static int retire_frame_graphs(pm3d_ctx *ctx) {
    if (ctx->capturing) return pm3d_fail(ctx, PM3D_ERR_ARG, "uploads and reconfiguration are not allowed inside a frame capture");
    bool any = false;
    for (auto &fg : ctx->graphs) any = any || fg.exec;
    if (!any) return PM3D_OK;
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &fg : ctx->graphs)
        if (fg.exec) { cudaGraphExecDestroy(fg.exec); fg.exec = nullptr; }
    return PM3D_OK;
}

template <typename T>
static int dev_realloc(pm3d_ctx *ctx, T **ptr, size_t count) {
    { int rc_ = retire_frame_graphs(ctx); if (rc_) return rc_; }
    if (*ptr) { PM3D_CUDA(ctx, cudaFree(*ptr)); *ptr = nullptr; }
    if (count == 0) count = 1;
    PM3D_CUDA(ctx, cudaMalloc((void **)ptr, count * sizeof(T)));
    return PM3D_OK;
}

// ---- small results of pipelined host renders (see pm3d_ctx::SmallPending)
__global__ void k_export_small(const uint4 *rects, int n_rects, const uint4 *conv, int n_conv16, uint4 *dst, size_t rect_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rects) dst[i] = rects[i];
    if (i < n_conv16) dst[rect_cap + i] = conv[i];
}
static void small_flush(pm3d_ctx *ctx, int s) {
    pm3d_ctx::SmallPending &sp = ctx->small[s];
    if (!sp.active) return;
    const uint8_t *src = ctx->small_ring + (size_t)s * ctx->small_slot_bytes;
    if (sp.dst_rects && sp.n_rects) memcpy(sp.dst_rects, src, (size_t)sp.n_rects * sizeof(pm3d_rect));
    if (sp.dst_conv && sp.n_conv) memcpy(sp.dst_conv, src + ctx->small_rect_cap * sizeof(pm3d_rect), (size_t)sp.n_conv * sizeof(pm3d_convergence));
    sp.active = false;
}
// after the main stream was synchronised: oldest first, so a buffer handed out twice ends with the newest result
static void small_flush_all(pm3d_ctx *ctx) {
    for (int k = 0; k < PM3D_SMALL_SLOTS; k++) small_flush(ctx, (ctx->small_pos + k) % PM3D_SMALL_SLOTS);
}

// every stream a frame batch may still be running on
static cudaError_t sync_frame_streams(pm3d_ctx *ctx) {
    cudaError_t e = cudaSuccess, r;
    if (ctx->front_stream && (r = cudaStreamSynchronize(ctx->front_stream)) != cudaSuccess) e = r;
    if (ctx->front_stream_lo && (r = cudaStreamSynchronize(ctx->front_stream_lo)) != cudaSuccess) e = r;
    for (int l = 1; l < PM3D_MAX_LANES; l++) {
        if (ctx->lane_fstream[l] && (r = cudaStreamSynchronize(ctx->lane_fstream[l])) != cudaSuccess) e = r;
        if (ctx->lane_cstream[l] && (r = cudaStreamSynchronize(ctx->lane_cstream[l])) != cudaSuccess) e = r;
    }
    if ((r = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) e = r;
    return e;
}
// readers of the last frame's results (rects, convergence, columns, checksum, presentation) enqueue on the main
// stream: when that frame ran on another lane they wait for it first
static cudaError_t wait_last_frame(pm3d_ctx *ctx) {
    if (ctx->last_cstream && ctx->last_cstream != ctx->stream) return cudaStreamSynchronize(ctx->last_cstream);
    return cudaSuccess;
}
int pm3d_wait_last_frame(pm3d_ctx *ctx) { return pm3d_main_enter(ctx); }  // a stream-side join, no host wait

// frame -> main: non-frame work enqueued on the main stream follows every frame dealt to another lane
int pm3d_main_enter(pm3d_ctx *ctx) {
    if (ctx->capturing) return pm3d_fail(ctx, PM3D_ERR_ARG, "only pm3d_render_resident may be called between pm3d_frames_capture_begin and pm3d_frames_capture_end");
    if (!ctx->lanes_dirty) return PM3D_OK;
    for (int l = 1; l < PM3D_MAX_LANES; l++) {
        if (!ctx->lane_cstream[l]) continue;
        PM3D_CUDA(ctx, cudaEventRecord(ctx->lane_join[l], ctx->lane_cstream[l]));
        PM3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_join[l], 0));
    }
    ctx->lanes_dirty = false;
    return PM3D_OK;
}
// main -> frame: the front stream of the frame being enqueued waits for everything the main stream holds, once
// per change of the main stream (a loop of frame calls alone never records or waits)
static int order_after_main(pm3d_ctx *ctx, cudaStream_t s, uint64_t *seen) {
    if (!ctx->own_stream) ctx->main_seq++;  // the caller may have enqueued anything on its stream
    if (s == ctx->stream || *seen == ctx->main_seq) return PM3D_OK;
    PM3D_CUDA(ctx, cudaEventRecord(ctx->ev_main, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_main, 0));
    *seen = ctx->main_seq;
    return PM3D_OK;
}
// device-side error words (mapped pinned memory), read once the streams have drained
int pm3d_check_status(pm3d_ctx *ctx) {
    if (!ctx->status_h) return PM3D_OK;
    volatile uint32_t *st = ctx->status_h;
    const uint32_t bad_view = st[PM3D_STATUS_VIEW_SPRITES], bad_list = st[PM3D_STATUS_LIST_TOTAL];
    if (!bad_view && !bad_list) return PM3D_OK;
    st[PM3D_STATUS_VIEW_SPRITES] = 0; st[PM3D_STATUS_LIST_TOTAL] = 0;
    if (bad_view)
        return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "view %u: sprite_offset entries out of range (negative, not monotone, more than 2048 sprites "
                         "in the view, or past n_sprites_total); the view was rendered without sprites", bad_view - 1);
    return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "pm3d_raycast_sprites_batch listed more sprites than max_total; the lists were cut");
}

extern "C" {

int pm3d_version(void) { return PM3D_VERSION; }

int pm3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *pm3d_last_error(const pm3d_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pm3d_create(int device, void *stream, pm3d_ctx **out) {
    if (!out) return pm3d_fail(nullptr, PM3D_ERR_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return pm3d_fail(nullptr, PM3D_ERR_CUDA, "no CUDA device (%s); libpm3d has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return pm3d_fail(nullptr, PM3D_ERR_ARG, "device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return pm3d_fail(nullptr, PM3D_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    pm3d_ctx *ctx = new pm3d_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        // work the caller orders on its own stream must precede every kernel of a frame: one stream, one set
        ctx->stream = (cudaStream_t)stream;
        ctx->n_sets = 1;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return pm3d_fail(nullptr, PM3D_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    if ((e = cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&ctx->status_h, PM3D_STATUS_WORDS * sizeof(uint32_t), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer((void **)&ctx->status_d, ctx->status_h, 0)) != cudaSuccess) {
        pm3d_destroy(ctx);
        return pm3d_fail(nullptr, PM3D_ERR_CUDA, "ctx ordering event / status words: %s", cudaGetErrorString(e));
    }
    memset(ctx->status_h, 0, PM3D_STATUS_WORDS * sizeof(uint32_t));
    *out = ctx;
    return PM3D_OK;
}

void pm3d_destroy(pm3d_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    sync_frame_streams(ctx);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    pm3d_comm_destroy(ctx);
    for (auto &fg : ctx->graphs)
        if (fg.exec) cudaGraphExecDestroy(fg.exec);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (cudaEvent_t e : ctx->ev_join)
        if (e) cudaEventDestroy(e);
    void *ptrs[] = {ctx->grid, ctx->dist, ctx->qdist, ctx->wall_lut, ctx->atlas, ctx->floor_tiles, ctx->sheet_pixels, ctx->sheet_info,
                    ctx->views, ctx->sprites, ctx->sprite_offset, ctx->frame, ctx->checksums, ctx->segs, ctx->seg_irect,
                    ctx->cell_start, ctx->cell_items, ctx->cell_rec, ctx->cell_code, ctx->los_list, ctx->los_count, ctx->path_scratch, ctx->colliders, ctx->cb_start, ctx->cb_items, ctx->cb_loose, ctx->present_map, ctx->present_buf, ctx->sparse_cells, ctx->sparse_count, ctx->clutter_class, ctx->clutter_tab, ctx->io_a, ctx->io_b, ctx->io_c, ctx->flush_buf};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto &fs : ctx->sets) {
        for (void *p : fs.p)
            if (p) cudaFree(p);
        if (fs.inbuf) cudaFree(fs.inbuf);
        if (fs.stage) cudaFreeHost(fs.stage);
        if (fs.ev_stage) cudaEventDestroy(fs.ev_stage);
        if (fs.ev_front) cudaEventDestroy(fs.ev_front);
        if (fs.ev_compose) cudaEventDestroy(fs.ev_compose);
    }
    if (ctx->front_stream) cudaStreamDestroy(ctx->front_stream);
    if (ctx->front_stream_lo) cudaStreamDestroy(ctx->front_stream_lo);
    for (int l = 1; l < PM3D_MAX_LANES; l++) {
        if (ctx->lane_fstream[l]) cudaStreamDestroy(ctx->lane_fstream[l]);
        if (ctx->lane_cstream[l]) cudaStreamDestroy(ctx->lane_cstream[l]);
        if (ctx->lane_join[l]) cudaEventDestroy(ctx->lane_join[l]);
    }
    for (int i = 0; i < 16; i++) {
        if (ctx->ev_start[i]) cudaEventDestroy(ctx->ev_start[i]);
        if (ctx->ev_stop[i]) cudaEventDestroy(ctx->ev_stop[i]);
    }
    for (int i = 0; i < 4; i++) {
        if (ctx->frame_ring[i]) cudaFree(ctx->frame_ring[i]);
        if (ctx->ev_rendered[i]) cudaEventDestroy(ctx->ev_rendered[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->small_ring) cudaFreeHost(ctx->small_ring);
    if (ctx->status_h) cudaFreeHost(ctx->status_h);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    for (auto &sp : ctx->small)
        if (sp.done) cudaEventDestroy(sp.done);
    for (int i = 0; i < ctx->kt_cap * 6; i++) cudaEventDestroy(ctx->kt_ev[i]);
    delete[] ctx->kt_ev;
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pm3d_sync(pm3d_ctx *ctx) {
    if (!ctx) return PM3D_ERR_ARG;
    if (ctx->capturing) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_sync inside a frame capture");
    PM3D_CUDA(ctx, sync_frame_streams(ctx));
    if (ctx->copy_stream) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    small_flush_all(ctx);
    ctx->lanes_dirty = false;
    return pm3d_check_status(ctx);
}

// ------------------------------------------------------------------ uploads

int pm3d_upload_grid(pm3d_ctx *ctx, const uint16_t *grid, int stored_levels, int map_w, int map_h,
                     int num_raycast_levels) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!grid || stored_levels < 1 || map_w < 1 || map_h < 1 || num_raycast_levels < 1 || map_w > 32768 ||
        map_h > 32768 || num_raycast_levels > 64)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "bad grid dimensions");
    size_t count = (size_t)stored_levels * map_w * map_h;
    for (size_t i = 0; i < count; i++)
        if (grid[i] > 32767) return pm3d_fail(ctx, PM3D_ERR_ARG, "grid cell %zu holds texture index %u > 32767 (math.MaxInt16 is the largest the reference uses)", i, (unsigned)grid[i]);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dev_realloc(ctx, &ctx->grid, count);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->grid, grid, count * 2, cudaMemcpyHostToDevice, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // free-run distance field (cast_kernel.cu): Chebyshev distance to the nearest wall cell or to the
    // outside of the map, capped at 255; two-pass 8-neighbour chamfer, exact for the L-inf metric
    {
        std::vector<uint8_t> dist(count);
        const int W = map_w, H = map_h;
        for (int l = 0; l < stored_levels; l++) {
            const uint16_t *g = grid + (size_t)l * W * H;
            uint8_t *d = dist.data() + (size_t)l * W * H;
            auto at = [&](int x, int y) -> int { return (x < 0 || y < 0 || x >= W || y >= H) ? 0 : d[(size_t)x * H + y]; };
            for (int x = 0; x < W; x++)
                for (int y = 0; y < H; y++) {
                    int v = 255;
                    if (g[(size_t)x * H + y] > 0) v = 0;
                    else {
                        v = std::min(v, at(x - 1, y - 1) + 1); v = std::min(v, at(x - 1, y) + 1);
                        v = std::min(v, at(x - 1, y + 1) + 1); v = std::min(v, at(x, y - 1) + 1);
                        // (x, y+1) and the x+1 column are not final yet: they may only be read as "outside"
                        if (y + 1 >= H || x + 1 >= W) v = std::min(v, 1);
                    }
                    d[(size_t)x * H + y] = (uint8_t)v;
                }
            for (int x = W - 1; x >= 0; x--)
                for (int y = H - 1; y >= 0; y--) {
                    int v = d[(size_t)x * H + y];
                    if (v == 0) continue;
                    v = std::min(v, at(x + 1, y + 1) + 1); v = std::min(v, at(x + 1, y) + 1);
                    v = std::min(v, at(x + 1, y - 1) + 1); v = std::min(v, at(x, y + 1) + 1);
                    d[(size_t)x * H + y] = (uint8_t)v;
                }
        }
        rc = dev_realloc(ctx, &ctx->dist, count);
        if (rc) return rc;
        PM3D_CUDA(ctx, cudaMemcpy(ctx->dist, dist.data(), count, cudaMemcpyHostToDevice));
    }
    // jump field (cast_kernel.cuh): per step quadrant q = (stepX > 0) | (stepY > 0) << 1, the number of DDA steps
    // a ray walking in that quadrant needs at least to reach a wall cell or leave the map: F = 0 on walls and
    // outside, else 1 + min(F(x + stepX, y), F(x, y + stepY)) — the L1 distance to the nearest wall cell inside the
    // quadrant's cone (every step moves one cell along one axis, never against the quadrant)
    {
        const int W = map_w, H = map_h;
        std::vector<uint16_t> q((size_t)stored_levels * 4 * W * H);
        for (int l = 0; l < stored_levels; l++) {
            const uint16_t *g = grid + (size_t)l * W * H;
            for (int quad = 0; quad < 4; quad++) {
                uint16_t *f = q.data() + ((size_t)l * 4 + quad) * W * H;
                const int sx = (quad & 1) ? 1 : -1, sy = (quad & 2) ? 1 : -1;
                auto at = [&](int x, int y) -> int { return (x < 0 || y < 0 || x >= W || y >= H) ? 0 : f[(size_t)x * H + y]; };
                for (int xi = 0; xi < W; xi++) {
                    const int x = sx > 0 ? W - 1 - xi : xi;
                    for (int yi = 0; yi < H; yi++) {
                        const int y = sy > 0 ? H - 1 - yi : yi;
                        int v = 0;
                        if (g[(size_t)x * H + y] == 0) v = std::min(1 + std::min(at(x + sx, y), at(x, y + sy)), 65535);
                        f[(size_t)x * H + y] = (uint16_t)v;
                    }
                }
            }
        }
        if ((rc = dev_realloc(ctx, &ctx->qdist, q.size()))) return rc;
        PM3D_CUDA(ctx, cudaMemcpy(ctx->qdist, q.data(), q.size() * 2, cudaMemcpyHostToDevice));
    }
    // sparse levels (cast_kernel.cuh): levels with at most PM3D_SPARSE_MAX wall cells list their cell centres, so
    // an upper-level ray that provably passes none of them skips the walk across the empty level
    {
        std::vector<double> cells((size_t)stored_levels * PM3D_SPARSE_MAX * 2, 0.0);
        std::vector<int32_t> cnt(stored_levels, -1);
        for (int l = 0; l < stored_levels; l++) {
            const uint16_t *g = grid + (size_t)l * map_w * map_h;
            int n = 0;
            for (int x = 0; x < map_w && n <= PM3D_SPARSE_MAX; x++)
                for (int y = 0; y < map_h && n <= PM3D_SPARSE_MAX; y++)
                    if (g[(size_t)x * map_h + y] > 0) {
                        if (n < PM3D_SPARSE_MAX) {
                            cells[((size_t)l * PM3D_SPARSE_MAX + n) * 2] = x + 0.5;
                            cells[((size_t)l * PM3D_SPARSE_MAX + n) * 2 + 1] = y + 0.5;
                        }
                        n++;
                    }
            cnt[l] = n <= PM3D_SPARSE_MAX ? n : -1;
        }
        if ((rc = dev_realloc(ctx, &ctx->sparse_cells, cells.size()))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->sparse_count, (size_t)stored_levels))) return rc;
        PM3D_CUDA(ctx, cudaMemcpy(ctx->sparse_cells, cells.data(), cells.size() * 8, cudaMemcpyHostToDevice));
        PM3D_CUDA(ctx, cudaMemcpy(ctx->sparse_count, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice));
    }
    // the clutter classes are per cell of the map they were uploaded for
    if (ctx->map_w != map_w || ctx->map_h != map_h) ctx->clutter_ready = false;
    ctx->stored_levels = stored_levels; ctx->map_w = map_w; ctx->map_h = map_h;
    pm3d_main_touch(ctx);
    const bool levels_changed = ctx->num_levels != num_raycast_levels;
    ctx->num_levels = num_raycast_levels;
    if (levels_changed && ctx->w > 0) return pm3d_set_view(ctx, ctx->w, ctx->h, ctx->max_views, ctx->max_sprites);
    return PM3D_OK;
}

int pm3d_upload_wall_textures(pm3d_ctx *ctx, const int32_t *tex_index, const int32_t *tile_side0,
                              const int32_t *tile_side1, int n) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || (n > 0 && (!tex_index || !tile_side0 || !tile_side1))) return pm3d_fail(ctx, PM3D_ERR_ARG, "null LUT");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int16_t> lut(32768 * 2, (int16_t)-1);
    ctx->max_wall_tile = -1;
    for (int i = 0; i < n; i++) {
        if (tex_index[i] < 0 || tex_index[i] > 32767) return pm3d_fail(ctx, PM3D_ERR_ARG, "texture index %d out of [0,32767]", tex_index[i]);
        if (tile_side0[i] < -1 || tile_side0[i] > 32767 || tile_side1[i] < -1 || tile_side1[i] > 32767)
            return pm3d_fail(ctx, PM3D_ERR_ARG, "tile out of range");
        lut[tex_index[i] * 2 + 0] = (int16_t)tile_side0[i];
        lut[tex_index[i] * 2 + 1] = (int16_t)tile_side1[i];
        ctx->max_wall_tile = std::max(ctx->max_wall_tile, std::max(tile_side0[i], tile_side1[i]));
    }
    int rc = dev_realloc(ctx, &ctx->wall_lut, lut.size());
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaMemcpy(ctx->wall_lut, lut.data(), lut.size() * 2, cudaMemcpyHostToDevice));
    return PM3D_OK;
}

int pm3d_upload_atlas(pm3d_ctx *ctx, const uint8_t *rgba, int n_tiles, int tex_size) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!rgba || n_tiles < 1 || n_tiles > 32767 || tex_size < 1 || tex_size > 4096 || (tex_size & (tex_size - 1)))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "atlas: tex_size must be a power of two <= 4096, n_tiles in [1,32767]");
    if ((uint64_t)n_tiles * tex_size * tex_size > (1ull << 32))
        return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "atlas of %d tiles of %dx%d exceeds 2^32 texels (the kernels index it with 32 bits)", n_tiles, tex_size, tex_size);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = (size_t)n_tiles * tex_size * tex_size * 4;
    int rc = dev_realloc(ctx, &ctx->atlas, bytes);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaMemcpy(ctx->atlas, rgba, bytes, cudaMemcpyHostToDevice));
    ctx->n_tiles = n_tiles; ctx->tex_size = tex_size;
    ctx->tex_shift = 0;
    while ((1 << ctx->tex_shift) < tex_size) ctx->tex_shift++;
    return PM3D_OK;
}

int pm3d_upload_floor_tiles(pm3d_ctx *ctx, const int16_t *tiles, int map_w, int map_h) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!tiles || map_w < 1 || map_h < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "bad floor grid");
    ctx->max_floor_tile = -1;
    for (size_t i = 0; i < (size_t)map_w * map_h; i++) ctx->max_floor_tile = std::max(ctx->max_floor_tile, (int)tiles[i]);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dev_realloc(ctx, &ctx->floor_tiles, (size_t)map_w * map_h);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaMemcpy(ctx->floor_tiles, tiles, (size_t)map_w * map_h * 2, cudaMemcpyHostToDevice));
    ctx->floor_w = map_w; ctx->floor_h = map_h;
    return PM3D_OK;
}

int pm3d_set_box_tiles(pm3d_ctx *ctx, int sky_tile, int floorbox_tile) {
    if (!ctx) return PM3D_ERR_ARG;
    ctx->sky_tile = sky_tile; ctx->floorbox_tile = floorbox_tile;
    return PM3D_OK;
}

int pm3d_upload_sprite_sheets(pm3d_ctx *ctx, const uint8_t *pixels, const int64_t *byte_offset, const int32_t *width,
                              const int32_t *height, int n_sheets) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n_sheets < 0 || (n_sheets > 0 && (!pixels || !byte_offset || !width || !height)))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "null sheet arrays");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t total = 0;
    std::vector<SheetInfo> info(std::max(n_sheets, 1));
    for (int i = 0; i < n_sheets; i++) {
        if (width[i] < 1 || height[i] < 1 || byte_offset[i] < 0 || (byte_offset[i] & 3))
            return pm3d_fail(ctx, PM3D_ERR_ARG, "sheet %d: bad size/offset", i);
        info[i].px_offset = byte_offset[i] / 4;
        info[i].w = width[i]; info[i].h = height[i];
        total = std::max(total, (size_t)byte_offset[i] + (size_t)width[i] * height[i] * 4);
    }
    int rc = dev_realloc(ctx, &ctx->sheet_pixels, std::max<size_t>(total, 4));
    if (rc) return rc;
    if (total) PM3D_CUDA(ctx, cudaMemcpy(ctx->sheet_pixels, pixels, total, cudaMemcpyHostToDevice));
    rc = dev_realloc(ctx, &ctx->sheet_info, info.size());
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaMemcpy(ctx->sheet_info, info.data(), info.size() * sizeof(SheetInfo), cudaMemcpyHostToDevice));
    ctx->n_sheets = n_sheets;
    return PM3D_OK;
}

int pm3d_set_option(pm3d_ctx *ctx, int option, int value) {
    if (!ctx) return PM3D_ERR_ARG;
    switch (option) {
    case PM3D_OPT_KEEP_COLUMNS:
        ctx->keep_columns = value ? 1 : 0;
        if (ctx->keep_columns && ctx->w > 0) return pm3d_set_view(ctx, ctx->w, ctx->h, ctx->max_views, ctx->max_sprites);
        return PM3D_OK;
    case PM3D_OPT_FORCE_PAINTER:
        ctx->force_painter = value ? 1 : 0;
        return PM3D_OK;
    case PM3D_OPT_DDA_JUMP:
        ctx->dda_jump = value ? 1 : 0;
        return PM3D_OK;
    case PM3D_OPT_OVERLAP:
        PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
        PM3D_CUDA(ctx, sync_frame_streams(ctx));
        ctx->n_sets = value ? 2 : 1;
        ctx->cur_set = 0;
        if (ctx->w > 0) return pm3d_set_view(ctx, ctx->w, ctx->h, ctx->max_views, ctx->max_sprites);
        return PM3D_OK;
    case PM3D_OPT_LANES:
        if (value < 1 || value > PM3D_MAX_LANES) return pm3d_fail(ctx, PM3D_ERR_ARG, "lanes must be 1..%d", PM3D_MAX_LANES);
        if (!ctx->own_stream && value > 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "a ctx on a caller stream has one lane");
        PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
        PM3D_CUDA(ctx, sync_frame_streams(ctx));
        ctx->n_lanes = value;
        ctx->lane_pos = 0;
        if (ctx->w > 0) return pm3d_set_view(ctx, ctx->w, ctx->h, ctx->max_views, ctx->max_sprites);
        return PM3D_OK;
    case PM3D_OPT_PIPELINE: {
        if (value < 1 || value > 4) return pm3d_fail(ctx, PM3D_ERR_ARG, "pipeline depth must be 1..4");
        if (ctx->w < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_set_view must precede PM3D_OPT_PIPELINE");
        PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
        PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->copy_stream) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        for (int i = 0; i < 4; i++)
            if (ctx->frame_ring[i]) { PM3D_CUDA(ctx, cudaFree(ctx->frame_ring[i])); ctx->frame_ring[i] = nullptr; }
        ctx->ring_n = value; ctx->ring_pos = 0;
        if (value > 1) {
            if (!ctx->copy_stream) PM3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            const size_t fb = (size_t)ctx->max_views * ctx->w * ctx->h * 4;
            for (int i = 0; i < value; i++) {
                PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->frame_ring[i], fb));
                if (!ctx->ev_rendered[i]) {
                    PM3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rendered[i], cudaEventDisableTiming));
                    PM3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
                }
            }
        }
        return PM3D_OK;
    }
    }
    return pm3d_fail(ctx, PM3D_ERR_ARG, "unknown option %d", option);
}

int pm3d_set_view(pm3d_ctx *ctx, int width, int height, int max_views, int max_sprites_total) {
    if (!ctx) return PM3D_ERR_ARG;
    if (width < 1 || height < 1 || width > 16384 || height > 16384 || max_views < 1 || max_views > 65535 ||
        max_sprites_total < 0)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "bad view size / capacity");
    if (ctx->num_levels < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_grid must precede pm3d_set_view");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, sync_frame_streams(ctx));
    if (ctx->copy_stream) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    small_flush_all(ctx);
    ctx->w = width; ctx->h = height; ctx->max_views = max_views; ctx->max_sprites = max_sprites_total;
    ctx->n_groups = (width + 63) / 64;
    const size_t V = max_views, L = ctx->num_levels, W = width, H = height, S = std::max(max_sprites_total, 1);
    int rc;
    if ((rc = dev_realloc(ctx, &ctx->views, V))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->sprites, S))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->sprite_offset, V + 1))) return rc;
    if (!ctx->front_stream) {
        // highest priority: the front launch is a handful of long-running CTAs; they must get SM slots as soon as any
        // free up instead of queueing behind the thousands of CTAs of the compose launch they overlap with
        int prio_lo = 0, prio_hi = 0;
        PM3D_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        PM3D_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->front_stream, cudaStreamNonBlocking, prio_hi));
        PM3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->front_stream_lo, cudaStreamNonBlocking));
    }
    for (int l = 1; l < ctx->n_lanes; l++) {
        if (ctx->lane_cstream[l]) continue;
        int prio_lo = 0, prio_hi = 0;
        PM3D_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        PM3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lane_cstream[l], cudaStreamNonBlocking));
        PM3D_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->lane_fstream[l], cudaStreamNonBlocking, prio_hi));
        PM3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_join[l], cudaEventDisableTiming));
    }
    const int sets_in_use = ctx->n_sets > 1 ? 2 * ctx->n_lanes : 1;
    for (int k = 0; k < 2 * PM3D_MAX_LANES; k++) {
        pm3d_ctx::FrameSet &fs = ctx->sets[k];
        pm3d_load_set(ctx, k);
        if (k >= sets_in_use) {  // PM3D_OPT_OVERLAP = 0 / fewer lanes: the other sets are released
#define X(f) if (ctx->f) { PM3D_CUDA(ctx, cudaFree(ctx->f)); ctx->f = nullptr; }
            PM3D_FRAME_FIELDS(X)
#undef X
            pm3d_save_set(ctx, k);
            continue;
        }
        auto alloc_set = [&]() -> int {
            if ((rc = dev_realloc(ctx, &ctx->col_span, V * L * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->col_tt, V * L * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->col_ps, V * L * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_xw, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_yw, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_dw, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_rc, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_dx, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_dy, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->flo_fs, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->agg, V * L * ctx->n_groups * 2))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->zbuf, V * W))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->row_rec, V * H))) return rc;
            if (ctx->keep_columns) {
                if ((rc = dev_realloc(ctx, &ctx->dbg_cols, V * L * W))) return rc;
                ctx->dbg_levels = (int)L; ctx->dbg_views = (int)V;
            }
            if ((rc = dev_realloc(ctx, &ctx->draw_rec, S))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->n_draw, V))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->bin_list, S * ctx->n_groups))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->bin_rect, S * ctx->n_groups))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->bin_count, V * ctx->n_groups))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->rects, S))) return rc;
            if ((rc = dev_realloc(ctx, &ctx->conv, V))) return rc;
            return PM3D_OK;
        };
        rc = alloc_set();
        pm3d_save_set(ctx, k);  // on failure too: the set must not keep pointers dev_realloc already freed
        if (rc) return rc;
        if (!fs.ev_front) {
            PM3D_CUDA(ctx, cudaEventCreateWithFlags(&fs.ev_front, cudaEventDisableTiming));
            PM3D_CUDA(ctx, cudaEventCreateWithFlags(&fs.ev_compose, cudaEventDisableTiming));
            PM3D_CUDA(ctx, cudaEventCreateWithFlags(&fs.ev_stage, cudaEventDisableTiming));
        }
    }
    ctx->cur_set = 0;
    for (int l = 0; l < PM3D_MAX_LANES; l++) ctx->lane_cur[l] = 0;
    ctx->lane_pos = 0;
    ctx->cur_cstream = ctx->last_cstream = ctx->stream;
    pm3d_load_set(ctx, 0);
    if ((rc = dev_realloc(ctx, &ctx->frame, V * W * H * 4))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->checksums, V))) return rc;
    ctx->last_frame = nullptr;
    ctx->last_views = 0;
    if (ctx->ring_n > 1) return pm3d_set_option(ctx, PM3D_OPT_PIPELINE, ctx->ring_n);  // re-size the slots
    return PM3D_OK;
}

// ------------------------------------------------------------------ frame path

// the three kernels of one frame batch; inputs come from ctx->cur_*
// (the caller selected the frame set: pm3d_begin_frame)
static int render_core(pm3d_ctx *ctx, int n_views, int n_sprites, uint8_t *target) {
    int rc;
    pm3d_ctx::FrameSet &fs = ctx->sets[ctx->cur_set];
    const bool overlap = ctx->n_sets > 1;
    cudaStream_t fstream = overlap ? ctx->cur_fstream : ctx->stream;
    cudaStream_t cstream = overlap ? ctx->cur_cstream : ctx->stream;
    ctx->cur_cstream = cstream;
    ctx->cur_n_sprites = n_sprites;
    if ((rc = pm3d_launch_front(ctx, n_views, fstream))) return rc;
    if (overlap) {
        PM3D_CUDA(ctx, cudaEventRecord(fs.ev_front, fstream));
        PM3D_CUDA(ctx, cudaStreamWaitEvent(cstream, fs.ev_front, 0));
    }
    if ((rc = pm3d_launch_compose(ctx, n_views, target))) return rc;
    if (overlap) PM3D_CUDA(ctx, cudaEventRecord(fs.ev_compose, cstream));
    if (ctx->capturing) ctx->set_in_capture[ctx->cur_set] = true;
    ctx->last_cstream = cstream;
    if (ctx->kt_on && ctx->kt_n < ctx->kt_cap) ctx->kt_n++;
    ctx->last_frame = target;
    ctx->last_views = n_views;
    ctx->last_sprites = n_sprites;
    return PM3D_OK;
}

// Select the frame set of the next frame: the named per-frame fields alias it from here on.  Its previous
// user (the compose kernel two frames back) must be done before the front stream writes into it, and
// everything enqueued on the main stream so far (sprite lists, clutter, uploads, timers: whatever the caller
// ordered before this frame) precedes the frame: order_after_main.
static int pm3d_begin_frame(pm3d_ctx *ctx, int n_views, cudaStream_t *fstream, int lane = 0) {
    if (ctx->n_sets > 1) {
        ctx->lane_cur[lane] ^= 1;
        ctx->cur_set = 2 * lane + ctx->lane_cur[lane];
        pm3d_load_set(ctx, ctx->cur_set);
        if (lane == 0) {
            // a batch whose front launch fills the chip by itself gains nothing from jumping the queue
            const long long front_ctas = (long long)n_views * ctx->num_levels * ctx->n_groups / 4;
            ctx->cur_fstream = front_ctas >= 4LL * ctx->sm_count ? ctx->front_stream_lo : ctx->front_stream;
            ctx->cur_cstream = ctx->stream;
        } else {
            ctx->cur_fstream = ctx->lane_fstream[lane];
            ctx->cur_cstream = ctx->lane_cstream[lane];
        }
        if (!ctx->capturing || ctx->set_in_capture[ctx->cur_set])
            PM3D_CUDA(ctx, cudaStreamWaitEvent(ctx->cur_fstream, ctx->sets[ctx->cur_set].ev_compose, 0));
        uint64_t *seen = lane ? &ctx->seen_lane[lane] : (ctx->cur_fstream == ctx->front_stream ? &ctx->seen_front : &ctx->seen_front_lo);
        int rc = order_after_main(ctx, ctx->cur_fstream, seen);
        if (rc) return rc;
        if (lane) ctx->lanes_dirty = true;
        *fstream = ctx->cur_fstream;
    } else {
        ctx->cur_cstream = ctx->stream;
        *fstream = ctx->stream;
    }
    return PM3D_OK;
}

static int render_ready(pm3d_ctx *ctx, int n_views) {
    if (!ctx->grid || !ctx->wall_lut || !ctx->atlas || !ctx->floor_tiles || ctx->w < 1)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "grid, wall textures, atlas, floor tiles and view size must be set first");
    if (ctx->floor_w != ctx->map_w || ctx->floor_h != ctx->map_h)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "floor tile grid %dx%d does not match map %dx%d", ctx->floor_w, ctx->floor_h, ctx->map_w, ctx->map_h);
    if (n_views < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "no views");
    if (n_views > ctx->max_views) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "%d views > max_views %d", n_views, ctx->max_views);
    if (ctx->sky_tile >= ctx->n_tiles || ctx->floorbox_tile >= ctx->n_tiles)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "sky/floor box tile outside the atlas");
    if (ctx->max_wall_tile >= ctx->n_tiles || ctx->max_floor_tile >= ctx->n_tiles)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "a wall (%d) or floor (%d) tile index lies outside the %d-tile atlas", ctx->max_wall_tile, ctx->max_floor_tile, ctx->n_tiles);
    return PM3D_OK;
}

int pm3d_render_resident(pm3d_ctx *ctx, const pm3d_view *d_views, int n_views, const pm3d_sprite *d_sprites,
                         const int32_t *d_sprite_offset, int n_sprites_total, uint8_t *d_out_rgba) {
    if (!ctx) return PM3D_ERR_ARG;
    int rc = render_ready(ctx, n_views);
    if (rc) return rc;
    if (!d_views || !d_sprite_offset || !d_out_rgba || n_sprites_total < 0 || (n_sprites_total > 0 && !d_sprites))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "null device pointer");
    if (n_sprites_total > ctx->max_sprites) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "%d sprites > max_sprites_total %d", n_sprites_total, ctx->max_sprites);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t fstream;
    int lane = 0;
    if (ctx->n_lanes > 1 && ctx->n_sets > 1) {  // independent frames: next lane
        lane = ctx->lane_pos;
        ctx->lane_pos = (lane + 1) % ctx->n_lanes;
    }
    if ((rc = pm3d_begin_frame(ctx, n_views, &fstream, lane))) return rc;
    ctx->cur_views = d_views; ctx->cur_sprites = d_sprites; ctx->cur_offset = d_sprite_offset;
    return render_core(ctx, n_views, n_sprites_total, d_out_rgba);
}

// ---- recorded frame sequences
static void capture_streams(pm3d_ctx *ctx, std::vector<cudaStream_t> &aux) {
    if (ctx->n_sets <= 1) return;  // PM3D_OPT_OVERLAP = 0: everything runs on the ctx stream
    aux.push_back(ctx->front_stream);
    aux.push_back(ctx->front_stream_lo);
    for (int l = 1; l < ctx->n_lanes; l++) { aux.push_back(ctx->lane_fstream[l]); aux.push_back(ctx->lane_cstream[l]); }
}

int pm3d_frames_capture_begin(pm3d_ctx *ctx) {
    if (!ctx) return PM3D_ERR_ARG;
    if (ctx->capturing) return pm3d_fail(ctx, PM3D_ERR_ARG, "a frame capture is already open");
    if (!ctx->own_stream) return pm3d_fail(ctx, PM3D_ERR_ARG, "a ctx on a caller stream cannot record frame sequences");
    if (ctx->w < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_set_view must be called first");
    if (ctx->kt_on) return pm3d_fail(ctx, PM3D_ERR_ARG, "per-kernel timing (pm3d_ktimer_begin) is on");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, sync_frame_streams(ctx));  // nothing of the capture may depend on work enqueued before it
    int rc = pm3d_check_status(ctx);
    if (rc) return rc;
    if (!ctx->ev_fork) {
        PM3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (cudaEvent_t &e : ctx->ev_join) PM3D_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    ctx->lanes_dirty = false;
    ctx->seen_front = ctx->seen_front_lo = ctx->main_seq;
    for (uint64_t &s : ctx->seen_lane) s = ctx->main_seq;
    for (bool &b : ctx->set_in_capture) b = false;
    PM3D_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    // fork: every stream a frame may run on joins the capture
    std::vector<cudaStream_t> aux;
    capture_streams(ctx, aux);
    cudaError_t e = cudaEventRecord(ctx->ev_fork, ctx->stream);
    for (cudaStream_t s : aux)
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, ctx->ev_fork, 0);
    if (e != cudaSuccess) {
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(ctx->stream, &g);
        if (g) cudaGraphDestroy(g);
        return pm3d_fail(ctx, PM3D_ERR_CUDA, "frame capture: %s", cudaGetErrorString(e));
    }
    ctx->capturing = true;
    ctx->capture_launches0 = ctx->launches;
    return PM3D_OK;
}

int pm3d_frames_capture_end(pm3d_ctx *ctx, int *out_graph) {
    if (!ctx || !out_graph) return PM3D_ERR_ARG;
    if (!ctx->capturing) return pm3d_fail(ctx, PM3D_ERR_ARG, "no frame capture is open");
    ctx->capturing = false;
    std::vector<cudaStream_t> aux;
    capture_streams(ctx, aux);
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < aux.size() && e == cudaSuccess; k++) {  // join: the graph ends when every stream's work has
        e = cudaEventRecord(ctx->ev_join[k], aux[k]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0);
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
    if (e == cudaSuccess) e = e2;
    pm3d_ctx::FrameGraph fg;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&fg.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    ctx->lanes_dirty = false;
    pm3d_main_touch(ctx);
    // an event last recorded inside a capture cannot be waited for outside of it: give the frame sets' events a real
    // (immediately complete) record again
    for (int k = 0; k < 2 * PM3D_MAX_LANES; k++)
        if (ctx->set_in_capture[k] && ctx->sets[k].ev_compose) {
            cudaEventRecord(ctx->sets[k].ev_compose, ctx->stream);
            cudaEventRecord(ctx->sets[k].ev_front, ctx->stream);
            ctx->set_in_capture[k] = false;
        }
    if (e != cudaSuccess) return pm3d_fail(ctx, PM3D_ERR_CUDA, "frame capture: %s", cudaGetErrorString(e));
    fg.launches = ctx->launches - ctx->capture_launches0;
    fg.last_frame = ctx->last_frame; fg.last_views = ctx->last_views; fg.last_sprites = ctx->last_sprites;
    ctx->launches = ctx->capture_launches0;  // nothing ran yet
    ctx->graphs.push_back(fg);
    *out_graph = (int)ctx->graphs.size() - 1;
    return PM3D_OK;
}

int pm3d_frames_replay(pm3d_ctx *ctx, int graph) {
    if (!ctx) return PM3D_ERR_ARG;
    if (graph < 0 || graph >= (int)ctx->graphs.size() || !ctx->graphs[graph].exec) return pm3d_fail(ctx, PM3D_ERR_ARG, "no such recorded frame sequence");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = pm3d_main_enter(ctx);  // after the frames dealt to the lanes so far
    if (rc) return rc;
    const pm3d_ctx::FrameGraph &fg = ctx->graphs[graph];
    PM3D_CUDA(ctx, cudaGraphLaunch(fg.exec, ctx->stream));
    ctx->launches += fg.launches;
    ctx->last_cstream = ctx->stream;
    ctx->last_frame = fg.last_frame; ctx->last_views = fg.last_views; ctx->last_sprites = fg.last_sprites;
    pm3d_main_touch(ctx);  // the next frame call waits for the ctx stream, i.e. for the replayed frames
    return PM3D_OK;
}

int pm3d_frames_graph_destroy(pm3d_ctx *ctx, int graph) {
    if (!ctx) return PM3D_ERR_ARG;
    if (graph < 0 || graph >= (int)ctx->graphs.size()) return pm3d_fail(ctx, PM3D_ERR_ARG, "no such recorded frame sequence");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->graphs[graph].exec) PM3D_CUDA(ctx, cudaGraphExecDestroy(ctx->graphs[graph].exec));
    ctx->graphs[graph].exec = nullptr;
    return PM3D_OK;
}

int pm3d_download_results(pm3d_ctx *ctx, pm3d_rect *out_rects, pm3d_convergence *out_conv) {
    if (!ctx) return PM3D_ERR_ARG;
    if (ctx->last_views < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "nothing rendered");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, wait_last_frame(ctx));
    if (out_rects && ctx->last_sprites)
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_rects, ctx->rects, (size_t)ctx->last_sprites * sizeof(pm3d_rect), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_conv)
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_conv, ctx->conv, (size_t)ctx->last_views * sizeof(pm3d_convergence), cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return pm3d_check_status(ctx);
}

int pm3d_ktimer_begin(pm3d_ctx *ctx, int max_frames) {
    if (!ctx || max_frames < 1 || max_frames > (1 << 20)) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->kt_cap < max_frames) {
        cudaEvent_t *ev = new cudaEvent_t[(size_t)max_frames * 6];
        for (int i = 0; i < ctx->kt_cap * 6; i++) ev[i] = ctx->kt_ev[i];
        for (int i = ctx->kt_cap * 6; i < max_frames * 6; i++) PM3D_CUDA(ctx, cudaEventCreate(&ev[i]));
        delete[] ctx->kt_ev;
        ctx->kt_ev = ev;
        ctx->kt_cap = max_frames;
    }
    ctx->kt_n = 0;
    ctx->kt_on = 1;
    return PM3D_OK;
}

int pm3d_ktimer_collect(pm3d_ctx *ctx, float *out_ms, int cap_frames) {
    if (!ctx || !out_ms) return PM3D_ERR_ARG;
    ctx->kt_on = 0;
    PM3D_CUDA(ctx, sync_frame_streams(ctx));  // every lane: the event pairs sit on the streams their frames ran on
    const int n = ctx->kt_n < cap_frames ? ctx->kt_n : cap_frames;
    for (int f = 0; f < n; f++)
        for (int k = 0; k < 3; k++) {
            out_ms[f * 3 + k] = 0.0f;  // slot 1 is unused since the cast and the view setup share one launch
            if (k != 1) PM3D_CUDA(ctx, cudaEventElapsedTime(&out_ms[f * 3 + k], ctx->kt_ev[(f * 3 + k) * 2], ctx->kt_ev[(f * 3 + k) * 2 + 1]));
        }
    ctx->kt_n = 0;
    return n;
}

int pm3d_render(pm3d_ctx *ctx, const pm3d_view *views, int n_views, const pm3d_sprite *sprites,
                const int32_t *sprite_offset, uint8_t *out_rgba, pm3d_rect *out_rects, pm3d_convergence *out_conv,
                uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!views) return pm3d_fail(ctx, PM3D_ERR_ARG, "no views");
    {
        int rdy = render_ready(ctx, n_views);
        if (rdy) return rdy;
    }
    std::vector<int32_t> zero_off;
    if (!sprite_offset) {
        zero_off.assign(n_views + 1, 0);
        sprite_offset = zero_off.data();
    }
    const int n_sprites = sprite_offset[n_views];
    if (sprite_offset[0] != 0 || n_sprites < 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "sprite_offset must start at 0");
    if (n_sprites > ctx->max_sprites) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "%d sprites > max_sprites_total %d", n_sprites, ctx->max_sprites);
    if (n_sprites > 0 && !sprites) return pm3d_fail(ctx, PM3D_ERR_ARG, "sprites is null");
    for (int v = 0; v < n_views; v++) {
        const int nv = sprite_offset[v + 1] - sprite_offset[v];
        if (nv < 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "sprite_offset not monotone");
        if (nv > 2048) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "view %d has %d sprites (limit 2048 per view)", v, nv);
    }
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));

    // stage the small per-frame inputs through pinned memory, one H2D each
    const size_t vb = (size_t)n_views * sizeof(pm3d_view), ob = (size_t)(n_views + 1) * 4,
                 sb = (size_t)n_sprites * sizeof(pm3d_sprite);
    const size_t ob_pad = (ob + 15) & ~(size_t)15;
    cudaStream_t fstream;
    int rc = pm3d_begin_frame(ctx, n_views, &fstream);
    if (rc) return rc;
    pm3d_ctx::FrameSet &fs = ctx->sets[ctx->cur_set];
    const size_t in_bytes = vb + ob_pad + sb;
    // this set's previous H2D (two frames back for a PM3D_NO_SYNC caller) must have left the staging block
    PM3D_CUDA(ctx, cudaEventSynchronize(fs.ev_stage));
    if (fs.stage_bytes < in_bytes) {
        if (fs.stage) { PM3D_CUDA(ctx, cudaFreeHost(fs.stage)); fs.stage = nullptr; fs.stage_bytes = 0; }
        const size_t cap = std::max<size_t>(in_bytes, 1 << 20);
        PM3D_CUDA(ctx, cudaMallocHost(&fs.stage, cap));
        fs.stage_bytes = cap;
    }
    uint8_t *st = (uint8_t *)fs.stage;
    memcpy(st, views, vb);
    memcpy(st + vb, sprite_offset, ob);
    if (sb) memcpy(st + vb + ob_pad, sprites, sb);
    // one H2D for the whole per-frame input block: [views | sprite_offset | sprites]
    if (fs.inbuf_bytes < in_bytes) {
        PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // frames that still read the old block
        if (fs.inbuf) PM3D_CUDA(ctx, cudaFree(fs.inbuf));
        fs.inbuf = nullptr;
        fs.inbuf_bytes = std::max<size_t>(in_bytes, (size_t)1 << 16) * 2;
        PM3D_CUDA(ctx, cudaMalloc((void **)&fs.inbuf, fs.inbuf_bytes));
    }
    PM3D_CUDA(ctx, cudaMemcpyAsync(fs.inbuf, st, in_bytes, cudaMemcpyHostToDevice, fstream));
    PM3D_CUDA(ctx, cudaEventRecord(fs.ev_stage, fstream));

    const bool piped = ctx->ring_n > 1 && out_rgba && !(flags & PM3D_OUT_DEVICE) && (flags & PM3D_NO_SYNC);
    const int slot = ctx->ring_pos;
    uint8_t *target = (out_rgba && (flags & PM3D_OUT_DEVICE)) ? out_rgba : (piped ? ctx->frame_ring[slot] : ctx->frame);
    // the slot's previous frame must have left the device before the compose kernel overwrites it
    if (piped) PM3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0));
    ctx->cur_views = (const pm3d_view *)fs.inbuf;
    ctx->cur_offset = (const int32_t *)(fs.inbuf + vb);
    ctx->cur_sprites = (const pm3d_sprite *)(fs.inbuf + vb + ob_pad);
    if ((rc = render_core(ctx, n_views, n_sprites, target))) return rc;

    const size_t fb = (size_t)n_views * ctx->w * ctx->h * 4;
    if (piped) {
        // pipelined: the frame copy goes to the copy stream, so the main stream never waits for the copy engine and the
        // next frame's kernels start at once.  The small results do not go through the copy engine at all (a few-KB
        // copy costs microseconds of DMA set-up between the 8 MB frame copies that bound this path): a kernel stores
        // them into mapped pinned memory and pm3d_sync hands them to the caller's buffers.
        PM3D_CUDA(ctx, cudaEventRecord(ctx->ev_rendered[slot], ctx->stream));
        PM3D_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[slot], 0));
        const int nr = out_rects ? n_sprites : 0, nc = out_conv ? n_views : 0;
        if (nr || nc) {
            const size_t rect_cap = (size_t)std::max(ctx->max_sprites, 1), need = rect_cap * sizeof(pm3d_rect) + (size_t)ctx->max_views * sizeof(pm3d_convergence);
            if (ctx->small_slot_bytes < need || ctx->small_rect_cap != rect_cap) {
                PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                small_flush_all(ctx);
                if (ctx->small_ring) PM3D_CUDA(ctx, cudaFreeHost(ctx->small_ring));
                ctx->small_ring = nullptr;
                PM3D_CUDA(ctx, cudaHostAlloc((void **)&ctx->small_ring, need * PM3D_SMALL_SLOTS, cudaHostAllocMapped));
                ctx->small_slot_bytes = need;
                ctx->small_rect_cap = rect_cap;
            }
            const int s = ctx->small_pos;
            ctx->small_pos = (s + 1) % PM3D_SMALL_SLOTS;
            pm3d_ctx::SmallPending &sp = ctx->small[s];
            if (!sp.done) PM3D_CUDA(ctx, cudaEventCreateWithFlags(&sp.done, cudaEventDisableTiming));
            if (sp.active) {  // PM3D_SMALL_SLOTS frames ago: long finished
                PM3D_CUDA(ctx, cudaEventSynchronize(sp.done));
                small_flush(ctx, s);
            }
            uint4 *d_slot = nullptr;
            PM3D_CUDA(ctx, cudaHostGetDevicePointer((void **)&d_slot, ctx->small_ring + (size_t)s * need, 0));
            const int n16 = std::max(nr, nc * 2);
            k_export_small<<<(n16 + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(ctx->rects), nr,
                                                                       reinterpret_cast<const uint4 *>(ctx->conv), nc * 2, d_slot, rect_cap);
            ctx->launches++;
            PM3D_CUDA(ctx, cudaGetLastError());
            PM3D_CUDA(ctx, cudaEventRecord(sp.done, ctx->stream));
            sp.dst_rects = out_rects; sp.dst_conv = out_conv; sp.n_rects = nr; sp.n_conv = nc; sp.active = true;
        }
        // the set's rects/conv have now been read: only after this point may its next frame rewrite them
        if (ctx->n_sets > 1) PM3D_CUDA(ctx, cudaEventRecord(fs.ev_compose, ctx->stream));
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_rgba, target, fb, cudaMemcpyDeviceToHost, ctx->copy_stream));
        PM3D_CUDA(ctx, cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
        ctx->ring_pos = (slot + 1) % ctx->ring_n;
    } else {
        if (out_rects && n_sprites)
            PM3D_CUDA(ctx, cudaMemcpyAsync(out_rects, ctx->rects, (size_t)n_sprites * sizeof(pm3d_rect), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_conv)
            PM3D_CUDA(ctx, cudaMemcpyAsync(out_conv, ctx->conv, (size_t)n_views * sizeof(pm3d_convergence), cudaMemcpyDeviceToHost, ctx->stream));
        // the set's rects/conv have now been read: only after this point may its next frame rewrite them
        if (ctx->n_sets > 1) PM3D_CUDA(ctx, cudaEventRecord(fs.ev_compose, ctx->stream));
        if (out_rgba && !(flags & PM3D_OUT_DEVICE))
            PM3D_CUDA(ctx, cudaMemcpyAsync(out_rgba, ctx->frame, fb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}

int pm3d_download_columns(pm3d_ctx *ctx, int view, pm3d_column *columns, double *zbuf) {
    if (!ctx) return PM3D_ERR_ARG;
    if (view < 0 || view >= ctx->last_views) return pm3d_fail(ctx, PM3D_ERR_ARG, "view %d not rendered", view);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, wait_last_frame(ctx));
    if (columns) {
        if (!ctx->keep_columns || !ctx->dbg_cols) return pm3d_fail(ctx, PM3D_ERR_ARG, "PM3D_OPT_KEEP_COLUMNS is off");
        const size_t cnt = (size_t)ctx->num_levels * ctx->w;
        PM3D_CUDA(ctx, cudaMemcpyAsync(columns, ctx->dbg_cols + (size_t)view * cnt, cnt * sizeof(pm3d_column), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (zbuf)
        PM3D_CUDA(ctx, cudaMemcpyAsync(zbuf, ctx->zbuf + (size_t)view * ctx->w, (size_t)ctx->w * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}

int pm3d_frame_checksums(pm3d_ctx *ctx, int n_views, uint64_t *out, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->last_frame || n_views < 1 || n_views > ctx->last_views || !out) return pm3d_fail(ctx, PM3D_ERR_ARG, "nothing rendered / bad arguments");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, wait_last_frame(ctx));
    int rc = pm3d_main_enter(ctx);
    if (rc) return rc;
    if ((rc = pm3d_launch_checksum(ctx, n_views, ctx->last_frame))) return rc;
    pm3d_main_touch(ctx);
    if (flags & PM3D_OUT_DEVICE)
        PM3D_CUDA(ctx, cudaMemcpyAsync(out, ctx->checksums, (size_t)n_views * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    else
        PM3D_CUDA(ctx, cudaMemcpyAsync(out, ctx->checksums, (size_t)n_views * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}

void *pm3d_last_frame_device(pm3d_ctx *ctx) { return ctx ? ctx->last_frame : nullptr; }

// ------------------------------------------------------------------ LOS segments + bins

static inline int64_t host_go_int(double v) {
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;
}

int pm3d_upload_segments(pm3d_ctx *ctx, const double *xyxy, int n) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || (n > 0 && !xyxy)) return pm3d_fail(ctx, PM3D_ERR_ARG, "null segments");
    if (n >= (1 << 27)) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "%d collision segments (the LOS kernel packs a segment id into 27 bits)", n);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    // bounding box of all endpoints -> bin area dilated by one cell plus one spare
    double minx = 0, miny = 0, maxx = 1, maxy = 1;
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 4; k++)
            if (!isfinite(xyxy[i * 4 + k])) return pm3d_fail(ctx, PM3D_ERR_ARG, "segment %d is not finite", i);
        const double *s = xyxy + (size_t)i * 4;
        if (i == 0) { minx = std::min(s[0], s[2]); maxx = std::max(s[0], s[2]); miny = std::min(s[1], s[3]); maxy = std::max(s[1], s[3]); }
        minx = std::min(minx, std::min(s[0], s[2])); maxx = std::max(maxx, std::max(s[0], s[2]));
        miny = std::min(miny, std::min(s[1], s[3])); maxy = std::max(maxy, std::max(s[1], s[3]));
    }
    if (maxx - minx > 30000 || maxy - miny > 30000) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "segment extent too large to bin");
    const int x0 = (int)floor(minx) - 2, y0 = (int)floor(miny) - 2;
    const int bw = (int)floor(maxx) + 3 - x0, bh = (int)floor(maxy) + 3 - y0;
    std::vector<int32_t> count((size_t)bw * bh + 1, 0);
    std::vector<int4> irect(std::max(n, 1));
    // A segment is registered in every 1x1 cell whose box, grown by BIN_EPS on each side, it touches.
    // BIN_EPS is ~8 orders of magnitude above the rounding error of the device's cell walk and of the
    // predicate at map-scale coordinates, so any segment the exact predicate could report against a ray
    // is listed in a cell that ray's walk visits (k_los); it is far below a cell, so lists stay short.
    const double BIN_EPS = 1e-3;
    auto cell_range = [&](const double *s, int &cx0, int &cy0, int &cx1, int &cy1) {
        cx0 = (int)floor(std::min(s[0], s[2]) - BIN_EPS) - x0; cx1 = (int)floor(std::max(s[0], s[2]) + BIN_EPS) - x0;
        cy0 = (int)floor(std::min(s[1], s[3]) - BIN_EPS) - y0; cy1 = (int)floor(std::max(s[1], s[3]) + BIN_EPS) - y0;
    };
    // conservative segment-vs-grown-box overlap (Liang-Barsky on the infinite precision-free side: any doubt keeps it)
    auto touches = [&](const double *s, int cx, int cy) -> bool {
        const double bx0 = (double)(cx + x0) - 2 * BIN_EPS, bx1 = (double)(cx + x0) + 1.0 + 2 * BIN_EPS;
        const double by0 = (double)(cy + y0) - 2 * BIN_EPS, by1 = (double)(cy + y0) + 1.0 + 2 * BIN_EPS;
        double t0 = 0.0, t1 = 1.0;
        const double dx = s[2] - s[0], dy = s[3] - s[1];
        const double pq[4][2] = {{-dx, s[0] - bx0}, {dx, bx1 - s[0]}, {-dy, s[1] - by0}, {dy, by1 - s[1]}};
        for (int k = 0; k < 4; k++) {
            const double pp = pq[k][0], q = pq[k][1];
            if (pp == 0) { if (q < 0) return false; continue; }
            const double r = q / pp;
            if (pp < 0) { if (r > t1 + 1e-9) return false; if (r > t0) t0 = r; }
            else { if (r < t0 - 1e-9) return false; if (r < t1) t1 = r; }
        }
        return true;
    };
    for (int i = 0; i < n; i++) {
        const double *s = xyxy + (size_t)i * 4;
        int cx0, cy0, cx1, cy1;
        cell_range(s, cx0, cy0, cx1, cy1);
        for (int cy = cy0; cy <= cy1; cy++)
            for (int cx = cx0; cx <= cx1; cx++)
                if (touches(s, cx, cy)) count[(size_t)cy * bw + cx + 1]++;
        int64_t ax = host_go_int(s[0]), ay = host_go_int(s[1]), bx = host_go_int(s[2]), by = host_go_int(s[3]);
        if (ax > bx) std::swap(ax, bx);
        if (ay > by) std::swap(ay, by);
        irect[i] = make_int4((int)ax, (int)ay, (int)bx, (int)by);
    }
    for (size_t c = 0; c < (size_t)bw * bh; c++) count[c + 1] += count[c];
    std::vector<int32_t> items(std::max<size_t>(count.back(), 1));
    std::vector<int32_t> fill(count.begin(), count.end() - 1);
    for (int i = 0; i < n; i++) {
        const double *s = xyxy + (size_t)i * 4;
        int cx0, cy0, cx1, cy1;
        cell_range(s, cx0, cy0, cx1, cy1);
        for (int cy = cy0; cy <= cy1; cy++)
            for (int cx = cx0; cx <= cx1; cx++)
                if (touches(s, cx, cy)) items[fill[(size_t)cy * bw + cx]++] = i;
    }
    int rc;
    if ((rc = dev_realloc(ctx, &ctx->segs, (size_t)std::max(n, 1) * 4))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->seg_irect, (size_t)std::max(n, 1)))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->cell_start, count.size()))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->cell_items, items.size()))) return rc;
    std::vector<int4> recs((size_t)bw * bh);
    for (size_t c = 0; c < (size_t)bw * bh; c++) {
        const int b = count[c], e = count[c + 1], k = e - b;
        recs[c] = k <= 3 ? make_int4(k, k > 0 ? items[b] : -1, k > 1 ? items[b + 1] : -1, k > 2 ? items[b + 2] : -1)
                         : make_int4(k, b, 0, 0);
    }
    if ((rc = dev_realloc(ctx, &ctx->cell_rec, recs.size()))) return rc;
    PM3D_CUDA(ctx, cudaMemcpy(ctx->cell_rec, recs.data(), recs.size() * sizeof(int4), cudaMemcpyHostToDevice));
    // code byte per bin cell for the first LOS pass (los_kernel.cu k_los_fast): bit 0 = the cell has candidates;
    // bits 1..4 = its W / E / S / N unit edge is covered by an axis-parallel segment lying exactly on that lattice line
    {
        std::vector<uint8_t> code((((size_t)bw * bh + 15) / 16) * 16, 0);
        for (size_t c = 0; c < (size_t)bw * bh; c++)
            if (count[c + 1] > count[c]) code[c] |= 1u;
        for (int i = 0; i < n; i++) {
            const double *s = xyxy + (size_t)i * 4;
            const bool vertical = s[0] == s[2], horizontal = s[1] == s[3];
            const double line = vertical ? s[0] : s[1];
            const double lo = vertical ? std::min(s[1], s[3]) : std::min(s[0], s[2]);
            const double hi = vertical ? std::max(s[1], s[3]) : std::max(s[0], s[2]);
            // CLEAN = axis-parallel, on a lattice line, integer end points: a union of whole unit edges.  Every other
            // segment (oblique, a point, off the lattice like the border rectangles inset by clipDistance, fractional
            // ends) makes the cells it is registered in DIRTY: the first pass never decides "visible" through them
            const bool clean = vertical != horizontal && line == floor(line) && lo == floor(lo) && hi == floor(hi);
            if (!clean) {
                int cx0, cy0, cx1, cy1;
                cell_range(s, cx0, cy0, cx1, cy1);
                for (int cy = cy0; cy <= cy1; cy++)
                    for (int cx = cx0; cx <= cx1; cx++)
                        if (touches(s, cx, cy)) code[(size_t)cy * bw + cx] |= 32u;
            }
            if (vertical == horizontal) continue;  // oblique, or a point
            if (line != floor(line)) continue;     // not on the lattice
            const int c = (int)line;
            for (int k = (int)ceil(lo); (double)(k + 1) <= hi; k++) {  // unit edges [k, k+1] inside [lo, hi]
                if (vertical) {
                    const int cy = k - y0, cxe = c - 1 - x0, cxw = c - x0;
                    if (cy < 0 || cy >= bh) continue;
                    if (cxe >= 0 && cxe < bw) code[(size_t)cy * bw + cxe] |= 4u;   // E edge of the cell on the left
                    if (cxw >= 0 && cxw < bw) code[(size_t)cy * bw + cxw] |= 2u;   // W edge of the cell on the right
                } else {
                    const int cx = k - x0, cyn = c - 1 - y0, cys = c - y0;
                    if (cx < 0 || cx >= bw) continue;
                    if (cyn >= 0 && cyn < bh) code[(size_t)cyn * bw + cx] |= 16u;  // N edge of the cell below
                    if (cys >= 0 && cys < bh) code[(size_t)cys * bw + cx] |= 8u;   // S edge of the cell above
                }
            }
        }
        // verify mode of the first pass pays when most cells have no candidates (long visible rays); on dense maps the
        // rays end at a covered edge after a few cells and the extra crossing tests would only cost
        size_t n_cand = 0;
        for (size_t c = 0; c < (size_t)bw * bh; c++) n_cand += code[c] & 1u;
        ctx->los_verify = n_cand * 4 < (size_t)bw * bh && bw <= 4096 && bh <= 4096;  // (the rounding bound of the walk's tX / tY)
        if ((rc = dev_realloc(ctx, &ctx->cell_code, code.size()))) return rc;
        PM3D_CUDA(ctx, cudaMemcpy(ctx->cell_code, code.data(), code.size(), cudaMemcpyHostToDevice));
    }
    if (n) {
        PM3D_CUDA(ctx, cudaMemcpy(ctx->segs, xyxy, (size_t)n * 32, cudaMemcpyHostToDevice));
        PM3D_CUDA(ctx, cudaMemcpy(ctx->seg_irect, irect.data(), (size_t)n * sizeof(int4), cudaMemcpyHostToDevice));
    }
    PM3D_CUDA(ctx, cudaMemcpy(ctx->cell_start, count.data(), count.size() * 4, cudaMemcpyHostToDevice));
    PM3D_CUDA(ctx, cudaMemcpy(ctx->cell_items, items.data(), items.size() * 4, cudaMemcpyHostToDevice));
    ctx->n_segs = n; ctx->bin_x0 = x0; ctx->bin_y0 = y0; ctx->bin_w = bw; ctx->bin_h = bh;
    return PM3D_OK;
}

// ------------------------------------------------------------------ memory + timing helpers

int pm3d_device_alloc(pm3d_ctx *ctx, int64_t bytes, void **out) {
    if (!ctx || !out || bytes < 0) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1)));
    return PM3D_OK;
}
int pm3d_device_free(pm3d_ctx *ctx, void *ptr) {
    if (!ctx) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaFree(ptr));
    return PM3D_OK;
}
int pm3d_host_alloc(pm3d_ctx *ctx, int64_t bytes, void **out) {
    if (!ctx || !out || bytes < 0) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1)));
    return PM3D_OK;
}
int pm3d_host_free(pm3d_ctx *ctx, void *ptr) {
    if (!ctx) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaFreeHost(ptr));
    return PM3D_OK;
}
int pm3d_memcpy_h2d(pm3d_ctx *ctx, void *dst, const void *src, int64_t bytes) {
    if (!ctx) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc = pm3d_main_enter(ctx); if (rc) return rc; }
    PM3D_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pm3d_main_touch(ctx);
    return PM3D_OK;
}
int pm3d_memcpy_d2h(pm3d_ctx *ctx, void *dst, const void *src, int64_t bytes) {
    if (!ctx) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc = pm3d_main_enter(ctx); if (rc) return rc; }
    PM3D_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pm3d_main_touch(ctx);
    return PM3D_OK;
}

int pm3d_timer_start(pm3d_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->ev_start[slot]) {
        PM3D_CUDA(ctx, cudaEventCreate(&ctx->ev_start[slot]));
        PM3D_CUDA(ctx, cudaEventCreate(&ctx->ev_stop[slot]));
    }
    int rc = pm3d_main_enter(ctx);  // the start event follows the frames already enqueued on every lane
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaEventRecord(ctx->ev_start[slot], ctx->stream));
    // fork: nothing of the timed region may start on a front stream or another lane before the start event —
    // the next frame of every lane (lane 0's front streams included) waits for the main stream
    pm3d_main_touch(ctx);
    return PM3D_OK;
}
int pm3d_timer_stop(pm3d_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16 || !ctx->ev_stop[slot]) return PM3D_ERR_ARG;
    // join: the stop event follows the last frame of every lane
    int rc = pm3d_main_enter(ctx);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaEventRecord(ctx->ev_stop[slot], ctx->stream));
    pm3d_main_touch(ctx);
    return PM3D_OK;
}
int pm3d_timer_elapsed_ms(pm3d_ctx *ctx, int slot, float *ms) {
    if (!ctx || slot < 0 || slot >= 16 || !ctx->ev_stop[slot] || !ms) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop[slot]));
    PM3D_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start[slot], ctx->ev_stop[slot]));
    return PM3D_OK;
}

int64_t pm3d_launch_count(const pm3d_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pm3d_flush_l2(pm3d_ctx *ctx) {
    if (!ctx) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
        PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->flush_buf, ctx->flush_bytes));
    }
    { int rc = pm3d_main_enter(ctx); if (rc) return rc; }
    PM3D_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    pm3d_main_touch(ctx);
    return PM3D_OK;
}

}  // extern "C"
