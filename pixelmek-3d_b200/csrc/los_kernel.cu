// los_kernel.cu — batched line-of-sight and wall-collision queries.
//
// Replaces Game.lineOfSight (game/collision.go:40-65) and the wall part of Game.getValidMove
// (game/collision.go:69-97, 202-203, 307-342) for batches of independent rays.  The reference
// tests every query segment against every segment of g.collisionMap; here the collision
// segments are binned once (at pm3d_upload_segments) into a uniform 1x1-cell grid with a
// one-cell dilation, and each query walks only the cells its segment crosses, running the
// reference's EXACT predicate (integer bounding-rect pre-filter + geom.LineIntersection in
// binary64) on the candidates.  The boolean per ray is therefore the brute-force answer:
// an intersection point lies on the query segment, hence inside a crossed cell (or, at a
// lattice corner, a neighbour of one), and the dilation covers that neighbour.
#include "pm3d_internal.cuh"

struct LosParams {
    const double4 *rays;  // x1,y1,x2,y2
    const double4 *segs;
    const int4 *seg_irect;  // image.Rect(int(x1),int(y1),int(x2),int(y2)) canonicalised: minx,miny,maxx,maxy
    const int32_t *cell_start;
    const int32_t *cell_items;
    int64_t n;
    int n_segs, bin_x0, bin_y0, bin_w, bin_h;
    // move queries
    const uint16_t *grid0;
    int map_w, map_h;
    double clip;
    uint8_t *out_flag;
    double2 *out_pos;
};

// geom.LineIntersection (raycaster-go geom; endpoints inclusive), boolean + point
__device__ __forceinline__ bool line_intersection(const double4 &a, const double4 &b, double *px, double *py) {
    const double x1 = a.x, y1 = a.y, x2 = a.z, y2 = a.w;
    const double x3 = b.x, y3 = b.y, x4 = b.z, y4 = b.w;
    const double denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    const double tNum = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4);
    const double uNum = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3));
    if (denom == 0) return false;
    const double t = tNum / denom;
    if (t > 1 || t < 0) return false;
    const double u = uNum / denom;
    if (u > 1 || u < 0) return false;
    if (px) { *px = x1 + t * (x2 - x1); *py = y1 + t * (y2 - y1); }
    return true;
}

__device__ __forceinline__ int4 go_rect(int64_t x0, int64_t y0, int64_t x1, int64_t y1) {
    if (x0 > x1) { int64_t t = x0; x0 = x1; x1 = t; }
    if (y0 > y1) { int64_t t = y0; y0 = y1; y1 = t; }
    const int64_t lo = -(1ll << 30), hi = 1ll << 30;  // map coordinates are far inside this range
    return make_int4((int)clamp_i64(x0, lo, hi), (int)clamp_i64(y0, lo, hi), (int)clamp_i64(x1, lo, hi),
                     (int)clamp_i64(y1, lo, hi));
}

// Visits the candidate segments of every bin cell crossed by `ray`; returns true as soon as
// `USE_PREFILTER ? lineOfSight's test : plain intersection` finds a hit.
template <bool USE_PREFILTER>
__device__ bool any_hit(const LosParams &p, const double4 &ray) {
    const int4 lr = go_rect(go_int(ray.x), go_int(ray.y), go_int(ray.z), go_int(ray.w));

    auto test = [&](int id) -> bool {
        if (USE_PREFILTER) {
            const int4 br = __ldg(p.seg_irect + id);
            if ((br.x > lr.z && br.y > lr.w) || (br.z < lr.x && br.w < lr.y)) return false;
        }
        const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
        const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
        const double4 sg = make_double4(s01.x, s01.y, s23.x, s23.y);
        return line_intersection(ray, sg, nullptr, nullptr);
    };

    // rays that are not finite or leave the binned area by more than a cell: brute force
    const double gx0 = (double)p.bin_x0, gy0 = (double)p.bin_y0;
    const double gx1 = (double)(p.bin_x0 + p.bin_w), gy1 = (double)(p.bin_y0 + p.bin_h);
    const bool inside = ray.x >= gx0 && ray.x < gx1 && ray.z >= gx0 && ray.z < gx1 && ray.y >= gy0 && ray.y < gy1 &&
                        ray.w >= gy0 && ray.w < gy1;
    if (!inside) {
        for (int id = 0; id < p.n_segs; id++)
            if (test(id)) return true;
        return false;
    }

    // cell walk (Amanatides-Woo) from the source cell to the target cell
    int cx = (int)floor(ray.x) - p.bin_x0, cy = (int)floor(ray.y) - p.bin_y0;
    const int ex = (int)floor(ray.z) - p.bin_x0, ey = (int)floor(ray.w) - p.bin_y0;
    const double dx = ray.z - ray.x, dy = ray.w - ray.y;
    const int sx = dx > 0 ? 1 : -1, sy = dy > 0 ? 1 : -1;
    const double inv_dx = dx != 0 ? 1.0 / fabs(dx) : 0.0, inv_dy = dy != 0 ? 1.0 / fabs(dy) : 0.0;
    // parametric distance to the next vertical / horizontal cell border
    double tx = dx != 0 ? (dx > 0 ? (floor(ray.x) + 1.0 - ray.x) : (ray.x - floor(ray.x))) * inv_dx : 1e300;
    double ty = dy != 0 ? (dy > 0 ? (floor(ray.y) + 1.0 - ray.y) : (ray.y - floor(ray.y))) * inv_dy : 1e300;
    int steps = abs(ex - cx) + abs(ey - cy);
    int last = -1;
    for (;;) {
        const int cell = cy * p.bin_w + cx;
        const int b = __ldg(p.cell_start + cell), e = __ldg(p.cell_start + cell + 1);
        for (int k = b; k < e; k++) {
            const int id = __ldg(p.cell_items + k);
            if (id == last) continue;
            last = id;
            if (test(id)) return true;
        }
        if (steps-- <= 0) break;
        // never step past the target cell on either axis: keeps the walk inside the bins
        if ((tx < ty && cx != ex) || cy == ey) { tx += inv_dx; cx += sx; }
        else { ty += inv_dy; cy += sy; }
    }
    return false;
}

__global__ void __launch_bounds__(256) k_los(LosParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const double4 ray = p.rays[i];
    p.out_flag[i] = any_hit<true>(p, ray) ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_move(LosParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const double4 m = p.rays[i];
    double posX = m.x, posY = m.y, newX = m.z, newY = m.w;
    uint8_t collided = 0;
    if (posX == newX && posY == newY) {
        p.out_pos[i] = make_double2(posX, posY);
        p.out_flag[i] = 0;
        return;
    }
    if (any_hit<false>(p, m)) {
        p.out_pos[i] = make_double2(posX, posY);
        p.out_flag[i] = 1;
        return;
    }
    int64_t ix = go_int(newX), iy = go_int(newY);
    if (ix < 0 || newX < 0) { newX = p.clip; ix = 0; }
    else if (ix >= p.map_w) { newX = (double)p.map_w - p.clip; ix = go_int(newX); }
    if (iy < 0 || newY < 0) { newY = p.clip; iy = 0; }
    else if (iy >= p.map_h) { newY = (double)p.map_h - p.clip; iy = go_int(newY); }
    if (ix > p.map_w - 1) ix = p.map_w - 1;
    if (iy > p.map_h - 1) iy = p.map_h - 1;
    if (__ldg(p.grid0 + ix * p.map_h + iy) == 0) { posX = newX; posY = newY; }
    else collided = 1;
    p.out_pos[i] = make_double2(posX, posY);
    p.out_flag[i] = collided;
}

static int ensure_io(pm3d_ctx *ctx, void **buf, size_t *have, size_t need) {
    if (*have >= need) return PM3D_OK;
    if (*buf) PM3D_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr; *have = 0;
    PM3D_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return PM3D_OK;
}

static int run_batch(pm3d_ctx *ctx, bool is_move, const double *rays, int64_t n, double clip, double *out_pos,
                     uint8_t *out_flag, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || (n > 0 && (!rays || !out_flag))) return pm3d_fail(ctx, PM3D_ERR_ARG, "null batch pointers");
    if (is_move && n > 0 && !out_pos) return pm3d_fail(ctx, PM3D_ERR_ARG, "null out_pos");
    if (!ctx->segs && ctx->n_segs != 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "segments not uploaded");
    if (!ctx->cell_start) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_segments must be called first");
    if (is_move && !ctx->grid) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_grid must be called first");
    if (n == 0) return PM3D_OK;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool in_dev = flags & PM3D_IN_DEVICE, out_dev = flags & PM3D_OUT_DEVICE;
    const double *d_rays = rays;
    if (!in_dev) {
        int rc = ensure_io(ctx, &ctx->io_a, &ctx->io_a_bytes, (size_t)n * 32);
        if (rc) return rc;
        PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->io_a, rays, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
        d_rays = (const double *)ctx->io_a;
    }
    uint8_t *d_flag = out_flag;
    double *d_pos = out_pos;
    if (!out_dev) {
        int rc = ensure_io(ctx, &ctx->io_b, &ctx->io_b_bytes, (size_t)n);
        if (rc) return rc;
        d_flag = (uint8_t *)ctx->io_b;
        if (is_move) {
            rc = ensure_io(ctx, &ctx->io_c, &ctx->io_c_bytes, (size_t)n * 16);
            if (rc) return rc;
            d_pos = (double *)ctx->io_c;
        }
    }
    LosParams p;
    p.rays = (const double4 *)d_rays;
    p.segs = (const double4 *)ctx->segs;
    p.seg_irect = ctx->seg_irect;
    p.cell_start = ctx->cell_start;
    p.cell_items = ctx->cell_items;
    p.n = n;
    p.n_segs = ctx->n_segs; p.bin_x0 = ctx->bin_x0; p.bin_y0 = ctx->bin_y0; p.bin_w = ctx->bin_w; p.bin_h = ctx->bin_h;
    p.grid0 = ctx->grid; p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.clip = clip;
    p.out_flag = d_flag;
    p.out_pos = (double2 *)d_pos;
    const int64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffll) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "batch too large");
    if (is_move) k_move<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
    else k_los<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    if (!out_dev) {
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_flag, d_flag, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (is_move) PM3D_CUDA(ctx, cudaMemcpyAsync(out_pos, d_pos, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}

extern "C" int pm3d_los_batch(pm3d_ctx *ctx, const double *rays, int64_t n, uint8_t *out, uint32_t flags) {
    return run_batch(ctx, false, rays, n, 0.0, nullptr, out, flags);
}

extern "C" int pm3d_move_batch(pm3d_ctx *ctx, const double *moves, int64_t n, double clip_distance, double *out_pos,
                               uint8_t *out_collided, uint32_t flags) {
    return run_batch(ctx, true, moves, n, clip_distance, out_pos, out_collided, flags);
}
