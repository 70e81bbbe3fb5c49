// sprite_list_kernel.cu — Game.getRaycastSprites (game/sprites.go:102-145) for a batch of cameras.
//
// The reference walks every map sprite per frame on one goroutine ("for much larger amounts of sprites may need
// goroutines to divide up the work", sprites.go:111) and appends clutter and a few extras.  Here one CTA per camera
// filters the shared sprite list with the reference's test and writes that camera's list — map sprites in host
// order, clutter, extras — straight into the concatenated `sprites` / `sprite_offset` layout the frame path reads:
// count per camera, one exclusive scan over the cameras, ordered compaction.
#include "pm3d_internal.cuh"

namespace {

constexpr int LIST_THREADS = 128;

struct ListParams {
    const pm3d_sprite *world;
    const pm3d_list_query *queries;
    const pm3d_sprite *clutter;
    const int32_t *clutter_count;
    const pm3d_sprite *extras;
    const int32_t *extra_offset;
    pm3d_sprite *out_sprites;
    int32_t *out_offset;  // [n + 1]; holds the per-camera counts between the first and the second launch
    int32_t *out_index;
    double render_distance, reach;  // reach = math.Ceil(renderDistance)
    int n_world, n, clutter_stride, max_total;
};

// renderDistance < 0 || g.player.Target() == sprite.Entity || model.PointInProximity(renderDistance, cam, sprite)
__device__ __forceinline__ bool do_sprite(const ListParams &p, const pm3d_list_query &q, int i) {
    if (p.render_distance < 0 || q.target == i) return true;
    const double sx = __ldg(&p.world[i].x), sy = __ldg(&p.world[i].y);
    return fabs(sx - q.cam_x) <= p.reach && fabs(sy - q.cam_y) <= p.reach;
}

__device__ __forceinline__ void copy_sprite(pm3d_sprite *dst, const pm3d_sprite *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2), e = __ldg(s + 3);
    d[0] = a; d[1] = b; d[2] = c; d[3] = e;
}

__global__ void __launch_bounds__(LIST_THREADS) k_list_count(const __grid_constant__ ListParams p) {
    const int v = blockIdx.x;
    const pm3d_list_query q = p.queries[v];
    int kept = 0;
    for (int i0 = 0; i0 < p.n_world; i0 += LIST_THREADS) {
        const int i = i0 + threadIdx.x;
        kept += __syncthreads_count(i < p.n_world && do_sprite(p, q, i));
    }
    if (threadIdx.x == 0) {
        int c = kept;
        if (p.clutter) c += max(0, min(p.clutter_count[v], p.clutter_stride));
        if (p.extras) c += max(0, p.extra_offset[v + 1] - p.extra_offset[v]);
        p.out_offset[v + 1] = c;
    }
}

// exclusive scan of the counts in place: out_offset[0] = 0, out_offset[v + 1] = sum of the counts of cameras <= v
__global__ void __launch_bounds__(1024) k_list_scan(int32_t *off, int n, int max_total, uint32_t *status) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry = 0; off[0] = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int c = i < n ? off[i + 1] : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;  // inclusive over warps
        }
        __syncthreads();
        const int base = carry + (warp ? warp_tot[warp - 1] : 0);
        if (i < n) off[i + 1] = base + incl;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    // an enqueue-only caller (PM3D_NO_SYNC) never sees the total: more sprites than max_total raises the ctx's error
    // word (k_list_write cuts the lists at max_total; pm3d_sync returns PM3D_ERR_CAPACITY)
    if (threadIdx.x == 0 && carry > max_total) {
        *reinterpret_cast<volatile uint32_t *>(status + PM3D_STATUS_LIST_TOTAL) = 1u;
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(LIST_THREADS) k_list_write(const __grid_constant__ ListParams p) {
    __shared__ int warp_cnt[2][LIST_THREADS / 32];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const pm3d_list_query q = p.queries[v];
    const int start = p.out_offset[v];
    auto put = [&](int pos, const pm3d_sprite *src, int index) {
        if (pos < p.max_total) {
            copy_sprite(p.out_sprites + pos, src);
            if (p.out_index) p.out_index[pos] = index;
        }
    };
    int n_kept = 0, pp = 0;
    for (int i0 = 0; i0 < p.n_world; i0 += LIST_THREADS, pp ^= 1) {
        const int i = i0 + tid;
        const bool keep = i < p.n_world && do_sprite(p, q, i);
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[pp][warp] = __popc(b);
        __syncthreads();
        int base = n_kept;
        for (int w = 0; w < LIST_THREADS / 32; w++) {
            const int c = warp_cnt[pp][w];
            if (w < warp) base += c;
            n_kept += c;
        }
        if (keep) put(start + base + __popc(b & ((1u << lane) - 1u)), p.world + i, i);
    }
    int pos = start + n_kept;
    if (p.clutter) {
        const int nc = max(0, min(p.clutter_count[v], p.clutter_stride));
        for (int k = tid; k < nc; k += LIST_THREADS) put(pos + k, p.clutter + (size_t)v * p.clutter_stride + k, -1 - k);
        pos += nc;
    }
    if (p.extras) {
        const int e0 = p.extra_offset[v], ne = max(0, p.extra_offset[v + 1] - e0);
        for (int k = tid; k < ne; k += LIST_THREADS) put(pos + k, p.extras + e0 + k, INT32_MIN + k);
    }
}

int ensure_dev(pm3d_ctx *ctx, void **buf, size_t *have, size_t need) {
    if (*have >= need) return PM3D_OK;
    if (*buf) PM3D_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr; *have = 0;
    PM3D_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return PM3D_OK;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" int pm3d_raycast_sprites_batch(pm3d_ctx *ctx, const pm3d_sprite *world, int n_world,
                                          const pm3d_list_query *queries, int n, double render_distance,
                                          const pm3d_sprite *clutter, const int32_t *clutter_count, int clutter_stride,
                                          const pm3d_sprite *extras, const int32_t *extra_offset, int max_total,
                                          pm3d_sprite *out_sprites, int32_t *out_offset, int32_t *out_index,
                                          int32_t *out_total, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || n_world < 0 || max_total < 1 || (n_world > 0 && !world) || (n > 0 && (!queries || !out_sprites || !out_offset)))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "sprite list batch: null pointers or negative sizes");
    if ((clutter != nullptr) != (clutter_count != nullptr) || (clutter && clutter_stride < 1) ||
        (extras != nullptr) != (extra_offset != nullptr))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "sprite list batch: clutter / extras need both their arrays");
    if (render_distance != render_distance) return pm3d_fail(ctx, PM3D_ERR_ARG, "render_distance is NaN");
    if (n == 0) { if (out_total) *out_total = 0; return PM3D_OK; }
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool in_dev = flags & PM3D_IN_DEVICE, out_dev = flags & PM3D_OUT_DEVICE;
    int rc;
    if ((rc = pm3d_main_enter(ctx))) return rc;
    ListParams p;
    p.world = world; p.queries = queries; p.clutter = clutter; p.clutter_count = clutter_count;
    p.extras = extras; p.extra_offset = extra_offset;
    if (!in_dev) {
        // one staging block: world | queries | clutter_count | clutter | extra_offset | extras
        int n_extras = 0;
        if (extras) {
            n_extras = extra_offset[n];
            if (n_extras < 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "extra_offset[n] < 0");
        }
        const size_t b_world = align256((size_t)n_world * sizeof(pm3d_sprite)), b_q = align256((size_t)n * sizeof(pm3d_list_query));
        const size_t b_cc = clutter ? align256((size_t)n * 4) : 0, b_cl = clutter ? align256((size_t)n * clutter_stride * sizeof(pm3d_sprite)) : 0;
        const size_t b_eo = extras ? align256((size_t)(n + 1) * 4) : 0, b_ex = extras ? align256((size_t)n_extras * sizeof(pm3d_sprite)) : 0;
        if ((rc = ensure_dev(ctx, &ctx->io_a, &ctx->io_a_bytes, b_world + b_q + b_cc + b_cl + b_eo + b_ex + 256))) return rc;
        uint8_t *d = (uint8_t *)ctx->io_a;
        auto up = [&](const void *src, size_t bytes, size_t slot) -> const void * {
            const void *at = d;
            if (bytes) cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
            d += slot;
            return at;
        };
        p.world = (const pm3d_sprite *)up(world, (size_t)n_world * sizeof(pm3d_sprite), b_world);
        p.queries = (const pm3d_list_query *)up(queries, (size_t)n * sizeof(pm3d_list_query), b_q);
        if (clutter) {
            p.clutter_count = (const int32_t *)up(clutter_count, (size_t)n * 4, b_cc);
            p.clutter = (const pm3d_sprite *)up(clutter, (size_t)n * clutter_stride * sizeof(pm3d_sprite), b_cl);
        }
        if (extras) {
            p.extra_offset = (const int32_t *)up(extra_offset, (size_t)(n + 1) * 4, b_eo);
            p.extras = (const pm3d_sprite *)up(extras, (size_t)n_extras * sizeof(pm3d_sprite), b_ex);
        }
        PM3D_CUDA(ctx, cudaGetLastError());
    }
    const size_t b_off = align256((size_t)(n + 1) * 4), b_spr = align256((size_t)max_total * sizeof(pm3d_sprite)),
                 b_idx = out_index ? align256((size_t)max_total * 4) : 0;
    p.out_sprites = out_sprites; p.out_offset = out_offset; p.out_index = out_index;
    if (!out_dev) {
        if ((rc = ensure_dev(ctx, &ctx->io_c, &ctx->io_c_bytes, b_off + b_spr + b_idx))) return rc;
        p.out_offset = (int32_t *)ctx->io_c;
        p.out_sprites = (pm3d_sprite *)((uint8_t *)ctx->io_c + b_off);
        p.out_index = out_index ? (int32_t *)((uint8_t *)ctx->io_c + b_off + b_spr) : nullptr;
    }
    p.render_distance = render_distance;
    p.reach = ceil(render_distance);
    p.n_world = n_world; p.n = n; p.clutter_stride = clutter_stride; p.max_total = max_total;
    k_list_count<<<n, LIST_THREADS, 0, ctx->stream>>>(p);
    k_list_scan<<<1, 1024, 0, ctx->stream>>>(p.out_offset, n, max_total, ctx->status_d);
    k_list_write<<<n, LIST_THREADS, 0, ctx->stream>>>(p);
    ctx->launches += 3;
    PM3D_CUDA(ctx, cudaGetLastError());
    pm3d_main_touch(ctx);
    if (flags & PM3D_NO_SYNC) {
        if (!out_dev) return pm3d_fail(ctx, PM3D_ERR_ARG, "PM3D_NO_SYNC needs PM3D_OUT_DEVICE here (the copy size is the device-side total)");
        return PM3D_OK;
    }
    int32_t total = 0;
    PM3D_CUDA(ctx, cudaMemcpyAsync(&total, p.out_offset + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_total) *out_total = total;
    if (total > max_total) ctx->status_h[PM3D_STATUS_LIST_TOTAL] = 0;  // reported right here
    if (total > max_total) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "the cameras list %d sprites, max_total is %d", total, max_total);
    if (!out_dev) {
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_offset, p.out_offset, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (total) PM3D_CUDA(ctx, cudaMemcpyAsync(out_sprites, p.out_sprites, (size_t)total * sizeof(pm3d_sprite), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_index && total) PM3D_CUDA(ctx, cudaMemcpyAsync(out_index, p.out_index, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PM3D_OK;
}
