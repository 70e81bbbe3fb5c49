// path_kernel.cu — batched Pathing.FindPath.
//
// Replaces, for a batch of independent (start, finish) queries — one per AI unit and tick, call site
// game/ai_movement.go:64 — Pathing.FindPath (game/model/pathing.go:123-186): astar.Path of go-astar over the level-0
// tile world (Tile.PathNeighbors / PathNeighborCost / PathEstimatedCost, pathing.go:52-92; a tile is a blocker iff
// Level(0)[x][y] != 0, pathing.go:94-112), the reversal into cell centres and the Bezier smoothing in segments of
// seven steps (game/common/bezier/bezier.go).
//
// WHICH shortest path comes back is decided by the reference's priority queue (container/heap ordered by `rank <`
// alone: on a uniform grid with the Manhattan estimate most of the queue ties), so the search cannot be replaced
// by a parallel wavefront: every query runs the reference's exact sequence of heap operations.  The batch is the
// parallel axis — one thread per query, a warp's 32 searches advance in lockstep so their (dependent, scattered)
// loads are in flight together; the chip holds thousands of searches.  Per query, in HBM:
//   idx[cell]  u32  node.index (position in the queue)
//   st[cell]   u32  node.cost (27 bits) | parent direction + 1 (3 bits) | open | closed   — 0 = a fresh node
//   heap[k]    u64  rank << 32 | cell
// All of it integer work: costs are step counts, ranks step counts + Manhattan distance.  The Bezier part is the
// reference's binary64 sequence (no contraction, -fmad=false).
#include <algorithm>

#include "pm3d_internal.cuh"

struct PathParams {
    const uint16_t *level0;
    const double4 *queries;
    uint32_t *idx, *st;      // [wave][cells]
    uint64_t *heap;          // [wave][cells]
    double2 *out_points;     // [n][max_points]
    int32_t *out_count;      // [n]
    int64_t n, first;        // queries of this wave: [first, first + wave)
    int map_w, map_h, max_points, wave;
};

#define ST_CLOSED 1u
#define ST_OPEN 2u
#define ST_PARENT_SHIFT 2
#define ST_COST_SHIFT 5

struct PathHeap {
    uint64_t *h;
    uint32_t *idx;
    int n;
    __device__ __forceinline__ static uint32_t rank_of(uint64_t e) { return (uint32_t)(e >> 32); }
    __device__ __forceinline__ void put(int i, uint64_t e) { h[i] = e; idx[(uint32_t)e] = (uint32_t)i; }
    // container/heap up(j): the entry `e` (held in a register) moves towards the root while its parent is strictly larger
    __device__ __forceinline__ void up(int j, uint64_t e) {
        while (j > 0) {
            const int i = (j - 1) / 2;
            const uint64_t pe = h[i];
            if (!(rank_of(e) < rank_of(pe))) break;
            put(j, pe);
            j = i;
        }
        put(j, e);
    }
    // container/heap down(i, n): returns whether the entry moved
    __device__ __forceinline__ bool down(int i0, uint64_t e, int n_) {
        int i = i0;
        for (;;) {
            const int j1 = 2 * i + 1;
            if (j1 >= n_) break;
            int j = j1;
            uint64_t ce = h[j1];
            if (j1 + 1 < n_) {
                const uint64_t c2 = h[j1 + 1];
                if (rank_of(c2) < rank_of(ce)) { j = j1 + 1; ce = c2; }
            }
            if (!(rank_of(ce) < rank_of(e))) break;
            put(i, ce);
            i = j;
        }
        put(i, e);
        return i > i0;
    }
    __device__ __forceinline__ void push(uint64_t e) { up(n++, e); }
    // heap.Pop: swap(0, n-1), down(0, n-1), take the last
    __device__ __forceinline__ uint64_t pop() {
        const uint64_t top = h[0];
        n--;
        if (n > 0) down(0, h[n], n);
        return top;
    }
    // heap.Remove(i): swap(i, n-1); if !down(i, n-1) up(i); take the last
    __device__ __forceinline__ void remove(int i) {
        n--;
        if (n != i) {
            const uint64_t e = h[n];
            if (!down(i, e, n)) up(i, e);
        }
    }
};

// bezier.New(cp).Curve(p), len(p) = 2 len(cp): the reference's binary64 sequence
__device__ void bezier_segment(const int32_t *steps, int map_h, int n, double2 *out, int64_t &n_out, int cap) {
    double cx[9], cy[9];
    double w = 0;
    for (int i = 0; i < n; i++) {
        if (i == 0) w = 1;
        else if (i == 1) w = (double)n - 1;
        else w *= (double)(n - i) / (double)i;
        const int c = steps[i];
        cx[i] = ((double)(c / map_h) + 0.5) * w;
        cy[i] = ((double)(c % map_h) + 0.5) * w;
    }
    const int np = 2 * n;
    const double nf = (double)(np - 1);
    for (int k = 0; k < np; k++) {
        const double t = (double)k / nf;
        double px[9], py[9];
        px[0] = cx[0]; py[0] = cy[0];
        double u = t;
        for (int i = 1; i < n; i++) { px[i] = cx[i] * u; py[i] = cy[i] * u; u *= t; }
        const double t1 = 1 - t;
        double tt = t1;
        double x = px[n - 1], y = py[n - 1];
        for (int i = n - 2; i >= 0; i--) { x += px[i] * tt; y += py[i] * tt; tt *= t1; }
        if (n_out < cap) out[n_out] = make_double2(x, y);
        n_out++;
    }
}

__global__ void __launch_bounds__(32) k_find_path(PathParams p) {
    const int slot = blockIdx.x * 32 + threadIdx.x;
    const int64_t qi = p.first + slot;
    if (slot >= p.wave || qi >= p.n) return;
    const size_t cells = (size_t)p.map_w * p.map_h;
    uint32_t *idx = p.idx + (size_t)slot * cells, *st = p.st + (size_t)slot * cells;
    PathHeap hp;
    hp.h = p.heap + (size_t)slot * cells; hp.idx = idx; hp.n = 0;
    double2 *out = p.out_points + (size_t)qi * p.max_points;
    const double4 q = p.queries[qi];
    const int64_t ix = go_int(q.x), iy = go_int(q.y), jx = go_int(q.z), jy = go_int(q.w);
    if (ix < 0 || iy < 0 || ix >= p.map_w || iy >= p.map_h || jx < 0 || jy < 0 || jx >= p.map_w || jy >= p.map_h) {
        p.out_count[qi] = -1;  // a nil tile
        return;
    }
    const int H = p.map_h;
    const uint32_t from = (uint32_t)(ix * H + iy), to = (uint32_t)(jx * H + jy);
    const int tx = (int)jx, ty = (int)jy;

    // ---- astar.Path
    st[from] = ST_OPEN;
    hp.push((uint64_t)from);  // rank 0
    bool found = false;
    while (hp.n > 0) {
        const uint32_t cur = (uint32_t)hp.pop();
        const uint32_t cs = st[cur];
        st[cur] = (cs & ~ST_OPEN) | ST_CLOSED;
        if (cur == to) { found = true; break; }
        const int cx = (int)(cur / (uint32_t)H), cy = (int)(cur % (uint32_t)H);
        const uint32_t cost = (cs >> ST_COST_SHIFT) + 1u;
        // the four neighbours' cells and states are fetched together, then handled in the reference's order
        int nxs[4], nys[4];
        uint32_t ns[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            nxs[k] = cx + (k == 0 ? -1 : (k == 1 ? 1 : 0));
            nys[k] = cy + (k == 2 ? -1 : (k == 3 ? 1 : 0));
            ok[k] = (unsigned)nxs[k] < (unsigned)p.map_w && (unsigned)nys[k] < (unsigned)p.map_h;
            const uint32_t nb = (uint32_t)(nxs[k] * H + nys[k]);
            ok[k] = ok[k] && __ldg(p.level0 + (ok[k] ? nb : 0u)) == 0;
            ns[k] = ok[k] ? st[nb] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!ok[k]) continue;
            const uint32_t nb = (uint32_t)(nxs[k] * H + nys[k]);
            uint32_t s = ns[k];
            if (cost < (s >> ST_COST_SHIFT)) {
                if (s & ST_OPEN) hp.remove((int)idx[nb]);
                s &= ~(ST_OPEN | ST_CLOSED);
            }
            if (!(s & (ST_OPEN | ST_CLOSED))) {
                const uint32_t rank = cost + (uint32_t)(abs(tx - nxs[k]) + abs(ty - nys[k]));
                st[nb] = (cost << ST_COST_SHIFT) | ((uint32_t)(k + 1) << ST_PARENT_SHIFT) | ST_OPEN;
                hp.push(((uint64_t)rank << 32) | nb);
            }
        }
    }
    if (!found) { p.out_count[qi] = -1; return; }

    // ---- the path, goal -> start along the parents; the queue's memory now holds the steps, start -> goal
    int64_t len = 0;
    for (uint32_t c = to;; ) {
        len++;
        const uint32_t d = (st[c] >> ST_PARENT_SHIFT) & 7u;
        if (d == 0u) break;
        const int k = (int)d - 1;
        c = (uint32_t)((int)c - (k == 0 ? -H : (k == 1 ? H : (k == 2 ? -1 : 1))));
    }
    int32_t *steps = reinterpret_cast<int32_t *>(hp.h);
    {
        int64_t k = len - 1;
        for (uint32_t c = to;; k--) {
            steps[k] = (int32_t)c;
            const uint32_t d = (st[c] >> ST_PARENT_SHIFT) & 7u;
            if (d == 0u) break;
            const int dk = (int)d - 1;
            c = (uint32_t)((int)c - (dk == 0 ? -H : (dk == 1 ? H : (dk == 2 ? -1 : 1))));
        }
    }
    // ---- FindPath: as-is below three steps, else Bezier curves over segments of seven steps
    int64_t n_out = 0;
    auto emit_step = [&](int64_t k) {
        const int c = steps[k];
        if (n_out < p.max_points) out[n_out] = make_double2((double)(c / H) + 0.5, (double)(c % H) + 0.5);
        n_out++;
    };
    if (len < 3) {
        for (int64_t k = 0; k < len; k++) emit_step(k);
    } else {
        int64_t segmentIndex = 0, segmentSize = 7, remainingSteps = len;
        while (remainingSteps > 0) {
            if (segmentSize <= 1) { emit_step(segmentIndex); remainingSteps--; segmentIndex++; continue; }
            if (remainingSteps < segmentSize) segmentSize = remainingSteps;
            remainingSteps -= segmentSize - 1;
            if (remainingSteps > 0 && remainingSteps < 3) { segmentSize += remainingSteps - 1; remainingSteps = 0; }
            bezier_segment(steps + segmentIndex, H, (int)segmentSize, out, n_out, p.max_points);
            segmentIndex += segmentSize - 1;
        }
    }
    p.out_count[qi] = (int32_t)min(n_out, (int64_t)0x7fffffff);
}

extern "C" int pm3d_find_path_batch(pm3d_ctx *ctx, const double *queries, int64_t n, int max_points, double *out_points,
                                    int32_t *out_count, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || max_points < 1 || (n > 0 && (!queries || !out_points || !out_count)))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_find_path_batch: null pointers / max_points < 1");
    if (!ctx->grid) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_grid must be called first");
    const size_t cells = (size_t)ctx->map_w * ctx->map_h;
    if (cells >= (1u << 27)) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "map too large for the path search (2^27 cells)");
    if (n == 0) return PM3D_OK;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc_ = pm3d_main_enter(ctx); if (rc_) return rc_; }
    pm3d_main_touch(ctx);
    // searches in flight: 16 bytes of scratch per cell and search, whole warps.  A wave ends with its slowest search, so
    // as many as possible run at once: up to a third of the free HBM (at most 48 GiB) — 180 GB per GPU is what makes
    // 4096 searches over a 500x500 map (16 GiB) one wave
    size_t free_b = 0, total_b = 0;
    PM3D_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const int64_t budget = std::max<int64_t>((int64_t)1 << 30, std::min<int64_t>((int64_t)((free_b + ctx->path_scratch_bytes) / 3), (int64_t)48 << 30));
    int64_t wave = std::min<int64_t>(n, std::max<int64_t>(32, budget / (int64_t)(16 * cells)));
    wave = std::min<int64_t>(wave, 65536);
    wave = (wave + 31) / 32 * 32;
    const size_t need = (size_t)wave * cells * 16;
    if (ctx->path_scratch_bytes < need) {
        if (ctx->path_scratch) PM3D_CUDA(ctx, cudaFree(ctx->path_scratch));
        ctx->path_scratch = nullptr; ctx->path_scratch_bytes = 0;
        PM3D_CUDA(ctx, cudaMalloc(&ctx->path_scratch, need));
        ctx->path_scratch_bytes = need;
    }
    const bool in_dev = flags & PM3D_IN_DEVICE, out_dev = flags & PM3D_OUT_DEVICE;
    const size_t out_bytes = (size_t)n * max_points * 16;
    void *d_q = nullptr, *d_pts = nullptr, *d_cnt = nullptr;
    if (!in_dev) {
        PM3D_CUDA(ctx, cudaMallocAsync(&d_q, (size_t)n * 32, ctx->stream));
        PM3D_CUDA(ctx, cudaMemcpyAsync(d_q, queries, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!out_dev) {
        PM3D_CUDA(ctx, cudaMallocAsync(&d_pts, out_bytes, ctx->stream));
        PM3D_CUDA(ctx, cudaMallocAsync(&d_cnt, (size_t)n * 4, ctx->stream));
    }
    PathParams p;
    p.level0 = ctx->grid;
    p.queries = (const double4 *)(in_dev ? (const void *)queries : d_q);
    p.idx = (uint32_t *)ctx->path_scratch;
    p.st = p.idx + (size_t)wave * cells;
    p.heap = (uint64_t *)(p.st + (size_t)wave * cells);
    p.out_points = (double2 *)(out_dev ? (void *)out_points : d_pts);
    p.out_count = (int32_t *)(out_dev ? (void *)out_count : d_cnt);
    p.n = n; p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.max_points = max_points; p.wave = (int)wave;
    for (int64_t first = 0; first < n; first += wave) {
        p.first = first;
        // fresh nodes read as zero (cost 0, neither open nor closed); the queue index and the queue need no clearing
        PM3D_CUDA(ctx, cudaMemsetAsync(p.st, 0, (size_t)wave * cells * 4, ctx->stream));
        k_find_path<<<(unsigned)(wave / 32), 32, 0, ctx->stream>>>(p);
        ctx->launches++;
        PM3D_CUDA(ctx, cudaGetLastError());
    }
    if (!out_dev) {
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_points, d_pts, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_count, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PM3D_CUDA(ctx, cudaFreeAsync(d_pts, ctx->stream));
        PM3D_CUDA(ctx, cudaFreeAsync(d_cnt, ctx->stream));
    }
    if (!in_dev) PM3D_CUDA(ctx, cudaFreeAsync(d_q, ctx->stream));
    if (!(flags & PM3D_NO_SYNC) || !out_dev || !in_dev) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}
