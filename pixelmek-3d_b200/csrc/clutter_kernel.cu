// clutter_kernel.cu — ClutterHandler.Update (game/clutter.go:25-117) for a batch of cameras.
//
// The reference walks 21 sample rays x (clutterDistance + 1) distances per camera, and for every NEW cell it lands on
// re-seeds a Go math/rand generator with Map.Seed + posId (607-word additive lagged Fibonacci state filled by 1841
// steps of a Lehmer generator, XORed with the library's rngCooked table) only to draw 1-3 numbers per clutter type.
// Here:
//   * one CTA per camera; the samples' cells go through a shared-memory hash table keyed by posId that keeps the
//     FIRST sample (in the reference's loop order) of every cell, so the output order is the creation order;
//   * the generator is never materialised.  Draw k (k <= 273) of a freshly seeded source is
//     vec[334-k] + vec[607-k] of the INITIAL state (the feed position only overwrites words that draws <= 273 never
//     read back), and an initial word is three Lehmer values x0*A^n mod (2^31-1) with fixed exponents n — so a draw
//     is six modular multiplications by table constants plus two table words: O(1) per draw instead of O(607) per
//     cell.  The table (A^n for the 2x3 exponents and the two rngCooked words of each k) is built once on the host;
//     rngCooked itself is regenerated the way the Go library's gen_cooked.go does (7.8e12 steps from seed 1) by
//     polynomial jump-ahead.  tests/test_clutter_*.py pin the generator against Go's published seed-1 outputs.
#include <string.h>

#include <mutex>
#include <vector>

#include "pm3d_internal.cuh"

namespace {

constexpr int RNG_LEN = 607, RNG_TAP = 273;
constexpr uint32_t LEHMER_M = 2147483647u, LEHMER_A = 48271u;
constexpr int MAX_DRAWS = RNG_TAP;  // draws of a fresh source that depend on its initial words only

// ---------------------------------------------------------------- host: rngCooked by jump-ahead

using Poly = std::vector<uint64_t>;  // degree < 607, coefficients mod 2^64, modulo x^607 = x^334 + 1

Poly poly_mul(const Poly &a, const Poly &b) {
    std::vector<uint64_t> c(2 * RNG_LEN - 1, 0);
    for (int i = 0; i < RNG_LEN; i++) {
        const uint64_t ai = a[i];
        if (!ai) continue;
        for (int j = 0; j < RNG_LEN; j++) c[i + j] += ai * b[j];
    }
    for (int d = 2 * RNG_LEN - 2; d >= RNG_LEN; d--) {
        c[d - RNG_TAP] += c[d];  // x^d = x^(d-607) * (x^334 + 1)
        c[d - RNG_LEN] += c[d];
    }
    c.resize(RNG_LEN);
    return c;
}

uint32_t lehmer_mul(uint32_t x, uint32_t t) { return (uint32_t)(((uint64_t)x * t) % LEHMER_M); }

// the state gen_cooked.go prints: srand(1) with the 20/10-bit packing, then `steps` calls of vrand()
void cooked_table(uint64_t steps, uint64_t *out) {
    // sequence view: y_k = y_{k-607} + y_{k-273}; call k (1-based) overwrites ring position (334 - k) mod 607
    uint64_t seq[RNG_LEN];  // y_{-606} .. y_0
    {
        uint32_t x = 1;
        for (int i = 0; i < 20; i++) x = lehmer_mul(x, LEHMER_A);
        uint64_t vec[RNG_LEN];
        for (int i = 0; i < RNG_LEN; i++) {
            const uint32_t a = lehmer_mul(x, LEHMER_A), b = lehmer_mul(a, LEHMER_A), c = lehmer_mul(b, LEHMER_A);
            vec[i] = ((uint64_t)a << 20) ^ ((uint64_t)b << 10) ^ (uint64_t)c;
            x = c;
        }
        for (int p = 0; p < RNG_LEN; p++) {
            const int k0 = p >= 334 ? 334 - p : -273 - p;  // the (virtual) call that wrote position p
            seq[k0 + 606] = vec[p];
        }
    }
    Poly q(RNG_LEN, 0), b(RNG_LEN, 0);
    q[0] = 1;
    b[1] = 1;
    for (uint64_t e = steps; e; e >>= 1) {
        if (e & 1) q = poly_mul(q, b);
        b = poly_mul(b, b);
    }
    for (int k = 0; k < RNG_LEN; k++) {  // y_{steps - 606 + k} = <x^(steps + k), seq>
        uint64_t acc = 0;
        for (int j = 0; j < RNG_LEN; j++) acc += q[j] * seq[j];
        const long long m = (long long)steps - 606 + k;
        out[(int)(((334 - m) % RNG_LEN + RNG_LEN) % RNG_LEN)] = acc;
        const uint64_t top = q[RNG_LEN - 1];
        for (int j = RNG_LEN - 1; j > 0; j--) q[j] = q[j - 1];
        q[0] = top;
        q[RNG_LEN - RNG_TAP] += top;
    }
}

const uint64_t *go_rand_cooked() {
    static uint64_t table[RNG_LEN];
    static std::once_flag once;
    std::call_once(once, [] { cooked_table(7800000000000ull, table); });
    return table;
}

// ---------------------------------------------------------------- device

// what draw k of a freshly seeded source needs: word 334-k ("feed") and word 607-k ("tap") of the initial state
struct DrawEntry {
    uint64_t cooked_feed, cooked_tap;
    uint32_t pow_feed[3], pow_tap[3];  // A^(21+3i), A^(22+3i), A^(23+3i) mod 2^31-1 for i = 334-k / 607-k
    uint32_t _pad[2];
};
static_assert(sizeof(DrawEntry) == 48, "three 16-byte loads");

struct ClutterParams {
    const pm3d_clutter_query *queries;
    const uint16_t *level0;
    const uint8_t *cell_class;
    const uint64_t *class_mask;
    const pm3d_clutter_def *defs;
    const DrawEntry *draws;  // [MAX_DRAWS + 1], entry 0 unused
    int32_t *out_count;
    pm3d_clutter_sprite *out_records;
    pm3d_sprite *out_sprites;
    int64_t map_seed;
    int map_w, map_h, n_defs, n_dist, slots, max_per_view;
};

__device__ __forceinline__ uint32_t lehmer_mulmod(uint32_t x, uint32_t t) {
    const uint64_t p = (uint64_t)x * t;  // < 2^62 = hi * 2^31 + lo with hi, lo < 2^31, and 2^31 = 1 (mod M)
    uint32_t r = ((uint32_t)p & LEHMER_M) + (uint32_t)(p >> 31);
    r = (r & LEHMER_M) + (r >> 31);
    return r >= LEHMER_M ? r - LEHMER_M : r;
}

__device__ __forceinline__ uint64_t initial_word(uint32_t x0, const uint32_t pw[3], uint64_t cooked) {
    return (((uint64_t)lehmer_mulmod(x0, pw[0]) << 40) ^ ((uint64_t)lehmer_mulmod(x0, pw[1]) << 20) ^
            (uint64_t)lehmer_mulmod(x0, pw[2])) ^ cooked;
}

// Rand.Float64 of a source seeded with x0, draw counter k (1-based, advanced): float64(Int63()) / 2^63, resampled on 1
__device__ __forceinline__ double go_float64(const DrawEntry *draws, uint32_t x0, int &k) {
    for (;;) {
        const int kk = k < MAX_DRAWS ? k : MAX_DRAWS;  // unreachable for n_defs <= 64 (<= 192 draws + O(never) resamples)
        k++;
        const uint4 *e = reinterpret_cast<const uint4 *>(draws + kk);
        const uint4 e0 = __ldg(e), e1 = __ldg(e + 1), e2 = __ldg(e + 2);
        const uint32_t pf[3] = {e1.x, e1.y, e1.z}, pt[3] = {e1.w, e2.x, e2.y};
        const uint64_t cf = (uint64_t)e0.x | ((uint64_t)e0.y << 32), ct = (uint64_t)e0.z | ((uint64_t)e0.w << 32);
        const uint64_t u = initial_word(x0, pf, cf) + initial_word(x0, pt, ct);
        const double f = __ll2double_rn((long long)(u & 0x7fffffffffffffffull)) * 0x1p-63;
        if (f != 1.0) return f;
    }
}

__device__ __forceinline__ int64_t pos_id(uint32_t cx, uint32_t cy, int map_w) {
    return ((int64_t)cx - 1) * (int64_t)map_w + (int64_t)cy;
}

constexpr uint32_t NO_CELL = 0xffffffffu;
constexpr int CLUTTER_THREADS = 128;

// dynamic shared memory: uint32 cell_x[S], cell_y[S] (the cell of every sample), order[S] (first-visit samples in loop
// order), table[slots] (posId -> first sample)
__global__ void __launch_bounds__(CLUTTER_THREADS) k_clutter(const __grid_constant__ ClutterParams p) {
    extern __shared__ uint32_t smem[];
    const int view = blockIdx.x, tid = threadIdx.x;
    const pm3d_clutter_query &q = p.queries[view];
    const int n_angles = min(max(q.n_angles, 0), PM3D_CLUTTER_MAX_ANGLES);
    const int S = n_angles * p.n_dist;
    uint32_t *cell_x = smem, *cell_y = smem + S, *order = smem + 2 * S, *table = smem + 3 * S;
    const uint32_t mask = (uint32_t)p.slots - 1u;
    __shared__ int warp_sum[2][CLUTTER_THREADS / 32];

    for (int i = tid; i < p.slots; i += CLUTTER_THREADS) table[i] = NO_CELL;
    // ---- the cell of every sample (clutter.go:46-57)
    const double camx = q.cam_x, camy = q.cam_y;
    for (int s = tid; s < S; s += CLUTTER_THREADS) {
        const int ai = s / p.n_dist;
        const double d = (double)(s - ai * p.n_dist);
        const double x2 = camx + d * q.cs[ai][0];  // geom.LineFromAngle
        const double y2 = camy + d * q.cs[ai][1];
        const int64_t x = go_int(x2), y = go_int(y2);
        uint32_t cx = NO_CELL, cy = NO_CELL;
        if (!(x < 0 || y < 0 || x >= (int64_t)p.map_w || y >= (int64_t)p.map_h) &&
            !(__ldg(p.level0 + (size_t)x * p.map_h + y) > 0)) {  // IsWallAt(0, x, y)
            cx = (uint32_t)x; cy = (uint32_t)y;
        }
        cell_x[s] = cx; cell_y[s] = cy;
    }
    __syncthreads();
    // ---- spritesByPosition: the first sample of every posId
    auto slot_of = [&](int64_t id) -> uint32_t { return (uint32_t)(((uint64_t)id * 0x9e3779b97f4a7c15ull) >> 40) & mask; };
    for (int s = tid; s < S; s += CLUTTER_THREADS) {
        if (cell_x[s] == NO_CELL) continue;
        const int64_t id = pos_id(cell_x[s], cell_y[s], p.map_w);
        uint32_t h = slot_of(id);
        for (;;) {
            const uint32_t prev = atomicCAS(&table[h], NO_CELL, (uint32_t)s);
            if (prev == NO_CELL) break;
            if (pos_id(cell_x[prev], cell_y[prev], p.map_w) == id) { atomicMin(&table[h], (uint32_t)s); break; }
            h = (h + 1u) & mask;
        }
    }
    __syncthreads();

    // ---- the first-visit samples, compacted in loop order (so the lanes that draw numbers below are all busy)
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CLUTTER_THREADS / 32;
    int n_cells = 0;  // (every thread keeps the same running totals)
    int pp = 0;       // ping-pong slot of warp_sum: one barrier per chunk
    for (int s0 = 0; s0 < S; s0 += CLUTTER_THREADS, pp ^= 1) {
        const int s = s0 + tid;
        bool first = false;
        if (s < S && cell_x[s] != NO_CELL) {
            const int64_t id = pos_id(cell_x[s], cell_y[s], p.map_w);
            uint32_t h = slot_of(id);
            for (;;) {
                const uint32_t t = table[h];
                if (pos_id(cell_x[t], cell_y[t], p.map_w) == id) { first = t == (uint32_t)s; break; }
                h = (h + 1u) & mask;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, first);
        if (lane == 0) warp_sum[pp][warp] = __popc(b);
        __syncthreads();
        int base = n_cells;
        for (int w = 0; w < NW; w++) {
            const int c = warp_sum[pp][w];
            if (w < warp) base += c;
            n_cells += c;
        }
        if (first) order[base + __popc(b & ((1u << lane) - 1u))] = (uint32_t)s;
    }
    __syncthreads();

    // ---- per cell: seed, draw, count; block scan; draw again and write in creation order
    int n_out = 0;
    for (int c0 = 0; c0 < n_cells; c0 += CLUTTER_THREADS, pp ^= 1) {
        const int c = c0 + tid;
        const bool live = c < n_cells;
        uint32_t cx = 0, cy = 0, x0 = 0;
        uint64_t match = 0;
        int cnt = 0;
        if (live) {
            const uint32_t s = order[c];
            cx = cell_x[s]; cy = cell_y[s];
            match = __ldg(p.class_mask + __ldg(p.cell_class + (size_t)cx * p.map_h + cy));
            // rngSource.Seed(Map.Seed + posId): seed %= 2^31-1; negative += 2^31-1; 0 -> 89482311
            int64_t seed = (int64_t)((uint64_t)p.map_seed + (uint64_t)pos_id(cx, cy, p.map_w)) % (int64_t)LEHMER_M;
            if (seed < 0) seed += (int64_t)LEHMER_M;
            if (seed == 0) seed = 89482311;
            x0 = (uint32_t)seed;
            int k = 1;
            for (uint64_t m = match; m; m &= m - 1) {
                const int i = __ffsll((long long)m) - 1;
                if (go_float64(p.draws, x0, k) <= p.defs[i].frequency) {
                    cnt++;
                    go_float64(p.draws, x0, k);  // the two position draws (a resample would move k by more than 1)
                    go_float64(p.draws, x0, k);
                }
            }
        }
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[pp][warp] = incl;
        __syncthreads();
        int off = n_out + incl - cnt;
        for (int w = 0; w < NW; w++) {
            const int t = warp_sum[pp][w];
            if (w < warp) off += t;
            n_out += t;
        }
        if (cnt > 0) {
            int k = 1;
            for (uint64_t m = match; m; m &= m - 1) {
                const int i = __ffsll((long long)m) - 1;
                const pm3d_clutter_def &df = p.defs[i];
                if (!(go_float64(p.draws, x0, k) <= df.frequency)) continue;
                const double sx = (double)cx + go_float64(p.draws, x0, k);
                const double sy = (double)cy + go_float64(p.draws, x0, k);
                if (off < p.max_per_view) {
                    const size_t o = (size_t)view * p.max_per_view + off;
                    pm3d_clutter_sprite r;
                    r.x = sx; r.y = sy; r.cell_x = (int32_t)cx; r.cell_y = (int32_t)cy; r.def = i; r._pad = 0;
                    p.out_records[o] = r;
                    if (p.out_sprites) {
                        pm3d_sprite sp;
                        sp.x = sx; sp.y = sy; sp.z = 0.0; sp.scale = df.scale; sp.illumination = 0.0;
                        sp.sheet = df.sheet; sp.rx0 = df.rx0; sp.ry0 = df.ry0; sp.rx1 = df.rx1; sp.ry1 = df.ry1;
                        sp.anchor = PM3D_ANCHOR_BOTTOM; sp.focusable = 1; sp._pad[0] = sp._pad[1] = 0;
                        p.out_sprites[o] = sp;
                    }
                }
                off++;
            }
        }
    }
    if (tid == 0) p.out_count[view] = n_out;
}

int ensure_buf(pm3d_ctx *ctx, void **buf, size_t *have, size_t need) {
    if (*have >= need) return PM3D_OK;
    if (*buf) PM3D_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr; *have = 0;
    PM3D_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return PM3D_OK;
}

}  // namespace

extern "C" int pm3d_go_rand_cooked(uint64_t *out607) {
    if (!out607) return PM3D_ERR_ARG;
    memcpy(out607, go_rand_cooked(), RNG_LEN * sizeof(uint64_t));
    return PM3D_OK;
}

extern "C" int pm3d_upload_clutter(pm3d_ctx *ctx, const pm3d_clutter_def *defs, int n_defs, const uint8_t *cell_class,
                                   const uint64_t *class_mask, int n_classes, int64_t map_seed) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->grid) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_grid must be called first");
    if (n_defs < 0 || n_defs > PM3D_CLUTTER_MAX_DEFS) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "n_defs %d outside [0, %d]", n_defs, PM3D_CLUTTER_MAX_DEFS);
    if (n_classes < 1 || n_classes > 256 || !cell_class || !class_mask || (n_defs > 0 && !defs))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "clutter: null pointer or n_classes outside [1, 256]");
    const size_t cells = (size_t)ctx->map_w * ctx->map_h;
    for (size_t i = 0; i < cells; i++)
        if (cell_class[i] >= n_classes) return pm3d_fail(ctx, PM3D_ERR_ARG, "cell_class[%zu] = %d >= n_classes %d", i, (int)cell_class[i], n_classes);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_buf(ctx, (void **)&ctx->clutter_class, &ctx->clutter_class_bytes, cells))) return rc;
    if ((rc = ensure_buf(ctx, (void **)&ctx->clutter_tab, &ctx->clutter_tab_bytes,
                         256 * sizeof(uint64_t) + PM3D_CLUTTER_MAX_DEFS * sizeof(pm3d_clutter_def) + (MAX_DRAWS + 1) * sizeof(DrawEntry))))
        return rc;
    // one block: class masks | defs | draw table
    std::vector<uint8_t> blob(ctx->clutter_tab_bytes, 0);
    memcpy(blob.data(), class_mask, (size_t)n_classes * sizeof(uint64_t));
    if (n_defs) memcpy(blob.data() + 256 * sizeof(uint64_t), defs, (size_t)n_defs * sizeof(pm3d_clutter_def));
    {
        std::vector<uint32_t> pw(24 + 3 * RNG_LEN);
        pw[0] = 1;
        for (size_t n = 1; n < pw.size(); n++) pw[n] = lehmer_mul(pw[n - 1], LEHMER_A);
        const uint64_t *cooked = go_rand_cooked();
        DrawEntry *d = reinterpret_cast<DrawEntry *>(blob.data() + 256 * sizeof(uint64_t) + PM3D_CLUTTER_MAX_DEFS * sizeof(pm3d_clutter_def));
        for (int k = 1; k <= MAX_DRAWS; k++) {
            const int fi = (RNG_LEN - RNG_TAP) - k, ti = RNG_LEN - k;
            d[k].cooked_feed = cooked[fi];
            d[k].cooked_tap = cooked[ti];
            for (int j = 0; j < 3; j++) { d[k].pow_feed[j] = pw[21 + 3 * fi + j]; d[k].pow_tap[j] = pw[21 + 3 * ti + j]; }
        }
    }
    PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->clutter_class, cell_class, cells, cudaMemcpyHostToDevice, ctx->stream));
    PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->clutter_tab, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->clutter_defs = n_defs;
    ctx->clutter_seed = map_seed;
    ctx->clutter_ready = true;
    return PM3D_OK;
}

extern "C" int pm3d_clutter_batch(pm3d_ctx *ctx, const pm3d_clutter_query *queries, int n, double clutter_distance,
                                  int max_per_view, int32_t *out_count, pm3d_clutter_sprite *out_records,
                                  pm3d_sprite *out_sprites, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->clutter_ready) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_clutter must be called first");
    if (n < 0 || max_per_view < 1 || (n > 0 && (!queries || !out_count || !out_records)))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "clutter batch: null pointers or max_per_view < 1");
    if (!(clutter_distance <= PM3D_CLUTTER_MAX_DISTANCE))  // also rejects NaN
        return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "clutter_distance %g > %g cells", clutter_distance, PM3D_CLUTTER_MAX_DISTANCE);
    if (n == 0) return PM3D_OK;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc_ = pm3d_main_enter(ctx); if (rc_) return rc_; }
    const bool in_dev = flags & PM3D_IN_DEVICE, out_dev = flags & PM3D_OUT_DEVICE;
    const size_t per = (size_t)n * max_per_view;
    const size_t b_cnt = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255, b_rec = per * sizeof(pm3d_clutter_sprite),
                 b_spr = out_sprites ? per * sizeof(pm3d_sprite) : 0;
    int rc;
    const pm3d_clutter_query *d_q = queries;
    if (!in_dev) {
        if ((rc = ensure_buf(ctx, &ctx->io_a, &ctx->io_a_bytes, (size_t)n * sizeof(pm3d_clutter_query)))) return rc;
        PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->io_a, queries, (size_t)n * sizeof(pm3d_clutter_query), cudaMemcpyHostToDevice, ctx->stream));
        d_q = (const pm3d_clutter_query *)ctx->io_a;
    }
    int32_t *d_cnt = out_count;
    pm3d_clutter_sprite *d_rec = out_records;
    pm3d_sprite *d_spr = out_sprites;
    if (!out_dev) {
        if ((rc = ensure_buf(ctx, &ctx->io_c, &ctx->io_c_bytes, b_cnt + b_rec + b_spr))) return rc;
        d_cnt = (int32_t *)ctx->io_c;
        d_rec = (pm3d_clutter_sprite *)((uint8_t *)ctx->io_c + b_cnt);
        d_spr = out_sprites ? (pm3d_sprite *)((uint8_t *)ctx->io_c + b_cnt + b_rec) : nullptr;
    }
    // ClutterHandler.Update returns before doing anything when there is no clutter or no distance (clutter.go:31-34)
    if (ctx->clutter_defs == 0 || !(clutter_distance > 0)) {
        PM3D_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    } else {
        ClutterParams p;
        p.queries = d_q;
        p.level0 = ctx->grid;
        p.cell_class = ctx->clutter_class;
        p.class_mask = (const uint64_t *)ctx->clutter_tab;
        p.defs = (const pm3d_clutter_def *)((const uint8_t *)ctx->clutter_tab + 256 * sizeof(uint64_t));
        p.draws = (const DrawEntry *)((const uint8_t *)ctx->clutter_tab + 256 * sizeof(uint64_t) + PM3D_CLUTTER_MAX_DEFS * sizeof(pm3d_clutter_def));
        p.out_count = d_cnt; p.out_records = d_rec; p.out_sprites = d_spr;
        p.map_seed = ctx->clutter_seed;
        p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.n_defs = ctx->clutter_defs;
        p.n_dist = (int)clutter_distance + 1;  // for d := 0.0; d <= clutterDistance; d++
        p.max_per_view = max_per_view;
        const int max_samples = PM3D_CLUTTER_MAX_ANGLES * p.n_dist;
        int slots = 256;
        while (slots < 2 * max_samples) slots <<= 1;
        p.slots = slots;
        const size_t smem = ((size_t)3 * max_samples + slots) * sizeof(uint32_t);
        if (smem > 48 * 1024 && !ctx->clutter_smem_optin) {
            PM3D_CUDA(ctx, cudaFuncSetAttribute(k_clutter, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            ctx->clutter_smem_optin = true;
        }
        k_clutter<<<n, CLUTTER_THREADS, smem, ctx->stream>>>(p);
        ctx->launches++;
        PM3D_CUDA(ctx, cudaGetLastError());
    }
    if (!out_dev) {
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_count, d_cnt, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PM3D_CUDA(ctx, cudaMemcpyAsync(out_records, d_rec, b_rec, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_sprites) PM3D_CUDA(ctx, cudaMemcpyAsync(out_sprites, d_spr, b_spr, cudaMemcpyDeviceToHost, ctx->stream));
    }
    pm3d_main_touch(ctx);
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}
