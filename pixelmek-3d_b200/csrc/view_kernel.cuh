// view_kernel.cuh — per-view setup beside the wall cast and before the pixel pass: one CTA per view.
//
//  (a) row table: the floor cast's per-row distance and light tint (they depend on the
//      screen row only) and the sky / floor-box texel row;
//  (b) sprite projection (raycaster.Camera sprite cast; Sprite accessors
//      game/render/sprites/sprite.go:49-91), far->near ordering and SetScreenRect rectangles
//      (sprite.go:85); the convergence sprite needs the z-buffer and is picked by the compose
//      kernel's CTA that owns the screen centre, so this setup does not depend on the cast;
//  (c) per 64-column group lists of the sprites that can touch the group, in paint order.
#pragma once
#include "pm3d_internal.cuh"

#define VIEW_THREADS 256
#define MAX_VIEW_SPRITES 2048

struct ViewParams {
    const pm3d_view *views;
    RowRec *row_rec;
    const pm3d_sprite *sprites;
    const int32_t *sprite_offset;
    const SheetInfo *sheets;
    DrawRec *draw_rec;
    int32_t *n_draw;
    uint16_t *bin_list;
    int4 *bin_rect;    // dsx, dex, dsy, dey of each bin_list entry
    int4 *bin_count;   // per (view, group): x = entries, y = min dsy, z = max dey
    pm3d_rect *rects;
    pm3d_convergence *conv;
    uint32_t *status;     // device error words (pm3d_ctx::status_d)
    int n_sprites_total;  // what the caller declared for this batch: offsets must stay inside it
    int w, h, tex_size, n_sheets, n_groups, sky_tile, floorbox_tile;
};

__device__ __forceinline__ double sprite_distance(const pm3d_view &c, const pm3d_sprite &sp) {
    double dx = c.pos_x - sp.x, dy = c.pos_y - sp.y;
    return sqrt(dx * dx + dy * dy);
}

__device__ __forceinline__ double anchor_offset(int anchor, double scale, int64_t h) {
    double half = (double)h / 2.0;
    switch (anchor) {
    case PM3D_ANCHOR_BOTTOM: return half - scale * half;
    case PM3D_ANCHOR_CENTER: return half;
    case PM3D_ANCHOR_TOP: return half + scale * half;
    }
    return 0.0;
}

// ordering key: far -> near, ties by list index (DESIGN.md "sprite order")
struct SortKey { double d; int idx; };
__device__ __forceinline__ bool key_before(const SortKey &a, const SortKey &b) {
    // a paints before b: strictly farther, or not comparable/equal and lower index.
    // NaN distances never move relative to their neighbours in the CPU insertion sort; here
    // they sort as "equal to everything", which matches for the tested (finite) domain.
    if (a.d > b.d) return true;
    if (a.d < b.d) return false;
    return a.idx < b.idx;
}

// (a) row table of one view: rows [row0, row0 + VIEW_THREADS) — its own CTAs of the front launch, beside the cast and
// the sprite setup (it is a third of the sprite setup's serial time otherwise)
__device__ __forceinline__ void view_rows(const ViewParams &p, const int view, const int row0) {
    const pm3d_view c = p.views[view];
    const int64_t h = p.h;
    const int64_t horizon = go_int((double)h * 0.5) + c.pitch_px;
    // the two box spans are row-invariant: one 64-bit division each, not one per row
    const uint64_t sky_step = horizon > 0 ? span_step(p.tex_size, horizon) : 0ull;
    const uint64_t box_step = h > horizon ? span_step(p.tex_size, h - horizon) : 0ull;
    const int y = row0 + (int)threadIdx.x;
    if (y < p.h) {
        RowRec r;
        r.cd = ((double)h + 2.0 * c.cam_z_px) / (2.0 * (double)(y - c.pitch_px) - (double)h);
        r.tint = light_tint(c, r.cd, 255, 0.0, false);
        bool valid = !(c.render_distance >= 0 && !(r.cd <= c.render_distance));
        // sky / floor-box texel row (blit model): sky spans [0,horizon), floor box [horizon,h)
        uint32_t v;
        if (y < horizon) v = (uint32_t)((((uint64_t)(2 * (int64_t)y + 1)) * sky_step) >> 33);
        else v = (uint32_t)((((uint64_t)(2 * ((int64_t)y - horizon) + 1)) * box_step) >> 33);
        if (v > (uint32_t)p.tex_size - 1) v = p.tex_size - 1;
        // bit 29: the row distance is positive and mid-range: weight = cd / distWall may use the reciprocal paths
        r.info = (valid ? 0x80000000u : 0u) | (y < horizon ? 0x40000000u : 0u) | ((exp_mid(r.cd) && r.cd > 0.0) ? 0x20000000u : 0u) | v;
        p.row_rec[(size_t)view * p.h + y] = r;
    }
}

// One sprite of a view: screen rectangle (SetScreenRect) and, when it can touch the screen, its draw record.
// Returns whether it is on screen.  (raycaster.Camera sprite cast in Go integer semantics.)
__device__ __forceinline__ bool project_sprite(const ViewParams &p, const pm3d_view &c, const pm3d_sprite &sp, const double dist,
                                               const int src, pm3d_rect &rc, DrawRec &d) {
    const int64_t h = p.h, w = p.w;
    const double spriteX = sp.x - c.pos_x, spriteY = sp.y - c.pos_y;
    const double invDet = 1.0 / (c.plane_x * c.dir_y - c.dir_x * c.plane_y);
    const double transformX = invDet * (c.dir_y * spriteX - c.dir_x * spriteY);
    const double transformY = invDet * (-c.plane_y * spriteX + c.plane_x * spriteY);
    const int64_t tex_w = (int64_t)sp.rx1 - sp.rx0, tex_h = (int64_t)sp.ry1 - sp.ry0;
    const int64_t spriteScreenX = go_int((double)w / 2.0 * (1.0 + transformX / transformY));
    const double ratioWH = (double)tex_w / (double)tex_h;
    const double uDiv = 1.0 / (sp.scale * ratioWH);
    const double vDiv = 1.0 / sp.scale;
    const double vMove = -sp.z * (double)h + anchor_offset(sp.anchor, sp.scale, h);
    const int64_t vMoveScreen = wadd(wadd(go_int(vMove / transformY), c.pitch_px), go_int(c.cam_z_px / transformY));
    const int64_t spriteHeight = go_int(fabs((double)h / transformY) / vDiv);
    const int64_t dsy = wadd(wadd(wneg(spriteHeight) / 2, h / 2), vMoveScreen);
    const int64_t dey = wadd(wadd(spriteHeight / 2, h / 2), vMoveScreen);
    const int64_t spriteWidth = go_int(fabs((double)h / transformY) / uDiv);
    const int64_t dsx = wadd(wneg(spriteWidth) / 2, spriteScreenX);
    const int64_t dex = wadd(spriteWidth / 2, spriteScreenX);

    bool ok = transformY > 0 && spriteWidth > 0 && spriteHeight > 0 && dex > dsx && dey > dsy && tex_w > 0 &&
              tex_h > 0 && sp.sheet >= 0 && sp.sheet < p.n_sheets;
    if (c.render_distance >= 0 && !(dist <= c.render_distance)) ok = false;
    if (ok && (spriteWidth > (1ll << 20) || spriteHeight > (1ll << 20) || tex_w > 4096 || tex_h > 4096)) ok = false;

    if (ok) {
        rc.x0 = (int32_t)clamp_i64(dsx, -(1 << 30), 1 << 30);
        rc.y0 = (int32_t)clamp_i64(dsy, -(1 << 30), 1 << 30);
        rc.x1 = (int32_t)clamp_i64(dex, -(1 << 30), 1 << 30);
        rc.y1 = (int32_t)clamp_i64(dey, -(1 << 30), 1 << 30);
    } else {
        rc.x0 = 0; rc.y0 = 0; rc.x1 = -1; rc.y1 = -1;
    }
    // only sprites that can touch a drawable stripe (0 < stripe < w) and row are kept
    const bool on_screen = ok && dex > 1 && dsx < w && dey > 0 && dsy < h;
    if (on_screen) {
        d.dsx = (int32_t)clamp_i64(dsx, -(1 << 30), 1 << 30);
        d.dex = (int32_t)clamp_i64(dex, -(1 << 30), 1 << 30);
        d.dsy = (int32_t)clamp_i64(dsy, -(1 << 30), 1 << 30);
        d.dey = (int32_t)clamp_i64(dey, -(1 << 30), 1 << 30);
        d.transformY = transformY;
        const uint64_t ystep = span_step(tex_h, dey - dsy);
        d.ystep2 = ystep * 2ull;
        d.ypos0 = (1ull - 2ull * (uint64_t)dsy) * ystep;
        // m = floor(2^64 / sw) + 1 for sw >= 2 (exact floor(n/sw) for n < 2^32); sw == 1 has no stripes
        d.xmagic = spriteWidth >= 2 ? (0xFFFFFFFFFFFFFFFFull / (uint64_t)spriteWidth) + 1ull : 0ull;
        const SheetInfo si = p.sheets[sp.sheet];
        d.sheet_px = (uint64_t)si.px_offset + (uint64_t)((int64_t)sp.ry0 * si.w + sp.rx0);
        d.tex_w = (uint32_t)tex_w;
        d.tex_h = (uint32_t)tex_h;
        d.sheet_w = (uint32_t)si.w;
        d.tint = light_tint(c, transformY, 255, sp.illumination, true);
        // texel ranges inside the sheet: rx0 + tx in [0, si.w), ry0 + ty in [0, si.h)
        int64_t txl = sp.rx0 < 0 ? -(int64_t)sp.rx0 : 0, txh = (int64_t)si.w - sp.rx0;
        int64_t tyl = sp.ry0 < 0 ? -(int64_t)sp.ry0 : 0, tyh = (int64_t)si.h - sp.ry0;
        if (txh > tex_w) txh = tex_w;
        if (tyh > tex_h) tyh = tex_h;
        if (txh < txl) txh = txl;
        if (tyh < tyl) tyh = tyl;
        d.tx_lo = (uint16_t)txl; d.tx_hi = (uint16_t)txh; d.ty_lo = (uint16_t)tyl; d.ty_hi = (uint16_t)tyh;
        d.dist = dist;
        d.src = src;
        d.flags = sp.focusable ? 1u : 0u;
        d._pad0 = 0; d._pad1 = 0;
    }
    return on_screen;
}

__device__ __forceinline__ void view_setup(const ViewParams &p, const int view) {
    const int tid = threadIdx.x;
    const pm3d_view c = p.views[view];

#ifdef PM3D_FRONT_STATS
    const long long vt0 = clock64();
    auto vstamp = [&](int k) { __syncthreads(); if (tid == 0) g_front_stats[k] = (unsigned long long)(clock64() - vt0); };
#endif
    // ---------------- (b) sprites
    // pm3d_render validates the offsets on the host; pm3d_render_resident cannot (they live in HBM, written by
    // pm3d_raycast_sprites_batch or by the caller), so they are checked here: a view whose entries are negative, not
    // monotone, hold more sprites than the shared-memory sort takes or point past the declared total renders without
    // sprites and raises the ctx's error word (pm3d_sync / pm3d_download_results then return PM3D_ERR_CAPACITY).
    int s0 = p.sprite_offset[view];
    int n = p.sprite_offset[view + 1] - s0;
    if (s0 < 0 || n < 0 || n > MAX_VIEW_SPRITES || s0 > p.n_sprites_total - n) {
        if (tid == 0) {
            *reinterpret_cast<volatile uint32_t *>(p.status + PM3D_STATUS_VIEW_SPRITES) = 1u + (uint32_t)view;
            __threadfence_system();
        }
        s0 = 0; n = 0;
    }
    const pm3d_sprite *sprites = p.sprites + s0;
    DrawRec *draw = p.draw_rec + s0;

    __shared__ double key_d[MAX_VIEW_SPRITES];     // paint order: distance of the i-th painted sprite
    __shared__ uint16_t key_i[MAX_VIEW_SPRITES];   // ... and its index in the view's sprite list (0xffff = pad)
    __shared__ int2 xspan[MAX_VIEW_SPRITES];     // dsx, dex of the compacted draw list
    __shared__ int warp_tot[VIEW_THREADS / 32];
    __shared__ int s_ndraw, s_nan;
    int4 *srect = reinterpret_cast<int4 *>(xspan);  // fast path: dsx, dex, dsy, dey of the compacted draw list (<= VIEW_THREADS)
    const int lane = tid & 31, wid = tid >> 5;

    // ---- fast path, at most one sprite per thread (the game's per-frame lists: a few hundred sprites).  Nothing is
    // sorted: every thread projects ITS sprite first (the long part: a dozen binary64 divisions, independent of the
    // order), then finds its place in the paint order by counting the on-screen sprites that paint before it
    // (farther, or as far with a lower list index) — one pass over the distances in shared memory instead of the
    // 36 barrier-separated passes of a 256-key bitonic sort.  A NaN distance (not a total order) takes the sort below.
    bool fast = n <= VIEW_THREADS;
    if (fast) {
        if (tid == 0) s_nan = 0;
        __syncthreads();
        DrawRec d;
        bool on = false;
        double dist = 0.0;
        if (tid < n) {
            const pm3d_sprite sp = sprites[tid];
            dist = sprite_distance(c, sp);
            pm3d_rect rc;
            on = project_sprite(p, c, sp, dist, tid, rc, d);
            p.rects[s0 + tid] = rc;
            if (dist != dist) s_nan = 1;
            key_d[tid] = on ? dist : -INFINITY;  // (off-screen sprites paint before nobody)
        }
        __syncthreads();
        fast = s_nan == 0;
        if (fast) {
            int pos = 0, tot = 0;
            for (int j = 0; j < n; j++) {
                const double dj = key_d[j];
                tot += dj != -INFINITY;
                pos += (dj > dist) || (dj == dist && j < tid);
            }
            if (on) {
                draw[pos] = d;
                srect[pos] = make_int4(d.dsx, d.dex, d.dsy, d.dey);
            }
            if (tid == 0) { s_ndraw = tot; p.n_draw[view] = tot; }
        }
        __syncthreads();
    }
    if (!fast) {
    int npad = 1;
    while (npad < n) npad <<= 1;

    if (tid == 0) s_nan = 0;
    __syncthreads();
    for (int i = tid; i < npad; i += VIEW_THREADS) {
        if (i >= n) { key_d[i] = -1.0; key_i[i] = 0xffff; continue; }  // pads paint last
        const double d = sprite_distance(c, sprites[i]);
        key_d[i] = d;
        key_i[i] = (uint16_t)i;
        if (d != d) s_nan = 1;
    }
    __syncthreads();
    // up to 1024 sprites with totally ordered (NaN-free) distances: rank sort — every sprite counts the sprites that
    // paint before it and is scattered to that place — n/256 passes over the distances instead of the ~50 barrier-
    // separated passes of the bitonic network (408 sprites, the bundled map's static list: 25 k -> 9 k cycles)
    const bool rank_sort = n <= 1024 && s_nan == 0;
    if (rank_sort) {
        // distances of the unsorted list move to the upper half of key_d (n <= 1024 of 2048 entries), the sorted
        // (distance, index) pairs are written to the lower half
        for (int i = tid; i < n; i += VIEW_THREADS) key_d[1024 + i] = key_d[i];
        __syncthreads();
        for (int i = tid; i < n; i += VIEW_THREADS) {
            const double di = key_d[1024 + i];
            int rank = 0;
            for (int j = 0; j < n; j++) {
                const double dj = key_d[1024 + j];
                rank += (dj > di) || (dj == di && j < i);
            }
            key_d[rank] = di;
            key_i[rank] = (uint16_t)i;
        }
        __syncthreads();
    }
    for (int k = 2; k <= npad && !rank_sort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += VIEW_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    SortKey a, b;
                    a.d = key_d[i]; a.idx = key_i[i]; b.d = key_d[ixj]; b.idx = key_i[ixj];
                    bool up = (i & k) == 0;
                    bool swap = up ? key_before(b, a) : key_before(a, b);
                    if (swap) { key_d[i] = b.d; key_i[i] = (uint16_t)b.idx; key_d[ixj] = a.d; key_i[ixj] = (uint16_t)a.idx; }
                }
            }
            __syncthreads();
        }
    }

    // ordered compaction of the on-screen sprites: chunk of VIEW_THREADS paint-order entries at a time
    int base_count = 0;
    for (int base = 0; base < n; base += VIEW_THREADS) {
        const int i = base + tid;
        DrawRec d;
        bool on_screen = false;
        if (i < n) {
            const int src = key_i[i];
            pm3d_rect rc;
            on_screen = project_sprite(p, c, sprites[src], key_d[i], src, rc, d);
            p.rects[s0 + src] = rc;
        }
        // block-wide ordered position of this entry among the on-screen ones
        const unsigned bal = __ballot_sync(0xffffffffu, on_screen);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = base_count;
        for (int wv = 0; wv < wid; wv++) before += warp_tot[wv];
        int chunk_total = 0;
        for (int wv = 0; wv < VIEW_THREADS / 32; wv++) chunk_total += warp_tot[wv];
        if (on_screen) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            draw[pos] = d;
            xspan[pos] = make_int2(d.dsx, d.dex);
        }
        base_count += chunk_total;
        __syncthreads();
    }
    if (tid == 0) { s_ndraw = base_count; p.n_draw[view] = base_count; }

    __syncthreads();
    }
#ifdef PM3D_FRONT_STATS
    vstamp(6);  // projection + ordering done
#endif

    // ---------------- (c) column-group lists, paint order preserved: one warp per group
    const int m = s_ndraw;
    for (int g = wid; g < p.n_groups; g += VIEW_THREADS / 32) {
        const int gx0 = g * GROUP_COLS, gx1 = gx0 + GROUP_COLS;
        const size_t lbase = (size_t)s0 * p.n_groups + (size_t)g * n;
        uint16_t *list = p.bin_list + lbase;
        int4 *rects = p.bin_rect + lbase;
        int cnt = 0, ymin = 0x7fffffff, ymax = -0x7fffffff;
        for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            bool keep = false;
            int4 rc = make_int4(0, 0, 0, 0);
            if (i < m) {
                if (fast) { rc = srect[i]; keep = rc.y > gx0 && rc.x < gx1; }
                else { const int2 xs = xspan[i]; keep = xs.y > gx0 && xs.x < gx1; }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                if (!fast) rc = *reinterpret_cast<const int4 *>(&draw[i]);  // dsx, dex, dsy, dey
                list[pos] = (uint16_t)i;
                rects[pos] = rc;
                ymin = min(ymin, rc.z);
                ymax = max(ymax, rc.w);
            }
            cnt += __popc(bal);
        }
        for (int o = 16; o > 0; o >>= 1) {
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        if (lane == 0) p.bin_count[view * p.n_groups + g] = make_int4(cnt, ymin, ymax, 0);
    }
}

