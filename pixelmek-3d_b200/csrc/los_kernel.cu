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
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "pm3d_internal.cuh"

struct LosParams {
    const double4 *rays;  // x1,y1,x2,y2
    const double4 *segs;
    const int4 *seg_irect;  // image.Rect(int(x1),int(y1),int(x2),int(y2)) canonicalised: minx,miny,maxx,maxy
    const int32_t *cell_start;
    const int32_t *cell_items;
    const int4 *cell_rec;  // per bin cell: x = count, y/z/w = the first three candidate ids (count <= 3) or y = offset into cell_items
    int64_t n;
    int n_segs, bin_x0, bin_y0, bin_w, bin_h, max_skip;
    // move queries
    const uint16_t *grid0;
    int map_w, map_h;
    double clip;
    uint8_t *out_flag;
    double2 *out_pos;
    // two-pass line of sight (k_los_fast -> k_los): per-cell code bytes, the list of rays the first pass left
    // undecided and its length (device memory; k_los reads its work from it when list != nullptr)
    const uint8_t *cell_code;
    uint32_t *list;
    unsigned *list_count;
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

// Same boolean as line_intersection(..., nullptr, nullptr) without the two divisions.
// With denom, tNum, uNum finite and their exponents within 2^+-500 the quotients cannot overflow or
// underflow, so for t = RN(tNum/denom):  t < 0  <=>  tNum != 0 and the signs differ;
// t > 1  <=>  |tNum| > |denom|  (two doubles a > b satisfy a/b >= 1 + ulp(b)/b > 1 + 2^-53, which rounds to at
// least 1 + 2^-52; a <= b gives a quotient <= 1).  Anything else (zero / tiny / huge / inf / NaN operands,
// where Go's comparisons with NaN all read false) takes the reference arithmetic.
__device__ __forceinline__ bool segments_intersect(const double4 &a, const double4 &b) {
    const double x1 = a.x, y1 = a.y, x2 = a.z, y2 = a.w;
    const double x3 = b.x, y3 = b.y, x4 = b.z, y4 = b.w;
    const double denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    const double tNum = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4);
    const double uNum = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3));
    if (exp_mid(denom) && (tNum == 0 || exp_mid(tNum)) && (uNum == 0 || exp_mid(uNum))) {
        // (the guard keeps every exponent within 2^+-500, so the products below neither overflow nor underflow:
        //  a product is negative exactly when the numerator is non-zero and its sign differs from denom's)
        const double ad = fabs(denom);
        if (tNum * denom < 0.0 || fabs(tNum) > ad) return false;
        if (uNum * denom < 0.0 || fabs(uNum) > ad) return false;
        return true;
    }
    if (denom == 0) return false;
    const double t = tNum / denom;
    if (t > 1 || t < 0) return false;
    const double u = uNum / denom;
    if (u > 1 || u < 0) return false;
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
        return segments_intersect(ray, sg);
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

// Persistent line-of-sight kernel.  Rays end after very different numbers of cells (most are blocked
// within a few cells, the visible ones cross the whole map), so a warp owns a contiguous range of rays
// and a lane that finishes its ray takes the next one of the range (refill once >= 8 lanes are idle)
// instead of idling until the slowest ray of a fixed group of 32 is done.  One loop iteration = one bin cell.
#define LOS_THREADS 256
#ifndef LOS_LANE_CAP
#define LOS_LANE_CAP 8
#endif
#ifndef LOS_MAX_SPLIT_SHIFT
#define LOS_MAX_SPLIT_SHIFT 6  // the exact second pass cuts a ray into at most 2^6 parts (see k_los)
#endif
#ifndef LOS_SPLIT_ROUNDS
#define LOS_SPLIT_ROUNDS 4     // ... as long as the parts are no more than this many per lane of the grid
#endif
#ifndef LOS_MIN_BLOCKS
#define LOS_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(LOS_THREADS, LOS_MIN_BLOCKS) k_los(LosParams p, long long rays_per_warp) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = (long long)blockIdx.x * (LOS_THREADS / 32) + (threadIdx.x >> 5);
    long long n_work = p.n;
    // Second pass (p.list): the rays k_los_fast could not decide; their number is only known on the device.  A ray's
    // walk is a chain of a few hundred dependent cell steps (~130 us for a ray across the map, however few rays are
    // left), and "blocked" is an OR over its cells — so when the list is shorter than the grid is wide, every ray is
    // cut into 2^kshift PARTS at the parameters k / 2^kshift, each part walks the cells of its own sub-segment and runs
    // the exact predicate on the WHOLE ray.  The parts of a ray meet at bitwise identical points, a sub-segment's
    // cells are those of the ray up to the rounding of its end points (1e-13 against the bins' 1e-3 dilation,
    // pm3d_upload_segments), and the answer byte holds "visible" from the first pass: a part that finds a hit clears it.
    int kshift = 0;
    if (p.list) {
        const long long n_rays = (long long)*p.list_count;
        const long long warps = (long long)gridDim.x * (LOS_THREADS / 32);
        while (kshift < LOS_MAX_SPLIT_SHIFT && (n_rays << (kshift + 1)) <= warps * 32 * LOS_SPLIT_ROUNDS) kshift++;
        n_work = n_rays << kshift;
        rays_per_warp = (n_work + warps - 1) / warps;
        if (rays_per_warp < 32) rays_per_warp = 32;
    }
    long long next = warp_id * rays_per_warp;                      // warp-uniform: next unassigned ray
    const long long end = min(next + rays_per_warp, n_work);
    if (next >= end) return;

    const double gx0 = (double)p.bin_x0, gy0 = (double)p.bin_y0;
    const double gx1 = (double)(p.bin_x0 + p.bin_w), gy1 = (double)(p.bin_y0 + p.bin_h);

    __shared__ double4 s_ray[LOS_THREADS];
    __shared__ int4 s_lr[LOS_THREADS];
    __shared__ uint32_t s_q[LOS_THREADS / 32][32 * LOS_LANE_CAP];
    __shared__ unsigned s_hit[LOS_THREADS / 32];
    double4 *wray = s_ray + (threadIdx.x & ~31);
    int4 *wlr = s_lr + (threadIdx.x & ~31);
    uint32_t *wq = s_q[threadIdx.x >> 5];
    unsigned *whit = &s_hit[threadIdx.x >> 5];
    if (lane == 0) *whit = 0u;
    __syncwarp();

    bool active = false;
    int cbase = 0;  // candidates of the current cell already tested
    long long idx = 0;
    double4 ray = make_double4(0, 0, 0, 0);
    int4 lr = make_int4(0, 0, 0, 0);
    int cx = 0, cy = 0, ex = 0, ey = 0, sx = 1, sy = 1, steps = 0;
    int4 cur = make_int4(0, -1, -1, -1);  // record of the current cell
    double tx = 0, ty = 0, inv_dx = 0, inv_dy = 0;

    auto test = [&](int id) -> bool {
        const int4 br = __ldg(p.seg_irect + id);  // lineOfSight's integer bounding-rect pre-filter, as written
        if ((br.x > lr.z && br.y > lr.w) || (br.z < lr.x && br.w < lr.y)) return false;
        const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
        const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
        return segments_intersect(ray, make_double4(s01.x, s01.y, s23.x, s23.y));
    };

    for (;;) {
        // ---- refill idle lanes from the warp's range
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && next < end && (__popc(idle) >= 8 || idle == 0xffffffffu)) {
            const long long mine = next + __popc(idle & ((1u << lane) - 1u));
            next += __popc(idle);
            if (!active && mine < end) {
                idx = p.list ? (long long)p.list[mine >> kshift] : mine;
                const int part = (int)(mine & ((1ll << kshift) - 1)), n_parts = 1 << kshift;
                ray = p.rays[idx];
                lr = go_rect(go_int(ray.x), go_int(ray.y), go_int(ray.z), go_int(ray.w));
                const bool inside = ray.x >= gx0 && ray.x < gx1 && ray.z >= gx0 && ray.z < gx1 && ray.y >= gy0 &&
                                    ray.y < gy1 && ray.w >= gy0 && ray.w < gy1;
                if (!inside) {
                    // not finite, or leaves the binned area: every segment is a candidate (the first part takes them all)
                    bool hit = false;
                    for (int id = 0; id < p.n_segs && !hit && part == 0; id++) hit = test(id);
                    if (hit) p.out_flag[idx] = 0;
                    else if (!p.list) p.out_flag[idx] = 1;
                } else {
                    // this part's sub-segment (the whole ray when it is not split; the ray's own end points at its ends)
                    const double t0 = (double)part / (double)n_parts, t1 = (double)(part + 1) / (double)n_parts;
                    const double ax = part == 0 ? ray.x : ray.x + t0 * (ray.z - ray.x), ay = part == 0 ? ray.y : ray.y + t0 * (ray.w - ray.y);
                    const double bx = part == n_parts - 1 ? ray.z : ray.x + t1 * (ray.z - ray.x), by = part == n_parts - 1 ? ray.w : ray.y + t1 * (ray.w - ray.y);
                    const double fx = floor(ax), fy = floor(ay);
                    cx = (int)fx - p.bin_x0; cy = (int)fy - p.bin_y0;
                    ex = (int)floor(bx) - p.bin_x0; ey = (int)floor(by) - p.bin_y0;
                    const double dx = bx - ax, dy = by - ay;
                    sx = dx > 0 ? 1 : -1; sy = dy > 0 ? 1 : -1;
                    inv_dx = dx != 0 ? 1.0 / fabs(dx) : 0.0;
                    inv_dy = dy != 0 ? 1.0 / fabs(dy) : 0.0;
                    tx = dx != 0 ? (dx > 0 ? (fx + 1.0 - ax) : (ax - fx)) * inv_dx : 1e300;
                    ty = dy != 0 ? (dy > 0 ? (fy + 1.0 - ay) : (ay - fy)) * inv_dy : 1e300;
                    steps = abs(ex - cx) + abs(ey - cy);
                    cur = __ldg(p.cell_rec + cy * p.bin_w + cx);
                    cbase = 0;
                    wray[lane] = ray;
                    wlr[lane] = lr;
                    active = true;
                }
            }
            __syncwarp();
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
            if (next >= end) break;
            continue;
        }
        // ---- one bin cell of every active ray.  Cells without candidates are walked through first (a few cheap
        //      instructions each) so the lanes reach the test phase together with work to do.
        if (active) {
#pragma unroll 1
            for (int skip = 0; skip < p.max_skip && cur.x == 0 && steps > 0; skip++) {
                if ((tx < ty && cx != ex) || cy == ey) { tx += inv_dx; cx += sx; }
                else { ty += inv_dy; cy += sy; }
                steps--;
                cur = __ldg(p.cell_rec + cy * p.bin_w + cx);
                cbase = 0;
            }
        }
        // The candidates of the lanes' current cells differ in number (0 .. a dozen), so a lane testing only its
        // own would leave most of the warp idle.  They are pooled instead: every lane queues up to LOS_LANE_CAP
        // (ray, segment) pairs in shared memory, then the warp's 32 lanes take the pairs round-robin and run the
        // reference's exact predicate on whichever ray the pair belongs to (rays are staged in shared memory at
        // refill).  A hit is reported to the owning lane through a bit mask.
        const bool more = active && steps > 0;
        int ncx = cx, ncy = cy;
        double ntx = tx, nty = ty;
        int4 nrec = make_int4(0, -1, -1, -1);
        if (more) {
            // never step past the target cell on either axis: keeps the walk inside the bins
            if ((tx < ty && cx != ex) || cy == ey) { ntx += inv_dx; ncx += sx; }
            else { nty += inv_dy; ncy += sy; }
            nrec = __ldg(p.cell_rec + ncy * p.bin_w + ncx);  // requested before the tests, used after them
        }
        const int left = active ? cur.x - cbase : 0;
        const int n_put = left < LOS_LANE_CAP ? left : LOS_LANE_CAP;
        int off = n_put;  // inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, off, o);
            if (lane >= o) off += v;
        }
        const int total = __shfl_sync(0xffffffffu, off, 31);
        off -= n_put;
        if (n_put > 0) {
            const uint32_t tag = (uint32_t)lane << 27;
            if (cur.x <= 3) {
                // the usual cell: its (<= 3) candidate ids sit in the record itself and go out in one round
                // (cbase == 0): three predicated stores instead of a loop the whole warp would walk
                wq[off] = tag | (uint32_t)cur.y;
                if (n_put > 1) wq[off + 1] = tag | (uint32_t)cur.z;
                if (n_put > 2) wq[off + 2] = tag | (uint32_t)cur.w;
            } else {
                const int32_t *items = p.cell_items + cur.y + cbase;
                for (int k = 0; k < n_put; k++) wq[off + k] = tag | (uint32_t)__ldg(items + k);
            }
        }
        __syncwarp();
        for (int q = lane; q < total; q += 32) {
            const uint32_t e = wq[q];
            const int owner = (int)(e >> 27), id = (int)(e & 0x7ffffffu);
            const int4 br = __ldg(p.seg_irect + id);  // lineOfSight's integer bounding-rect pre-filter, as written
            const int4 olr = wlr[owner];
            if ((br.x > olr.z && br.y > olr.w) || (br.z < olr.x && br.w < olr.y)) continue;
            const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
            const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
            if (segments_intersect(wray[owner], make_double4(s01.x, s01.y, s23.x, s23.y))) atomicOr(whit, 1u << owner);
        }
        __syncwarp();
        const unsigned hits = *whit;
        __syncwarp();
        if (lane == 0) *whit = 0u;
        if (active) {
            const bool hit = (hits >> lane) & 1u;
            if (hit) {
                p.out_flag[idx] = 0;
                active = false;
            } else if (left > n_put) {
                cbase += n_put;  // more candidates in this cell than one round takes
            } else if (!more) {
                if (!p.list) p.out_flag[idx] = 1;  // (second pass: the first pass wrote "visible unless a part finds a hit")
                active = false;
            } else {
                steps--;
                cx = ncx; cy = ncy; tx = ntx; ty = nty;
                cur = nrec;
                cbase = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// First pass of pm3d_los_batch: decide a ray WITHOUT touching a collision segment whenever that is provable.
//
// pm3d_upload_segments derives one code byte per bin cell from the segment list:
//   bit 0      the cell has candidate segments (cell_rec.x > 0; bins are dilated by BIN_EPS)
//   bits 1..4  the cell's W / E / S / N unit edge (x = cx, x = cx + 1, y = cy, y = cy + 1 between the two lattice
//              points) is COVERED: some axis-parallel segment lies on that lattice line and contains the whole edge.
// The pass walks the cells of the ray (same Amanatides-Woo order as k_los, in the division-free cross-multiplied
// form: tX = ax*|dy| against tY = ay*|dx|) and looks only at these bytes (staged in shared memory when the bin
// area fits, north_star's "map grid staged in shared memory"):
//
// BLOCKED proof.  The walk is about to leave a cell through a covered edge, say the edge on x = c between
// (c, k) and (c, k+1), covered by the segment S = (c, a)-(c, b) with [k, k+1] inside [min(a,b), max(a,b)].
// c lies between x1 and x2 (x1 < c <= x2 walking right, x2 < c <= x1 walking left: the walk only steps in x
// while the target column is not reached).  The test `lo < P < hi` below proves that the TRUE ray crosses x = c at
// yc with k + m/2 < yc < k + 1 - m/2 (m = 2^-20; P, lo, hi carry relative errors <= 4u, u = 2^-53, coordinates
// are below 2^15, so the slack it can lose is < 2^-36).  Then geom.LineIntersection(ray, S) as the reference
// evaluates it in binary64 returns true:
//   * x3 - x4 = 0 exactly, so denom = RN(RN(x1-x2)*L), tNum = RN(RN(x1-c)*L) with the same L = RN(a-b);
//     |x1-c| <= |x1-x2| and RN is monotone, so |tNum| <= |denom|, both have the sign of L*(x1-x2) (or tNum = 0):
//     0 <= t = RN(tNum/denom) <= 1;
//   * u_true = (yc - a)/(b - a) lies in [m/2|L|, 1 - m/2|L|]; the computed uNum carries an absolute error
//     <= 3.1u(|x1-x2||y1-a| + |y1-y2||x1-c|) + u|uNum|, i.e. after the division an error <= (6.2u*2^16 + 3u|L|)/|L|
//     < 2^-33/|L|, far inside the margin: 0 <= u <= 1.
//   and lineOfSight's integer bounding-rect pre-filter cannot skip S (int() is monotone: int(c) lies between the
//   truncated ray endpoints in x), so the reference's loop returns false: not visible.  (Mirror image for edges on
//   y = c.)  |dx| >= 1e-90 keeps denom normal and makes an underflowing product an absolute error far below the margin.
// VISIBLE proof.  No visited cell has a candidate: no segment can intersect (the k_los argument).
// VISIBLE proof, verify mode (VERIFY, chosen at upload for maps where few cells have candidates, i.e. where most rays are
// visible and long).  A cell whose candidates are all CLEAN — axis-parallel, on a lattice line, integer end points,
// the collision outlines of wall cells — can be crossed without testing them: such a segment is a union of whole unit
// edges, so the ray intersects it exactly when it crosses one of its unit edges (covered: blocked, above) or touches a
// lattice point.  So every crossing is kept 2^-14 away from the lattice points of the line it crosses (one
// subtraction per step in the walk's own units, see the loop), the ray's end points must lie 2^-20 off every lattice line, and a cell
// holding a segment that is not clean (the border rectangles inset by clipDistance, arbitrary caller segments) is
// DIRTY: any of these failing makes the ray undecided.  With them holding, every clean segment misses the ray by a
// margin the reference's binary64 evaluation cannot bridge: on its lattice line x = c either c lies 2^-20 outside
// [x1, x2] (t = (x1-c)/(x1-x2) is outside [0, 1] by more than 2^-35, against 2^-52 of rounding) or the crossing is
// 2^-21 away from the segment's extent (u outside [0, 1] by 2^-21/|L| against an evaluation error < 2^-33/|L|).
// Everything else — a failed margin (the ray grazes a lattice corner), candidates seen but the target reached, a ray
// outside the binned area — is appended to the undecided list and decided by k_los with the exact predicate.
#define LOSF_THREADS 512
#define LOSF_MARGIN 9.5367431640625e-07  // 2^-20
#define LOSF_CAND 1u
#define LOSF_W 2u
#define LOSF_E 4u
#define LOSF_S 8u
#define LOSF_N 16u
#define LOSF_DIRTY 32u

// A lane that stops walking only clears `active`; the margin test, the answer byte and the append to the undecided
// list run later for all stopped lanes of the warp together, in the same convergent block that hands them their
// next rays (a lane finishing alone ran that tail at 1.7 of 32 lanes).  Why it stopped is read off its registers:
// no steps left (n2 == 0) = it reached the target cell, otherwise it stands in front of a covered edge.
#ifndef LOSF_STEPS
#define LOSF_STEPS 8     // cells per lane between two looks at the warp's idle lanes
#endif
#ifndef LOSF_REFILL
#define LOSF_REFILL 12   // idle lanes that trigger a refill
#endif

template <bool SMEM, bool VERIFY>
__global__ void __launch_bounds__(LOSF_THREADS, 2) k_los_fast(LosParams p, long long rays_per_warp) {
    extern __shared__ uint8_t s_code[];
    const int n_cells = p.bin_w * p.bin_h;
    if (SMEM) {
        const int words = (n_cells + 3) >> 2;  // the code array is padded to a multiple of 16 bytes
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.cell_code);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_code);
        for (int i = threadIdx.x; i < words; i += LOSF_THREADS) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const uint8_t *code = SMEM ? s_code : p.cell_code;
    auto code_at = [&](int cell) -> unsigned { return SMEM ? (unsigned)code[cell] : (unsigned)__ldg(code + cell); };

    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const long long warp_id = (long long)blockIdx.x * (LOSF_THREADS / 32) + (threadIdx.x >> 5);
    const long long begin = warp_id * rays_per_warp;
    const long long end = min(begin + rays_per_warp, (long long)p.n);
    if (begin >= end) return;

    // The warp's rays are fetched 32 at a time, one per lane (a coalesced 1 KB load issued when the previous
    // 32 are used up, so it is in flight while the lanes walk); a refill hands them out by shuffle.
    long long pf_base = begin;            // ray number held by lane 0's prefetch registers
    int pf_used = 0;                      // how many of the 32 have been handed out
    double4 pf = make_double4(0, 0, 0, 0);
    if (pf_base + lane < end) pf = p.rays[pf_base + lane];

    bool active = false;
    long long idx = -1;                   // < 0: the lane holds no ray to finalize
    double x1 = 0, y1 = 0, dx = 0, dy = 0, tX = 0, tY = 0;
    unsigned n2 = 0;                      // steps left: x steps in the low half, y steps in the high half
    int cell = 0, exy = 0;                // target cell (x | y << 16)
    int dcx = 0, dcy = 0;                 // cell index increments of an x / y step
    unsigned xmask = 0, ymask = 0, cur = 0, seen = 0;
    bool risky = false;                   // verify mode: something on the way could not be verified -> undecided
    double mthr = 0.0;                    // verify mode: 2^-14 * max(|dx|, |dy|)

    // margin test of the crossing of the lattice line the walk meets NEXT (x step: xs): true when the real crossing lies
    // 2^-21 inside the unit edge between the two cells
    auto crossing_inside = [&](bool xs, int nx, int ny) -> bool {
        const int cx = (exy & 0xffff) - (dcx > 0 ? nx : -nx), cy = (exy >> 16) - (dcy > 0 ? ny : -ny);
        const int fwd = xs ? (dcx > 0) : (dcy > 0);
        const double c = (double)((xs ? cx + p.bin_x0 : cy + p.bin_y0) + fwd);  // the lattice line crossed
        const double k = (double)(xs ? cy + p.bin_y0 : cx + p.bin_x0);          // the edge spans [k, k+1] on it
        const double a0 = xs ? x1 : y1, b0 = xs ? y1 : x1;
        const double da = xs ? dx : dy, db = xs ? dy : dx;
        const double ada = fabs(da);
        const double P = fabs(c - a0) * db;
        const double lo = ((k + LOSF_MARGIN) - b0) * ada, hi = ((k + (1.0 - LOSF_MARGIN)) - b0) * ada;
        return lo < P && P < hi && ada >= 1e-90;
    };

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        const bool more = pf_base + pf_used < end;
        if (idle == 0xffffffffu || (__popc(idle) >= LOSF_REFILL && more)) {
            // ---- finalize the rays of the stopped lanes (together)
            int result = -1;  // 0 blocked, 1 visible, 2 undecided
            if (!active && idx >= 0) {
                if (risky) {
                    result = 2;
                } else if (n2 == 0u) {
                    // target reached (or the ray left the bins: seen = CAND): visible when no cell had a candidate, or,
                    // in verify mode, when every crossing next to one was verified
                    result = (!(seen & LOSF_CAND) || (VERIFY && !(seen & LOSF_DIRTY))) ? 1 : 2;
                } else {
                    // in front of a covered edge: prove that the crossing lies inside it, margin 2^-20 from its ends
                    const int nx = (int)(n2 & 0xffffu), ny = (int)(n2 >> 16);
                    const bool xs = (tX < tY) ? (nx > 0) : (ny == 0);
                    result = crossing_inside(xs, nx, ny) ? 0 : 2;
                }
                p.out_flag[idx] = result == 0 ? 0 : 1;  // undecided (2): visible unless the exact pass finds a hit
            }
            const unsigned und = __ballot_sync(0xffffffffu, result == 2);
            if (und) {
                const int leader = __ffs(und) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(p.list_count, (unsigned)__popc(und));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (result == 2) p.list[base + __popc(und & below)] = (uint32_t)idx;
            }
            if (!active) idx = -1;
            // ---- hand out the next rays
            const int avail = (int)min((long long)(32 - pf_used), end - pf_base - pf_used);
            const int rank = __popc(idle & below);
            const int src = (pf_used + rank) & 31;
            double4 ray;
            ray.x = __shfl_sync(0xffffffffu, pf.x, src); ray.y = __shfl_sync(0xffffffffu, pf.y, src);
            ray.z = __shfl_sync(0xffffffffu, pf.z, src); ray.w = __shfl_sync(0xffffffffu, pf.w, src);
            const bool take = !active && rank < avail;
            if (take) idx = pf_base + pf_used + rank;
            pf_used += min(__popc(idle), avail);
            if (pf_used == 32) {
                pf_base += 32; pf_used = 0;
                if (pf_base + lane < end) pf = p.rays[pf_base + lane];
            }
            if (take) {
                // finite and small enough for the int conversions; then inside the bins as cell numbers
                const bool sane = fabs(ray.x) + fabs(ray.y) + fabs(ray.z) + fabs(ray.w) < 1e9;
                x1 = ray.x; y1 = ray.y;
                const double fx = floor(x1), fy = floor(y1);
                const int cx = sane ? (int)fx - p.bin_x0 : -1, cy = (int)fy - p.bin_y0;
                const int ex = (int)floor(ray.z) - p.bin_x0, ey = (int)floor(ray.w) - p.bin_y0;
                if (!((unsigned)cx < (unsigned)p.bin_w && (unsigned)ex < (unsigned)p.bin_w &&
                      (unsigned)cy < (unsigned)p.bin_h && (unsigned)ey < (unsigned)p.bin_h)) {
                    n2 = 0u; seen = LOSF_CAND | LOSF_DIRTY;  // not finite, or outside the binned area: undecided
                    risky = false;
                } else {
                    dx = ray.z - x1; dy = ray.w - y1;
                    const bool right = dx > 0, up = dy > 0;
                    n2 = (unsigned)abs(ex - cx) | ((unsigned)abs(ey - cy) << 16);
                    exy = ex | (ey << 16);
                    dcx = right ? 1 : -1; dcy = up ? p.bin_w : -p.bin_w;
                    xmask = right ? LOSF_E : LOSF_W; ymask = up ? LOSF_N : LOSF_S;
                    // distances to the first vertical / horizontal cell border, cross-multiplied (no division)
                    tX = (right ? (fx + 1.0 - x1) : (x1 - fx)) * fabs(dy);
                    tY = (up ? (fy + 1.0 - y1) : (y1 - fy)) * fabs(dx);
                    cell = cy * p.bin_w + cx;
                    cur = code_at(cell);
                    seen = cur;
                    if (VERIFY) {
                        // both end points 2^-20 off every lattice line (a point ON a line may touch a segment without crossing)
                        const double fzx = ray.z - floor(ray.z), fzy = ray.w - floor(ray.w), fxx = x1 - fx, fyy = y1 - fy;
                        const double m_ = LOSF_MARGIN, M_ = 1.0 - LOSF_MARGIN;
                        risky = !(fxx > m_ && fxx < M_ && fyy > m_ && fyy < M_ && fzx > m_ && fzx < M_ && fzy > m_ && fzy < M_) ||
                                (cur & LOSF_DIRTY) != 0u;
                        mthr = 6.103515625e-05 * fmax(fabs(dx), fabs(dy));
                    } else {
                        risky = false;
                    }
                    active = n2 != 0u && !(VERIFY && risky);
                }
            }
            if (idle == 0xffffffffu && __ballot_sync(0xffffffffu, idx >= 0) == 0u) break;
            continue;
        }
        // ---- a few cells of every walking lane
#pragma unroll
        for (int it = 0; it < LOSF_STEPS; it++) {
            if (active) {
                const bool xs = (tX < tY) ? ((n2 & 0xffffu) != 0u) : (n2 < 0x10000u);
                if (cur & (xs ? xmask : ymask)) {
                    active = false;
                } else {
                    // verify mode: this crossing and the next crossing of the OTHER axis must be 2^-14 apart along the
                    // crossed line, i.e. the ray keeps clear of the lattice point between them.  In the walk's own units
                    // that is one subtraction: the x line is met at tX/(|dx||dy|), the y line at tY/(|dx||dy|), and the ray
                    // moves |tX - tY|/|dx| along the x line (|tX - tY|/|dy| along the y line) in between.  Every pair of
                    // crossings adjacent in the merged order is a pair of heads at some step, so testing the heads at
                    // every step covers them all (tX, tY carry < 2^-16 of accumulated rounding on bin areas up to 4096^2).
                    if (VERIFY && !(fabs(tX - tY) > mthr)) risky = true;
                    if (xs) tX += fabs(dy); else tY += fabs(dx);
                    n2 -= xs ? 1u : 0x10000u;
                    cell += xs ? dcx : dcy;
                    cur = code_at(cell);
                    seen |= cur;
                    if (VERIFY && (cur & LOSF_DIRTY)) risky = true;  // a segment that is not a union of unit edges
                    active = n2 != 0u && !(VERIFY && risky);
                }
            }
        }
    }
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

// ------------------------------------------------------------------------------------------------
// Full Game.getValidMove(entity, x, y, z, checkAlternate=false) (game/collision.go:69-203, 307-355): one
// thread per moving entity.  Walls: the bin walk above, collecting the DISTINCT segments the move segment
// intersects (a segment is registered in several cells).  Colliders: every lane of a warp reads the same
// collider at the same time (one broadcast load per warp) and applies the reference's proximity and Z
// pre-filters before the two geom.LineCircleIntersection stages.  The sorted collision list is reduced to
// what its consumers read — collisions[0] and len(collisions) (game/combat.go:335-341,
// game/update_units.go:31-33): the record nearest to the entity, earliest inserted on ties.
struct MoveParams {
    LosParams los;
    const pm3d_mover *movers;
    const pm3d_collider *colliders;
    pm3d_move_result *out;
    int n_colliders, player;
    // colliders binned by the 1x1 cell their position falls in (pm3d_upload_colliders); cb_w == 0: not binned
    const int32_t *cb_start;  // CSR over [cb_x0, cb_x0 + cb_w) x [cb_y0, cb_y0 + cb_h), ids ascending per cell
    const int32_t *cb_items;
    const int32_t *cb_loose;  // colliders outside the binned area (or not finite): always candidates
    int cb_x0, cb_y0, cb_w, cb_h, n_loose;
};

#define GEOM_EPS 1e-14
__device__ __forceinline__ double sqd(double x) { return x * x; }

// geom.LineCircleIntersection(li, ci, isSegment=true) [R] — the Rosetta-code routine raycaster-go adopted;
// operations in the routine's own order, no contraction (-fmad=false).  Returns 0..2 points.
__device__ __noinline__ int line_circle_intersection(double x1, double y1, double x2, double y2, double x0, double y0,
                                                    double radius, double *out) {
    int n = 0;
    const double A = y2 - y1, B = x1 - x2, C = x2 * y1 - x1 * y2;
    const double a = sqd(A) + sqd(B);
    double b, c;
    const bool bnz = fabs(B) >= GEOM_EPS;
    if (bnz) {
        b = 2 * (A * C + A * B * y0 - sqd(B) * x0);
        c = sqd(C) + 2 * B * C * y0 - sqd(B) * (sqd(radius) - sqd(x0) - sqd(y0));
    } else {
        b = 2 * (B * C + A * B * x0 - sqd(A) * y0);
        c = sqd(C) + 2 * A * C * x0 - sqd(A) * (sqd(radius) - sqd(x0) - sqd(y0));
    }
    double d = sqd(b) - 4 * a * c;
    if (d < 0) return 0;
    const double d1 = sqrt(sqd(x2 - x1) + sqd(y2 - y1));
    auto within = [&](double x, double y) -> bool {
        const double d2 = sqrt(sqd(x - x1) + sqd(y - y1)), d3 = sqrt(sqd(x2 - x) + sqd(y2 - y));
        return fabs(d1 - d2 - d3) < GEOM_EPS;
    };
    const int roots = d == 0 ? 1 : 2;
    if (d != 0) d = sqrt(d);
    for (int k = 0; k < roots; k++) {
        const double num = roots == 1 ? -b : (k == 0 ? -b + d : -b - d);
        double x, y;
        if (bnz) { x = num / (2 * a); y = -(A * x + C) / B; }
        else     { y = num / (2 * a); x = -(B * y + C) / A; }
        if (within(x, y)) { out[n * 2] = x; out[n * 2 + 1] = y; n++; }
    }
    return n;
}

// zEntityMinMax (game/collision.go:380-394)
__device__ __forceinline__ void z_min_max(double z, double height, int anchor, double &lo, double &hi) {
    lo = 0; hi = 0;
    if (anchor == PM3D_ANCHOR_BOTTOM) { lo = z; hi = z + height; }
    else if (anchor == PM3D_ANCHOR_CENTER) { lo = z - height / 2; hi = z + height / 2; }
    else if (anchor == PM3D_ANCHOR_TOP) { lo = z - height; hi = z; }
}

#define MOVE_MAX_WALL_HITS 8
#define MOVE_MAX_CANDIDATES 16
#ifndef MOVE_MIN_BLOCKS
#define MOVE_MIN_BLOCKS 8  // 64 registers: the spills cost less than the warps they make room for (4 Mi movers: 0.89 -> 0.74 ms)
#endif
__global__ void __launch_bounds__(128, MOVE_MIN_BLOCKS) k_valid_move(MoveParams q) {
    const LosParams &p = q.los;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const pm3d_mover m = q.movers[i];
    pm3d_move_result r;
    double posX = m.pos_x, posY = m.pos_y, posZ = m.pos_z;
    r.x = posX; r.y = posY; r.z = posZ;
    r.collided = 0; r.n_collisions = 0; r.first_entity = -1; r.n_points = 0;
    r.first_x = r.first_y = r.first_z = 0; r.near_x = r.near_y = __longlong_as_double(0x7ff8000000000000ll);
    if (m.radius <= 0 || m.height <= 0 || (posX == m.move_x && posY == m.move_y && posZ == m.move_z)) {
        q.out[i] = r;
        return;
    }
    double newX = m.move_x, newY = m.move_y, newZ = m.move_z;
    const double4 ray = make_double4(posX, posY, newX, newY);
    const double checkDist = sqrt(sqd(newX - posX) + sqd(newY - posY) + sqd(newZ - posZ));  // geom3d.Line3d.Distance
    const double prox = ceil(checkDist);                                                 // model.PointInProximity

    int n_points = 0, n_collisions = 0;
    double near_d2 = INFINITY, best_d2 = 0;
    bool have_best = false;
    auto add_point = [&](double x, double y) {
        n_points++;
        const double d2 = sqd(x - posX) + sqd(y - posY);
        if (d2 < near_d2) { near_d2 = d2; r.near_x = x; r.near_y = y; }
    };
    auto consider = [&](int entity, double x, double y, double z) {
        n_collisions++;
        const double d2 = sqd(x - posX) + sqd(y - posY);
        if (!have_best || d2 < best_d2 || (best_d2 != best_d2 && d2 == d2)) {
            have_best = true; best_d2 = d2;
            r.first_entity = entity; r.first_x = x; r.first_y = y; r.first_z = z;
        }
    };

    // ---- walls (:93-97): distinct intersected segments, then their points in segment order
    int wall_id[MOVE_MAX_WALL_HITS];
    int n_wall = 0;
    bool brute = false;
    {
        auto visit = [&](int id) {
            const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
            const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
            if (!segments_intersect(ray, make_double4(s01.x, s01.y, s23.x, s23.y))) return;
            for (int k = 0; k < n_wall && k < MOVE_MAX_WALL_HITS; k++)
                if (wall_id[k] == id) return;
            if (n_wall < MOVE_MAX_WALL_HITS) wall_id[n_wall] = id;
            n_wall++;
        };
        const double gx0 = (double)p.bin_x0, gy0 = (double)p.bin_y0;
        const double gx1 = (double)(p.bin_x0 + p.bin_w), gy1 = (double)(p.bin_y0 + p.bin_h);
        const bool inside = ray.x >= gx0 && ray.x < gx1 && ray.z >= gx0 && ray.z < gx1 && ray.y >= gy0 && ray.y < gy1 &&
                            ray.w >= gy0 && ray.w < gy1;
        if (inside) {
            int cx = (int)floor(ray.x) - p.bin_x0, cy = (int)floor(ray.y) - p.bin_y0;
            const int ex = (int)floor(ray.z) - p.bin_x0, ey = (int)floor(ray.w) - p.bin_y0;
            const double dx = ray.z - ray.x, dy = ray.w - ray.y;
            const int sx = dx > 0 ? 1 : -1, sy = dy > 0 ? 1 : -1;
            const double inv_dx = dx != 0 ? 1.0 / fabs(dx) : 0.0, inv_dy = dy != 0 ? 1.0 / fabs(dy) : 0.0;
            double tx = dx != 0 ? (dx > 0 ? (floor(ray.x) + 1.0 - ray.x) : (ray.x - floor(ray.x))) * inv_dx : 1e300;
            double ty = dy != 0 ? (dy > 0 ? (floor(ray.y) + 1.0 - ray.y) : (ray.y - floor(ray.y))) * inv_dy : 1e300;
            int steps = abs(ex - cx) + abs(ey - cy);
            for (;;) {
                const int cell = cy * p.bin_w + cx;
                const int b = __ldg(p.cell_start + cell), e = __ldg(p.cell_start + cell + 1);
                for (int k = b; k < e; k++) visit(__ldg(p.cell_items + k));
                if (steps-- <= 0) break;
                if ((tx < ty && cx != ex) || cy == ey) { tx += inv_dx; cx += sx; }
                else { ty += inv_dy; cy += sy; }
            }
        }
        brute = !inside || n_wall > MOVE_MAX_WALL_HITS;  // outside the bins, or more hits than the list holds
    }
    const bool player_block = q.player >= 0 && m.self != q.player && m.parent != q.player && !m.destroyed;
    // the player block (:100-132) re-reads every wall point, so it needs them in segment order
    auto for_each_wall_point = [&](auto &&fn) {
        if (brute) {
            for (int id = 0; id < p.n_segs; id++) {
                const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
                const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
                double px, py;
                if (line_intersection(ray, make_double4(s01.x, s01.y, s23.x, s23.y), &px, &py)) fn(px, py);
            }
        } else {
            // ascending segment id (insertion order of the reference loop): selection over <= 8 ids
            int prev = -1;
            for (int k = 0; k < n_wall; k++) {
                int id = 0x7fffffff;
                for (int t = 0; t < n_wall; t++)
                    if (wall_id[t] > prev && wall_id[t] < id) id = wall_id[t];
                prev = id;
                const double2 s01 = __ldg(reinterpret_cast<const double2 *>(p.segs + id));
                const double2 s23 = __ldg(reinterpret_cast<const double2 *>(p.segs + id) + 1);
                double px, py;
                line_intersection(ray, make_double4(s01.x, s01.y, s23.x, s23.y), &px, &py);
                fn(px, py);
            }
        }
    };
    for_each_wall_point([&](double px, double py) { add_point(px, py); });
    const int walls = n_points;

    // ---- g.player.Unit (:100-132)
    if (player_block) {
        const pm3d_collider pl = q.colliders[q.player];
        if (fabs(pl.x - newX) <= prox && fabs(pl.y - newY) <= prox) {
            double smin, smax, tmin, tmax, zI = -1;
            z_min_max(newZ, m.height, m.anchor, smin, smax);
            z_min_max(pl.z, pl.height, pl.anchor, tmin, tmax);
            if (!(smin > tmax || tmin > smax)) {
                const double mid = smin + (smax - smin) / 2;
                zI = mid < tmin ? tmin : (mid > tmax ? tmax : mid);
            }
            double comb[4];
            const int nc = line_circle_intersection(posX, posY, newX, newY, pl.x, pl.y, pl.radius + m.radius, comb);
            if (zI >= 0 && nc > 0) {
                double extra[8];
                int n_extra = 0;
                for (int k = 0; k < nc; k++) {
                    double got[4];
                    const int ng = line_circle_intersection(comb[k * 2], comb[k * 2 + 1], pl.x, pl.y, pl.x, pl.y, pl.radius, got);
                    for (int g = 0; g < ng; g++) {
                        extra[n_extra * 2] = got[g * 2]; extra[n_extra * 2 + 1] = got[g * 2 + 1]; n_extra++;
                        add_point(got[g * 2], got[g * 2 + 1]);
                    }
                    // :123-127 appends one collision per point of intersectPoints so far — wall points included
                    if (walls > 0) for_each_wall_point([&](double px, double py) { consider(q.player, px, py, zI); });
                    for (int e = 0; e < n_extra; e++) consider(q.player, extra[e * 2], extra[e * 2 + 1], zI);
                }
            }
        }
    }

    // ---- collidable sprites (:135-194), list order.  The reference tests every sprite against the proximity box
    //      |sprite - new| <= ceil(checkDist) first; with the colliders binned by cell only the cells that box touches
    //      are read.  The candidates are put in ascending list order before they are processed, so counts, nearest
    //      point and collisions[0] (ties: earliest inserted) are those of the full loop.
    auto collide = [&](int s) {
        if (s == q.player || s == m.self || s == m.parent) return;
        const pm3d_collider *sp = q.colliders + s;
        const double sx = sp->x, sy = sp->y;
        if (!(fabs(sx - newX) <= prox && fabs(sy - newY) <= prox)) return;
        const double srad = sp->radius;
        if (srad <= 0 || sp->destroyed) return;
        double smin, smax, tmin, tmax;
        z_min_max(newZ, m.height, m.anchor, smin, smax);
        z_min_max(sp->z, sp->height, sp->anchor, tmin, tmax);
        if (smin > tmax || tmin > smax) return;
        const double mid = smin + (smax - smin) / 2;
        const double zI = mid < tmin ? tmin : (mid > tmax ? tmax : mid);
        if (zI < 0) return;
        const double comb_r = srad + m.radius;
        double comb[4];
        const int nc = line_circle_intersection(posX, posY, newX, newY, sx, sy, comb_r, comb);
        bool circle_hit = false;
        for (int k = 0; k < nc; k++) {
            double got[4];
            const int ng = line_circle_intersection(comb[k * 2], comb[k * 2 + 1], sx, sy, sx, sy, srad, got);
            for (int g = 0; g < ng; g++) {
                circle_hit = true;
                add_point(got[g * 2], got[g * 2 + 1]);
                consider(s, got[g * 2], got[g * 2 + 1], zI);
            }
        }
        if (!circle_hit && sqrt(sqd(sx - newX) + sqd(sy - newY)) <= comb_r) {  // inside the circle without touching it
            add_point(newX, newY);
            consider(s, newX, newY, zI);
        }
    };
    bool binned = false;
    if (q.cb_w > 0 && prox <= 3.0 && fabs(newX) < 1e9 && fabs(newY) < 1e9) {
        int cand[MOVE_MAX_CANDIDATES];
        int nc = 0;
        bool overflow = false;
        const int cx0 = max((int)floor(newX - prox) - q.cb_x0, 0), cx1 = min((int)floor(newX + prox) - q.cb_x0, q.cb_w - 1);
        const int cy0 = max((int)floor(newY - prox) - q.cb_y0, 0), cy1 = min((int)floor(newY + prox) - q.cb_y0, q.cb_h - 1);
        for (int cx = cx0; cx <= cx1 && !overflow; cx++) {
            // the cells of one column are consecutive in the CSR: one range per column
            if (cy0 > cy1) break;
            const int b = __ldg(q.cb_start + cx * q.cb_h + cy0), e = __ldg(q.cb_start + cx * q.cb_h + cy1 + 1);
            for (int k = b; k < e; k++) {
                if (nc == MOVE_MAX_CANDIDATES) { overflow = true; break; }
                cand[nc++] = __ldg(q.cb_items + k);
            }
        }
        for (int k = 0; k < q.n_loose && !overflow; k++) {
            if (nc == MOVE_MAX_CANDIDATES) { overflow = true; break; }
            cand[nc++] = __ldg(q.cb_loose + k);
        }
        if (!overflow) {
            for (int a = 1; a < nc; a++) {  // ascending list order (insertion sort of <= 16 ids)
                const int v = cand[a];
                int b = a - 1;
                for (; b >= 0 && cand[b] > v; b--) cand[b + 1] = cand[b];
                cand[b + 1] = v;
            }
            for (int k = 0; k < nc; k++) collide(cand[k]);
            binned = true;
        }
    }
    if (!binned)
        for (int s = 0; s < q.n_colliders; s++) collide(s);

    r.n_points = n_points;
    r.n_collisions = n_collisions;
    if (n_points > 0) {  // :202, :307-311 "cannot move"
        r.collided = 1;
        q.out[i] = r;
        return;
    }
    // ---- :314-352 map clamp, level-0 cell test, floor clamp
    int64_t ix = go_int(newX), iy = go_int(newY);
    if (ix < 0 || newX < 0) { newX = p.clip; ix = 0; }
    else if (ix >= p.map_w) { newX = (double)p.map_w - p.clip; ix = go_int(newX); }
    if (iy < 0 || newY < 0) { newY = p.clip; iy = 0; }
    else if (iy >= p.map_h) { newY = (double)p.map_h - p.clip; iy = go_int(newY); }
    ix = ix > p.map_w - 1 ? p.map_w - 1 : (ix < 0 ? 0 : ix);
    iy = iy > p.map_h - 1 ? p.map_h - 1 : (iy < 0 ? 0 : iy);
    if (__ldg(p.grid0 + ix * p.map_h + iy) == 0) { posX = newX; posY = newY; }
    else r.collided = 1;
    posZ = newZ;
    double zMin, zMax;
    z_min_max(0.0, m.height, m.anchor, zMin, zMax);
    zMin = fabs(zMin);
    if (posZ < zMin) { posZ = zMin; r.collided = 1; }
    r.x = posX; r.y = posY; r.z = posZ;
    q.out[i] = r;
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
    { int rc_ = pm3d_main_enter(ctx); if (rc_) return rc_; }
    pm3d_main_touch(ctx);
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
    p.cell_rec = ctx->cell_rec;
    p.n = n;
    p.n_segs = ctx->n_segs; p.bin_x0 = ctx->bin_x0; p.bin_y0 = ctx->bin_y0; p.bin_w = ctx->bin_w; p.bin_h = ctx->bin_h;
    p.max_skip = 2;
    if (const char *e = getenv("PM3D_LOS_SKIP")) p.max_skip = atoi(e);  // tuning override (tools/)
    p.grid0 = ctx->grid; p.map_w = ctx->map_w; p.map_h = ctx->map_h; p.clip = clip;
    p.out_flag = d_flag;
    p.out_pos = (double2 *)d_pos;
    const int64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffll) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "batch too large");
    p.cell_code = nullptr; p.list = nullptr; p.list_count = nullptr;
    int fast = ctx->cell_code != nullptr;
    if (const char *e = getenv("PM3D_LOS_FAST")) fast = fast && atoi(e) != 0;  // tuning / test override
    if (is_move) {
        k_move<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
    } else if (fast) {
        // two passes: k_los_fast decides what the cell codes alone can prove, k_los takes the rest from the list
        if (!ctx->los_count) PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->los_count, 64));
        const int64_t chunk_max = 1ll << 30;  // the list holds 32-bit ray numbers
        {
            void *lp = ctx->los_list;
            int rc = ensure_io(ctx, &lp, &ctx->los_list_bytes, (size_t)std::min<int64_t>(n, chunk_max) * 4);
            ctx->los_list = (uint32_t *)lp;
            if (rc) return rc;
        }
        const size_t code_bytes = (((size_t)ctx->bin_w * ctx->bin_h + 15) / 16) * 16;
        const bool smem = code_bytes <= 100 * 1024;
        if (smem && !ctx->los_smem_optin) {
            PM3D_CUDA(ctx, cudaFuncSetAttribute(k_los_fast<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            PM3D_CUDA(ctx, cudaFuncSetAttribute(k_los_fast<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            ctx->los_smem_optin = true;
        }
        int per_sm_a = 1, per_sm_b = 2;
        if (smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_los_fast<true, true>, LOSF_THREADS, code_bytes);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_los_fast<false, true>, LOSF_THREADS, 0);
        int verify = ctx->los_verify;
        if (const char *e = getenv("PM3D_LOS_VERIFY")) verify = atoi(e) != 0;  // tuning / test override
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_los, LOS_THREADS, 0);
        if (per_sm_a < 1) per_sm_a = 1;
        if (per_sm_b < 1) per_sm_b = 1;
        for (int64_t off = 0; off < n; off += chunk_max) {
            LosParams q = p;
            q.n = std::min<int64_t>(n - off, chunk_max);
            q.rays = p.rays + off;
            q.out_flag = p.out_flag + off;
            q.cell_code = ctx->cell_code;
            q.list = ctx->los_list;
            q.list_count = ctx->los_count;
            PM3D_CUDA(ctx, cudaMemsetAsync(ctx->los_count, 0, 4, ctx->stream));
            long long warps = (long long)ctx->sm_count * per_sm_a * (LOSF_THREADS / 32);
            long long per_warp = (q.n + warps - 1) / warps;
            if (per_warp < 32) per_warp = 32;
            warps = (q.n + per_warp - 1) / per_warp;
            const unsigned blocks_a = (unsigned)((warps + LOSF_THREADS / 32 - 1) / (LOSF_THREADS / 32));
            if (smem && verify) k_los_fast<true, true><<<blocks_a, LOSF_THREADS, code_bytes, ctx->stream>>>(q, per_warp);
            else if (smem) k_los_fast<true, false><<<blocks_a, LOSF_THREADS, code_bytes, ctx->stream>>>(q, per_warp);
            else if (verify) k_los_fast<false, true><<<blocks_a, LOSF_THREADS, 0, ctx->stream>>>(q, per_warp);
            else k_los_fast<false, false><<<blocks_a, LOSF_THREADS, 0, ctx->stream>>>(q, per_warp);
            // the second pass sizes its ranges from the device-side count; a grid of one wave is enough for any count
            long long warps_b = std::min<long long>((long long)ctx->sm_count * per_sm_b * (LOS_THREADS / 32), (q.n + 31) / 32);
            const unsigned blocks_b = (unsigned)((warps_b + LOS_THREADS / 32 - 1) / (LOS_THREADS / 32));
            k_los<<<blocks_b, LOS_THREADS, 0, ctx->stream>>>(q, 0);
            ctx->launches += 2;
        }
        ctx->launches--;  // (the common epilogue below counts one)
    } else {
        // persistent: 8 CTAs per SM, every warp owns a contiguous range of rays (at least 32)
        int per_sm = 2;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_los, LOS_THREADS, 0);
        if (per_sm < 1) per_sm = 1;
        if (const char *e = getenv("PM3D_LOS_WAVES")) per_sm *= atoi(e);  // tuning override (tools/)
        long long warps = (long long)ctx->sm_count * per_sm * (LOS_THREADS / 32);
        long long per_warp = (n + warps - 1) / warps;
        if (per_warp < 32) per_warp = 32;
        warps = (n + per_warp - 1) / per_warp;
        const unsigned los_blocks = (unsigned)((warps + LOS_THREADS / 32 - 1) / (LOS_THREADS / 32));
        k_los<<<los_blocks, LOS_THREADS, 0, ctx->stream>>>(p, per_warp);
    }
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

extern "C" int pm3d_upload_colliders(pm3d_ctx *ctx, const pm3d_collider *colliders, int n, int player_index) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || (n > 0 && !colliders) || player_index < -1 || player_index >= n)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "bad collider list (n = %d, player_index = %d)", n, player_index);
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    void *old[] = {ctx->colliders, ctx->cb_start, ctx->cb_items, ctx->cb_loose};
    for (void *p : old)
        if (p) PM3D_CUDA(ctx, cudaFree(p));
    ctx->colliders = nullptr; ctx->cb_start = ctx->cb_items = ctx->cb_loose = nullptr;
    ctx->cb_w = ctx->cb_h = ctx->n_loose = 0;
    PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->colliders, sizeof(pm3d_collider) * (size_t)(n > 0 ? n : 1)));
    if (n) PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->colliders, colliders, sizeof(pm3d_collider) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    // bin the colliders by the cell their position falls in: the area is the bounding box of the finite positions,
    // at most 4096 x 4096 cells; what lies outside (or is not finite) goes to the "loose" list
    if (n >= 32) {
        double minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
        for (int i = 0; i < n; i++)
            if (isfinite(colliders[i].x) && isfinite(colliders[i].y) && fabs(colliders[i].x) < 1e6 && fabs(colliders[i].y) < 1e6) {
                minx = std::min(minx, colliders[i].x); maxx = std::max(maxx, colliders[i].x);
                miny = std::min(miny, colliders[i].y); maxy = std::max(maxy, colliders[i].y);
            }
        if (maxx >= minx) {
            const int x0 = (int)floor(minx), y0 = (int)floor(miny);
            const int w = std::min((int)floor(maxx) - x0 + 1, 4096), h = std::min((int)floor(maxy) - y0 + 1, 4096);
            std::vector<int32_t> start((size_t)w * h + 1, 0), loose;
            std::vector<int32_t> cell(n, -1);
            for (int i = 0; i < n; i++) {
                const double x = colliders[i].x, y = colliders[i].y;
                if (isfinite(x) && isfinite(y) && x >= x0 && y >= y0 && x < (double)x0 + w && y < (double)y0 + h)
                    cell[i] = ((int)floor(x) - x0) * h + ((int)floor(y) - y0);
                if (cell[i] >= 0) start[cell[i] + 1]++; else loose.push_back(i);
            }
            for (size_t c = 0; c < (size_t)w * h; c++) start[c + 1] += start[c];
            std::vector<int32_t> items(std::max(n, 1)), fill(start.begin(), start.end() - 1);
            for (int i = 0; i < n; i++)  // ascending ids per cell
                if (cell[i] >= 0) items[fill[cell[i]]++] = i;
            if (loose.size() <= 8) {
                PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->cb_start, start.size() * 4));
                PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->cb_items, items.size() * 4));
                PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->cb_loose, std::max<size_t>(loose.size(), 1) * 4));
                PM3D_CUDA(ctx, cudaMemcpy(ctx->cb_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
                PM3D_CUDA(ctx, cudaMemcpy(ctx->cb_items, items.data(), items.size() * 4, cudaMemcpyHostToDevice));
                if (!loose.empty()) PM3D_CUDA(ctx, cudaMemcpy(ctx->cb_loose, loose.data(), loose.size() * 4, cudaMemcpyHostToDevice));
                ctx->cb_x0 = x0; ctx->cb_y0 = y0; ctx->cb_w = w; ctx->cb_h = h; ctx->n_loose = (int)loose.size();
            }
        }
    }
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_colliders = n;
    ctx->player_index = player_index;
    return PM3D_OK;
}

extern "C" int pm3d_valid_move_batch(pm3d_ctx *ctx, const pm3d_mover *movers, int64_t n, double clip_distance,
                                     pm3d_move_result *out, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (n < 0 || (n > 0 && (!movers || !out))) return pm3d_fail(ctx, PM3D_ERR_ARG, "null batch pointers");
    if (!ctx->cell_start) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_segments must be called first");
    if (!ctx->grid) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_grid must be called first");
    if (!ctx->colliders) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_upload_colliders must be called first (n = 0 is fine)");
    if (n == 0) return PM3D_OK;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc_ = pm3d_main_enter(ctx); if (rc_) return rc_; }
    pm3d_main_touch(ctx);
    const bool in_dev = flags & PM3D_IN_DEVICE, out_dev = flags & PM3D_OUT_DEVICE;
    const pm3d_mover *d_in = movers;
    pm3d_move_result *d_out = out;
    int rc;
    if (!in_dev) {
        if ((rc = ensure_io(ctx, &ctx->io_a, &ctx->io_a_bytes, (size_t)n * sizeof(pm3d_mover)))) return rc;
        PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->io_a, movers, (size_t)n * sizeof(pm3d_mover), cudaMemcpyHostToDevice, ctx->stream));
        d_in = (const pm3d_mover *)ctx->io_a;
    }
    if (!out_dev) {
        if ((rc = ensure_io(ctx, &ctx->io_c, &ctx->io_c_bytes, (size_t)n * sizeof(pm3d_move_result)))) return rc;
        d_out = (pm3d_move_result *)ctx->io_c;
    }
    MoveParams q;
    LosParams &p = q.los;
    p.rays = nullptr;
    p.segs = (const double4 *)ctx->segs;
    p.seg_irect = ctx->seg_irect;
    p.cell_start = ctx->cell_start;
    p.cell_items = ctx->cell_items;
    p.cell_rec = ctx->cell_rec;
    p.n = n;
    p.n_segs = ctx->n_segs;
    p.bin_x0 = ctx->bin_x0; p.bin_y0 = ctx->bin_y0; p.bin_w = ctx->bin_w; p.bin_h = ctx->bin_h;
    p.max_skip = 2;
    if (const char *e = getenv("PM3D_LOS_SKIP")) p.max_skip = atoi(e);  // tuning override (tools/)
    p.grid0 = ctx->grid;
    p.map_w = ctx->map_w; p.map_h = ctx->map_h;
    p.clip = clip_distance;
    p.out_flag = nullptr; p.out_pos = nullptr;
    q.movers = d_in; q.colliders = ctx->colliders; q.out = d_out;
    q.n_colliders = ctx->n_colliders; q.player = ctx->player_index;
    q.cb_start = ctx->cb_start; q.cb_items = ctx->cb_items; q.cb_loose = ctx->cb_loose;
    q.cb_x0 = ctx->cb_x0; q.cb_y0 = ctx->cb_y0; q.cb_w = ctx->cb_w; q.cb_h = ctx->cb_h; q.n_loose = ctx->n_loose;
    const long long blocks = (n + 127) / 128;
    if (blocks > 0x7fffffffll) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "batch too large");
    k_valid_move<<<(unsigned)blocks, 128, 0, ctx->stream>>>(q);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    if (!out_dev) PM3D_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(pm3d_move_result), cudaMemcpyDeviceToHost, ctx->stream));
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}
