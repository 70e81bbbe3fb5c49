// pm3d_internal.cuh — shared device helpers, HBM record layouts and the ctx of libpm3d.so.
// sm_100a only.  Compiled with -fmad=false: the path's binary64 arithmetic must not be
// contracted into FMAs (Go on amd64 never fuses), see DESIGN.md "numeric contract".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "pm3d.h"

// ------------------------------------------------------------------ HBM layouts

// The cast writes its per-column results as struct-of-arrays, [view][level][column] each, so that
// the compose kernel reads a lane's four adjacent columns as one contiguous 32 B / 64 B load:
//   col_span  int2        ya, yb     visible rows [ya, yb) of the slice, clipped to the screen; ya >= yb: nothing
//   col_tt    uint2       (tile << 16 | texX), tint (R | G<<8 | B<<16 | 255<<24)
//   col_ps    ulonglong2  pos0, step2: texel row of screen row y = (pos0 + y*step2) >> 33 (mod 2^64)
// and the level-0 floor-cast parameters as [view][column]:
//   flo_xw, flo_yw, flo_dw  double  floorXWall, floorYWall, distWall
//   flo_rc                  double  RN(1/distWall) when the column is eligible for the fast floor cast, else 0
//   flo_dx, flo_dy          double  RN(floorXWall*texSize - posX*texSize), same for Y (fast floor cast)
//   flo_fs                  int     first floor row (drawEnd + 1, or h when none)
// plus one aggregate per (view, level, 64-column group, 32-column half):
//   agg  int4  x: min ya over drawn slices (INT_MAX if none); y: min fs; z: max fs (level 0 only);
//              w: float bits, upper bound of the z-buffer over the 32 columns (level 0 only, -inf if every z is NaN)
#define GROUP_COLS 64

// Per (view, screen row) values shared by every column.  16 B.
struct __align__(16) RowRec {
    double cd;       // currentDist of the floor cast
    uint32_t tint;   // floor light tint of this row
    uint32_t info;   // bit31: floor row valid (render distance); bits 0..15: sky/floor-box texel row
};
static_assert(sizeof(RowRec) == 16, "RowRec");

// One visible sprite of a view, in paint order (far -> near).  112 B.
struct __align__(16) DrawRec {
    int32_t dsx, dex, dsy, dey;  // unclipped screen rect [dsx,dex) x [dsy,dey)
    double transformY;           // depth for the z test
    uint64_t ypos0, ystep2;      // sprite texel row of screen row y = (ypos0 + y*ystep2) >> 33
    uint64_t xmagic;             // ceil(2^64 / spriteWidth): texX = mulhi(k*texW, xmagic)
    uint64_t sheet_px;           // pixel index of the rect's (rx0, ry0) inside sheet_pixels (may wrap)
    uint32_t tex_w, tex_h;       // TextureRect size
    uint32_t sheet_w;            // sheet row pitch in pixels
    uint32_t tint;
    uint16_t tx_lo, tx_hi, ty_lo, ty_hi;  // texel ranges that fall inside the sheet
    double dist;                 // euclidean camera distance (convergence pick)
    int32_t src;                 // index into the view's sprite list
    uint32_t flags;              // bit0: IsFocusable()
    uint64_t _pad0, _pad1;
};
static_assert(sizeof(DrawRec) == 112, "DrawRec");

struct SheetInfo {
    int64_t px_offset;  // first pixel of the sheet inside sheet_pixels (in pixels)
    int32_t w, h;
};

#define PM3D_SPARSE_MAX 32

// ------------------------------------------------------------------ ctx

#define PM3D_MAX_LANES 4
#define PM3D_SMALL_SLOTS 8
struct pm3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;

    // mission data
    uint16_t *grid = nullptr;
    uint8_t *dist = nullptr;  // [stored_levels][map_w][map_h] free-run distance field of the grid
    uint16_t *qdist = nullptr;  // [stored_levels][4][map_w][map_h] per step quadrant: steps to the nearest wall cell / map exit
    double *sparse_cells = nullptr;   // [stored_levels][PM3D_SPARSE_MAX][2] wall-cell centres of sparse levels
    int32_t *sparse_count = nullptr;  // [stored_levels] number of wall cells, -1 = more than PM3D_SPARSE_MAX
    int stored_levels = 0, map_w = 0, map_h = 0, num_levels = 0;
    int16_t *wall_lut = nullptr;  // [32768][2]
    uint8_t *atlas = nullptr;
    int n_tiles = 0, tex_size = 0, tex_shift = 0;
    int16_t *floor_tiles = nullptr;  // [map_w][map_h]
    int floor_w = 0, floor_h = 0;
    int sky_tile = -1, floorbox_tile = -1;
    int max_wall_tile = -1, max_floor_tile = -1;  // validated against n_tiles before every render
    uint8_t *sheet_pixels = nullptr;
    SheetInfo *sheet_info = nullptr;
    int n_sheets = 0;

    // view configuration
    int w = 0, h = 0, max_views = 0, max_sprites = 0, n_groups = 0;
    int keep_columns = 0, force_painter = 0, dda_jump = 1;
    int last_views = 0;

    // per-call device buffers
    pm3d_view *views = nullptr;
    int2 *col_span = nullptr;        // [views][levels][w]
    uint2 *col_tt = nullptr;
    ulonglong2 *col_ps = nullptr;
    double *flo_xw = nullptr, *flo_yw = nullptr, *flo_dw = nullptr, *flo_rc = nullptr, *flo_dx = nullptr, *flo_dy = nullptr;  // [views][w]
    int32_t *flo_fs = nullptr;
    int4 *agg = nullptr;             // [views][n_groups][levels][2]
    double *zbuf = nullptr;          // [views][w]
    RowRec *row_rec = nullptr;      // [views][h]
    pm3d_column *dbg_cols = nullptr;  // [views][levels][w] when keep_columns
    int dbg_levels = 0, dbg_views = 0;
    pm3d_sprite *sprites = nullptr;  // [max_sprites]
    int32_t *sprite_offset = nullptr;  // [max_views + 1]
    DrawRec *draw_rec = nullptr;     // [max_sprites], per view contiguous at sprite_offset[v]
    int32_t *n_draw = nullptr;       // [max_views]
    uint16_t *bin_list = nullptr;    // [max_sprites * n_groups]
    int4 *bin_rect = nullptr;        // [max_sprites * n_groups]
    int4 *bin_count = nullptr;       // [max_views * n_groups]: entries, min dsy, max dey
    pm3d_rect *rects = nullptr;      // [max_sprites]
    pm3d_convergence *conv = nullptr;  // [max_views]
    uint8_t *frame = nullptr;        // [max_views][h][w][4]
    uint8_t *last_frame = nullptr;   // where the last render wrote (frame or caller memory)
    // PM3D_OPT_PIPELINE: ring of internal frame slots so the D2H of frame i (copy stream) overlaps
    // the kernels of frame i+1 (compute stream)
    cudaStream_t copy_stream = nullptr;
    int ring_n = 1, ring_pos = 0;
    uint8_t *frame_ring[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_rendered[4] = {}, ev_copied[4] = {};
    uint64_t *checksums = nullptr;   // [max_views]
    // Small results (rectangles, convergence) of pipelined host renders: a kernel stores them into mapped pinned
    // memory right after compose, so no few-KB copy-engine operation sits between the 8 MB frame copies; the host
    // hands them to the caller's buffers at pm3d_sync (or when the slot comes round again).
    struct SmallPending {
        pm3d_rect *dst_rects = nullptr;
        pm3d_convergence *dst_conv = nullptr;
        int n_rects = 0, n_conv = 0;
        cudaEvent_t done = nullptr;
        bool active = false;
    } small[PM3D_SMALL_SLOTS];
    uint8_t *small_ring = nullptr;   // [PM3D_SMALL_SLOTS][small_slot_bytes], cudaHostAllocMapped
    size_t small_slot_bytes = 0, small_rect_cap = 0;
    int small_pos = 0;

    // Two sets of the per-frame buffers above (PM3D_FRAME_FIELDS) + input staging, so that the front kernel of
    // frame i+1 (front_stream) overlaps the compose kernel of frame i (stream).  The named fields above always
    // alias the set of the frame rendered last (downloads read them); PM3D_OPT_OVERLAP = 0 uses one set/stream.
    struct FrameSet {
        void *p[24] = {};
        uint8_t *inbuf = nullptr;      // device copy of one call's [views | sprite_offset | sprites]
        size_t inbuf_bytes = 0;
        void *stage = nullptr;         // pinned host staging of the same block
        size_t stage_bytes = 0;
        cudaEvent_t ev_stage = nullptr;    // the H2D out of `stage` has executed
        cudaEvent_t ev_front = nullptr;    // front kernel done (recorded on front_stream)
        cudaEvent_t ev_compose = nullptr;  // compose done and the set's results read: the set may be rewritten
    } sets[2 * PM3D_MAX_LANES];
    int n_sets = 2, cur_set = 0;
    // PM3D_OPT_LANES: independent frames of enqueue-only callers (pm3d_render_resident) are dealt round-robin to
    // n_lanes lanes, each with its own compose stream, front stream and pair of frame sets (lane l owns sets 2l and
    // 2l+1; lane 0 is `stream` / `front_stream`), so the tail of one frame's compose launch and the 24-CTA front
    // launch of another no longer leave most of the chip idle.
    int n_lanes = 1, lane_pos = 0;
    int lane_cur[PM3D_MAX_LANES] = {};
    cudaStream_t lane_cstream[PM3D_MAX_LANES] = {}, lane_fstream[PM3D_MAX_LANES] = {};
    cudaEvent_t lane_join[PM3D_MAX_LANES] = {};
    cudaStream_t cur_cstream = nullptr;      // compose stream of the frame batch being enqueued
    cudaStream_t last_cstream = nullptr;     // ... of the frame rendered last (readers of its results wait for it)
    cudaStream_t front_stream = nullptr;     // highest priority: single frames (a handful of long-running front CTAs)
    cudaStream_t front_stream_lo = nullptr;  // default priority: large batches (both launches fill the chip)
    cudaStream_t cur_fstream = nullptr;      // the one the frame batch being enqueued uses

    // ---- ordering between the main stream and the frame streams.  The ctx promises ONE stream's semantics to its
    // caller (include/pm3d.h): whatever was enqueued on `stream` before a frame call precedes that frame, and
    // whatever is enqueued on `stream` after frame calls follows them.  Frames run on other streams (front streams,
    // lanes), so the two directions are made explicit:
    //   main -> frame: every entry point that enqueues non-frame work on `stream` bumps main_seq (pm3d_main_touch);
    //                  pm3d_begin_frame records ev_main on `stream` and makes the frame's front stream wait for it
    //                  whenever that stream has not yet seen the current main_seq (a pure frame loop pays nothing);
    //   frame -> main: pm3d_main_enter joins the compose streams of lanes >= 1 into `stream` when frames were
    //                  dealt to them since the last join.
    // A ctx on a CALLER stream cannot see what the caller enqueues there: every frame then counts as "main touched".
    cudaEvent_t ev_main = nullptr;
    uint64_t main_seq = 1;
    uint64_t seen_front = 0, seen_front_lo = 0, seen_lane[PM3D_MAX_LANES] = {};
    bool lanes_dirty = false;
    // ---- recorded frame sequences (pm3d_frames_capture_begin / _end / pm3d_frames_replay): the enqueue-only frame
    // calls between begin and end are captured from all the frame streams into ONE CUDA graph; a replay is a single
    // launch on the ctx stream instead of ~8 runtime calls per frame.  While capturing, a frame set only waits for
    // its previous user when that user is part of the capture (an event recorded before it is not capturable).
    bool capturing = false;
    bool set_in_capture[2 * PM3D_MAX_LANES] = {};
    int64_t capture_launches0 = 0;
    cudaEvent_t ev_fork = nullptr, ev_join[2 + 2 * PM3D_MAX_LANES] = {};
    struct FrameGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; uint8_t *last_frame = nullptr; int last_views = 0, last_sprites = 0; };
    std::vector<FrameGraph> graphs;
    // device -> host error words in mapped pinned memory (kernels store, pm3d_sync reads after the streams drained):
    //   [0] 1 + index of a view whose sprite_offset entries are out of range (resident path: offsets live in HBM)
    //   [1] 1 when pm3d_raycast_sprites_batch listed more than max_total sprites (PM3D_NO_SYNC path)
    uint32_t *status_h = nullptr, *status_d = nullptr;

    // inputs the kernels read this call (ctx-owned copies, or caller device memory)
    const pm3d_view *cur_views = nullptr;
    const pm3d_sprite *cur_sprites = nullptr;
    const int32_t *cur_offset = nullptr;
    int last_sprites = 0;
    int cur_n_sprites = 0;           // sprites of the frame batch being enqueued

    // per-kernel event timing (pm3d_ktimer_*)
    cudaEvent_t *kt_ev = nullptr;  // [kt_cap][3][2]
    int kt_cap = 0, kt_n = 0, kt_on = 0;


    // LOS / collision
    double *segs = nullptr;          // [n_segs][4]
    int4 *seg_irect = nullptr;       // Go image.Rect of int-truncated endpoints
    int n_segs = 0;
    int32_t *cell_start = nullptr;   // CSR over bin cells
    int32_t *cell_items = nullptr;
    int4 *cell_rec = nullptr;        // count + first three candidate ids (or offset) per bin cell
    int bin_x0 = 0, bin_y0 = 0, bin_w = 0, bin_h = 0;
    // two-pass LOS (los_kernel.cu k_los_fast): one code byte per bin cell (candidates present + which of its four
    // unit edges an axis-parallel segment covers), the list of rays the first pass leaves to k_los, its length
    uint8_t *cell_code = nullptr;
    uint32_t *los_list = nullptr;
    size_t los_list_bytes = 0;
    unsigned *los_count = nullptr;
    bool los_smem_optin = false;
    int los_verify = 0;  // first pass verifies crossings next to candidate cells (maps where few cells have candidates)
    // batched Pathing.FindPath (path_kernel.cu): node states and queues of the searches in flight
    void *path_scratch = nullptr;
    size_t path_scratch_bytes = 0;
    // presentation step (present_kernel.cu)
    int32_t *present_map = nullptr;  // [present_w + present_h] output -> source column / row, -1 = outside
    int present_w = 0, present_h = 0, present_src_w = 0, present_src_h = 0;
    uint8_t *present_buf = nullptr;
    size_t present_bytes = 0;
    pm3d_collider *colliders = nullptr;  // g.player.Unit + collidable sprites (pm3d_upload_colliders)
    int n_colliders = 0, player_index = -1;
    int32_t *cb_start = nullptr, *cb_items = nullptr, *cb_loose = nullptr;  // colliders binned by cell (los_kernel.cu)
    int cb_x0 = 0, cb_y0 = 0, cb_w = 0, cb_h = 0, n_loose = 0;
    // clutter (clutter_kernel.cu)
    uint8_t *clutter_class = nullptr;   // [map_w][map_h] floor-path class per cell
    void *clutter_tab = nullptr;        // class masks | clutter defs | per-draw generator constants
    size_t clutter_class_bytes = 0, clutter_tab_bytes = 0;
    int clutter_defs = 0;
    int64_t clutter_seed = 0;
    bool clutter_ready = false, clutter_smem_optin = false;
    void *io_a = nullptr, *io_b = nullptr, *io_c = nullptr;  // batch staging on device
    size_t io_a_bytes = 0, io_b_bytes = 0, io_c_bytes = 0;

    // multi-GPU (comm.cu): the NCCL communicator this ctx belongs to (ncclComm_t, opaque here) and the gather staging
    void *comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    bool comm_owned = false;
    void *comm_send = nullptr, *comm_recv = nullptr;
    size_t comm_send_bytes = 0, comm_recv_bytes = 0;

    // timers / L2 flush
    cudaEvent_t ev_start[16] = {}, ev_stop[16] = {};
    uint8_t *flush_buf = nullptr;
    size_t flush_bytes = 0;
};

int pm3d_fail(pm3d_ctx *ctx, int code, const char *fmt, ...);
enum { PM3D_STATUS_VIEW_SPRITES = 0, PM3D_STATUS_LIST_TOTAL = 1, PM3D_STATUS_WORDS = 16 };
int pm3d_main_enter(pm3d_ctx *ctx);                 // before enqueueing non-frame work on ctx->stream (see ev_main)
inline void pm3d_main_touch(pm3d_ctx *ctx) { ctx->main_seq++; }  // after it
int pm3d_check_status(pm3d_ctx *ctx);               // after a host sync: device error words -> PM3D_ERR_CAPACITY
// event pair around kernel `k` (0 cast, 1 view setup, 2 compose) of the frame being sampled
inline void pm3d_kt_mark(pm3d_ctx *ctx, int k, int end, cudaStream_t stream) {
    if (ctx->kt_on && ctx->kt_n < ctx->kt_cap) cudaEventRecord(ctx->kt_ev[(ctx->kt_n * 3 + k) * 2 + end], stream);
}
// the per-frame device buffers that exist once per frame set
#define PM3D_FRAME_FIELDS(X)                                                                                       \
    X(col_span) X(col_tt) X(col_ps) X(flo_xw) X(flo_yw) X(flo_dw) X(flo_rc) X(flo_dx) X(flo_dy) X(flo_fs) X(agg) \
    X(zbuf) X(row_rec) X(dbg_cols) X(draw_rec) X(n_draw) X(bin_list) X(bin_rect) X(bin_count) X(rects) X(conv)
inline void pm3d_load_set(pm3d_ctx *ctx, int k) {
    int i = 0;
#define X(f) ctx->f = (decltype(ctx->f))ctx->sets[k].p[i++];
    PM3D_FRAME_FIELDS(X)
#undef X
}
inline void pm3d_save_set(pm3d_ctx *ctx, int k) {
    int i = 0;
#define X(f) ctx->sets[k].p[i++] = (void *)ctx->f;
    PM3D_FRAME_FIELDS(X)
#undef X
}
#define PM3D_CUDA(ctx, call)                                                                         \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) return pm3d_fail((ctx), PM3D_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// kernel launchers (one per .cu)
int pm3d_launch_front(pm3d_ctx *ctx, int n_views, cudaStream_t stream);  // column cast + per-view setup, one launch
int pm3d_launch_compose(pm3d_ctx *ctx, int n_views, uint8_t *out);
int pm3d_launch_checksum(pm3d_ctx *ctx, int n_views, const uint8_t *frames);
int pm3d_wait_last_frame(pm3d_ctx *ctx);  // host-waits for the lane the last frame batch ran on (if not the main stream)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Go int(float64) on amd64 (CVTTSD2SQ): truncation, NaN / out of range -> INT64_MIN.
__device__ __forceinline__ int64_t go_int(double v) {
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;  // cvt.rzi.s64.f64
}
// Go int arithmetic wraps silently.
__device__ __forceinline__ int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
__device__ __forceinline__ int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
__device__ __forceinline__ int64_t wneg(int64_t a) { return (int64_t)(0ull - (uint64_t)a); }
__device__ __forceinline__ int64_t wmul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
__device__ __forceinline__ int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// blit model: nearest sample of n texels over len pixels
__device__ __forceinline__ uint64_t span_step(int64_t n_src, int64_t len_dst) {
    return ((uint64_t)n_src << 32) / (uint64_t)len_dst;
}

// true when v is finite, non-zero and its exponent lies in [2^-500, 2^500]: quotients and products of
// such values can neither overflow nor underflow
__device__ __forceinline__ bool exp_mid(double v) {
    const unsigned e = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
    return e - 523u <= 1000u;
}

// RN(a / b) from r = RN(1 / b) without a division: two Newton corrections with exact FMA residuals
// (Markstein's sequence: q1 is within an ulp of a/b, and a correction q1 + r*(a - b*q1) of a faithful
// quotient by the correctly rounded reciprocal rounds correctly).  Requires exp_mid(a) && exp_mid(b);
// callers fall back to the plain division otherwise.  pm3d_selfcheck_division compares it with `/`
// on billions of random and adversarial operand pairs (tests/test_parity_gpu.py).
__device__ __forceinline__ double div_by_rcp(double a, double b, double r) {
    const double q0 = a * r;
    const double e0 = __fma_rn(-b, q0, a);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-b, q1, a);
    return __fma_rn(e1, r, q1);
}

// ---- fast floor cast --------------------------------------------------------------------------
// The reference evaluates, per floor pixel (all binary64, every operation rounded):
//     weight = currentDist / distWall
//     fx     = weight * floorXWall + (1 - weight) * posX          (same for y)
//     cell   = int(fx),  texel = int(fx * texSize) % texSize
// Only the INTEGER int(fx * texSize) =: int(t) matters (cell = int(t) >> log2 texSize, texSize a power of two, so
// the scaling is exact).  The kernel evaluates t' = fma(w', D, P) with w' = RN(cd * RN(1/distWall)),
// D = RN(X1 - P), X1 = floorXWall * texSize, P = posX * texSize, then m = RN(t' + 2^32): for t' in
// [0, 2^32) the low 20 mantissa bits of m are the fraction of t' rounded to nearest in units of 2^-20 and the
// next 32 bits are its integer part (negative t' decode to >= 2^32 - 2^29, i.e. "out of bounds").
// Error bound, for |w| <= 4 and |X1|, |P| <= 2^24 (cast_kernel.cuh marks eligible columns, the compose kernel
// checks cd <= 4 |distWall| per lane row), with u = 2^-53 and T the real-number value:
//     |t  - T| <= u (2|W||X1| + |P|(|W| + 2|1-W|) + |t|)  <= 38 * 2^24 u
//     |t' - T| <= u (3|W||D| + |t'|)                       <= 40 * 2^24 u        => |t - t'| < 2^-22.7
// so whenever the rounded fraction is NOT zero (t' farther than 2^-21 from every integer) int(t') = int(t) and
// both lie on the same side of the map bounds.  A zero fraction (probability ~2^-20 per coordinate) sends the
// lane to the reference sequence.  tests: parity suite (every frame is compared bit for bit with the oracle,
// PM3D_OPT_FORCE_PAINTER runs the reference sequence everywhere) + pm3d_selfcheck_floor (selfcheck.cu).
#define PM3D_FLOOR_MAGIC 4294967296.0
__device__ __forceinline__ bool floor_fast_index(double w, double D, double P, uint32_t bound, uint32_t &idx) {
    const double m = __fma_rn(w, D, P) + PM3D_FLOOR_MAGIC;
    const uint32_t lo = (uint32_t)__double2loint(m), hi = (uint32_t)__double2hiint(m);
    idx = __funnelshift_r(lo, hi, 20);
    return (lo & 0xfffffu) != 0u && idx < bound;
}

// ---- DDA jump --------------------------------------------------------------------------------
// The reference's wall cast (raycaster-go Camera.castLevel: `if sideDistX < sideDistY { sideDistX += deltaDistX; mapX += stepX }
// else { sideDistY += deltaDistY; mapY += stepY }`) is a MERGE of two non-decreasing binary64 sequences
//     A_0 = sideDistX, A_{k+1} = RN(A_k + deltaDistX)        B_0 = sideDistY, B_{k+1} = RN(B_k + deltaDistY):
// every step takes an x step when A_i < B_j and a y step otherwise.  Only the cells it visits matter (the wall
// distance comes from the integer hit cell), and the state (i, j) after N steps is the unique split i + j = N with
//     (i == 0 or A_{i-1} < B_j)  and  (j == 0 or B_{j-1} <= A_i)                                  (merge path)
// whatever happened in between; the N-th step was an x step iff A_{i-1} >= B_{j-1} (a tie takes y first).
// dda_jump evaluates that split for a given N in closed form: i = ceil((B_0 - A_0 + (N-1) dY) / (dX + dY)) is the
// split of the real-number progressions, and the inequalities are then CERTIFIED on the true sequences through
//     |A_k - fma(k, dX, A_0)| <= (k + 1) 2^-53 A_k (1 + 2^-30)        (k rounded additions + one fma rounding)
// by requiring gaps larger than T = (N + 8) 2^-50 max(A_i, B_j) — four times the worst case.  A failed
// certificate (a near tie: rays on exact diagonals, or non-finite operands, whose comparisons all read false)
// returns false and the caller runs the reference loop instead.  Requires A_0, B_0 >= 0.
// pm3d_selfcheck_dda compares it with the sequential loop on random and tie-prone rays (selfcheck.cu).
// The two halves are separate so that the caller can fetch the landing cell's probe distance between them: the
// load then overlaps the certificate instead of waiting for it.
__device__ __forceinline__ void dda_jump_split(double dY, double c0, double inv, int N, int &i_out, int &j_out) {
    double r = ceil(__fma_rn((double)(N - 1), dY, c0) * inv);
    r = fmin(fmax(r, 0.0), (double)N);  // (NaN -> 0)
    i_out = (int)r;
    j_out = N - i_out;
}
__device__ __forceinline__ bool dda_jump_certify(double sX0, double sY0, double dX, double dY, int N, int i, int j, int &side_out) {
    const double a0 = __fma_rn((double)i, dX, sX0), b0 = __fma_rn((double)j, dY, sY0);  // A_i, B_j
    const double a1 = i > 0 ? __fma_rn((double)(i - 1), dX, sX0) : -INFINITY;           // A_{i-1}
    const double b1 = j > 0 ? __fma_rn((double)(j - 1), dY, sY0) : -INFINITY;           // B_{j-1}
    const double T = (double)(N + 8) * 8.8817841970012523e-16 * fmax(a0, b0);           // 2^-50
    if (a1 - b1 > T) {          // the last step was an x step; B_{j-1} < A_{i-1} <= A_i holds, certify A_{i-1} < B_j
        side_out = 0;
        return b0 - a1 > T;
    }
    if (b1 - a1 > T) {          // the last step was a y step; A_{i-1} < B_{j-1} <= B_j holds, certify B_{j-1} <= A_i
        side_out = 1;
        return a0 - b1 > T;
    }
    return false;
}
__device__ __forceinline__ bool dda_jump(double sX0, double sY0, double dX, double dY, double c0, double inv, int N,
                                         int &i_out, int &j_out, int &side_out) {
    dda_jump_split(dY, c0, inv, N, i_out, j_out);
    return dda_jump_certify(sX0, sY0, dX, dY, N, i_out, j_out, side_out);
}

// lighting: byte(ClampInt(int(base + sqrt(dist)*falloff + illumination (+ extra)), min, max)) per channel
__device__ __forceinline__ uint32_t light_tint(const pm3d_view &c, double dist, int base, double extra, bool use_extra) {
    double shadow = sqrt(dist) * c.light_falloff;
    double v = (double)base + shadow + c.global_illumination;
    if (use_extra) v = v + extra;
    int64_t iv = go_int(v);
    uint32_t r = (uint32_t)clamp_i64(iv, c.min_light[0], c.max_light[0]) & 255u;
    uint32_t g = (uint32_t)clamp_i64(iv, c.min_light[1], c.max_light[1]) & 255u;
    uint32_t b = (uint32_t)clamp_i64(iv, c.min_light[2], c.max_light[2]) & 255u;
    return r | (g << 8) | (b << 16) | (255u << 24);
}

// out = (c*tint + dst*(255-a) + 127) / 255 per channel (nvcc lowers /255u to an exact mul.hi + shift)
__device__ __forceinline__ uint32_t div255(uint32_t v) { return v / 255u; }
__device__ __forceinline__ uint32_t blend_over(uint32_t dst, uint32_t src, uint32_t tint) {
    uint32_t a = src >> 24;
    if (a == 0) return dst;
    uint32_t ia = 255u - a;
    uint32_t r = div255((src & 255u) * (tint & 255u) + (dst & 255u) * ia + 127u);
    uint32_t g = div255(((src >> 8) & 255u) * ((tint >> 8) & 255u) + ((dst >> 8) & 255u) * ia + 127u);
    uint32_t b = div255(((src >> 16) & 255u) * ((tint >> 16) & 255u) + ((dst >> 16) & 255u) * ia + 127u);
    uint32_t o = div255(a * 255u + (dst >> 24) * ia + 127u);
    r = min(r, 255u); g = min(g, 255u); b = min(b, 255u); o = min(o, 255u);
    return r | (g << 8) | (b << 16) | (o << 24);
}
#endif
