// present_kernel.cu — the step right after the raycast in GameScene.Draw (game/scene_game.go:169-185), fused
// into one pass over the output: nearest-neighbour scale of rayScreen by 1/renderScale onto the screen-sized
// image (:170-177) and then either the plain copy (:184) or the CRT shader (:179-182,
// game/resources/shaders/crt.kage, game/render/effects/crt.go:40-53).  The scaled renderScreen is never
// materialised: a shader tap reads the source texel the scaled image would hold there.
//
// One thread owns four adjacent output pixels of one row (one 128-bit store); a warp writes 512 contiguous
// bytes.  The source frame was just written by the compose kernel and is L2-resident, so the pass costs its
// output: 4 bytes per screen pixel (HBM-bound).
// Arithmetic: float32 like a Kage shader, in crt.kage's operation order without contraction (-fmad=false), so
// sampling positions are bit-identical with the CPU oracle; only powf / sinf differ from the C library's.
#include "pm3d_internal.cuh"

struct PresentParams {
    const uint8_t *src;   // [views][sh][sw][4]
    uint8_t *out;         // [views][oh][ow][4]
    const int32_t *xmap;  // [ow] source column of an output column, -1 = outside
    const int32_t *ymap;  // [oh]
    int sw, sh, ow, oh, mode, n_views;
    long long n_quads;    // views * oh * ceil(ow / 4)
};

__device__ __forceinline__ float clampf01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

// imageSrc0At(uv * size + origin) over the scaled image, split by axis: the source row / column an image
// coordinate falls on, or -1 outside the image or outside the scaled source
__device__ __forceinline__ int tap_row(const PresentParams &p, float v) {
    const float py = v * (float)p.oh;
    if (!(py >= 0.0f && py < (float)p.oh)) return -1;
    return __ldg(p.ymap + (int)floorf(py));
}
__device__ __forceinline__ int tap_col(const PresentParams &p, float u) {
    const float px = u * (float)p.ow;
    if (!(px >= 0.0f && px < (float)p.ow)) return -1;
    return __ldg(p.xmap + (int)floorf(px));
}
// channel c (0..2) of the texel as float(byte) / 255 (s_unorm holds exactly that quotient for every byte)
__device__ __forceinline__ float tap(const PresentParams &p, const uint32_t *__restrict__ frame, const float *s_unorm,
                                     int sx, int sy, int c) {
    if ((sx | sy) < 0) return 0.0f;
    const uint32_t t = __ldg(frame + (size_t)sy * p.sw + sx);
    return s_unorm[(t >> (8 * c)) & 255u];
}

// x^y for x >= 0 through exp2f(y * log2f(x)) (CUDA's full-accuracy exp2f / log2f, ~2 ulp each): the shader's pow
// is the GPU's own approximation anyway and the result is quantised to 8 bits (tests: <= 1 LSB vs the C library)
__device__ __forceinline__ float pow_pos(float x, float y) { return x > 0.0f ? exp2f(y * log2f(x)) : 0.0f; }

// scanline factor 0.4 + 0.7 * pow(clamp(0.35 + 0.35 * sin(uv.y * size.y * 1.5)), 1.7) of crt.kage:60-62
__device__ __forceinline__ float scan_factor(const PresentParams &p, float uy) {
    const float scans = clampf01(0.35f + 0.35f * sinf(uy * (float)p.oh * 1.5f));
    return 0.4f + 0.7f * pow_pos(scans, 1.7f);
}

// `rows`: the six source rows of the taps and the scanline factor when they depend on the row only (no curvature)
struct RowTaps { int r[6]; float scan; };
__device__ __forceinline__ void row_taps(const PresentParams &p, float uy, RowTaps &t) {
    t.r[0] = tap_row(p, uy + 0.0002f);
    t.r[1] = tap_row(p, uy - 0.0005f);
    t.r[2] = tap_row(p, uy + 0.0000f);
    t.r[3] = tap_row(p, -0.02025f + (uy + 0.0002f));
    t.r[4] = tap_row(p, -0.015f + (uy - 0.0005f));
    t.r[5] = tap_row(p, -0.0135f + (uy + 0.0000f));
    t.scan = scan_factor(p, uy);
}

__device__ __forceinline__ uint32_t crt_pixel(const PresentParams &p, const uint32_t *__restrict__ frame, const float *s_unorm,
                                              int X, int Y, const RowTaps *rows) {
    const float fx = (float)X + 0.5f, fy = (float)Y + 0.5f;
    float ux = fx / (float)p.ow, uy = fy / (float)p.oh;      // crt.kage:30-33
    RowTaps own;
    if (p.mode == 2) {                                       // crt.kage:18-27 curve(); pow(v, 2) = v * v
        ux = (ux - 0.5f) * 2.0f; uy = (uy - 0.5f) * 2.0f;
        ux *= 1.1f; uy *= 1.1f;
        const float ay = fabsf(uy) / 5.0f;
        ux *= 1.0f + ay * ay;
        const float ax = fabsf(ux) / 4.0f;
        uy *= 1.0f + ax * ax;
        ux = ux / 2.0f + 0.5f; uy = uy / 2.0f + 0.5f;
        ux = ux * 0.92f + 0.04f; uy = uy * 0.92f + 0.04f;
        if (ux < 0.0f || ux > 1.0f || uy < 0.0f || uy > 1.0f) return 0xff000000u;          // crt.kage:70-72
        row_taps(p, uy, own);
        rows = &own;
    }
    float col[3];                                            // crt.kage:41-48
    col[0] = tap(p, frame, s_unorm, tap_col(p, ux + 0.0002f), rows->r[0], 0) + 0.05f;
    col[1] = tap(p, frame, s_unorm, tap_col(p, ux + 0.0000f), rows->r[1], 1) + 0.05f;
    col[2] = tap(p, frame, s_unorm, tap_col(p, ux - 0.0005f), rows->r[2], 2) + 0.05f;
    col[0] += 0.08f * tap(p, frame, s_unorm, tap_col(p, 0.01875f + (ux + 0.0002f)), rows->r[3], 0);
    col[1] += 0.05f * tap(p, frame, s_unorm, tap_col(p, -0.0165f + (ux + 0.0000f)), rows->r[4], 1);
    col[2] += 0.08f * tap(p, frame, s_unorm, tap_col(p, -0.015f + (ux - 0.0005f)), rows->r[5], 2);
    const float vig = 16.0f * ux * uy * (1.0f - ux) * (1.0f - uy);                         // :54
    const float vp = pow_pos(vig, 0.3f);
    // :74  1 - 0.65 * clamp((mod(src.x, 2) - 1) * 2, 0, 1): src.x = X + 0.5, so the clamp is 0 on even, 1 on odd columns
    const float mask = (X & 1) ? 1.0f - 0.65f : 1.0f;
    uint32_t o = 0xff000000u;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float v = clampf01(col[c] * 0.6f + 0.4f * col[c] * col[c]);                        // :51
        v *= vp;
        v *= (c == 1 ? 1.05f : 0.95f);
        v *= 2.8f;
        v *= rows->scan;                                                                   // :60-62
        v *= mask;
        o |= (uint32_t)(int)floorf(clampf01(v) * 255.0f + 0.5f) << (8 * c);
    }
    return o;
}

// grid: x = 256-thread chunks of a row's pixel quads, y = output row, z = view
__global__ void __launch_bounds__(256) k_present(const PresentParams p) {
    __shared__ float s_unorm[256];
    if (p.mode != 0) {
        s_unorm[threadIdx.x] = (float)threadIdx.x / 255.0f;
        __syncthreads();
    }
    const int qx = blockIdx.x * blockDim.x + threadIdx.x;
    const int X0 = qx * 4;
    if (X0 >= p.ow) return;
    const int Y = blockIdx.y, view = blockIdx.z;
    const uint32_t *frame = reinterpret_cast<const uint32_t *>(p.src) + (size_t)view * p.sh * p.sw;
    uint32_t *orow = reinterpret_cast<uint32_t *>(p.out) + ((size_t)view * p.oh + Y) * p.ow;
    uint32_t px[4];
    if (p.mode == 0) {
        const int sy = __ldg(p.ymap + Y);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int X = X0 + j;
            const int sx = X < p.ow ? __ldg(p.xmap + X) : -1;
            px[j] = (sx >= 0 && sy >= 0) ? __ldg(frame + (size_t)sy * p.sw + sx) : 0u;
        }
    } else {
        RowTaps rows;
        if (p.mode == 1) row_taps(p, ((float)Y + 0.5f) / (float)p.oh, rows);
#pragma unroll
        for (int j = 0; j < 4; j++) px[j] = X0 + j < p.ow ? crt_pixel(p, frame, s_unorm, X0 + j, Y, &rows) : 0u;
    }
    if (X0 + 3 < p.ow && (p.ow & 3) == 0) {
        *reinterpret_cast<uint4 *>(orow + X0) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (X0 + j < p.ow) orow[X0 + j] = px[j];
    }
}

extern "C" int pm3d_set_present(pm3d_ctx *ctx, int out_w, int out_h, double render_scale) {
    if (!ctx) return PM3D_ERR_ARG;
    if (ctx->w < 1) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_set_view must precede pm3d_set_present");
    if (out_w < 1 || out_h < 1 || out_w > 32768 || out_h > 32768 || !(render_scale > 0.0) || !isfinite(render_scale))
        return pm3d_fail(ctx, PM3D_ERR_ARG, "bad present size / render scale");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> map((size_t)out_w + out_h);
    for (int i = 0; i < out_w; i++) {
        const double s = floor(((double)i + 0.5) * render_scale);
        map[i] = (s >= 0 && s < (double)ctx->w) ? (int32_t)s : -1;
    }
    for (int i = 0; i < out_h; i++) {
        const double s = floor(((double)i + 0.5) * render_scale);
        map[(size_t)out_w + i] = (s >= 0 && s < (double)ctx->h) ? (int32_t)s : -1;
    }
    if (ctx->present_map) PM3D_CUDA(ctx, cudaFree(ctx->present_map));
    ctx->present_map = nullptr;
    PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->present_map, map.size() * 4));
    PM3D_CUDA(ctx, cudaMemcpyAsync(ctx->present_map, map.data(), map.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->present_w = out_w; ctx->present_h = out_h;
    ctx->present_src_w = ctx->w; ctx->present_src_h = ctx->h;
    return PM3D_OK;
}

extern "C" int pm3d_present(pm3d_ctx *ctx, const uint8_t *d_src, int n_views, int mode, uint8_t *out, uint32_t flags) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->present_map || ctx->present_src_w != ctx->w || ctx->present_src_h != ctx->h)
        return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_set_present must follow pm3d_set_view");
    if (mode < 0 || mode > 2) return pm3d_fail(ctx, PM3D_ERR_ARG, "present mode %d not in 0..2", mode);
    if (!d_src) {
        d_src = ctx->last_frame;
        if (n_views < 1) n_views = ctx->last_views;
        int rc = pm3d_wait_last_frame(ctx);  // it may have run on another lane
        if (rc) return rc;
    }
    if (!d_src || n_views < 1 || !out) return pm3d_fail(ctx, PM3D_ERR_ARG, "nothing to present / null output");
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ob = (size_t)n_views * ctx->present_w * ctx->present_h * 4;
    uint8_t *d_out = out;
    if (!(flags & PM3D_OUT_DEVICE)) {
        if (ctx->present_bytes < ob) {
            if (ctx->present_buf) PM3D_CUDA(ctx, cudaFree(ctx->present_buf));
            ctx->present_buf = nullptr; ctx->present_bytes = 0;
            PM3D_CUDA(ctx, cudaMalloc((void **)&ctx->present_buf, ob));
            ctx->present_bytes = ob;
        }
        d_out = ctx->present_buf;
    }
    PresentParams p;
    p.src = d_src; p.out = d_out;
    p.xmap = ctx->present_map; p.ymap = ctx->present_map + ctx->present_w;
    p.sw = ctx->w; p.sh = ctx->h; p.ow = ctx->present_w; p.oh = ctx->present_h; p.mode = mode; p.n_views = n_views;
    p.n_quads = (long long)n_views * p.oh * ((p.ow + 3) / 4);
    if (p.oh > 65535 || n_views > 65535) return pm3d_fail(ctx, PM3D_ERR_CAPACITY, "present: too many rows / views for one launch");
    const dim3 grid((unsigned)(((p.ow + 3) / 4 + 255) / 256), (unsigned)p.oh, (unsigned)n_views);
    { int rc_ = pm3d_main_enter(ctx); if (rc_) return rc_; }
    k_present<<<grid, 256, 0, ctx->stream>>>(p);
    pm3d_main_touch(ctx);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    if (!(flags & PM3D_OUT_DEVICE)) PM3D_CUDA(ctx, cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
    if (!(flags & PM3D_NO_SYNC)) PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM3D_OK;
}
