// selfcheck.cu — device self-checks of arithmetic shortcuts the kernels rely on for bit-exactness.
//
// pm3d_selfcheck_division: div_by_rcp(a, b, RN(1/b)) (pm3d_internal.cuh; the floor cast's
// weight = currentDist / distWall) against the IEEE division on random and adversarial operands:
// uniform mantissas over a wide exponent range, divisors with all-ones / all-zero / single-bit mantissas,
// quotients engineered to sit next to rounding boundaries (a = b * m with m +- a few ulps, and
// a = RN(b * (m + 1/2 ulp))), and operands drawn from the path's own value range.
//
// pm3d_selfcheck_floor: floor_fast_index (pm3d_internal.cuh "fast floor cast") against the reference sequence
// int((weight*floorXWall + (1-weight)*posX) * texSize) on eligible operands — uniformly random ones and ones
// steered to within a few thousand ulps of an integer texel boundary, where the two roundings could disagree.
// Whenever the fast path accepts (non-zero rounded fraction, index below the bound) its index must equal the
// reference's and the reference coordinate must be non-negative.
#include "pm3d_internal.cuh"

__device__ __forceinline__ uint64_t sm64(uint64_t &s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double make_double(uint64_t mant52, int exp, bool neg) {
    const uint64_t bits = ((uint64_t)neg << 63) | ((uint64_t)(exp + 1023) << 52) | (mant52 & 0xFFFFFFFFFFFFFull);
    return __longlong_as_double((long long)bits);
}

__global__ void k_selfcheck_division(uint64_t seed, int iters, unsigned long long *mismatches, unsigned long long *tested) {
    uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
    unsigned long long bad = 0, n = 0;
    for (int it = 0; it < iters; it++) {
        const uint64_t r0 = sm64(s), r1 = sm64(s), r2 = sm64(s);
        const int mode = (int)(r2 & 7);
        double a, b;
        const int eb = (int)((r2 >> 8) % 61) - 30, ea = (int)((r2 >> 16) % 61) - 30;
        b = make_double(r1, eb, (r2 >> 30) & 1);
        a = make_double(r0, ea, (r2 >> 31) & 1);
        if (mode == 1) b = make_double(0xFFFFFFFFFFFFFull, eb, false);                  // all-ones divisor
        else if (mode == 2) b = make_double(0, eb, false);                              // power of two
        else if (mode == 3) b = make_double(1ull << ((r2 >> 40) % 52), eb, false);      // single mantissa bit
        else if (mode == 4 || mode == 5) {
            // a = b * m rounded, nudged by up to 3 ulps: quotients beside representable values / midpoints
            const double m = make_double(r0, ea, false);
            a = b * m;
            long long ab = __double_as_longlong(a) + (long long)((r2 >> 44) % 7) - 3;
            a = __longlong_as_double(ab);
        } else if (mode == 7) {
            // the whole exp_mid range: exponents up to +-400 on either side
            b = make_double(r1, (int)((r2 >> 8) % 801) - 400, (r2 >> 30) & 1);
            a = make_double(r0, (int)((r2 >> 20) % 801) - 400, (r2 >> 31) & 1);
        } else if (mode == 6) {
            // the path's own range: row distances 1e-3 .. 1e3, wall distances 1e-2 .. 1e3
            a = 1e-3 + (double)(r0 >> 11) * (1.0 / 9007199254740992.0) * 1e3;
            b = 1e-2 + (double)(r1 >> 11) * (1.0 / 9007199254740992.0) * 1e3;
        }
        if (!exp_mid(a) || !exp_mid(b)) continue;
        const double r = 1.0 / b;
        const double want = a / b;
        const double got = div_by_rcp(a, b, r);
        n++;
        if (__double_as_longlong(want) != __double_as_longlong(got)) bad++;
    }
    atomicAdd(mismatches, bad);
    atomicAdd(tested, n);
}

__global__ void k_selfcheck_floor(uint64_t seed, int iters, unsigned long long *counters) {
    uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
    unsigned long long bad = 0, n = 0, accepted = 0;
    const double U = 1.0 / 9007199254740992.0;
    for (int it = 0; it < iters; it++) {
        const uint64_t r0 = sm64(s), r1 = sm64(s), r2 = sm64(s), r3 = sm64(s), r4 = sm64(s);
        const int shift = 4 + (int)(r4 % 7);                        // texSize 16 .. 1024
        const double S = (double)(1 << shift);
        const double span = 16777216.0 / S;                          // coordinates up to 2^24 texels
        const double scale = (r4 & 0x100) ? span : ((r4 & 0x200) ? 512.0 : 8.0);
        const double posx = (double)(r0 >> 11) * U * scale;
        double fxw = (double)(r1 >> 11) * U * scale;
        if (r4 & 0x400) fxw = floor(fxw);                            // wall faces sit on integers
        const double dw = 1e-2 + (double)(r2 >> 11) * U * ((r4 & 0x800) ? 1e3 : 10.0);
        double cd = (double)(r3 >> 11) * U * 4.0 * dw;
        const double rc = 1.0 / dw;
        const double X1 = fxw * S, P = posx * S, D = X1 - P;
        if ((r4 >> 16) & 1) {
            // steer t = P + w*D next to an integer n: w = (n - P)/D, cd = w*dw, then nudge by up to 2^14 ulps
            const double t0 = P + (cd * rc) * D;
            const double nn = floor(t0 + 0.5);
            const double w = (nn - P) / D;
            if (w > 0.0 && w <= 3.999 && fabs(D) > 1e-6) {
                cd = w * dw;
                long long cb = __double_as_longlong(cd) + (long long)((r4 >> 20) % 32769) - 16384;
                cd = __longlong_as_double(cb);
            }
        }
        if (!(cd > 0.0) || !exp_mid(cd) || !(cd <= 4.0 * dw)) continue;
        // reference sequence (game path: no contraction)
        const double weight = cd / dw;
        const double fx = weight * fxw + (1.0 - weight) * posx;
        const double t = fx * S;
        uint32_t idx;
        const bool ok = floor_fast_index(cd * rc, D, P, 0x80000000u, idx);
        n++;
        if (ok) {
            accepted++;
            if (!(fx >= 0.0) || (uint32_t)(long long)t != idx) bad++;
        }
    }
    atomicAdd(counters, bad);
    atomicAdd(counters + 1, n);
    atomicAdd(counters + 2, accepted);
}

// pm3d_selfcheck_dda: dda_jump against the sequential merge it replaces.
__global__ void k_selfcheck_dda(uint64_t seed, int steps, unsigned long long *counters) {
    uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
    const double U = 1.0 / 9007199254740992.0;
    const uint64_t r0 = sm64(s), r1 = sm64(s), r2 = sm64(s), r3 = sm64(s), r4 = sm64(s);
    const int mode = (int)(r4 & 7);
    // position inside a cell, direction: as the cast builds them (cast_kernel.cuh)
    double fracX = (double)(r0 >> 11) * U, fracY = (double)(r1 >> 11) * U;
    double dirX = 2.0 * (double)(r2 >> 11) * U - 1.0, dirY = 2.0 * (double)(r3 >> 11) * U - 1.0;
    if (mode == 1) { dirY = dirX; }                                             // exact diagonal: ties everywhere
    else if (mode == 2) { dirY = dirX * (double)(1 + (r4 >> 8) % 7) / (double)(1 + (r4 >> 16) % 7); }  // rational slopes
    else if (mode == 3) { fracX = 0.5; fracY = 0.5; }                            // cell centre (the bundled poses)
    else if (mode == 4) { fracX = 0.0; fracY = (r4 & 0x100) ? 0.0 : 0.25; dirY = dirX * ((r4 & 0x200) ? 1.0 : 2.0); }
    else if (mode == 5) { dirX *= 1e-9; }                                        // almost axis-parallel
    else if (mode == 6) { dirX = (r4 & 0x100) ? 0.0 : dirX; dirY = (r4 & 0x200) ? 0.0 : dirY; }  // axis-parallel: inf / NaN
    const double pos = (double)((r4 >> 24) % 500);
    const double posX = pos + fracX, posY = (double)((r4 >> 40) % 500) + fracY;
    const long long mapX = (long long)posX, mapY = (long long)posY;
    const double dX = fabs(1.0 / dirX), dY = fabs(1.0 / dirY);
    const double sX0 = dirX < 0 ? (posX - (double)mapX) * dX : ((double)mapX + 1.0 - posX) * dX;
    const double sY0 = dirY < 0 ? (posY - (double)mapY) * dY : ((double)mapY + 1.0 - posY) * dY;
    const bool x_dead = !(sX0 < INFINITY) || sY0 != sY0;
    const bool y_dead = !x_dead && sY0 == INFINITY && dX < 1e290;
    const double c0 = sY0 - sX0, inv = 1.0 / (dX + dY);
    double a = sX0, b = sY0;
    int i = 0, j = 0, side = -1;
    unsigned long long bad = 0, n = 0, cert = 0;
    int next_check = 1 + (int)(sm64(s) % 4);
    for (int step = 1; step <= steps; step++) {
        if (a < b) { a += dX; i++; side = 0; } else { b += dY; j++; side = 1; }
        if (step == next_check) {
            next_check += 1 + (int)(sm64(s) % 37);
            int gi = 0, gj = step, gs = 1;
            bool ok = true;
            if (x_dead) { gi = 0; gj = step; gs = 1; }
            else if (y_dead) { gi = step; gj = 0; gs = 0; }
            else ok = dda_jump(sX0, sY0, dX, dY, c0, inv, step, gi, gj, gs);
            n++;
            if (ok) {
                cert++;
                if (gi != i || gj != j || gs != side) bad++;
            }
        }
    }
    atomicAdd(counters, bad);
    atomicAdd(counters + 1, n);
    atomicAdd(counters + 2, cert);
}

extern "C" int pm3d_selfcheck_dda(pm3d_ctx *ctx, uint64_t seed, int blocks, int steps, uint64_t *out_mismatches,
                                  uint64_t *out_tested, uint64_t *out_certified) {
    if (!ctx || blocks < 1 || steps < 1 || !out_mismatches || !out_tested || !out_certified) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    PM3D_CUDA(ctx, cudaMalloc((void **)&d, 24));
    PM3D_CUDA(ctx, cudaMemsetAsync(d, 0, 24, ctx->stream));
    k_selfcheck_dda<<<blocks, 256, 0, ctx->stream>>>(seed, steps, d);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    unsigned long long h[3] = {0, 0, 0};
    PM3D_CUDA(ctx, cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PM3D_CUDA(ctx, cudaFree(d));
    *out_mismatches = h[0];
    *out_tested = h[1];
    *out_certified = h[2];
    return PM3D_OK;
}

extern "C" int pm3d_selfcheck_floor(pm3d_ctx *ctx, uint64_t seed, int blocks, int iters, uint64_t *out_mismatches,
                                    uint64_t *out_tested, uint64_t *out_accepted) {
    if (!ctx || blocks < 1 || iters < 1 || !out_mismatches || !out_tested || !out_accepted) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    PM3D_CUDA(ctx, cudaMalloc((void **)&d, 24));
    PM3D_CUDA(ctx, cudaMemsetAsync(d, 0, 24, ctx->stream));
    k_selfcheck_floor<<<blocks, 256, 0, ctx->stream>>>(seed, iters, d);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    unsigned long long h[3] = {0, 0, 0};
    PM3D_CUDA(ctx, cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PM3D_CUDA(ctx, cudaFree(d));
    *out_mismatches = h[0];
    *out_tested = h[1];
    *out_accepted = h[2];
    return PM3D_OK;
}

extern "C" int pm3d_selfcheck_division(pm3d_ctx *ctx, uint64_t seed, int blocks, int iters, uint64_t *out_mismatches,
                                       uint64_t *out_tested) {
    if (!ctx || blocks < 1 || iters < 1 || !out_mismatches || !out_tested) return PM3D_ERR_ARG;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    PM3D_CUDA(ctx, cudaMalloc((void **)&d, 16));
    PM3D_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    k_selfcheck_division<<<blocks, 256, 0, ctx->stream>>>(seed, iters, d, d + 1);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    PM3D_CUDA(ctx, cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PM3D_CUDA(ctx, cudaFree(d));
    *out_mismatches = h[0];
    *out_tested = h[1];
    return PM3D_OK;
}
