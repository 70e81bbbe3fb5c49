// Store-pattern microbenchmark (tools only): how fast can 4096 frames of 320x200 RGBA8 (1.05 GB) be WRITTEN with the
// compose kernel's store pattern (a warp owns a 64-column x 16-row strip: 256 B pieces at a 1280 B pitch) compared with
// whole rows and with fully contiguous warps?   nvcc -arch=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_strip(uint2 *out, int w, int h, int n_views, unsigned *ctr, unsigned n_tasks, int rows_per_step) {
    const int lane = threadIdx.x & 31;
    const int n_groups = w / 64, n_yb = (h + 15) / 16;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned t = gw; t < n_tasks; t += nw) {
        const int view = t / (n_groups * n_yb), r = t % (n_groups * n_yb), yb = r / n_groups, g = r % n_groups;
        const int y0 = yb * 16, y1 = min(y0 + 16, h);
        uint2 *o = out + (((size_t)view * h + y0) * w + g * 64) / 2 + lane;
        for (int y = y0; y < y1; y++) o[(size_t)(y - y0) * (w / 2)] = make_uint2(t, y);
    }
}
// a warp writes whole rows: 16 rows x w pixels contiguous (w*4*16 bytes)
__global__ void k_rows(uint4 *out, int w, int h, int n_views, unsigned *ctr, unsigned n_tasks) {
    const int lane = threadIdx.x & 31;
    const int n_yb = (h + 15) / 16;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned t = gw; t < n_tasks; t += nw) {
        const int view = t / n_yb, yb = t % n_yb;
        const int y0 = yb * 16, y1 = min(y0 + 16, h);
        uint4 *o = out + (((size_t)view * h + y0) * w) / 4;
        const int n = (y1 - y0) * w / 4;
        for (int i = lane; i < n; i += 32) o[i] = make_uint4(t, i, 0, 0);
    }
}
// a warp owns full-width row blocks and writes them the way the compose runs would: per step one row pair, 64-column
// group by group (STG.64 per lane: 256 B per instruction, a row pair = 2 * w * 4 contiguous bytes)
__global__ void k_rowpairs(uint2 *out, int w, int h, int n_views, unsigned *ctr, unsigned n_tasks, int rows) {
    const int lane = threadIdx.x & 31;
    const int n_groups = w / 64, n_yb = (h + rows - 1) / rows;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned t = gw; t < n_tasks; t += nw) {
        const int view = t / n_yb, yb = t % n_yb;
        const int y0 = yb * rows, y1 = min(y0 + rows, h);
        uint2 *o = out + (((size_t)view * h + y0) * w) / 2 + lane;
        for (int y = y0; y < y1; y++)
            for (int g = 0; g < n_groups; g++) o[(size_t)(y - y0) * (w / 2) + g * 32] = make_uint2(t, y);
    }
}
// strips, but group-major inside one warp: the warp walks group 0's 16 rows, then group 1's, ... (full-width task)
__global__ void k_groupmajor(uint2 *out, int w, int h, int n_views, unsigned *ctr, unsigned n_tasks) {
    const int lane = threadIdx.x & 31;
    const int n_groups = w / 64, n_yb = (h + 15) / 16;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned t = gw; t < n_tasks; t += nw) {
        const int view = t / n_yb, yb = t % n_yb;
        const int y0 = yb * 16, y1 = min(y0 + 16, h);
        uint2 *o = out + (((size_t)view * h + y0) * w) / 2 + lane;
        for (int g = 0; g < n_groups; g++)
            for (int y = y0; y < y1; y++) o[(size_t)(y - y0) * (w / 2) + g * 32] = make_uint2(t, y);
    }
}
int main() {
    const int w = 320, h = 200, nv = 4096;
    const size_t bytes = (size_t)nv * w * h * 4;
    void *out; unsigned *ctr;
    cudaMalloc(&out, bytes); cudaMalloc(&ctr, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int mode = 0; mode < 8; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 6; rep++) {
            cudaMemsetAsync(ctr, 0, 4);
            cudaEventRecord(e0);
            if (mode == 0) k_strip<<<sm * 8, 128>>>((uint2 *)out, w, h, nv, ctr, nv * (w / 64) * ((h + 15) / 16), 2);
            else if (mode == 1) k_rows<<<sm * 8, 128>>>((uint4 *)out, w, h, nv, ctr, nv * ((h + 15) / 16));
            else if (mode == 2) cudaMemsetAsync(out, 1, bytes);
            else if (mode == 3) k_rowpairs<<<sm * 8, 128>>>((uint2 *)out, w, h, nv, ctr, nv * ((h + 15) / 16), 16);
            else if (mode == 4) k_rowpairs<<<sm * 8, 128>>>((uint2 *)out, w, h, nv, ctr, nv * ((h + 3) / 4), 4);
            else if (mode == 5) k_groupmajor<<<sm * 8, 128>>>((uint2 *)out, w, h, nv, ctr, nv * ((h + 15) / 16));
            else if (mode == 6) k_strip<<<sm * 2, 128>>>((uint2 *)out, w, h, nv, ctr, nv * (w / 64) * ((h + 15) / 16), 2);
            else k_rowpairs<<<sm * 8, 128>>>((uint2 *)out, w, h, nv, ctr, nv * ((h + 1) / 2), 2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%s: %.3f ms = %.0f GB/s\n", mode == 0 ? "strips 64x16 (compose pattern)" : mode == 1 ? "whole rows, 16 per warp" : mode == 2 ? "cudaMemset" : mode == 3 ? "full-width, 16 rows per task, row by row, STG.64 per group" : mode == 4 ? "full-width, 4 rows per task" : mode == 5 ? "full-width task, group-major inside" : mode == 6 ? "strips, 8 warps per SM" : "full-width, 2 rows per task", best, bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
