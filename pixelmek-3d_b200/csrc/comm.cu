// comm.cu — multi-GPU side of the C-ABI: the path shards without an exchange step — camera views
// round-robin, LOS rays in contiguous ranges, mission data replicated — and the only communication is the gather of
// the COMPACT per-unit results (convergence + checksum per view, one byte per ray) to every rank over NVLink.
//
// The library owns the NCCL communicator so that a Go host needs neither an NCCL nor a torch binding:
//   * one process per GPU:  pm3d_comm_unique_id on rank 0, the 128 bytes travel by whatever means the host has,
//     pm3d_comm_init(ctx, id, rank, n) on every rank;
//   * one process, n GPUs:  pm3d_group_create(devices, n) = one ctx per device + ncclCommInitAll.
// libnccl is not a link dependency: it is dlopen'ed on first use (libnccl.so.2; PM3D_NCCL_LIB overrides), so a host
// without NCCL still renders on one GPU and every pm3d_comm_* / pm3d_group_* call fails with PM3D_ERR_NCCL.
// Every collective is enqueued on the ctx stream, after the kernels whose results it carries: no host round trip
// between the compute and the gather.
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "pm3d_internal.cuh"

// ---- the slice of the NCCL 2.x ABI this file uses (declared here: nccl.h is not needed to build)
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess = 0
}
static_assert(sizeof(ncclUniqueId) == PM3D_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
namespace {
constexpr int NCCL_UINT8 = 1;  // ncclDataType_t::ncclUint8

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *names[] = {getenv("PM3D_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
        api.error = dlerror();
    }
    if (!api.handle) return &api;
    bool ok = true;
    auto sym = [&](const char *name) -> void * {
        void *p = dlsym(api.handle, name);
        if (!p) { ok = false; api.error = std::string("libnccl lacks ") + name; }
        return p;
    };
    *(void **)&api.GetUniqueId = sym("ncclGetUniqueId");
    *(void **)&api.CommInitRank = sym("ncclCommInitRank");
    *(void **)&api.CommInitAll = sym("ncclCommInitAll");
    *(void **)&api.CommDestroy = sym("ncclCommDestroy");
    *(void **)&api.AllGather = sym("ncclAllGather");
    *(void **)&api.Broadcast = sym("ncclBroadcast");
    *(void **)&api.GroupStart = sym("ncclGroupStart");
    *(void **)&api.GroupEnd = sym("ncclGroupEnd");
    *(void **)&api.GetErrorString = sym("ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    return &api;
}

int need_nccl(pm3d_ctx *ctx, NcclApi **out) {
    NcclApi *a = nccl_api();
    if (!a->handle) return pm3d_fail(ctx, PM3D_ERR_NCCL, "NCCL is not available (%s); multi-GPU calls need libnccl.so.2", a->error.c_str());
    *out = a;
    return PM3D_OK;
}
#define PM3D_NCCL(ctx, api, call)                                                                              \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != 0) return pm3d_fail((ctx), PM3D_ERR_NCCL, "%s: %s", #call, (api)->GetErrorString(r__));     \
    } while (0)

int ensure_buf(pm3d_ctx *ctx, void **buf, size_t *have, size_t need) {
    if (*have >= need) return PM3D_OK;
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*buf) PM3D_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr; *have = 0;
    PM3D_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return PM3D_OK;
}

// One gathered record per view: what a host needs back from a headless unit camera.
struct __align__(8) ViewResult {
    pm3d_convergence conv;  // 32 B
    uint64_t checksum;
};
static_assert(sizeof(ViewResult) == 40, "ViewResult");

__global__ void k_pack_views(const pm3d_convergence *conv, const uint64_t *sums, ViewResult *send, int n_local, int per) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    ViewResult r;
    memset(&r, 0, sizeof r);
    r.conv.sprite = -1;
    if (i < n_local) { r.conv = conv[i]; r.checksum = sums ? sums[i] : 0ull; }
    send[i] = r;
}
// recv is rank-major [n_ranks][per]; view i of the batch is slot i / n_ranks of rank i % n_ranks
__global__ void k_unshard_views(const ViewResult *recv, pm3d_convergence *all_conv, uint64_t *all_sums, int n_views, int n_ranks, int per) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_views) return;
    const ViewResult r = recv[(size_t)(i % n_ranks) * per + i / n_ranks];
    if (all_conv) all_conv[i] = r.conv;
    if (all_sums) all_sums[i] = r.checksum;
}

// ---- LOS answers travel as BITS: a rank's byte-per-ray answers are packed 32 to a word, the words are all-gathered
// (1/8 of the bytes: 2 MiB instead of 16 MiB for the c5 batch, where the collective is a third of an 8-GPU batch) and
// every rank expands the other ranks' words back into the caller's byte array (its own range is already in place).
__global__ void k_pack_bits(const uint8_t *bytes, uint32_t *words, int64_t n_words) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint4 a = reinterpret_cast<const uint4 *>(bytes)[2 * i], b = reinterpret_cast<const uint4 *>(bytes)[2 * i + 1];
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w = v[k];  // four answers, one per byte (any non-zero byte reads as "visible")
        const uint32_t nib = ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 2u : 0u) | ((w & 0xff0000u) ? 4u : 0u) | ((w & 0xff000000u) ? 8u : 0u);
        m |= nib << (4 * k);
    }
    words[i] = m;
}
__global__ void k_unpack_bits(const uint32_t *words, uint8_t *bytes, int64_t n_words, int64_t own_lo, int64_t own_hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words || (i >= own_lo && i < own_hi)) return;
    const uint32_t m = words[i];
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t nib = (m >> (4 * k)) & 15u;
        v[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
    }
    reinterpret_cast<uint4 *>(bytes)[2 * i] = make_uint4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<uint4 *>(bytes)[2 * i + 1] = make_uint4(v[4], v[5], v[6], v[7]);
}
// equal shards of whole 32-ray words in a 16 B aligned array, more than one rank
// (PM3D_COMM_FORCE_BITS=1, a test override: a single rank packs, gathers and unpacks its own range too, so that one GPU
//  exercises both kernels)
static bool force_bits() { const char *e = getenv("PM3D_COMM_FORCE_BITS"); return e && atoi(e) != 0; }
bool rays_packable(const pm3d_ctx *ctx, const uint8_t *d_all, int64_t n_global) {
    const int n = ctx->comm_size;
    return (n > 1 || force_bits()) && n_global % n == 0 && (n_global / n) % 32 == 0 && ((uintptr_t)d_all & 15u) == 0;
}
int rays_gather_packed(pm3d_ctx *ctx, NcclApi *api, uint8_t *d_all, int64_t n_global) {
    const int n = ctx->comm_size;
    const int64_t cnt = n_global / n, words = cnt / 32, all_words = n_global / 32;
    int rc;
    if ((rc = ensure_buf(ctx, &ctx->comm_send, &ctx->comm_send_bytes, (size_t)words * 4))) return rc;
    if ((rc = ensure_buf(ctx, &ctx->comm_recv, &ctx->comm_recv_bytes, (size_t)all_words * 4))) return rc;
    k_pack_bits<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(d_all + cnt * ctx->comm_rank, (uint32_t *)ctx->comm_send, words);
    PM3D_NCCL(ctx, api, api->AllGather(ctx->comm_send, ctx->comm_recv, (size_t)words * 4, NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream));
    const bool own_too = n == 1;  // (forced single rank: nothing else to unpack)
    k_unpack_bits<<<(unsigned)((all_words + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)ctx->comm_recv, d_all, all_words, own_too ? 0 : words * ctx->comm_rank, own_too ? 0 : words * (ctx->comm_rank + 1));
    ctx->launches += 2;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}

// ---- the three phases of a sharded view batch on one ctx (the group calls run each phase over all its devices)
int views_pre(pm3d_ctx *ctx, const pm3d_view *d_views, int n_local, const pm3d_sprite *d_sprites, const int32_t *d_off,
              int n_sprites_total, uint8_t *d_out_rgba, int n_views_global, bool want_sums, bool rendered_already) {
    const int n = ctx->comm_size, per = (n_views_global + n - 1) / n;
    const int expect = n_views_global / n + (ctx->comm_rank < n_views_global % n ? 1 : 0);
    if (n_local != expect) return pm3d_fail(ctx, PM3D_ERR_ARG, "rank %d of %d owns %d of %d views (i %% n), got %d", ctx->comm_rank, n, expect, n_views_global, n_local);
    int rc;
    if (!rendered_already && n_local > 0 && (rc = pm3d_render_resident(ctx, d_views, n_local, d_sprites, d_off, n_sprites_total, d_out_rgba))) return rc;
    if ((rc = pm3d_main_enter(ctx))) return rc;
    if (want_sums && n_local > 0 && (rc = pm3d_launch_checksum(ctx, n_local, ctx->last_frame))) return rc;
    if ((rc = ensure_buf(ctx, &ctx->comm_send, &ctx->comm_send_bytes, (size_t)per * sizeof(ViewResult)))) return rc;
    if ((rc = ensure_buf(ctx, &ctx->comm_recv, &ctx->comm_recv_bytes, (size_t)per * n * sizeof(ViewResult)))) return rc;
    k_pack_views<<<(per + 127) / 128, 128, 0, ctx->stream>>>(ctx->conv, want_sums ? ctx->checksums : nullptr, (ViewResult *)ctx->comm_send, n_local, per);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    return PM3D_OK;
}
int views_gather(pm3d_ctx *ctx, NcclApi *api, int n_views_global) {
    const int per = (n_views_global + ctx->comm_size - 1) / ctx->comm_size;
    PM3D_NCCL(ctx, api, api->AllGather(ctx->comm_send, ctx->comm_recv, (size_t)per * sizeof(ViewResult), NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream));
    return PM3D_OK;
}
int views_post(pm3d_ctx *ctx, int n_views_global, pm3d_convergence *d_all_conv, uint64_t *d_all_sums) {
    const int per = (n_views_global + ctx->comm_size - 1) / ctx->comm_size;
    k_unshard_views<<<(n_views_global + 127) / 128, 128, 0, ctx->stream>>>((const ViewResult *)ctx->comm_recv, d_all_conv, d_all_sums, n_views_global, ctx->comm_size, per);
    ctx->launches++;
    PM3D_CUDA(ctx, cudaGetLastError());
    // (no pm3d_main_touch: checksum, pack, gather and unshard only CONSUME a frame batch — nothing they write is an
    //  input of a frame, so the front launch of the next batch need not wait for them and runs beside this batch's
    //  compose launch and gather; the set-reuse events still keep the conv / checksum buffers intact until read)
    return PM3D_OK;
}

void ray_range(int64_t n, int rank, int n_ranks, int64_t *lo, int64_t *hi) {
    const int64_t base = n / n_ranks, rem = n % n_ranks;
    *lo = rank * base + std::min<int64_t>(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}
// every rank's contiguous byte range of d_all travels to every other rank: one in-place all-gather when the ranges are
// equal, else one broadcast per range inside one NCCL group
int rays_gather(pm3d_ctx *ctx, NcclApi *api, uint8_t *d_all, int64_t n_global, bool own_group) {
    const int n = ctx->comm_size;
    if (n_global % n == 0) {
        const size_t cnt = (size_t)(n_global / n);
        PM3D_NCCL(ctx, api, api->AllGather(d_all + cnt * ctx->comm_rank, d_all, cnt, NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream));
        return PM3D_OK;
    }
    if (own_group) PM3D_NCCL(ctx, api, api->GroupStart());
    for (int r = 0; r < n; r++) {
        int64_t lo, hi;
        ray_range(n_global, r, n, &lo, &hi);
        if (hi > lo) PM3D_NCCL(ctx, api, api->Broadcast(d_all + lo, d_all + lo, (size_t)(hi - lo), NCCL_UINT8, r, (ncclComm_t)ctx->comm, ctx->stream));
    }
    if (own_group) PM3D_NCCL(ctx, api, api->GroupEnd());
    return PM3D_OK;
}

}  // namespace

struct pm3d_group {
    int n = 0;
    pm3d_ctx *ctx[PM3D_GROUP_MAX] = {};
    std::string err;
    // per-device staging of the host-facing group calls
    void *d_in[PM3D_GROUP_MAX] = {}, *d_out[PM3D_GROUP_MAX] = {};
    size_t in_bytes[PM3D_GROUP_MAX] = {}, out_bytes[PM3D_GROUP_MAX] = {};
};

static thread_local std::string g_group_error;
static int group_fail(pm3d_group *g, int code, const std::string &msg) {
    if (g) g->err = msg; else g_group_error = msg;
    return code;
}
static int group_fail_ctx(pm3d_group *g, int i, int code) {
    return group_fail(g, code, "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err);
}

extern "C" {

int pm3d_comm_unique_id(uint8_t *id128) {
    NcclApi *api = nullptr;
    if (!id128) return PM3D_ERR_ARG;
    int rc = need_nccl(nullptr, &api);
    if (rc) return rc;
    ncclUniqueId id;
    PM3D_NCCL(nullptr, api, api->GetUniqueId(&id));
    memcpy(id128, id.internal, PM3D_COMM_ID_BYTES);
    return PM3D_OK;
}

int pm3d_comm_init(pm3d_ctx *ctx, const uint8_t *id128, int rank, int n_ranks) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_comm_init: bad rank %d of %d", rank, n_ranks);
    if (ctx->comm) return pm3d_fail(ctx, PM3D_ERR_ARG, "the ctx already belongs to a communicator");
    NcclApi *api = nullptr;
    int rc = need_nccl(ctx, &api);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, PM3D_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    PM3D_NCCL(ctx, api, api->CommInitRank(&comm, n_ranks, id, rank));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_size = n_ranks; ctx->comm_owned = true;
    return PM3D_OK;
}

int pm3d_comm_destroy(pm3d_ctx *ctx) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->comm) return PM3D_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    NcclApi *api = nccl_api();
    if (ctx->comm_owned && api->handle) api->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_size = 1; ctx->comm_owned = false;
    if (ctx->comm_send) cudaFree(ctx->comm_send);
    if (ctx->comm_recv) cudaFree(ctx->comm_recv);
    ctx->comm_send = ctx->comm_recv = nullptr;
    ctx->comm_send_bytes = ctx->comm_recv_bytes = 0;
    return PM3D_OK;
}

int pm3d_comm_rank(const pm3d_ctx *ctx) { return ctx ? ctx->comm_rank : -1; }
int pm3d_comm_size(const pm3d_ctx *ctx) { return ctx ? ctx->comm_size : 0; }

int64_t pm3d_shard_view_count(int64_t n_views, int rank, int n_ranks) {
    if (n_views < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return -1;
    return n_views / n_ranks + (rank < n_views % n_ranks ? 1 : 0);
}
int pm3d_shard_ray_range(int64_t n_rays, int rank, int n_ranks, int64_t *lo, int64_t *hi) {
    if (n_rays < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !lo || !hi) return PM3D_ERR_ARG;
    ray_range(n_rays, rank, n_ranks, lo, hi);
    return PM3D_OK;
}

int pm3d_comm_all_gather(pm3d_ctx *ctx, const void *d_send, void *d_recv, int64_t bytes_per_rank) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->comm) return pm3d_fail(ctx, PM3D_ERR_NCCL, "pm3d_comm_init must be called first");
    if (!d_send || !d_recv || bytes_per_rank < 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_comm_all_gather: null buffer");
    NcclApi *api = nullptr;
    int rc = need_nccl(ctx, &api);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = pm3d_main_enter(ctx))) return rc;
    PM3D_NCCL(ctx, api, api->AllGather(d_send, d_recv, (size_t)bytes_per_rank, NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream));
    pm3d_main_touch(ctx);
    return PM3D_OK;
}

int pm3d_comm_render_views(pm3d_ctx *ctx, const pm3d_view *d_views, int n_local, const pm3d_sprite *d_sprites,
                           const int32_t *d_sprite_offset, int n_sprites_total, uint8_t *d_out_rgba, int n_views_global,
                           pm3d_convergence *d_all_conv, uint64_t *d_all_checksums) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->comm) return pm3d_fail(ctx, PM3D_ERR_NCCL, "pm3d_comm_init must be called first");
    if (n_views_global < 1 || n_local < 0) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_comm_render_views: no views");
    NcclApi *api = nullptr;
    int rc = need_nccl(ctx, &api);
    if (rc) return rc;
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = views_pre(ctx, d_views, n_local, d_sprites, d_sprite_offset, n_sprites_total, d_out_rgba, n_views_global, d_all_checksums != nullptr, false))) return rc;
    if ((rc = views_gather(ctx, api, n_views_global))) return rc;
    return views_post(ctx, n_views_global, d_all_conv, d_all_checksums);
}

int pm3d_comm_los_batch(pm3d_ctx *ctx, const double *d_rays, int64_t n_local, int64_t n_global, uint8_t *d_all_out) {
    if (!ctx) return PM3D_ERR_ARG;
    if (!ctx->comm) return pm3d_fail(ctx, PM3D_ERR_NCCL, "pm3d_comm_init must be called first");
    if (n_global < 1 || !d_all_out) return pm3d_fail(ctx, PM3D_ERR_ARG, "pm3d_comm_los_batch: no rays / null output");
    int64_t lo, hi;
    ray_range(n_global, ctx->comm_rank, ctx->comm_size, &lo, &hi);
    if (n_local != hi - lo) return pm3d_fail(ctx, PM3D_ERR_ARG, "rank %d of %d owns rays [%lld, %lld) of %lld, got %lld", ctx->comm_rank, ctx->comm_size, (long long)lo, (long long)hi, (long long)n_global, (long long)n_local);
    NcclApi *api = nullptr;
    int rc = need_nccl(ctx, &api);
    if (rc) return rc;
    if (n_local > 0 && (rc = pm3d_los_batch(ctx, d_rays, n_local, d_all_out + lo, PM3D_IN_DEVICE | PM3D_OUT_DEVICE | PM3D_NO_SYNC))) return rc;
    if (rays_packable(ctx, d_all_out, n_global)) rc = rays_gather_packed(ctx, api, d_all_out, n_global);
    else rc = rays_gather(ctx, api, d_all_out, n_global, true);
    if (rc) return rc;
    pm3d_main_touch(ctx);
    return PM3D_OK;
}

// ------------------------------------------------------------------ one process, n GPUs

const char *pm3d_group_last_error(const pm3d_group *g) { return g ? g->err.c_str() : g_group_error.c_str(); }

int pm3d_group_create(const int *devices, int n, pm3d_group **out) {
    if (!out) return PM3D_ERR_ARG;
    *out = nullptr;
    if (!devices || n < 1 || n > PM3D_GROUP_MAX) return group_fail(nullptr, PM3D_ERR_ARG, "pm3d_group_create: 1..8 devices");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return group_fail(nullptr, PM3D_ERR_ARG, "pm3d_group_create: a device is listed twice");
    NcclApi *api = nccl_api();
    if (!api->handle) return group_fail(nullptr, PM3D_ERR_NCCL, "NCCL is not available (" + api->error + ")");
    pm3d_group *g = new pm3d_group();
    g->n = n;
    for (int i = 0; i < n; i++) {
        int rc = pm3d_create(devices[i], nullptr, &g->ctx[i]);
        if (rc) {
            std::string msg = pm3d_last_error(nullptr);
            pm3d_group_destroy(g);
            return group_fail(nullptr, rc, msg);
        }
    }
    ncclComm_t comms[PM3D_GROUP_MAX];
    ncclResult_t r = api->CommInitAll(comms, n, devices);
    if (r != 0) {
        std::string msg = std::string("ncclCommInitAll: ") + api->GetErrorString(r);
        pm3d_group_destroy(g);
        return group_fail(nullptr, PM3D_ERR_NCCL, msg);
    }
    for (int i = 0; i < n; i++) {
        g->ctx[i]->comm = comms[i]; g->ctx[i]->comm_rank = i; g->ctx[i]->comm_size = n; g->ctx[i]->comm_owned = true;
    }
    *out = g;
    return PM3D_OK;
}

void pm3d_group_destroy(pm3d_group *g) {
    if (!g) return;
    for (int i = 0; i < g->n; i++) {
        if (!g->ctx[i]) continue;
        cudaSetDevice(g->ctx[i]->device);
        if (g->d_in[i]) cudaFree(g->d_in[i]);
        if (g->d_out[i]) cudaFree(g->d_out[i]);
        pm3d_destroy(g->ctx[i]);  // destroys its communicator too
    }
    delete g;
}

int pm3d_group_size(const pm3d_group *g) { return g ? g->n : 0; }
pm3d_ctx *pm3d_group_ctx(pm3d_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }

int pm3d_group_sync(pm3d_group *g) {
    if (!g) return PM3D_ERR_ARG;
    int first = PM3D_OK;
    for (int i = 0; i < g->n; i++) {
        const int rc = pm3d_sync(g->ctx[i]);
        if (rc && !first) first = group_fail_ctx(g, i, rc);
    }
    return first;
}

static int group_stage(pm3d_group *g, int i, void **buf, size_t *have, size_t need) {
    if (*have >= need) return PM3D_OK;
    pm3d_ctx *ctx = g->ctx[i];
    PM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    PM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*buf) PM3D_CUDA(ctx, cudaFree(*buf));
    *buf = nullptr; *have = 0;
    PM3D_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return PM3D_OK;
}

int pm3d_group_render_views(pm3d_group *g, const pm3d_view *views, int n_views, const pm3d_sprite *sprites,
                            const int32_t *sprite_offset, pm3d_convergence *out_conv, uint64_t *out_checksums) {
    if (!g) return PM3D_ERR_ARG;
    if (!views || n_views < 1 || (!out_conv && !out_checksums)) return group_fail(g, PM3D_ERR_ARG, "pm3d_group_render_views: no views / no output");
    NcclApi *api = nccl_api();
    const int n = g->n;
    int rc;
    // shard on the host: device r gets views r, r + n, r + 2n, ... with their sprite lists, in that order
    for (int r = 0; r < n; r++) {
        std::vector<pm3d_view> v;
        std::vector<pm3d_sprite> s;
        std::vector<int32_t> off(1, 0);
        for (int i = r; i < n_views; i += n) {
            v.push_back(views[i]);
            if (sprite_offset) {
                if (sprite_offset[i + 1] < sprite_offset[i]) return group_fail(g, PM3D_ERR_ARG, "sprite_offset not monotone");
                s.insert(s.end(), sprites + sprite_offset[i], sprites + sprite_offset[i + 1]);
            }
            off.push_back((int32_t)s.size());
        }
        pm3d_ctx *ctx = g->ctx[r];
        if (!v.empty()) {
            // host arrays in (copied to the ctx's pinned staging before the call returns), the frames stay in the device's
            // own frame buffer; enqueue only, so the devices render side by side
            if ((rc = pm3d_render(ctx, v.data(), (int)v.size(), s.empty() ? nullptr : s.data(), off.data(), nullptr, nullptr, nullptr, PM3D_NO_SYNC)))
                return group_fail_ctx(g, r, rc);
        }
        cudaSetDevice(ctx->device);
        if ((rc = views_pre(ctx, nullptr, (int)v.size(), nullptr, nullptr, 0, nullptr, n_views, out_checksums != nullptr, true))) return group_fail_ctx(g, r, rc);
    }
    if (api->GroupStart() != 0) return group_fail(g, PM3D_ERR_NCCL, "ncclGroupStart failed");
    for (int r = 0; r < n; r++) {
        cudaSetDevice(g->ctx[r]->device);
        if ((rc = views_gather(g->ctx[r], api, n_views))) { api->GroupEnd(); return group_fail_ctx(g, r, rc); }
    }
    { const ncclResult_t e = api->GroupEnd(); if (e != 0) return group_fail(g, PM3D_ERR_NCCL, std::string("ncclGroupEnd: ") + api->GetErrorString(e)); }
    // every device now holds every view's record; the host reads device 0's copy
    const size_t cb = (size_t)n_views * sizeof(pm3d_convergence), sb = (size_t)n_views * 8;
    for (int r = 0; r < n; r++) {
        pm3d_ctx *ctx = g->ctx[r];
        cudaSetDevice(ctx->device);
        if ((rc = group_stage(g, r, &g->d_out[r], &g->out_bytes[r], cb + sb))) return group_fail_ctx(g, r, rc);
        if ((rc = views_post(ctx, n_views, (pm3d_convergence *)g->d_out[r], (uint64_t *)((uint8_t *)g->d_out[r] + cb)))) return group_fail_ctx(g, r, rc);
    }
    pm3d_ctx *c0 = g->ctx[0];
    cudaSetDevice(c0->device);
    if (out_conv && cudaMemcpyAsync(out_conv, g->d_out[0], cb, cudaMemcpyDeviceToHost, c0->stream) != cudaSuccess) return group_fail(g, PM3D_ERR_CUDA, "D2H of the gathered convergence records failed");
    if (out_checksums && cudaMemcpyAsync(out_checksums, (uint8_t *)g->d_out[0] + cb, sb, cudaMemcpyDeviceToHost, c0->stream) != cudaSuccess) return group_fail(g, PM3D_ERR_CUDA, "D2H of the gathered checksums failed");
    return pm3d_group_sync(g);
}

int pm3d_group_los_batch(pm3d_group *g, const double *rays, int64_t n_rays, uint8_t *out) {
    if (!g) return PM3D_ERR_ARG;
    if (!rays || !out || n_rays < 1) return group_fail(g, PM3D_ERR_ARG, "pm3d_group_los_batch: no rays / null output");
    NcclApi *api = nccl_api();
    const int n = g->n;
    int rc;
    for (int r = 0; r < n; r++) {
        pm3d_ctx *ctx = g->ctx[r];
        int64_t lo, hi;
        ray_range(n_rays, r, n, &lo, &hi);
        cudaSetDevice(ctx->device);
        if ((rc = group_stage(g, r, &g->d_in[r], &g->in_bytes[r], (size_t)std::max<int64_t>(hi - lo, 1) * 32))) return group_fail_ctx(g, r, rc);
        if ((rc = group_stage(g, r, &g->d_out[r], &g->out_bytes[r], (size_t)n_rays))) return group_fail_ctx(g, r, rc);
        if (hi > lo) {
            if (cudaMemcpyAsync(g->d_in[r], rays + lo * 4, (size_t)(hi - lo) * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return group_fail(g, PM3D_ERR_CUDA, "H2D of a ray shard failed");
            if ((rc = pm3d_los_batch(ctx, (const double *)g->d_in[r], hi - lo, (uint8_t *)g->d_out[r] + lo, PM3D_IN_DEVICE | PM3D_OUT_DEVICE | PM3D_NO_SYNC))) return group_fail_ctx(g, r, rc);
        }
    }
    if (api->GroupStart() != 0) return group_fail(g, PM3D_ERR_NCCL, "ncclGroupStart failed");
    for (int r = 0; r < n; r++) {
        cudaSetDevice(g->ctx[r]->device);
        if ((rc = rays_gather(g->ctx[r], api, (uint8_t *)g->d_out[r], n_rays, false))) { api->GroupEnd(); return group_fail_ctx(g, r, rc); }
    }
    { const ncclResult_t e = api->GroupEnd(); if (e != 0) return group_fail(g, PM3D_ERR_NCCL, std::string("ncclGroupEnd: ") + api->GetErrorString(e)); }
    pm3d_ctx *c0 = g->ctx[0];
    cudaSetDevice(c0->device);
    if (cudaMemcpyAsync(out, g->d_out[0], (size_t)n_rays, cudaMemcpyDeviceToHost, c0->stream) != cudaSuccess) return group_fail(g, PM3D_ERR_CUDA, "D2H of the gathered LOS results failed");
    return pm3d_group_sync(g);
}

}  // extern "C"
