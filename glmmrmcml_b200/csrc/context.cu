// context.cu — context, error reporting and the (dlopen'ed) NCCL communicator.
#include "common.cuh"
#include <dlfcn.h>

static thread_local char g_err[1024] = "";

int gmb_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* gmb_last_error(void) { return g_err; }
extern "C" const char* gmb_version(void) { return "glmmrmcml_b200 0.1 (sm_100a)"; }

extern "C" int gmb_ctx_create(int device, gmb_ctx** out) {
    if (!out) return gmb_set_error(GMB_EINVAL, "gmb_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return gmb_set_error(GMB_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return gmb_set_error(GMB_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    GMB_CUDA(cudaSetDevice(device));
    gmb_ctx* ctx = new gmb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    GMB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sms = prop.multiProcessorCount;
    GMB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        // keep freed blocks in the device's default memory pool instead of returning them to the driver
        cudaMemPool_t pool;
        GMB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        GMB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    {
        int lo = 0, hi = 0;
        GMB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GMB_CUDA(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
        GMB_CUDA(cudaEventCreateWithFlags(&ctx->evp, cudaEventDisableTiming));
        GMB_CUDA(cudaEventCreateWithFlags(&ctx->evn, cudaEventDisableTiming));
        GMB_CUDA(cudaEventCreateWithFlags(&ctx->evj, cudaEventDisableTiming));
        GMB_CUDA(cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, hi));
        for (int k = 0; k < 4; k++) GMB_CUDA(cudaEventCreateWithFlags(&ctx->evd[k], cudaEventDisableTiming));
        GMB_CUDA(cudaEventCreateWithFlags(&ctx->evx, cudaEventDisableTiming));
    }
    GMB_CUDA(cudaEventCreate(&ctx->ev0));
    GMB_CUDA(cudaEventCreate(&ctx->ev1));
    GMB_CUDA(cudaEventCreate(&ctx->ev2));
    GMB_CUDA(cudaEventCreate(&ctx->ev3));
    ctx->pinned_doubles = 1 << 16;
    GMB_CUDA(cudaMallocHost(&ctx->h_pinned, ctx->pinned_doubles * sizeof(double)));
    GMB_CUDA(cudaMalloc(&ctx->d_result, GMB_RESULT_DOUBLES * sizeof(double)));
    GMB_CUDA(cudaMalloc(&ctx->d_counter, 64 * sizeof(unsigned int)));
    GMB_CUDA(cudaMemset(ctx->d_counter, 0, 64 * sizeof(unsigned int)));
    int rc = gmb_ctx_scratch(ctx, 1 << 16);
    if (rc) { delete ctx; return rc; }
    *out = ctx;
    return GMB_OK;
}

cudaError_t gmb_dmalloc_raw(gmb_ctx* ctx, void** p, size_t bytes) {
    return cudaMallocAsync(p, bytes > 0 ? bytes : 1, ctx->stream);
}

void gmb_dfree(gmb_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

int gmb_ctx_scratch(gmb_ctx* ctx, size_t doubles) {
    if (doubles <= ctx->scratch_doubles) return GMB_OK;
    if (ctx->d_scratch) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, ctx->d_scratch); ctx->d_scratch = nullptr; }
    size_t want = round_up_sz(doubles, 1 << 12);
    GMB_CUDA(gmb_dmalloc(ctx, &ctx->d_scratch, want * sizeof(double)));
    ctx->scratch_doubles = want;
    return GMB_OK;
}

extern "C" int gmb_ctx_sync(gmb_ctx* ctx) {
    if (!ctx) return gmb_set_error(GMB_EINVAL, "ctx is NULL");
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMB_OK;
}

// CUDA-event timer on the context's own stream (torch.cuda.Event only sees torch's current stream)
extern "C" int gmb_ctx_timer_start(gmb_ctx* ctx) {
    if (!ctx) return gmb_set_error(GMB_EINVAL, "ctx is NULL");
    GMB_CUDA(cudaEventRecord(ctx->ev2, ctx->stream));
    return GMB_OK;
}
extern "C" int gmb_ctx_timer_stop(gmb_ctx* ctx, double* ms) {
    if (!ctx || !ms) return gmb_set_error(GMB_EINVAL, "ctx or ms is NULL");
    GMB_CUDA(cudaEventRecord(ctx->ev3, ctx->stream));
    GMB_CUDA(cudaEventSynchronize(ctx->ev3));
    float f = 0.f;
    GMB_CUDA(cudaEventElapsedTime(&f, ctx->ev2, ctx->ev3));
    *ms = f;
    return GMB_OK;
}
// Overwrites a buffer larger than the L2 cache (126 MB on B200) so that the next kernel starts cold.
extern "C" int gmb_ctx_flush_l2(gmb_ctx* ctx) {
    if (!ctx) return gmb_set_error(GMB_EINVAL, "ctx is NULL");
    const size_t bytes = (size_t)256 << 20;
    if (!ctx->d_flush) GMB_CUDA(cudaMalloc(&ctx->d_flush, bytes));
    ctx->flush_val ^= 1;
    GMB_CUDA(cudaMemsetAsync(ctx->d_flush, ctx->flush_val, bytes, ctx->stream));
    return GMB_OK;
}

extern "C" int64_t gmb_ctx_launch_count(gmb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* gmb_ctx_stream(gmb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library binds to whichever libnccl.so.2 the process already has (torch's
// bundled 2.28 when driven from Python, the system 2.27 otherwise) without a link-time dependency.
// ---------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_getuid)(NcclUniqueId*);
typedef int (*fn_initrank)(void**, int, NcclUniqueId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* h = nullptr;
    fn_getuid getuid = nullptr; fn_initrank initrank = nullptr; fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr; fn_bcast bcast = nullptr; fn_errstr errstr = nullptr;
} g_nccl;

int load_nccl() {
    if (g_nccl.h) return GMB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
    if (!g_nccl.h) return gmb_set_error(GMB_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.getuid = (fn_getuid)dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.initrank = (fn_initrank)dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.bcast = (fn_bcast)dlsym(g_nccl.h, "ncclBroadcast");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.getuid || !g_nccl.initrank || !g_nccl.allreduce || !g_nccl.bcast)
        return gmb_set_error(GMB_ENCCL, "libnccl is missing required symbols");
    return GMB_OK;
}
const int kNcclFloat64 = 8;  // ncclDouble
const int kNcclSum = 0;      // ncclSum
}  // namespace

#define GMB_NCCL(call)                                                                                    \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != 0) return gmb_set_error(GMB_ENCCL, "%s failed: %s", #call, g_nccl.errstr ? g_nccl.errstr(r__) : "?"); \
    } while (0)

extern "C" int gmb_comm_unique_id(void* id128) {
    if (!id128) return gmb_set_error(GMB_EINVAL, "id128 is NULL");
    GMB_TRY(load_nccl());
    NcclUniqueId id;
    GMB_NCCL(g_nccl.getuid(&id));
    memcpy(id128, &id, 128);
    return GMB_OK;
}

extern "C" int gmb_comm_init(gmb_ctx* ctx, const void* id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return gmb_set_error(GMB_EINVAL, "gmb_comm_init: bad arguments");
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return GMB_OK; }
    GMB_TRY(load_nccl());
    GMB_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, id128, 128);
    GMB_NCCL(g_nccl.initrank(&ctx->nccl_comm, world, id, rank));
    ctx->rank = rank; ctx->world = world;
    return GMB_OK;
}

extern "C" int gmb_comm_rank(gmb_ctx* ctx, int* rank, int* world) {
    if (!ctx) return gmb_set_error(GMB_EINVAL, "ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    return GMB_OK;
}

int gmb_comm_allreduce_dev(gmb_ctx* ctx, double* dbuf, int count) {
    if (ctx->world == 1) return GMB_OK;
    GMB_NCCL(g_nccl.allreduce(dbuf, dbuf, (size_t)count, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream));
    return GMB_OK;
}

extern "C" int gmb_comm_allreduce_host(gmb_ctx* ctx, double* buf, int count) {
    if (!ctx || !buf || count < 0) return gmb_set_error(GMB_EINVAL, "gmb_comm_allreduce_host: bad arguments");
    if (ctx->world == 1 || count == 0) return GMB_OK;
    GMB_TRY(gmb_ctx_scratch(ctx, (size_t)count));
    GMB_CUDA(cudaMemcpyAsync(ctx->d_scratch, buf, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    GMB_TRY(gmb_comm_allreduce_dev(ctx, ctx->d_scratch, count));
    GMB_CUDA(cudaMemcpyAsync(buf, ctx->d_scratch, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMB_OK;
}

extern "C" int gmb_comm_bcast_host(gmb_ctx* ctx, double* buf, int count) {
    if (!ctx || !buf || count < 0) return gmb_set_error(GMB_EINVAL, "gmb_comm_bcast_host: bad arguments");
    if (ctx->world == 1 || count == 0) return GMB_OK;
    GMB_TRY(gmb_ctx_scratch(ctx, (size_t)count));
    GMB_CUDA(cudaMemcpyAsync(ctx->d_scratch, buf, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    GMB_NCCL(g_nccl.bcast(ctx->d_scratch, ctx->d_scratch, (size_t)count, kNcclFloat64, 0, ctx->nccl_comm, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(buf, ctx->d_scratch, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMB_OK;
}

void gmb_api_release_ctx(gmb_ctx* ctx);   // api.cpp: device objects the entry points keep between calls

extern "C" void gmb_ctx_destroy(gmb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    gmb_api_release_ctx(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm && g_nccl.destroy) g_nccl.destroy(ctx->nccl_comm);
    if (ctx->d_scratch) { gmb_dfree(ctx, ctx->d_scratch); cudaStreamSynchronize(ctx->stream); }
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    if (ctx->evp) cudaEventDestroy(ctx->evp);
    if (ctx->evn) cudaEventDestroy(ctx->evn);
    if (ctx->evj) cudaEventDestroy(ctx->evj);
    for (int k = 0; k < 4; k++) if (ctx->evd[k]) cudaEventDestroy(ctx->evd[k]);
    if (ctx->evx) cudaEventDestroy(ctx->evx);
    if (ctx->stream3) { cudaStreamSynchronize(ctx->stream3); cudaStreamDestroy(ctx->stream3); }
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
