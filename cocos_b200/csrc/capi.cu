// C ABI of libcocos_b200.so (declared in include/cocos_b200.h).
// Host-side glue only: argument checks, launch-constant preparation in double
// precision, launch geometry, NCCL (loaded lazily with dlopen), error strings.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <cstdlib>
#include <tuple>

#include "kernels.h"

using namespace cb;

// ---------------------------------------------------------------- errors

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(CB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                      \
    } while (0)

#define CB_REQUIRE(cond, ...)                                \
    do {                                                     \
        if (!(cond)) return fail(CB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---------------------------------------------------------------- context

struct cb_ctx {
    int dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    ReduceWorkspace ws{};
    unsigned long long *d_inside = nullptr;
    double *h_result = nullptr;              // pinned, kMaxAcc doubles + one u64
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    void *scratch = nullptr;                 // grow-only staging buffer for the *_host entry points
    size_t scratch_bytes = 0;
    int threads = 0, blocks_per_sm = 0, ppt = 0;   // 0 = default
    int variant = 0;
    int pi_blocks_per_sm[2] = {0, 0};        // occupancy of pi_count_kernel<plain / antithetic>
    std::map<std::tuple<int, int, int, int, int>, std::pair<int, int>> occ_cache;   // (is64, ppt, multi, threads, variant) -> (bps, regs)
};

struct cb_comm {
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
    // fused peer exchange (PeerExchange in kernels.h): this rank's buffer, the peers' mappings, the launch counter
    int dev = -1;
    void *xchg = nullptr;                 // owned: kExchangeBytes on `dev`
    void *imported[kMaxPeers] = {};       // cudaIpcOpenMemHandle mappings (other processes' buffers)
    PeerExchange px = {};                 // host copy of the mapping table
    PeerExchange *d_px = nullptr;         // the same on `dev` (what the kernels read)
    bool p2p = false;
    unsigned int epoch = 0;
};

extern "C" CB_API const char *cb_last_error(void) { return g_err; }
extern "C" CB_API const char *cb_version(void) { return "cocos_b200 0.1 (sm_100a)"; }

extern "C" CB_API int cb_device_count(int *n) {
    CB_REQUIRE(n != nullptr, "n is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(CB_ERR_CUDA, "cudaGetDeviceCount failed: %s (no CPU fallback exists)", cudaGetErrorString(e));
    }
    *n = c;
    return CB_OK;
}

extern "C" CB_API int cb_device_name(int dev, char *buf, size_t buflen) {
    CB_REQUIRE(buf != nullptr && buflen > 0, "bad buffer");
    cudaDeviceProp p;
    CB_CUDA(cudaGetDeviceProperties(&p, dev));
    snprintf(buf, buflen, "%s", p.name);
    return CB_OK;
}

extern "C" CB_API int cb_device_attributes(int dev, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    cudaDeviceProp p;
    CB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return CB_OK;
}

static int ctx_allocate(cb_ctx *c, void *stream_or_null) {
    if (stream_or_null != nullptr) {
        c->stream = (cudaStream_t)stream_or_null;
    } else {
        CB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->dev));
    {   // cb_malloc/cb_free: keep freed blocks in the device's default pool instead of returning them to the driver
        cudaMemPool_t pool;
        CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->dev));
        unsigned long long keep = ~0ull;
        CB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CB_CUDA(cudaMalloc(&c->ws.partials, sizeof(double) * (size_t)kMaxGridBlocks * kMaxAcc));
    CB_CUDA(cudaMalloc(&c->ws.ticket, sizeof(unsigned int)));
    CB_CUDA(cudaMalloc(&c->ws.result, sizeof(double) * kMaxAcc));
    CB_CUDA(cudaMalloc(&c->d_inside, sizeof(unsigned long long)));
    CB_CUDA(cudaMemset(c->ws.ticket, 0, sizeof(unsigned int)));
    CB_CUDA(cudaMemset(c->ws.result, 0, sizeof(double) * kMaxAcc));
    CB_CUDA(cudaMallocHost(&c->h_result, sizeof(double) * (kMaxAcc + 1)));
    CB_CUDA(cudaEventCreate(&c->ev_start));
    CB_CUDA(cudaEventCreate(&c->ev_stop));
    return CB_OK;
}

extern "C" CB_API int cb_ctx_destroy(cb_ctx *c);

extern "C" CB_API int cb_ctx_create(int dev, void *stream_or_null, cb_ctx **out) {
    CB_REQUIRE(out != nullptr, "out is NULL");
    int n = 0;
    int rc = cb_device_count(&n);
    if (rc != CB_OK) return rc;
    CB_REQUIRE(dev >= 0 && dev < n, "device %d out of range (%d devices)", dev, n);
    CB_CUDA(cudaSetDevice(dev));
    cb_ctx *c = new cb_ctx();
    c->dev = dev;
    rc = ctx_allocate(c, stream_or_null);
    if (rc != CB_OK) {                       // release whatever was allocated; the error text is already set
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        cb_ctx_destroy(c);
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    *out = c;
    return CB_OK;
}

extern "C" CB_API int cb_ctx_destroy(cb_ctx *c) {
    if (c == nullptr) return CB_OK;
    cudaSetDevice(c->dev);
    if (c->stream != nullptr || !c->own_stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->ws.partials);
    cudaFree(c->ws.ticket);
    cudaFree(c->ws.result);
    cudaFree(c->d_inside);
    cudaFree(c->scratch);
    cudaFreeHost(c->h_result);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CB_OK;
}

#define CB_CTX(c)                                   \
    CB_REQUIRE((c) != nullptr, "context is NULL"); \
    CB_CUDA(cudaSetDevice((c)->dev))

extern "C" CB_API int cb_ctx_sync(cb_ctx *c) {
    CB_CTX(c);
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_ctx_device(cb_ctx *c, int *dev) {
    CB_REQUIRE(c != nullptr && dev != nullptr, "NULL argument");
    *dev = c->dev;
    return CB_OK;
}

extern "C" CB_API void *cb_ctx_stream(cb_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" CB_API int cb_timer_start(cb_ctx *c) {
    CB_CTX(c);
    CB_CUDA(cudaEventRecord(c->ev_start, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_timer_stop(cb_ctx *c, float *elapsed_ms) {
    CB_CTX(c);
    CB_REQUIRE(elapsed_ms != nullptr, "elapsed_ms is NULL");
    CB_CUDA(cudaEventRecord(c->ev_stop, c->stream));
    CB_CUDA(cudaEventSynchronize(c->ev_stop));
    CB_CUDA(cudaEventElapsedTime(elapsed_ms, c->ev_start, c->ev_stop));
    return CB_OK;
}

extern "C" CB_API int cb_ctx_set_launch(cb_ctx *c, int threads, int blocks_per_sm, int pairs_per_thread) {
    CB_REQUIRE(c != nullptr, "context is NULL");
    CB_REQUIRE(threads == 0 || (threads >= 32 && threads <= 256 && threads % 32 == 0), "threads must be 0 or a multiple of 32 in [32,256]");
    CB_REQUIRE(blocks_per_sm >= 0 && blocks_per_sm <= 64, "blocks_per_sm must be in [0,64]");
    CB_REQUIRE(pairs_per_thread == 0 || pairs_per_thread == 1 || pairs_per_thread == 2 || pairs_per_thread == 4,
               "pairs_per_thread must be 0, 1, 2 or 4");
    c->threads = threads;
    c->blocks_per_sm = blocks_per_sm;
    c->ppt = pairs_per_thread;
    return CB_OK;
}

extern "C" CB_API int cb_ctx_set_variant(cb_ctx *c, int variant) {
    CB_REQUIRE(c != nullptr, "context is NULL");
    CB_REQUIRE(variant >= 0 && variant <= 2, "variant must be in [0,2]");
    c->variant = variant;
    return CB_OK;
}

// Device memory for the caller comes from the device's stream-ordered pool (cudaMallocAsync): allocation and release
// are enqueued on the context's stream, cost no device synchronisation, and the pool keeps freed blocks for reuse
// (release threshold = unlimited, set when the context is created) -- a caching allocator without a framework.
extern "C" CB_API int cb_malloc(cb_ctx *c, size_t bytes, void **d_ptr) {
    CB_CTX(c);
    CB_REQUIRE(d_ptr != nullptr, "d_ptr is NULL");
    CB_CUDA(cudaMallocAsync(d_ptr, bytes ? bytes : 1, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_free(cb_ctx *c, void *d_ptr) {
    CB_CTX(c);
    if (d_ptr == nullptr) return CB_OK;
    CB_CUDA(cudaFreeAsync(d_ptr, c->stream));      // ordered after every kernel already enqueued on the stream
    return CB_OK;
}

extern "C" CB_API int cb_memcpy_h2d(cb_ctx *c, void *d_dst, const void *h_src, size_t bytes) {
    CB_CTX(c);
    CB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_memcpy_d2h(cb_ctx *c, void *h_dst, const void *d_src, size_t bytes) {
    CB_CTX(c);
    CB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_memcpy_d2d(cb_ctx *c, void *d_dst, const void *d_src, size_t bytes) {
    CB_CTX(c);
    CB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return CB_OK;
}

// copy from another context's device (NVLink peer copy); ordered on the DESTINATION context's stream after the
// source context's stream has drained
extern "C" CB_API int cb_memcpy_peer(cb_ctx *dst_ctx, void *d_dst, cb_ctx *src_ctx, const void *d_src, size_t bytes) {
    CB_REQUIRE(dst_ctx != nullptr && src_ctx != nullptr, "context is NULL");
    CB_CUDA(cudaSetDevice(src_ctx->dev));
    CB_CUDA(cudaStreamSynchronize(src_ctx->stream));
    CB_CUDA(cudaSetDevice(dst_ctx->dev));
    CB_CUDA(cudaMemcpyPeerAsync(d_dst, dst_ctx->dev, d_src, src_ctx->dev, bytes, dst_ctx->stream));
    return CB_OK;
}

extern "C" CB_API int cb_cast(cb_ctx *c, void *d_dst, int dst_kind, const void *d_src, int src_kind, uint64_t n) {
    CB_CTX(c);
    CB_REQUIRE((d_dst != nullptr && d_src != nullptr) || n == 0, "NULL device pointer");
    CB_REQUIRE(dst_kind >= 0 && dst_kind <= 4 && src_kind >= 0 && src_kind <= 4,
               "element kinds are 0 float32, 1 float64, 2 int32, 3 int64, 4 uint32");
    CB_CUDA(launch_cast(d_dst, dst_kind, d_src, src_kind, n, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_unary(cb_ctx *c, void *d_dst, const void *d_src, uint64_t n, int op, int is_f64) {
    CB_CTX(c);
    CB_REQUIRE((d_dst != nullptr && d_src != nullptr) || n == 0, "NULL device pointer");
    CB_REQUIRE(op == 0 || op == 1, "op must be 0 (exp) or 1 (log)");
    CB_CUDA(launch_unary(d_dst, d_src, n, op, is_f64, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_transpose(cb_ctx *c, void *d_dst, const void *d_src, uint64_t rows, uint64_t cols,
                                   int elem_bytes) {
    CB_CTX(c);
    CB_REQUIRE((d_dst != nullptr && d_src != nullptr) || rows * cols == 0, "NULL device pointer");
    CB_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "element size must be 4 or 8 bytes");
    CB_CUDA(launch_transpose(d_dst, d_src, rows, cols, elem_bytes, c->stream));
    return CB_OK;
}

static int ensure_scratch(cb_ctx *c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return CB_OK;
    CB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->scratch) CB_CUDA(cudaFree(c->scratch));
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    CB_CUDA(cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return CB_OK;
}

// ---------------------------------------------------------------- NCCL (lazy)

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

std::string g_nccl_path;   // cb_nccl_library(); else $COCOS_B200_NCCL_LIBRARY
std::string g_nccl_why;    // why loading failed: the dlerror() text, captured right where it happened

void load_nccl() {
    // 1. a libnccl already mapped into the process (the host framework's copy);
    // 2. the path the host named: a process that later loads a framework built against a NEWER libnccl.so.2 must
    //    not find an older system copy already resident under the same SONAME;
    // 3. the system library.
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        const char *env = getenv("COCOS_B200_NCCL_LIBRARY");
        const std::string path = !g_nccl_path.empty() ? g_nccl_path : std::string(env ? env : "");
        if (!path.empty()) g_nccl.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
    }
    if (!g_nccl.handle)
        for (const char *nm : names) {
            g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
            const char *why = dlerror();                    // reading it clears it: read once, keep the text
            g_nccl_why = why ? why : "dlopen failed";
        }
    if (!g_nccl.handle) return;
#define CB_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);         \
    if (!g_nccl.field) {                                            \
        const char *why = dlerror();                                \
        g_nccl_why = why ? why : std::string("missing symbol ") + name; \
        return;                                                     \
    }
    CB_SYM(GetUniqueId, "ncclGetUniqueId")
    CB_SYM(CommInitRank, "ncclCommInitRank")
    CB_SYM(CommInitAll, "ncclCommInitAll")
    CB_SYM(CommDestroy, "ncclCommDestroy")
    CB_SYM(AllReduce, "ncclAllReduce")
    CB_SYM(GroupStart, "ncclGroupStart")
    CB_SYM(GroupEnd, "ncclGroupEnd")
    CB_SYM(GetErrorString, "ncclGetErrorString")
#undef CB_SYM
    g_nccl.ok = true;
}

int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return fail(CB_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", g_nccl_why.c_str());
    return CB_OK;
}
}  // namespace

#define CB_NCCL(expr)                                                                                     \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
            return fail(CB_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

static_assert(sizeof(ncclUniqueId) == CB_UNIQUE_ID_BYTES, "unique id size");

extern "C" CB_API int cb_nccl_library(const char *path) {
    CB_REQUIRE(path != nullptr, "path is NULL");
    CB_REQUIRE(!g_nccl.ok, "NCCL is already loaded");
    g_nccl_path = path;
    return CB_OK;
}

extern "C" CB_API int cb_comm_unique_id(void *h_id) {
    CB_REQUIRE(h_id != nullptr, "h_id is NULL");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    CB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(h_id, &id, sizeof id);
    return CB_OK;
}

// ---- fused peer exchange: set-up ------------------------------------------------------------------------------

static void exchange_point(cb_comm *m, int peer, void *base) {
    m->px.slots[peer] = reinterpret_cast<unsigned long long *>(base);
    m->px.flags[peer] = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(base) + kExchangeSlotBytes);
}

static int exchange_allocate(cb_comm *m, int dev) {
    CB_CUDA(cudaSetDevice(dev));
    m->dev = dev;
    CB_CUDA(cudaMalloc(&m->xchg, kExchangeBytes));
    CB_CUDA(cudaMemset(m->xchg, 0, kExchangeBytes));
    CB_CUDA(cudaDeviceSynchronize());
    m->px.nranks = m->nranks;
    m->px.rank = m->rank;
    exchange_point(m, m->rank, m->xchg);
    return CB_OK;
}

// copies the finished mapping table to the device; the kernels read it from there
static int exchange_publish_table(cb_comm *m) {
    CB_CUDA(cudaSetDevice(m->dev));
    if (m->d_px == nullptr) CB_CUDA(cudaMalloc(&m->d_px, sizeof(PeerExchange)));
    CB_CUDA(cudaMemcpy(m->d_px, &m->px, sizeof(PeerExchange), cudaMemcpyHostToDevice));
    return CB_OK;
}

static void exchange_release(cb_comm *m) {
    if (m->dev >= 0) cudaSetDevice(m->dev);
    for (int p = 0; p < kMaxPeers; ++p)
        if (m->imported[p]) cudaIpcCloseMemHandle(m->imported[p]);
    if (m->xchg) cudaFree(m->xchg);
    if (m->d_px) cudaFree(m->d_px);
    m->xchg = nullptr;
    m->d_px = nullptr;
    m->p2p = false;
}

// one process, n devices: peer access both ways between every pair, then plain pointers
static bool exchange_connect_local(int n, cb_ctx *const *ctxs, cb_comm **comms) {
    if (n < 2 || n > kMaxPeers || getenv("COCOS_B200_NO_P2P") != nullptr) return false;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            if (ctxs[i]->dev == ctxs[j]->dev) return false;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctxs[i]->dev, ctxs[j]->dev) != cudaSuccess || !can) return false;
        }
    for (int i = 0; i < n; ++i) {
        if (cudaSetDevice(ctxs[i]->dev) != cudaSuccess) return false;
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->dev, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); return false; }
        }
    }
    for (int i = 0; i < n; ++i)
        if (exchange_allocate(comms[i], ctxs[i]->dev) != CB_OK) return false;
    for (int i = 0; i < n; ++i) {
        for (int p = 0; p < n; ++p) exchange_point(comms[i], p, comms[p]->xchg);
        if (exchange_publish_table(comms[i]) != CB_OK) return false;
    }
    for (int i = 0; i < n; ++i) comms[i]->p2p = true;
    return true;
}

extern "C" CB_API int cb_comm_p2p_export(cb_ctx *c, cb_comm *m, void *h_handle) {
    CB_CTX(c);
    CB_REQUIRE(m != nullptr && h_handle != nullptr, "NULL argument");
    CB_REQUIRE(m->nranks >= 2 && m->nranks <= kMaxPeers, "the fused exchange serves 2..%d ranks", kMaxPeers);
    static_assert(sizeof(cudaIpcMemHandle_t) == CB_P2P_HANDLE_BYTES, "ipc handle size");
    if (m->xchg == nullptr) {
        int rc = exchange_allocate(m, c->dev);
        if (rc) return rc;
    }
    cudaIpcMemHandle_t h;
    CB_CUDA(cudaIpcGetMemHandle(&h, m->xchg));
    memcpy(h_handle, &h, sizeof h);
    return CB_OK;
}

extern "C" CB_API int cb_comm_p2p_import(cb_ctx *c, cb_comm *m, const void *h_handles) {
    CB_CTX(c);
    CB_REQUIRE(m != nullptr && h_handles != nullptr, "NULL argument");
    CB_REQUIRE(m->xchg != nullptr, "cb_comm_p2p_export comes first");
    for (int p = 0; p < m->nranks; ++p) {
        if (p == m->rank || m->imported[p] != nullptr) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(h_handles) + (size_t)p * sizeof h, sizeof h);
        void *base = nullptr;
        CB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        m->imported[p] = base;
        exchange_point(m, p, base);
    }
    return CB_OK;
}

// All ranks must agree: enable only after EVERY rank imported successfully (the host side checks that).
extern "C" CB_API int cb_comm_p2p_enable(cb_comm *m, int enable) {
    CB_REQUIRE(m != nullptr, "comm is NULL");
    if (enable) {
        CB_REQUIRE(m->xchg != nullptr, "no exchange buffer: export/import first");
        for (int p = 0; p < m->nranks; ++p) CB_REQUIRE(m->px.slots[p] != nullptr, "rank %d is not mapped", p);
        int rc = exchange_publish_table(m);
        if (rc) return rc;
    }
    m->p2p = enable != 0;
    return CB_OK;
}

// Unmaps the other processes' buffers and turns the fused exchange off.  An exported buffer must not be freed while
// another process still has it mapped: every rank disconnects, the host synchronises the ranks, then cb_comm_destroy.
extern "C" CB_API int cb_comm_p2p_disconnect(cb_comm *m) {
    CB_REQUIRE(m != nullptr, "comm is NULL");
    m->p2p = false;
    if (m->dev >= 0) CB_CUDA(cudaSetDevice(m->dev));
    for (int p = 0; p < kMaxPeers; ++p)
        if (m->imported[p]) {
            cudaIpcCloseMemHandle(m->imported[p]);
            m->imported[p] = nullptr;
            m->px.slots[p] = nullptr;
            m->px.flags[p] = nullptr;
        }
    return CB_OK;
}

extern "C" CB_API int cb_comm_p2p_active(const cb_comm *m, int *active) {
    CB_REQUIRE(m != nullptr && active != nullptr, "NULL argument");
    *active = m->p2p ? 1 : 0;
    return CB_OK;
}

extern "C" CB_API int cb_comm_init_rank(cb_ctx *c, int nranks, int rank, const void *h_id, cb_comm **out) {
    CB_CTX(c);
    CB_REQUIRE(out != nullptr && h_id != nullptr, "NULL argument");
    CB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, h_id, sizeof id);
    cb_comm *m = new cb_comm();
    m->nranks = nranks;
    m->rank = rank;
    ncclResult_t r = g_nccl.CommInitRank(&m->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete m;
        return fail(CB_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    *out = m;
    return CB_OK;
}

extern "C" CB_API int cb_comm_init_all(int n, cb_ctx *const *ctxs, cb_comm **out) {
    CB_REQUIRE(n >= 1 && ctxs != nullptr && out != nullptr, "bad arguments");
    int rc = need_nccl();
    if (rc) return rc;
    int devs[64];
    CB_REQUIRE(n <= 64, "too many devices");
    for (int i = 0; i < n; ++i) {
        CB_REQUIRE(ctxs[i] != nullptr, "context %d is NULL", i);
        devs[i] = ctxs[i]->dev;
    }
    ncclComm_t comms[64];
    CB_NCCL(g_nccl.CommInitAll(comms, n, devs));
    for (int i = 0; i < n; ++i) {
        out[i] = new cb_comm();
        out[i]->comm = comms[i];
        out[i]->nranks = n;
        out[i]->rank = i;
    }
    if (!exchange_connect_local(n, ctxs, out))              // no peer access: the NCCL allreduce stays in the path
        for (int i = 0; i < n; ++i) exchange_release(out[i]);
    return CB_OK;
}

extern "C" CB_API int cb_comm_destroy(cb_comm *m) {
    if (m == nullptr) return CB_OK;
    exchange_release(m);
    if (g_nccl.ok && m->comm) g_nccl.CommDestroy(m->comm);
    delete m;
    return CB_OK;
}

extern "C" CB_API int cb_allreduce_sum_f64(cb_ctx *c, cb_comm *m, double *d_buf, size_t count) {
    CB_CTX(c);
    CB_REQUIRE(m != nullptr && d_buf != nullptr, "NULL argument");
    CB_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, m->comm, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_group_start(void) {
    int rc = need_nccl();
    if (rc) return rc;
    CB_NCCL(g_nccl.GroupStart());
    return CB_OK;
}

extern "C" CB_API int cb_group_end(void) {
    int rc = need_nccl();
    if (rc) return rc;
    CB_NCCL(g_nccl.GroupEnd());
    return CB_OK;
}

// ---------------------------------------------------------------- random arrays

static int check_out(cb_ctx *c, const void *d_out, uint64_t n) {
    CB_CTX(c);
    CB_REQUIRE(d_out != nullptr || n == 0, "d_out is NULL");
    return CB_OK;
}

extern "C" CB_API int cb_philox_words(cb_ctx *c, uint32_t *d_out, uint64_t n_blocks, uint64_t first_index, uint32_t block,
                               uint64_t seed, uint32_t subsequence) {
    int rc = check_out(c, d_out, n_blocks);
    if (rc) return rc;
    CB_CUDA(launch_philox_words(d_out, n_blocks, first_index, block, philox_keys_from_seed(seed), subsequence,
                                c->stream));
    return CB_OK;
}

#define CB_FILL(name, real, normal)                                                                        \
    extern "C" CB_API int name(cb_ctx *c, real *d_out, uint64_t n, uint64_t seed, uint32_t subsequence) {        \
        int rc = check_out(c, d_out, n);                                                                   \
        if (rc) return rc;                                                                                 \
        CB_CUDA((launch_random_fill<real, normal>(d_out, n, philox_keys_from_seed(seed), subsequence,      \
                                                  c->stream)));                                            \
        return CB_OK;                                                                                      \
    }
CB_FILL(cb_rand_f32, float, false)
CB_FILL(cb_rand_f64, double, false)
CB_FILL(cb_randn_f32, float, true)
CB_FILL(cb_randn_f64, double, true)

// mirrors verify_shape_and_antithetic_dimension (cocos/numerics/random.py:103-120)
static int split_shape(const uint64_t *dims, int ndim, int axis, uint64_t *outer, uint64_t *d, uint64_t *inner) {
    CB_REQUIRE(dims != nullptr, "dims is NULL");
    CB_REQUIRE(ndim <= 4, "arrays with more than 4 axes are not supported");
    CB_REQUIRE(ndim >= 1, "array must have at least one axis");
    CB_REQUIRE(axis >= 0 && axis <= ndim - 1, "antithetic dimension must be None or between 0 and %d", ndim);
    *outer = 1;
    *inner = 1;
    for (int i = 0; i < axis; ++i) *outer *= dims[i];
    for (int i = axis + 1; i < ndim; ++i) *inner *= dims[i];
    *d = dims[axis];
    return CB_OK;
}

#define CB_ANTI(name, real, normal)                                                                           \
    extern "C" CB_API int name(cb_ctx *c, real *d_out, const uint64_t *dims, int ndim, int axis, uint64_t seed,      \
                        uint32_t subsequence) {                                                               \
        uint64_t outer, d, inner;                                                                             \
        int rc = split_shape(dims, ndim, axis, &outer, &d, &inner);                                           \
        if (rc) return rc;                                                                                    \
        rc = check_out(c, d_out, outer * d * inner);                                                          \
        if (rc) return rc;                                                                                    \
        CB_CUDA((launch_random_antithetic<real, normal>(d_out, outer, d, inner, philox_keys_from_seed(seed),  \
                                                        subsequence, c->stream)));                            \
        return CB_OK;                                                                                         \
    }
CB_ANTI(cb_randn_antithetic_f32, float, true)
CB_ANTI(cb_randn_antithetic_f64, double, true)
CB_ANTI(cb_rand_antithetic_f32, float, false)
CB_ANTI(cb_rand_antithetic_f64, double, false)

// ---------------------------------------------------------------- derived distributions

extern "C" CB_API int cb_random_distribution_f32(cb_ctx *c, float *d_out, uint64_t n, int kind, double a, double b,
                                                 uint64_t seed, uint32_t subsequence) {
    int rc = check_out(c, d_out, n);
    if (rc) return rc;
    CB_REQUIRE(kind >= 0 && kind <= 7, "unknown distribution kind %d", kind);
    float fa = (float)a, fb = (float)b;
    switch (kind) {
        case 0: CB_REQUIRE(b >= a, "high must not be less than low"); fb = (float)(b - a); break;   // (low, high-low)
        case 1: CB_REQUIRE(a >= 0.0, "scale must be non-negative"); break;
        case 2: case 3: case 4: CB_REQUIRE(b >= 0.0, "scale must be non-negative"); break;
        case 5: CB_REQUIRE(a > 0.0 && b > 0.0, "mean and scale must be positive"); break;
        case 6: CB_REQUIRE(a > 0.0 && b > 0.0, "shape and scale must be positive"); break;
        case 7: CB_REQUIRE(a > 0.0 && b > 0.0, "a and b must be positive"); break;
    }
    CB_CUDA(launch_dist_fill(kind, d_out, n, fa, fb, philox_keys_from_seed(seed), subsequence, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_randint(cb_ctx *c, void *d_out, int out_is_64, uint64_t n, int64_t low, int64_t high,
                                 uint64_t seed, uint32_t subsequence) {
    int rc = check_out(c, d_out, n);
    if (rc) return rc;
    CB_REQUIRE(high > low, "high must be greater than low");
    const float divisor = (float)(1.0 / (double)(high - low));
    CB_CUDA(launch_randint(d_out, out_is_64, n, divisor, (long long)low, philox_keys_from_seed(seed), subsequence,
                           c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_gather(cb_ctx *c, void *d_out, const void *d_src, const int32_t *d_index, uint64_t n,
                                int elem_bytes) {
    int rc = check_out(c, d_out, n);
    if (rc) return rc;
    CB_REQUIRE(d_src != nullptr && d_index != nullptr, "NULL argument");
    CB_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "element size must be 4 or 8 bytes");
    CB_CUDA(launch_gather(d_out, d_src, d_index, n, elem_bytes, c->stream));
    return CB_OK;
}

template <typename real>
static int mvn(cb_ctx *c, real *d_out, uint64_t rows, int d, const real *h_chol, const real *h_mean, uint64_t seed,
               uint32_t subsequence) {
    int rc = check_out(c, d_out, rows * (uint64_t)d);
    if (rc) return rc;
    CB_REQUIRE(d >= 1 && d <= 64, "dimension must be in [1,64]");
    CB_REQUIRE(h_chol != nullptr && h_mean != nullptr, "NULL argument");
    const size_t bytes = sizeof(real) * ((size_t)d * d + d);
    rc = ensure_scratch(c, bytes);
    if (rc) return rc;
    real *d_chol = (real *)c->scratch, *d_mean = d_chol + (size_t)d * d;
    CB_CUDA(cudaMemcpyAsync(d_chol, h_chol, sizeof(real) * d * d, cudaMemcpyHostToDevice, c->stream));
    CB_CUDA(cudaMemcpyAsync(d_mean, h_mean, sizeof(real) * d, cudaMemcpyHostToDevice, c->stream));
    CB_CUDA((launch_random_fill<real, true>(d_out, rows * (uint64_t)d, philox_keys_from_seed(seed), subsequence,
                                            c->stream)));
    CB_CUDA(launch_mvn_transform<real>(d_out, rows, d, d_chol, d_mean, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));      // h_chol / h_mean are caller memory: done with them on return
    return CB_OK;
}

extern "C" CB_API int cb_multivariate_normal_f32(cb_ctx *c, float *d_out, uint64_t rows, int d, const float *h_chol,
                                                 const float *h_mean, uint64_t seed, uint32_t subsequence) {
    return mvn<float>(c, d_out, rows, d, h_chol, h_mean, seed, subsequence);
}

extern "C" CB_API int cb_multivariate_normal_f64(cb_ctx *c, double *d_out, uint64_t rows, int d, const double *h_chol,
                                                 const double *h_mean, uint64_t seed, uint32_t subsequence) {
    return mvn<double>(c, d_out, rows, d, h_chol, h_mean, seed, subsequence);
}

// ---------------------------------------------------------------- Gaussian KDE

extern "C" CB_API int cb_kde_moments_f32(cb_ctx *c, const float *d_points, const float *d_weights_or_null, uint64_t n,
                                         int d, double *h_out /* 1 + d + d*d */) {
    CB_CTX(c);
    CB_REQUIRE(d_points != nullptr && h_out != nullptr, "NULL argument");
    CB_REQUIRE(d >= 1 && d <= 8, "dimension must be in [1,8]");
    const size_t bytes = sizeof(double) * (1 + d + d * d);
    int rc = ensure_scratch(c, bytes);
    if (rc) return rc;
    CB_CUDA(launch_kde_moments(d_points, d_weights_or_null, n, d, (double *)c->scratch, c->stream));
    CB_CUDA(cudaMemcpyAsync(h_out, c->scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_kde_whiten_f32(cb_ctx *c, const float *d_in, uint64_t n, int d, const float *h_matrix,
                                        float *d_out) {
    CB_CTX(c);
    CB_REQUIRE(d_in != nullptr && d_out != nullptr && h_matrix != nullptr, "NULL argument");
    CB_REQUIRE(d >= 1 && d <= 8, "dimension must be in [1,8]");
    int rc = ensure_scratch(c, sizeof(float) * 64);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(c->scratch, h_matrix, sizeof(float) * d * d, cudaMemcpyHostToDevice, c->stream));
    CB_CUDA(launch_kde_whiten(d_in, n, d, (const float *)c->scratch, d_out, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));      // h_matrix is caller memory
    return CB_OK;
}

extern "C" CB_API int cb_kde_evaluate_f32(cb_ctx *c, const float *d_wpoints, const float *d_weights_or_null,
                                          uint64_t n, int d, const float *d_wxi, uint64_t m, double scale,
                                          float *d_out) {
    CB_CTX(c);
    CB_REQUIRE(d_wpoints != nullptr && (d_wxi != nullptr || m == 0) && (d_out != nullptr || m == 0), "NULL argument");
    CB_REQUIRE(d >= 1 && d <= 4, "the evaluation kernel supports 1 to 4 dimensions");
    CB_REQUIRE(n >= 1, "empty dataset");
    if (m == 0) return CB_OK;
    int pj = 1;
    const int chunks = kde_plan_chunks(n, m, c->sm_count, &pj);
    int rc = ensure_scratch(c, sizeof(float) * (size_t)chunks * m);
    if (rc) return rc;
    CB_CUDA(launch_kde_evaluate(d_wpoints, d_weights_or_null, n, d, d_wxi, m, scale, (float *)c->scratch, chunks, pj,
                                d_out, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_kde_integrate_box_1d_f32(cb_ctx *c, const float *d_points, const float *d_weights_or_null,
                                                  uint64_t n, double low, double high, double stdev, double *h_out) {
    CB_CTX(c);
    CB_REQUIRE(d_points != nullptr && h_out != nullptr, "NULL argument");
    CB_REQUIRE(n >= 1 && stdev > 0.0, "empty dataset or non-positive bandwidth");
    int rc = ensure_scratch(c, sizeof(double));
    if (rc) return rc;
    CB_CUDA(launch_kde_box_1d(d_points, d_weights_or_null, n, low, high, stdev, (double *)c->scratch, c->stream));
    CB_CUDA(cudaMemcpyAsync(h_out, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_kde_resample_f32(cb_ctx *c, const float *d_points, const double *d_cumweights_or_null, uint64_t n,
                                          int d, const float *h_chol, uint64_t size, uint64_t seed, uint32_t subsequence,
                                          float *d_out) {
    CB_CTX(c);
    CB_REQUIRE(d_points != nullptr && h_chol != nullptr && (d_out != nullptr || size == 0), "NULL argument");
    CB_REQUIRE(d >= 1 && d <= 4, "the resampling kernel supports 1 to 4 dimensions");
    CB_REQUIRE(n >= 1, "empty dataset");
    int rc = ensure_scratch(c, sizeof(float) * 16);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(c->scratch, h_chol, sizeof(float) * d * d, cudaMemcpyHostToDevice, c->stream));
    CB_CUDA(launch_kde_resample(d_points, d_cumweights_or_null, n, d, (const float *)c->scratch, size,
                                philox_keys_from_seed(seed), subsequence, d_out, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));      // h_chol is caller memory
    return CB_OK;
}

// ---------------------------------------------------------------- Heston: parity kernels

static int check_heston(const cb_heston_params *p, uint64_t R, uint32_t N) {
    CB_REQUIRE(p != nullptr, "params is NULL");
    CB_REQUIRE(R >= 1, "R must be >= 1");
    CB_REQUIRE(N >= 2, "N must be >= 2 (N-1 Euler steps)");
    CB_REQUIRE(p->rho >= -1.0 && p->rho <= 1.0, "rho must be in [-1,1]");
    return CB_OK;
}

extern "C" CB_API int cb_heston_terminal_injected_f32(cb_ctx *c, const float *d_Z, const cb_heston_params *p, uint64_t R,
                                               uint32_t N, float *d_x, float *d_v) {
    CB_CTX(c);
    int rc = check_heston(p, R, N);
    if (rc) return rc;
    CB_REQUIRE(d_Z && d_x && d_v, "NULL device pointer");
    const double dt = p->T / (double)(N - 1);
    InjectedConsts32 k;
    k.dt = (float)dt;
    k.root_dt = (float)std::sqrt(dt);
    k.m0 = (float)p->rho;
    k.m1 = (float)std::sqrt(1.0 - p->rho * p->rho);
    k.mu = (float)p->mu;
    k.kappa = (float)p->kappa;
    k.v_bar = (float)p->v_bar;
    k.sigma_v = (float)p->sigma_v;
    k.x0 = (float)p->x0;
    k.v0 = (float)p->v0;
    CB_CUDA(launch_heston_injected_f32(d_Z, k, R, N, d_x, d_v, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_heston_terminal_injected_f64(cb_ctx *c, const double *d_Z, const cb_heston_params *p, uint64_t R,
                                               uint32_t N, double *d_x, double *d_v) {
    CB_CTX(c);
    int rc = check_heston(p, R, N);
    if (rc) return rc;
    CB_REQUIRE(d_Z && d_x && d_v, "NULL device pointer");
    const double dt = p->T / (double)(N - 1);
    InjectedConsts64 k;
    k.dt = dt;
    k.root_dt = std::sqrt(dt);
    k.m0 = p->rho;
    k.m1 = std::sqrt(1.0 - p->rho * p->rho);
    k.mu = p->mu;
    k.kappa = p->kappa;
    k.v_bar = p->v_bar;
    k.sigma_v = p->sigma_v;
    k.x0 = p->x0;
    k.v0 = p->v0;
    CB_CUDA(launch_heston_injected_f64(d_Z, k, R, N, d_x, d_v, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_heston_terminal_injected_host_f32(cb_ctx *c, const float *h_Z, const cb_heston_params *p,
                                                    uint64_t R, uint32_t N, float *h_x, float *h_v) {
    CB_CTX(c);
    int rc = check_heston(p, R, N);
    if (rc) return rc;
    CB_REQUIRE(h_Z && h_x && h_v, "NULL host pointer");
    const uint64_t H = (R + 1) / 2;
    const size_t z_bytes = sizeof(float) * 2 * H * (size_t)(N - 1);
    const size_t out_bytes = sizeof(float) * R;
    const size_t z_pad = (z_bytes + 255) & ~(size_t)255, o_pad = (out_bytes + 255) & ~(size_t)255;
    rc = ensure_scratch(c, z_pad + 2 * o_pad);
    if (rc) return rc;
    float *d_Z = (float *)c->scratch;
    float *d_x = (float *)((char *)c->scratch + z_pad);
    float *d_v = (float *)((char *)c->scratch + z_pad + o_pad);
    CB_CUDA(cudaMemcpyAsync(d_Z, h_Z, z_bytes, cudaMemcpyHostToDevice, c->stream));
    rc = cb_heston_terminal_injected_f32(c, d_Z, p, R, N, d_x, d_v);
    if (rc) return rc;
    CB_CUDA(cudaMemcpyAsync(h_x, d_x, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaMemcpyAsync(h_v, d_v, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return CB_OK;
}

// ---------------------------------------------------------------- Heston: fused kernel

// Hands the kernel this launch's view of the communicator's peer mappings; false = no fused exchange.
// The launch number is only COMMITTED (exchange_commit) once the launch has been accepted by the runtime: a launch
// that fails leaves the communicator where it was, so the ranks' counters cannot drift apart on an error path.
// `sig` identifies the call (what is launched, on which seed/subsequence and sizes); every rank must pass the same
// one, the slots' tags carry it and a rank that folds another call's data gets the poison value instead of a sum.
static bool exchange_arm(cb_comm *comm, ReduceWorkspace *ws, unsigned int sig) {
    if (comm == nullptr || !comm->p2p) return false;
    ws->px = comm->d_px;
    ws->epoch = comm->epoch + 1;
    ws->sig = sig;
    return true;
}
static void exchange_commit(cb_comm *comm) { ++comm->epoch; }

static unsigned int call_signature(unsigned int kind, uint64_t seed, uint32_t subsequence, uint64_t a, uint64_t b) {
    uint64_t h = 0x9E3779B97F4A7C15ull * (kind + 1);
    for (uint64_t v : {seed, (uint64_t)subsequence, a, b}) {
        h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    }
    return (unsigned int)(h & 0xFFFFFFu);
}

// `result` after a fused exchange: the poison values say what went wrong
static int check_exchange_poison(unsigned long long first_bits) {
    if (first_bits == kExchangeTimeoutBits)
        return fail(CB_ERR_CUDA, "fused peer exchange timed out: a rank of the communicator never launched its shard");
    if (first_bits == kExchangeMismatchBits)
        return fail(CB_ERR_CUDA, "fused peer exchange: a rank published another launch's data (the ranks' launch "
                                 "sequences differ: every rank must issue the same calls in the same order)");
    return CB_OK;
}

template <typename real>
static FusedConsts<real> make_consts(const cb_heston_params *p, uint64_t R, uint32_t N, const double *strikes, int nK,
                                     uint64_t seed, uint32_t subsequence, uint64_t pair_first, uint64_t pair_count) {
    FusedConsts<real> k;
    const double dt = p->T / (double)(N - 1);
    k.drift_v = (real)(-0.5 * dt);
    k.decay = (real)(1.0 - p->kappa * dt);
    k.pull = (real)(p->kappa * p->v_bar * dt);
    k.w0 = (real)(p->sigma_v * p->rho);
    k.w1 = (real)(p->sigma_v * std::sqrt(1.0 - p->rho * p->rho));
    k.v0 = (real)p->v0;
    k.x0 = (real)p->x0;
    k.mu_dt = (real)(p->mu * dt);
    k.floor_v = (real)1.1920928955078125e-07f;
    k.neg2ln2_dt = (float)(-2.0 * 0.6931471805599453 * dt);
    k.steps = N - 1;
    k.subsequence = subsequence;
    k.pair_first = pair_first;
    k.pair_count = pair_count;
    k.partnered = R / 2;
    k.nK = nK;
    k.payoff_kind = 0;
    k.payoff_sign = (real)1;
    k.barrier_sign = (real)1;
    k.barrier_log = (real)0;
    k.keys = philox_keys_from_seed(seed);
    for (int i = 0; i < kMaxStrikes; ++i) k.strikes[i] = (real)(i < nK ? strikes[i] : 0.0);
    return k;
}

// Geometry: a grid of (resident blocks per SM) x (SM count) blocks, grid-stride over the shard's pairs.  The last
// round is usually partial: warps that lie wholly past the end skip it (heston_fused_kernel), so its cost is
// proportional to what is left and the largest residency is always the best choice.
// `xchg`: the launch goes through a communicator with NVLink peer mappings.  Every rank must then run the SAME
// instantiation, so the per-context tuning overrides (pairs per thread, variant) are ignored and the default
// geometry of the type -- the one the exchange kernels are built for -- is used; eligibility for the fused
// exchange is thereby a property of the communicator, not of a context.
template <typename real>
static int choose_shape(cb_ctx *c, uint64_t pair_count, bool multi, int payoff_kind, bool xchg, LaunchShape *s) {
    // defaults from the tools/tune_fused.py sweeps: float 256 threads x 1 pair, double 128 threads x 2 pairs
    const int default_ppt = sizeof(real) == 8 ? 2 : 1;
    const int threads = c->threads ? c->threads : (sizeof(real) == 8 ? 128 : 256);
    const int ppt = payoff_kind != 0 ? 1 : (xchg || !c->ppt) ? default_ppt : c->ppt;
    const int variant = (sizeof(real) == 4 && !multi && payoff_kind == 0 && !xchg) ? c->variant : 0;
    const bool with_exchange = xchg && payoff_kind == 0;      // the path-dependent kernels carry the exchange anyway
    const auto key = std::make_tuple((int)(sizeof(real) == 8), ppt, (int)multi, threads,
                                     variant + 16 * payoff_kind + 64 * (int)with_exchange);
    auto it = c->occ_cache.find(key);
    if (it == c->occ_cache.end()) {
        int bps = 0, regs = 0;
        CB_CUDA((fused_occupancy<real>(threads, ppt, multi, payoff_kind, variant, with_exchange, &bps, &regs)));
        it = c->occ_cache.emplace(key, std::make_pair(bps, regs)).first;
    }
    int bps = it->second.first;
    CB_REQUIRE(bps >= 1, "fused kernel does not fit on an SM with %d threads", threads);
    const uint64_t per_block = (uint64_t)threads * ppt;
    const uint64_t items = std::max<uint64_t>(1, (pair_count + per_block - 1) / per_block);
    const uint64_t resident = (uint64_t)c->sm_count * bps;
    uint64_t blocks;
    if (c->blocks_per_sm > 0) {
        // explicit grid: sm_count * blocks_per_sm blocks (more than fit at once => the hardware back-fills)
        blocks = std::min<uint64_t>(items, (uint64_t)c->sm_count * c->blocks_per_sm);
    } else {
        blocks = std::min(items, resident);
    }
    blocks = std::min<uint64_t>(blocks, kMaxGridBlocks);
    s->threads = threads;
    s->blocks = (int)blocks;
    s->pairs_per_thread = ppt;
    s->variant = variant;
    return CB_OK;
}

template <typename real>
static int heston_price_launch(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N, const double *h_strikes,
                               int nK, uint64_t seed, uint32_t subsequence, uint64_t pair_first, uint64_t pair_count,
                               cb_comm *comm, real *d_x, real *d_v, real *d_traj, real *d_traj_v,
                               const cb_payoff_spec *spec = nullptr) {
    CB_CTX(c);
    int rc = check_heston(p, R, N);
    if (rc) return rc;
    CB_REQUIRE(h_strikes != nullptr, "h_strikes is NULL");
    CB_REQUIRE(nK >= 1 && nK <= CB_MAX_STRIKES, "nK must be in [1,%d]", CB_MAX_STRIKES);
    const uint64_t H = (R + 1) / 2;
    CB_REQUIRE(pair_first <= H && pair_count <= H - pair_first, "pair range [%llu,+%llu) outside [0,%llu)",
               (unsigned long long)pair_first, (unsigned long long)pair_count, (unsigned long long)H);
    CB_REQUIRE((d_x == nullptr) == (d_v == nullptr), "d_x and d_v must both be given or both be NULL");
    CB_REQUIRE(d_traj_v == nullptr || d_traj != nullptr, "the variance trajectory comes with the log-price trajectory");
    const bool p2p = comm != nullptr && comm->p2p;
    if (comm != nullptr && !p2p) {
        rc = need_nccl();
        if (rc) return rc;
    }
    FusedConsts<real> k = make_consts<real>(p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count);
    if (spec != nullptr) {
        CB_REQUIRE(spec->kind >= 0 && spec->kind <= 3, "payoff kind must be 0 (european), 1 (asian), 2 (up-and-out), 3 (down-and-out)");
        CB_REQUIRE(d_traj == nullptr, "path-dependent payoffs and trajectory output are separate launches");
        k.payoff_kind = spec->kind >= 2 ? 2 : spec->kind;
        k.payoff_sign = (real)(spec->is_put ? -1.0 : 1.0);
        if (spec->kind >= 2) {
            CB_REQUIRE(spec->barrier > 0.0, "the barrier level must be positive");
            const double sgn = spec->kind == 2 ? 1.0 : -1.0;
            k.barrier_sign = (real)sgn;
            k.barrier_log = (real)(sgn * std::log(spec->barrier));
        }
    }
    // The cross-GPU sum is fused into the kernel's last block over NVLink peer mappings whenever the communicator
    // has them -- every launch type carries it (an empty shard still launches one block to publish its zeros);
    // a communicator without peer mappings gets one ncclAllReduce behind the kernel instead.
    ReduceWorkspace ws = c->ws;
    const unsigned int sig = call_signature(10u + (unsigned)k.payoff_kind + (d_traj ? 4u : 0u) + (sizeof(real) == 8 ? 8u : 0u),
                                            seed, subsequence, R, ((uint64_t)N << 8) | (uint64_t)nK);
    const bool fused_exchange = exchange_arm(comm, &ws, sig);
    if (d_traj != nullptr) {
        int threads = 0, bps = 0;
        const bool bulk = trajectory_bulk_eligible<real>(k, d_traj, d_traj_v) && getenv("COCOS_B200_NO_BULK") == nullptr;
        CB_CUDA((trajectory_occupancy<real>(d_traj_v != nullptr, nK > 1, bulk, &threads, &bps)));
        CB_REQUIRE(bps >= 1, "trajectory kernel does not fit on an SM");
        const uint64_t tiles = std::max<uint64_t>(1, (pair_count + threads - 1) / threads);
        uint64_t blocks = std::min<uint64_t>(tiles, (uint64_t)c->sm_count * bps);
        blocks = std::min<uint64_t>(blocks, kMaxGridBlocks);
        if (pair_count > 0 || fused_exchange)
            CB_CUDA((launch_heston_trajectory<real>(k, (int)blocks, bulk, ws, d_x, d_v, d_traj, d_traj_v, c->stream)));
    } else {
        LaunchShape shape;
        rc = choose_shape<real>(c, pair_count, nK > 1, k.payoff_kind, fused_exchange, &shape);
        if (rc) return rc;
        if (k.payoff_sign < (real)0) shape.variant = 0;
        if (pair_count > 0 || fused_exchange)
            CB_CUDA(launch_heston_fused<real>(k, shape, ws, d_x, d_v, c->stream));
    }
    if (fused_exchange) exchange_commit(comm);
    if (pair_count == 0 && !fused_exchange)
        CB_CUDA(cudaMemsetAsync(c->ws.result, 0, sizeof(double) * 2 * nK, c->stream));   // an empty shard adds nothing
    if (comm != nullptr && !fused_exchange)
        CB_NCCL(g_nccl.AllReduce(c->ws.result, c->ws.result, (size_t)2 * nK, ncclDouble, ncclSum, comm->comm,
                                 c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_heston_price_launch_f32(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                          const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                          uint64_t pair_first, uint64_t pair_count, cb_comm *comm, float *d_x,
                                          float *d_v, float *d_traj) {
    return heston_price_launch<float>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm, d_x,
                                      d_v, d_traj, nullptr);
}

extern "C" CB_API int cb_heston_payoff_launch_f32(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                  const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                                  uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                                  const cb_payoff_spec *spec) {
    CB_REQUIRE(spec != nullptr, "spec is NULL");
    return heston_price_launch<float>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm,
                                      nullptr, nullptr, nullptr, nullptr, spec);
}

extern "C" CB_API int cb_heston_payoff_launch_f64(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                  const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                                  uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                                  const cb_payoff_spec *spec) {
    CB_REQUIRE(spec != nullptr, "spec is NULL");
    return heston_price_launch<double>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm,
                                       nullptr, nullptr, nullptr, nullptr, spec);
}

extern "C" CB_API int cb_heston_price_launch_f64(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                          const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                          uint64_t pair_first, uint64_t pair_count, cb_comm *comm, double *d_x,
                                          double *d_v, double *d_traj) {
    return heston_price_launch<double>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm,
                                       d_x, d_v, d_traj, nullptr);
}

// Path-materialising mode with both state variables: x_t into d_traj_x and v_t into d_traj_v (NULL: log prices only),
// each [N][2*pair_count] in the [first halves | partners] layout of the shard; d_x/d_v (terminal state) are optional.
extern "C" CB_API int cb_heston_trajectory_launch_f32(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                      const double *h_strikes, int nK, uint64_t seed,
                                                      uint32_t subsequence, uint64_t pair_first, uint64_t pair_count,
                                                      cb_comm *comm, float *d_x, float *d_v, float *d_traj_x,
                                                      float *d_traj_v) {
    CB_REQUIRE(d_traj_x != nullptr, "d_traj_x is NULL");
    return heston_price_launch<float>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm, d_x,
                                      d_v, d_traj_x, d_traj_v);
}

extern "C" CB_API int cb_heston_trajectory_launch_f64(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                      const double *h_strikes, int nK, uint64_t seed,
                                                      uint32_t subsequence, uint64_t pair_first, uint64_t pair_count,
                                                      cb_comm *comm, double *d_x, double *d_v, double *d_traj_x,
                                                      double *d_traj_v) {
    CB_REQUIRE(d_traj_x != nullptr, "d_traj_x is NULL");
    return heston_price_launch<double>(c, p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count, comm, d_x,
                                       d_v, d_traj_x, d_traj_v);
}

extern "C" CB_API int cb_heston_price_conditional_launch_f32(cb_ctx *c, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                            const double *h_strikes, int nK, uint64_t seed,
                                                            uint32_t subsequence, uint64_t pair_first,
                                                            uint64_t pair_count, cb_comm *comm, float *d_x, float *d_v) {
    CB_CTX(c);
    int rc = check_heston(p, R, N);
    if (rc) return rc;
    CB_REQUIRE(h_strikes != nullptr, "h_strikes is NULL");
    CB_REQUIRE(nK >= 1 && nK <= CB_MAX_STRIKES, "nK must be in [1,%d]", CB_MAX_STRIKES);
    CB_REQUIRE(p->sigma_v > 0.0, "the conditional scheme needs sigma_v > 0 (use the Euler kernel otherwise)");
    const uint64_t H = (R + 1) / 2;
    CB_REQUIRE(pair_first <= H && pair_count <= H - pair_first, "pair range outside [0,%llu)", (unsigned long long)H);
    CB_REQUIRE((d_x == nullptr) == (d_v == nullptr), "d_x and d_v must both be given or both be NULL");
    if (comm != nullptr && !comm->p2p) {
        rc = need_nccl();
        if (rc) return rc;
    }
    const FusedConsts<float> k = make_consts<float>(p, R, N, h_strikes, nK, seed, subsequence, pair_first, pair_count);
    ReduceWorkspace ws = c->ws;
    const bool fused_exchange = exchange_arm(comm, &ws, call_signature(3u, seed, subsequence, R, ((uint64_t)N << 8) | (uint64_t)nK));
    if (pair_count > 0 || fused_exchange) {
        const int threads = c->threads ? c->threads : 256;
        int bps = 0;
        CB_CUDA(conditional_occupancy(threads, nK > 1, &bps));
        CB_REQUIRE(bps >= 1, "kernel does not fit");
        if (c->blocks_per_sm > 0) bps = std::min(bps, c->blocks_per_sm);
        const uint64_t items = std::max<uint64_t>(1, (pair_count + threads - 1) / threads);
        uint64_t blocks = std::min<uint64_t>(items, (uint64_t)c->sm_count * bps);
        blocks = std::min<uint64_t>(blocks, kMaxGridBlocks);
        const double dt = p->T / (double)(N - 1);
        CB_CUDA(launch_heston_conditional(k, (float)p->sigma_v, (float)p->rho, (float)dt, threads, (int)blocks, ws,
                                          d_x, d_v, c->stream));
        if (fused_exchange) exchange_commit(comm);
    } else {
        CB_CUDA(cudaMemsetAsync(c->ws.result, 0, sizeof(double) * 2 * nK, c->stream));
    }
    if (comm != nullptr && !fused_exchange)
        CB_NCCL(g_nccl.AllReduce(c->ws.result, c->ws.result, (size_t)2 * nK, ncclDouble, ncclSum, comm->comm,
                                 c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_heston_price_finish(cb_ctx *c, int nK, double *h_sum, double *h_sumsq) {
    CB_CTX(c);
    CB_REQUIRE(nK >= 1 && nK <= CB_MAX_STRIKES, "nK must be in [1,%d]", CB_MAX_STRIKES);
    CB_REQUIRE(h_sum != nullptr && h_sumsq != nullptr, "NULL output");
    CB_CUDA(cudaMemcpyAsync(c->h_result, c->ws.result, sizeof(double) * 2 * nK, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long first_bits;
    memcpy(&first_bits, c->h_result, sizeof first_bits);
    int rc = check_exchange_poison(first_bits);
    if (rc) return rc;
    memcpy(h_sum, c->h_result, sizeof(double) * nK);
    memcpy(h_sumsq, c->h_result + nK, sizeof(double) * nK);
    return CB_OK;
}

// One call for the whole multi-GPU job (SURVEY.md section 8(b) "cb_heston_price_multi"): contiguous pair shards,
// kernel + allreduce enqueued on every context's stream inside one NCCL group, a single host read at the end.
template <typename real>
static int heston_price_multi(int G, cb_ctx *const *ctxs, cb_comm *const *comms, const cb_heston_params *p, uint64_t R,
                              uint32_t N, const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                              const cb_payoff_spec *spec, double *h_sum, double *h_sumsq) {
    CB_REQUIRE(G >= 1 && ctxs != nullptr, "need at least one context");
    CB_REQUIRE(G == 1 || comms != nullptr, "G > 1 needs one communicator per context");
    for (int g = 0; g < G; ++g) {
        CB_REQUIRE(ctxs[g] != nullptr, "context %d is NULL", g);
        CB_REQUIRE(G == 1 || comms[g] != nullptr, "communicator %d is NULL", g);
    }
    const bool grouped = G > 1;
    const bool nccl_group = grouped && !comms[0]->p2p;       // the fused exchange needs no collective library call
    const uint64_t pairs = (R + 1) / 2, per = (pairs + (uint64_t)G - 1) / (uint64_t)G;
    int rc = CB_OK;
    if (nccl_group) {
        rc = cb_group_start();
        if (rc) return rc;
    }
    for (int g = 0; g < G && rc == CB_OK; ++g) {
        const uint64_t first = std::min<uint64_t>((uint64_t)g * per, pairs);
        const uint64_t count = std::min<uint64_t>(per, pairs - first);
        rc = heston_price_launch<real>(ctxs[g], p, R, N, h_strikes, nK, seed, subsequence, first, count,
                                       grouped ? comms[g] : nullptr, nullptr, nullptr, nullptr, nullptr, spec);
    }
    if (nccl_group) {
        const int rc_end = cb_group_end();          // always close the group, keep the first error
        if (rc == CB_OK) rc = rc_end;
    }
    if (rc) return rc;
    // every rank holds the reduced sums; each one is read back so that a rank whose exchange failed (poison value)
    // is reported, not just rank 0's.  The first context's copy is the one returned.
    double other_sum[CB_MAX_STRIKES], other_sq[CB_MAX_STRIKES];
    for (int g = G - 1; g >= 0; --g) {
        const int rc_g = cb_heston_price_finish(ctxs[g], nK, g == 0 ? h_sum : other_sum, g == 0 ? h_sumsq : other_sq);
        if (rc == CB_OK) rc = rc_g;
    }
    return rc;
}

extern "C" CB_API int cb_heston_price_multi_f32(int G, cb_ctx *const *ctxs, cb_comm *const *comms,
                                                const cb_heston_params *p, uint64_t R, uint32_t N,
                                                const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                                const cb_payoff_spec *spec, double *h_sum, double *h_sumsq) {
    return heston_price_multi<float>(G, ctxs, comms, p, R, N, h_strikes, nK, seed, subsequence, spec, h_sum, h_sumsq);
}

extern "C" CB_API int cb_heston_price_multi_f64(int G, cb_ctx *const *ctxs, cb_comm *const *comms,
                                                const cb_heston_params *p, uint64_t R, uint32_t N,
                                                const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                                const cb_payoff_spec *spec, double *h_sum, double *h_sumsq) {
    return heston_price_multi<double>(G, ctxs, comms, p, R, N, h_strikes, nK, seed, subsequence, spec, h_sum, h_sumsq);
}

template <typename real>
static int payoff_sums(cb_ctx *c, const real *d_x, uint64_t R, const double *h_strikes, int nK, double *h_sum,
                       double *h_sumsq) {
    CB_CTX(c);
    CB_REQUIRE(d_x != nullptr && h_strikes != nullptr, "NULL argument");
    CB_REQUIRE(R >= 1, "R must be >= 1");
    CB_REQUIRE(nK >= 1 && nK <= CB_MAX_STRIKES, "nK must be in [1,%d]", CB_MAX_STRIKES);
    CB_CUDA(launch_call_payoff_sums<real>(d_x, R, h_strikes, nK, c->ws, c->sm_count, c->stream));
    return cb_heston_price_finish(c, nK, h_sum, h_sumsq);
}

extern "C" CB_API int cb_call_payoff_sums_f32(cb_ctx *c, const float *d_x, uint64_t R, const double *h_strikes, int nK,
                                       double *h_sum, double *h_sumsq) {
    return payoff_sums<float>(c, d_x, R, h_strikes, nK, h_sum, h_sumsq);
}

extern "C" CB_API int cb_call_payoff_sums_f64(cb_ctx *c, const double *d_x, uint64_t R, const double *h_strikes, int nK,
                                       double *h_sum, double *h_sumsq) {
    return payoff_sums<double>(c, d_x, R, h_strikes, nK, h_sum, h_sumsq);
}

// ---------------------------------------------------------------- pi

extern "C" CB_API int cb_pi_count_launch(cb_ctx *c, uint64_t n, uint64_t seed, uint32_t subsequence, int antithetic,
                                  uint64_t draw_first, uint64_t draw_count, cb_comm *comm) {
    CB_CTX(c);
    const uint64_t ndraw = antithetic ? (n + 1) / 2 : n;
    CB_REQUIRE(draw_first <= ndraw && draw_count <= ndraw - draw_first, "draw range outside [0,%llu)",
               (unsigned long long)ndraw);
    const bool p2p = comm != nullptr && comm->p2p;
    if (comm != nullptr && !p2p) {
        int rc = need_nccl();
        if (rc) return rc;
    }
    int &bps = c->pi_blocks_per_sm[antithetic ? 1 : 0];
    if (bps == 0) {
        CB_CUDA(pi_occupancy(antithetic, &bps));
        CB_REQUIRE(bps >= 1, "pi kernel does not fit on an SM");
    }
    ReduceWorkspace ws = c->ws;
    const bool fused_exchange = exchange_arm(comm, &ws, call_signature(antithetic ? 2u : 1u, seed, subsequence, n, 0));
    CB_CUDA(launch_pi_count(n, antithetic, draw_first, draw_count, philox_keys_from_seed(seed), subsequence,
                            c->d_inside, c->sm_count * bps, ws, c->stream));
    if (fused_exchange) exchange_commit(comm);
    if (comm != nullptr && !fused_exchange)
        CB_NCCL(g_nccl.AllReduce(c->d_inside, c->d_inside, 1, ncclUint64, ncclSum, comm->comm, c->stream));
    return CB_OK;
}

extern "C" CB_API int cb_pi_count_finish(cb_ctx *c, uint64_t *h_inside) {
    CB_CTX(c);
    CB_REQUIRE(h_inside != nullptr, "h_inside is NULL");
    unsigned long long *h = reinterpret_cast<unsigned long long *>(c->h_result + kMaxAcc);
    CB_CUDA(cudaMemcpyAsync(h, c->d_inside, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    int rc = check_exchange_poison(*h);
    if (rc) return rc;
    *h_inside = *h;
    return CB_OK;
}

extern "C" CB_API int cb_pi_count_multi(int G, cb_ctx *const *ctxs, cb_comm *const *comms, uint64_t n, uint64_t seed,
                                        uint32_t subsequence, int antithetic, uint64_t *h_inside) {
    CB_REQUIRE(G >= 1 && ctxs != nullptr, "need at least one context");
    CB_REQUIRE(G == 1 || comms != nullptr, "G > 1 needs one communicator per context");
    for (int g = 0; g < G; ++g) {
        CB_REQUIRE(ctxs[g] != nullptr, "context %d is NULL", g);
        CB_REQUIRE(G == 1 || comms[g] != nullptr, "communicator %d is NULL", g);
    }
    const bool grouped = G > 1;
    const bool nccl_group = grouped && !comms[0]->p2p;
    const uint64_t ndraw = antithetic ? (n + 1) / 2 : n, per = (ndraw + (uint64_t)G - 1) / (uint64_t)G;
    int rc = CB_OK;
    if (nccl_group) {
        rc = cb_group_start();
        if (rc) return rc;
    }
    for (int g = 0; g < G && rc == CB_OK; ++g) {
        const uint64_t first = std::min<uint64_t>((uint64_t)g * per, ndraw);
        rc = cb_pi_count_launch(ctxs[g], n, seed, subsequence, antithetic, first, std::min<uint64_t>(per, ndraw - first),
                                grouped ? comms[g] : nullptr);
    }
    if (nccl_group) {
        const int rc_end = cb_group_end();
        if (rc == CB_OK) rc = rc_end;
    }
    if (rc) return rc;
    for (int g = G - 1; g >= 0; --g) {                       // every rank is read: a failed exchange anywhere is reported
        const int rc_g = cb_pi_count_finish(ctxs[g], h_inside);
        if (rc == CB_OK) rc = rc_g;
    }
    return rc;
}

// ---------------------------------------------------------------- calibration

extern "C" CB_API int cb_peak_measure(cb_ctx *c, int kind, int repeats, double *ops_per_s, float *ms_best) {
    CB_CTX(c);
    CB_REQUIRE(ops_per_s != nullptr, "ops_per_s is NULL");
    CB_REQUIRE(kind >= 0 && kind <= 38, "kind must be in [0,38]");
    if (repeats < 1) repeats = 1;
    size_t bytes = 4096;
    if (kind == 8 || kind == 10) bytes = (size_t)4 << 30;   // far larger than the 126 MB L2
    void *scratch = nullptr;
    CB_CUDA(cudaMalloc(&scratch, bytes));
    float best = 1e30f;
    double work = 0.0;
    int rc = CB_OK;
    for (int r = 0; r < repeats + 1; ++r) {   // first pass is a warm-up
        cudaEventRecord(c->ev_start, c->stream);
        cudaError_t e = launch_peak_kernel(kind, c->sm_count, scratch, bytes, &work, c->stream);
        cudaEventRecord(c->ev_stop, c->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(c->ev_stop);
        if (e != cudaSuccess) {
            rc = fail(CB_ERR_CUDA, "peak kernel %d failed: %s", kind, cudaGetErrorString(e));
            break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop);
        if (r > 0 && ms < best) best = ms;
    }
    cudaFree(scratch);
    if (rc) return rc;
    *ops_per_s = work / ((double)best * 1e-3);
    if (ms_best) *ms_best = best;
    return CB_OK;
}
