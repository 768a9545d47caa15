// qoa_api.cu — host side of libqoa_b200.so: the C ABI of include/qoa_b200.h on top of the sm_100a kernels.
//
// Mirrors the reference crate's host logic for the hot path (paths into rafaelcaricio/qoaudio):
//   QoaEncoder::new validation order      src/lib.rs:335-345
//   QoaEncoder::encode length check       src/lib.rs:368, size formula :374-378
//   QoaEncoder::encode_frame validation   src/lib.rs:424-430
//   write_header                          src/lib.rs:404-408
// There is deliberately no CPU implementation of encode/decode in this file: without a CUDA device every compute
// entry point fails with QOA_B200_E_CUDA.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "qoa_kernels.h"

using namespace qoa;

namespace {

thread_local std::string g_create_error; // qoa_b200_last_error(NULL): why the last ctx_create of THIS thread failed

struct PoolBlock {
    void *ptr;
    size_t size;
    bool free;
    int slot = 0; // while a pipelined host call runs: the pipeline stream (1..3) whose earlier stage last used the block
};

} // namespace

struct qoa_b200_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    std::vector<PoolBlock> pool;
    cudaStream_t last_stream = nullptr;
    bool used = false;
    // Pipelined host calls keep three internal streams in flight.  A block released by a stage that was enqueued on
    // pipeline stream k is parked in deferred[k] and handed out again only to LATER stages of the same stream (stream
    // order makes that safe without waiting), so a call needs the arenas of about three stages, not of the whole batch.
    bool pipelining = false;
    int cur_slot = 0;
    std::vector<void *> deferred[4];
    void begin_stage(int slot) // slot = 1..3: the stage about to be enqueued runs on pipe_streams[slot - 1]
    {
        cur_slot = slot;
        for (void *p : deferred[slot])
            for (auto &b : pool)
                if (b.ptr == p) {
                    b.free = true;
                    b.slot = slot;
                    break;
                }
        deferred[slot].clear();
    }
    void end_pipeline() // every pipeline stream has been synchronised
    {
        pipelining = false;
        cur_slot = 0;
        for (auto &d : deferred) {
            for (void *p : d) release(p);
            d.clear();
        }
        for (auto &b : pool) b.slot = 0;
    }
    cudaStream_t pipe_streams[3] = {nullptr, nullptr, nullptr};
    // a mixed batch launches one decode kernel per channel-count class: classes after the first run on these side
    // streams (forked from / joined to the caller's stream by events) so that small classes overlap instead of each
    // draining the machine on its own
    cudaStream_t cls_streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t cls_fork = nullptr, cls_join[3] = {nullptr, nullptr, nullptr};
    bool cls_ready()
    {
        if (cls_fork) return true;
        for (auto &cs : cls_streams)
            if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) return cudaGetLastError(), false;
        for (auto &e : cls_join)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return cudaGetLastError(), false;
        if (cudaEventCreateWithFlags(&cls_fork, cudaEventDisableTiming) != cudaSuccess) return cudaGetLastError(), false;
        return true;
    }
    // host-mapped (pinned) result slots the index kernels write directly: one region per pipeline stream
    struct Mapped {
        void *ptr = nullptr;
        size_t size = 0;
    } mapped[4];
    void *mapped_get(int slot, size_t bytes)
    {
        Mapped &m = mapped[slot];
        if (m.size < bytes) {
            if (m.ptr) cudaFreeHost(m.ptr);
            m.ptr = nullptr;
            m.size = 0;
            const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
            if (cudaHostAlloc(&m.ptr, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            m.size = want;
        }
        return m.ptr;
    }
    int slot_of(cudaStream_t s) const
    {
        for (int i = 0; i < 3; ++i)
            if (pipelining && s == pipe_streams[i]) return i + 1;
        return 0;
    }

    void *alloc(size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        PoolBlock *best = nullptr;
        for (auto &b : pool)
            if (b.free && b.size >= bytes && b.size <= bytes * 2 + 4096 && (!best || b.size < best->size) &&
                (!pipelining || b.slot == 0 || b.slot == cur_slot))
                best = &b;
        if (best) {
            best->free = false;
            return best->ptr;
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            // drop cached free blocks and retry once
            cudaGetLastError();
            for (auto &b : pool)
                if (b.free && b.ptr && (!pipelining || b.slot == 0 || b.slot == cur_slot)) { // not in use by another stream
                    cudaFree(b.ptr);
                    b.ptr = nullptr;
                }
            pool.erase(std::remove_if(pool.begin(), pool.end(), [](const PoolBlock &b) { return b.ptr == nullptr; }),
                       pool.end());
            if (cudaMalloc(&p, bytes) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        pool.push_back({p, bytes, false, 0});
        return p;
    }
    void release(void *p)
    {
        if (!p) return;
        if (pipelining) {
            deferred[cur_slot].push_back(p);
            return;
        }
        for (auto &b : pool)
            if (b.ptr == p) {
                b.free = true;
                return;
            }
    }
    // Stream of the synchronous host-buffer entry points: an internal non-blocking stream rather than the legacy NULL
    // stream (cudaMemcpyBatchAsync refuses the latter); nullptr if it cannot be created.
    cudaStream_t host_stream()
    {
        if (!pipe_streams[0] && cudaStreamCreateWithFlags(&pipe_streams[0], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            pipe_streams[0] = nullptr;
        }
        return pipe_streams[0];
    }
    // Scratch blocks are recycled without waiting: safe as long as work is enqueued on one stream at a time.
    void enter(cudaStream_t s)
    {
        if (pipelining) return;
        if (used && s != last_stream) cudaStreamSynchronize(last_stream);
        last_stream = s;
        used = true;
    }
};

#define CHECK_CUDA(ctx, expr)                                                                        \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
            cudaGetLastError();                                                                      \
            return QOA_B200_E_CUDA;                                                                  \
        }                                                                                            \
    } while (0)

// RAII for pooled scratch: blocks go back to the pool when the call returns (the stream still owns their contents
// until it drains; see qoa_b200_ctx::enter).
struct Scratch {
    qoa_b200_ctx *ctx;
    std::vector<void *> ptrs;
    explicit Scratch(qoa_b200_ctx *c) : ctx(c) {}
    ~Scratch()
    {
        for (void *p : ptrs) ctx->release(p);
    }
    template <typename T>
    T *get(size_t count)
    {
        void *p = ctx->alloc(count * sizeof(T));
        if (p) ptrs.push_back(p);
        return static_cast<T *>(p);
    }
};

// ---- batched host <-> device copies ---------------------------------------------------------------------------------
// Items whose host side is pinned (device-accessible through UVA) and 16-byte aligned are moved by ONE kernel
// (launch_copy_items); the rest fall back to one cudaMemcpyAsync each.  `host_is_src` tells which side is the host's.
namespace {

bool device_accessible(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) &&
           a.devicePointer != nullptr;
}

// `known` (optional, one entry per item): 1 = the host side is known to be device-accessible, 0 = known not to be;
// without it every item's host pointer is queried (cudaPointerGetAttributes, ~1 us each).
int copy_batch(qoa_b200_ctx *ctx, Scratch &sc, const std::vector<CopyItem> &items, bool host_is_src, cudaStream_t s,
               const std::vector<char> *known = nullptr)
{
    // Three mechanisms, chosen per buffer (crossovers measured with tools/e2e_small_streams.py and
    // tools/e2e_encode_probe.py; the environment variables are development aids):
    //  (1) runs of >= 8 equal-sized buffers at a constant pitch on both sides (rows of a batch tensor): ONE strided copy
    //      on a copy engine -- full PCIe rate, no per-buffer cost, overlaps the neighbouring stages' kernels;
    //  (2) small pinned buffers: one copy kernel through mapped pinned memory (SM loads read host memory at ~22 GB/s;
    //      device -> host goes through the bulk-copy engine at 35-42 GB/s), no per-buffer cost at all;
    //  (3) the rest in ONE cudaMemcpyBatchAsync (CUDA 12.8+; ~0.9 us per buffer inside the driver, 45-49 GB/s for
    //      128-512 KB buffers, on the copy engines) -- or, on the legacy NULL stream / an older driver, one
    //      cudaMemcpyAsync per buffer (4-8 us of host time each).
    static const bool kernel_on = !(getenv("QOA_B200_COPY_KERNEL") && atoi(getenv("QOA_B200_COPY_KERNEL")) == 0);
    static const bool use_2d = !(getenv("QOA_B200_COPY_2D") && atoi(getenv("QOA_B200_COPY_2D")) == 0);
    static std::atomic<bool> use_batch{!(getenv("QOA_B200_COPY_BATCH") && atoi(getenv("QOA_B200_COPY_BATCH")) == 0)}; // written by any host thread
    static const long env_h2d = getenv("QOA_B200_COPY_KERNEL_MAX_KB_H2D") ? atol(getenv("QOA_B200_COPY_KERNEL_MAX_KB_H2D")) : -1;
    static const long env_d2h = getenv("QOA_B200_COPY_KERNEL_MAX_KB_D2H") ? atol(getenv("QOA_B200_COPY_KERNEL_MAX_KB_D2H")) : -1;
#if CUDART_VERSION >= 12080
    const bool batch_ok = use_batch && s != nullptr && s != cudaStreamLegacy;
#else
    const bool batch_ok = false;
#endif
    // kernel-vs-copy-engine crossover: against the batched call 32 KB (host -> device) / 96 KB (device -> host),
    // against one call per buffer 256 KB / 1 MB
    const long env_max = host_is_src ? env_h2d : env_d2h;
    const uint64_t kernel_max =
        (env_max >= 0 ? (uint64_t)env_max : batch_ok ? (host_is_src ? 32ull : 96ull) : (host_is_src ? 256ull : 1024ull)) << 10;
    const bool many = kernel_on && items.size() >= 32; // a handful of copies is cheaper through the copy engines
    const cudaMemcpyKind kind = host_is_src ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    std::vector<CopyItem> fast, rest;
    uint64_t max_bytes = 0;
    for (size_t k = 0; k < items.size(); ++k) {
        const CopyItem &it = items[k];
        if (it.bytes == 0) continue;
        if (use_2d && it.bytes >= (16u << 10) && k + 1 < items.size() && items[k + 1].bytes == it.bytes) {
            const ptrdiff_t ps = static_cast<const char *>(items[k + 1].src) - static_cast<const char *>(it.src);
            const ptrdiff_t pd = static_cast<char *>(items[k + 1].dst) - static_cast<char *>(it.dst);
            const ptrdiff_t lim = 0x7fffffff; // cudaDeviceProp::memPitch
            if (ps >= (ptrdiff_t)it.bytes && pd >= (ptrdiff_t)it.bytes && ps <= lim && pd <= lim) {
                size_t r = k + 1;
                while (r + 1 < items.size() && items[r + 1].bytes == it.bytes &&
                       static_cast<const char *>(items[r + 1].src) - static_cast<const char *>(items[r].src) == ps &&
                       static_cast<char *>(items[r + 1].dst) - static_cast<char *>(items[r].dst) == pd)
                    ++r;
                if (r - k + 1 >= 8) {
                    CHECK_CUDA(ctx, cudaMemcpy2DAsync(it.dst, (size_t)pd, it.src, (size_t)ps, it.bytes, r - k + 1, kind, s));
                    k = r;
                    continue;
                }
            }
        }
        const void *host = host_is_src ? it.src : it.dst;
        const bool aligned = ((reinterpret_cast<uintptr_t>(it.src) | reinterpret_cast<uintptr_t>(it.dst)) & 15) == 0;
        if (many && aligned && it.bytes < kernel_max && (known ? (*known)[k] != 0 : device_accessible(host))) {
            fast.push_back(it);
            max_bytes = std::max(max_bytes, it.bytes);
        } else {
            rest.push_back(it);
        }
    }
    if (!fast.empty()) {
        CopyItem *d_items = sc.get<CopyItem>(fast.size());
        if (!d_items) {
            ctx->err = "out of device memory (copy table)";
            return QOA_B200_E_NOMEM;
        }
        CHECK_CUDA(ctx, cudaMemcpyAsync(d_items, fast.data(), fast.size() * sizeof(CopyItem), cudaMemcpyHostToDevice, s));
        CHECK_CUDA(ctx, launch_copy_items(d_items, (uint32_t)fast.size(), max_bytes, !host_is_src, s));
        ctx->launches += 1;
    }
#if CUDART_VERSION >= 12080
    if (batch_ok && rest.size() >= 2) {
        std::vector<void *> dsts(rest.size()), srcs(rest.size());
        std::vector<size_t> sizes(rest.size());
        for (size_t k = 0; k < rest.size(); ++k) {
            dsts[k] = rest[k].dst;
            srcs[k] = const_cast<void *>(rest[k].src);
            sizes[k] = rest[k].bytes;
        }
        cudaMemcpyAttributes attr;
        memset(&attr, 0, sizeof attr);
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t first = 0, fail = 0;
        if (cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), rest.size(), &attr, &first, 1, &fail, s) == cudaSuccess)
            return 0;
        cudaGetLastError();
        use_batch = false; // e.g. an older driver: one call per buffer, now and on every later call
    }
#endif
    for (const CopyItem &it : rest) CHECK_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.bytes, kind, s));
    return 0;
}

} // namespace

extern "C" {

const char *qoa_b200_version(void) { return "qoa_b200 0.1 (sm_100a)"; }

int qoa_b200_ctx_create(int device, qoa_b200_ctx **out)
{
    if (!out) return QOA_B200_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
        cudaGetLastError();
        return QOA_B200_E_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        return QOA_B200_E_ARG;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess || (e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return QOA_B200_E_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (compute capability " + std::to_string(prop.major) + "." +
                         std::to_string(prop.minor) + "); this library carries sm_100a code only";
        return QOA_B200_E_CUDA;
    }
    auto *ctx = new qoa_b200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return 0;
}

void qoa_b200_ctx_destroy(qoa_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &b : ctx->pool)
        if (b.ptr) cudaFree(b.ptr);
    for (auto &ps : ctx->pipe_streams)
        if (ps) cudaStreamDestroy(ps);
    for (auto &cs : ctx->cls_streams)
        if (cs) cudaStreamDestroy(cs);
    for (auto &e : ctx->cls_join)
        if (e) cudaEventDestroy(e);
    if (ctx->cls_fork) cudaEventDestroy(ctx->cls_fork);
    for (auto &m : ctx->mapped)
        if (m.ptr) cudaFreeHost(m.ptr);
    delete ctx;
}

const char *qoa_b200_last_error(const qoa_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t qoa_b200_kernel_launches(const qoa_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

size_t qoa_b200_encoded_size(uint32_t channels, uint32_t samples) // src/lib.rs:374-378
{
    const size_t total = samples, ch = channels;
    const size_t num_slices = (total + QOA_B200_SLICE_LEN - 1) / QOA_B200_SLICE_LEN;
    const size_t num_frames = (total + QOA_B200_FRAME_LEN - 1) / QOA_B200_FRAME_LEN;
    return QOA_B200_HEADER_SIZE + num_frames * (8 + 16 * ch) + num_slices * 8 * ch;
}

size_t qoa_b200_frame_size(uint32_t channels, uint32_t frame_samples) // src/lib.rs:602-604
{
    const size_t slices = ((size_t)frame_samples + QOA_B200_SLICE_LEN - 1) / QOA_B200_SLICE_LEN;
    return 8 + 16 * (size_t)channels + 8 * slices * channels;
}

} // extern "C"

// ================================================================================================================
// encode
// ================================================================================================================

namespace {

int validate_desc(const qoa_b200_desc &d) // QoaEncoder::new :336-345, in this order
{
    if (d.channels == 0 || d.channels > QOA_B200_MAX_CHANNELS) return QOA_B200_ENC_INVALID_CHANNELS;
    if (d.sample_rate == 0) return QOA_B200_ENC_INVALID_SAMPLE_RATE;
    if (d.samples == 0) return QOA_B200_ENC_INVALID_SAMPLES;
    return QOA_B200_OK;
}

const int32_t kInitLms[8] = {0, 0, 0, 0, 0, 0, -(1 << 13), 1 << 14}; // :348-351

// Uploads jobs + the chain table and launches the encode kernel.  d_lms holds 8 x i32 per chain at job.lms_base + c.
int run_encode_jobs(qoa_b200_ctx *ctx, Scratch &sc, const std::vector<EncJob> &jobs, const int16_t *d_pcm, uint8_t *d_out,
                    int32_t *d_lms, cudaStream_t s)
{
    std::vector<EncChain> chains;
    for (uint32_t j = 0; j < jobs.size(); ++j)
        if (jobs[j].nsamples)
            for (uint32_t c = 0; c < jobs[j].channels; ++c) chains.push_back({j, c});
    if (chains.empty()) return 0;
    EncJob *d_jobs = sc.get<EncJob>(jobs.size());
    EncChain *d_chains = sc.get<EncChain>(chains.size());
    if (!d_jobs || !d_chains) {
        ctx->err = "out of device memory (encode tables)";
        return QOA_B200_E_NOMEM;
    }
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(EncJob), cudaMemcpyHostToDevice, s));
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_chains, chains.data(), chains.size() * sizeof(EncChain), cudaMemcpyHostToDevice, s));
    int nl = 0;
    CHECK_CUDA(ctx, launch_encode(d_pcm, d_out, d_jobs, d_chains, (uint32_t)chains.size(), d_lms, s, &nl));
    ctx->launches += nl;
    return 0;
}

} // namespace

extern "C" int qoa_b200_encode_all_batch_dev(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *d_pcm,
                                             const uint64_t *pcm_off, const size_t *pcm_len, uint8_t *d_out,
                                             const uint64_t *out_off, const size_t *out_cap, size_t *out_len,
                                             int32_t *status, size_t n, void *stream)
{
    if (!ctx) return QOA_B200_E_ARG;
    if (n == 0) return 0;
    if (!descs || !d_pcm || !pcm_off || !pcm_len || !d_out || !out_off || !out_cap || !status) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_out) & 7) || (reinterpret_cast<uintptr_t>(d_pcm) & 1)) { // 8-byte slice stores
        ctx->err = "d_out must be 8-byte aligned and d_pcm 2-byte aligned";
        return QOA_B200_E_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->enter(s);
    Scratch sc(ctx);
    std::vector<EncJob> jobs(n);
    std::vector<int32_t> lms_init;
    uint32_t chains = 0;
    for (size_t i = 0; i < n; ++i) {
        EncJob &j = jobs[i];
        memset(&j, 0, sizeof j);
        if (out_len) out_len[i] = 0;
        int st = validate_desc(descs[i]);
        if (st == QOA_B200_OK && pcm_len[i] != (size_t)descs[i].samples * descs[i].channels) st = QOA_B200_ENC_INVALID_SAMPLES; // :368
        const size_t need = st == QOA_B200_OK ? qoa_b200_encoded_size(descs[i].channels, descs[i].samples) : 0;
        if (st == QOA_B200_OK && out_cap[i] < need) st = QOA_B200_ENC_IO_ERROR;
        if (st == QOA_B200_OK && (out_off[i] & 7)) {
            ctx->err = "out_off must be a multiple of 8";
            return QOA_B200_E_ARG;
        }
        status[i] = st;
        if (st != QOA_B200_OK) continue;
        j.pcm_off = pcm_off[i];
        j.out_off = out_off[i];
        j.channels = descs[i].channels;
        j.rate = descs[i].sample_rate;
        j.nsamples = descs[i].samples;
        j.lms_base = chains;
        j.file_samples = descs[i].samples;
        j.flags = 1; // write the 8-byte file header first
        chains += j.channels;
        if (out_len) out_len[i] = need;
    }
    if (chains == 0) return 0;
    lms_init.resize((size_t)chains * 8);
    for (uint32_t c = 0; c < chains; ++c) memcpy(&lms_init[(size_t)c * 8], kInitLms, sizeof kInitLms);
    int32_t *d_lms = sc.get<int32_t>(lms_init.size());
    if (!d_lms) {
        ctx->err = "out of device memory (LMS state)";
        return QOA_B200_E_NOMEM;
    }
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_lms, lms_init.data(), lms_init.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    return run_encode_jobs(ctx, sc, jobs, d_pcm, d_out, d_lms, s);
}

static int encode_all_batch_pipelined(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *const *pcm,
                                      const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                      size_t *out_len, int32_t *status, size_t n, uint64_t total_pcm_values);

// PCM values (i16) per time chunk of the pipelined host encode path; QOA_B200_ENC_CHUNK_MB (MB of PCM) is a development aid.
static uint64_t enc_chunk_values()
{
    static const uint64_t v = (getenv("QOA_B200_ENC_CHUNK_MB") ? (uint64_t)atoi(getenv("QOA_B200_ENC_CHUNK_MB")) : 176ull) << 19;
    return std::max<uint64_t>(v, 1);
}

extern "C" int qoa_b200_encode_all_batch(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *const *pcm,
                                         const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                         size_t *out_len, int32_t *status, size_t n)
{
    if (!ctx) return QOA_B200_E_ARG;
    if (n == 0) return 0;
    if (!descs || !pcm || !pcm_len || !out || !out_cap || !out_len || !status) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    {
        // >= 3 chunks of PCM: overlap the copies with the kernel in time chunks of ~176 MB (measured on the C4 shape:
        // 512 MB chunks 47.6 ms, 256 MB 42.7 ms, 84-170 MB 41.4 ms for 2 GB of PCM; a chunk holds whole frames of every
        // stream, at least one).
        uint64_t total = 0, per_frame_step = 0;
        for (size_t i = 0; i < n; ++i) {
            total += pcm_len[i];
            if (validate_desc(descs[i]) == QOA_B200_OK) per_frame_step += (uint64_t)QOA_B200_FRAME_LEN * descs[i].channels;
        }
        if (total >= 3 * enc_chunk_values() && per_frame_step && per_frame_step <= enc_chunk_values())
            return encode_all_batch_pipelined(ctx, descs, pcm, pcm_len, out, out_cap, out_len, status, n, total);
    }
    cudaStream_t s = ctx->host_stream();
    // lay the valid streams out in two device arenas
    std::vector<uint64_t> pcm_off(n, 0), out_off(n, 0);
    std::vector<size_t> caps(n, 0);
    uint64_t pcm_total = 0, out_total = 0;
    for (size_t i = 0; i < n; ++i) {
        int st = validate_desc(descs[i]);
        if (st == QOA_B200_OK && pcm_len[i] != (size_t)descs[i].samples * descs[i].channels) st = QOA_B200_ENC_INVALID_SAMPLES;
        if (st != QOA_B200_OK || !pcm[i] || !out[i]) continue;
        pcm_off[i] = pcm_total;
        pcm_total += (pcm_len[i] + 7) & ~(uint64_t)7;
        out_off[i] = out_total;
        const size_t need = qoa_b200_encoded_size(descs[i].channels, descs[i].samples);
        caps[i] = std::min(out_cap[i], need);
        out_total += (need + 15) & ~(uint64_t)15;
    }
    ctx->enter(s);
    int rc;
    {
        Scratch sc(ctx);
        int16_t *d_pcm = sc.get<int16_t>(pcm_total + 8);
        uint8_t *d_out = sc.get<uint8_t>(out_total + 16);
        if (!d_pcm || !d_out) {
            ctx->err = "out of device memory (encode arenas)";
            return QOA_B200_E_NOMEM;
        }
        std::vector<CopyItem> items;
        for (size_t i = 0; i < n; ++i)
            if (caps[i]) items.push_back({pcm[i], d_pcm + pcm_off[i], pcm_len[i] * sizeof(int16_t)});
        rc = copy_batch(ctx, sc, items, true, s);
        if (rc != 0) return rc;
        rc = qoa_b200_encode_all_batch_dev(ctx, descs, d_pcm, pcm_off.data(), pcm_len, d_out, out_off.data(), caps.data(),
                                           out_len, status, n, s);
        if (rc != 0) return rc;
        items.clear();
        for (size_t i = 0; i < n; ++i)
            if (status[i] == QOA_B200_OK && out_len[i]) items.push_back({d_out + out_off[i], out[i], out_len[i]});
        rc = copy_batch(ctx, sc, items, false, s);
        if (rc != 0) return rc;
        CHECK_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return 0;
}

// ---- streaming encoder ------------------------------------------------------------------------------------------

struct qoa_b200_encoder_batch {
    qoa_b200_ctx *ctx = nullptr;
    std::vector<qoa_b200_desc> descs;
    std::vector<int32_t> new_status;
    std::vector<uint32_t> lms_base;
    int32_t *d_lms = nullptr;
    uint32_t chains = 0;
};

extern "C" int qoa_b200_encoder_batch_new(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, size_t n, int32_t *status,
                                          qoa_b200_encoder_batch **out)
{
    if (!ctx || !out || (n && !descs)) return QOA_B200_E_ARG;
    *out = nullptr;
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    auto *enc = new qoa_b200_encoder_batch();
    enc->ctx = ctx;
    enc->descs.assign(descs, descs + n);
    enc->new_status.resize(n);
    enc->lms_base.resize(n);
    for (size_t i = 0; i < n; ++i) {
        enc->new_status[i] = validate_desc(descs[i]);
        if (status) status[i] = enc->new_status[i];
        enc->lms_base[i] = enc->chains;
        if (enc->new_status[i] == QOA_B200_OK) enc->chains += descs[i].channels;
    }
    std::vector<int32_t> init((size_t)std::max(enc->chains, 1u) * 8);
    for (uint32_t c = 0; c < enc->chains; ++c) memcpy(&init[(size_t)c * 8], kInitLms, sizeof kInitLms);
    cudaError_t e = cudaMalloc(&enc->d_lms, init.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(enc->d_lms, init.data(), init.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ctx->err = std::string("encoder_batch_new: ") + cudaGetErrorString(e);
        cudaGetLastError();
        if (enc->d_lms) cudaFree(enc->d_lms);
        delete enc;
        return QOA_B200_E_CUDA;
    }
    *out = enc;
    return 0;
}

extern "C" void qoa_b200_encoder_batch_free(qoa_b200_encoder_batch *enc)
{
    if (!enc) return;
    cudaSetDevice(enc->ctx->device);
    cudaDeviceSynchronize();
    if (enc->d_lms) cudaFree(enc->d_lms);
    delete enc;
}

extern "C" int qoa_b200_encoder_batch_write_header(const qoa_b200_encoder_batch *enc, uint8_t *const *out) // :404-408
{
    if (!enc || !out) return QOA_B200_E_ARG;
    for (size_t i = 0; i < enc->descs.size(); ++i) {
        if (enc->new_status[i] != QOA_B200_OK || !out[i]) continue;
        const uint32_t s = enc->descs[i].samples;
        const uint8_t h[8] = {'q', 'o', 'a', 'f', (uint8_t)(s >> 24), (uint8_t)(s >> 16), (uint8_t)(s >> 8), (uint8_t)s};
        memcpy(out[i], h, 8);
    }
    return 0;
}

extern "C" int qoa_b200_encoder_batch_encode_chunk_dev(qoa_b200_encoder_batch *enc, const int16_t *d_pcm,
                                                       const uint64_t *pcm_off, const uint32_t *chunk_samples,
                                                       uint8_t *d_out, const uint64_t *out_off, void *stream)
{
    if (!enc) return QOA_B200_E_ARG;
    qoa_b200_ctx *ctx = enc->ctx;
    if (!d_pcm || !pcm_off || !chunk_samples || !d_out || !out_off) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_out) & 7) || (reinterpret_cast<uintptr_t>(d_pcm) & 1)) { // 8-byte slice stores
        ctx->err = "d_out must be 8-byte aligned and d_pcm 2-byte aligned";
        return QOA_B200_E_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->enter(s);
    Scratch sc(ctx);
    const size_t n = enc->descs.size();
    std::vector<EncJob> jobs(n);
    for (size_t i = 0; i < n; ++i) {
        EncJob &j = jobs[i];
        memset(&j, 0, sizeof j);
        if (enc->new_status[i] != QOA_B200_OK || chunk_samples[i] == 0) continue;
        if (out_off[i] & 7) {
            ctx->err = "out_off must be a multiple of 8";
            return QOA_B200_E_ARG;
        }
        j.pcm_off = pcm_off[i];
        j.out_off = out_off[i];
        j.channels = enc->descs[i].channels;
        j.rate = enc->descs[i].sample_rate;
        j.nsamples = chunk_samples[i];
        j.lms_base = enc->lms_base[i];
    }
    return run_encode_jobs(ctx, sc, jobs, d_pcm, d_out, enc->d_lms, s);
}

// Shared body of encode_frame (frame_rules: one frame of <= 5120 samples, :424-430) and encode_chunk (any length).
static int encode_host(qoa_b200_encoder_batch *enc, const int16_t *const *pcm, const size_t *pcm_len,
                       uint8_t *const *out, const size_t *out_cap, size_t *written, int32_t *status, bool frame_rules)
{
    if (!enc) return QOA_B200_E_ARG;
    qoa_b200_ctx *ctx = enc->ctx;
    if (!pcm || !pcm_len || !out || !out_cap || !written || !status) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = enc->ctx->host_stream();
    const size_t n = enc->descs.size();
    std::vector<uint64_t> pcm_off(n, 0), out_off(n, 0);
    std::vector<uint32_t> chunk(n, 0);
    std::vector<size_t> fsize(n, 0);
    uint64_t pcm_total = 0, out_total = 0;
    for (size_t i = 0; i < n; ++i) {
        written[i] = 0;
        if (enc->new_status[i] != QOA_B200_OK) {
            status[i] = enc->new_status[i];
            continue;
        }
        if (pcm[i] == nullptr && pcm_len[i] == 0) continue; // stream not fed this round
        const size_t ch = enc->descs[i].channels;
        // :424-430
        if (pcm_len[i] == 0 || pcm_len[i] % ch != 0 || (frame_rules && pcm_len[i] / ch > QOA_B200_FRAME_LEN) ||
            pcm_len[i] / ch > 0xffffffffull) {
            status[i] = QOA_B200_ENC_INVALID_SAMPLES;
            continue;
        }
        const uint32_t frame_len = (uint32_t)(pcm_len[i] / ch);
        const size_t fs = qoa_b200_encoded_size((uint32_t)ch, frame_len) - QOA_B200_HEADER_SIZE; // whole frames, no file header
        if (out_cap[i] < fs || !out[i]) {
            status[i] = QOA_B200_ENC_IO_ERROR;
            continue;
        }
        status[i] = QOA_B200_OK;
        chunk[i] = frame_len;
        fsize[i] = fs;
        pcm_off[i] = pcm_total;
        pcm_total += (pcm_len[i] + 7) & ~(uint64_t)7;
        out_off[i] = out_total;
        out_total += (fs + 15) & ~(uint64_t)15;
    }
    if (pcm_total == 0) return 0;
    ctx->enter(s);
    {
        Scratch sc(ctx);
        int16_t *d_pcm = sc.get<int16_t>(pcm_total + 8);
        uint8_t *d_out = sc.get<uint8_t>(out_total + 16);
        if (!d_pcm || !d_out) {
            ctx->err = "out of device memory (encode_frame arenas)";
            return QOA_B200_E_NOMEM;
        }
        for (size_t i = 0; i < n; ++i)
            if (chunk[i])
                CHECK_CUDA(ctx, cudaMemcpyAsync(d_pcm + pcm_off[i], pcm[i], pcm_len[i] * sizeof(int16_t), cudaMemcpyHostToDevice, s));
        int rc = qoa_b200_encoder_batch_encode_chunk_dev(enc, d_pcm, pcm_off.data(), chunk.data(), d_out, out_off.data(), s);
        if (rc != 0) return rc;
        for (size_t i = 0; i < n; ++i)
            if (chunk[i]) {
                CHECK_CUDA(ctx, cudaMemcpyAsync(out[i], d_out + out_off[i], fsize[i], cudaMemcpyDeviceToHost, s));
                written[i] = fsize[i];
            }
        CHECK_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return 0;
}

extern "C" int qoa_b200_encoder_batch_encode_frame(qoa_b200_encoder_batch *enc, const int16_t *const *pcm,
                                                   const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                                   size_t *written, int32_t *status)
{
    return encode_host(enc, pcm, pcm_len, out, out_cap, written, status, true);
}

extern "C" int qoa_b200_encoder_batch_encode_chunk(qoa_b200_encoder_batch *enc, const int16_t *const *pcm,
                                                   const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                                   size_t *written, int32_t *status)
{
    return encode_host(enc, pcm, pcm_len, out, out_cap, written, status, false);
}

extern "C" int qoa_b200_encoder_batch_get_lms(qoa_b200_encoder_batch *enc, size_t i, uint32_t c, int32_t history[4],
                                              int32_t weights[4])
{
    if (!enc || i >= enc->descs.size() || enc->new_status[i] != QOA_B200_OK || c >= enc->descs[i].channels)
        return QOA_B200_E_ARG;
    qoa_b200_ctx *ctx = enc->ctx;
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    CHECK_CUDA(ctx, cudaDeviceSynchronize());
    int32_t st[8];
    CHECK_CUDA(ctx, cudaMemcpy(st, enc->d_lms + (size_t)(enc->lms_base[i] + c) * 8, sizeof st, cudaMemcpyDeviceToHost));
    memcpy(history, st, 16);
    memcpy(weights, st + 4, 16);
    return 0;
}

// Large host batches: time-chunked software pipeline.  Chunk t holds the next F frames of EVERY stream (the encoder's
// parallelism is streams x channels, so chunks must keep all streams), the LMS state persists on the device between
// chunks exactly like encode_frame (src/lib.rs:482), and three internal streams overlap the H2D copy of chunk t + 1,
// the kernel of chunk t and the D2H copy of chunk t - 1.  Kernels are chained with events (they share the LMS state).
static int encode_all_batch_pipelined(qoa_b200_ctx *ctx, const qoa_b200_desc *descs, const int16_t *const *pcm,
                                      const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap,
                                      size_t *out_len, int32_t *status, size_t n, uint64_t total_pcm_values)
{
    if (ctx->used) CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->last_stream));
    for (auto &ps : ctx->pipe_streams)
        if (!ps) CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    // validation exactly as the one-shot path
    std::vector<qoa_b200_desc> d(descs, descs + n);
    uint64_t values_per_frame_step = 0; // PCM values one frame of every valid stream holds
    uint32_t max_frames = 0;
    for (size_t i = 0; i < n; ++i) {
        out_len[i] = 0;
        int st = validate_desc(descs[i]);
        if (st == QOA_B200_OK && pcm_len[i] != (size_t)descs[i].samples * descs[i].channels) st = QOA_B200_ENC_INVALID_SAMPLES;
        const size_t need = st == QOA_B200_OK ? qoa_b200_encoded_size(descs[i].channels, descs[i].samples) : 0;
        if (st == QOA_B200_OK && (out_cap[i] < need || !out[i] || !pcm[i])) st = QOA_B200_ENC_IO_ERROR;
        status[i] = st;
        if (st != QOA_B200_OK) {
            d[i].channels = 0; // skipped by the encoder batch
            continue;
        }
        out_len[i] = need;
        values_per_frame_step += (uint64_t)QOA_B200_FRAME_LEN * descs[i].channels;
        max_frames = std::max<uint32_t>(max_frames, (uint32_t)(((uint64_t)descs[i].samples + QOA_B200_FRAME_LEN - 1) / QOA_B200_FRAME_LEN));
    }
    if (values_per_frame_step == 0) return 0;
    (void)total_pcm_values;
    const uint32_t frames_per_chunk = (uint32_t)std::max<uint64_t>(1, enc_chunk_values() / values_per_frame_step);
    qoa_b200_encoder_batch *enc = nullptr;
    std::vector<int32_t> new_status(n);
    int rc = qoa_b200_encoder_batch_new(ctx, d.data(), n, new_status.data(), &enc);
    if (rc != 0) return rc;
    {
        std::vector<uint8_t *> hdr(n, nullptr);
        for (size_t i = 0; i < n; ++i)
            if (status[i] == QOA_B200_OK) hdr[i] = out[i];
        qoa_b200_encoder_batch_write_header(enc, hdr.data());
    }
    ctx->pipelining = true;
    std::vector<cudaEvent_t> events;
    std::vector<uint64_t> pcm_off(n), out_off(n);
    std::vector<uint32_t> chunk(n);
    std::vector<char> pcm_pinned(n, 0), out_pinned(n, 0), known;
    if (n >= 32)
        for (size_t i = 0; i < n; ++i)
            if (status[i] == QOA_B200_OK) {
                pcm_pinned[i] = device_accessible(pcm[i]);
                out_pinned[i] = device_accessible(out[i]);
            }
    for (uint32_t f0 = 0, t = 0; f0 < max_frames && rc == 0; f0 += frames_per_chunk, ++t) {
        cudaStream_t s = ctx->pipe_streams[t % 3];
        ctx->begin_stage((int)(t % 3) + 1); // the arenas of chunk t - 3 (same stream) become reusable
        uint64_t pcm_total = 0, out_total = 0;
        for (size_t i = 0; i < n; ++i) {
            chunk[i] = 0;
            if (status[i] != QOA_B200_OK) continue;
            const uint64_t done = (uint64_t)f0 * QOA_B200_FRAME_LEN;
            if (done >= descs[i].samples) continue;
            chunk[i] = (uint32_t)std::min<uint64_t>((uint64_t)frames_per_chunk * QOA_B200_FRAME_LEN, descs[i].samples - done);
            pcm_off[i] = pcm_total;
            pcm_total += ((uint64_t)chunk[i] * descs[i].channels + 7) & ~7ull;
            out_off[i] = out_total;
            out_total += (qoa_b200_encoded_size(descs[i].channels, chunk[i]) - QOA_B200_HEADER_SIZE + 15) & ~15ull;
        }
        int16_t *d_pcm = static_cast<int16_t *>(ctx->alloc((pcm_total + 8) * sizeof(int16_t)));
        uint8_t *d_out = static_cast<uint8_t *>(ctx->alloc(out_total + 16));
        if (!d_pcm || !d_out) {
            ctx->err = "out of device memory (pipelined encode arenas)";
            rc = QOA_B200_E_NOMEM;
            break;
        }
        ctx->release(d_pcm); // parked until the next stage on this stream
        ctx->release(d_out);
        cudaError_t e = cudaSuccess;
        Scratch sc(ctx);
        std::vector<CopyItem> items;
        known.clear();
        for (size_t i = 0; i < n; ++i)
            if (chunk[i]) {
                items.push_back({pcm[i] + (uint64_t)f0 * QOA_B200_FRAME_LEN * descs[i].channels, d_pcm + pcm_off[i],
                                 (uint64_t)chunk[i] * descs[i].channels * sizeof(int16_t)});
                known.push_back(pcm_pinned[i]);
            }
        rc = copy_batch(ctx, sc, items, true, s, &known);
        if (rc != 0) break;
        if (e == cudaSuccess && !events.empty()) e = cudaStreamWaitEvent(s, events.back(), 0); // LMS state order
        if (e == cudaSuccess) {
            rc = qoa_b200_encoder_batch_encode_chunk_dev(enc, d_pcm, pcm_off.data(), chunk.data(), d_out, out_off.data(), s);
            if (rc != 0) break;
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e == cudaSuccess) {
                events.push_back(ev);
                e = cudaEventRecord(ev, s);
            }
        }
        items.clear();
        known.clear();
        for (size_t i = 0; i < n; ++i)
            if (chunk[i]) {
                const uint32_t ch = descs[i].channels;
                const size_t dst = QOA_B200_HEADER_SIZE + (size_t)f0 * qoa_b200_frame_size(ch, QOA_B200_FRAME_LEN);
                items.push_back({d_out + out_off[i], out[i] + dst, qoa_b200_encoded_size(ch, chunk[i]) - QOA_B200_HEADER_SIZE});
                known.push_back(out_pinned[i]);
            }
        if (e == cudaSuccess) {
            rc = copy_batch(ctx, sc, items, false, s, &known);
            if (rc != 0) break;
        }
        if (e != cudaSuccess) {
            ctx->err = std::string("pipelined encode: ") + cudaGetErrorString(e);
            cudaGetLastError();
            rc = QOA_B200_E_CUDA;
        }
    }
    for (auto &ps : ctx->pipe_streams) {
        cudaError_t e = cudaStreamSynchronize(ps);
        if (e != cudaSuccess && rc == 0) {
            ctx->err = std::string("pipelined encode: ") + cudaGetErrorString(e);
            cudaGetLastError();
            rc = QOA_B200_E_CUDA;
        }
    }
    for (auto ev : events) cudaEventDestroy(ev);
    ctx->end_pipeline();
    qoa_b200_encoder_batch_free(enc);
    return rc;
}

// ================================================================================================================
// decode
// ================================================================================================================

struct qoa_b200_decode_plan {
    qoa_b200_ctx *ctx = nullptr;
    size_t n = 0;
    int mode = 0;
    const uint8_t *d_arena = nullptr; // input bytes (borrowed for *_dev plans, owned otherwise)
    uint8_t *owned_arena = nullptr;
    FrameRec *d_recs = nullptr;
    uint64_t n_frames = 0;
    std::vector<qoa_b200_stream_info> info;
    std::vector<uint64_t> frame_base; // first FrameRec of stream i
    // The frame table is grouped by kernel class so that a mixed batch still runs every stream on its fast kernel:
    // class c = 1..8: streams whose frames all have c channels and start 16-byte aligned in the output;
    // class 0: everything else (per-frame channel changes, > 8 channels, odd-sized interior frames) -> generic kernel.
    std::vector<uint8_t> cls;       // class of stream i
    std::vector<size_t> order;      // stream indices in table order (class 1..8, then 0)
    uint64_t cls_start[9] = {0}, cls_frames[9] = {0};
    uint32_t cls_max_ch[9] = {0};
};

namespace {

int plan_index(qoa_b200_ctx *ctx, qoa_b200_decode_plan *plan, const uint64_t *in_off, const size_t *len, cudaStream_t s,
               int speculate = 1)
{
    const size_t n = plan->n;
    Scratch sc(ctx);
    uint64_t *d_off = sc.get<uint64_t>(n);
    uint64_t *d_len = sc.get<uint64_t>(n);
    uint64_t *d_fbase = sc.get<uint64_t>(n + 1);
    qoa_b200_stream_info *d_info = sc.get<qoa_b200_stream_info>(n);
    // host-mapped results: [int bad flag (64 bytes reserved)] [n x stream_info]
    uint8_t *mapped = static_cast<uint8_t *>(ctx->mapped_get(ctx->slot_of(s), 64 + n * sizeof(qoa_b200_stream_info)));
    if (!d_off || !d_len || !d_fbase || !d_info || !mapped) {
        ctx->err = "out of memory (index tables)";
        return QOA_B200_E_NOMEM;
    }
    volatile int *h_bad = reinterpret_cast<volatile int *>(mapped);
    qoa_b200_stream_info *h_info = reinterpret_cast<qoa_b200_stream_info *>(mapped + 64);
    *h_bad = 0;
    std::vector<uint64_t> len64(len, len + n);
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_off, in_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_len, len64.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CHECK_CUDA(ctx, launch_index_count(plan->d_arena, d_off, d_len, n, plan->mode, speculate, d_info, h_info, s));
    ctx->launches += 1;
    CHECK_CUDA(ctx, cudaStreamSynchronize(s));
    plan->info.assign(h_info, h_info + n);
    plan->frame_base.assign(n, 0);
    plan->cls.assign(n, 0);
    plan->order.clear();
    uint32_t max_regular_frames = 0;
    size_t n_regular = 0;
    for (int c = 0; c < 9; ++c) plan->cls_start[c] = plan->cls_frames[c] = plan->cls_max_ch[c] = 0;
    for (size_t i = 0; i < n; ++i) {
        const auto &inf = plan->info[i];
        if (inf.flags & QOA_B200_INFO_REGULAR) {
            ++n_regular;
            max_regular_frames = std::max(max_regular_frames, inf.index_frames);
        }
        const bool fast = inf.index_frames && inf.max_channels >= 1 && inf.max_channels <= 8 &&
                          (inf.flags & QOA_B200_INFO_UNIFORM_CH) && (inf.flags & QOA_B200_INFO_ALIGNED16);
        plan->cls[i] = fast ? (uint8_t)inf.max_channels : 0;
    }
    uint64_t next = 0;
    static const int kClassOrder[9] = {1, 2, 3, 4, 5, 6, 7, 8, 0};
    for (int c : kClassOrder) {
        plan->cls_start[c] = next;
        for (size_t i = 0; i < n; ++i)
            if (plan->cls[i] == c) {
                plan->frame_base[i] = next;
                next += plan->info[i].index_frames;
                plan->order.push_back(i);
                plan->cls_max_ch[c] = std::max(plan->cls_max_ch[c], plan->info[i].max_channels);
            }
        plan->cls_frames[c] = next - plan->cls_start[c];
    }
    plan->n_frames = next;
    if (plan->n_frames) {
        if (!plan->d_recs) plan->d_recs = static_cast<FrameRec *>(ctx->alloc(plan->n_frames * sizeof(FrameRec)));
        if (!plan->d_recs) {
            ctx->err = "out of device memory (frame index)";
            return QOA_B200_E_NOMEM;
        }
        CHECK_CUDA(ctx, cudaMemcpyAsync(d_fbase, plan->frame_base.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        if (n_regular < n) {
            CHECK_CUDA(ctx, launch_index_fill(plan->d_arena, d_off, d_len, n, plan->mode, d_info, d_fbase, plan->d_recs, s));
            ctx->launches += 1;
        }
        if (n_regular) {
            CHECK_CUDA(ctx, launch_index_fill_regular(plan->d_arena, d_off, n, max_regular_frames, d_info, d_fbase,
                                                      plan->d_recs, const_cast<int *>(h_bad), s));
            ctx->launches += 1;
            CHECK_CUDA(ctx, cudaStreamSynchronize(s));
            if (*h_bad) { // some frame header is not the closed-form one: index the batch by walking
                ctx->release(plan->d_recs);
                plan->d_recs = nullptr;
                return plan_index(ctx, plan, in_off, len, s, 0);
            }
        }
    }
    return 0;
}

} // namespace

extern "C" int qoa_b200_probe(const uint8_t *qoa, size_t len, int mode, qoa_b200_stream_info *info)
{
    if (!info || (len && !qoa) || mode < QOA_B200_DECODE_ITER || mode > QOA_B200_DECODE_FRAMES) return QOA_B200_E_ARG;
    qoa::probe_stream(qoa, len, mode, info);
    return 0;
}

extern "C" int qoa_b200_decode_plan_create_dev(qoa_b200_ctx *ctx, const uint8_t *d_qoa, const uint64_t *in_off,
                                               const size_t *len, size_t n, int mode, void *stream,
                                               qoa_b200_decode_plan **out)
{
    if (!ctx || !out) return QOA_B200_E_ARG;
    *out = nullptr;
    if (n && (!d_qoa || !in_off || !len)) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    if (mode < QOA_B200_DECODE_ITER || mode > QOA_B200_DECODE_FRAMES) {
        ctx->err = "bad decode mode";
        return QOA_B200_E_ARG;
    }
    if (n && (reinterpret_cast<uintptr_t>(d_qoa) & 15)) { // the kernels read the arena with 16-byte vectors
        ctx->err = "d_qoa must be 16-byte aligned";
        return QOA_B200_E_ARG;
    }
    for (size_t i = 0; i < n; ++i)
        if (in_off[i] & 15) {
            ctx->err = "in_off must be a multiple of 16";
            return QOA_B200_E_ARG;
        }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->enter(s);
    auto *plan = new qoa_b200_decode_plan();
    plan->ctx = ctx;
    plan->n = n;
    plan->mode = mode;
    plan->d_arena = d_qoa;
    if (n) {
        int rc = plan_index(ctx, plan, in_off, len, s);
        if (rc != 0) {
            qoa_b200_decode_plan_destroy(plan);
            return rc;
        }
    }
    *out = plan;
    return 0;
}

static int plan_create_host(qoa_b200_ctx *ctx, const uint8_t *const *qoa, const size_t *len, size_t n, int mode,
                            cudaStream_t s, qoa_b200_decode_plan **out)
{
    if (!ctx || !out) return QOA_B200_E_ARG;
    *out = nullptr;
    if (n && (!qoa || !len)) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->enter(s);
    std::vector<uint64_t> off(n, 0);
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        off[i] = total;
        total += (len[i] + 15) & ~(uint64_t)15;
    }
    uint8_t *arena = static_cast<uint8_t *>(ctx->alloc(total + 16));
    if (!arena) {
        ctx->err = "out of device memory (input arena)";
        return QOA_B200_E_NOMEM;
    }
    {
        Scratch sc(ctx);
        std::vector<CopyItem> items;
        for (size_t i = 0; i < n; ++i)
            if (len[i]) items.push_back({qoa[i], arena + off[i], len[i]});
        int rcc = copy_batch(ctx, sc, items, true, s);
        if (rcc != 0) {
            ctx->release(arena);
            return rcc;
        }
    }
    int rc = qoa_b200_decode_plan_create_dev(ctx, arena, off.data(), len, n, mode, s, out);
    if (rc != 0) {
        ctx->release(arena);
        return rc;
    }
    (*out)->owned_arena = arena;
    return 0;
}

extern "C" int qoa_b200_decode_plan_create(qoa_b200_ctx *ctx, const uint8_t *const *qoa, const size_t *len, size_t n,
                                           int mode, qoa_b200_decode_plan **out)
{
    if (!ctx || !out) return QOA_B200_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) cudaGetLastError(); // reported by plan_create_host
    return plan_create_host(ctx, qoa, len, n, mode, ctx->host_stream(), out);
}

extern "C" void qoa_b200_decode_plan_destroy(qoa_b200_decode_plan *plan)
{
    if (!plan) return;
    plan->ctx->release(plan->d_recs);
    plan->ctx->release(plan->owned_arena);
    delete plan;
}

extern "C" int qoa_b200_decode_plan_info(const qoa_b200_decode_plan *plan, qoa_b200_stream_info *info)
{
    if (!plan || (plan->n && !info)) return QOA_B200_E_ARG;
    if (plan->n) memcpy(info, plan->info.data(), plan->n * sizeof(qoa_b200_stream_info));
    return 0;
}

extern "C" size_t qoa_b200_decode_plan_frame_headers(const qoa_b200_decode_plan *plan, size_t i,
                                                     qoa_b200_frame_header *hdrs, size_t cap)
{
    if (!plan || i >= plan->n) return 0;
    const uint64_t f0 = plan->frame_base[i];
    const size_t count = plan->info[i].index_frames;
    const size_t m = std::min(count, cap);
    if (m && hdrs) {
        std::vector<FrameRec> recs(m);
        cudaSetDevice(plan->ctx->device);
        if (cudaStreamSynchronize(plan->ctx->last_stream) != cudaSuccess ||
            cudaMemcpy(recs.data(), plan->d_recs + f0, m * sizeof(FrameRec), cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        for (size_t k = 0; k < m; ++k) {
            hdrs[k].num_channels = recs[k].ch;
            hdrs[k].sample_rate = recs[k].rate;
            hdrs[k].num_samples_per_channel = recs[k].nspc;
        }
    }
    return count;
}

extern "C" int qoa_b200_decode_plan_run_dev(qoa_b200_decode_plan *plan, int16_t *d_out, const uint64_t *out_off,
                                            const size_t *out_cap, int32_t *status, void *stream)
{
    if (!plan) return QOA_B200_E_ARG;
    qoa_b200_ctx *ctx = plan->ctx;
    const size_t n = plan->n;
    if (n == 0) return 0;
    if (!d_out || !out_off || !out_cap) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->enter(s);
    Scratch sc(ctx);
    std::vector<uint64_t> base(n);
    bool cls_aligned[9] = {true, true, true, true, true, true, true, true, true};
    if (reinterpret_cast<uintptr_t>(d_out) & 15) // the fast kernels store 16-byte rows: a misaligned arena (e.g. a view with
        for (bool &a : cls_aligned) a = false;   // a storage offset) takes the generic kernel, which stores single samples
    if (reinterpret_cast<uintptr_t>(d_out) & 1) {
        ctx->err = "d_out must be 2-byte aligned";
        return QOA_B200_E_ARG;
    }
    uint64_t out_bytes = 0;
    for (size_t i = 0; i < n; ++i) {
        int st = plan->info[i].status;
        if (plan->info[i].samples > out_cap[i]) {
            st = QOA_B200_DEC_OUTPUT_TOO_SMALL;
            base[i] = ~0ull;
        } else {
            base[i] = out_off[i];
            if (out_off[i] & 7) cls_aligned[plan->cls[i]] = false;
            out_bytes = std::max<uint64_t>(out_bytes, (out_off[i] + plan->info[i].samples) * 2);
        }
        if (status) status[i] = st;
    }
    if (plan->n_frames == 0) return 0;
    uint64_t *d_base = sc.get<uint64_t>(n);
    if (!d_base) {
        ctx->err = "out of device memory (output table)";
        return QOA_B200_E_NOMEM;
    }
    CHECK_CUDA(ctx, cudaMemcpyAsync(d_base, base.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    int n_cls = 0;
    for (int c = 0; c < 9; ++c) n_cls += plan->cls_frames[c] != 0;
    // several classes: fork the launches over side streams (not inside a pipelined host call, whose stages already overlap)
    const bool fork = n_cls > 1 && !ctx->pipelining && ctx->cls_ready();
    if (fork) CHECK_CUDA(ctx, cudaEventRecord(ctx->cls_fork, s));
    int k = 0;
    bool used[3] = {false, false, false};
    cudaError_t e = cudaSuccess; // no early return between fork and join: a side stream is never left unjoined
    for (int c = 0; c < 9 && e == cudaSuccess; ++c) { // one launch per kernel class present in the batch
        if (!plan->cls_frames[c]) continue;
        cudaStream_t ls = s;
        if (fork && k > 0) {
            const int side = (k - 1) % 3;
            if (!used[side] && (e = cudaStreamWaitEvent(ctx->cls_streams[side], ctx->cls_fork, 0)) != cudaSuccess) break;
            ls = ctx->cls_streams[side];
            used[side] = true;
        }
        ++k;
        int nl = 0;
        e = launch_decode(plan->d_arena, plan->d_recs + plan->cls_start[c], plan->cls_frames[c], plan->cls_max_ch[c],
                          (c && cls_aligned[c]) ? (uint32_t)c : 0u, d_base, d_out, out_bytes, ls, &nl);
        ctx->launches += nl;
    }
    for (int side = 0; side < 3; ++side) // join: everything the caller enqueues behind this call sees every class done
        if (used[side]) {
            cudaError_t ej = cudaEventRecord(ctx->cls_join[side], ctx->cls_streams[side]);
            if (ej == cudaSuccess) ej = cudaStreamWaitEvent(s, ctx->cls_join[side], 0);
            if (e == cudaSuccess) e = ej;
        }
    if (e != cudaSuccess) {
        ctx->err = std::string("decode launch: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return QOA_B200_E_CUDA;
    }
    return 0;
}

// Decode + D2H on stream s.  With sync == false the copies are left in flight (the caller drains the stream).
static int plan_run_host(qoa_b200_decode_plan *plan, int16_t *const *out, const size_t *out_cap, int32_t *status,
                         cudaStream_t s, bool sync)
{
    if (!plan) return QOA_B200_E_ARG;
    qoa_b200_ctx *ctx = plan->ctx;
    const size_t n = plan->n;
    if (n == 0) return 0;
    if (!out || !out_cap) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint64_t> off(n, 0);
    uint64_t total = 0;
    // The arena follows the frame table's order (grouped by kernel class).  Mono / stereo streams get slots that start
    // on a multiple of one full frame of PCM (5120 x channels values): frames of consecutive streams then stay on the
    // 2-D tensor map's row lattice, so boxes that span a stream boundary still leave through one tensor store (+4 %).
    for (size_t i : plan->order) {
        const uint64_t slot = (plan->cls[i] == 1 || plan->cls[i] == 2) ? (uint64_t)QOA_B200_FRAME_LEN * plan->cls[i] : 8;
        total = (total + slot - 1) / slot * slot;
        off[i] = total;
        if (plan->info[i].samples <= out_cap[i]) total += (plan->info[i].samples + 7) & ~7ull;
    }
    int16_t *d_out = static_cast<int16_t *>(ctx->alloc((total + 8) * sizeof(int16_t)));
    if (!d_out) {
        ctx->err = "out of device memory (output arena)";
        return QOA_B200_E_NOMEM;
    }
    int rc = qoa_b200_decode_plan_run_dev(plan, d_out, off.data(), out_cap, status, s);
    if (rc == 0) {
        {
            Scratch sc(ctx);
            std::vector<CopyItem> items;
            for (size_t i = 0; i < n; ++i) {
                const uint64_t cnt = plan->info[i].samples;
                if (cnt == 0 || cnt > out_cap[i] || !out[i]) continue;
                items.push_back({d_out + off[i], out[i], cnt * sizeof(int16_t)});
            }
            rc = copy_batch(ctx, sc, items, false, s);
        }
        cudaError_t e = sync ? cudaStreamSynchronize(s) : cudaSuccess;
        if (e != cudaSuccess && rc == 0) {
            ctx->err = std::string("decode: ") + cudaGetErrorString(e);
            cudaGetLastError();
            rc = QOA_B200_E_CUDA;
        }
    }
    ctx->release(d_out); // deferred while pipelining
    return rc;
}

extern "C" int qoa_b200_decode_plan_run(qoa_b200_decode_plan *plan, int16_t *const *out, const size_t *out_cap,
                                        int32_t *status)
{
    if (!plan) return QOA_B200_E_ARG;
    return plan_run_host(plan, out, out_cap, status, plan->ctx->last_stream, true);
}

extern "C" int qoa_b200_decode_all_batch(qoa_b200_ctx *ctx, const uint8_t *const *qoa, const size_t *len, size_t n,
                                         int mode, int16_t *const *out, const size_t *out_cap,
                                         qoa_b200_stream_info *info, int32_t *status)
{
    if (!ctx) return QOA_B200_E_ARG;
    if (n && (!qoa || !len)) {
        ctx->err = "null argument";
        return QOA_B200_E_ARG;
    }
    uint64_t total_in = 0;
    for (size_t i = 0; i < n; ++i) total_in += len[i];
    constexpr uint64_t kChunkBytes = 24ull << 20; // input bytes per pipeline stage (~5x that in PCM coming back)
    if (n < 8 || total_in < 3 * kChunkBytes) { // small batch: one shot
        qoa_b200_decode_plan *plan = nullptr;
        int rc = qoa_b200_decode_plan_create(ctx, qoa, len, n, mode, &plan);
        if (rc != 0) return rc;
        if (info) qoa_b200_decode_plan_info(plan, info);
        rc = qoa_b200_decode_plan_run(plan, out, out_cap, status);
        qoa_b200_decode_plan_destroy(plan);
        return rc;
    }
    // Large batch: software pipeline over chunks of streams on three internal streams, so that the H2D copy of chunk
    // c + 1, the index + decode of chunk c and the D2H copy of chunk c - 1 overlap (PCIe is full duplex; the path is
    // D2H-bound: 2 bytes out per 0.4 bytes in).
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->used) CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->last_stream));
    for (auto &ps : ctx->pipe_streams)
        if (!ps) CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    ctx->pipelining = true;
    std::vector<qoa_b200_decode_plan *> plans;
    int rc = 0;
    size_t b = 0;
    for (int c = 0; b < n && rc == 0; ++c) {
        size_t e = b;
        uint64_t bytes = 0;
        while (e < n && (e == b || bytes + len[e] <= kChunkBytes)) bytes += len[e++];
        cudaStream_t s = ctx->pipe_streams[c % 3];
        ctx->cur_slot = c % 3 + 1;
        if (c >= 3) { // the plan of chunk c - 3 ran on this stream: its arenas are reusable by this chunk (stream order)
            qoa_b200_decode_plan_destroy(plans[c - 3]);
            plans[c - 3] = nullptr;
        }
        ctx->begin_stage(c % 3 + 1);
        qoa_b200_decode_plan *plan = nullptr;
        rc = plan_create_host(ctx, qoa + b, len + b, e - b, mode, s, &plan);
        if (rc != 0) break;
        plans.push_back(plan);
        if (info) qoa_b200_decode_plan_info(plan, info + b);
        rc = plan_run_host(plan, out + b, out_cap + b, status ? status + b : nullptr, s, false);
        b = e;
    }
    for (auto &ps : ctx->pipe_streams) {
        cudaError_t e = cudaStreamSynchronize(ps);
        if (e != cudaSuccess && rc == 0) {
            ctx->err = std::string("decode_all_batch: ") + cudaGetErrorString(e);
            cudaGetLastError();
            rc = QOA_B200_E_CUDA;
        }
    }
    for (auto *p : plans)
        if (p) qoa_b200_decode_plan_destroy(p);
    ctx->end_pipeline();
    return rc;
}

// ================================================================================================================
// generators / microbenchmarks
// ================================================================================================================

extern "C" int qoa_b200_gen_pcm_dev(qoa_b200_ctx *ctx, int16_t *d_pcm, uint64_t n_streams, uint64_t first_stream,
                                    uint32_t channels, uint64_t samples, uint64_t start, uint64_t seed, int kind,
                                    void *stream)
{
    if (!ctx || !d_pcm) return QOA_B200_E_ARG;
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    CHECK_CUDA(ctx, launch_gen_pcm(d_pcm, n_streams, first_stream, channels, samples, start, seed, kind,
                                   static_cast<cudaStream_t>(stream)));
    return 0;
}

extern "C" int qoa_b200_gen_qoa_dev(qoa_b200_ctx *ctx, uint8_t *d_qoa, uint64_t stride, uint64_t n_streams,
                                    uint64_t first_stream, uint32_t channels, uint32_t sample_rate, uint32_t samples,
                                    uint64_t seed, void *stream)
{
    if (!ctx || !d_qoa || channels == 0 || channels > 255 || (stride & 15) ||
        stride < qoa_b200_encoded_size(channels, samples))
        return QOA_B200_E_ARG;
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    CHECK_CUDA(ctx, launch_gen_qoa(d_qoa, stride, n_streams, first_stream, channels, sample_rate, samples, seed,
                                   static_cast<cudaStream_t>(stream)));
    return 0;
}

extern "C" int qoa_b200_int_issue_peak(qoa_b200_ctx *ctx, int kind, double *thread_ops_per_s)
{
    if (!ctx || !thread_ops_per_s) return QOA_B200_E_ARG;
    CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    CHECK_CUDA(ctx, run_int_issue_peak(kind, ctx->sm_count, thread_ops_per_s));
    return 0;
}
