// qoa_multi.cu — one context that drives several GPUs of a box from ONE process.
//
// The QOA path has no cross-stream dependency (QoaEncoder / QoaDecoder hold only their own state, src/lib.rs:59-68,
// :85-91), so a batch is cut into contiguous ranges of streams, one per device, balanced by work, and every device
// runs the single-device host path (qoa_b200_encode_all_batch / qoa_b200_decode_all_batch: pinned-buffer pipelines on
// three CUDA streams) on its range from its own host thread.  No collective, no peer access, results land in the
// caller's arrays in caller order because the ranges are contiguous.  This is what a crate user who binds
// encode_all (src/lib.rs:760) / decode_all (:713) for a batch gets on an 8-GPU box.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "qoa_kernels.h"

struct qoa_b200_mctx {
    std::vector<qoa_b200_ctx *> ctx;
    std::vector<int> devices;
    std::string err;
};

extern "C" {

// Splits range(n) into `parts` contiguous ranges whose summed weights are as even as one left-to-right sweep can make
// them: first[p] .. first[p + 1] is range p (parts + 1 entries; ranges may be empty).  Pure host arithmetic (no GPU).
int qoa_b200_partition(const uint64_t *weight, size_t n, int parts, size_t *first)
{
    if (parts <= 0 || !first || (n && !weight)) return QOA_B200_E_ARG;
    long double total = 0;
    for (size_t i = 0; i < n; ++i) total += (long double)weight[i];
    long double acc = 0;
    size_t i = 0;
    first[0] = 0;
    for (int p = 1; p < parts; ++p) {
        const long double target = total * p / parts;
        while (i < n && acc + (long double)weight[i] / 2 <= target) acc += (long double)weight[i++];
        first[p] = i;
    }
    first[parts] = n;
    return 0;
}

int qoa_b200_ctx_create_multi(const int *devices, int n, qoa_b200_mctx **out)
{
    if (!out) return QOA_B200_E_ARG;
    *out = nullptr;
    std::vector<int> devs;
    if (devices && n > 0) {
        devs.assign(devices, devices + n);
    } else { // every visible device
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess) {
            cudaGetLastError();
            count = 0;
        }
        for (int d = 0; d < count; ++d) devs.push_back(d);
    }
    if (devs.empty()) return QOA_B200_E_CUDA;
    auto *m = new qoa_b200_mctx();
    for (int d : devs) {
        qoa_b200_ctx *c = nullptr;
        const int rc = qoa_b200_ctx_create(d, &c);
        if (rc != 0) { // qoa_b200_last_error(NULL) holds the reason (this thread)
            for (auto *p : m->ctx) qoa_b200_ctx_destroy(p);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
        m->devices.push_back(d);
    }
    *out = m;
    return 0;
}

void qoa_b200_mctx_destroy(qoa_b200_mctx *m)
{
    if (!m) return;
    for (auto *c : m->ctx) qoa_b200_ctx_destroy(c);
    delete m;
}

int qoa_b200_mctx_device_count(const qoa_b200_mctx *m) { return m ? (int)m->ctx.size() : 0; }
int qoa_b200_mctx_device(const qoa_b200_mctx *m, int i) { return (m && i >= 0 && i < (int)m->devices.size()) ? m->devices[i] : -1; }
qoa_b200_ctx *qoa_b200_mctx_ctx(qoa_b200_mctx *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
const char *qoa_b200_mctx_last_error(const qoa_b200_mctx *m) { return m ? m->err.c_str() : qoa_b200_last_error(nullptr); }
uint64_t qoa_b200_mctx_kernel_launches(const qoa_b200_mctx *m)
{
    uint64_t t = 0;
    if (m)
        for (auto *c : m->ctx) t += qoa_b200_kernel_launches(c);
    return t;
}

} // extern "C"

namespace {

// Runs fn(device slot, first stream, stream count) for every non-empty range, one host thread per device; returns the
// first non-zero call status and records which device it came from.
template <typename F>
int run_sharded(qoa_b200_mctx *m, const std::vector<size_t> &first, F fn)
{
    const int parts = (int)m->ctx.size();
    std::vector<int> rc(parts, 0);
    std::vector<std::thread> th;
    for (int p = 1; p < parts; ++p)
        if (first[p + 1] > first[p]) th.emplace_back([&, p] { rc[p] = fn(p, first[p], first[p + 1] - first[p]); });
    if (first[1] > first[0]) rc[0] = fn(0, first[0], first[1] - first[0]); // the caller's thread serves the first device
    for (auto &t : th) t.join();
    for (int p = 0; p < parts; ++p)
        if (rc[p] != 0) {
            m->err = "device " + std::to_string(m->devices[p]) + ": " + qoa_b200_last_error(m->ctx[p]);
            return rc[p];
        }
    return 0;
}

} // namespace

extern "C" {

int qoa_b200_mctx_partition(const qoa_b200_mctx *m, const uint64_t *weight, size_t n, size_t *first)
{
    if (!m) return QOA_B200_E_ARG;
    return qoa_b200_partition(weight, n, (int)m->ctx.size(), first);
}

int qoa_b200_mctx_encode_all_batch(qoa_b200_mctx *m, const qoa_b200_desc *descs, const int16_t *const *pcm,
                                   const size_t *pcm_len, uint8_t *const *out, const size_t *out_cap, size_t *out_len,
                                   int32_t *status, size_t n)
{
    if (!m) return QOA_B200_E_ARG;
    if (n == 0) return 0;
    if (!descs || !pcm || !pcm_len || !out || !out_cap || !out_len || !status) {
        m->err = "null argument";
        return QOA_B200_E_ARG;
    }
    std::vector<uint64_t> w(n);
    for (size_t i = 0; i < n; ++i) w[i] = std::max<uint64_t>(1, (uint64_t)descs[i].samples * descs[i].channels); // 16-candidate search per sample
    std::vector<size_t> first(m->ctx.size() + 1);
    qoa_b200_partition(w.data(), n, (int)m->ctx.size(), first.data());
    return run_sharded(m, first, [&](int p, size_t b, size_t cnt) {
        return qoa_b200_encode_all_batch(m->ctx[p], descs + b, pcm + b, pcm_len + b, out + b, out_cap + b, out_len + b,
                                         status + b, cnt);
    });
}

int qoa_b200_mctx_decode_all_batch(qoa_b200_mctx *m, const uint8_t *const *qoa, const size_t *len, size_t n, int mode,
                                   int16_t *const *out, const size_t *out_cap, qoa_b200_stream_info *info,
                                   int32_t *status)
{
    if (!m) return QOA_B200_E_ARG;
    if (n == 0) return 0;
    if (!qoa || !len) {
        m->err = "null argument";
        return QOA_B200_E_ARG;
    }
    std::vector<uint64_t> w(n);
    for (size_t i = 0; i < n; ++i) w[i] = std::max<uint64_t>(1, len[i]); // decode work ~ slice count ~ encoded length
    std::vector<size_t> first(m->ctx.size() + 1);
    qoa_b200_partition(w.data(), n, (int)m->ctx.size(), first.data());
    return run_sharded(m, first, [&](int p, size_t b, size_t cnt) {
        return qoa_b200_decode_all_batch(m->ctx[p], qoa + b, len + b, cnt, mode, out ? out + b : nullptr,
                                         out_cap ? out_cap + b : nullptr, info ? info + b : nullptr,
                                         status ? status + b : nullptr);
    });
}

} // extern "C"
