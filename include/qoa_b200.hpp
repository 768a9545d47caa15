// qoa_b200.hpp — C++ host-side mirror of the `qoaudio` crate's API over the C ABI in qoa_b200.h.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host side a user of the crate would
// write against is offered in C++ with the crate's names, argument meaning and error behaviour (src/lib.rs of
// rafaelcaricio/qoaudio; line numbers cited per item).  Header-only, C++17, links against libqoa_b200.so.
// Every call runs on the GPU; there is no CPU path behind these functions.
//
//   crate item (src/lib.rs)                          here
//   QOA_* constants (:12-19)                         qoaudio::QOA_*
//   QoaDesc (:94-102)                                qoaudio::QoaDesc
//   EncodeError (:105-115), DecodeError (:866-873)   qoaudio::EncodeError / DecodeError (.kind)
//   ProcessingMode (:30-45)                          qoaudio::ProcessingMode
//   FrameHeader (:773-781), QoaItem (:655-659)       qoaudio::FrameHeader / QoaItem
//   encode_all (:760)                                qoaudio::encode_all (+ encode_all_batch)
//   QoaEncoder::{new, encode, write_header,          qoaudio::QoaEncoder
//                encode_frame} (:335-439)
//   QoaDecoder::{new, open, new_streaming,           qoaudio::QoaDecoder (next() replaces Iterator::next)
//                decode_frame, mode, ...} (:127-201, :619-693)
//   decode_all (:713), open_and_decode_all (:750)    qoaudio::decode_all / open_and_decode_all (+ decode_all_batch)
//   DecodedQoa (:696-704)                            qoaudio::DecodedQoa
#ifndef QOA_B200_HPP
#define QOA_B200_HPP

#include <cstdint>
#include <fstream>
#include <istream>
#include <iterator>
#include <memory>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "qoa_b200.h"

namespace qoaudio {

constexpr uint32_t QOA_SLICE_LEN = QOA_B200_SLICE_LEN;
constexpr uint32_t QOA_LMS_LEN = QOA_B200_LMS_LEN;
constexpr uint32_t QOA_HEADER_SIZE = QOA_B200_HEADER_SIZE;
constexpr uint32_t QOA_MAGIC = QOA_B200_MAGIC;
constexpr uint32_t QOA_SLICES_PER_FRAME = QOA_B200_SLICES_PER_FRAME;
constexpr uint32_t QOA_FRAME_LEN = QOA_B200_FRAME_LEN;
constexpr uint32_t QOA_MAX_CHANNELS = QOA_B200_MAX_CHANNELS;

struct QoaDesc { // :94-102
    uint8_t channels = 0;
    uint32_t sample_rate = 0;
    uint32_t samples = 0; // per channel
};

// A call-level failure of the engine itself (no device, out of memory, CUDA error): not part of the crate's API.
class EngineError : public std::runtime_error {
public:
    EngineError(int code, const std::string &what) : std::runtime_error(what), code(code) {}
    int code;
};

class EncodeError : public std::runtime_error { // :105-115
public:
    enum Kind { InvalidChannels = 1, InvalidSampleRate = 2, InvalidSamples = 3, IoError = 4 };
    explicit EncodeError(Kind k, const std::string &detail = "") : std::runtime_error(message(k, detail)), kind(k) {}
    Kind kind;

private:
    static std::string message(Kind k, const std::string &detail)
    {
        switch (k) {
        case InvalidChannels: return "Invalid number of channels (must be 1-8)";
        case InvalidSampleRate: return "Invalid sample rate (must be > 0)";
        case InvalidSamples: return "Invalid number of samples (must be > 0)";
        default: return "IO error: " + detail;
        }
    }
};

class DecodeError : public std::runtime_error { // :866-873
public:
    enum Kind { NotQoaFile = 11, NoSamples = 12, InvalidFrameHeader = 13, IncompatibleFrame = 14, IoError = 15 };
    explicit DecodeError(Kind k, const std::string &detail = "failed to fill whole buffer")
        : std::runtime_error(message(k, detail)), kind(k)
    {
    }
    Kind kind;
    static Kind from_status(int32_t st) { return st == QOA_B200_DEC_OUTPUT_TOO_SMALL ? IoError : static_cast<Kind>(st); }

private:
    static std::string message(Kind k, const std::string &detail)
    {
        switch (k) {
        case NotQoaFile: return "File is not a qoa file";
        case NoSamples: return "File has no samples";
        case InvalidFrameHeader: return "File has invalid frame header";
        case IncompatibleFrame: return "Incompatible frame header";
        default: return "IO error: " + detail;
        }
    }
};

struct FrameHeader { // :773-781
    uint8_t num_channels = 0;
    uint32_t sample_rate = 0;
    uint16_t num_samples_per_channel = 0;
    bool operator==(const FrameHeader &o) const
    {
        return num_channels == o.num_channels && sample_rate == o.sample_rate &&
               num_samples_per_channel == o.num_samples_per_channel;
    }
};

struct ProcessingMode { // :30-45: FixedSamples{channels, sample_rate, samples} | Streaming
    bool streaming = true;
    uint8_t channels = 0;
    uint32_t sample_rate = 0;
    uint32_t samples = 0;
    static ProcessingMode Streaming() { return {}; }
    static ProcessingMode FixedSamples(uint8_t ch, uint32_t rate, uint32_t samples) { return {false, ch, rate, samples}; }
    bool operator==(const ProcessingMode &o) const
    {
        return streaming == o.streaming && (streaming || (channels == o.channels && sample_rate == o.sample_rate && samples == o.samples));
    }
};

using QoaItem = std::variant<int16_t, FrameHeader>; // :655-659: Sample(i16) | FrameHeader(FrameHeader)

struct DecodedQoa { // :696-704
    uint8_t num_channels = 0;
    uint32_t sample_rate = 0;
    std::vector<int16_t> samples;
};

// One engine context per GPU (RAII over qoa_b200_ctx).  Engine::shared() is a process-wide context on device 0.
class Engine {
public:
    explicit Engine(int device = 0)
    {
        const int rc = qoa_b200_ctx_create(device, &ctx_);
        if (rc != 0) throw EngineError(rc, std::string("qoa_b200_ctx_create: ") + qoa_b200_last_error(nullptr));
    }
    ~Engine() { qoa_b200_ctx_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    qoa_b200_ctx *handle() const { return ctx_; }
    void check(int rc, const char *what) const
    {
        if (rc != 0) throw EngineError(rc, std::string(what) + ": " + qoa_b200_last_error(ctx_));
    }
    static Engine &shared()
    {
        static Engine e(0);
        return e;
    }

private:
    qoa_b200_ctx *ctx_ = nullptr;
};

// Every GPU of the box from one process: batches are cut into contiguous ranges of streams, one per device, balanced by
// work; no collective (QoaEncoder / QoaDecoder hold only their own state, :59-68, :85-91).  Pass it to encode_all_batch /
// decode_all_batch instead of an Engine.
class MultiEngine {
public:
    explicit MultiEngine(const std::vector<int> &devices = {}) // empty: every visible device
    {
        const int rc = qoa_b200_ctx_create_multi(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &m_);
        if (rc != 0) throw EngineError(rc, std::string("qoa_b200_ctx_create_multi: ") + qoa_b200_mctx_last_error(nullptr));
    }
    ~MultiEngine() { qoa_b200_mctx_destroy(m_); }
    MultiEngine(const MultiEngine &) = delete;
    MultiEngine &operator=(const MultiEngine &) = delete;
    qoa_b200_mctx *handle() const { return m_; }
    int device_count() const { return qoa_b200_mctx_device_count(m_); }
    void check(int rc, const char *what) const
    {
        if (rc != 0) throw EngineError(rc, std::string(what) + ": " + qoa_b200_mctx_last_error(m_));
    }

private:
    qoa_b200_mctx *m_ = nullptr;
};

namespace detail {
inline qoa_b200_desc c_desc(const QoaDesc &d) { return {d.channels, d.sample_rate, d.samples}; }
inline void validate(const QoaDesc &d) // QoaEncoder::new, :336-345 (this order)
{
    if (d.channels == 0 || d.channels > QOA_MAX_CHANNELS) throw EncodeError(EncodeError::InvalidChannels);
    if (d.sample_rate == 0) throw EncodeError(EncodeError::InvalidSampleRate);
    if (d.samples == 0) throw EncodeError(EncodeError::InvalidSamples);
}
inline std::vector<uint8_t> read_all(std::istream &in)
{
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
}
inline void put_file_header(uint8_t *p, uint32_t samples) // :404-408
{
    p[0] = 'q', p[1] = 'o', p[2] = 'a', p[3] = 'f';
    p[4] = uint8_t(samples >> 24), p[5] = uint8_t(samples >> 16), p[6] = uint8_t(samples >> 8), p[7] = uint8_t(samples);
}
} // namespace detail

// ---- encode ----------------------------------------------------------------------------------------------------

using EncodeResult = std::variant<std::vector<uint8_t>, EncodeError>;

namespace detail {
inline int call_encode_all(Engine &e, const qoa_b200_desc *d, const int16_t *const *pcm, const size_t *pl, uint8_t *const *out,
                           const size_t *cap, size_t *len, int32_t *st, size_t n)
{
    return qoa_b200_encode_all_batch(e.handle(), d, pcm, pl, out, cap, len, st, n);
}
inline int call_encode_all(MultiEngine &e, const qoa_b200_desc *d, const int16_t *const *pcm, const size_t *pl,
                           uint8_t *const *out, const size_t *cap, size_t *len, int32_t *st, size_t n)
{
    return qoa_b200_mctx_encode_all_batch(e.handle(), d, pcm, pl, out, cap, len, st, n);
}
} // namespace detail

// Batched encode_all: one result per stream; a failing stream never aborts the batch.  EngineT = Engine | MultiEngine.
template <typename EngineT>
inline std::vector<EncodeResult> encode_all_batch(const std::vector<std::pair<const std::vector<int16_t> *, QoaDesc>> &items,
                                                  EngineT &eng)
{
    const size_t n = items.size();
    std::vector<qoa_b200_desc> descs(n);
    std::vector<const int16_t *> pcm(n);
    std::vector<size_t> pcm_len(n), cap(n), len(n);
    std::vector<std::vector<uint8_t>> out(n);
    std::vector<uint8_t *> outp(n);
    std::vector<int32_t> status(n);
    for (size_t i = 0; i < n; ++i) {
        const QoaDesc &d = items[i].second;
        descs[i] = detail::c_desc(d);
        pcm[i] = items[i].first->data();
        pcm_len[i] = items[i].first->size();
        const bool ok = d.channels >= 1 && d.channels <= QOA_MAX_CHANNELS && d.sample_rate && d.samples;
        out[i].resize(ok ? qoa_b200_encoded_size(d.channels, d.samples) : 0);
        outp[i] = out[i].data();
        cap[i] = out[i].size();
    }
    if (n)
        eng.check(detail::call_encode_all(eng, descs.data(), pcm.data(), pcm_len.data(), outp.data(), cap.data(), len.data(),
                                          status.data(), n),
                  "qoa_b200_encode_all_batch");
    std::vector<EncodeResult> res;
    res.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != QOA_B200_OK) {
            res.emplace_back(EncodeError(static_cast<EncodeError::Kind>(status[i])));
        } else {
            out[i].resize(len[i]);
            res.emplace_back(std::move(out[i]));
        }
    }
    return res;
}

inline std::vector<EncodeResult> encode_all_batch(const std::vector<std::pair<const std::vector<int16_t> *, QoaDesc>> &items)
{
    return encode_all_batch(items, Engine::shared());
}

inline std::vector<uint8_t> encode_all(const std::vector<int16_t> &sample_data, const QoaDesc &desc,
                                       Engine &eng = Engine::shared()) // :760-763
{
    auto res = encode_all_batch({{&sample_data, desc}}, eng);
    if (auto *e = std::get_if<EncodeError>(&res[0])) throw *e;
    return std::move(std::get<std::vector<uint8_t>>(res[0]));
}

class QoaEncoder { // :85-91, :333-439; the LMS state lives on the GPU between encode_frame calls
public:
    explicit QoaEncoder(const QoaDesc &desc, Engine &eng = Engine::shared()) : desc_(desc), eng_(&eng)
    {
        detail::validate(desc);
    }
    ~QoaEncoder()
    {
        if (batch_) qoa_b200_encoder_batch_free(batch_);
    }
    QoaEncoder(const QoaEncoder &) = delete;
    QoaEncoder &operator=(const QoaEncoder &) = delete;

    // :367-398 — encoding starts from this encoder's CURRENT LMS state, as in the reference
    std::vector<uint8_t> encode(const std::vector<int16_t> &sample_data)
    {
        if (sample_data.size() != size_t(desc_.samples) * desc_.channels) throw EncodeError(EncodeError::InvalidSamples);
        std::vector<uint8_t> out(qoa_b200_encoded_size(desc_.channels, desc_.samples));
        detail::put_file_header(out.data(), desc_.samples);
        const int16_t *pcm = sample_data.data();
        const size_t pcm_len = sample_data.size(), cap = out.size() - QOA_HEADER_SIZE;
        uint8_t *dst = out.data() + QOA_HEADER_SIZE;
        size_t written = 0;
        int32_t status = 0;
        eng_->check(qoa_b200_encoder_batch_encode_chunk(state(), &pcm, &pcm_len, &dst, &cap, &written, &status),
                    "qoa_b200_encoder_batch_encode_chunk");
        if (status != QOA_B200_OK) throw EncodeError(static_cast<EncodeError::Kind>(status));
        out.resize(QOA_HEADER_SIZE + written);
        return out;
    }

    void write_header(std::ostream &writer) const // :404-408
    {
        uint8_t h[8];
        detail::put_file_header(h, desc_.samples);
        if (!writer.write(reinterpret_cast<const char *>(h), 8)) throw EncodeError(EncodeError::IoError, "write failed");
    }

    // Returns frame_len, the samples per channel encoded (Ok(frame_len), :438).
    size_t encode_frame(const std::vector<int16_t> &sample_data, std::ostream &writer) // :418-439
    {
        std::vector<uint8_t> out(qoa_b200_frame_size(desc_.channels, QOA_FRAME_LEN));
        static const int16_t none = 0; // a null pointer would mean "this stream sits the call out"; empty input is an error (:424)
        const int16_t *pcm = sample_data.empty() ? &none : sample_data.data();
        const size_t pcm_len = sample_data.size(), cap = out.size();
        uint8_t *dst = out.data();
        size_t written = 0;
        int32_t status = 0;
        eng_->check(qoa_b200_encoder_batch_encode_frame(state(), &pcm, &pcm_len, &dst, &cap, &written, &status),
                    "qoa_b200_encoder_batch_encode_frame");
        if (status != QOA_B200_OK) throw EncodeError(static_cast<EncodeError::Kind>(status));
        if (!writer.write(reinterpret_cast<const char *>(out.data()), std::streamsize(written)))
            throw EncodeError(EncodeError::IoError, "write failed");
        return sample_data.size() / desc_.channels;
    }

private:
    qoa_b200_encoder_batch *state()
    {
        if (!batch_) {
            const qoa_b200_desc d = detail::c_desc(desc_);
            int32_t status = 0;
            eng_->check(qoa_b200_encoder_batch_new(eng_->handle(), &d, 1, &status, &batch_), "qoa_b200_encoder_batch_new");
        }
        return batch_;
    }
    QoaDesc desc_;
    Engine *eng_;
    qoa_b200_encoder_batch *batch_ = nullptr;
};

// ---- decode ----------------------------------------------------------------------------------------------------

using DecodeResult = std::variant<DecodedQoa, DecodeError>;

// Batched decode_all (:713-747).
inline std::vector<DecodeResult> decode_all_batch(const std::vector<std::vector<uint8_t>> &streams, Engine &eng = Engine::shared())
{
    const size_t n = streams.size();
    std::vector<DecodeResult> res;
    if (n == 0) return res;
    std::vector<const uint8_t *> qoa(n);
    std::vector<size_t> len(n);
    for (size_t i = 0; i < n; ++i) qoa[i] = streams[i].data(), len[i] = streams[i].size();
    qoa_b200_decode_plan *plan = nullptr;
    eng.check(qoa_b200_decode_plan_create(eng.handle(), qoa.data(), len.data(), n, QOA_B200_DECODE_ALL, &plan),
              "qoa_b200_decode_plan_create");
    std::unique_ptr<qoa_b200_decode_plan, void (*)(qoa_b200_decode_plan *)> guard(plan, qoa_b200_decode_plan_destroy);
    std::vector<qoa_b200_stream_info> info(n);
    eng.check(qoa_b200_decode_plan_info(plan, info.data()), "qoa_b200_decode_plan_info");
    std::vector<std::vector<int16_t>> out(n);
    std::vector<int16_t *> outp(n);
    std::vector<size_t> cap(n);
    std::vector<int32_t> status(n);
    for (size_t i = 0; i < n; ++i) {
        out[i].resize(info[i].samples);
        outp[i] = out[i].data();
        cap[i] = out[i].size();
    }
    eng.check(qoa_b200_decode_plan_run(plan, outp.data(), cap.data(), status.data()), "qoa_b200_decode_plan_run");
    res.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != QOA_B200_OK)
            res.emplace_back(DecodeError(DecodeError::from_status(status[i])));
        else
            res.emplace_back(DecodedQoa{uint8_t(info[i].channels), info[i].sample_rate, std::move(out[i])});
    }
    return res;
}

// The same on every device of a MultiEngine: output sizes come from qoa_b200_probe (a host-side walk of the frame headers,
// no sample decoded), the samples from the sharded one-call path.
inline std::vector<DecodeResult> decode_all_batch(const std::vector<std::vector<uint8_t>> &streams, MultiEngine &eng)
{
    const size_t n = streams.size();
    std::vector<DecodeResult> res;
    if (n == 0) return res;
    std::vector<const uint8_t *> qoa(n);
    std::vector<size_t> len(n), cap(n);
    std::vector<qoa_b200_stream_info> info(n);
    std::vector<std::vector<int16_t>> out(n);
    std::vector<int16_t *> outp(n);
    std::vector<int32_t> status(n);
    for (size_t i = 0; i < n; ++i) {
        qoa[i] = streams[i].data(), len[i] = streams[i].size();
        eng.check(qoa_b200_probe(qoa[i], len[i], QOA_B200_DECODE_ALL, &info[i]), "qoa_b200_probe");
        out[i].resize(info[i].samples);
        outp[i] = out[i].data();
        cap[i] = out[i].size();
    }
    eng.check(qoa_b200_mctx_decode_all_batch(eng.handle(), qoa.data(), len.data(), n, QOA_B200_DECODE_ALL, outp.data(),
                                             cap.data(), info.data(), status.data()),
              "qoa_b200_mctx_decode_all_batch");
    res.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != QOA_B200_OK)
            res.emplace_back(DecodeError(DecodeError::from_status(status[i])));
        else
            res.emplace_back(DecodedQoa{uint8_t(info[i].channels), info[i].sample_rate, std::move(out[i])});
    }
    return res;
}

inline DecodedQoa decode_all(const std::vector<uint8_t> &bytes, Engine &eng = Engine::shared()) // :713-747
{
    auto res = decode_all_batch({bytes}, eng);
    if (auto *e = std::get_if<DecodeError>(&res[0])) throw *e;
    return std::move(std::get<DecodedQoa>(res[0]));
}

inline DecodedQoa decode_all(std::istream &reader, Engine &eng = Engine::shared()) { return decode_all(detail::read_all(reader), eng); }

inline DecodedQoa open_and_decode_all(const std::string &path, Engine &eng = Engine::shared()) // :750-754
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw DecodeError(DecodeError::IoError, "cannot open " + path);
    return decode_all(f, eng);
}

// :59-68, :117-331, Iterator :661-693.  The whole stream is indexed and decoded on the GPU when the decoder is created;
// next() replays the reference's item order (FrameHeader, then that frame's samples, ...) and throws the reference's
// error at the position where the reference's iterator would return it.
class QoaDecoder {
public:
    explicit QoaDecoder(std::vector<uint8_t> bytes, Engine &eng = Engine::shared()) : eng_(&eng)
    {
        const uint8_t *q = bytes.data();
        const size_t len = bytes.size();
        qoa_b200_decode_plan *plan = nullptr;
        eng.check(qoa_b200_decode_plan_create(eng.handle(), &q, &len, 1, QOA_B200_DECODE_ITER, &plan), "qoa_b200_decode_plan_create");
        std::unique_ptr<qoa_b200_decode_plan, void (*)(qoa_b200_decode_plan *)> guard(plan, qoa_b200_decode_plan_destroy);
        eng.check(qoa_b200_decode_plan_info(plan, &info_), "qoa_b200_decode_plan_info");
        if (info_.status != QOA_B200_OK && info_.frames == 0) // QoaDecoder::new failed (:127-169)
            throw DecodeError(DecodeError::from_status(info_.status));
        std::vector<qoa_b200_frame_header> hdrs(info_.index_frames);
        hdrs.resize(qoa_b200_decode_plan_frame_headers(plan, 0, hdrs.data(), hdrs.size()));
        for (const auto &h : hdrs) headers_.push_back({uint8_t(h.num_channels), h.sample_rate, uint16_t(h.num_samples_per_channel)});
        samples_.resize(info_.samples);
        int16_t *outp = samples_.data();
        const size_t cap = samples_.size();
        int32_t status = 0;
        eng.check(qoa_b200_decode_plan_run(plan, &outp, &cap, &status), "qoa_b200_decode_plan_run");
        if (info_.streaming) {
            mode_ = ProcessingMode::Streaming();
            pending_default_ = true; // :674-677: a default FrameHeader comes first
        } else {
            mode_ = ProcessingMode::FixedSamples(uint8_t(info_.channels), info_.sample_rate, info_.header_samples);
            if (!headers_.empty()) current_ = headers_[0];
        }
        inner_ = std::move(bytes);
    }
    explicit QoaDecoder(std::istream &reader, Engine &eng = Engine::shared()) : QoaDecoder(detail::read_all(reader), eng) {}

    static QoaDecoder open(const std::string &path, Engine &eng = Engine::shared()) // :621-624
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw DecodeError(DecodeError::IoError, "cannot open " + path);
        return QoaDecoder(f, eng);
    }
    static QoaDecoder new_streaming(Engine &eng = Engine::shared()) { return QoaDecoder(StreamingTag{}, eng); } // :632-638

    const ProcessingMode &mode() const { return mode_; }                       // :180
    const FrameHeader &current_frame_header() const { return current_; }       // :186
    std::optional<double> total_duration() const                               // :190-201, seconds
    {
        if (mode_.streaming) return std::nullopt;
        return double(mode_.samples) / double(mode_.sample_rate);
    }
    std::vector<uint8_t> into_inner() { return std::move(inner_); }            // :172

    // Iterator::next (:661-693): nullopt at the clean end; throws DecodeError where the reference yields Err.
    std::optional<QoaItem> next()
    {
        if (pending_default_) {
            pending_default_ = false;
            return QoaItem(FrameHeader{});
        }
        if (left_in_frame_ == 0) {
            if (frame_ < headers_.size()) {
                current_ = headers_[frame_++];
                const size_t want = size_t(current_.num_samples_per_channel) * current_.num_channels;
                left_in_frame_ = want < samples_.size() - pos_ ? want : samples_.size() - pos_;
                return QoaItem(current_);
            }
            if (info_.status != QOA_B200_OK && !failed_) {
                failed_ = true;
                throw DecodeError(DecodeError::from_status(info_.status));
            }
            return std::nullopt;
        }
        --left_in_frame_;
        return QoaItem(samples_[pos_++]);
    }

    // :641-651: decode one (or more) whole frames given as bytes; only on a new_streaming() decoder
    std::vector<int16_t> decode_frame(const std::vector<uint8_t> &frame_data)
    {
        if (!frame_mode_) throw std::logic_error("decode_frame is only available on QoaDecoder::new_streaming()");
        const uint8_t *q = frame_data.data();
        const size_t len = frame_data.size();
        qoa_b200_decode_plan *plan = nullptr;
        eng_->check(qoa_b200_decode_plan_create(eng_->handle(), &q, &len, 1, QOA_B200_DECODE_FRAMES, &plan), "qoa_b200_decode_plan_create");
        std::unique_ptr<qoa_b200_decode_plan, void (*)(qoa_b200_decode_plan *)> guard(plan, qoa_b200_decode_plan_destroy);
        qoa_b200_stream_info info;
        eng_->check(qoa_b200_decode_plan_info(plan, &info), "qoa_b200_decode_plan_info");
        if (info.status != QOA_B200_OK) throw DecodeError(DecodeError::from_status(info.status));
        std::vector<int16_t> out(info.samples);
        int16_t *outp = out.data();
        const size_t cap = out.size();
        int32_t status = 0;
        eng_->check(qoa_b200_decode_plan_run(plan, &outp, &cap, &status), "qoa_b200_decode_plan_run");
        if (status != QOA_B200_OK) throw DecodeError(DecodeError::from_status(status));
        return out;
    }

private:
    struct StreamingTag {};
    QoaDecoder(StreamingTag, Engine &eng) : eng_(&eng), frame_mode_(true) { info_ = qoa_b200_stream_info{}; }
    Engine *eng_;
    bool frame_mode_ = false;
    qoa_b200_stream_info info_;
    ProcessingMode mode_;
    FrameHeader current_;
    std::vector<FrameHeader> headers_;
    std::vector<int16_t> samples_;
    std::vector<uint8_t> inner_;
    size_t frame_ = 0, pos_ = 0, left_in_frame_ = 0;
    bool pending_default_ = false, failed_ = false;
};

} // namespace qoaudio

#endif
