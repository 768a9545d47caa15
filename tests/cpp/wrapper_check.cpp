// Exercises include/qoa_b200.hpp (the C++ mirror of the crate's API) on the GPU and leaves its outputs in a
// directory for tests/test_gpu_cpp.py to compare with the oracle.  Usage: wrapper_check <out_dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "qoa_b200.hpp"

using namespace qoaudio;

static int failures = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
            ++failures;                                                              \
        }                                                                            \
    } while (0)

static std::vector<int16_t> make_pcm(uint32_t seed, size_t n) // triangle wave + LCG noise; the test reads it back from disk
{
    std::vector<int16_t> v(n);
    uint32_t s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        const int tri = int((i * (37 + seed)) % 4096) - 2048;
        v[i] = int16_t(tri * 6 + int(s >> 22) - 512);
    }
    return v;
}

template <typename T>
static void dump(const std::string &path, const std::vector<T> &v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), std::streamsize(v.size() * sizeof(T)));
}

template <typename E, typename F>
static int kind_of(F &&f)
{
    try {
        f();
    } catch (const E &e) {
        return int(e.kind);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc != 2) return 2;
    const std::string dir = argv[1];
    const QoaDesc desc{2, 44100, 5120 * 3 + 777};
    const auto pcm = make_pcm(1, size_t(desc.samples) * desc.channels);
    dump(dir + "/stereo.pcm", pcm);

    // encode_all == QoaEncoder::encode == write_header + encode_frame x N (src/lib.rs:1261-1297)
    const auto one_shot = encode_all(pcm, desc);
    dump(dir + "/stereo.qoa", one_shot);
    CHECK(one_shot.size() == qoa_b200_encoded_size(2, desc.samples));
    {
        QoaEncoder enc(desc);
        CHECK(enc.encode(pcm) == one_shot);
    }
    {
        QoaEncoder enc(desc);
        std::ostringstream os;
        enc.write_header(os);
        size_t total = 0; // encode_frame returns Ok(frame_len) = samples per channel (src/lib.rs:438)
        for (size_t at = 0; at < pcm.size(); at += size_t(QOA_FRAME_LEN) * 2) {
            const size_t end = std::min(pcm.size(), at + size_t(QOA_FRAME_LEN) * 2);
            total += enc.encode_frame(std::vector<int16_t>(pcm.begin() + at, pcm.begin() + end), os);
        }
        const std::string s = os.str();
        CHECK(total == desc.samples);
        CHECK(std::vector<uint8_t>(s.begin(), s.end()) == one_shot);
        CHECK(kind_of<EncodeError>([&] { enc.encode_frame({}, os); }) == EncodeError::InvalidSamples);        // :424
        CHECK(kind_of<EncodeError>([&] { enc.encode_frame(std::vector<int16_t>(3), os); }) == EncodeError::InvalidSamples);
    }
    // error precedence (:1148-1183)
    CHECK(kind_of<EncodeError>([] { QoaEncoder e(QoaDesc{0, 0, 0}); }) == EncodeError::InvalidChannels);
    CHECK(kind_of<EncodeError>([] { QoaEncoder e(QoaDesc{9, 44100, 10}); }) == EncodeError::InvalidChannels);
    CHECK(kind_of<EncodeError>([] { QoaEncoder e(QoaDesc{2, 0, 0}); }) == EncodeError::InvalidSampleRate);
    CHECK(kind_of<EncodeError>([] { QoaEncoder e(QoaDesc{2, 44100, 0}); }) == EncodeError::InvalidSamples);
    CHECK(kind_of<EncodeError>([&] { encode_all(std::vector<int16_t>(10), QoaDesc{2, 44100, 6}); }) == EncodeError::InvalidSamples);

    // decode_all, and the iterator's item order (:998-1032)
    const DecodedQoa dec = decode_all(one_shot);
    dump(dir + "/stereo.dec", dec.samples);
    CHECK(dec.num_channels == 2 && dec.sample_rate == 44100 && dec.samples.size() == pcm.size());
    {
        QoaDecoder d(one_shot);
        CHECK(d.mode() == ProcessingMode::FixedSamples(2, 44100, desc.samples));
        CHECK(d.total_duration() && *d.total_duration() == double(desc.samples) / 44100.0);
        size_t headers = 0, samples = 0;
        bool same = true;
        FrameHeader last;
        while (auto item = d.next()) {
            if (auto *h = std::get_if<FrameHeader>(&*item)) {
                ++headers;
                last = *h;
                CHECK(d.current_frame_header() == *h);
            } else {
                same = same && samples < dec.samples.size() && std::get<int16_t>(*item) == dec.samples[samples];
                ++samples;
            }
        }
        CHECK(headers == 4 && samples == pcm.size() && same);
        CHECK(last.num_samples_per_channel == 777 && last.num_channels == 2 && last.sample_rate == 44100);
        CHECK(!d.next());
    }
    // truncated input: the items before the failing frame come first, then IoError (:206-216, :783-795)
    {
        std::vector<uint8_t> cut(one_shot.begin(), one_shot.begin() + 8 + 4136 + 100);
        QoaDecoder d(cut);
        size_t samples = 0;
        int kind = 0;
        try {
            while (auto item = d.next()) samples += std::holds_alternative<int16_t>(*item);
        } catch (const DecodeError &e) {
            kind = e.kind;
        }
        CHECK(kind == DecodeError::IoError && samples >= 5120 * 2);
        std::ofstream(dir + "/results.txt") << "trunc_len=" << cut.size() << "\ntrunc_samples=" << samples << "\n";
        CHECK(kind_of<DecodeError>([&] { decode_all(cut); }) == DecodeError::IoError);
    }
    CHECK(kind_of<DecodeError>([] { QoaDecoder d(std::vector<uint8_t>{'n', 'o', 'p', 'e', 0, 0, 0, 1, 0, 0}); }) == DecodeError::NotQoaFile);
    CHECK(kind_of<DecodeError>([&] { QoaDecoder d(std::vector<uint8_t>(one_shot.begin(), one_shot.begin() + 8)); }) == DecodeError::NoSamples);

    // new_streaming + decode_frame (:1034-1066)
    {
        QoaDecoder d = QoaDecoder::new_streaming();
        CHECK(d.mode().streaming && !d.total_duration());
        const std::vector<uint8_t> frame(one_shot.begin() + 8, one_shot.begin() + 8 + 4136);
        const auto s = d.decode_frame(frame);
        CHECK(s.size() == 5120 * 2 && std::equal(s.begin(), s.end(), dec.samples.begin()));
        CHECK(kind_of<DecodeError>([&] { d.decode_frame(std::vector<uint8_t>(frame.begin(), frame.begin() + 100)); }) == DecodeError::IoError);
    }

    // batches: a failing stream never aborts the batch
    {
        const QoaDesc mono{1, 8000, 4321};
        const auto mpcm = make_pcm(2, mono.samples);
        dump(dir + "/mono.pcm", mpcm);
        const std::vector<int16_t> none;
        auto enc = encode_all_batch({{&pcm, desc}, {&none, QoaDesc{0, 8000, 1}}, {&mpcm, mono}});
        CHECK(std::holds_alternative<std::vector<uint8_t>>(enc[0]) && std::get<std::vector<uint8_t>>(enc[0]) == one_shot);
        CHECK(std::holds_alternative<EncodeError>(enc[1]) && std::get<EncodeError>(enc[1]).kind == EncodeError::InvalidChannels);
        CHECK(std::holds_alternative<std::vector<uint8_t>>(enc[2]));
        const auto &mq = std::get<std::vector<uint8_t>>(enc[2]);
        dump(dir + "/mono.qoa", mq);
        auto decs = decode_all_batch({mq, std::vector<uint8_t>{1, 2, 3}, one_shot});
        CHECK(std::holds_alternative<DecodedQoa>(decs[0]) && std::get<DecodedQoa>(decs[0]).samples.size() == mono.samples);
        CHECK(std::holds_alternative<DecodeError>(decs[1]));
        CHECK(std::holds_alternative<DecodedQoa>(decs[2]) && std::get<DecodedQoa>(decs[2]).samples == dec.samples);
        dump(dir + "/mono.dec", std::get<DecodedQoa>(decs[0]).samples);
        // the same batch through a multi-device engine (two contexts; with one visible GPU both sit on device 0)
        int ndev = 0;
        {
            MultiEngine all; // no list: every visible device
            ndev = all.device_count();
            CHECK(ndev >= 1);
        }
        MultiEngine multi(ndev >= 2 ? std::vector<int>{0, 1} : std::vector<int>{0, 0});
        CHECK(multi.device_count() == 2);
        auto menc = encode_all_batch({{&pcm, desc}, {&none, QoaDesc{0, 8000, 1}}, {&mpcm, mono}}, multi);
        CHECK(std::holds_alternative<std::vector<uint8_t>>(menc[0]) && std::get<std::vector<uint8_t>>(menc[0]) == one_shot);
        CHECK(std::holds_alternative<EncodeError>(menc[1]) && std::get<EncodeError>(menc[1]).kind == EncodeError::InvalidChannels);
        CHECK(std::holds_alternative<std::vector<uint8_t>>(menc[2]) && std::get<std::vector<uint8_t>>(menc[2]) == mq);
        auto mdecs = decode_all_batch({mq, std::vector<uint8_t>{1, 2, 3}, one_shot}, multi);
        CHECK(std::holds_alternative<DecodedQoa>(mdecs[0]) &&
              std::get<DecodedQoa>(mdecs[0]).samples == std::get<DecodedQoa>(decs[0]).samples);
        CHECK(std::holds_alternative<DecodeError>(mdecs[1]));
        CHECK(std::holds_alternative<DecodedQoa>(mdecs[2]) && std::get<DecodedQoa>(mdecs[2]).samples == dec.samples);
    }
    {
        dump(dir + "/file.qoa", one_shot);
        CHECK(open_and_decode_all(dir + "/file.qoa").samples == dec.samples);
        CHECK(kind_of<DecodeError>([&] { open_and_decode_all(dir + "/missing.qoa"); }) == DecodeError::IoError);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
