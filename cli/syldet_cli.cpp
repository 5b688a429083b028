// syldet — command line detector, Linux counterpart of SyllableDetectorCLI/main.swift.
//   syldet -n <network.txt> -a <audio.wav> [-a ...] [-d <seconds>]
// Same flags (main.swift:21-23), same CSV rows on stdout: channel,sample,seconds,out0[,out1...]
// (TrackDetector.swift:92-96, help text main.swift:31-39). Differences, both forced by leaving macOS:
//   * audio is read from RIFF/WAVE (PCM16, PCM24, PCM32 or float32) instead of AVFoundation. 16- and 24-bit PCM at the network's
//     rate goes to the device as it is in the file (interleaved integers, converted by the ingest kernel); a file at another rate is
//     converted to the network's rate on the device first, with the polyphase converter (upstream asks AVFoundation for
//     config.samplingRate, SyllableDetector.swift:19-23, and gets its converter);
//   * every channel of a file is a "track": upstream reads channel 0 of each AVAssetTrack (main.swift:86-89), and rows are grouped
//     by channel, then time (upstream interleaves tracks buffer by buffer, main.swift:126-130).
// All compute goes through the C-ABI of libsyldet_cuda.so; there is no CPU path.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/syldet.h"

namespace {

struct Wav {
    int channels = 0, rate = 0;
    int64_t frames = 0;
    int pcm_format = SYLDET_PCM_F32;          // what `raw` holds when it is usable as is (int16 / packed 24-bit), else F32
    std::vector<unsigned char> raw;           // interleaved samples exactly as in the file (PCM16 / PCM24 only)
    std::vector<float> interleaved;           // float32 samples (always for PCM32 / float files; on demand otherwise)
    const unsigned char *data = nullptr;      // file bytes of the data chunk (valid while `file` lives)
    int bits = 0, fmt = 0;
    std::vector<unsigned char> file;
    void to_float() {                          // AVAssetReader's LPCM Float32 conversion (SyllableDetector.swift:19-23)
        if (!interleaved.empty() || frames == 0) return;
        const size_t total = (size_t)frames * channels;
        interleaved.resize(total);
        if (bits == 16) for (size_t i = 0; i < total; ++i) interleaved[i] = (float)(int16_t)(data[2 * i] | (data[2 * i + 1] << 8)) / 32768.0f;
        else for (size_t i = 0; i < total; ++i) {
            int32_t v = data[3 * i] | (data[3 * i + 1] << 8) | ((int32_t)(int8_t)data[3 * i + 2] << 16);
            interleaved[i] = (float)v / 8388608.0f;
        }
    }
};

uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

bool read_wav(const std::string &path, Wav &w, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open file"; return false; }
    std::vector<unsigned char> &buf = w.file;
    unsigned char tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    if (buf.size() < 12 || std::memcmp(buf.data(), "RIFF", 4) || std::memcmp(buf.data() + 8, "WAVE", 4)) { err = "not a RIFF/WAVE file"; return false; }
    int fmt = 0, bits = 0, align = 0;
    size_t pos = 12;
    const unsigned char *data = nullptr;
    size_t data_len = 0;
    while (pos + 8 <= buf.size()) {
        const uint32_t len = rd32(&buf[pos + 4]);
        const unsigned char *body = &buf[pos + 8];
        const size_t avail = buf.size() - pos - 8;
        if (!std::memcmp(&buf[pos], "fmt ", 4) && len >= 16 && avail >= 16) {
            fmt = rd16(body);
            w.channels = rd16(body + 2);
            w.rate = (int)rd32(body + 4);
            align = rd16(body + 12);
            bits = rd16(body + 14);
            if (fmt == 0xFFFE && len >= 26 && avail >= 26) fmt = rd16(body + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
        } else if (!std::memcmp(&buf[pos], "data", 4)) {
            data = body;
            data_len = len < avail ? len : avail;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    if (!data || w.channels <= 0 || align <= 0) { err = "missing fmt or data chunk"; return false; }
    w.frames = (int64_t)(data_len / align);
    w.data = data;
    w.bits = bits;
    w.fmt = fmt;
    const size_t total = (size_t)w.frames * w.channels;
    if (fmt == 1 && (bits == 16 || bits == 24)) {
        w.pcm_format = bits == 16 ? SYLDET_PCM_S16 : SYLDET_PCM_S24;   // uploaded as they are, converted on the device
    } else if (fmt == 1 && bits == 32) {
        w.interleaved.resize(total);
        for (size_t i = 0; i < total; ++i) w.interleaved[i] = (float)((double)(int32_t)rd32(data + 4 * i) / 2147483648.0);
    } else if (fmt == 3 && bits == 32) {
        w.interleaved.resize(total);
        std::memcpy(w.interleaved.data(), data, total * 4);
    } else { err = "unsupported sample format (want PCM 16/24/32 or float32)"; return false; }
    return true;
}

void usage() {
    std::puts("Usage: syldet -n <net> [-a <audio>]... [-d <seconds>] [-s <trace.wav>]");
    std::puts("  -n, --net <net>          Path to trained network file.");
    std::puts("  -a, --audio <audio>      Path to the audio file to process.");
    std::puts("  -d, --debounce <seconds> Number of seconds to debounce triggers.");
    std::puts("  -s, --simulate <wav>     Write the simulator trace (output 0 / threshold 0, clamped to [0, 1], 16-bit) instead of events.");
    std::puts("The command line will write a comma-separated list of detection events (when the network has at least one output above threshold) to standard out. For example, it might output:");
    std::puts("");
    std::puts("\t0,1593298,36.1292063492063,0.918557");
    std::puts("");
    std::puts("The columns are:");
    std::puts("1. The track or channel number from the audio file (starting with 0).");
    std::puts("2. The sample number from the audio when detection occurred.");
    std::puts("3. The timestamp from the audio when detection occurred.");
    std::puts("4. The first neural network output. Note that there may be additional columns for additional outputs.");
}

// shortest decimal that round-trips, like Swift's description of Double / Float (TrackDetector.swift:92-96)
std::string shortest(double v, bool single) {
    char b[64];
    for (int prec = 1; prec <= 17; ++prec) {
        std::snprintf(b, sizeof b, "%.*g", prec, v);
        if (single ? (std::strtof(b, nullptr) == (float)v) : (std::strtod(b, nullptr) == v)) break;
    }
    std::string s(b);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}

bool write_wav16(const std::string &path, const std::vector<int16_t> &interleaved, int channels, int rate) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t data = (uint32_t)(interleaved.size() * 2), riff = 36 + data, fmt_len = 16, byte_rate = (uint32_t)rate * channels * 2;
    const uint16_t pcm = 1, ch = (uint16_t)channels, align = (uint16_t)(channels * 2), bits = 16;
    const uint32_t r = (uint32_t)rate;
    bool ok = std::fwrite("RIFF", 1, 4, f) == 4 && std::fwrite(&riff, 4, 1, f) == 1 && std::fwrite("WAVEfmt ", 1, 8, f) == 8 &&
              std::fwrite(&fmt_len, 4, 1, f) == 1 && std::fwrite(&pcm, 2, 1, f) == 1 && std::fwrite(&ch, 2, 1, f) == 1 &&
              std::fwrite(&r, 4, 1, f) == 1 && std::fwrite(&byte_rate, 4, 1, f) == 1 && std::fwrite(&align, 2, 1, f) == 1 &&
              std::fwrite(&bits, 2, 1, f) == 1 && std::fwrite("data", 1, 4, f) == 4 && std::fwrite(&data, 4, 1, f) == 1 &&
              std::fwrite(interleaved.data(), 2, interleaved.size(), f) == interleaved.size();
    return std::fclose(f) == 0 && ok;
}

}  // namespace

int main(int argc, char **argv) {
    std::string net;
    std::vector<std::string> audio;
    bool have_debounce = false;
    double debounce = 0.0;
    std::string simulate;  // -s <out.wav>: write the simulator's output track instead of CSV rows
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](std::string &dst) { if (i + 1 >= argc) return false; dst = argv[++i]; return true; };
        std::string v;
        if (a == "-n" || a == "--net") { if (!value(net)) { usage(); return 64; } }
        else if (a == "-a" || a == "--audio") { if (!value(v)) { usage(); return 64; } audio.push_back(v); }
        else if (a == "-d" || a == "--debounce") { if (!value(v)) { usage(); return 64; } char *e; debounce = std::strtod(v.c_str(), &e); have_debounce = (*e == 0 && !v.empty()); }
        else if (a == "-s" || a == "--simulate") { if (!value(simulate)) { usage(); return 64; } }
        else { usage(); return 64; }  // EX_USAGE, main.swift:40
    }
    if (net.empty()) { usage(); return 64; }

    syldet_config *cfg = nullptr;
    if (syldet_config_load_text(net.c_str(), &cfg) != SYLDET_OK || syldet_config_validate(cfg) != SYLDET_OK) {
        std::fprintf(stderr, "Unable to load the network configuration: %s\n", syldet_last_error());
        return 1;
    }
    syldet_batch *batch = nullptr;
    if (syldet_batch_create(cfg, 0, &batch) != SYLDET_OK) {
        std::fprintf(stderr, "Unable to create the detector: %s\n", syldet_last_error());
        return 1;
    }
    const double fs = syldet_config_sampling_rate(cfg);
    const int64_t debounce_frames = have_debounce ? syldet_config_debounce_frames(cfg, debounce) : 0;

    for (const std::string &path : audio) {
        Wav w;
        std::string err;
        if (!read_wav(path, w, err)) { std::fprintf(stderr, "Unable to read %s: %s\n", path.c_str(), err.c_str()); continue; }
        const void *pcm = nullptr;
        int pcm_format = SYLDET_PCM_F32, layout = SYLDET_LAYOUT_INTERLEAVED;
        std::vector<float> converted;   // planar, at the network's rate
        if (std::fabs((double)w.rate - fs) > 1.0 && w.frames > 0) {
            // another sampling rate: convert on the device (upstream: the asset reader delivers config.samplingRate, SyllableDetector.swift:19-23)
            w.to_float();
            std::vector<float> planar((size_t)w.channels * w.frames);
            for (int c = 0; c < w.channels; ++c)
                for (int64_t i = 0; i < w.frames; ++i) planar[(size_t)c * w.frames + i] = w.interleaved[(size_t)i * w.channels + c];
            const int64_t n_out = syldet_resample_output_length(SYLDET_RESAMPLE_POLYPHASE, w.frames, (double)w.rate, fs);
            if (n_out <= 0) {
                std::fprintf(stderr, "Can not read audio tracks found in %s: no conversion from %d Hz to %g Hz.\n", path.c_str(), w.rate, fs);
                continue;
            }
            converted.resize((size_t)w.channels * n_out);
            int64_t got = 0;
            if (syldet_resample_host(SYLDET_RESAMPLE_POLYPHASE, planar.data(), w.channels, w.frames, w.frames, (double)w.rate, fs, converted.data(), n_out,
                                     &got, 0) != SYLDET_OK) {
                std::fprintf(stderr, "Can not convert %s to the network's sampling rate: %s.\n", path.c_str(), syldet_last_error());
                continue;
            }
            w.frames = got;
            w.rate = (int)std::lround(fs);
            pcm = converted.data();
            layout = SYLDET_LAYOUT_PLANAR;
        } else if (w.pcm_format != SYLDET_PCM_F32) {
            pcm = w.data;                    // 16- / 24-bit samples exactly as in the file
            pcm_format = w.pcm_format;
        } else {
            pcm = w.interleaved.data();
        }
        if (audio.size() > 1) std::printf("%s\n", path.c_str());  // main.swift:122-124
        if (w.frames <= 0) continue;
        if (!simulate.empty()) {
            // the GUI simulator's output (ViewControllerSimulator.swift:135-376): 16-bit PCM, one trace channel per input channel
            std::vector<int16_t> planar((size_t)w.channels * w.frames), inter((size_t)w.channels * w.frames);
            if (syldet_batch_simulate_host(batch, pcm, pcm_format, w.channels, w.frames, w.frames, layout, SYLDET_PCM_S16, planar.data()) != SYLDET_OK) {
                std::fprintf(stderr, "Can not simulate %s: %s.\n", path.c_str(), syldet_last_error());
                continue;
            }
            for (int c = 0; c < w.channels; ++c)
                for (int64_t i = 0; i < w.frames; ++i) inter[(size_t)i * w.channels + c] = planar[(size_t)c * w.frames + i];
            const std::string out = audio.size() > 1 ? simulate + "." + std::to_string(&path - &audio[0]) + ".wav" : simulate;
            if (!write_wav16(out, inter, w.channels, w.rate)) std::fprintf(stderr, "Unable to write %s\n", out.c_str());
            continue;
        }
        syldet_events *ev = nullptr;
        if (syldet_batch_run_host(batch, pcm, pcm_format, w.channels, w.frames, w.frames, layout, debounce_frames, SYLDET_DETECT_ANY_OUTPUT, nullptr,
                                  &ev) != SYLDET_OK) {
            std::fprintf(stderr, "Can not start reading %s: %s.\n", path.c_str(), syldet_last_error());
            continue;
        }
        const int64_t n = syldet_events_count(ev);
        const int O = syldet_events_outputs_per_event(ev);
        const syldet_event *rows = syldet_events_data(ev);
        const float *outs = syldet_events_outputs(ev);
        // upstream interleaves tracks buffer by buffer (main.swift:126-130); rows here are grouped by channel, then time
        for (int64_t r = 0; r < n; ++r) {
            std::printf("%d,%lld,%s", rows[r].channel, (long long)rows[r].sample, shortest((double)rows[r].sample / fs, false).c_str());
            for (int o = 0; o < O; ++o) std::printf(",%s", shortest((double)outs[r * O + o], true).c_str());
            std::printf("\n");
        }
        syldet_events_free(ev);
    }
    syldet_batch_destroy(batch);
    syldet_config_free(cfg);
    return 0;
}
