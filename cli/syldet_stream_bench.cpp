// syldet_stream_bench — per-buffer latency of the live path (BASELINE config 5: many concurrent channels, small buffers).
//   syldet_stream_bench -n <network.txt> [-c channels=64] [-b buffer=32] [-s stream_seconds=60] [-p paced_seconds=5] [-d device=0]
// Drives syldet_stream_submit exactly as Processor.swift drives its detectors (one buffer per channel per audio callback,
// SyllableDetector/Processor.swift:102-149; 32-frame buffers, AudioInterface.swift:474) and times every call on the host:
// latency = submit() entry -> return, i.e. until the tick's outputs and `seen` flags are host-visible.
// Two passes: "burst" submits buffers back to back (sustained real-time factor); "paced" releases each buffer at its
// real-time arrival instant (the GPU idles between ticks, as it would live). Prints one JSON object on stdout.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/syldet.h"

namespace {

using Clock = std::chrono::steady_clock;

double us_between(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }

struct Dist {
    std::vector<double> v;
    void add(double x) { v.push_back(x); }
    double q(double p) {
        if (v.empty()) return 0.0;
        std::sort(v.begin(), v.end());
        size_t i = (size_t)std::ceil(p * (double)v.size());
        if (i > 0) --i;
        return v[std::min(i, v.size() - 1)];
    }
    double mean() const {
        double s = 0;
        for (double x : v) s += x;
        return v.empty() ? 0.0 : s / (double)v.size();
    }
    void print(const char *name) {
        std::printf("\"%s\": {\"n\": %zu, \"mean_us\": %.3f, \"p50_us\": %.3f, \"p90_us\": %.3f, \"p99_us\": %.3f, \"p999_us\": %.3f, \"max_us\": %.3f}",
                    name, v.size(), mean(), q(0.50), q(0.90), q(0.99), q(0.999), q(1.0));
    }
};

// Gaussian-ish noise, sigma 1e-3 (sum of 4 uniforms); latency does not depend on the content.
struct Rng {
    uint64_t s;
    float next() {
        float acc = 0.f;
        for (int i = 0; i < 4; ++i) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            acc += (float)((s >> 40) & 0xFFFFFF) * (1.0f / 16777216.0f) - 0.5f;
        }
        return acc * 1.732e-3f;
    }
};

struct Pass {
    Dist all, with_evals;
    double wall_s = 0.0, stream_s = 0.0;
    int64_t evals = 0, seen = 0, late = 0;
};

bool run_pass(syldet_stream *st, int nch, int n_out, int nbuf, int64_t ticks, double rate, bool paced, const std::vector<float> &audio,
              int64_t audio_ticks, Pass &out) {
    std::vector<const float *> ptrs(nch);
    std::vector<uint8_t> seen(nch);
    std::vector<int32_t> n_new(nch);
    std::vector<float> last((size_t)nch * n_out);
    const double tick_s = nbuf / rate;
    const Clock::time_point t0 = Clock::now();
    for (int64_t t = 0; t < ticks; ++t) {
        const int64_t a = t % audio_ticks;
        for (int ch = 0; ch < nch; ++ch) ptrs[ch] = audio.data() + ((size_t)ch * audio_ticks + a) * nbuf;
        if (paced) {  // the buffer exists once its last sample has been captured
            const auto due = t0 + std::chrono::duration_cast<Clock::duration>(std::chrono::duration<double>((t + 1) * tick_s));
            if (Clock::now() > due) ++out.late;
            while (Clock::now() < due) {
            }
        }
        const Clock::time_point a0 = Clock::now();
        if (syldet_stream_submit(st, ptrs.data(), nbuf, seen.data(), n_new.data(), last.data()) != SYLDET_OK) {
            std::fprintf(stderr, "submit failed: %s\n", syldet_last_error());
            return false;
        }
        const double us = us_between(a0, Clock::now());
        out.all.add(us);
        if (n_new[0] > 0) out.with_evals.add(us);
        out.evals += n_new[0];
        for (int ch = 0; ch < nch; ++ch) out.seen += seen[ch];
    }
    out.wall_s = us_between(t0, Clock::now()) * 1e-6;
    out.stream_s = (double)ticks * tick_s;
    return true;
}

void print_pass(const char *name, Pass &p, int nch) {
    std::printf("\"%s\": {\"ticks\": %zu, \"stream_seconds\": %.3f, \"wall_seconds\": %.4f, \"realtime_factor\": %.2f, "
                "\"channel_audio_seconds_per_second\": %.1f, \"evaluations_per_channel\": %lld, \"seen_flags\": %lld, \"late_ticks\": %lld, ",
                name, p.all.v.size(), p.stream_s, p.wall_s, p.stream_s / p.wall_s, nch * p.stream_s / p.wall_s,
                (long long)p.evals, (long long)p.seen, (long long)p.late);
    p.all.print("per_buffer");
    std::printf(", ");
    p.with_evals.print("per_buffer_with_new_outputs");
    std::printf("}");
}

}  // namespace

int main(int argc, char **argv) {
    std::string net;
    int nch = 64, nbuf = 32, device = 0;
    double stream_s = 60.0, paced_s = 5.0;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        if (k == "-n") net = argv[i + 1];
        else if (k == "-c") nch = std::atoi(argv[i + 1]);
        else if (k == "-b") nbuf = std::atoi(argv[i + 1]);
        else if (k == "-s") stream_s = std::atof(argv[i + 1]);
        else if (k == "-p") paced_s = std::atof(argv[i + 1]);
        else if (k == "-d") device = std::atoi(argv[i + 1]);
    }
    if (net.empty() || nch <= 0 || nbuf <= 0) {
        std::fprintf(stderr, "usage: syldet_stream_bench -n network.txt [-c channels] [-b buffer] [-s seconds] [-p paced_seconds] [-d device]\n");
        return 2;
    }
    syldet_config *cfg = nullptr;
    if (syldet_config_load_text(net.c_str(), &cfg) != SYLDET_OK) {
        std::fprintf(stderr, "config: %s\n", syldet_last_error());
        return 1;
    }
    const double rate = syldet_config_sampling_rate(cfg);
    const int n_out = std::max(1, syldet_config_net_outputs(cfg));
    // two seconds of distinct audio per channel, replayed (the stream itself never repeats state: counters keep running)
    const int64_t audio_ticks = std::max<int64_t>(1, (int64_t)(2.0 * rate / nbuf));
    std::vector<float> audio((size_t)nch * audio_ticks * nbuf);
    Rng rng{0x5eed5eedULL};
    for (float &v : audio) v = rng.next();

    std::printf("{\"channels\": %d, \"buffer_frames\": %d, \"sampling_rate\": %.1f, \"buffer_ms\": %.4f, ", nch, nbuf, rate, 1e3 * nbuf / rate);
    const char *names[2] = {"burst", "paced"};
    const double secs[2] = {stream_s, paced_s};
    bool first = true;
    int64_t launches = 0, fast_launches = 0, resident_ticks = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (secs[pass] <= 0) continue;
        syldet_stream *st = nullptr;
        if (syldet_stream_create(cfg, nch, nbuf, device, &st) != SYLDET_OK) {
            std::fprintf(stderr, "stream: %s\n", syldet_last_error());
            return 1;
        }
        Pass warm, p;
        const int64_t ticks = (int64_t)(secs[pass] * rate / nbuf);
        if (!run_pass(st, nch, n_out, nbuf, std::min<int64_t>(2000, ticks), rate, false, audio, audio_ticks, warm)) return 1;  // warm-up, untimed
        if (!run_pass(st, nch, n_out, nbuf, ticks, rate, pass == 1, audio, audio_ticks, p)) return 1;
        launches += syldet_stream_launch_count(st);
        fast_launches += syldet_stream_fast_tick_count(st);
        resident_ticks += syldet_stream_resident_tick_count(st);
        if (!first) std::printf(", ");
        first = false;
        print_pass(names[pass], p, nch);
        syldet_stream_destroy(st);
    }
    std::printf(", \"kernel_launches\": %lld, \"latency_shaped_tick_launches\": %lld, \"resident_ticks\": %lld}\n", (long long)launches,
                (long long)fast_launches, (long long)resident_ticks);
    syldet_config_free(cfg);
    return 0;
}
