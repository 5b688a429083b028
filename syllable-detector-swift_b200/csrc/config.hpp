// Host-side model of SyllableDetectorConfig + NeuralNet (reference: Common/SyllableDetectorConfig.swift:11-45,
// Common/NeuralNet.swift:233-326) and the derived STFT geometry (Common/CircularShortTimeFourierTransform.swift:61-129,
// Common/SyllableDetector.swift:37-60).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/syldet.h"

namespace syldet {

struct Processing {
    int function = SYLDET_PROC_MAPMINMAX;
    std::vector<float> x_offsets, gains;
    float y = 0.0f;  // yMin (mapminmax) / yMean (mapstd)
};

struct Layer {
    int inputs = 0, outputs = 0, transfer = SYLDET_TF_PURELIN;
    std::vector<float> weights;  // row-major [outputs][inputs] (NeuralNet.swift:368, convert_to_text.m:202)
    std::vector<float> biases;
};

struct Config {
    // parsed fields
    double sampling_rate = 0.0;
    int fourier_length = 0, window_length = 0, window_overlap = 0;
    double freq_lo = 0.0, freq_hi = 0.0;
    int time_range = 0;
    int scaling = SYLDET_SCALING_LINEAR;
    std::vector<double> thresholds;
    std::vector<Layer> layers;
    std::vector<Processing> input_processing, output_processing;

    // derived by validate()
    bool valid = false;
    int gap = 0, overlap = 0, hop = 0;   // CSTFT.swift:66-73; hop = gap + W - overlap
    int k0 = 0, k1 = 0, band = 0;        // frequencyIndexRange, band = k1 - k0
    int inputs = 0, outputs = 0;         // NeuralNet.inputs / outputs

    int64_t first_output_sample() const { return (int64_t)gap + window_length + (int64_t)hop * (time_range - 1); }
    int64_t num_columns(int64_t n) const {
        int64_t need = (int64_t)gap + window_length;
        return n < need ? 0 : (n - need) / hop + 1;
    }
    int64_t num_evals(int64_t n) const {
        int64_t e = num_columns(n) - time_range + 1;
        return e > 0 ? e : 0;
    }
    // samples the evaluations [0, E) touch: (E + T - 2) * hop + gap + W
    int64_t samples_for_evals(int64_t evals) const {
        return evals <= 0 ? 0 : (evals + time_range - 2) * (int64_t)hop + gap + window_length;
    }
};

// Error plumbing shared by the whole library (thread-local, see syldet_last_error()).
syldet_status set_error(syldet_status st, const std::string &msg, const std::string &key = std::string());
const std::string &last_error_message();
const std::string &last_error_key();

syldet_status parse_config_text(const char *text, size_t len, Config &out);
syldet_status load_config_file(const char *path, Config &out);
syldet_status validate_config(Config &cfg);
bool frequency_index_range(int fft_len, double f_lo, double f_hi, double rate, int &start, int &end);

}  // namespace syldet

struct syldet_config {
    syldet::Config c;
};
